#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 attention kernel variants (CUDA events, L2 flushed): POLY (exp2 on the FMA pipe for
POLY of every 8 column pairs), on the trunk's shapes at 1 and 8 frames per call.

    python tools/perf_attn.py            # parent: one child process per variant (the choice is read once per process)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    from cra5_b200 import _lib as L
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    for name, heads, nseg, seg in (("global_b1", 16, 1, 10368), ("global_b8", 16, 8, 10368), ("window_b8", 16, 144, 576)):
        rows = nseg * seg
        g = torch.Generator(device="cuda").manual_seed(1)
        q = (torch.randn(heads, rows, 64, device="cuda", generator=g) * 0.125 * 2.0).to(torch.bfloat16)   # scores ~ N(0, 2^2)
        k = torch.randn(heads, rows, 64, device="cuda", generator=g).to(torch.bfloat16)
        vt = torch.randn(heads, 64, rows, device="cuda", generator=g).to(torch.bfloat16)
        out = torch.empty(rows, heads * 64, device="cuda", dtype=torch.bfloat16)

        def f():
            L.check(L.lib.cra5_op_attention(L.ptr(q), L.ptr(k), L.ptr(vt), L.ptr(out), heads * 64, heads, rows, seg,
                                            L.stream_ptr()))
        for _ in range(3):
            f()
        ts = []
        for _ in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        ms = ts[len(ts) // 2]
        res[name] = {"ms": round(ms, 4), "tflops": round(4.0 * heads * nseg * seg * seg * 64 / ms / 1e9, 1),
                     "checksum": float(out.float().abs().sum().item())}
    print(json.dumps(res))


def main():
    for poly in ("2", "3", "4"):
        e = dict(os.environ, CRA5_ATTN_POLY=poly)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=e, capture_output=True, text=True,
                           timeout=300)
        print(f"POLY={poly}", r.stdout.strip().splitlines()[-1] if r.returncode == 0 else r.stderr[-800:], flush=True)


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
