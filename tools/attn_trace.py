#!/usr/bin/env python
"""Phase timeline of the attention kernel's softmax warps (development aid). Builds a TRACE copy of the library
(-DCRA5_ATTN_TRACE, cra5_b200/lib/libcra5b200_trace.so, never shipped or loaded by the package), runs the global
attention shape once and prints, for CTA 0 / tile A and B / lane quarter 0, the clock64 deltas between phase boundaries
of KV steps 8..23:  wait S | TMEM load (+ maximum on the exact path, + wait for the ping-pong turn) | exponentials | wait PV |
store P.

    python tools/attn_trace.py --build      # here (nvcc)
    python tools/attn_trace.py              # on the GPU box
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SRC = os.path.join(ROOT, "cra5_b200", "csrc")
OUT = os.path.join(ROOT, "cra5_b200", "lib", "libcra5b200_trace.so")


def build():
    from cra5_b200 import build as B
    objs = []
    od = os.path.join(ROOT, "cra5_b200", "lib", "obj_trace")
    os.makedirs(od, exist_ok=True)
    for f in B._sources():
        o = os.path.join(od, f + ".o")
        objs.append(o)
        flags = B.NVCC_FLAGS + (["-DCRA5_ATTN_TRACE"] if f == "attn_tc4.cu" else [])
        if f != "attn_tc4.cu" and os.path.exists(o):
            continue
        subprocess.run(["nvcc"] + flags + ["-x", "cu", "-c", os.path.join(SRC, f), "-o", o], check=True)
    subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs, check=True)
    print("built", OUT)


def run():
    import torch
    lib = ctypes.CDLL(OUT)
    heads, nseg, seg = 16, 1, 10368
    if len(sys.argv) > 1 and sys.argv[1] == "window":
        heads, nseg, seg = 16, 144, 576
    rows = nseg * seg
    g = torch.Generator(device="cuda").manual_seed(1)
    q = (torch.randn(heads, rows, 64, device="cuda", generator=g) * 0.25).to(torch.bfloat16)
    k = torch.randn(heads, rows, 64, device="cuda", generator=g).to(torch.bfloat16)
    vt = torch.randn(heads, 64, rows, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.empty(rows, heads * 64, device="cuda", dtype=torch.bfloat16)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    for _ in range(2):
        rc = lib.cra5_op_attention(P(q), P(k), P(vt), P(out), heads * 64, heads, rows, seg,
                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (2 * 64 * 8))()
    assert lib.cra5_debug_attn_trace(buf) == 0
    t = [[[buf[(x * 64 + s) * 8 + ph] for ph in range(8)] for s in range(64)] for x in range(2)]
    t00 = t[0][8][0]
    names = ["waitS", "ld/max/turn", "exps", "waitPV", "storeP"]
    print("tile step  start  " + "  ".join(f"{n:>9s}" for n in names) + "   step_total")
    for s in range(8, 24):
        for x in range(2):
            r = t[x][s]
            # stamps: 0 before wait S, 1 after, 2 exponentials start (load issued / maximum known / turn taken), 4 after the
            # exponentials (turn passed), 5 after wait PV, 6 after the P store and the p_full arrive
            d = [r[1] - r[0], r[2] - r[1], r[4] - r[2], r[5] - r[4], r[6] - r[5]]
            nxt = t[x][s + 1][0]
            print(f"{'AB'[x]}    {s:3d}  {r[0] - t00:6d}  " + "  ".join(f"{v:9d}" for v in d) + f"   {nxt - r[0]:8d}")


if __name__ == "__main__":
    build() if "--build" in sys.argv else run()
