#!/usr/bin/env python
"""Where the wall time of one bench step goes, stage by stage: the four public calls of a compress + decompress round
trip (bench.py's `step`) timed with a device synchronisation after each (host wall clock), next to the device time
of the kernels each stage launched (the library's CUDA-event profiler) -- the difference is host work and idle GPU.
Also prints the un-synchronised step time (what bench.py's `value` is made of).

    python tools/step_timeline.py [--batch 8] [--steps 5]
"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch
    from cra5_b200 import _lib, config as C
    from cra5_b200.synthetic import bench_regime
    from cra5_b200.vaeformer import VAEformer
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg, B = C.cra5_268(), args.batch
    net = VAEformer(268, cfg=cfg, device=dev, init_seed=1234, max_batch=B)
    sd = bench_regime({k: v for k, v in net.state_dict().items() if k in C.param_shapes(cfg)}, cfg)
    net.load_state_dict(sd)
    net.update(force=True)
    g = torch.Generator(device=dev).manual_seed(1000)
    frames = [torch.randn(B, cfg.in_chans, 721, 1440, device=dev, generator=g) for _ in range(2)]
    sync = lambda: torch.cuda.synchronize(dev)

    def staged(i, rec):
        t = [time.perf_counter()]
        y, _, _ = net.encode_latent(frames[i % 2], type="float"); sync(); t.append(time.perf_counter())
        out = net.compress_from_latent(y); sync(); t.append(time.perf_counter())
        y_hat = net.decompress(out["strings"], out["z_shape"], return_format="latent"); sync(); t.append(time.perf_counter())
        x_hat = net.decode_latent(y_hat); sync(); t.append(time.perf_counter())
        if rec is not None:
            rec.append([1e3 * (b - a) for a, b in zip(t, t[1:])])

    def plain(i):
        out = net.compress(frames[i % 2])
        net.decompress(out["strings"], out["z_shape"])

    for i in range(3):
        plain(i)
    sync()
    t0 = time.perf_counter()
    for i in range(args.steps):
        plain(i)
    sync()
    plain_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    rec = []
    for i in range(args.steps):
        staged(i, rec)
    names = ["encode_latent", "compress_from_latent", "decompress(latent)", "decode_latent"]
    wall = [sum(r[k] for r in rec) / len(rec) for k in range(4)]
    # device time per stage: the library's own profiler (CUDA events around every launch), one staged step
    dev_ms = []
    L = _lib.lib
    def prof(fn):
        _lib.check(L.cra5_profile_enable(1))
        r = fn()
        sync()
        need = ctypes.c_uint64()
        buf = ctypes.create_string_buffer(1 << 20)
        _lib.check(L.cra5_profile_report(buf, ctypes.c_uint64(len(buf)), ctypes.byref(need)))
        _lib.check(L.cra5_profile_enable(0))
        return r, sum(k["ms"] for k in json.loads(buf.value.decode()).values())
    y, a = prof(lambda: net.encode_latent(frames[0], type="float")[0])
    out, b_ = prof(lambda: net.compress_from_latent(y))
    y_hat, c = prof(lambda: net.decompress(out["strings"], out["z_shape"], return_format="latent"))
    _, d = prof(lambda: net.decode_latent(y_hat))
    dev_ms = [a, b_, c, d]
    print(f"batch {B}: un-synchronised step {plain_ms:.2f} ms ({plain_ms / B:.2f} per frame); staged sum {sum(wall):.2f} ms")
    for k, n in enumerate(names):
        extra = f"  kernels {dev_ms[k]:8.2f} ms  host/idle {wall[k] - dev_ms[k]:6.2f} ms" if dev_ms else ""
        print(f"  {n:22s} wall {wall[k]:8.2f} ms{extra}")


if __name__ == "__main__":
    main()
