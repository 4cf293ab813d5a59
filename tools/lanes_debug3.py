#!/usr/bin/env python
"""GPU diagnostic, third step: two encoders running CONCURRENTLY on one GPU give different results than one alone.
Host threading or GPU concurrency? Which stage? CRA5_DEBUG_STOP_AFTER=k stops the encoder after k trunk blocks.

    python tools/lanes_debug3.py
"""
import json
import os
import subprocess
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    from cra5_b200 import config as C
    from cra5_b200.vaeformer import VAEformer
    cfg = C.cra5_268()
    net = VAEformer(268, cfg=cfg, init_seed=1234)
    net.update(force=True)
    rep = net.replica()
    x = torch.randn(1, cfg.in_chans, 721, 1440, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    torch.cuda.synchronize()

    def enc(codec):
        with torch.no_grad():
            codec.encode_latent(x, type="float")

    NAMES = ("blk.q", "blk.k", "blk.vt", "blk.o", "blk.a", "blk.h", "tokens")

    def snap(codec):
        return {n: codec.tap(n).clone() for n in NAMES}

    def cmp(a, b):
        out = {}
        for n in NAMES:
            out[n] = int((a[n] != b[n]).sum().item())
        return out

    res = {}
    enc(net)
    t0 = snap(net)
    enc(rep)
    res["replica_alone"] = cmp(snap(rep), t0)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    # (1) ONE host thread, two streams, no synchronisation between the calls: GPU concurrency without host threading
    torch.cuda.synchronize()
    for _ in range(4):
        with torch.cuda.stream(streams[0]):
            enc(net)
        with torch.cuda.stream(streams[1]):
            enc(rep)
    torch.cuda.synchronize()
    res["one_thread_concurrent_a"] = cmp(snap(net), t0)
    res["one_thread_concurrent_b"] = cmp(snap(rep), t0)
    # (2) two host threads, each with its own stream
    def lane(k, codec):
        torch.cuda.set_device(0)
        with torch.cuda.stream(streams[k]):
            for _ in range(4):
                enc(codec)
    th = threading.Thread(target=lane, args=(1, rep))
    th.start()
    lane(0, net)
    th.join()
    torch.cuda.synchronize()
    res["two_threads_a"] = cmp(snap(net), t0)
    res["two_threads_b"] = cmp(snap(rep), t0)
    # (3) two host threads but the GPU work serialised: both lanes on the SAME stream
    def lane_same(codec):
        torch.cuda.set_device(0)
        with torch.cuda.stream(streams[0]):
            for _ in range(4):
                enc(codec)
    th = threading.Thread(target=lane_same, args=(rep,))
    th.start()
    lane_same(net)
    th.join()
    torch.cuda.synchronize()
    res["two_threads_same_stream_a"] = cmp(snap(net), t0)
    res["two_threads_same_stream_b"] = cmp(snap(rep), t0)
    print(json.dumps(res))


def main():
    for k in ("3", "4"):
        e = dict(os.environ)
        e["CRA5_DEBUG_STOP_AFTER"] = k
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=e, capture_output=True, text=True,
                           timeout=300)
        print("STOP_AFTER", k)
        print(r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else r.stderr[-2000:], flush=True)


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
