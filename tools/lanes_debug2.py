#!/usr/bin/env python
"""GPU diagnostic, second step: the full-size encoder output differs between a single-lane run and a multi-lane run.
Is it the stream (non-default) or the concurrency, how large is the difference, and does it survive CRA5_GEMM_PAIR=0?

    python tools/lanes_debug2.py            # parent: runs itself with CRA5_GEMM_PAIR unset / 0
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
    x = torch.randn(1, cfg.in_chans, 721, 1440, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    torch.cuda.synchronize()

    def enc(codec):
        with torch.no_grad():
            y, _, _ = codec.encode_latent(x, type="float")
        tok = codec.tap("tokens")
        return y.clone(), tok.clone()

    def cmp(a, b):
        d = (a - b).abs()
        return {"n_diff": int((a != b).sum().item()), "max_abs": float(d.max().item()),
                "rel_rms": float((d.pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item())}

    res = {}
    y0, t0 = enc(net)
    y1, t1 = enc(net)
    res["default_stream_repeat"] = {"y": cmp(y1, y0), "tokens": cmp(t1, t0)}
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        y2, t2 = enc(net)
    torch.cuda.synchronize()
    res["side_stream_alone"] = {"y": cmp(y2, y0), "tokens": cmp(t2, t0)}
    rep = net.replica()
    torch.cuda.synchronize()
    y3, t3 = enc(rep)
    res["replica_default_stream"] = {"y": cmp(y3, y0), "tokens": cmp(t3, t0)}
    # two host threads, two streams, concurrently
    out = [None, None]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for st in streams:
        st.wait_stream(torch.cuda.current_stream())

    def lane(k, codec):
        torch.cuda.set_device(0)
        with torch.cuda.stream(streams[k]):
            r = None
            for _ in range(3):
                r = enc(codec)
            out[k] = r
    th = threading.Thread(target=lane, args=(1, rep))
    th.start()
    lane(0, net)
    th.join()
    torch.cuda.synchronize()
    res["two_lanes_lane0"] = {"y": cmp(out[0][0], y0), "tokens": cmp(out[0][1], t0)}
    res["two_lanes_lane1"] = {"y": cmp(out[1][0], y0), "tokens": cmp(out[1][1], t0)}
    # same thread, two streams interleaved (no host threading)
    with torch.cuda.stream(streams[0]):
        ya, ta = enc(net)
    with torch.cuda.stream(streams[1]):
        yb, tb = enc(rep)
    torch.cuda.synchronize()
    res["one_thread_two_streams_a"] = {"y": cmp(ya, y0), "tokens": cmp(ta, t0)}
    res["one_thread_two_streams_b"] = {"y": cmp(yb, y0), "tokens": cmp(tb, t0)}
    print(json.dumps(res))


def main():
    for env in ({}, {"CRA5_GEMM_PAIR": "0"}):
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=e, capture_output=True, text=True,
                           timeout=400)
        print("ENV", env)
        print(r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else r.stderr[-2000:])


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
