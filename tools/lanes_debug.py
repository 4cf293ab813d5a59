#!/usr/bin/env python
"""GPU diagnostic: where do results diverge when several codec lanes (cra5_b200.stream.CodecLanes) share one GPU?

Every kernel is deterministic, so each stage of each item must hash to the single-lane value. Prints, per model and lane
count, how many items differed at each stage (y latent / y string / z string / y_hat / x_hat) and the first stage that
differed, which names the kernel family to look at.

    python tools/lanes_debug.py [--lanes 2,3] [--items 24] [--big]
"""
import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from cra5_b200 import config as C
from cra5_b200.stream import CodecLanes
from cra5_b200.vaeformer import VAEformer
from oracle import weights

STAGES = ("y", "y_str", "z_str", "y_hat", "x_hat")


def sha(t):
    if isinstance(t, bytes):
        return hashlib.sha256(t).hexdigest()[:12]
    return hashlib.sha256(t.detach().cpu().numpy().tobytes()).hexdigest()[:12]


def stages(codec, x):
    with torch.no_grad():
        y, _, _ = codec.encode_latent(x, type="float")
        out = codec.compress_from_latent(y)
        y_hat = codec.decompress(out["strings"], out["z_shape"], return_format="latent")
        x_hat = codec.decode_latent(y_hat)
    return (sha(y), sha(out["strings"][0][0]), sha(out["strings"][1][0]), sha(y_hat), sha(x_hat))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lanes", default="2,3")
    ap.add_argument("--items", type=int, default=24)
    ap.add_argument("--big", action="store_true")
    a = ap.parse_args()
    models = [("small", C.small_lowres(5), 11, 3), ("tiny69", C.tiny_fullres(69), 7, 1)]
    summary = {}
    for name, cfg, wseed, fseed in models + ([("full268", C.cra5_268(), None, None)] if a.big else []):
        if wseed is None:
            net = VAEformer(268, cfg=cfg, init_seed=1234)
            x = torch.randn(1, cfg.in_chans, 721, 1440, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
        else:
            net = VAEformer(268, cfg=cfg, init_seed=None)
            net.load_state_dict(weights.seeded_state_dict(C.param_shapes(cfg), wseed))
            x = weights.seeded_frame(cfg, fseed).unsqueeze(0).cuda()
        net.update(force=True)
        torch.cuda.synchronize()
        base = stages(net, x)
        again = stages(net, x)
        summary[name] = {"single_lane_repeatable": base == again}
        for L in [int(v) for v in a.lanes.split(",")]:
            lanes = CodecLanes(net, lanes=L)
            n = a.items if wseed is not None else max(2 * L, a.items // 4)
            res = lanes.run(lambda codec, i: stages(codec, x), n)
            bad = {s: 0 for s in STAGES}
            first = {}
            for r in res:
                f = None
                for s, v, b in zip(STAGES, r, base):
                    if v != b:
                        bad[s] += 1
                        f = f or s
                if f:
                    first[f] = first.get(f, 0) + 1
            summary[name][f"lanes{L}"] = {"items": n, "differ": bad, "first_stage": first}
            del lanes
        del net
        torch.cuda.empty_cache()
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
