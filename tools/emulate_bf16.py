#!/usr/bin/env python
"""CPU experiment (no GPU needed): how far does bf16-operand arithmetic move the FULL-SIZE 268-variable model away
from the fp32 oracle? Every linear / conv input and weight is rounded to bf16 (fp32 accumulation), q, k, v and the
un-normalised softmax numerators are rounded to bf16 -- the precision the tensor-core path of libcra5b200 works in --
and the eval-mode forward (oracle/vaeformer_oracle.py::OracleCodec.forward) is run in both precisions on one frame.

Prints the relative rms error of y, scales, means, x_hat, the fraction of latent symbols that change, and the north-star
figure: max over variables of |RMSE_bf16(c) - RMSE_fp32(c)| (RMSE against the input frame, normalised units).
It predicts what a full-size GPU-vs-oracle test would see; the GPU parity tests themselves run on the reduced-width
full-resolution fixture (tests/test_gpu_model.py), where the oracle takes seconds.

    python tools/emulate_bf16.py [--channels 268] [--threads 8]
"""
import argparse
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.nn.functional as F

from cra5_b200 import config as C
from cra5_b200.vaeformer import init_state_dict
from oracle import vaeformer_oracle as VO


def r(t):
    return t.bfloat16().float()


def make_bf16_functional():
    ns = types.SimpleNamespace(**{k: getattr(F, k) for k in ("layer_norm", "gelu", "pad")})
    ns.linear = lambda x, w, b=None: F.linear(r(x), r(w), b)
    ns.conv2d = lambda x, w, b=None, stride=1: F.conv2d(r(x), r(w), b, stride=stride)
    ns.conv_transpose2d = lambda x, w, b=None, stride=1: F.conv_transpose2d(r(x), r(w), b, stride=stride)
    return ns


def mhsa_bf16(qkv, heads):
    B, N, D3 = qkv.shape
    D = D3 // 3
    hd = D // heads
    qkv = qkv.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = r(qkv[0] * hd ** -0.5), r(qkv[1]), r(qkv[2])
    s = q @ k.transpose(-2, -1)
    p = torch.exp(s - s.amax(dim=-1, keepdim=True))
    out = (r(p) @ v) / p.sum(dim=-1, keepdim=True)
    return r(out.transpose(1, 2).reshape(B, N, D))      # the attention output feeds the projection GEMM as bf16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=268)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    a = ap.parse_args()
    torch.set_num_threads(a.threads)
    cfg = C.cra5_268() if a.channels == 268 else C.variant(a.channels)
    sd = init_state_dict(cfg, 3)
    sd["quant_conv.weight"] = sd["quant_conv.weight"] * 6.0      # the bench / API-test entropy regime
    sd["h_s.final.weight"] = sd["h_s.final.weight"] * 12.0
    codec = VO.OracleCodec(sd, cfg)
    x = torch.randn(1, cfg.in_chans, *cfg.img_size, generator=torch.Generator().manual_seed(1000))
    with torch.no_grad():
        t0 = time.time()
        ref = codec.forward(x)
        t1 = time.time()
        keepF, keepM = VO.F, VO.mhsa
        VO.F, VO.mhsa = make_bf16_functional(), mhsa_bf16
        try:
            emu = codec.forward(x)
        finally:
            VO.F, VO.mhsa = keepF, keepM
        t2 = time.time()
    print(f"fp32 forward {t1 - t0:.1f} s, bf16-emulated forward {t2 - t1:.1f} s, {a.threads} threads")

    def rel(k):
        d = (emu[k] - ref[k]).pow(2).mean().sqrt().item()
        return d / ref[k].pow(2).mean().sqrt().item()

    for k in ("y", "z", "scales", "means", "x_hat"):
        print(f"  rel rms error {k:7s} {rel(k):.3e}   (rms of reference {ref[k].pow(2).mean().sqrt().item():.3e})")
    flips = (torch.round(emu["y"] - emu["means"]) != torch.round(ref["y"] - ref["means"])).float().mean().item()
    print(f"  latent symbols that differ: {100 * flips:.2f} %")
    rm_e = ((emu["x_hat"][0] - x[0]) ** 2).mean(dim=(1, 2)).sqrt()
    rm_r = ((ref["x_hat"][0] - x[0]) ** 2).mean(dim=(1, 2)).sqrt()
    print(f"  per-variable RMSE: mean {rm_r.mean().item():.4f}, max |RMSE_bf16 - RMSE_fp32| = "
          f"{(rm_e - rm_r).abs().max().item():.3e}  (north-star bound 1e-4)")


if __name__ == "__main__":
    main()
