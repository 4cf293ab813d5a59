"""Seeded synthetic parameters for parity work (TEST INFRASTRUCTURE ONLY).

The pretrained checkpoint (zoo/image.py:72) is unreachable offline, so fixtures are generated from weights that
both the reference (here, in the build container) and the tests (on the GPU box, without the reference) can
regenerate bit-identically from a seed: torch's CPU generator is platform independent.

Scales follow the reference initialisation (trunc_normal std 0.02 for Linear, vit_nlc.py:446-453) but biases and
LayerNorm affine terms are made non-trivial so a dropped bias or a swapped gamma/beta cannot go unnoticed, and
`quant_conv` / `h_s.final` are widened so that y spans many quantisation bins and sigma-hat spans many rows of
the scale table (last bullet).
"""
import math
from collections import OrderedDict

import torch


def _std_for(key: str, shape):
    if key.endswith("pos_embed"):
        return 0.02, 0.0
    if ".norm" in key and key.endswith("weight"):
        return 0.1, 1.0
    if ".norm" in key and key.endswith("bias"):
        return 0.1, 0.0
    if key.endswith("bias"):
        return 0.05, 0.0
    if key.startswith("quant_conv.weight"):
        return 0.12, 0.0
    if key.startswith("h_s.final.weight"):
        return 0.25, 0.0
    if key.startswith("h_a.quan_mlp.fc2.weight"):
        return 0.6, 0.0
    if "patch_embed.proj.weight" in key or key.startswith("g_s.final.weight"):
        return 0.02, 0.0
    return 0.04, 0.0


def seeded_state_dict(shapes: "OrderedDict[str, tuple]", seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for key, shape in shapes.items():
        if key.startswith("entropy_bottleneck."):
            continue
        std, mean = _std_for(key, shape)
        sd[key] = torch.randn(shape, generator=g, dtype=torch.float32) * std + mean
    # EntropyBottleneck: keep the reference's structural init (entropy_models.py:364-385) but give every channel its
    # own median / support so the per-channel CDF table is ragged.
    import numpy as np
    nz = shapes["entropy_bottleneck.quantiles"][0]
    filt = (1, 3, 3, 3, 3, 1)
    scale = 10.0 ** (1 / 5)
    for i in range(5):
        init = float(np.log(np.expm1(1 / scale / filt[i + 1])))
        sd[f"entropy_bottleneck._matrix{i}"] = torch.full(shapes[f"entropy_bottleneck._matrix{i}"], init) \
            + 0.05 * torch.randn(shapes[f"entropy_bottleneck._matrix{i}"], generator=g)
        sd[f"entropy_bottleneck._bias{i}"] = torch.rand(shapes[f"entropy_bottleneck._bias{i}"], generator=g) - 0.5
        if i < 4:
            sd[f"entropy_bottleneck._factor{i}"] = 0.1 * torch.randn(shapes[f"entropy_bottleneck._factor{i}"],
                                                                     generator=g)
    med = 0.5 * torch.randn(nz, generator=g)
    lo = med - (4.0 + 8.0 * torch.rand(nz, generator=g))
    hi = med + (4.0 + 8.0 * torch.rand(nz, generator=g))
    sd["entropy_bottleneck.quantiles"] = torch.stack([lo, med, hi], dim=1).reshape(nz, 1, 3).contiguous()
    return OrderedDict((k, sd[k]) for k in shapes)  # reference order


def seeded_frame(cfg, seed: int, smooth: bool = True) -> torch.Tensor:
    """normalised-space input frame (C, H, W): white noise plus a smooth large-scale component"""
    g = torch.Generator().manual_seed(1000 + seed)
    C, (H, W) = cfg.in_chans, cfg.img_size
    x = torch.randn(C, H, W, generator=g, dtype=torch.float32)
    if smooth:
        low = torch.randn(1, C, (H + 15) // 16 + 1, (W + 15) // 16 + 1, generator=g)
        low = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=True)[0]
        x = 0.5 * x + low
    return x.contiguous()


def synth_entropy_case(seed, n):
    """synthetic (y, sigma, mu) hitting every row of the 64-entry scale table and both bypass branches"""
    g = torch.Generator().manual_seed(seed)
    sig = torch.exp(torch.empty(n).uniform_(math.log(0.02), math.log(400.0), generator=g))
    mu = torch.randn(n, generator=g) * 3
    y = mu + torch.randn(n, generator=g) * sig * 1.3
    y[::97] += 4000.0 * torch.randn(y[::97].shape, generator=g)  # far tail -> bypass with many nibbles
    return y, sig, mu
