"""Synthetic stand-ins for what cannot be fetched offline (the cra5_268v_300k.pth checkpoint, ERA5 frames): the weight
regime bench.py and the CPU reference arm both run, so that the entropy stage sees trained-like statistics.

With plain random-init weights the hyperprior is unrelated to the latent: half of the sigma-hat values are negative
(clamped to the 0.11 lower bound), 44 % of the symbols leave their CDF row and take the bypass path, and a frame codes
to 4.9 MB -- twice a real frame, with the coder timing dominated by nibble loops a trained model rarely enters
(VERDICT round 1, weak item 7). `bench_regime` rescales four tensors so that, on N(0,1) frames,

    y      std 8, |y| up to ~40            (the notebook's real latents reach ~35)
    sigma  positive, log-uniform per channel over [2, 30] -- comparable to the spread of y - mu
    rate   ~2.1 MB per 268-variable frame  (real CRA5 frames: ~2.4 MB), ~0.7 % bypass symbols

(numbers from the fp32 oracle on the 268-variable model; the recipe is a fixed function of the config, no data).
"""
import math

import torch


def bench_regime(sd, cfg, seed=7):
    """returns a copy of state dict `sd` (reference layout, fp32) in the bench's entropy regime"""
    sd = {k: v.clone() for k, v in sd.items()}
    lat, Dh = cfg.latent_chans, cfg.hyper_dim
    g = torch.Generator().manual_seed(seed)
    # y: random-init g_a + quant_conv give std 0.5; x16 -> std 8
    sd["quant_conv.weight"] = sd["quant_conv.weight"] * 16.0
    # sigma-hat = W_sigma . LayerNorm(t): positive rows on a positive LayerNorm offset make it positive, one log-uniform
    # level per latent channel (E|w| = 0.0159 for the trunc-normal(0.02) init, summed over hyper_dim inputs)
    w = sd["h_s.final.weight"].reshape(-1, 2 * lat, Dh).clone()
    level = torch.exp(torch.empty(lat).uniform_(math.log(2.0), math.log(30.0), generator=g)) / (Dh * 0.0159)
    w[:, :lat] = w[:, :lat].abs() * level[None, :, None]
    sd["h_s.final.weight"] = w.reshape(-1, Dh)
    sd["h_s.norm.bias"] = torch.ones_like(sd["h_s.norm.bias"])
    return sd
