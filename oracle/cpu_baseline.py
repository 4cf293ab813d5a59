"""CPU timing of the reference algorithm for the hot path (TEST / MEASUREMENT INFRASTRUCTURE ONLY).

Used by bench.py for the `cpu_baseline` object and for `--impl reference`. The reference Python package cannot travel
to the GPU box, so the transforms are timed through the oracle port (oracle/vaeformer_oracle.py: the same torch CPU
fp32 operators the reference executes, `ATTENTION_MODE='math'`), and the entropy coder through the reference's own
compiled C++ coder when oracle/_ref is present (list-in / list-out pybind calls, exactly what
entropy_models.py:264-270 pays) or the C restatement otherwise.

A whole 268x721x1440 frame costs ~45 s on 8 cores, so one *sample* times every distinct layer type once on the full
frame and scales by the layer counts of the real model:
    encode = patch-embed + 3*(W24x24 + W12x48 + W48x12) + 4*G + quant_conv + h_a + EB + h_s + indexes + rANS(y,z)
    decode = rANS decode(z,y) + h_s + indexes + post_quant_conv + 3*(W24x24 + W12x48 + W48x12) + 3*G + LN + ConvTranspose
(13 encoder blocks = 9 windowed + 4 global, 12 decoder blocks = 9 windowed + 3 global; vit_nlc.py:401-422, 613-623).
"""
from __future__ import annotations

import os
import sys
import time
from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import entropy_oracle as EO, vaeformer_oracle as VO

_HERE = os.path.dirname(os.path.abspath(__file__))


def _block_sd(prefix, D, mlp, g):
    sd = OrderedDict()
    for k, shape in ((".norm1.weight", (D,)), (".norm1.bias", (D,)), (".attn.qkv.weight", (3 * D, D)),
                     (".attn.qkv.bias", (3 * D,)), (".attn.proj.weight", (D, D)), (".attn.proj.bias", (D,)),
                     (".norm2.weight", (D,)), (".norm2.bias", (D,)), (".mlp.fc1.weight", (mlp * D, D)),
                     (".mlp.fc1.bias", (mlp * D,)), (".mlp.fc2.weight", (D, mlp * D)), (".mlp.fc2.bias", (D,))):
        sd[prefix + k] = torch.randn(shape, generator=g) * (0.02 if "weight" in k and "norm" not in k else 0.1) \
            + (1.0 if k.endswith("norm1.weight") or k.endswith("norm2.weight") else 0.0)
    return sd


def _ref_coder():
    """the reference's own compiled coder (oracle/_ref), if it travelled with the repo"""
    ref = os.path.join(_HERE, "_ref")
    if not os.path.isdir(os.path.join(ref, "compressai")):
        return None
    try:
        import importlib.util
        import glob
        so = glob.glob(os.path.join(ref, "compressai", "ans*.so"))
        if not so:
            return None
        spec = importlib.util.spec_from_file_location("ans", so[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    except Exception:
        return None


class Sampler:
    def __init__(self, cfg, threads=None, seed=0):
        self.cfg = cfg
        self.threads = threads or os.cpu_count()
        torch.set_num_threads(self.threads)
        g = torch.Generator().manual_seed(seed)
        D, C, mlp = cfg.dim, cfg.in_chans, cfg.mlp_ratio
        ph, pw = cfg.patch_size
        sd = OrderedDict()
        sd["g_a.pos_embed"] = torch.randn(1, cfg.tokens, D, generator=g) * 0.02
        sd["g_a.patch_embed.proj.weight"] = torch.randn(D, C, ph, pw, generator=g) * 0.02
        sd["g_a.patch_embed.proj.bias"] = torch.zeros(D)
        wins = [w for w in cfg.enc_block_windows()]
        self.kinds = []  # (window or None, count_enc, count_dec)
        enc, dec = cfg.enc_block_windows(), cfg.dec_block_windows()
        for w in dict.fromkeys(enc + dec):
            self.kinds.append((w, enc.count(w), dec.count(w)))
        for i, (w, _, _) in enumerate(self.kinds):
            sd.update(_block_sd(f"blk.{i}", D, mlp, g))
        sd["g_s.norm.weight"] = torch.ones(D)
        sd["g_s.norm.bias"] = torch.zeros(D)
        sd["g_s.final.weight"] = torch.randn(D, C, ph, pw, generator=g) * 0.02
        sd["quant_conv.weight"] = torch.randn(2 * cfg.latent_chans, 2 * D, 1, 1, generator=g) * 0.12
        sd["quant_conv.bias"] = torch.zeros(2 * cfg.latent_chans)
        sd["post_quant_conv.weight"] = torch.randn(D, cfg.latent_chans, 1, 1, generator=g) * 0.02
        sd["post_quant_conv.bias"] = torch.zeros(D)
        # full hyperprior (cheap) from the regular seeded generator
        from cra5_b200 import config as Cfg
        from . import weights
        shapes = OrderedDict((k, v) for k, v in Cfg.param_shapes(cfg).items()
                             if k.startswith(("h_a.", "h_s.", "entropy_bottleneck.")))
        full = OrderedDict(shapes)
        sd.update(weights.seeded_state_dict(_with_dummy(shapes), seed))
        self.sd = sd
        self.eb = EO.entropy_bottleneck_tables(sd)
        self.gc = EO.gaussian_conditional_tables()
        self.ans = _ref_coder()
        self.coder_kind = "reference C++ coder (oracle/_ref)" if self.ans is not None else "C restatement (oracle/rans_oracle.c)"
        self.x = torch.randn(1, C, *cfg.img_size, generator=g)

    def _code(self, sym, idx, tab):
        """encode + decode one tensor the way EntropyModel.compress/decompress do (entropy_models.py:263-272, 317-327)"""
        if self.ans is not None:
            cdf, ln, off = tab.cdf.tolist(), tab.cdf_length.tolist(), tab.offset.tolist()
            s = self.ans.RansEncoder().encode_with_indexes(sym.reshape(-1).int().tolist(), idx.reshape(-1).int().tolist(),
                                                           cdf, ln, off)
            out = self.ans.RansDecoder().decode_with_indexes(s, idx.reshape(-1).int().tolist(), cdf, ln, off)
            return len(s), torch.tensor(out, dtype=torch.int32)
        s = EO.rans_encode(sym.reshape(-1), idx.reshape(-1), *tab.coder_args())
        return len(s), EO.rans_decode(s, idx.reshape(-1), *tab.coder_args())

    @torch.no_grad()
    def sample(self):
        """-> dict(encode_s, decode_s, total_s, parts) for ONE frame, extrapolated from one pass over each layer type"""
        cfg, sd = self.cfg, self.sd
        H, W = cfg.grid
        parts = {}

        def timed(name, fn):
            t0 = time.perf_counter()
            r = fn()
            parts[name] = time.perf_counter() - t0
            return r

        t = timed("patch_embed", lambda: F.conv2d(self.x, sd["g_a.patch_embed.proj.weight"], sd["g_a.patch_embed.proj.bias"],
                                                  stride=cfg.patch_stride).flatten(2).transpose(1, 2) + sd["g_a.pos_embed"])
        enc_blocks = dec_blocks = 0.0
        for i, (w, n_enc, n_dec) in enumerate(self.kinds):
            t = timed(f"block[{w}]", lambda: VO.block(sd, f"blk.{i}", t, cfg.num_heads, H, W, w, cfg.ln_eps))
            enc_blocks += n_enc * parts[f"block[{w}]"]
            dec_blocks += n_dec * parts[f"block[{w}]"]
        moments = torch.cat([t, t], 2).reshape(1, H, W, -1).permute(0, 3, 1, 2)
        y = timed("quant_conv", lambda: F.conv2d(moments, sd["quant_conv.weight"], sd["quant_conv.bias"]))[:, :cfg.latent_chans]
        y = y * (4.0 / y.std())  # spread over several quantisation bins so the coder does real work
        z = timed("h_a", lambda: VO.h_a(sd, cfg, y))
        med = sd["entropy_bottleneck.quantiles"][:, 0, 1].reshape(1, -1, 1, 1)
        zsym = EO.quantize_symbols(z, med)
        zidx = EO.eb_indexes(z.shape)
        _, zdec = timed("eb_code+decode", lambda: self._code(zsym, zidx, self.eb))
        z_hat = zdec.reshape(z.shape).float() + med
        scales, means = timed("h_s", lambda: VO.h_s(sd, cfg, z_hat))
        idx = timed("build_indexes", lambda: EO.build_indexes(scales, self.gc.scale_table))
        ysym = EO.quantize_symbols(y, means)
        nbytes, ydec = timed("gc_code+decode", lambda: self._code(ysym, idx, self.gc))
        assert torch.equal(ydec.reshape(-1), ysym.reshape(-1))
        y_hat = ydec.reshape(y.shape).float() + means
        t2 = timed("post_quant_conv", lambda: F.conv2d(y_hat, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"]))
        tok = t2.reshape(1, cfg.dim, -1).permute(0, 2, 1)
        tok = timed("final_norm", lambda: VO.layer_norm(tok, sd["g_s.norm.weight"], sd["g_s.norm.bias"], cfg.ln_eps))
        timed("conv_transpose", lambda: F.conv_transpose2d(tok.reshape(1, H, W, cfg.dim).permute(0, 3, 1, 2),
                                                           sd["g_s.final.weight"], None, stride=cfg.patch_stride))
        half = 0.5 * (parts["eb_code+decode"] + parts["gc_code+decode"])
        encode = parts["patch_embed"] + enc_blocks + parts["quant_conv"] + parts["h_a"] + parts["h_s"] + \
            parts["build_indexes"] + half
        decode = half + parts["h_s"] + parts["build_indexes"] + parts["post_quant_conv"] + dec_blocks + \
            parts["final_norm"] + parts["conv_transpose"]
        return dict(encode_s=encode, decode_s=decode, total_s=encode + decode, parts=parts, y_bytes=nbytes,
                    measured_s=sum(parts.values()))

    def describe(self):
        kinds = ", ".join(f"{'global' if w is None else 'window%dx%d' % w} x({e} enc,{d} dec)" for w, e, d in self.kinds)
        return (f"one pass over each distinct layer type on a full {self.cfg.in_chans}x{self.cfg.img_size[0]}x"
                f"{self.cfg.img_size[1]} frame (patch-embed, blocks [{kinds}], quant/post-quant conv, whole hyperprior, "
                f"entropy code+decode of y and z with the {self.coder_kind}, LayerNorm, ConvTranspose), scaled by the "
                f"model's layer counts; torch CPU fp32, math attention, {self.threads} threads")


def _with_dummy(shapes):
    return shapes
