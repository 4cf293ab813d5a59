"""CPU fp32 restatement of the VAEformer transforms and codec orchestration (TEST INFRASTRUCTURE ONLY).

Functional (state-dict in, tensors out), torch CPU fp32, no dependency on the reference package, so it travels to
the GPU box where /root/reference does not exist. Pinned against the real reference, imported through the shims in
the build container, by tools/make_golden.py (which also stores the reference's own outputs under tests/golden/).

Every function cites the reference lines it restates (paths relative to the CRA5 repository).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import entropy_oracle as EO


# ------------------------------------------------------------------------------------------------ blocks
def layer_norm(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def mlp(sd, p, x):
    """Mlp.forward, vit_nlc.py:62-69: fc1 -> exact-erf GELU -> fc2"""
    h = F.linear(x, sd[f"{p}.fc1.weight"], sd[f"{p}.fc1.bias"])
    h = F.gelu(h)
    return F.linear(h, sd[f"{p}.fc2.weight"], sd.get(f"{p}.fc2.bias"))


def mhsa(qkv, heads):
    """softmax((q*scale) k^T) v on a (B, N, 3*D) qkv tensor laid out [q|k|v][head][dim]
    (vit_nlc.py:99-103 and :242-246)"""
    B, N, D3 = qkv.shape
    D = D3 // 3
    hd = D // heads
    qkv = qkv.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    attn = torch.softmax(attn, dim=-1)
    return (attn @ v).transpose(1, 2).reshape(B, N, D)


def global_attention(sd, p, x, heads):
    """Attention.forward (math mode), vit_nlc.py:94-112"""
    qkv = F.linear(x, sd[f"{p}.qkv.weight"], sd[f"{p}.qkv.bias"])
    return F.linear(mhsa(qkv, heads), sd[f"{p}.proj.weight"], sd[f"{p}.proj.bias"])


def window_attention(sd, p, x, heads, H, W, win):
    """WindowAttention.forward, vit_nlc.py:219-258. The normalised tokens are zero-padded on the bottom/right to a
    multiple of the window and the pad tokens take part in the softmax UNMASKED (their q/k/v are the qkv bias)."""
    B, N, D = x.shape
    wh, ww = win
    x = x.reshape(B, H, W, D)
    pad_b = (wh - H % wh) % wh
    pad_r = (ww - W % ww) % ww
    x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))
    Hp, Wp = H + pad_b, W + pad_r
    # window_partition, vit_nlc.py:115-126
    x = x.reshape(B, Hp // wh, wh, Wp // ww, ww, D).permute(0, 1, 3, 2, 4, 5).reshape(-1, wh * ww, D)
    qkv = F.linear(x, sd[f"{p}.qkv.weight"], sd[f"{p}.qkv.bias"])
    x = F.linear(mhsa(qkv, heads), sd[f"{p}.proj.weight"], sd[f"{p}.proj.bias"])
    # window_reverse + crop, vit_nlc.py:129-142, 250-256
    x = x.reshape(B, Hp // wh, Wp // ww, wh, ww, D).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, D)
    return x[:, :H, :W, :].reshape(B, H * W, D)


def block(sd, p, x, heads, H, W, win, eps):
    """Block.forward, vit_nlc.py:282-287 (drop-path rate 0)"""
    h = layer_norm(x, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], eps)
    if win is None:
        x = x + global_attention(sd, f"{p}.attn", h, heads)
    else:
        x = x + window_attention(sd, f"{p}.attn", h, heads, H, W, win)
    h = layer_norm(x, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], eps)
    return x + mlp(sd, f"{p}.mlp", h)


# ------------------------------------------------------------------------------------------------ transforms
def g_a(sd, cfg, x, taps=None):
    """ViT_Encoder.forward, vit_nlc.py:458-486: patch-embed conv + pos-embed, depth//2 - 1 sequential blocks, then the
    last two blocks run IN PARALLEL on the same input and are concatenated (mean || logvar)."""
    B = x.shape[0]
    t = F.conv2d(x, sd["g_a.patch_embed.proj.weight"], sd["g_a.patch_embed.proj.bias"], stride=cfg.patch_stride)
    H, W = t.shape[2], t.shape[3]
    t = t.flatten(2).transpose(1, 2) + sd["g_a.pos_embed"]
    if taps is not None:
        taps["g_a.embed"] = t
    wins = cfg.enc_block_windows()
    n = cfg.enc_blocks
    for i in range(n - 2):
        t = block(sd, f"g_a.blocks.{i}", t, cfg.num_heads, H, W, wins[i], cfg.ln_eps)
        if taps is not None:
            taps[f"g_a.blocks.{i}"] = t
    mean = block(sd, f"g_a.blocks.{n - 2}", t, cfg.num_heads, H, W, wins[n - 2], cfg.ln_eps)
    logvar = block(sd, f"g_a.blocks.{n - 1}", t, cfg.num_heads, H, W, wins[n - 1], cfg.ln_eps)
    t = torch.cat([mean, logvar], 2)
    return t.reshape(B, H, W, -1).permute(0, 3, 1, 2)


def encode_y(sd, cfg, x, taps=None):
    """VAEformer.encode_latent(type='float'), vaeformer.py:272-292: quant_conv then
    DiagonalGaussianDistribution.mode() == first half of the channels (distributions.py:30-32, 71-72)"""
    moments = F.conv2d(g_a(sd, cfg, x, taps), sd["quant_conv.weight"], sd["quant_conv.bias"])
    return moments[:, : cfg.latent_chans]


def g_s(sd, cfg, feat, taps=None):
    """ViT_Decoder.forward, vit_nlc.py:655-693: no pos-embed, depth//2 blocks, LayerNorm, ConvTranspose2d head
    (721x1440 only) or Linear + (p1 p2 c) rearrange."""
    B, D, H, W = feat.shape
    t = feat.reshape(B, D, -1).permute(0, 2, 1)
    wins = cfg.dec_block_windows()
    for i in range(cfg.dec_blocks):
        t = block(sd, f"g_s.blocks.{i}", t, cfg.num_heads, H, W, wins[i], cfg.ln_eps)
        if taps is not None:
            taps[f"g_s.blocks.{i}"] = t
    t = layer_norm(t, sd["g_s.norm.weight"], sd["g_s.norm.bias"], cfg.ln_eps)
    if cfg.conv_head:
        t = t.reshape(B, H, W, D).permute(0, 3, 1, 2)
        return F.conv_transpose2d(t, sd["g_s.final.weight"], None, stride=cfg.patch_stride)
    t = F.linear(t, sd["g_s.final.weight"])
    p1, p2 = cfg.patch_size
    C = cfg.in_chans
    t = t.reshape(B, H, W, p1, p2, C).permute(0, 5, 1, 3, 2, 4)
    return t.reshape(B, C, H * p1, W * p2)


def decode_y(sd, cfg, y_hat, taps=None):
    """VAEformer.decode_latent, vaeformer.py:294-300"""
    return g_s(sd, cfg, F.conv2d(y_hat, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"]), taps)


def h_a(sd, cfg, y):
    """HyperpriorEncoder (vit_nlc.py:488-551) through ViT_Encoder.forward (:477-486): conv patch-embed + pos-embed,
    global blocks, quan_mlp, back to NCHW"""
    B = y.shape[0]
    t = F.conv2d(y, sd["h_a.patch_embed.proj.weight"], sd["h_a.patch_embed.proj.bias"], stride=cfg.hyper_patch)
    H, W = t.shape[2], t.shape[3]
    t = t.flatten(2).transpose(1, 2) + sd["h_a.pos_embed"]
    for i in range(cfg.hyper_depth // 2):
        t = block(sd, f"h_a.blocks.{i}", t, cfg.hyper_heads, H, W, None, cfg.ln_eps)
    t = mlp(sd, "h_a.quan_mlp", t)
    return t.reshape(B, H, W, -1).permute(0, 3, 1, 2)


def h_s(sd, cfg, z_hat):
    """HyperpriorDecoder (vit_nlc.py:696-748) through ViT_Decoder.forward (:682-693, :671-680): post_quan_mlp, global
    blocks, LayerNorm, Linear(no bias) to 2*C*p1*p2, pixel-shuffle rearrange '(p1 p2 c)'. Returns (scales, means)
    = chunk(2, dim=1) (vaeformer.py:369)."""
    B, Cz, H, W = z_hat.shape
    t = z_hat.reshape(B, Cz, -1).permute(0, 2, 1)
    t = mlp(sd, "h_s.post_quan_mlp", t)
    for i in range(cfg.hyper_depth - cfg.hyper_depth // 2):
        t = block(sd, f"h_s.blocks.{i}", t, cfg.hyper_heads, H, W, None, cfg.ln_eps)
    t = layer_norm(t, sd["h_s.norm.weight"], sd["h_s.norm.bias"], cfg.ln_eps)
    t = F.linear(t, sd["h_s.final.weight"])
    p1, p2 = cfg.hyper_patch
    C2 = 2 * cfg.latent_chans
    t = t.reshape(B, H, W, p1, p2, C2).permute(0, 5, 1, 3, 2, 4).reshape(B, C2, H * p1, W * p2)
    return t[:, : cfg.latent_chans], t[:, cfg.latent_chans:]


# ------------------------------------------------------------------------------------------------ codec
class OracleCodec:
    """compress / decompress exactly as VAEformer does (vaeformer.py:334-400) with the reference stream format:
    one sequential rANS stream per tensor, produced by the C restatement of the coder (oracle/rans_oracle.c)."""

    def __init__(self, sd, cfg):
        self.sd = {k: v.float() for k, v in sd.items() if v.is_floating_point()}
        self.cfg = cfg
        self.eb = EO.entropy_bottleneck_tables(self.sd)
        self.gc = EO.gaussian_conditional_tables()

    # ---- the pieces
    def z_symbols(self, z):
        med = self.sd["entropy_bottleneck.quantiles"][:, 0, 1]
        return EO.quantize_symbols(z, med.reshape(1, -1, 1, 1))

    def compress_from_latent(self, y):
        """VAEformer.compress_from_latent, vaeformer.py:334-348"""
        cfg, sd = self.cfg, self.sd
        z = h_a(sd, cfg, y)
        zsym = self.z_symbols(z)
        zidx = EO.eb_indexes(z.shape)
        z_strings = [EO.rans_encode(zsym[i].reshape(-1), zidx[i].reshape(-1), *self.eb.coder_args())
                     for i in range(z.shape[0])]
        med = sd["entropy_bottleneck.quantiles"][:, 0, 1].reshape(1, -1, 1, 1)
        z_hat = zsym.float() + med  # == decompress(compress(z)) since the coder is lossless
        scales, means = h_s(sd, cfg, z_hat)
        idx = EO.build_indexes(scales, self.gc.scale_table)
        ysym = EO.quantize_symbols(y, means)
        y_strings = [EO.rans_encode(ysym[i].reshape(-1), idx[i].reshape(-1), *self.gc.coder_args())
                     for i in range(y.shape[0])]
        return {"strings": [y_strings, z_strings], "z_shape": tuple(z.shape[-2:]),
                "debug": dict(z=z, z_symbols=zsym, z_hat=z_hat, scales=scales, means=means, indexes=idx,
                              y_symbols=ysym)}

    def compress(self, x):
        """VAEformer.compress, vaeformer.py:350-376"""
        y = encode_y(self.sd, self.cfg, x)
        out = self.compress_from_latent(y)
        out["debug"]["y"] = y
        return out

    def decompress(self, strings, shape, return_format="reconstructed"):
        """VAEformer.decompress, vaeformer.py:378-400"""
        cfg, sd = self.cfg, self.sd
        B = len(strings[1])
        zshape = (B, cfg.z_chans, shape[0], shape[1])
        zidx = EO.eb_indexes(zshape)
        med = sd["entropy_bottleneck.quantiles"][:, 0, 1].reshape(1, -1, 1, 1)
        zsym = torch.stack([EO.rans_decode(strings[1][i], zidx[i].reshape(-1), *self.eb.coder_args())
                            .reshape(zshape[1:]) for i in range(B)])
        z_hat = zsym.float() + med
        scales, means = h_s(sd, cfg, z_hat)
        idx = EO.build_indexes(scales, self.gc.scale_table)
        ysym = torch.stack([EO.rans_decode(strings[0][i], idx[i].reshape(-1), *self.gc.coder_args())
                            .reshape(idx.shape[1:]) for i in range(B)])
        y_hat = ysym.float() + means  # EntropyModel.dequantize, entropy_models.py:193-201
        if return_format == "latent":
            return y_hat
        return {"x_hat": decode_y(sd, cfg, y_hat)}

    def forward(self, x):
        """VAEformer.forward (eval), vaeformer.py:302-333: dequantize path without the coder"""
        cfg, sd = self.cfg, self.sd
        y = encode_y(sd, cfg, x)
        z = h_a(sd, cfg, y)
        med = sd["entropy_bottleneck.quantiles"][:, 0, 1].reshape(1, -1, 1, 1)
        z_hat = torch.round(z - med) + med
        scales, means = h_s(sd, cfg, z_hat)
        y_hat = torch.round(y - means) + means
        return {"x_hat": decode_y(sd, cfg, y_hat), "y": y, "y_hat": y_hat, "z": z, "z_hat": z_hat,
                "scales": scales, "means": means,
                "likelihoods": {"y": EO.gc_likelihood(y_hat, scales, means), "z": EO.eb_likelihood(sd, z_hat)}}
