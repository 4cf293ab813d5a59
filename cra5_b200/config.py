"""Model geometry of the VAEformer codec and the state-dict schema that goes with it.

The reference hard-codes its one shipped variant inside `VAEformer.__init__`
(cra5/models/vaeformer/vaeformer.py:93-142) and derives the transformer sizes from the `vit_large`
defaults of `Encoder()` / `Decoder()` (cra5/models/vaeformer/vit_nlc.py:1009-1015). Here the same numbers
live in one dataclass so that other channel counts (159, 69 -- BASELINE.json configs 1 and 5) and the
reduced-width test models are first-class.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field, asdict
from typing import List, Tuple


@dataclass
class VaeformerConfig:
    in_chans: int = 268
    img_size: Tuple[int, int] = (721, 1440)
    patch_size: Tuple[int, int] = (11, 10)
    patch_stride: Tuple[int, int] = (10, 10)
    dim: int = 1024              # transformer width ("y_channels" in the reference, vaeformer.py:96)
    depth: int = 24              # encoder uses depth//2 (+1 extra head block), decoder depth//2
    num_heads: int = 16
    mlp_ratio: int = 4
    window_sizes: List[Tuple[int, int]] = field(default_factory=lambda: [(24, 24), (12, 48), (48, 12)])
    interval: int = 4            # every interval-th block is global (vit_nlc.py:406-407)
    latent_chans: int = 256      # channels of y ("embed_dim" in the reference, vaeformer.py:94)
    z_chans: int = 256           # channels of z (vaeformer.py:95)
    hyper_dim: int = 360
    hyper_depth: int = 8         # h_a and h_s get hyper_depth//2 global blocks each
    hyper_heads: int = 5
    hyper_patch: Tuple[int, int] = (4, 4)
    ln_eps: float = 1e-6

    # ---- derived geometry
    @property
    def grid(self) -> Tuple[int, int]:
        """token grid of g_a / g_s = conv output size, floor((img - k)/s) + 1 (vit_nlc.py:302-307)"""
        return ((self.img_size[0] - self.patch_size[0]) // self.patch_stride[0] + 1,
                (self.img_size[1] - self.patch_size[1]) // self.patch_stride[1] + 1)

    @property
    def tokens(self) -> int:
        return self.grid[0] * self.grid[1]

    @property
    def hyper_grid(self) -> Tuple[int, int]:
        return (self.grid[0] // self.hyper_patch[0], self.grid[1] // self.hyper_patch[1])

    @property
    def hyper_tokens(self) -> int:
        return self.hyper_grid[0] * self.hyper_grid[1]

    @property
    def conv_head(self) -> bool:
        """g_s ends in ConvTranspose2d only for the (721,1440) geometry, else Linear + rearrange
        (vit_nlc.py:628-632, 665-680)."""
        return tuple(self.img_size) == (721, 1440)

    @property
    def enc_blocks(self) -> int:
        return self.depth // 2 + 1

    @property
    def dec_blocks(self) -> int:
        return self.depth - self.depth // 2

    def block_window(self, i: int):
        """window (h, w) of trunk block with absolute index i, or None for a global block
        (vit_nlc.py:401-407 encoder, :613-619 decoder)."""
        if (i + 1) % self.interval == 0:
            return None
        return tuple(self.window_sizes[min(i % self.interval, len(self.window_sizes) - 1)])

    def enc_block_windows(self):
        n = self.depth // 2
        w = [self.block_window(i) for i in range(n)]
        return w + [w[-1]]  # the extra "logvar" block copies the last block's setting (vit_nlc.py:413-422)

    def dec_block_windows(self):
        return [self.block_window(i) for i in range(self.depth // 2, self.depth)]

    @property
    def hyper_hidden(self) -> int:
        """hidden width of quan_mlp / post_quan_mlp: int(sqrt(embed//z)) * z (vit_nlc.py:544-546, 609-611)"""
        import math
        return int(math.sqrt(self.hyper_dim // self.z_chans)) * self.z_chans

    def validate(self):
        if self.dim % self.num_heads or self.hyper_dim % self.hyper_heads:
            raise ValueError("width must be divisible by the number of heads")
        if self.patch_size[1] != self.patch_stride[1]:
            raise ValueError("unsupported geometry: horizontal patch overlap")
        if not (self.patch_stride[0] <= self.patch_size[0] <= 2 * self.patch_stride[0]):
            raise ValueError("unsupported geometry: vertical patch size must be in [stride, 2*stride]")
        if not self.conv_head:
            if self.patch_size != self.patch_stride:
                raise ValueError("Linear head (non-721x1440 images) needs patch_size == patch_stride")
        if self.grid[0] % self.hyper_patch[0] or self.grid[1] % self.hyper_patch[1]:
            raise ValueError("token grid must be divisible by the hyperprior patch")
        return self

    def to_dict(self):
        d = asdict(self)
        d["window_sizes"] = [list(w) for w in self.window_sizes]
        return d


def cra5_268() -> VaeformerConfig:
    """the shipped model (vaeformer.py:93-142)"""
    return VaeformerConfig().validate()


def variant(in_chans: int) -> VaeformerConfig:
    """same architecture, different variable count (159: config/vaeformer_era5_159v_1h.py:41-50; 69)"""
    return VaeformerConfig(in_chans=in_chans).validate()


def tiny_fullres(in_chans: int = 69) -> VaeformerConfig:
    """reduced-width model at FULL resolution: every shape quirk of the real model (721x1440, (11,10)/(10,10)
    patches with the one-row overlap, all three windows incl. the padded (48,12) one, ConvTranspose head, 18x36
    hyper grid) at ~1/200 of the compute. Used for golden fixtures."""
    return VaeformerConfig(in_chans=in_chans, dim=128, depth=8, num_heads=2, latent_chans=32, z_chans=32,
                           hyper_dim=48, hyper_depth=4, hyper_heads=2).validate()


def small_lowres(in_chans: int = 5) -> VaeformerConfig:
    """small image, Linear head path (vit_nlc.py:632, 671-680): whole tensors fit in a committed fixture"""
    return VaeformerConfig(in_chans=in_chans, img_size=(160, 320), patch_size=(10, 10), patch_stride=(10, 10),
                           dim=128, depth=8, num_heads=2, latent_chans=32, z_chans=32,
                           window_sizes=[(8, 8), (4, 16), (12, 4)], hyper_dim=48, hyper_depth=4,
                           hyper_heads=2).validate()


# --------------------------------------------------------------------------------------------------------------
# state-dict schema. Keys and shapes are those of the reference's `state_dict()`.
# --------------------------------------------------------------------------------------------------------------

def _block_shapes(prefix: str, dim: int, mlp_ratio: int) -> "OrderedDict[str, tuple]":
    d = OrderedDict()
    d[f"{prefix}.norm1.weight"] = (dim,)
    d[f"{prefix}.norm1.bias"] = (dim,)
    d[f"{prefix}.attn.qkv.weight"] = (3 * dim, dim)
    d[f"{prefix}.attn.qkv.bias"] = (3 * dim,)
    d[f"{prefix}.attn.proj.weight"] = (dim, dim)
    d[f"{prefix}.attn.proj.bias"] = (dim,)
    d[f"{prefix}.norm2.weight"] = (dim,)
    d[f"{prefix}.norm2.bias"] = (dim,)
    d[f"{prefix}.mlp.fc1.weight"] = (mlp_ratio * dim, dim)
    d[f"{prefix}.mlp.fc1.bias"] = (mlp_ratio * dim,)
    d[f"{prefix}.mlp.fc2.weight"] = (dim, mlp_ratio * dim)
    d[f"{prefix}.mlp.fc2.bias"] = (dim,)
    return d


EB_FILTERS = (3, 3, 3, 3)  # entropy_models.py:353


def param_shapes(cfg: VaeformerConfig) -> "OrderedDict[str, tuple]":
    """float parameters of the model, in the reference's state_dict order (int CDF buffers and the constant
    `target` / `*.bound` buffers excluded -- see `buffer_keys`)."""
    C, D = cfg.in_chans, cfg.dim
    ph, pw = cfg.patch_size
    d = OrderedDict()
    filt = (1,) + EB_FILTERS + (1,)
    for i in range(len(EB_FILTERS) + 1):
        d[f"entropy_bottleneck._matrix{i}"] = (cfg.z_chans, filt[i + 1], filt[i])
        d[f"entropy_bottleneck._bias{i}"] = (cfg.z_chans, filt[i + 1], 1)
        if i < len(EB_FILTERS):
            d[f"entropy_bottleneck._factor{i}"] = (cfg.z_chans, filt[i + 1], 1)
    d["entropy_bottleneck.quantiles"] = (cfg.z_chans, 1, 3)
    d["g_a.pos_embed"] = (1, cfg.tokens, D)
    d["g_a.patch_embed.proj.weight"] = (D, C, ph, pw)
    d["g_a.patch_embed.proj.bias"] = (D,)
    for i in range(cfg.enc_blocks):
        d.update(_block_shapes(f"g_a.blocks.{i}", D, cfg.mlp_ratio))
    for i in range(cfg.dec_blocks):
        d.update(_block_shapes(f"g_s.blocks.{i}", D, cfg.mlp_ratio))
    d["g_s.norm.weight"] = (D,)
    d["g_s.norm.bias"] = (D,)
    if cfg.conv_head:
        d["g_s.final.weight"] = (D, C, ph, pw)          # ConvTranspose2d layout (in, out, kh, kw)
    else:
        d["g_s.final.weight"] = (C * ph * pw, D)        # Linear, no bias
    d["quant_conv.weight"] = (2 * cfg.latent_chans, 2 * D, 1, 1)
    d["quant_conv.bias"] = (2 * cfg.latent_chans,)
    d["post_quant_conv.weight"] = (D, cfg.latent_chans, 1, 1)
    d["post_quant_conv.bias"] = (D,)
    Dh, hp = cfg.hyper_dim, cfg.hyper_patch
    d["h_a.pos_embed"] = (1, cfg.hyper_tokens, Dh)
    d["h_a.patch_embed.proj.weight"] = (Dh, cfg.latent_chans, hp[0], hp[1])
    d["h_a.patch_embed.proj.bias"] = (Dh,)
    for i in range(cfg.hyper_depth // 2):
        d.update(_block_shapes(f"h_a.blocks.{i}", Dh, cfg.mlp_ratio))
    hid = cfg.hyper_hidden
    d["h_a.quan_mlp.fc1.weight"] = (hid, Dh)
    d["h_a.quan_mlp.fc1.bias"] = (hid,)
    d["h_a.quan_mlp.fc2.weight"] = (cfg.z_chans, hid)
    d["h_a.quan_mlp.fc2.bias"] = (cfg.z_chans,)
    d["h_s.post_quan_mlp.fc1.weight"] = (hid, cfg.z_chans)
    d["h_s.post_quan_mlp.fc1.bias"] = (hid,)
    d["h_s.post_quan_mlp.fc2.weight"] = (Dh, hid)
    d["h_s.post_quan_mlp.fc2.bias"] = (Dh,)
    for i in range(cfg.hyper_depth - cfg.hyper_depth // 2):
        d.update(_block_shapes(f"h_s.blocks.{i}", Dh, cfg.mlp_ratio))
    d["h_s.norm.weight"] = (Dh,)
    d["h_s.norm.bias"] = (Dh,)
    d["h_s.final.weight"] = (2 * cfg.latent_chans * hp[0] * hp[1], Dh)
    return d


# buffers that ride along in a reference checkpoint (models/base.py:73-87); the CDF ones are sized by update()
BUFFER_KEYS = (
    "entropy_bottleneck._offset", "entropy_bottleneck._quantized_cdf", "entropy_bottleneck._cdf_length",
    "entropy_bottleneck.target", "entropy_bottleneck.likelihood_lower_bound.bound",
    "gaussian_conditional._offset", "gaussian_conditional._quantized_cdf", "gaussian_conditional._cdf_length",
    "gaussian_conditional.scale_table", "gaussian_conditional.scale_bound",
    "gaussian_conditional.likelihood_lower_bound.bound", "gaussian_conditional.lower_bound_scale.bound",
)


def config_from_state_dict(sd) -> VaeformerConfig:
    """infer the geometry from a reference checkpoint (mirrors VAEformer.from_state_dict reading in_chans off
    `g_a.patch_embed.proj.weight`, vaeformer.py:172)."""
    w = sd["g_a.patch_embed.proj.weight"]
    D, C, ph, pw = tuple(w.shape)
    n_enc = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("g_a.blocks."))
    depth = 2 * (n_enc - 1)
    lat = sd["post_quant_conv.weight"].shape[1]
    Dh = sd["h_a.patch_embed.proj.weight"].shape[0]
    n_h = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("h_a.blocks."))
    cfg = VaeformerConfig(in_chans=C, dim=D, depth=depth, patch_size=(ph, pw), latent_chans=lat,
                          z_chans=sd["entropy_bottleneck.quantiles"].shape[0], hyper_dim=Dh,
                          hyper_depth=2 * n_h)
    if D != 1024 or Dh != 360:
        raise ValueError("cannot infer head counts for a non-standard width; pass a VaeformerConfig explicitly")
    return cfg.validate()
