"""Model factory with the reference's entry point (cra5/models/compressai/zoo/image.py:302-324 `vaeformer_pretrained`,
:275-300 `_load_model`): same signature, same ValueErrors for bad metric / quality."""
from __future__ import annotations

import torch

from . import config as C
from .vaeformer import VAEformer

# the one published checkpoint (zoo/image.py:69-75)
model_urls = {"vaeformer-pretrained": {"mse": {268: "https://openmmlab.oss-cn-hangzhou.aliyuncs.com/cra5/cra5_268v_300k.pth"}}}

# reference ships only quality 268 (zoo/image.py:202-205); 159 / 69 are the same architecture at other channel counts
cfgs = {"vaeformer-pretrained": {268: C.cra5_268, 159: lambda: C.variant(159), 69: lambda: C.variant(69)}}


def _load_model(architecture, metric, quality, pretrained=False, progress=True, checkpoint=None, **kwargs):
    if architecture not in cfgs:
        raise ValueError(f'Invalid architecture name "{architecture}"')
    if quality not in cfgs[architecture]:
        raise ValueError(f'Invalid quality value "{quality}"')
    if pretrained or checkpoint is not None:
        if checkpoint is None:
            if metric not in model_urls[architecture] or quality not in model_urls[architecture][metric]:
                raise RuntimeError("Pre-trained model not yet available")
            state_dict = torch.hub.load_state_dict_from_url(model_urls[architecture][metric][quality], progress=progress,
                                                            map_location="cpu")
        else:
            state_dict = torch.load(checkpoint, map_location="cpu")
        if "state_dict" in state_dict:
            state_dict = state_dict["state_dict"]
        state_dict = {k[len("module."):] if k.startswith("module.") else k: v for k, v in state_dict.items()}
        return VAEformer.from_state_dict(state_dict, cfg=cfgs[architecture][quality](), **kwargs)
    return VAEformer(quality, cfg=cfgs[architecture][quality](), **kwargs)


def vaeformer_pretrained(quality, metric="mse", pretrained=False, progress=True, **kwargs):
    if metric not in ("mse", "ms-ssim"):
        raise ValueError(f'Invalid metric "{metric}"')
    if quality < 1 or quality > 999:
        raise ValueError(f'Invalid quality "{quality}", should be between (1, 999)')
    return _load_model("vaeformer-pretrained", metric, quality, pretrained, progress, **kwargs)
