"""oracle shim (TEST INFRASTRUCTURE ONLY) for `timm.models.layers`.

The reference needs only `drop_path`, `to_2tuple`, `trunc_normal_`
(/root/reference/cra5/models/vaeformer/vit_nlc.py:23). None of them takes part in
inference arithmetic (drop-path rate is 0, trunc_normal_ is weight init).
"""
import collections.abc
import torch


def to_2tuple(x):
    if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
        return tuple(x)
    return (x, x)


def drop_path(x, drop_prob=0.0, training=False):
    if drop_prob == 0.0 or not training:
        return x
    keep = 1.0 - drop_prob
    mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
    return x * mask / keep


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class DropPath(torch.nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training)
