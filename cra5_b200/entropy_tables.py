"""Construction of the integer CDF tables consumed by the GPU range coder (run once per model, at `update()`).

Host-side mirror of
  GaussianConditional.update / _standardized_cumulative   cra5/models/compressai/entropy_models/entropy_models.py:598-643
  EntropyBottleneck.update / _logits_cumulative           entropy_models.py:394-463
  EntropyModel._pmf_to_cdf                                entropy_models.py:208-216
  get_scale_table                                         cra5/models/compressai/models/base.py:54-61

The float pmf is evaluated with the same torch CPU fp32 operations the reference uses (the integer tables must match
it to the last count, and that pins the erfc / sigmoid implementation); the float -> integer step is the C-ABI call
cra5_pmf_to_quantized_cdf of libcra5b200.so.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64
SCALE_BOUND = 0.11
TAIL_MASS = 1e-9
PRECISION = 16
EB_FILTERS = (3, 3, 3, 3)


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS) -> torch.Tensor:
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


def pmf_to_quantized_cdf(pmf: torch.Tensor, precision: int = PRECISION) -> torch.Tensor:
    p = np.ascontiguousarray(pmf.detach().cpu().numpy().astype(np.float32))
    out = np.zeros(p.size + 1, dtype=np.uint32)
    _lib.check(_lib.lib.cra5_pmf_to_quantized_cdf(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(p.size),
                                                  ctypes.c_int(precision), out.ctypes.data_as(ctypes.c_void_p)))
    return torch.from_numpy(out.astype(np.int32))


@dataclass
class CdfTables:
    quantized_cdf: torch.Tensor  # (rows, max_len + 2) int32
    cdf_length: torch.Tensor     # (rows,) int32
    offset: torch.Tensor         # (rows,) int32


def _rows_to_table(pmf, tail_mass, pmf_length, max_length) -> torch.Tensor:
    table = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
    for r in range(len(pmf_length)):
        n = int(pmf_length[r])
        row = pmf_to_quantized_cdf(torch.cat((pmf[r, :n], tail_mass[r].reshape(1))))
        table[r, : row.numel()] = row
    return table


def _phi(x: torch.Tensor) -> torch.Tensor:
    # standard normal CDF through erfc, as the reference does for numerical precision in the tails
    return 0.5 * torch.erfc(float(-(2 ** -0.5)) * x)


def gaussian_conditional_tables(scale_table: torch.Tensor) -> CdfTables:
    """one zero-mean discretised Gaussian per scale level, support +-ceil(sigma * Phi^-1(1 - tail/2))"""
    import scipy.stats
    scale_table = scale_table.detach().float().cpu()
    multiplier = -scipy.stats.norm.ppf(TAIL_MASS / 2)
    center = torch.ceil(scale_table * multiplier).int()
    length = 2 * center + 1
    max_length = int(length.max())
    dist = torch.abs(torch.arange(max_length).int() - center[:, None]).float()
    sigma = scale_table.unsqueeze(1)
    upper = _phi((0.5 - dist) / sigma)
    lower = _phi((-0.5 - dist) / sigma)
    pmf = upper - lower
    tail = 2 * lower[:, 0]
    return CdfTables(_rows_to_table(pmf, tail, length, max_length), (length + 2).int(), (-center).int())


def _logits_cumulative(sd, x: torch.Tensor) -> torch.Tensor:
    logits = x
    n = len(EB_FILTERS) + 1
    for i in range(n):
        logits = torch.matmul(F.softplus(sd[f"entropy_bottleneck._matrix{i}"].float().cpu()), logits)
        logits = logits + sd[f"entropy_bottleneck._bias{i}"].float().cpu()
        if i < n - 1:
            logits = logits + torch.tanh(sd[f"entropy_bottleneck._factor{i}"].float().cpu()) * torch.tanh(logits)
    return logits


def entropy_bottleneck_tables(sd) -> CdfTables:
    """per-channel factorised density of z on the integer grid around the channel median"""
    q = sd["entropy_bottleneck.quantiles"].detach().float().cpu()
    medians = q[:, 0, 1]
    minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
    maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
    start = medians - minima
    length = maxima + minima + 1
    max_length = int(length.max())
    samples = torch.arange(max_length)[None, :] + start[:, None, None]
    lower = _logits_cumulative(sd, samples - 0.5)
    upper = _logits_cumulative(sd, samples + 0.5)
    pmf = (torch.sigmoid(upper) - torch.sigmoid(lower))[:, 0, :]
    tail = (torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:]))[:, 0]
    return CdfTables(_rows_to_table(pmf, tail, length, max_length), (length + 2).int(), (-minima).int())
