"""CPU restatement of the reference's entropy-model arithmetic (TEST INFRASTRUCTURE ONLY).

  quantize / dequantize        entropy_models.py:155-201
  GaussianConditional tables   entropy_models.py:598-643, models/base.py:54-61
  EntropyBottleneck tables     entropy_models.py:394-463
  build_indexes                entropy_models.py:679-685 (+ LowerBound forward, ops/bound_ops.py:35-36)
  pmf_to_quantized_cdf, rANS   -> oracle/rans_oracle.c through ctypes

Integer results (symbols, indexes, CDF tables, byte streams) must agree with the reference exactly; they are pinned
by tests/test_oracle_pins.py against the reference's compiled coder and by tests/golden/*.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64  # models/base.py:54-56
SCALE_BOUND = 0.11                                      # entropy_models.py:560
TAIL_MASS = 1e-9                                        # entropy_models.py:561, 351
PRECISION = 16


def build_c_oracle():
    so = os.path.join(_HERE, "_build", "liboracle.so")
    src = os.path.join(_HERE, "rans_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build_c_oracle())
        L.oracle_rans_encode.restype = ctypes.c_long
        L.oracle_rans_encode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_long]
        L.oracle_rans_decode.restype = ctypes.c_int
        L.oracle_rans_decode.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long,
                                         ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]
        L.oracle_pmf_to_quantized_cdf.restype = ctypes.c_int
        L.oracle_pmf_to_quantized_cdf.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _LIB = L
    return _LIB


# ---------------------------------------------------------------------------------------------- native wrappers
def pmf_to_quantized_cdf(pmf, precision: int = PRECISION) -> np.ndarray:
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    out = np.zeros(p.size + 1, dtype=np.uint32)
    rc = _lib().oracle_pmf_to_quantized_cdf(p.ctypes.data, p.size, precision, out.ctypes.data)
    if rc != 0:
        raise ValueError("Invalid `pmf`")
    return out.astype(np.int32)


def _i32(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def rans_encode(symbols, indexes, cdfs, cdf_sizes, offsets) -> bytes:
    s, i, c, z, o = _i32(symbols).ravel(), _i32(indexes).ravel(), _i32(cdfs), _i32(cdf_sizes), _i32(offsets)
    cap = 16 + 8 * s.size + 64
    out = np.empty(cap, dtype=np.uint8)
    n = _lib().oracle_rans_encode(s.ctypes.data, i.ctypes.data, s.size, c.ctypes.data, c.shape[1], z.ctypes.data,
                                  o.ctypes.data, out.ctypes.data, cap)
    if n < 0:
        raise RuntimeError(f"oracle_rans_encode failed ({n})")
    return out[:n].tobytes()


def rans_decode(stream: bytes, indexes, cdfs, cdf_sizes, offsets) -> torch.Tensor:
    i, c, z, o = _i32(indexes).ravel(), _i32(cdfs), _i32(cdf_sizes), _i32(offsets)
    buf = np.frombuffer(stream + b"\0" * 16, dtype=np.uint8).copy()
    out = np.empty(i.size, dtype=np.int32)
    rc = _lib().oracle_rans_decode(buf.ctypes.data, len(stream), i.ctypes.data, i.size, c.ctypes.data, c.shape[1],
                                   z.ctypes.data, o.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle_rans_decode failed")
    return torch.from_numpy(out)


# ---------------------------------------------------------------------------------------------- quantisation
def quantize_symbols(x: torch.Tensor, means: torch.Tensor) -> torch.Tensor:
    """EntropyModel.quantize(mode='symbols'), entropy_models.py:167-184: fp32 subtract, round half to even, to int32"""
    return torch.round(x - means).int()


def dequantize(symbols: torch.Tensor, means: torch.Tensor) -> torch.Tensor:
    """EntropyModel.dequantize, entropy_models.py:193-201"""
    return symbols.to(means.dtype) + means


def eb_indexes(size) -> torch.Tensor:
    """EntropyBottleneck._build_indexes, entropy_models.py:513-523: index == channel"""
    N, C = size[0], size[1]
    v = [1] * len(size)
    v[1] = -1
    return torch.arange(C, dtype=torch.int32).view(*v).repeat(N, 1, *size[2:])


def get_scale_table() -> torch.Tensor:
    """models/base.py:59-61"""
    return torch.exp(torch.linspace(math.log(SCALES_MIN), math.log(SCALES_MAX), SCALES_LEVELS))


def build_indexes(scales: torch.Tensor, scale_table: torch.Tensor) -> torch.Tensor:
    """GaussianConditional.build_indexes, entropy_models.py:679-685"""
    s = torch.max(scales, torch.tensor(SCALE_BOUND, dtype=scales.dtype))  # LowerBound fwd, bound_ops.py:35-36
    idx = torch.full(s.shape, len(scale_table) - 1, dtype=torch.int32)
    for t in scale_table[:-1]:
        idx -= (s <= t).int()
    return idx


# ---------------------------------------------------------------------------------------------- CDF tables
@dataclass
class Tables:
    cdf: torch.Tensor          # (rows, max_len + 2) int32
    cdf_length: torch.Tensor   # (rows,) int32
    offset: torch.Tensor       # (rows,) int32
    scale_table: torch.Tensor = None

    def coder_args(self):
        return self.cdf, self.cdf_length, self.offset


def _pmf_to_cdf(pmf, tail_mass, pmf_length, max_length):
    """EntropyModel._pmf_to_cdf, entropy_models.py:208-216"""
    cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
    for i, p in enumerate(pmf):
        prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
        c = torch.from_numpy(pmf_to_quantized_cdf(prob.numpy(), PRECISION))
        cdf[i, : c.numel()] = c
    return cdf


def _std_cumulative(x):
    """GaussianConditional._standardized_cumulative, entropy_models.py:598-602"""
    return 0.5 * torch.erfc(float(-(2 ** -0.5)) * x)


def gaussian_conditional_tables(scale_table: torch.Tensor = None) -> Tables:
    """GaussianConditional.update, entropy_models.py:619-643"""
    import scipy.stats
    if scale_table is None:
        scale_table = get_scale_table()
    multiplier = -scipy.stats.norm.ppf(TAIL_MASS / 2)
    pmf_center = torch.ceil(scale_table * multiplier).int()
    pmf_length = 2 * pmf_center + 1
    max_length = torch.max(pmf_length).item()
    samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
    scale = scale_table.unsqueeze(1).float()
    upper = _std_cumulative((0.5 - samples) / scale)
    lower = _std_cumulative((-0.5 - samples) / scale)
    pmf = upper - lower
    tail_mass = 2 * lower[:, :1]
    cdf = _pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
    return Tables(cdf, (pmf_length + 2).int(), (-pmf_center).int(), scale_table)


def _logits_cumulative(sd, x):
    """EntropyBottleneck._logits_cumulative, entropy_models.py:434-453"""
    logits = x
    for i in range(5):
        logits = torch.matmul(F.softplus(sd[f"entropy_bottleneck._matrix{i}"]), logits)
        logits = logits + sd[f"entropy_bottleneck._bias{i}"]
        if i < 4:
            logits = logits + torch.tanh(sd[f"entropy_bottleneck._factor{i}"]) * torch.tanh(logits)
    return logits


def entropy_bottleneck_tables(sd) -> Tables:
    """EntropyBottleneck.update, entropy_models.py:394-427"""
    q = sd["entropy_bottleneck.quantiles"].float()
    medians = q[:, 0, 1]
    minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
    maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
    offset = -minima
    pmf_start = medians - minima
    pmf_length = maxima + minima + 1
    max_length = pmf_length.max().item()
    samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
    lower = _logits_cumulative(sd, samples - 0.5)
    upper = _logits_cumulative(sd, samples + 0.5)
    pmf = (torch.sigmoid(upper) - torch.sigmoid(lower))[:, 0, :]
    tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
    cdf = _pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
    return Tables(cdf, (pmf_length + 2).int(), offset.int())


# ---------------------------------------------------------------------------------------------- likelihoods (forward)
LIKELIHOOD_BOUND = 1e-9  # entropy_models.py:111


def gc_likelihood(y_hat: torch.Tensor, scales: torch.Tensor, means: torch.Tensor) -> torch.Tensor:
    """GaussianConditional._likelihood + lower bound, entropy_models.py:645-677 (eval mode: inputs already dequantised)"""
    values = torch.abs(y_hat - means)
    s = torch.max(scales, torch.tensor(SCALE_BOUND, dtype=scales.dtype))
    upper = _std_cumulative((0.5 - values) / s)
    lower = _std_cumulative((-0.5 - values) / s)
    return torch.max(upper - lower, torch.tensor(LIKELIHOOD_BOUND))


def eb_likelihood(sd, z_hat: torch.Tensor) -> torch.Tensor:
    """EntropyBottleneck.forward likelihood path, entropy_models.py:456-510: per-channel factorised density"""
    B, C = z_hat.shape[:2]
    v = z_hat.permute(1, 0, 2, 3).reshape(C, 1, -1)
    lower = _logits_cumulative(sd, v - 0.5)
    upper = _logits_cumulative(sd, v + 0.5)
    lik = torch.sigmoid(upper) - torch.sigmoid(lower)
    lik = torch.max(lik, torch.tensor(LIKELIHOOD_BOUND))
    return lik.reshape(C, B, *z_hat.shape[2:]).permute(1, 0, 2, 3).contiguous()
