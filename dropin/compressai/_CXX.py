"""compressai._CXX.pmf_to_quantized_cdf (cpp_exts/ops/ops.cpp:40-118) through the C ABI (host code, no GPU needed)."""
import ctypes

from cra5_b200 import _lib


def pmf_to_quantized_cdf(pmf, precision):
    pmf = [float(v) for v in pmf]
    n = len(pmf)
    src = (ctypes.c_float * n)(*pmf)
    dst = (ctypes.c_uint32 * (n + 1))()
    _lib.check(_lib.lib.cra5_pmf_to_quantized_cdf(src, n, int(precision), dst))
    return list(dst)
