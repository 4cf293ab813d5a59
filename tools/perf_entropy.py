"""Micro-benchmark of the entropy-stage kernels through the C ABI profiler."""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cra5_b200 import _lib as L
from oracle import entropy_oracle as EO

t = EO.gaussian_conditional_tables()
dev = {k: getattr(t, k).cuda().contiguous() for k in ("cdf", "cdf_length", "offset")}
n_ch, Lc = 256, 10368
g = torch.Generator().manual_seed(0)
idx = torch.randint(0, 40, (n_ch * Lc,), generator=g, dtype=torch.int32)
sym = torch.round(torch.randn(n_ch * Lc, generator=g) * t.scale_table[idx.long()]).int().cuda()
idx8 = idx.to(torch.uint8).cuda()
cap = 64 + 4 * n_ch * 64 + 8 * sym.numel() + 16 * n_ch * 64
out = (ctypes.c_uint8 * cap)()
n = ctypes.c_uint64()
dec = torch.empty_like(sym)

def report():
    buf = ctypes.create_string_buffer(1 << 20); need = ctypes.c_uint64()
    L.check(L.lib.cra5_profile_report(buf, ctypes.c_uint64(len(buf)), ctypes.byref(need)))
    return json.loads(buf.value.decode())

for spc in (1, 2, 4, 8, 16, 32, 64):
    for it in range(2):
        L.check(L.lib.cra5_profile_enable(1))
        L.check(L.lib.cra5_op_rans_encode(L.ptr(sym), L.ptr(idx8), L.ptr(dev["cdf"]), dev["cdf"].shape[1], L.ptr(dev["cdf_length"]),
                                          L.ptr(dev["offset"]), n_ch, Lc, spc, out, ctypes.c_uint64(cap), ctypes.byref(n), L.stream_ptr()))
        b = bytes(out[: n.value])
        L.check(L.lib.cra5_op_rans_decode(b, ctypes.c_uint64(len(b)), L.ptr(idx8), L.ptr(dev["cdf"]), dev["cdf"].shape[1],
                                          L.ptr(dev["cdf_length"]), L.ptr(dev["offset"]), n_ch, Lc, L.ptr(dec), L.stream_ptr()))
        r = report()
    assert torch.equal(dec, sym)
    print(f"spc={spc:3d} streams={n_ch*spc:6d} bytes={n.value:8d} " + " ".join(f"{k}={v['ms']:.3f}ms" for k, v in r.items()))
