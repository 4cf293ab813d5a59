"""GPU parity of the entropy stage in isolation, on synthetic (y, sigma, mu) that reach every row of the 64-level
scale table and both bypass branches (last bullet): bit-exact against the oracle, which is pinned to
the reference's coder (tests/golden/rans_kat.json)."""
import ctypes
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import entropy_oracle as EO, weights
from tests import cr5b

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def L():
    from cra5_b200 import _lib
    return _lib


@pytest.fixture(scope="module")
def tabs():
    t = EO.gaussian_conditional_tables()
    dev = {k: getattr(t, k).cuda().contiguous() for k in ("cdf", "cdf_length", "offset")}
    dev["scale_table"] = t.scale_table.cuda().contiguous()
    return t, dev


def gpu_quantize(y, sig, mu, dev, want_hat=True):
    lib = L()
    n = y.numel()
    sym = torch.empty(n, dtype=torch.int32, device="cuda")
    idx = torch.empty(n, dtype=torch.uint8, device="cuda")
    hat = torch.empty(n, dtype=torch.float32, device="cuda") if want_hat else None
    lib.check(lib.lib.cra5_op_gc_quantize(lib.ptr(y), lib.ptr(sig), lib.ptr(mu), lib.ptr(dev["scale_table"]), 64,
                                          ctypes.c_float(0.11), lib.ptr(sym), lib.ptr(idx), lib.ptr(hat),
                                          ctypes.c_uint64(n), lib.stream_ptr()))
    torch.cuda.synchronize()
    return sym, idx, hat


def gpu_encode(sym, idx, dev, n_ch, Lc, spc, table=False):
    """table=True: the shared-memory-table kernels (row count stated), the path the model takes"""
    lib = L()
    cap = 64 + 4 * n_ch * spc + 8 * max(sym.numel(), 1) + 16 * n_ch * spc
    out = (ctypes.c_uint8 * cap)()
    n = ctypes.c_uint64()
    if table:
        lib.check(lib.lib.cra5_op_rans_encode_table(lib.ptr(sym), lib.ptr(idx), lib.ptr(dev["cdf"]), dev["cdf"].shape[0],
                                                    dev["cdf"].shape[1], lib.ptr(dev["cdf_length"]), lib.ptr(dev["offset"]),
                                                    n_ch, Lc, spc, out, ctypes.c_uint64(cap), ctypes.byref(n),
                                                    lib.stream_ptr()))
    else:
        lib.check(lib.lib.cra5_op_rans_encode(lib.ptr(sym), lib.ptr(idx), lib.ptr(dev["cdf"]), dev["cdf"].shape[1],
                                              lib.ptr(dev["cdf_length"]), lib.ptr(dev["offset"]), n_ch, Lc, spc, out,
                                              ctypes.c_uint64(cap), ctypes.byref(n), lib.stream_ptr()))
    return bytes(out[: n.value])


def gpu_decode(b, idx, dev, n_ch, Lc, table=False):
    lib = L()
    sym = torch.full((max(n_ch * Lc, 1),), -12345, dtype=torch.int32, device="cuda")
    if table:
        lib.check(lib.lib.cra5_op_rans_decode_table(b, ctypes.c_uint64(len(b)), lib.ptr(idx), lib.ptr(dev["cdf"]),
                                                    dev["cdf"].shape[0], dev["cdf"].shape[1], lib.ptr(dev["cdf_length"]),
                                                    lib.ptr(dev["offset"]), n_ch, Lc, lib.ptr(sym), lib.stream_ptr()))
    else:
        lib.check(lib.lib.cra5_op_rans_decode(b, ctypes.c_uint64(len(b)), lib.ptr(idx), lib.ptr(dev["cdf"]),
                                              dev["cdf"].shape[1], lib.ptr(dev["cdf_length"]), lib.ptr(dev["offset"]),
                                              n_ch, Lc, lib.ptr(sym), lib.stream_ptr()))
    torch.cuda.synchronize()
    return sym[: n_ch * Lc]


@pytest.mark.parametrize("case", [0, 1])
def test_quantize_index_matches_reference_fixture(tabs, case):
    """symbols / indexes / dequantised values equal the sha256 recorded from the REAL reference modules"""
    import hashlib
    t, dev = tabs
    with open(os.path.join(GOLD, "rans_kat.json")) as f:
        rec = json.load(f)["synthetic"][case]
    y, sig, mu = weights.synth_entropy_case(rec["seed"], rec["n"])
    sym, idx, hat = gpu_quantize(y.cuda(), sig.cuda(), mu.cuda(), dev)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert sha(sym.cpu().numpy()) == rec["symbols_sha"]
    assert sha(idx.cpu().int().numpy()) == rec["indexes_sha"]
    assert sha(hat.cpu().numpy()) == rec["yhat_sha"]
    assert torch.bincount(idx.long().cpu(), minlength=64).tolist() == rec["index_hist"]  # all 64 rows touched
    # one channel, one stream == the reference's single sequential stream, byte for byte
    b = gpu_encode(sym, idx, dev, 1, rec["n"], 1)
    c = cr5b.parse(b)
    assert len(c["streams"]) == 1 and len(c["streams"][0]) == rec["nbytes"]
    assert hashlib.sha256(c["streams"][0]).hexdigest() == rec["stream_sha"]
    assert torch.equal(gpu_decode(b, idx, dev, 1, rec["n"]), sym)


def test_round_half_even_and_scale_boundaries(tabs):
    t, dev = tabs
    st = t.scale_table
    # exact .5 ties (round half to even), values straddling every table entry by one ulp, negative / zero scales
    y = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 1e6 + 0.5, 3.4999998, 3.5000002], dtype=torch.float32)
    mu = torch.zeros_like(y)
    below = torch.nextafter(st, torch.zeros_like(st))
    above = torch.nextafter(st, torch.full_like(st, 1e9))
    sig = torch.cat([st, below, above, torch.tensor([-1.0, 0.0, 0.11, 0.10999, 1e9])])
    n = 256
    yy = torch.zeros(n); yy[: y.numel()] = y
    ss = torch.ones(n); ss[: sig.numel()] = sig
    mm = torch.zeros(n)
    sym, idx, _ = gpu_quantize(yy.cuda(), ss.cuda(), mm.cuda(), dev)
    assert torch.equal(sym.cpu(), EO.quantize_symbols(yy, mm))
    assert torch.equal(idx.cpu().int(), EO.build_indexes(ss, st))


@pytest.mark.parametrize("kind", ["default64", "log40", "irregular", "linear", "tiny"])
def test_scale_index_paths_equal_reference_bucketisation(kind):
    """GaussianConditional.build_indexes (entropy_models.py:679-685) for log-spaced tables (the kernel's log2 + two-compare
    path) and for tables that are not log-spaced (binary-search fallback): a dense sweep plus every table entry and its
    two float neighbours must give the reference's integer, whichever path the kernel picks."""
    import math
    lib = L()
    g = torch.Generator().manual_seed(5)
    if kind == "default64":
        st = EO.get_scale_table()
    elif kind == "log40":
        st = torch.exp(torch.linspace(math.log(0.2), math.log(100.0), 40))
    elif kind == "irregular":
        st = torch.cumsum(torch.rand(50, generator=g) * 3 + 0.01, 0) + 0.11          # uneven steps
    elif kind == "linear":
        st = torch.linspace(0.11, 64.0, 64)
    else:
        st = torch.tensor([0.5, 2.0, 9.0])                                             # fewer than 4 levels
    st = st.float()
    dense = torch.exp(torch.empty(200000).uniform_(math.log(0.01), math.log(2000.0), generator=g))
    below = torch.nextafter(st, torch.zeros_like(st))
    above = torch.nextafter(st, torch.full_like(st, 1e9))
    sig = torch.cat([dense, st, below, above, torch.tensor([-3.0, 0.0, 0.11, float("inf")])]).float()
    pad = (-sig.numel()) % 4
    sig = torch.cat([sig, torch.ones(pad)])
    n = sig.numel()
    idx = torch.empty(n, dtype=torch.uint8, device="cuda")
    tab = st.cuda().contiguous()
    lib.check(lib.lib.cra5_op_gc_quantize(ctypes.c_void_p(0), lib.ptr(sig.cuda()), ctypes.c_void_p(0), lib.ptr(tab),
                                          int(st.numel()), ctypes.c_float(0.11), ctypes.c_void_p(0), lib.ptr(idx),
                                          ctypes.c_void_p(0), ctypes.c_uint64(n), lib.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(idx.cpu().int(), EO.build_indexes(sig, st).int())


@pytest.mark.parametrize("n_ch,Lc,spc", [(1, 1, 1), (3, 7, 8), (4, 1000, 8), (256, 648, 1), (16, 10368, 8), (5, 333, 64),
                                         (2, 0, 4), (0, 5, 2)])
@pytest.mark.parametrize("table", [False, True], ids=["global-table", "smem-table"])
def test_chunked_coder_substreams_and_roundtrip(tabs, n_ch, Lc, spc, table):
    t, dev = tabs
    n = n_ch * Lc
    g = torch.Generator().manual_seed(n_ch * 1000 + Lc + spc)
    idx = torch.randint(0, 64, (max(n, 1),), generator=g, dtype=torch.int32)[:n]
    scale = t.scale_table[idx.long()] if n else torch.zeros(0)
    sym = torch.round(torch.randn(n, generator=g) * scale * 1.5).int()
    if n > 10:
        sym[::13] = (torch.randn(sym[::13].shape, generator=g) * 30000).int()  # bypass, several nibbles
    sym_d = sym.cuda() if n else torch.zeros(1, dtype=torch.int32, device="cuda")
    idx_d = idx.to(torch.uint8).cuda() if n else torch.zeros(1, dtype=torch.uint8, device="cuda")
    b = gpu_encode(sym_d, idx_d, dev, n_ch, Lc, spc, table)
    if table:  # both kernel families emit the same container
        assert b == gpu_encode(sym_d, idx_d, dev, n_ch, Lc, spc, False)
    c = cr5b.parse(b)
    assert (c["n_channels"], c["L"], c["spc"]) == (n_ch, Lc, spc)
    s2, i2 = sym.reshape(n_ch, Lc), idx.reshape(n_ch, Lc)
    for ch in range(min(n_ch, 6)):
        for k in range(spc):
            if Lc == 0:
                assert c["streams"][ch * spc + k] == b""
                continue
            ref = EO.rans_encode(s2[ch, k::spc], i2[ch, k::spc], *t.coder_args())
            assert c["streams"][ch * spc + k] == ref, (ch, k)
    out = gpu_decode(b, idx_d, dev, n_ch, Lc, table)
    assert torch.equal(out.cpu(), sym)


@pytest.mark.parametrize("table", [False, True], ids=["global-table", "smem-table"])
def test_entropy_bottleneck_mode_index_is_channel(table):
    """idx == NULL -> CDF row = channel (EntropyBottleneck._build_indexes, entropy_models.py:513-523)"""
    from cra5_b200 import config as C
    cfg = C.tiny_fullres(69)
    sd = weights.seeded_state_dict(C.param_shapes(cfg), 7)
    eb = EO.entropy_bottleneck_tables(sd)
    dev = {"cdf": eb.cdf.cuda().contiguous(), "cdf_length": eb.cdf_length.cuda().contiguous(),
           "offset": eb.offset.cuda().contiguous()}
    n_ch, Lc = cfg.z_chans, 648
    g = torch.Generator().manual_seed(5)
    sym = torch.round(torch.randn(n_ch, Lc, generator=g) * 4).int()
    sym[:, ::50] = 40  # outside every channel's support -> bypass
    b = gpu_encode(sym.cuda().reshape(-1), None, dev, n_ch, Lc, 1, table)
    c = cr5b.parse(b)
    idx = EO.eb_indexes((1, n_ch, Lc)).reshape(n_ch, Lc)
    for ch in range(n_ch):
        assert c["streams"][ch] == EO.rans_encode(sym[ch], idx[ch], *eb.coder_args())
    assert torch.equal(gpu_decode(b, None, dev, n_ch, Lc, table).cpu(), sym.reshape(-1))


def test_corrupt_streams_are_rejected(tabs):
    t, dev = tabs
    g = torch.Generator().manual_seed(9)
    idx = torch.randint(20, 50, (64,), generator=g).to(torch.uint8).cuda()
    sym = torch.round(torch.randn(64, generator=g) * 20).int().cuda()
    b = gpu_encode(sym, idx, dev, 2, 32, 2)
    for bad in (b[:10], b[:-4], b + b"\0\0\0\0", b"CR5B" + b"\x09" + b[5:]):
        with pytest.raises(ValueError):
            gpu_decode(bad, idx, dev, 2, 32)
    # without the magic the bytes are taken for a reference-format stream: garbage in, garbage (or an error) out
    try:
        assert not torch.equal(gpu_decode(b"XXXX" + b[4:], idx, dev, 2, 32), sym)
    except ValueError:
        pass
    with pytest.raises(ValueError):
        gpu_decode(b, idx, dev, 2, 33)  # shape mismatch
    assert torch.equal(gpu_decode(b, idx, dev, 2, 32), sym)
    for bad in (b[:-4], b + b"\0\0\0\0"):   # the shared-memory-table decoder rejects them the same way
        with pytest.raises(ValueError):
            gpu_decode(bad, idx, dev, 2, 32, table=True)
    assert torch.equal(gpu_decode(b, idx, dev, 2, 32, table=True), sym)


def test_reference_format_streams_interoperate(tabs):
    """spc == 0: one sequential stream per tensor, byte-identical to the reference coder (oracle pinned to it), and
    reference-written streams decode on the GPU"""
    t, dev = tabs
    g = torch.Generator().manual_seed(77)
    n_ch, Lc = 7, 501
    idx = torch.randint(0, 64, (n_ch * Lc,), generator=g, dtype=torch.int32)
    sym = torch.round(torch.randn(n_ch * Lc, generator=g) * t.scale_table[idx.long()] * 1.2).int()
    sym[::29] = (torch.randn(sym[::29].shape, generator=g) * 20000).int()
    ref_stream = EO.rans_encode(sym, idx, *t.coder_args())             # == compressai.ans output (test_oracle_pins)
    b = gpu_encode(sym.cuda(), idx.to(torch.uint8).cuda(), dev, n_ch, Lc, 0)
    assert b == ref_stream
    assert torch.equal(gpu_decode(ref_stream, idx.to(torch.uint8).cuda(), dev, n_ch, Lc).cpu(), sym)
