"""GPU parity of the whole hot path against the CPU oracle (oracle/ is pinned bit-exactly to the real reference by
tests/test_oracle_pins.py and tools/make_golden.py), on seeded weights and frames.

Bars (north star): integer work -- symbols, scale indexes, every rANS sub-stream, decoded symbols -- bit-exact;
floating-point transforms within bf16-tensor-core tolerance of the fp32 oracle, stated per check; per-variable RMSE of
the reconstruction within 1e-4 of the reference's RMSE at full resolution.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cra5_b200 import config as C
from oracle import entropy_oracle as EO, vaeformer_oracle as VO, weights
from tests import cr5b

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"small": (C.small_lowres(5), 11, 3), "tiny69": (C.tiny_fullres(69), 7, 1)}


class Ctx:
    pass


@pytest.fixture(scope="module", params=["small", "tiny69"])
def ctx(request):
    from cra5_b200.vaeformer import VAEformer
    name = request.param
    cfg, wseed, fseed = CASES[name]
    c = Ctx()
    c.name, c.cfg = name, cfg
    c.sd = weights.seeded_state_dict(C.param_shapes(cfg), wseed)
    c.x = weights.seeded_frame(cfg, fseed).unsqueeze(0)
    c.codec = VO.OracleCodec(c.sd, cfg)
    c.net = VAEformer(268, cfg=cfg, init_seed=None)
    c.net.load_state_dict(c.sd)
    with pytest.raises(ValueError, match="Uninitialized CDFs"):  # entropy_models.py:218-220
        c.net.compress(c.x.cuda())
    assert c.net.update(force=True) is True
    assert c.net.update() is False
    c.gold = np.load(os.path.join(GOLD, f"{name}.npz"), allow_pickle=False)
    with torch.no_grad():
        c.y_o = VO.encode_y(c.codec.sd, cfg, c.x)
        c.y_g, none1, none2 = c.net.encode_latent(c.x.cuda(), type="float")
        assert none1 is None and none2 is None
        c.out = c.net.compress_from_latent(c.y_g)
        c.z_g = c.net.tap("z").reshape(1, cfg.z_chans, *cfg.hyper_grid).cpu()
        c.zhat_g = c.net.tap("z_hat").reshape(1, cfg.z_chans, *cfg.hyper_grid).cpu()
        c.zsym_g = c.net.tap("z_symbols").cpu()
        c.sc_g = c.net.tap("scales").reshape(1, cfg.latent_chans, *cfg.grid).cpu()
        c.mu_g = c.net.tap("means").reshape(1, cfg.latent_chans, *cfg.grid).cpu()
        c.ysym_g = c.net.tap("y_symbols").cpu()
        c.yidx_g = c.net.tap("y_indexes").cpu()
    return c


def rel_rms(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()


def test_cdf_tables_match_reference_fixture(ctx):
    # integer tables built by the product's update() == the reference's (golden) == oracle
    sd = ctx.net.state_dict()
    for tag, mod, tab in (("gc", "gaussian_conditional", ctx.codec.gc), ("eb", "entropy_bottleneck", ctx.codec.eb)):
        assert torch.equal(sd[f"{mod}._quantized_cdf"], tab.cdf)
        assert sd[f"{mod}._cdf_length"].tolist() == ctx.gold[f"{tag}_cdf_length"].tolist()
        assert sd[f"{mod}._offset"].tolist() == ctx.gold[f"{tag}_offset"].tolist()
    assert np.array_equal(sd["entropy_bottleneck._quantized_cdf"].numpy(), ctx.gold["eb_cdf"])
    assert np.array_equal(sd["gaussian_conditional.scale_table"].numpy(), ctx.gold["gc_scale_table"])


def test_g_a_latent_close_to_oracle(ctx):
    # bf16 tensor-core operands with fp32 accumulation vs the fp32 oracle: relative rms error <= 1.5 %
    assert tuple(ctx.y_g.shape) == (1, ctx.cfg.latent_chans, *ctx.cfg.grid)
    assert rel_rms(ctx.y_g, ctx.y_o) <= 1.5e-2
    assert (ctx.y_g.cpu() - ctx.y_o).abs().max().item() <= 0.02 * ctx.y_o.abs().max().item() + 0.05


def test_hyper_transforms_close_to_oracle(ctx):
    with torch.no_grad():
        z_o = VO.h_a(ctx.codec.sd, ctx.cfg, ctx.y_g.cpu())
        sc_o, mu_o = VO.h_s(ctx.codec.sd, ctx.cfg, ctx.zhat_g)
    assert rel_rms(ctx.z_g, z_o) <= 1.5e-2
    assert rel_rms(ctx.sc_g, sc_o) <= 1.5e-2
    assert rel_rms(ctx.mu_g, mu_o) <= 1.5e-2


def test_quantisation_and_indexes_bit_exact(ctx):
    """integer results from the GPU's own float tensors must equal the reference arithmetic exactly
    (entropy_models.py:167-184 round-half-even, :679-685 bucketisation, :390/:529-535 medians)"""
    med = ctx.codec.sd["entropy_bottleneck.quantiles"][:, 0, 1].reshape(1, -1, 1, 1)
    assert torch.equal(EO.quantize_symbols(ctx.z_g, med).reshape(-1), ctx.zsym_g)
    assert torch.equal(EO.quantize_symbols(ctx.z_g, med).float() + med, ctx.zhat_g)
    assert torch.equal(EO.quantize_symbols(ctx.y_g.cpu(), ctx.mu_g).reshape(-1), ctx.ysym_g)
    idx_o = EO.build_indexes(ctx.sc_g, ctx.codec.gc.scale_table)
    assert torch.equal(idx_o.reshape(-1).to(torch.uint8), ctx.yidx_g)
    assert len(torch.unique(ctx.yidx_g)) >= 20  # the fixture exercises many rows of the scale table


def test_every_substream_equals_reference_coder(ctx):
    """CR5B sub-stream (c, k) is byte-identical to the reference coder run on symbols[c, k::spc]"""
    cfg = ctx.cfg
    for which, sym, idx, tab, n_ch in (
            (0, ctx.ysym_g, ctx.yidx_g.int(), ctx.codec.gc, cfg.latent_chans),
            (1, ctx.zsym_g, EO.eb_indexes((1, cfg.z_chans, *cfg.hyper_grid)).reshape(-1), ctx.codec.eb, cfg.z_chans)):
        cont = cr5b.parse(ctx.out["strings"][which][0])
        assert cont["n_channels"] == n_ch and cont["L"] * n_ch == sym.numel()
        L, spc = cont["L"], cont["spc"]
        sym2, idx2 = sym.reshape(n_ch, L), idx.reshape(n_ch, L)
        channels = range(n_ch) if n_ch * spc <= 64 else list(range(0, n_ch, max(1, n_ch // 6)))[:6] + [n_ch - 1]
        for c in channels:
            for k in range(spc):
                ref = EO.rans_encode(sym2[c, k::spc], idx2[c, k::spc], *tab.coder_args())
                assert cont["streams"][c * spc + k] == ref, (which, c, k)
    assert tuple(ctx.out["z_shape"]) == tuple(cfg.hyper_grid)


def test_decompress_round_trip_bit_exact(ctx):
    with torch.no_grad():
        y_hat = ctx.net.decompress(ctx.out["strings"], ctx.out["z_shape"], return_format="latent")
    assert torch.equal(ctx.net.tap("y_symbols").cpu(), ctx.ysym_g)
    assert torch.equal(ctx.net.tap("z_symbols").cpu(), ctx.zsym_g)
    assert torch.equal(ctx.net.tap("y_indexes").cpu(), ctx.yidx_g)  # h_s is deterministic encode- vs decode-side
    expect = ctx.ysym_g.reshape(ctx.mu_g.shape).float() + ctx.mu_g   # EntropyModel.dequantize, entropy_models.py:193-201
    assert torch.equal(y_hat.cpu(), expect)
    # encode_latent(type='quantized') tail == coded path (reference invariant)
    with torch.no_grad():
        _, y_hat_q, _ = ctx.net.encode_latent(ctx.x.cuda(), type="quantized")
    assert torch.equal(y_hat_q.cpu(), expect)
    ctx.y_hat = y_hat


def test_g_s_reconstruction_close_to_oracle(ctx):
    with torch.no_grad():
        y_hat = ctx.net.decompress(ctx.out["strings"], ctx.out["z_shape"], return_format="latent")
        x_g = ctx.net.decode_latent(y_hat)
        x_o = VO.decode_y(ctx.codec.sd, ctx.cfg, y_hat.cpu())
    assert tuple(x_g.shape) == (1, ctx.cfg.in_chans, *ctx.cfg.img_size)
    assert rel_rms(x_g, x_o) <= 1.5e-2
    assert torch.isfinite(x_g).all()


def test_per_variable_rmse_within_tolerance_of_reference(ctx):
    """north-star gate: |RMSE_new(c) - RMSE_ref(c)| <= 1e-4 (normalised units, RMSE against the input), no waiver.
    The full-resolution fixture meets it with plain bf16 operands; the low-resolution one has 20x fewer pixels per
    variable, so individual symbol flips show (1.1e-4 at bf16) -- it meets the same 1e-4 at precision level "encoder"
    (split-bf16 GEMMs on every layer the symbols depend on), which is what the gate is asserted on there."""
    level = 0 if ctx.name == "tiny69" else 2
    ctx.net.set_precision(level)
    try:
        with torch.no_grad():
            out = ctx.net.compress(ctx.x.cuda())
            x_g = ctx.net.decompress(out["strings"], out["z_shape"])["x_hat"].cpu()
    finally:
        ctx.net.set_precision(0)
    rmse_g = ((x_g[0] - ctx.x[0]) ** 2).mean(dim=(1, 2)).sqrt().numpy()
    rmse_ref = ctx.gold["rmse_per_var"]  # produced by the real reference
    print(f"\n[{ctx.name}] precision level {level}: max |dRMSE| vs the reference {np.abs(rmse_g - rmse_ref).max():.2e}")
    assert np.abs(rmse_g - rmse_ref).max() <= 1e-4, np.abs(rmse_g - rmse_ref).max()
    # and the reconstruction itself is close to the reference's (symbol flips near .5 allowed): rms diff <= 10 %
    assert np.abs(x_g[0].std(dim=(1, 2)).numpy() - ctx.gold["xhat_std_per_var"]).max() <= 5e-3
    # rate within 1 % + container overhead of the reference's single-stream coder
    ref_bytes = len(ctx.gold["y_string"]) + len(ctx.gold["z_string"])
    new_bytes = len(out["strings"][0][0]) + len(out["strings"][1][0])
    n_streams = ctx.cfg.latent_chans * 16 + ctx.cfg.z_chans * 4
    assert new_bytes <= 1.01 * ref_bytes + 12 * n_streams + 64


def test_errors_mirror_reference(ctx):
    net = ctx.net
    with pytest.raises(ValueError):
        net.encode_latent(torch.zeros(1, 3, 8, 8))
    with pytest.raises(ValueError):
        net.decompress([[b"nonsense"], [b"nonsense"]], ctx.out["z_shape"], return_format="latent")
    bad = bytearray(ctx.out["strings"][0][0])
    with pytest.raises((ValueError, RuntimeError)):
        net.decompress([[bytes(bad[:-4])], ctx.out["strings"][1]], ctx.out["z_shape"], return_format="latent")
    with pytest.raises(ValueError):
        net.decompress(ctx.out["strings"], (1, 1), return_format="latent")
    # model still usable afterwards
    y_hat = net.decompress(ctx.out["strings"], ctx.out["z_shape"], return_format="latent")
    assert torch.isfinite(y_hat).all()


def test_coder_stream_count_knob(ctx):
    net = ctx.net
    sizes = {}
    for spc in (1, 4, 32):
        net.set_coder(spc, 1)
        out = net.compress_from_latent(ctx.y_g)
        assert cr5b.parse(out["strings"][0][0])["spc"] == spc
        y_hat = net.decompress(out["strings"], out["z_shape"], return_format="latent")
        assert torch.equal(net.tap("y_symbols").cpu(), ctx.ysym_g)
        sizes[spc] = len(out["strings"][0][0])
    net.set_coder(16, 4)
    assert sizes[1] < sizes[4] < sizes[32]
    with pytest.raises(ValueError):
        net.set_coder(65, 1)
    with pytest.raises(ValueError):
        net.set_coder(4, -1)
    with pytest.raises(ValueError):
        net.set_coder(format="zip")


def test_reference_coder_format_end_to_end(ctx):
    """format="ref": strings are the reference's own single-stream format -- byte-identical to the reference coder fed
    the GPU's symbols -- and a stream written by the reference coder decodes to the same latent"""
    net, cfg = ctx.net, ctx.cfg
    net.set_coder(format="ref")
    try:
        out = net.compress_from_latent(ctx.y_g)
        ysym, yidx, zsym = net.tap("y_symbols").cpu(), net.tap("y_indexes").cpu().int(), net.tap("z_symbols").cpu()
        assert torch.equal(ysym, ctx.ysym_g) and torch.equal(zsym, ctx.zsym_g)
        zidx = EO.eb_indexes((1, cfg.z_chans, *cfg.hyper_grid)).reshape(-1)
        assert out["strings"][0][0] == EO.rans_encode(ysym, yidx, *ctx.codec.gc.coder_args())
        assert out["strings"][1][0] == EO.rans_encode(zsym, zidx, *ctx.codec.eb.coder_args())
        y_hat = net.decompress(out["strings"], out["z_shape"], return_format="latent")
        assert torch.equal(y_hat.cpu(), ysym.reshape(ctx.mu_g.shape).float() + ctx.mu_g)
    finally:
        net.set_coder(16, 4)
    # auto-detection: the chunked container still decodes after the switch back
    y_hat2 = net.decompress(ctx.out["strings"], ctx.out["z_shape"], return_format="latent")
    assert torch.equal(y_hat2.cpu(), y_hat.cpu())


def test_reference_written_archive_is_flagged_not_trusted(ctx):
    """strings written by the fp32 reference path (here: the oracle codec, byte-identical to the real
    reference, tools/make_golden.py). The z stream depends on integer tables and the channel index only, so it decodes
    exactly. The y stream needs the writer's h_s floats bit for bit; this library's h_s is a different implementation,
    so the y latent is NOT trusted: decompress warns, and either reports a corrupt stream or returns symbols that may
    differ from the writer's -- never silently claims success."""
    from cra5_b200.vaeformer import VAEformer
    with torch.no_grad():
        o = ctx.codec.compress(ctx.x)
    VAEformer._warned_ref_stream = False
    with pytest.warns(RuntimeWarning, match="reference-format"):
        try:
            ctx.net.decompress(o["strings"], o["z_shape"], return_format="latent")
        except ValueError:
            pass                                  # stream exhausted: reported as a corrupt bitstream
    assert torch.equal(ctx.net.tap("z_symbols").cpu(), o["debug"]["z_symbols"].reshape(-1).int())
    # a second call does not repeat the warning (once per process)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        try:
            ctx.net.decompress(o["strings"], o["z_shape"], return_format="latent")
        except ValueError:
            pass


def test_forward_likelihoods_match_reference_arithmetic(ctx):
    """rate-estimation path (VAEformer.forward, vaeformer.py:302-333): likelihood tensors against the oracle's fp32
    erfc / logistic arithmetic on the GPU's own (y, sigma, mu, z_hat); total bits against the reference's own forward()"""
    net, cfg = ctx.net, ctx.cfg
    with torch.no_grad():
        out = net.forward(ctx.x.cuda())
    lik_y, lik_z = out["likelihoods"]["y"].cpu(), out["likelihoods"]["z"].cpu()
    assert tuple(lik_y.shape) == (1, cfg.latent_chans, *cfg.grid) and tuple(lik_z.shape) == (1, cfg.z_chans, *cfg.hyper_grid)
    y_hat = ctx.ysym_g.reshape(ctx.mu_g.shape).float() + ctx.mu_g
    ref_y = EO.gc_likelihood(y_hat, ctx.sc_g, ctx.mu_g)
    ref_z = EO.eb_likelihood(ctx.codec.sd, ctx.zhat_g)
    # fp32 erfc / tanh / exp of two libraries: elementwise within 1e-6 absolute + 2e-4 relative
    assert ((lik_y - ref_y).abs() <= 1e-6 + 2e-4 * ref_y).all()
    assert ((lik_z - ref_z).abs() <= 1e-6 + 2e-4 * ref_z).all()
    bits_y = float(-torch.log2(lik_y.double()).sum())
    bits_z = float(-torch.log2(lik_z.double()).sum())
    assert abs(bits_y - float(-torch.log2(ref_y.double()).sum())) <= 1e-4 * bits_y
    assert abs(bits_z - float(-torch.log2(ref_z.double()).sum())) <= 1e-4 * bits_z
    # against the reference's own forward() on the same frame: bf16 transforms move individual symbols, the estimated
    # rate stays within 1 %
    assert abs(bits_y - float(ctx.gold["bits_y"])) <= 1e-2 * float(ctx.gold["bits_y"])
    assert abs(bits_z - float(ctx.gold["bits_z"])) <= 2e-2 * float(ctx.gold["bits_z"])
    assert torch.isfinite(out["x_hat"]).all() and out["posterior"] is None


def test_codec_lanes_are_bit_identical(ctx):
    """several codec lanes on one GPU (cra5_b200.stream.CodecLanes: own handle / stream / host thread, shared weights)
    reproduce the single-lane bitstreams and reconstruction bit for bit. (The full-size model is checked the same way in
    test_zz_fullsize_parity.py: that is where a timing-dependent race in the attention kernel showed in round 2.)"""
    import hashlib
    from cra5_b200.stream import CodecLanes

    def roundtrip(codec, i):
        with torch.no_grad():
            o = codec.compress(ctx.x.cuda())
            rec = codec.decompress(o["strings"], o["z_shape"])["x_hat"]
        return hashlib.sha256(o["strings"][0][0] + o["strings"][1][0] + rec.cpu().numpy().tobytes()).hexdigest()

    base = roundtrip(ctx.net, 0)
    got = CodecLanes(ctx.net, lanes=3).run(roundtrip, 9)
    assert got == [base] * 9


def test_likelihoods_entry_accepts_null_outputs(ctx):
    """include/cra5_b200.h: "Any output may be NULL" for cra5_latent_likelihoods. y_lik = NULL used to make the
    quantise kernel read a null scale table; y_hat alone, z_lik alone and nothing at all must work."""
    import ctypes
    from cra5_b200 import _lib
    net, cfg = ctx.net, ctx.cfg
    y = ctx.y_g.contiguous()
    y_hat = torch.empty_like(y)
    z_lik = torch.empty((cfg.z_chans, *cfg.hyper_grid), device="cuda")
    null = ctypes.c_void_p(0)
    s = _lib.stream_ptr()
    _lib.check(_lib.lib.cra5_latent_likelihoods(net._handle, _lib.ptr(y[0]), _lib.ptr(y_hat[0]), null, null, s))
    torch.cuda.synchronize()
    expect = ctx.ysym_g.reshape(ctx.mu_g.shape).float() + ctx.mu_g
    assert torch.equal(y_hat.cpu(), expect)
    _lib.check(_lib.lib.cra5_latent_likelihoods(net._handle, _lib.ptr(y[0]), null, null, _lib.ptr(z_lik), s))
    _lib.check(_lib.lib.cra5_latent_likelihoods(net._handle, _lib.ptr(y[0]), null, null, null, s))
    torch.cuda.synchronize()
    assert torch.isfinite(z_lik).all() and (z_lik > 0).all()
    # the operator entry refuses an index request without a scale table instead of faulting
    idx = torch.empty(16, dtype=torch.uint8, device="cuda")
    v = torch.zeros(16, device="cuda")
    with pytest.raises(ValueError):
        _lib.check(_lib.lib.cra5_op_gc_quantize(null, _lib.ptr(v), null, null, 64, ctypes.c_float(0.11), null,
                                                _lib.ptr(idx), null, ctypes.c_uint64(16), s))
