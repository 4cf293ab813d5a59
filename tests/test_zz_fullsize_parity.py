"""GPU parity at BASELINE.json's FULL size: the shipped 268-variable, 1024-wide, 25-block VAEformer on one
268x721x1440 frame against the fp32 CPU oracle (about a minute and a half of oracle time on the host cores; the file
name sorts last so the quick suites run first).

Bars: latent within bf16-tensor-core tolerance of the fp32 oracle (relative rms <= 1.5e-2; tools/emulate_bf16.py predicts
4.4e-3 for bf16 operands with fp32 accumulation), integer work bit-exact on the GPU's own floats, rANS round trip
bit-exact, reconstruction given the same latent within 1.5e-2, and the north-star gate: per-variable RMSE within 1e-4 of
the reference's RMSE (prediction 9e-6).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from cra5_b200 import config as C
from oracle import entropy_oracle as EO, vaeformer_oracle as VO


def rel_rms(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()


def test_full_size_268_round_trip_against_fp32_oracle():
    from cra5_b200.vaeformer import VAEformer, init_state_dict
    cfg = C.cra5_268()
    sd = init_state_dict(cfg, 3)
    sd["quant_conv.weight"] = sd["quant_conv.weight"] * 6.0      # the bench's entropy regime (about 5 MB per frame)
    sd["h_s.final.weight"] = sd["h_s.final.weight"] * 12.0
    net = VAEformer(268, cfg=cfg, init_seed=None)
    net.load_state_dict(sd)
    net.update(force=True)
    codec = VO.OracleCodec(sd, cfg)
    x = torch.randn(1, cfg.in_chans, *cfg.img_size, generator=torch.Generator().manual_seed(1000))
    shape = (1, cfg.latent_chans, *cfg.grid)
    with torch.no_grad():
        # ---- GPU path
        y_g, _, _ = net.encode_latent(x.cuda(), type="float")
        out = net.compress_from_latent(y_g)
        mu_g = net.tap("means").reshape(shape).cpu()
        sc_g = net.tap("scales").reshape(shape).cpu()
        ysym_g, yidx_g = net.tap("y_symbols").cpu(), net.tap("y_indexes").cpu()
        y_hat_g = net.decompress(out["strings"], out["z_shape"], return_format="latent")
        ysym_dec = net.tap("y_symbols").cpu()
        x_g = net.decode_latent(y_hat_g).cpu()
        y_g = y_g.cpu()
        # ---- fp32 oracle
        y_o = VO.encode_y(codec.sd, cfg, x)
        dbg = codec.compress_from_latent(y_o)["debug"]
        y_hat_o = dbg["y_symbols"].float() + dbg["means"]        # == decompress(compress(.)): the coder is lossless
        x_o = VO.decode_y(codec.sd, cfg, y_hat_o)
        x_o_given_g = VO.decode_y(codec.sd, cfg, y_hat_g.cpu())
    # floating point, bf16 operands / fp32 accumulation against fp32
    assert rel_rms(y_g, y_o) <= 1.5e-2
    assert rel_rms(x_g, x_o_given_g) <= 1.5e-2
    # integer work: bit-exact given the GPU's own float tensors; coder round trip bit-exact
    assert torch.equal(EO.quantize_symbols(y_g, mu_g).reshape(-1), ysym_g)
    assert torch.equal(EO.build_indexes(sc_g, codec.gc.scale_table).reshape(-1).to(torch.uint8), yidx_g)
    assert torch.equal(ysym_dec, ysym_g)
    assert torch.equal(y_hat_g.cpu(), ysym_g.reshape(shape).float() + mu_g)
    # north star: per-variable RMSE (against the input frame) within 1e-4 of the reference path's RMSE
    rm_g = ((x_g[0] - x[0]) ** 2).mean(dim=(1, 2)).sqrt()
    rm_o = ((x_o[0] - x[0]) ** 2).mean(dim=(1, 2)).sqrt()
    assert (rm_g - rm_o).abs().max().item() <= 1e-4
    # ---- what the bf16 operands cost in SYMBOLS against the fp32 reference, per precision level (reported; the tests
    #      that bound it run on the reduced-width fixtures, tests/test_gpu_precision.py)
    sym_o = dbg["y_symbols"].reshape(-1)
    flips = {0: (ysym_g != sym_o).float().mean().item()}
    direct = {0: ((x_g[0] - x_o[0]) ** 2).mean(dim=(1, 2)).sqrt().max().item()}
    for level in (1, 2):
        net.set_precision(level)
        with torch.no_grad():
            o2 = net.compress(x.cuda())
            flips[level] = (net.tap("y_symbols").cpu() != sym_o).float().mean().item()
            x2 = net.decompress(o2["strings"], o2["z_shape"])["x_hat"].cpu()
        direct[level] = ((x2[0] - x_o[0]) ** 2).mean(dim=(1, 2)).sqrt().max().item()
        assert torch.equal(net.tap("y_symbols").cpu().float().reshape(shape) + net.tap("means").reshape(shape).cpu(),
                           net.decompress(o2["strings"], o2["z_shape"], return_format="latent").cpu())
    net.set_precision(0)
    print("\n[full size 268] symbol flips vs the fp32 reference: " +
          ", ".join(f"level {k}: {100 * v:.3f} %" for k, v in flips.items()) +
          " | direct max RMSE(x_hat - x_hat_ref): " + ", ".join(f"level {k}: {v:.2e}" for k, v in direct.items()))
    assert flips[2] < 0.6 * flips[0] and flips[1] <= flips[0]   # (what remains at level 2 is the bf16 attention)
    # ---- codec lanes at full size: bit-identical to the single-lane result (regression for the attention kernel's
    #      absent-tile phase tracking, csrc/attn_tc4.cu)
    import hashlib
    from cra5_b200.stream import CodecLanes

    def roundtrip(codec, i):
        with torch.no_grad():
            o = codec.compress(x.cuda())
            rec = codec.decompress(o["strings"], o["z_shape"])["x_hat"]
        return hashlib.sha256(o["strings"][0][0] + o["strings"][1][0] + rec.cpu().numpy().tobytes()).hexdigest()

    base = roundtrip(net, 0)
    assert CodecLanes(net, lanes=2).run(roundtrip, 6) == [base] * 6
    assert 1e6 < len(out["strings"][0][0]) < 12e6


def test_full_size_batch_of_4_equals_frame_by_frame():
    """batch invariance at BASELINE.json's size: 4 x 268 x 721 x 1440 frames in one call (one launch per kernel; the GEMMs
    see 41 472 rows) reproduce four single-frame calls bit for bit -- latents, every container, the decoded latents and
    the reconstructions. (The reduced-width fixtures in test_gpu_batch.py never reach the tile counts of this size.)"""
    from cra5_b200.synthetic import bench_regime
    from cra5_b200.vaeformer import VAEformer, init_state_dict
    cfg = C.cra5_268()
    sd = bench_regime(init_state_dict(cfg, 5), cfg)
    x = torch.randn(4, cfg.in_chans, *cfg.img_size, device="cuda", generator=torch.Generator(device="cuda").manual_seed(11))
    outs = []
    for mb in (1, 4):
        net = VAEformer(268, cfg=cfg, init_seed=None, max_batch=mb)
        net.load_state_dict(sd)
        net.update(force=True)
        with torch.no_grad():
            y, _, _ = net.encode_latent(x, type="float")
            o = net.compress_from_latent(y)
            y_hat = net.decompress(o["strings"], o["z_shape"], return_format="latent")
            x_hat = net.decode_latent(y_hat)
        outs.append((y.cpu(), list(o["strings"][0]), list(o["strings"][1]), y_hat.cpu(), x_hat.cpu()))
        del net
        torch.cuda.empty_cache()
    (y1, sy1, sz1, yh1, xh1), (y4, sy4, sz4, yh4, xh4) = outs
    assert torch.equal(y4, y1) and sy4 == sy1 and sz4 == sz1 and torch.equal(yh4, yh1) and torch.equal(xh4, xh1)
    # the bench's entropy regime (cra5_b200/synthetic.py): about 2 MB per frame, like real CRA5 frames
    per_frame = [len(a) + len(b) for a, b in zip(sy1, sz1)]
    print(f"\n[full size 268, bench regime] bytes per frame {per_frame}")
    assert all(1.2e6 < n < 3.5e6 for n in per_frame)


@pytest.mark.parametrize("chans", [159, 69])
def test_full_width_other_channel_counts_against_fp32_oracle(chans):
    """SURVEY 8f-4 / BASELINE.json configs[1] and [4]: the same architecture at 159 variables
    (config/vaeformer_era5_159v_1h.py:41-50) and at 69, full width and depth, one frame against the fp32 oracle: latent and
    reconstruction within bf16 tolerance, integer stage bit-exact on the GPU's own floats, coder round trip bit-exact,
    per-variable RMSE within 1e-4 of the reference path's."""
    from cra5_b200.synthetic import bench_regime
    from cra5_b200.vaeformer import init_state_dict
    from cra5_b200.zoo import vaeformer_pretrained
    cfg = C.variant(chans)
    sd = bench_regime(init_state_dict(cfg, 4), cfg)
    net = vaeformer_pretrained(quality=chans, pretrained=False, init_seed=None)     # the zoo entry itself
    assert net.cfg.in_chans == chans
    net.load_state_dict(sd)
    net.update(force=True)
    codec = VO.OracleCodec(sd, cfg)
    x = torch.randn(1, chans, *cfg.img_size, generator=torch.Generator().manual_seed(2000 + chans))
    shape = (1, cfg.latent_chans, *cfg.grid)
    with torch.no_grad():
        y_g, _, _ = net.encode_latent(x.cuda(), type="float")
        out = net.compress_from_latent(y_g)
        mu_g, sc_g = net.tap("means").reshape(shape).cpu(), net.tap("scales").reshape(shape).cpu()
        ysym_g, yidx_g = net.tap("y_symbols").cpu(), net.tap("y_indexes").cpu()
        y_hat_g = net.decompress(out["strings"], out["z_shape"], return_format="latent")
        assert torch.equal(net.tap("y_symbols").cpu(), ysym_g)
        x_g = net.decode_latent(y_hat_g).cpu()
        y_o = VO.encode_y(codec.sd, cfg, x)
        dbg = codec.compress_from_latent(y_o)["debug"]
        x_o = VO.decode_y(codec.sd, cfg, dbg["y_symbols"].float() + dbg["means"])
        x_o_given_g = VO.decode_y(codec.sd, cfg, y_hat_g.cpu())
    assert rel_rms(y_g, y_o) <= 1.5e-2 and rel_rms(x_g, x_o_given_g) <= 1.5e-2
    assert torch.equal(EO.quantize_symbols(y_g.cpu(), mu_g).reshape(-1), ysym_g)
    assert torch.equal(EO.build_indexes(sc_g, codec.gc.scale_table).reshape(-1).to(torch.uint8), yidx_g)
    rm_g = ((x_g[0] - x[0]) ** 2).mean(dim=(1, 2)).sqrt()
    rm_o = ((x_o[0] - x[0]) ** 2).mean(dim=(1, 2)).sqrt()
    print(f"\n[full width, {chans} variables] latent rel-rms {rel_rms(y_g, y_o):.2e}, max |dRMSE| {(rm_g - rm_o).abs().max():.2e}, "
          f"{len(out['strings'][0][0]) + len(out['strings'][1][0])} bytes")
    assert (rm_g - rm_o).abs().max().item() <= 1e-4
