"""Batch axis through the C ABI (missing item 1; BASELINE.json configs[4]): a (B, C, H, W) input runs
as ONE launch of every kernel -- B x tokens rows through each GEMM / LayerNorm / attention, B x channels sub-streams
through the entropy kernels -- and must reproduce frame-by-frame calls BIT FOR BIT (the hyperprior
branch has to be batch-size invariant or encode- and decode-side indexes diverge). The reference API is batched the
same way: vaeformer.py:350-400 take (B, C, H, W); entropy_models.py:263-272 loops the batch items."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from cra5_b200 import config as C
from oracle import weights

CASES = {"small": (C.small_lowres(5), 11), "tiny69": (C.tiny_fullres(69), 7)}


@pytest.fixture(scope="module", params=["small", "tiny69"])
def nets(request):
    from cra5_b200.vaeformer import VAEformer
    cfg, wseed = CASES[request.param]
    sd = weights.seeded_state_dict(C.param_shapes(cfg), wseed)
    one = VAEformer(268, cfg=cfg, init_seed=None)
    one.load_state_dict(sd)
    one.update(force=True)
    many = VAEformer(268, cfg=cfg, init_seed=None, max_batch=3)
    many.load_state_dict(sd)
    many.update(force=True)
    x = torch.stack([weights.seeded_frame(cfg, s) for s in (1, 2, 3, 4, 5)]).cuda()
    return cfg, one, many, x


def _per_frame(net, x):
    ys, strings_y, strings_z, yh, xh = [], [], [], [], []
    with torch.no_grad():
        for b in range(x.shape[0]):
            y, _, _ = net.encode_latent(x[b:b + 1], type="float")
            o = net.compress_from_latent(y)
            y_hat = net.decompress(o["strings"], o["z_shape"], return_format="latent")
            ys.append(y); strings_y.append(o["strings"][0][0]); strings_z.append(o["strings"][1][0])
            yh.append(y_hat); xh.append(net.decode_latent(y_hat))
    return torch.cat(ys), strings_y, strings_z, torch.cat(yh), torch.cat(xh)


@pytest.mark.parametrize("B", [1, 2, 3, 5])
def test_batched_calls_equal_frame_by_frame_calls(nets, B):
    cfg, one, many, x = nets
    xb = x[:B].contiguous()
    y1, sy1, sz1, yh1, xh1 = _per_frame(one, xb)
    with torch.no_grad():
        yB, _, _ = many.encode_latent(xb, type="float")           # B = 5 with max_batch = 3: chunks of 3 + 2
        oB = many.compress_from_latent(yB)
        yhB = many.decompress(oB["strings"], oB["z_shape"], return_format="latent")
        xhB = many.decode_latent(yhB)
    assert torch.equal(yB, y1)
    assert list(oB["strings"][0]) == sy1 and list(oB["strings"][1]) == sz1       # per-frame containers, byte for byte
    assert torch.equal(yhB, yh1)
    assert torch.equal(xhB, xh1)
    # and the batched model decodes what the single-frame model wrote (same containers either way)
    with torch.no_grad():
        assert torch.equal(many.decompress([sy1, sz1], oB["z_shape"], return_format="latent"), yh1)
        out = many.compress(xb)
        rec = many.decompress(out["strings"], out["z_shape"])["x_hat"]
    assert torch.equal(rec, xh1)


def test_batched_reference_format_and_precision_levels(nets):
    cfg, one, many, x = nets
    xb = x[:3].contiguous()
    # reference-format single streams: one frame per call underneath, same bytes as the single-frame model
    for net in (one, many):
        net.set_coder(format="ref")
    try:
        with torch.no_grad():
            o1 = [one.compress(xb[b:b + 1]) for b in range(3)]
            oB = many.compress(xb)
            assert [o["strings"][0][0] for o in o1] == list(oB["strings"][0])
            with pytest.warns(RuntimeWarning):
                from cra5_b200.vaeformer import VAEformer
                VAEformer._warned_ref_stream = False
                yh = many.decompress(oB["strings"], oB["z_shape"], return_format="latent")
            assert torch.equal(yh, torch.cat([one.decompress(o["strings"], o["z_shape"], return_format="latent") for o in o1]))
    finally:
        for net in (one, many):
            net.set_coder(16, 4)
    # split-bf16 precision levels are batch invariant too
    for net in (one, many):
        net.set_precision(2)
    try:
        with torch.no_grad():
            y1 = torch.cat([one.encode_latent(xb[b:b + 1], type="float")[0] for b in range(3)])
            yB = many.encode_latent(xb, type="float")[0]
            assert torch.equal(yB, y1)
            oB = many.compress_from_latent(yB)
            assert list(oB["strings"][0]) == [one.compress_from_latent(y1[b:b + 1])["strings"][0][0] for b in range(3)]
    finally:
        for net in (one, many):
            net.set_precision(0)


def test_batch_limits(nets):
    from cra5_b200.vaeformer import VAEformer
    cfg, one, many, x = nets
    with pytest.raises(ValueError):
        VAEformer(268, cfg=cfg, init_seed=None, max_batch=0)
    import ctypes
    from cra5_b200 import _lib
    y = torch.empty((4, cfg.latent_chans, *cfg.grid), device="cuda")
    with pytest.raises(ValueError, match="max_batch"):      # straight through the ABI: 4 frames into a 3-frame workspace
        _lib.check(_lib.lib.cra5_encode_to_latent_batch(many._handle, _lib.ptr(x), _lib.ptr(y), None, None, 4,
                                                        _lib.stream_ptr()))
