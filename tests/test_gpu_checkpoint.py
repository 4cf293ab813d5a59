"""Loader for the published checkpoint schema (missing item 2): `VAEformer.from_state_dict` strips the
`backbone.` prefix and drops `kl_loss.logvar` (vaeformer.py:168-185); `load_state_dict` honours CDF buffers shipped in
the checkpoint -- whatever their size -- without recomputing them (models/base.py:69-89, models/utils.py
update_registered_buffers). The real cra5_268v_300k.pth cannot be fetched offline, so a synthetic checkpoint with the
same key schema and NON-DEFAULT table sizes (40 scale levels instead of 64) pins the behaviour."""
import math
from collections import OrderedDict

import pytest
import torch

pytestmark = pytest.mark.gpu

from cra5_b200 import config as C
from oracle import entropy_oracle as EO, vaeformer_oracle as VO, weights


@pytest.fixture(scope="module")
def ckpt():
    cfg = C.tiny_fullres(69)
    sd = weights.seeded_state_dict(C.param_shapes(cfg), 7)
    scale_table = torch.exp(torch.linspace(math.log(0.2), math.log(100.0), 40))      # not the default 64 levels
    gc = EO.gaussian_conditional_tables(scale_table)
    eb = EO.entropy_bottleneck_tables(sd)
    ck = OrderedDict(("backbone." + k, v.clone()) for k, v in sd.items())
    ck["backbone.kl_loss.logvar"] = torch.zeros(())                                     # dropped by from_state_dict
    for mod, t in (("gaussian_conditional", gc), ("entropy_bottleneck", eb)):
        ck[f"backbone.{mod}._quantized_cdf"] = t.cdf.clone()
        ck[f"backbone.{mod}._cdf_length"] = t.cdf_length.clone()
        ck[f"backbone.{mod}._offset"] = t.offset.clone()
        ck[f"backbone.{mod}.likelihood_lower_bound.bound"] = torch.tensor([1e-9])
    ck["backbone.gaussian_conditional.scale_table"] = scale_table.clone()
    ck["backbone.gaussian_conditional.scale_bound"] = torch.tensor([0.11])
    ck["backbone.gaussian_conditional.lower_bound_scale.bound"] = torch.tensor([0.11])
    ck["backbone.entropy_bottleneck.target"] = torch.tensor([-6.9068, 0.0, 6.9068])
    return cfg, sd, ck, gc, eb, scale_table


def test_from_state_dict_strips_prefix_and_honours_shipped_tables(ckpt):
    from cra5_b200.vaeformer import VAEformer
    cfg, sd, ck, gc, eb, scale_table = ckpt
    net = VAEformer.from_state_dict(ck, cfg=cfg)
    out = net.state_dict()
    assert not any(k.startswith("backbone.") or "kl_loss" in k for k in out)
    for k, v in sd.items():
        assert torch.equal(out[k], v), k
    # shipped buffers are installed as they are: 40 rows, not update()'s 64
    assert torch.equal(out["gaussian_conditional._quantized_cdf"], gc.cdf) and gc.cdf.shape[0] == 40
    assert torch.equal(out["gaussian_conditional._cdf_length"], gc.cdf_length)
    assert torch.equal(out["gaussian_conditional._offset"], gc.offset)
    assert torch.equal(out["entropy_bottleneck._quantized_cdf"], eb.cdf)
    assert torch.equal(out["gaussian_conditional.scale_table"], scale_table)
    assert net.update() is False            # tables present, no force: nothing to do (models/base.py:91-115)

    # ...and they are what the coder uses, with no update() call: indexes follow the 40-level table, every stream decodes
    x = weights.seeded_frame(cfg, 1).unsqueeze(0)
    with torch.no_grad():
        y_dev, _, _ = net.encode_latent(x.cuda(), type="float")      # (held: the "y" tap points at this tensor)
        o = net.compress_from_latent(y_dev)
        sc = net.tap("scales").reshape(1, cfg.latent_chans, *cfg.grid).cpu()
        mu = net.tap("means").reshape(sc.shape).cpu()
        y = y_dev.cpu()
        idx, sym = net.tap("y_indexes").cpu(), net.tap("y_symbols").cpu()
        assert int(idx.max()) <= 39
        assert torch.equal(EO.build_indexes(sc, scale_table).reshape(-1).to(torch.uint8), idx)
        assert torch.equal(EO.quantize_symbols(y, mu).reshape(-1), sym)
        y_hat = net.decompress(o["strings"], o["z_shape"], return_format="latent")
        assert torch.equal(net.tap("y_symbols").cpu(), sym)
        assert torch.equal(y_hat.cpu(), sym.reshape(mu.shape).float() + mu)
    # reference-format stream of the same symbols with the shipped tables == the oracle coder's bytes
    net.set_coder(format="ref")
    with torch.no_grad():
        o_ref = net.compress_from_latent(y.cuda())
    assert o_ref["strings"][0][0] == EO.rans_encode(sym, idx.int(), *gc.coder_args())


def test_update_force_replaces_shipped_tables(ckpt):
    from cra5_b200.vaeformer import VAEformer
    cfg, sd, ck, gc, eb, scale_table = ckpt
    net = VAEformer.from_state_dict(ck, cfg=cfg)
    assert net.update(force=True) is True
    out = net.state_dict()
    assert out["gaussian_conditional._quantized_cdf"].shape[0] == 64
    assert torch.equal(out["gaussian_conditional._quantized_cdf"], EO.gaussian_conditional_tables().cdf)


def test_load_state_dict_errors_like_nn_module(ckpt):
    from cra5_b200.vaeformer import VAEformer
    cfg, sd, ck, gc, eb, scale_table = ckpt
    bad = OrderedDict(ck)
    del bad["backbone.g_a.blocks.0.attn.qkv.weight"]
    with pytest.raises(RuntimeError, match="missing keys"):
        VAEformer.from_state_dict(bad, cfg=cfg)
    bad = OrderedDict(ck)
    bad["backbone.quant_conv.bias"] = torch.zeros(3)
    with pytest.raises(RuntimeError, match="size mismatch"):
        VAEformer.from_state_dict(bad, cfg=cfg)
    bad = OrderedDict(ck)
    bad["backbone.not_a_parameter"] = torch.zeros(1)
    with pytest.raises(RuntimeError, match="unexpected keys"):
        VAEformer.from_state_dict(bad, cfg=cfg)
