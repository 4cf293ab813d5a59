"""CPU tests: the oracle restatement (oracle/) is pinned to fixtures produced by the REAL reference
(tools/make_golden.py, run in the build container) -- and, when the reference mount is present, to the reference's
own compiled coder directly."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from cra5_b200 import config as C
from oracle import entropy_oracle as EO, vaeformer_oracle as VO, weights, ref_import

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(t):
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().contiguous().numpy()
    if isinstance(t, np.ndarray):
        t = np.ascontiguousarray(t).tobytes()
    return hashlib.sha256(t).hexdigest()


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLD, "rans_kat.json")) as f:
        return json.load(f)


def test_pmf_to_quantized_cdf_kat(kat):
    assert EO.pmf_to_quantized_cdf(kat["pmf_kat"]["pmf"]).tolist() == kat["pmf_kat"]["cdf"]
    with pytest.raises(ValueError):
        EO.pmf_to_quantized_cdf([0.5, -0.1])
    with pytest.raises(ValueError):
        EO.pmf_to_quantized_cdf([0.0, 0.0])


def test_rans_kat0_bytes_and_roundtrip(kat):
    k = kat["kat0"]
    b = EO.rans_encode(k["symbols"], k["indexes"], k["cdfs"], k["sizes"], k["offsets"])
    assert b.hex() == k["hex"] == "8203223d9dcac616"
    assert EO.rans_decode(b, k["indexes"], k["cdfs"], k["sizes"], k["offsets"]).tolist() == k["symbols"]


def test_rans_kat1_sha(kat):
    k = kat["kat1"]
    rng = np.random.default_rng(k["rng_seed"])
    idx = rng.integers(0, 4, k["n"])
    sym = np.rint(rng.normal(0, 1, k["n"]) * (idx + 1) * 2).astype(np.int64)
    b = EO.rans_encode(sym, idx, k["cdfs"], k["sizes"], k["offsets"])
    assert len(b) == k["nbytes"] == 59068
    assert sha(b) == k["sha256"] == "947e47b80c531cd3338761e906a873237004c059375717a5756be4546254d478"
    assert b[:16].hex() == k["first16"]
    assert EO.rans_decode(b, idx, k["cdfs"], k["sizes"], k["offsets"]).tolist() == sym.tolist()


def test_rans_empty_and_single():
    cdfs, sizes, offs = [[0, 6554, 19661, 65536]], [4], [0]
    b = EO.rans_encode([], [], cdfs, sizes, offs)
    assert len(b) == 8  # just the flushed state
    assert EO.rans_decode(b, [], cdfs, sizes, offs).numel() == 0
    for s in (0, 1, 2, 3, -1, 100000, -100000):
        b = EO.rans_encode([s], [0], cdfs, sizes, offs)
        assert EO.rans_decode(b, [0], cdfs, sizes, offs).tolist() == [s]


def test_gaussian_tables_and_synthetic_cases(kat):
    tabs = EO.gaussian_conditional_tables()
    assert tuple(tabs.cdf.shape) == (64, 3133)
    assert sha(tabs.cdf) == kat["gc_cdf_sha"]
    assert tabs.cdf_length.min().item() == 5 and tabs.cdf_length.max().item() == 3133
    assert tabs.offset.max().item() == -1 and tabs.offset.min().item() == -1565
    for case in kat["synthetic"]:
        y, sig, mu = weights.synth_entropy_case(case["seed"], case["n"])
        idx = EO.build_indexes(sig, tabs.scale_table)
        sym = EO.quantize_symbols(y, mu)
        assert sha(idx.int()) == case["indexes_sha"]
        assert sha(sym.int()) == case["symbols_sha"]
        assert torch.bincount(idx.long(), minlength=64).tolist() == case["index_hist"]
        s = EO.rans_encode(sym, idx, *tabs.coder_args())
        assert len(s) == case["nbytes"] and sha(s) == case["stream_sha"]
        yhat = EO.dequantize(EO.rans_decode(s, idx, *tabs.coder_args()), mu)
        assert sha(yhat) == case["yhat_sha"]


def _load(name):
    g = np.load(os.path.join(GOLD, f"{name}.npz"), allow_pickle=False)
    meta = json.loads(str(g["meta"]))
    cfgd = dict(meta["config"])
    for k in ("img_size", "patch_size", "patch_stride", "hyper_patch"):
        cfgd[k] = tuple(cfgd[k])
    cfgd["window_sizes"] = [tuple(w) for w in cfgd["window_sizes"]]
    cfg = C.VaeformerConfig(**cfgd).validate()
    return g, meta, cfg


def _check_sample(g, tag, t, tol=2e-5):
    step = int(g[f"s_{tag}_info"][0])
    assert list(t.shape) == g[f"s_{tag}_shape"].tolist(), tag
    v = t.detach().reshape(-1)[::step][: len(g[f"s_{tag}_values"])].float().numpy()
    ref = g[f"s_{tag}_values"]
    assert np.abs(v - ref).max() <= tol * max(1.0, np.abs(ref).max()), tag
    assert abs(float(t.double().abs().sum()) - g[f"s_{tag}_info"][2]) <= 1e-5 * g[f"s_{tag}_info"][2] + 1e-6, tag


@pytest.mark.parametrize("name", ["small", "tiny69"])
def test_oracle_codec_matches_reference_fixture(name):
    g, meta, cfg = _load(name)
    sd = weights.seeded_state_dict(C.param_shapes(cfg), meta["weight_seed"])
    x = weights.seeded_frame(cfg, meta["frame_seed"]).unsqueeze(0)
    _check_sample(g, "x", x, 1e-6)
    codec = VO.OracleCodec(sd, cfg)
    for tag, tab in (("gc", codec.gc), ("eb", codec.eb)):
        assert sha(tab.cdf) == str(g[f"{tag}_cdf_sha"])
        assert tab.cdf_length.tolist() == g[f"{tag}_cdf_length"].tolist()
        assert tab.offset.tolist() == g[f"{tag}_offset"].tolist()
    assert np.array_equal(codec.eb.cdf.numpy(), g["eb_cdf"])
    with torch.no_grad():
        out = codec.compress(x)
        d = out["debug"]
        _check_sample(g, "y", d["y"])
        _check_sample(g, "z", d["z"])
        _check_sample(g, "scales", d["scales"])
        _check_sample(g, "means", d["means"])
        assert sha(d["y_symbols"].int()) == str(g["y_symbols_sha"])
        assert sha(d["z_symbols"].int()) == str(g["z_symbols_sha"])
        assert sha(d["indexes"].int()) == str(g["indexes_sha"])
        # byte-identical to the streams the reference's own coder wrote
        assert out["strings"][0][0] == g["y_string"].tobytes()
        assert out["strings"][1][0] == g["z_string"].tobytes()
        assert tuple(out["z_shape"]) == tuple(g["z_shape"].tolist())
        y_hat = codec.decompress(out["strings"], out["z_shape"], return_format="latent")
        _check_sample(g, "y_hat", y_hat)
        x_hat = codec.decompress(out["strings"], out["z_shape"])["x_hat"]
        _check_sample(g, "x_hat", x_hat)
        rmse = ((x_hat[0] - x[0]) ** 2).mean(dim=(1, 2)).sqrt().numpy()
        assert np.abs(rmse - g["rmse_per_var"]).max() <= 1e-5
        if "full_x_hat" in g.files:
            assert np.abs(x_hat[0].numpy() - g["full_x_hat"]).max() <= 2e-5
        # dequantize path == coded path (reference invariant)
        f = codec.forward(x)
        assert (f["x_hat"] - x_hat).abs().max().item() <= 1e-4
        # rate-estimation outputs pinned to the reference's forward()
        _check_sample(g, "lik_y", f["likelihoods"]["y"], 1e-6)
        _check_sample(g, "lik_z", f["likelihoods"]["z"], 1e-6)
        assert abs(float(-torch.log2(f["likelihoods"]["y"]).double().sum()) - float(g["bits_y"])) <= 1e-6 * float(g["bits_y"])


@pytest.mark.skipif(not ref_import.available(), reason="reference mount not present (GPU box)")
def test_oracle_coder_equals_reference_coder_random():
    r = ref_import.load()
    tabs = EO.gaussian_conditional_tables()
    cdfs, sizes, offs = tabs.cdf.tolist(), tabs.cdf_length.tolist(), tabs.offset.tolist()
    rng = np.random.default_rng(7)
    # n >= 4: the reference sizes its output buffer as one word per coded symbol (rans_interface.cpp:179), which
    # under-allocates (and corrupts the heap) when fewer than two words of payload exist
    for n in (4, 31, 1000, 20000):
        idx = rng.integers(0, 64, n)
        sym = np.rint(rng.normal(0, 1, n) * np.exp(rng.uniform(-2, 6, n))).astype(np.int64)
        ref = r.ans.RansEncoder().encode_with_indexes(sym.tolist(), idx.tolist(), cdfs, sizes, offs)
        assert EO.rans_encode(sym, idx, *tabs.coder_args()) == ref
        assert r.ans.RansDecoder().decode_with_indexes(ref, idx.tolist(), cdfs, sizes, offs) == sym.tolist()
        assert EO.rans_decode(ref, idx, *tabs.coder_args()).tolist() == sym.tolist()
