#!/usr/bin/env python
"""Generate tests/golden/* by running the REAL reference (imported from /root/reference through oracle/shims) and, in
the same pass, pin the oracle restatement (oracle/) against it.

Runs only in the build container (the reference mount does not exist on the GPU box). Fixtures are compact: integer
artefacts are stored whole or as sha256, float tensors as strided samples + sums, because inputs and weights are
regenerated from seeds (oracle/weights.py) at test time.

    python tools/make_golden.py            # all configs
    python tools/make_golden.py small      # one config
"""
import hashlib
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cra5_b200 import config as C  # noqa: E402
from oracle import ref_import, weights, entropy_oracle as EO, vaeformer_oracle as VO  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(t):
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().contiguous().numpy()
    if isinstance(t, np.ndarray):
        t = np.ascontiguousarray(t).tobytes()
    return hashlib.sha256(t).hexdigest()


def sample(t, n=4096):
    """deterministic strided sample + moments of a float tensor"""
    f = t.detach().reshape(-1).double()
    step = max(1, f.numel() // n)
    return {"shape": list(t.shape), "step": step, "values": f[::step][:n].float().numpy(),
            "sum": float(f.sum()), "abssum": float(f.abs().sum())}


def ref_model(r, cfg):
    dd = dict(arch="vit_large", pretrained_model="", patch_size=tuple(cfg.patch_size),
              patch_stride=tuple(cfg.patch_stride), in_chans=cfg.in_chans, out_chans=cfg.in_chans,
              kwargs=dict(z_dim=None, learnable_pos=True, window=True, window_size=[tuple(w) for w in cfg.window_sizes],
                          interval=cfg.interval, drop_path_rate=0., round_padding=True, pad_attn_mask=True,
                          test_pos_mode="learnable_simple_interpolate", lms_checkpoint_train=True,
                          img_size=tuple(cfg.img_size), embed_dim=cfg.dim, depth=cfg.depth,
                          num_heads=cfg.num_heads))
    pr = dict(pretrained_model="", patch_size=tuple(cfg.hyper_patch), in_chans=cfg.latent_chans,
              out_chans=cfg.latent_chans,
              kwargs=dict(z_dim=cfg.z_chans, embed_dim=cfg.hyper_dim, depth=cfg.hyper_depth,
                          num_heads=cfg.hyper_heads, interval=1, learnable_pos=True, window=False,
                          drop_path_rate=0., round_padding=True, pad_attn_mask=True,
                          test_pos_mode="learnable_simple_interpolate", lms_checkpoint_train=False,
                          img_size=tuple(cfg.grid)))
    net = r.VAEformer(cfg.in_chans, embed_dim=cfg.latent_chans, z_channels=cfg.z_chans, y_channels=cfg.dim,
                      lower_dim=True, sample_posterior=False, frozen_encoder=False, ddconfig=dd, priorconfig=pr)
    return net.eval()


def run(name, cfg, wseed, fseed):
    r = ref_import.load()
    torch.set_num_threads(os.cpu_count())
    net = ref_model(r, cfg)
    # 1. schema pin: my param_shapes == the reference's float parameters, same order
    shapes = C.param_shapes(cfg)
    ref_sd = net.state_dict()
    ref_float = [(k, tuple(v.shape)) for k, v in ref_sd.items() if k not in C.BUFFER_KEYS]
    assert ref_float == [(k, tuple(s)) for k, s in shapes.items()], "param_shapes() disagrees with the reference"
    sd = weights.seeded_state_dict(shapes, wseed)
    full = dict(ref_sd)  # the reference's loader insists on the (still empty) CDF buffers being present
    full.update(sd)
    net.load_state_dict(full, strict=True)
    net.update(force=True)

    x = weights.seeded_frame(cfg, fseed).unsqueeze(0)
    out = {"meta": json.dumps({"name": name, "config": cfg.to_dict(), "weight_seed": wseed, "frame_seed": fseed,
                               "torch": torch.__version__})}
    codec = VO.OracleCodec(sd, cfg)

    # 2. CDF tables (int): reference == oracle
    gcr, ebr = net.gaussian_conditional, net.entropy_bottleneck
    for tag, mod, tab in (("gc", gcr, codec.gc), ("eb", ebr, codec.eb)):
        assert torch.equal(mod._quantized_cdf.cpu(), tab.cdf), f"{tag} cdf differs"
        assert torch.equal(mod._cdf_length.cpu().int(), tab.cdf_length), f"{tag} cdf_length differs"
        assert torch.equal(mod._offset.cpu().int(), tab.offset), f"{tag} offset differs"
        out[f"{tag}_cdf_sha"] = sha(tab.cdf)
        out[f"{tag}_cdf_shape"] = np.array(tab.cdf.shape)
        out[f"{tag}_cdf_length"] = tab.cdf_length.numpy()
        out[f"{tag}_offset"] = tab.offset.numpy()
    out["eb_cdf"] = codec.eb.cdf.numpy()
    out["gc_cdf_row0"] = codec.gc.cdf[0, :8].numpy()
    out["gc_cdf_row20"] = codec.gc.cdf[20, :64].numpy()
    out["gc_scale_table"] = codec.gc.scale_table.numpy()
    assert torch.equal(gcr.scale_table.cpu(), codec.gc.scale_table)

    with torch.no_grad():
        # 3. the reference's own run
        t0 = time.perf_counter()
        ref_c = net.compress(x)
        t1 = time.perf_counter()
        ref_d = net.decompress(ref_c["strings"], ref_c["z_shape"])
        t2 = time.perf_counter()
        ref_y, _, _ = net.encode_latent(x, type="float")
        ref_yhat = net.decompress(ref_c["strings"], ref_c["z_shape"], return_format="latent")
        ref_fwd = net(x)
        ref_rec = net.decode_latent(ref_yhat)
        # intermediate tensors of the reference for pinning
        z = net.h_a(ref_y)
        z_hat = net.entropy_bottleneck.decompress(ref_c["strings"][1], z.size()[-2:])
        sc, mu = net.h_s(z_hat).chunk(2, 1)
        idx = net.gaussian_conditional.build_indexes(sc)
        ysym = net.gaussian_conditional.quantize(ref_y, "symbols", mu)
        zsym = torch.round(z - net.entropy_bottleneck._get_medians().reshape(1, -1, 1, 1)).int()
        print(f"[{name}] reference compress {t1 - t0:.2f}s decompress {t2 - t1:.2f}s "
              f"y_bytes {len(ref_c['strings'][0][0])} z_bytes {len(ref_c['strings'][1][0])}")

        # 4. the oracle on the same input
        t0 = time.perf_counter()
        o_c = codec.compress(x)
        o_d = codec.decompress(o_c["strings"], o_c["z_shape"])
        print(f"[{name}] oracle compress+decompress {time.perf_counter() - t0:.2f}s")
        dbg = o_c["debug"]

    def close(a, b, what, tol=2e-5):
        err = (a - b).abs().max().item()
        ref = b.abs().max().item()
        print(f"    {what:14s} max|diff| {err:.3e} (max|ref| {ref:.3e})")
        assert err <= tol * max(1.0, ref), what

    # pins: float tensors close, integer artefacts identical
    close(dbg["y"], ref_y, "y")
    close(dbg["z"], z, "z")
    close(dbg["scales"], sc, "scales")
    close(dbg["means"], mu, "means")
    n_sym_diff = (dbg["y_symbols"] != ysym).sum().item()
    n_idx_diff = (dbg["indexes"] != idx).sum().item()
    print(f"    y symbol mismatches {n_sym_diff}, index mismatches {n_idx_diff}, z symbol mismatches "
          f"{(dbg['z_symbols'] != zsym).sum().item()}")
    assert torch.equal(dbg["z_symbols"], zsym)
    assert n_sym_diff == 0 and n_idx_diff == 0
    assert o_c["strings"][0][0] == ref_c["strings"][0][0], "y stream differs from the reference coder"
    assert o_c["strings"][1][0] == ref_c["strings"][1][0], "z stream differs from the reference coder"
    close(o_d["x_hat"], ref_d["x_hat"], "x_hat")
    # rate estimation path (forward() likelihoods, vaeformer.py:302-333): oracle == reference
    o_f = codec.forward(x)
    close(o_f["likelihoods"]["y"], ref_fwd["likelihoods"]["y"], "y likelihood", 1e-6)
    close(o_f["likelihoods"]["z"], ref_fwd["likelihoods"]["z"], "z likelihood", 1e-6)
    out["bits_y"] = np.array(float(-torch.log2(ref_fwd["likelihoods"]["y"]).double().sum()))
    out["bits_z"] = np.array(float(-torch.log2(ref_fwd["likelihoods"]["z"]).double().sum()))
    for tag, tt in (("lik_y", ref_fwd["likelihoods"]["y"]), ("lik_z", ref_fwd["likelihoods"]["z"])):
        sm = sample(tt)
        out[f"s_{tag}_values"] = sm["values"]
        out[f"s_{tag}_info"] = np.array([sm["step"], sm["sum"], sm["abssum"]], dtype=np.float64)
        out[f"s_{tag}_shape"] = np.array(sm["shape"])
    # the reference's own self-consistency invariants
    close(ref_fwd["x_hat"], ref_d["x_hat"], "fwd vs coded", 1e-4)
    close(ref_rec, ref_d["x_hat"], "decode_latent")

    xh = ref_d["x_hat"][0]
    rmse = ((xh - x[0]) ** 2).mean(dim=(1, 2)).sqrt()
    out.update({
        "y": ref_y[0].numpy() if ref_y.numel() <= 400_000 else np.zeros(0, np.float32),
        "rmse_per_var": rmse.numpy(),
        "xhat_mean_per_var": xh.mean(dim=(1, 2)).numpy(),
        "xhat_std_per_var": xh.std(dim=(1, 2)).numpy(),
        "y_string": np.frombuffer(ref_c["strings"][0][0], dtype=np.uint8),
        "z_string": np.frombuffer(ref_c["strings"][1][0], dtype=np.uint8),
        "y_string_sha": sha(ref_c["strings"][0][0]), "z_string_sha": sha(ref_c["strings"][1][0]),
        "z_shape": np.array(ref_c["z_shape"]),
        "y_symbols_sha": sha(ysym.int()), "z_symbols_sha": sha(zsym.int()), "indexes_sha": sha(idx.int()),
        "index_hist": torch.bincount(idx.reshape(-1).long(), minlength=64).numpy(),
        "y_symbols_absmax": np.array(int(ysym.abs().max())),
    })
    for tag, t in (("x", x), ("y", ref_y), ("z", z), ("z_hat", z_hat), ("scales", sc), ("means", mu),
                   ("y_hat", ref_yhat), ("x_hat", ref_d["x_hat"])):
        s = sample(t)
        out[f"s_{tag}_values"] = s["values"]
        out[f"s_{tag}_info"] = np.array([s["step"], s["sum"], s["abssum"]], dtype=np.float64)
        out[f"s_{tag}_shape"] = np.array(s["shape"])
    if x.numel() <= 400_000:  # small config: whole tensors
        out["full_x_hat"] = ref_d["x_hat"][0].numpy()
        out["full_scales"] = sc[0].numpy()
        out["full_means"] = mu[0].numpy()
        out["full_z"] = z[0].numpy()
    path = os.path.join(GOLD, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"[{name}] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


def kat_tables(k_count=4):
    """c = 4(k+1), pmf_i = exp(-|i-c|/(k+1)), normalised to 1-1e-3, tail 1e-3"""
    r = ref_import.load()
    cdfs, sizes, offsets = [], [], []
    for k in range(k_count):
        c = 4 * (k + 1)
        pmf = np.exp(-np.abs(np.arange(2 * c + 1) - c) / (k + 1.0))
        pmf = (pmf / pmf.sum() * (1 - 1e-3)).astype(np.float32)
        full = [float(v) for v in pmf] + [1e-3]
        cdf = r._CXX.pmf_to_quantized_cdf(full, 16)
        assert list(EO.pmf_to_quantized_cdf(np.array(full, dtype=np.float32))) == list(cdf)
        cdfs.append(list(cdf)); sizes.append(len(cdf)); offsets.append(-c)
    width = max(sizes)
    cdfs = [c + [0] * (width - len(c)) for c in cdfs]
    return cdfs, sizes, offsets


synth_entropy_case = weights.synth_entropy_case


def kats():
    r = ref_import.load()
    enc, dec = r.ans.RansEncoder(), r.ans.RansDecoder()
    out = {}
    # KAT-0
    cdf0 = [[0, 6554, 19661, 65536]]
    sym0 = [0, 1, 2, -5, 9, 1]
    b0 = enc.encode_with_indexes(sym0, [0] * 6, cdf0, [4], [0])
    assert b0.hex() == "8203223d9dcac616" and dec.decode_with_indexes(b0, [0] * 6, cdf0, [4], [0]) == sym0
    assert EO.rans_encode(sym0, [0] * 6, cdf0, [4], [0]) == b0
    out["kat0"] = {"symbols": sym0, "indexes": [0] * 6, "cdfs": cdf0, "sizes": [4], "offsets": [0], "hex": b0.hex()}
    assert list(r._CXX.pmf_to_quantized_cdf([0.1, 0.2, 0.7], 16)) == [0, 6554, 19661, 65536]
    out["pmf_kat"] = {"pmf": [0.1, 0.2, 0.7], "cdf": [0, 6554, 19661, 65536]}
    # KAT-1
    cdfs, sizes, offsets = kat_tables()
    rng = np.random.default_rng(1234)
    idx = rng.integers(0, 4, 100000)
    sym = np.rint(rng.normal(0, 1, 100000) * (idx + 1) * 2).astype(np.int64)
    b1 = enc.encode_with_indexes(sym.tolist(), idx.tolist(), cdfs, sizes, offsets)
    assert dec.decode_with_indexes(b1, idx.tolist(), cdfs, sizes, offsets) == sym.tolist()
    ob1 = EO.rans_encode(sym, idx, cdfs, sizes, offsets)
    assert ob1 == b1, "oracle coder differs from the reference coder on KAT-1"
    assert EO.rans_decode(b1, idx, cdfs, sizes, offsets).tolist() == sym.tolist()
    out["kat1"] = {"cdfs": cdfs, "sizes": sizes, "offsets": offsets, "n": 100000, "rng_seed": 1234,
                   "nbytes": len(b1), "sha256": sha(b1), "first16": b1[:16].hex()}
    print(f"[kat1] {len(b1)} bytes sha256 {sha(b1)}")
    # synthetic Gaussian-conditional cases through the reference's own modules
    gc = r.GaussianConditional(None)
    gc.update_scale_table(EO.get_scale_table(), force=True)
    tabs = EO.gaussian_conditional_tables()
    assert torch.equal(gc._quantized_cdf, tabs.cdf)
    out["gc_cdf_sha"] = sha(tabs.cdf)
    cases = []
    for seed, n in ((101, 50000), (102, 200000)):
        y, sig, mu = synth_entropy_case(seed, n)
        y4, s4, m4 = (t.reshape(1, 1, 1, -1) for t in (y, sig, mu))
        idx_r = gc.build_indexes(s4)
        strings = gc.compress(y4, idx_r, means=m4)
        sym_r = gc.quantize(y4, "symbols", m4)
        yhat_r = gc.decompress(strings, idx_r, means=m4)
        idx_o = EO.build_indexes(s4, tabs.scale_table)
        sym_o = EO.quantize_symbols(y4, m4)
        assert torch.equal(idx_o, idx_r) and torch.equal(sym_o, sym_r)
        so = EO.rans_encode(sym_o.reshape(-1), idx_o.reshape(-1), *tabs.coder_args())
        assert so == strings[0]
        assert torch.equal(EO.dequantize(EO.rans_decode(so, idx_o.reshape(-1), *tabs.coder_args()).reshape(y4.shape), m4), yhat_r)
        hist = torch.bincount(idx_r.reshape(-1).long(), minlength=64)
        assert (hist > 0).all(), "synthetic case must touch all 64 tables"
        cases.append({"seed": seed, "n": n, "symbols_sha": sha(sym_r.int()), "indexes_sha": sha(idx_r.int()),
                      "stream_sha": sha(strings[0]), "nbytes": len(strings[0]), "yhat_sha": sha(yhat_r),
                      "index_hist": hist.tolist()})
        print(f"[synth {seed}] n={n} stream {len(strings[0])} bytes")
    out["synthetic"] = cases
    with open(os.path.join(GOLD, "rans_kat.json"), "w") as f:
        json.dump(out, f)
    print("wrote rans_kat.json")


CONFIGS = {
    "small": (C.small_lowres(5), 11, 3),
    "tiny69": (C.tiny_fullres(69), 7, 1),
}

if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    todo = sys.argv[1:] or (["kats"] + list(CONFIGS))
    for n in todo:
        if n == "kats":
            kats()
            continue
        cfg, ws, fs = CONFIGS[n]
        run(n, cfg, ws, fs)
