"""Exploratory: run the CUDA model against the CPU oracle and print error statistics (not a test)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cra5_b200 import config as C
from cra5_b200.vaeformer import VAEformer
from oracle import weights, vaeformer_oracle as VO, entropy_oracle as EO

name = sys.argv[1] if len(sys.argv) > 1 else "small"
cfg = {"small": C.small_lowres(5), "tiny69": C.tiny_fullres(69)}[name]
sd = weights.seeded_state_dict(C.param_shapes(cfg), 11 if name == "small" else 7)
x = weights.seeded_frame(cfg, 3 if name == "small" else 1).unsqueeze(0)
codec = VO.OracleCodec(sd, cfg)
net = VAEformer(268, cfg=cfg, init_seed=None)
net.load_state_dict(sd)
net.update(force=True)

def stats(tag, a, b):
    a = a.float().cpu(); b = b.float().cpu()
    d = (a - b)
    print(f"  {tag:12s} max|d|={d.abs().max():.4e} rms(d)={d.pow(2).mean().sqrt():.4e} rms(ref)={b.pow(2).mean().sqrt():.4e} max|ref|={b.abs().max():.3e}")

with torch.no_grad():
    taps = {}
    y_o = VO.encode_y(codec.sd, cfg, x, taps)
    y_g, _, _ = net.encode_latent(x.cuda(), type="float")
    stats("tokens", net.tap("tokens").reshape(cfg.tokens, cfg.dim), taps["g_a.embed"][0])
    stats("y", y_g, y_o)
    out = net.compress_from_latent(y_g)
    print("  bytes y/z:", len(out["strings"][0][0]), len(out["strings"][1][0]))
    z_g = net.tap("z").reshape(1, cfg.z_chans, *cfg.hyper_grid)
    stats("z(h_a)", z_g, VO.h_a(codec.sd, cfg, y_g.cpu()))
    zhat_g = net.tap("z_hat").reshape(1, cfg.z_chans, *cfg.hyper_grid).cpu()
    sc_o, mu_o = VO.h_s(codec.sd, cfg, zhat_g)
    sc_g = net.tap("scales").reshape(1, cfg.latent_chans, *cfg.grid)
    mu_g = net.tap("means").reshape(1, cfg.latent_chans, *cfg.grid)
    stats("scales", sc_g, sc_o); stats("means", mu_g, mu_o)
    # integer parity given the GPU's own float tensors
    med = codec.sd["entropy_bottleneck.quantiles"][:, 0, 1].reshape(1, -1, 1, 1)
    zs_o = EO.quantize_symbols(z_g.cpu(), med)
    print("  z symbols equal:", torch.equal(zs_o.reshape(-1), net.tap("z_symbols").cpu()))
    ys_o = EO.quantize_symbols(y_g.cpu(), mu_g.cpu())
    idx_o = EO.build_indexes(sc_g.cpu(), codec.gc.scale_table)
    print("  y symbols equal:", torch.equal(ys_o.reshape(-1), net.tap("y_symbols").cpu()),
          " indexes equal:", torch.equal(idx_o.reshape(-1).to(torch.uint8), net.tap("y_indexes").cpu()))
    yhat_g = net.decompress(out["strings"], out["z_shape"], return_format="latent")
    print("  decoded y symbols equal encoded:", torch.equal(net.tap("y_symbols").cpu(), ys_o.reshape(-1)))
    stats("y_hat", yhat_g, ys_o.float() + mu_g.cpu())
    xh_g = net.decode_latent(yhat_g)
    xh_o = VO.decode_y(codec.sd, cfg, yhat_g.cpu())
    stats("x_hat|yhat", xh_g, xh_o)
    full_o = codec.decompress(codec.compress(x)["strings"], cfg.hyper_grid)["x_hat"]
    stats("x_hat e2e", xh_g, full_o)
    rm_g = ((xh_g.cpu()[0] - x[0]) ** 2).mean(dim=(1, 2)).sqrt()
    rm_o = ((full_o[0] - x[0]) ** 2).mean(dim=(1, 2)).sqrt()
    print("  RMSE per var: max|diff| =", (rm_g - rm_o).abs().max().item(), " mean rmse", rm_o.mean().item())
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(3):
        o = net.compress(x.cuda()); r = net.decompress(o["strings"], o["z_shape"])
    torch.cuda.synchronize(); print("  3x compress+decompress: %.1f ms each" % ((time.time() - t0) / 3 * 1e3))
