"""How the synthetic weight widening of bench.py shapes the entropy workload: bytes per frame, share of bypass symbols,
coder kernel times, for a few (quant_conv, h_s.final) scale pairs."""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cra5_b200 import _lib, config as C
from cra5_b200.vaeformer import VAEformer
cfg = C.cra5_268()
net = VAEformer(268, cfg=cfg, device="cuda:0", init_seed=1234)
sd0 = {k: v.clone() for k, v in net.state_dict().items() if k in C.param_shapes(cfg)}
x = torch.randn(1, 268, 721, 1440, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1000))
for qs, hs in [(6, 12), (3, 12), (2, 24), (1.5, 40), (1, 40)]:
    sd = dict(sd0)
    sd["quant_conv.weight"] = sd0["quant_conv.weight"] * qs
    sd["h_s.final.weight"] = sd0["h_s.final.weight"] * hs
    net.load_state_dict(sd); net.update(force=True)
    out = net.compress(x)
    rec = net.decompress(out["strings"], out["z_shape"])
    sym = net.tap("y_symbols").long(); idx = net.tap("y_indexes").long()
    sdn = net.state_dict(); off = sdn["gaussian_conditional._offset"].long().cuda(); ln = sdn["gaussian_conditional._cdf_length"].long().cuda()
    v = sym - off[idx]
    byp = ((v < 0) | (v >= ln[idx] - 2)).float().mean().item()
    _lib.check(_lib.lib.cra5_profile_enable(1))
    out = net.compress(x); rec = net.decompress(out["strings"], out["z_shape"])
    buf = ctypes.create_string_buffer(1 << 20); need = ctypes.c_uint64()
    _lib.check(_lib.lib.cra5_profile_report(buf, ctypes.c_uint64(len(buf)), ctypes.byref(need)))
    _lib.check(_lib.lib.cra5_profile_enable(0))
    k = json.loads(buf.value.decode())
    t = {n: round(sum(v["ms"] for kk, v in k.items() if kk.split(":")[0] == n), 3) for n in ("rans_encode", "rans_decode")}
    print(f"quant x{qs} h_s x{hs}: y {len(out['strings'][0][0])/1e6:.2f} MB z {len(out['strings'][1][0])/1e3:.0f} KB  bypass {100*byp:.1f}%  "
          f"sym std {sym.float().std().item():.2f} idx mean {idx.float().mean().item():.1f}  {t}")
