"""Run-to-run determinism probes: attention kernel, encoder latent, entropy containers."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cra5_b200 import _lib as L

def attn_det(heads, nseg, seg):
    rows = nseg * seg
    g = torch.Generator(device="cuda").manual_seed(1)
    q = (torch.randn(heads, rows, 64, device="cuda", generator=g) * 0.25).to(torch.bfloat16)
    k = (torch.randn(heads, rows, 64, device="cuda", generator=g) * 2).to(torch.bfloat16)
    vt = torch.randn(heads, 64, rows, device="cuda", generator=g).to(torch.bfloat16)
    outs = []
    for _ in range(4):
        out = torch.zeros(rows, heads * 64, device="cuda", dtype=torch.bfloat16)
        L.check(L.lib.cra5_op_attention(L.ptr(q), L.ptr(k), L.ptr(vt), L.ptr(out), heads * 64, heads, rows, seg, L.stream_ptr()))
        torch.cuda.synchronize()
        outs.append(out)
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    print(f"attn heads={heads} nseg={nseg} seg={seg}: deterministic={same}", "" if same else (outs[0].float() - outs[1].float()).abs().max().item())

attn_det(16, 1, 10368); attn_det(16, 18, 576); attn_det(16, 24, 576)
if len(sys.argv) > 1 and sys.argv[1] == "model":
    from cra5_b200 import config as C
    from cra5_b200.vaeformer import VAEformer
    cfg = C.cra5_268()
    net = VAEformer(268, cfg=cfg, device="cuda:0", init_seed=3)
    sd = {k: v for k, v in net.state_dict().items() if k in C.param_shapes(cfg)}
    sd["quant_conv.weight"] = sd["quant_conv.weight"] * 6.0
    sd["h_s.final.weight"] = sd["h_s.final.weight"] * 12.0
    net.load_state_dict(sd); net.update(force=True)
    x = torch.randn(1, 268, 721, 1440, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    ys = [net.encode_latent(x, type="float")[0].clone() for _ in range(3)]
    print("encoder latent deterministic:", all(torch.equal(ys[0], y) for y in ys[1:]), (ys[0] - ys[1]).abs().max().item())
    ss = [net.compress_from_latent(ys[0]) for _ in range(3)]
    print("containers deterministic:", all(s["strings"] == ss[0]["strings"] for s in ss[1:]), [len(s["strings"][0][0]) for s in ss])
