"""Run one kernel configuration a few times (target for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cra5_b200 import _lib as L
what = sys.argv[1]
if what == "gemm":
    M, N, K, epi = (int(v) for v in sys.argv[2:6])
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16); B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if epi in (1, 2) else torch.float32)
    resid = torch.randn(M, N, device="cuda") if epi == 4 else None
    for _ in range(3):
        L.check(L.lib.cra5_op_gemm(L.ptr(A), K, L.ptr(B), K, M, N, K, L.ptr(bias), epi, L.ptr(out), N, L.ptr(resid), L.stream_ptr()))
elif what == "attn":
    heads, nseg, seg = (int(v) for v in sys.argv[2:5])
    rows = nseg * seg
    q = torch.randn(heads, rows, 64, device="cuda").to(torch.bfloat16); k = torch.randn(heads, rows, 64, device="cuda").to(torch.bfloat16)
    vt = torch.randn(heads, 64, rows, device="cuda").to(torch.bfloat16)
    out = torch.empty(rows, heads * 64, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        L.check(L.lib.cra5_op_attention(L.ptr(q), L.ptr(k), L.ptr(vt), L.ptr(out), heads * 64, heads, rows, seg, L.stream_ptr()))
torch.cuda.synchronize()
