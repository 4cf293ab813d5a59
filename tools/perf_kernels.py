"""Micro-benchmarks of the individual kernels (CUDA events, L2 flushed between iterations)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cra5_b200 import _lib as L

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

def gemm(M, N, K, epi=1):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if epi in (1, 2) else torch.float32)
    resid = torch.randn(M, N, device="cuda") if epi == 4 else None
    def f():
        L.check(L.lib.cra5_op_gemm(L.ptr(A), K, L.ptr(B), K, M, N, K, L.ptr(bias), epi, L.ptr(out), N, L.ptr(resid), L.stream_ptr()))
    ms = timeit(f)
    tf = 2.0 * M * N * K / ms / 1e9
    print(f"gemm M={M} N={N} K={K} epi={epi}: {ms:.3f} ms  {tf:.1f} TFLOP/s")
    def g():
        torch.matmul(A, B.t())
    ms2 = timeit(g)
    print(f"   cublas bf16: {ms2:.3f} ms {2.0*M*N*K/ms2/1e9:.1f} TFLOP/s")

def attn(heads, nseg, seg):
    rows = nseg * seg
    q = torch.randn(heads, rows, 64, device="cuda").to(torch.bfloat16)
    k = torch.randn(heads, rows, 64, device="cuda").to(torch.bfloat16)
    vt = torch.randn(heads, 64, rows, device="cuda").to(torch.bfloat16)
    out = torch.empty(rows, heads * 64, device="cuda", dtype=torch.bfloat16)
    def f():
        L.check(L.lib.cra5_op_attention(L.ptr(q), L.ptr(k), L.ptr(vt), L.ptr(out), heads * 64, heads, rows, seg, L.stream_ptr()))
    ms = timeit(f)
    fl = 4.0 * heads * nseg * seg * seg * 64
    print(f"attn heads={heads} nseg={nseg} seg={seg}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s")

if __name__ == "__main__" and len(sys.argv) == 1:
    gemm(10368, 3072, 1024, 1)
    gemm(10368, 4096, 1024, 2)
    gemm(10368, 1024, 4096, 4)
    gemm(10368, 1024, 1024, 4)
    gemm(13824, 3072, 1024, 1)
    gemm(648, 1080, 360, 1)
    attn(16, 1, 10368)
    attn(16, 18, 576)
    attn(16, 24, 576)

if len(sys.argv) > 1 and sys.argv[1] == "epi":
    for epi in (0, 1, 4):
        gemm(10368, 1024, 1024, epi)
        gemm(10368, 1024, 4096, epi)
