"""GPU parity of the individual sm_100a kernels, called through the C ABI (ctypes).

Floating-point kernels: compared against a plain torch fp32 reference of the same op fed the SAME bf16-rounded
operands, so the tolerance only has to absorb fp32 summation order (GEMM) or bf16 rounding of P (attention).
"""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from cra5_b200 import _lib
    return _lib


def _gemm(A, B, bias, epi, out, resid=None):
    L = _lib()
    M, K = A.shape
    N = B.shape[0]
    ldo = out.stride(0)
    L.check(L.lib.cra5_op_gemm(L.ptr(A), A.stride(0), L.ptr(B), B.stride(0), M, N, K, L.ptr(bias), epi, L.ptr(out),
                               ldo, L.ptr(resid), L.stream_ptr()))
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 128, 128), (256, 512, 1024), (10368, 1024, 1024),
                                   (648, 360, 360), (648, 1080, 360), (300, 200, 72), (10368, 3072, 1024),
                                   (129, 257, 200)])
def test_gemm_f32_matches_fp32_reference(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    Kp = (K + 7) // 8 * 8  # row stride must be a multiple of 8 elements (16 B) for TMA
    A = torch.zeros(M, Kp, device="cuda", dtype=torch.bfloat16)
    B = torch.zeros(N, Kp, device="cuda", dtype=torch.bfloat16)
    A[:, :K] = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    B[:, :K] = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda")
    _gemm(A[:, :K], B[:, :K], bias, 0, out)
    ref = A[:, :K].float() @ B[:, :K].float().t() + bias
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-4 * max(scale, 1.0), (err, scale)
    # the SIMT self-check kernel agrees too
    L = _lib()
    chk = torch.empty_like(out)
    L.check(L.lib.cra5_op_gemm_check(L.ptr(A), A.stride(0), L.ptr(B), B.stride(0), M, N, K, L.ptr(bias), L.ptr(chk),
                                     chk.stride(0), L.stream_ptr()))
    torch.cuda.synchronize()
    assert (chk - ref).abs().max().item() <= 2e-4 * max(scale, 1.0)


def test_gemm_epilogues():
    torch.manual_seed(0)
    M, N, K = 520, 384, 256
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda") * 0.1).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    ref = A.float() @ B.float().t() + bias
    # bf16 out
    o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    _gemm(A, B, bias, 1, o)
    assert (o.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    # gelu (exact erf, vit_nlc.py:53 nn.GELU default)
    _gemm(A, B, bias, 2, o)
    rg = torch.nn.functional.gelu(ref)
    assert (o.float() - rg).abs().max().item() <= 2 ** -7 * rg.abs().max().item()
    # the trunk's fc1 form of the same function (one MUFU.TANH of a quintic fitted to the erf form): against fp64 erf-GELU
    # of the GPU's own pre-activation it must stay within a bf16 rounding (2^-8 relative) plus the fit's 3e-5
    _gemm(A, B, bias, 9, o)
    o_exact = torch.empty_like(o)
    _gemm(A, B, bias, 2, o_exact)
    pre = torch.empty(M, N, device="cuda")
    _gemm(A, B, bias, 0, pre)
    g64 = torch.nn.functional.gelu(pre.double())
    tol = 2.0 ** -8 * g64.abs() + 3.5e-5
    assert ((o.double() - g64).abs() <= tol).all()
    assert ((o_exact.double() - g64).abs() <= 2.0 ** -8 * g64.abs() + 1e-6).all()
    assert (o.float() - o_exact.float()).abs().max().item() <= 2 ** -7 * rg.abs().max().item()
    # residual
    resid = torch.randn(M, N, device="cuda")
    of = torch.empty(M, N, device="cuda")
    _gemm(A, B, bias, 4, of, resid)
    assert (of - (ref + resid)).abs().max().item() <= 2e-4 * ref.abs().max().item()
    # in-place residual
    r2 = resid.clone()
    _gemm(A, B, bias, 4, r2, r2)
    assert torch.equal(r2, of)
    # transposed (channel-major) output
    ot = torch.empty(N, M, device="cuda")
    _gemm(A, B, bias, 5, ot)
    assert (ot.t() - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()


def test_gemm_cta_pair_sites():
    """the shapes / epilogues that take the CTA-pair kernel (csrc/gemm_tc.cu::gemm_use_pair: fc1's GELU with N >= 3072,
    the residual epilogue with N in [512, 1024], K >= 1024, M >= 4096), with an M that is not a multiple of the 256-row
    pair tile, against torch on the same bf16 operands"""
    torch.manual_seed(3)
    M, K = 4096 + 200, 1024
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    # fc1-like: GELU, bf16 out (both the erf form and the trunk's tanh-fitted form)
    N = 3072
    B = (torch.randn(N, K, device="cuda") * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    ref = torch.nn.functional.gelu(A.float() @ B.float().t() + bias)
    for epi in (2, 9):
        o = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
        _gemm(A, B, bias, epi, o)
        assert (o.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    # proj / fc2-like: residual, fp32 out, in place
    N = 1024
    B = (torch.randn(N, K, device="cuda") * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda")
    ref = resid + A.float() @ B.float().t() + bias
    _gemm(A, B, bias, 4, resid, resid)
    assert (resid - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()


def test_gemm_is_deterministic():
    torch.manual_seed(1)
    M, N, K = 648, 8192, 360
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda") * 0.1).to(torch.bfloat16)
    o1 = torch.empty(M, N, device="cuda")
    o2 = torch.empty(M, N, device="cuda")
    _gemm(A, B, None, 0, o1)
    _gemm(A, B, None, 0, o2)
    assert torch.equal(o1, o2)


def _attention_ref(q, k, v):  # [H, S, d] fp32, q pre-scaled
    s = q @ k.transpose(-1, -2)
    return torch.softmax(s, dim=-1) @ v


@pytest.mark.parametrize("heads,nseg,seg_len", [(2, 1, 128), (2, 3, 576), (4, 1, 1024), (16, 2, 576), (2, 1, 10368)])
def test_attention_matches_fp32_reference(heads, nseg, seg_len):
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(heads * 100 + seg_len)
    rows = nseg * seg_len
    q = (torch.randn(heads, rows, 64, device="cuda", generator=g) * 0.125 * 2.0).to(torch.bfloat16)
    k = (torch.randn(heads, rows, 64, device="cuda", generator=g) * 2.0).to(torch.bfloat16)
    v = torch.randn(heads, rows, 64, device="cuda", generator=g).to(torch.bfloat16)
    vt = v.transpose(1, 2).contiguous()
    out = torch.full((rows, heads * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(L.lib.cra5_op_attention(L.ptr(q), L.ptr(k), L.ptr(vt), L.ptr(out), heads * 64, heads, rows, seg_len,
                                    L.stream_ptr()))
    torch.cuda.synchronize()
    ref = torch.empty(rows, heads * 64, device="cuda")
    for s in range(nseg):
        sl = slice(s * seg_len, (s + 1) * seg_len)
        r = _attention_ref(q[:, sl].float(), k[:, sl].float(), v[:, sl].float())  # [H, S, 64]
        ref[sl] = r.permute(1, 0, 2).reshape(seg_len, heads * 64)
    err = (out.float() - ref).abs().max().item()
    # P and the output are rounded to bf16 (2^-8 relative); values are O(1)
    assert err <= 2.5e-2, err
    assert (out.float() - ref).abs().mean().item() <= 2e-3


def test_attention_is_deterministic_and_handles_peaked_rows():
    """run-to-run bit equality (the decoder re-runs the hyperprior and must reproduce the encoder's floats; the trunk
    kernel is held to the same standard), and rows whose maximum jumps by far more than the lazy-rescale threshold
    between KV tiles"""
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(99)
    heads, nseg, seg_len = 4, 3, 576
    rows = nseg * seg_len
    q = (torch.randn(heads, rows, 64, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    k = (torch.randn(heads, rows, 64, device="cuda", generator=g) * 2.0).to(torch.bfloat16)
    # one very strong key late in every segment: the running maximum moves by ~40 in the last tile
    k[:, 500::seg_len] = q[:, 100::seg_len] * 16
    v = torch.randn(heads, rows, 64, device="cuda", generator=g).to(torch.bfloat16)
    vt = v.transpose(1, 2).contiguous()
    outs = []
    for _ in range(3):
        out = torch.full((rows, heads * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        L.check(L.lib.cra5_op_attention(L.ptr(q), L.ptr(k), L.ptr(vt), L.ptr(out), heads * 64, heads, rows, seg_len,
                                        L.stream_ptr()))
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    ref = torch.empty(rows, heads * 64, device="cuda")
    for s_ in range(nseg):
        sl = slice(s_ * seg_len, (s_ + 1) * seg_len)
        r = _attention_ref(q[:, sl].float(), k[:, sl].float(), v[:, sl].float())
        ref[sl] = r.permute(1, 0, 2).reshape(seg_len, heads * 64)
    assert torch.isfinite(outs[0].float()).all()
    assert (outs[0].float() - ref).abs().max().item() <= 3e-2


@pytest.mark.parametrize("heads,hd,nseg,seg_len", [(5, 72, 1, 648), (2, 24, 1, 648), (2, 24, 1, 32), (3, 72, 2, 100),
                                                   (2, 80, 1, 4000)])
def test_generic_attention_matches_fp32_reference(heads, hd, nseg, seg_len):
    """hyperprior attention (h_a / h_s: 648 tokens, 5 heads x 72) -- fp32 SIMT kernels, bf16 operands"""
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(heads * 100 + seg_len + hd)
    rows = nseg * seg_len
    q = (torch.randn(heads, rows, hd, device="cuda", generator=g) * hd ** -0.5 * 2.0).to(torch.bfloat16)
    k = (torch.randn(heads, rows, hd, device="cuda", generator=g) * 2.0).to(torch.bfloat16)
    v = torch.randn(heads, rows, hd, device="cuda", generator=g).to(torch.bfloat16)
    vt = v.transpose(1, 2).contiguous()
    out = torch.full((rows, heads * hd), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(L.lib.cra5_op_attention_generic(L.ptr(q), L.ptr(k), L.ptr(vt), L.ptr(out), heads * hd, heads, hd, rows,
                                            seg_len, L.stream_ptr()))
    torch.cuda.synchronize()
    ref = torch.empty(rows, heads * hd, device="cuda")
    for s in range(nseg):
        sl = slice(s * seg_len, (s + 1) * seg_len)
        r = _attention_ref(q[:, sl].float(), k[:, sl].float(), v[:, sl].float())
        ref[sl] = r.permute(1, 0, 2).reshape(seg_len, heads * hd)
    # fp32 accumulation throughout; P (warp-MMA path) and the output are rounded to bf16: 2^-8 relative
    assert (out.float() - ref).abs().max().item() <= 2 ** -7 * max(1.0, ref.abs().max().item())
    # the decoder re-runs h_s and must reproduce the encoder's sigma / mu bit for bit: no run-to-run variation allowed
    out2 = torch.full_like(out, float("nan"))
    L.check(L.lib.cra5_op_attention_generic(L.ptr(q), L.ptr(k), L.ptr(vt), L.ptr(out2), heads * hd, heads, hd, rows,
                                            seg_len, L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
