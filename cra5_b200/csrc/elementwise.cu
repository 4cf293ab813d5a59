// elementwise.cu -- HBM-bound layout / normalisation kernels around the tensor-core GEMMs.
//   frame_to_patches   NCHW fp32 frame -> bf16 [h][patch column][c*pw+s] (+ fused (x-mean)/std, cra5_api.py:264-266)
//   layernorm_bf16     fp32 rows -> normalised bf16 rows, optionally gathered into window-partitioned order with
//                      zero pad rows (LayerNorm vit_nlc.py:266,278 + F.pad/window_partition vit_nlc.py:229-237)
//   im2col_latent      NCHW fp32 -> bf16 patches for the hyperprior conv (vit_nlc.py:302 with k=s=4)
//   transpose_cast     [C][T] fp32 -> [T][C] bf16 (NCHW feature -> token rows, vit_nlc.py:683-684)
//   cast_bf16          fp32 -> bf16
//   affine_channels    in-place x*std+mean (cra5_api.py:268-271) and its inverse
#include "gemm_tc.cuh"
#include "host_util.h"
#include "kernels.h"
#include <algorithm>

namespace cra5 {

// split-bf16 precision mode (GemmShape::a_split): v ~ hi + lo with hi = bf16(v), lo = bf16(v - hi). Every producer of a
// GEMM A operand below takes an optional `lo` output (same layout as the bf16 output); null = plain bf16.
__device__ __forceinline__ __nv_bfloat16 bf16_lo(float v) {
  return __float2bfloat16(v - __bfloat162float(__float2bfloat16(v)));
}

// ------------------------------------------------------------------------------------------------ frame_to_patches
// grid: (Wp / TOK, H, ceil(C / CH)); block 256. Stages a [CH][TOK*pw] tile through shared memory so that both the
// NCHW reads (runs of TOK*pw floats) and the patch-major writes (runs of CH*pw bf16) are coalesced.
template <int TOK, int CH>
__global__ void __launch_bounds__(256) frame_to_patches_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                               __nv_bfloat16* __restrict__ out_lo,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ std_, int C, int H, int W,
                                                               int Wp, int pw, int cs_pad) {
  extern __shared__ float tile[];  // [CH][TOK*pw]
  const int j0 = blockIdx.x * TOK;
  const int h = blockIdx.y;
  const int c0 = blockIdx.z * CH;
  const int run = TOK * pw;
  const int nch = min(CH, C - c0);
  for (int e = threadIdx.x; e < nch * run; e += blockDim.x) {
    const int cl = e / run, wl = e - cl * run;
    const int c = c0 + cl;
    float v = x[((size_t)c * H + h) * W + (size_t)j0 * pw + wl];
    if (mean != nullptr) v = (v - mean[c]) / std_[c];
    tile[cl * run + wl] = v;
  }
  __syncthreads();
  const int seg = nch * pw;  // contiguous output run per token
  for (int e = threadIdx.x; e < TOK * seg; e += blockDim.x) {
    const int jl = e / seg, q = e - jl * seg;
    const int cl = q / pw, s = q - cl * pw;
    const size_t o = ((size_t)h * Wp + j0 + jl) * cs_pad + (size_t)c0 * pw + q;
    const float v = tile[cl * run + jl * pw + s];
    out[o] = __float2bfloat16(v);
    if (out_lo != nullptr) out_lo[o] = bf16_lo(v);
  }
}

// specialised for a compile-time patch width (PW = 10 for ERA5): 128-bit loads of the NCHW rows, 128-bit stores of the
// patch-major rows, constant-divisor index arithmetic.
template <int TOK, int CH, int PW>
__global__ void __launch_bounds__(256) frame_to_patches_vec_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                                   const float* __restrict__ mean,
                                                                   const float* __restrict__ std_, int C, int H, int W,
                                                                   int Wp, int cs_pad) {
  constexpr int RUN = TOK * PW;       // floats per channel row segment (80)
  constexpr int RUN4 = RUN / 4;
  constexpr int LD = RUN + 1;         // padded smem row
  __shared__ float tile[CH * LD];
  const int j0 = blockIdx.x * TOK;
  const int h = blockIdx.y;
  const int c0 = blockIdx.z * CH;
  const int nch = min(CH, C - c0);
  for (int e = threadIdx.x; e < nch * RUN4; e += blockDim.x) {
    const int cl = e / RUN4, w4 = e - cl * RUN4;
    const int c = c0 + cl;
    float4 v = *reinterpret_cast<const float4*>(x + ((size_t)c * H + h) * W + (size_t)j0 * PW + 4 * w4);
    if (mean != nullptr) {
      const float m = mean[c], sd = std_[c];
      v.x = (v.x - m) / sd; v.y = (v.y - m) / sd; v.z = (v.z - m) / sd; v.w = (v.w - m) / sd;
    }
    float* t = tile + cl * LD + 4 * w4;
    t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
  }
  __syncthreads();
  const int seg = nch * PW;           // contiguous output run per token (bf16), multiple of 2
  const int seg8 = seg / 8;
  for (int e = threadIdx.x; e < TOK * seg8; e += blockDim.x) {
    const int jl = e / seg8, q0 = (e - jl * seg8) * 8;
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int qa = q0 + 2 * k, qb = qa + 1;
      const int ca = qa / PW, sa = qa - ca * PW, cb = qb / PW, sb = qb - cb * PW;
      w[k] = pack_bf16x2(tile[ca * LD + jl * PW + sa], tile[cb * LD + jl * PW + sb]);
    }
    *reinterpret_cast<uint4*>(out + ((size_t)h * Wp + j0 + jl) * cs_pad + (size_t)c0 * PW + q0) =
        make_uint4(w[0], w[1], w[2], w[3]);
  }
  // tail of the run when nch * PW is not a multiple of 8 (last channel block of e.g. 69 channels)
  for (int e = threadIdx.x + seg8 * 8; e < seg; e += blockDim.x) {
    const int ca = e / PW, sa = e - ca * PW;
#pragma unroll
    for (int jl = 0; jl < TOK; ++jl)
      out[((size_t)h * Wp + j0 + jl) * cs_pad + (size_t)c0 * PW + e] = __float2bfloat16(tile[ca * LD + jl * PW + sa]);
  }
}

void frame_to_patches(cudaStream_t st, const float* x, __nv_bfloat16* out, const float* mean, const float* std_,
                      int C, int H, int W, int Wp, int pw, int cs_pad, __nv_bfloat16* out_lo) {
  constexpr int TOK = 8, CH = 64;
  CRA5_CHECK(Wp % TOK == 0, ERR_INVALID, "unsupported geometry: patches per row must be a multiple of 8");
  CRA5_CHECK(Wp * pw <= W, ERR_INVALID, "frame_to_patches: geometry");
  dim3 grid(Wp / TOK, H, (C + CH - 1) / CH);
  if (out_lo == nullptr && pw == 10 && (W & 3) == 0 && (cs_pad & 7) == 0) {
    LaunchScope scope(st, "frame_to_patches", 0.0, 4.0 * C * H * (double)W + 2.0 * H * (double)Wp * C * pw);
    frame_to_patches_vec_kernel<TOK, CH, 10><<<grid, 256, 0, st>>>(x, out, mean, std_, C, H, W, Wp, cs_pad);
    CRA5_CUDA(cudaGetLastError());
    return;
  }
  const size_t smem = (size_t)CH * TOK * pw * sizeof(float);
  CRA5_CHECK(smem <= 48 * 1024, ERR_INVALID, "unsupported geometry: patch width too large");
  LaunchScope scope(st, "frame_to_patches", 0.0, 4.0 * C * H * (double)W + 2.0 * H * (double)Wp * C * pw);
  frame_to_patches_kernel<TOK, CH><<<grid, 256, smem, st>>>(x, out, out_lo, mean, std_, C, H, W, Wp, pw, cs_pad);
  CRA5_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ layernorm
// one warp per OUTPUT row. D <= 32 * MAXV.
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps,
                                                             __nv_bfloat16* __restrict__ out,
                                                             __nv_bfloat16* __restrict__ out_lo, int rows_out, int D,
                                                             WinMap wm) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows_out) return;
  __nv_bfloat16* o = out + (size_t)warp * D;
  __nv_bfloat16* ol = out_lo != nullptr ? out_lo + (size_t)warp * D : nullptr;
  const int t = wm.to_token(warp);
  if (t < 0) {  // zero pad token (F.pad after the norm, vit_nlc.py:233)
    for (int i = lane; i < D; i += 32) {
      o[i] = __float2bfloat16(0.f);
      if (ol != nullptr) ol[i] = __float2bfloat16(0.f);
    }
    return;
  }
  const float* r = x + (size_t)t * D;
  float v[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    v[i] = (idx < D) ? r[idx] : 0.f;
    sum += v[i];
  }
#pragma unroll
  for (int o_ = 16; o_ > 0; o_ >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o_);
  const float mean = sum / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    const float d = (idx < D) ? (v[i] - mean) : 0.f;
    sq += d * d;
  }
#pragma unroll
  for (int o_ = 16; o_ > 0; o_ >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o_);
  const float rstd = rsqrtf(sq / (float)D + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < D) {
      const float r_ = (v[i] - mean) * rstd * gamma[idx] + beta[idx];
      o[idx] = __float2bfloat16(r_);
      if (ol != nullptr) ol[idx] = bf16_lo(r_);
    }
  }
}

// vectorised variant for D % 128 == 0 (the 1024-wide trunk): each lane owns float4 groups at columns 4*lane + 128*i, so
// every load is a 512-byte warp transaction and every store 256 bytes.
template <int NV>  // NV = D / 128 float4 groups per lane
__global__ void __launch_bounds__(256) layernorm_bf16_vec_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float eps,
                                                                 __nv_bfloat16* __restrict__ out, int rows_out, WinMap wm) {
  constexpr int D = NV * 128;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows_out) return;
  uint2* o = reinterpret_cast<uint2*>(out + (size_t)warp * D);
  const int t = wm.to_token(warp);
  if (t < 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) o[lane + 32 * i] = make_uint2(0u, 0u);
    return;
  }
  const float4* r = reinterpret_cast<const float4*>(x + (size_t)t * D);
  float4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = r[lane + 32 * i];
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o_ = 16; o_ > 0; o_ >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o_);
  const float mean = sum * (1.0f / D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    sq += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int o_ = 16; o_ > 0; o_ >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o_);
  const float rstd = rsqrtf(sq * (1.0f / D) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i), b = __ldg(b4 + lane + 32 * i);
    uint2 u;
    u.x = pack_bf16x2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
    u.y = pack_bf16x2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
    o[lane + 32 * i] = u;
  }
}

void layernorm_bf16(cudaStream_t st, const float* x, const float* gamma, const float* beta, float eps,
                    __nv_bfloat16* out, int rows_out, int D, const WinMap& wm, __nv_bfloat16* out_lo) {
  const int blocks = (rows_out + 7) / 8;
  LaunchScope scope(st, "layernorm_bf16", 0.0, (out_lo != nullptr ? 8.0 : 6.0) * rows_out * (double)D);
  if (out_lo == nullptr && D == 1024) {
    layernorm_bf16_vec_kernel<8><<<blocks, 256, 0, st>>>(x, gamma, beta, eps, out, rows_out, wm);
  } else if (out_lo == nullptr && D == 128) {
    layernorm_bf16_vec_kernel<1><<<blocks, 256, 0, st>>>(x, gamma, beta, eps, out, rows_out, wm);
  } else if (D <= 128) {
    layernorm_bf16_kernel<4><<<blocks, 256, 0, st>>>(x, gamma, beta, eps, out, out_lo, rows_out, D, wm);
  } else if (D <= 512) {
    layernorm_bf16_kernel<16><<<blocks, 256, 0, st>>>(x, gamma, beta, eps, out, out_lo, rows_out, D, wm);
  } else if (D <= 1024) {
    layernorm_bf16_kernel<32><<<blocks, 256, 0, st>>>(x, gamma, beta, eps, out, out_lo, rows_out, D, wm);
  } else if (D <= 2048) {
    layernorm_bf16_kernel<64><<<blocks, 256, 0, st>>>(x, gamma, beta, eps, out, out_lo, rows_out, D, wm);
  } else {
    throw Error(ERR_INVALID, "layernorm: width > 2048 unsupported");
  }
  CRA5_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ im2col (hyper conv)
// y [C][Hy][Wy] fp32 -> A [(i,j)][(c, r, s)] bf16 with patch == stride == (p1, p2)
__global__ void im2col_latent_kernel(const float* __restrict__ y, __nv_bfloat16* __restrict__ A,
                                     __nv_bfloat16* __restrict__ A_lo, int C, int Hy, int Wy, int p1, int p2, int lda) {
  const int Wh = Wy / p2;
  const int K = C * p1 * p2;
  const size_t total = (size_t)(Hy / p1) * Wh * K;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(e % K);
    const int tok = (int)(e / K);
    const int i = tok / Wh, j = tok - i * Wh;
    const int c = k / (p1 * p2), rs = k - c * p1 * p2;
    const int r = rs / p2, s = rs - r * p2;
    const float v = y[((size_t)c * Hy + i * p1 + r) * Wy + j * p2 + s];
    A[(size_t)tok * lda + k] = __float2bfloat16(v);
    if (A_lo != nullptr) A_lo[(size_t)tok * lda + k] = bf16_lo(v);
  }
}

void im2col_latent(cudaStream_t st, const float* y, __nv_bfloat16* A, int C, int Hy, int Wy, int p1, int p2, int lda,
                   __nv_bfloat16* A_lo) {
  const size_t total = (size_t)(Hy / p1) * (Wy / p2) * C * p1 * p2;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  LaunchScope scope(st, "im2col_latent", 0.0, 6.0 * (double)total);
  im2col_latent_kernel<<<blocks, 256, 0, st>>>(y, A, A_lo, C, Hy, Wy, p1, p2, lda);
  CRA5_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ transpose + cast
// in [C][T] fp32 -> out [T][ldo] bf16 (columns 0..C-1)
__global__ void transpose_cast_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                      __nv_bfloat16* __restrict__ out_lo, int C, int T, int ldo) {
  __shared__ float tile[32][33];
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, t = t0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && t < T) ? in[(size_t)c * T + t] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int t = t0 + r, c = c0 + threadIdx.x;
    if (t < T && c < C) {
      out[(size_t)t * ldo + c] = __float2bfloat16(tile[threadIdx.x][r]);
      if (out_lo != nullptr) out_lo[(size_t)t * ldo + c] = bf16_lo(tile[threadIdx.x][r]);
    }
  }
}

void transpose_cast(cudaStream_t st, const float* in, __nv_bfloat16* out, int C, int T, int ldo, __nv_bfloat16* out_lo) {
  dim3 grid((T + 31) / 32, (C + 31) / 32), block(32, 8);
  LaunchScope scope(st, "transpose_cast", 0.0, 6.0 * C * (double)T);
  transpose_cast_kernel<<<grid, block, 0, st>>>(in, out, out_lo, C, T, ldo);
  CRA5_CUDA(cudaGetLastError());
}

__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
    out[e] = __float2bfloat16(in[e]);
}

void cast_bf16(cudaStream_t st, const float* in, __nv_bfloat16* out, size_t n) {
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
  LaunchScope scope(st, "cast_bf16", 0.0, 6.0 * (double)n);
  cast_bf16_kernel<<<blocks, 256, 0, st>>>(in, out, n);
  CRA5_CUDA(cudaGetLastError());
}

// fp32 rows -> split bf16 rows (hi, lo), optionally through the exact-erf GELU (nn.GELU default, vit_nlc.py:53): the
// A-operand producer of the split-precision mode wherever the default path fuses the bf16 cast into a GEMM epilogue
// (fc1 -> GELU -> fc2, the mean || logvar concat in front of quant_conv) or uses cast_bf16.
__global__ void split_rows_kernel(const float* __restrict__ in, int ld_in, int rows, int cols,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int ld_out, int gelu) {
  const size_t total = (size_t)rows * cols;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t r = e / cols;
    const int c = (int)(e - r * cols);
    float v = in[r * ld_in + c];
    if (gelu) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    hi[r * ld_out + c] = __float2bfloat16(v);
    if (lo != nullptr) lo[r * ld_out + c] = bf16_lo(v);
  }
}

void split_rows(cudaStream_t st, const float* in, int ld_in, int rows, int cols, __nv_bfloat16* hi, __nv_bfloat16* lo,
                int ld_out, bool gelu) {
  const size_t total = (size_t)rows * cols;
  if (total == 0) return;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  LaunchScope scope(st, "split_rows", 0.0, 8.0 * (double)total);
  split_rows_kernel<<<blocks, 256, 0, st>>>(in, ld_in, rows, cols, hi, lo, ld_out, gelu ? 1 : 0);
  CRA5_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ per-channel affine
// forward == 1: (x - a[c]) / b[c]  (cra5_api.normalization); forward == 0: x * b[c] + a[c] (de_normalization)
__global__ void affine_channels_kernel(const float* __restrict__ in, float* __restrict__ out,
                                       const float* __restrict__ a, const float* __restrict__ b, size_t hw, int C,
                                       int forward) {
  const int c = blockIdx.y;
  const float m = a[c], s = b[c];
  const float* src = in + (size_t)c * hw;
  float* dst = out + (size_t)c * hw;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < hw; e += (size_t)gridDim.x * blockDim.x) {
    const float v = src[e];
    dst[e] = forward ? (v - m) / s : v * s + m;
  }
}

void affine_channels(cudaStream_t st, const float* in, float* out, const float* a, const float* b, size_t hw, int C,
                     int forward) {
  dim3 grid((unsigned)std::min<size_t>((hw + 1023) / 1024, 64), C);
  LaunchScope scope(st, "affine_channels", 0.0, 8.0 * C * (double)hw);
  affine_channels_kernel<<<grid, 256, 0, st>>>(in, out, a, b, hw, C, forward);
  CRA5_CUDA(cudaGetLastError());
}

}  // namespace cra5
