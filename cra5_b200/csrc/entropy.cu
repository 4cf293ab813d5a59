// entropy.cu -- the entropy stage on the GPU: fused quantise + scale-index kernel, and a chunk-parallel rANS coder.
//
// Reference semantics restated (bit-exact integer work):
//   symbols  = round_half_even(y - mu).int()                   EntropyModel.quantize, entropy_models.py:167-184
//   index    = 63 - sum_{s in table[:63]} [max(sigma,0.11) <= s] GaussianConditional.build_indexes, :679-685
//   rANS     64-bit state, 32-bit renormalisation, 16-bit precision, 4-bit-nibble bypass for out-of-range symbols
//            rans_interface.cpp:108-200 (encode), :215-284 (decode); primitives as in ryg_rans rans64.h
//
// Parallel format: the reference codes a tensor as ONE sequential stream. Here every latent channel is split into
// `spc` interleaved sub-streams (sub-stream k of channel c holds symbols c*L + k, c*L + k + spc, ...); each sub-stream
// is an independent stream with EXACTLY the reference's arithmetic (so any sub-stream can be checked byte-for-byte
// against the reference coder fed the same strided symbols), coded by one thread. Lanes of a warp read neighbouring
// symbols, so symbol/index loads are sector-coalesced.
#include "host_util.h"
#include "kernels.h"
#include <algorithm>

namespace cra5 {

// ------------------------------------------------------------------------------------------------ quantise + index
// 4 elements per thread, 128-bit loads of y / sigma / mu, 128-bit store of symbols, 32-bit store of indexes.
__device__ __forceinline__ int scale_index(float sigma, const float* __restrict__ tab, int levels, float bound) {
  const float s = fmaxf(sigma, bound);  // LowerBound forward: torch.max(x, bound), bound_ops.py:35-36
  // first t in [0, levels-1) with s <= tab[t], else levels-1  (== levels-1 - #{t < levels-1 : s <= tab[t]})
  int lo = 0, hi = levels - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (s <= tab[mid]) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// The same index for a (near) log-spaced table -- get_scale_table() is exp(linspace(ln 0.11, ln 256, 64)),
// models/base.py:54-61 -- without the 6 dependent shared-memory probes: log2(s) * a + b lands within +-0.25 of the
// fractional position (checked per launch against every table entry, see the kernel prologue), so ceil() is within one
// step of the answer and two compares against the EXACT table entries finish it. Same integer as scale_index() for
// every input; the kernel was issue-bound on the search (22 instructions per element, ALU pipe 58 %, ncu round 2).
__device__ __forceinline__ int scale_index_log(float sigma, const float* __restrict__ tab, int levels, float bound,
                                               float a, float b) {
  const float s = fmaxf(sigma, bound);
  int g = __float2int_ru(fmaf(__log2f(s), a, b));
  g = min(max(g, 0), levels - 1);
  g -= (g > 0 && s <= tab[g - 1]) ? 1 : 0;
  g += (g < levels - 1 && s > tab[g]) ? 1 : 0;
  return g;
}

__global__ void __launch_bounds__(256)
gc_quantize_index_kernel(const float* __restrict__ y, const float* __restrict__ sigma, const float* __restrict__ mu,
                         const float* __restrict__ scale_table, int levels, float bound, int32_t* __restrict__ sym,
                         uint8_t* __restrict__ idx, float* __restrict__ y_hat, size_t n, int frames,
                         size_t param_stride) {
  __shared__ float tab[256];
  if (scale_table != nullptr)   // only needed for the index output; callers that want symbols / y_hat alone pass null
    for (int i = threadIdx.x; i < levels; i += blockDim.x) tab[i] = scale_table[i];
  __syncthreads();
  // is the table log-spaced (every entry within a quarter step of its predicted position)? block-uniform answer
  float la = 0.f, lb = 0.f;
  int fast = 0;
  if (idx != nullptr && levels >= 4 && tab[0] > 0.f && tab[levels - 1] > tab[0]) {
    const float l0 = __log2f(tab[0]), l1 = __log2f(tab[levels - 1]);
    la = (float)(levels - 1) / (l1 - l0);
    lb = -l0 * la;
    int ok = 1;
    for (int i = threadIdx.x; i < levels; i += blockDim.x)
      ok = ok && (fabsf(fmaf(__log2f(tab[i]), la, lb) - (float)i) < 0.25f) && (i == 0 || tab[i] > tab[i - 1]);
    fast = __syncthreads_and(ok);
  }
  const size_t n4 = n >> 2;
  // frames > 1 (host guarantees n % 4 == 0 and param_stride % 4 == 0): element group g of the batch lives at e = g in
  // the per-frame-contiguous tensors (y, sym, idx, y_hat) and at ep = frame * param_stride / 4 + g % n4 in sigma / mu
  const size_t total4 = n4 * (size_t)frames;
  // (unrolled: the loads of two iterations are independent and issue back to back -- the kernel is bound by memory-level
  // parallelism, not by the 6-step table search)
#pragma unroll 2
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total4; e += (size_t)gridDim.x * blockDim.x) {
    size_t ep = e;
    if (frames > 1) {
      const size_t f = e / n4;
      ep = f * (param_stride >> 2) + (e - f * n4);
    }
    int4 s = make_int4(0, 0, 0, 0);
    float4 mv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y != nullptr) {
      const float4 yv = reinterpret_cast<const float4*>(y)[e];
      mv = reinterpret_cast<const float4*>(mu)[ep];
      s.x = __float2int_rn(__fsub_rn(yv.x, mv.x));
      s.y = __float2int_rn(__fsub_rn(yv.y, mv.y));
      s.z = __float2int_rn(__fsub_rn(yv.z, mv.z));
      s.w = __float2int_rn(__fsub_rn(yv.w, mv.w));
      if (sym != nullptr) reinterpret_cast<int4*>(sym)[e] = s;
    }
    if (idx != nullptr) {
      const float4 sv = reinterpret_cast<const float4*>(sigma)[ep];
      uchar4 q;
      if (fast) {
        q.x = (unsigned char)scale_index_log(sv.x, tab, levels, bound, la, lb);
        q.y = (unsigned char)scale_index_log(sv.y, tab, levels, bound, la, lb);
        q.z = (unsigned char)scale_index_log(sv.z, tab, levels, bound, la, lb);
        q.w = (unsigned char)scale_index_log(sv.w, tab, levels, bound, la, lb);
      } else {
        q.x = (unsigned char)scale_index(sv.x, tab, levels, bound);
        q.y = (unsigned char)scale_index(sv.y, tab, levels, bound);
        q.z = (unsigned char)scale_index(sv.z, tab, levels, bound);
        q.w = (unsigned char)scale_index(sv.w, tab, levels, bound);
      }
      reinterpret_cast<uchar4*>(idx)[e] = q;
    }
    if (y_hat != nullptr && y != nullptr)  // "dequantize": round(y - mu) + mu, entropy_models.py:173-178
      reinterpret_cast<float4*>(y_hat)[e] = make_float4(__fadd_rn((float)s.x, mv.x), __fadd_rn((float)s.y, mv.y),
                                                       __fadd_rn((float)s.z, mv.z), __fadd_rn((float)s.w, mv.w));
  }
  // tail (single frame only)
  if (frames == 1 && blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const size_t e = (n4 << 2) + threadIdx.x;
    if (y != nullptr) {
      const int s = __float2int_rn(__fsub_rn(y[e], mu[e]));
      if (sym != nullptr) sym[e] = s;
      if (y_hat != nullptr) y_hat[e] = __fadd_rn((float)s, mu[e]);
    }
    if (idx != nullptr) idx[e] = (uint8_t)scale_index(sigma[e], tab, levels, bound);
  }
}

void gc_quantize_index(cudaStream_t st, const float* y, const float* sigma, const float* mu, const float* scale_table,
                       int levels, float bound, int32_t* sym, uint8_t* idx, float* y_hat, size_t n, int frames,
                       size_t param_stride) {
  CRA5_CHECK(levels >= 1 && levels <= 256, ERR_INVALID, "scale table must have 1..256 levels");
  CRA5_CHECK(idx == nullptr || (scale_table != nullptr && sigma != nullptr), ERR_INVALID,
             "gc_quantize_index: the index output needs sigma and the scale table");
  CRA5_CHECK(y == nullptr || mu != nullptr, ERR_INVALID, "gc_quantize_index: y needs mu");
  CRA5_CHECK(frames >= 1, ERR_INVALID, "gc_quantize_index: frames");
  CRA5_CHECK(frames == 1 || ((n & 3) == 0 && (param_stride & 3) == 0), ERR_INVALID,
             "gc_quantize_index: batched launches need element counts that are multiples of 4");
  if (n == 0) return;
  const size_t n_all = n * (size_t)frames;
  const int blocks = (int)std::min<size_t>(((n_all >> 2) + 255) / 256 + 1, 148 * 8);
  // algorithmic bytes: read y, sigma, mu (12 B) + write int32 symbol and uint8 index (5 B) per element
  LaunchScope scope(st, "gc_quantize_index", 0.0,
                    (double)n_all * ((y ? 8.0 : 0.0) + (idx ? 5.0 : 0.0) + (sym ? 4.0 : 0.0) + (y_hat ? 4.0 : 0.0)));
  gc_quantize_index_kernel<<<blocks, 256, 0, st>>>(y, sigma, mu, scale_table, levels, bound, sym, idx, y_hat, n, frames,
                                                   param_stride);
  CRA5_CUDA(cudaGetLastError());
}

// EntropyBottleneck: per-channel median, index == channel (entropy_models.py:529-542, 513-523)
__global__ void eb_quantize_kernel(const float* __restrict__ z, const float* __restrict__ median, int L, int n_ch,
                                   int32_t* __restrict__ sym, float* __restrict__ z_hat, size_t n) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const float m = median[(e / L) % n_ch];
    const int s = __float2int_rn(__fsub_rn(z[e], m));
    if (sym != nullptr) sym[e] = s;
    if (z_hat != nullptr) z_hat[e] = __fadd_rn((float)s, m);
  }
}

void eb_quantize(cudaStream_t st, const float* z, const float* median, int L, int n_ch, int32_t* sym, float* z_hat,
                 size_t n) {
  if (n == 0) return;
  CRA5_CHECK(L > 0 && n_ch > 0, ERR_INVALID, "eb_quantize: sizes");
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
  LaunchScope scope(st, "eb_quantize", 0.0, (double)n * 12.0);
  eb_quantize_kernel<<<blocks, 256, 0, st>>>(z, median, L, n_ch, sym, z_hat, n);
  CRA5_CUDA(cudaGetLastError());
}

// sym (int32) + mean -> float; mean per element (mu != null) or per channel (median != null)
__global__ void dequantize_kernel(const int32_t* __restrict__ sym, const float* __restrict__ mu,
                                  const float* __restrict__ median, int L, float* __restrict__ out, size_t n) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const float m = (mu != nullptr) ? mu[e] : median[e / L];
    out[e] = __fadd_rn((float)sym[e], m);
  }
}

void dequantize(cudaStream_t st, const int32_t* sym, const float* mu, const float* median, int L, float* out,
                size_t n) {
  if (n == 0) return;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
  LaunchScope scope(st, "dequantize", 0.0, (double)n * 12.0);
  dequantize_kernel<<<blocks, 256, 0, st>>>(sym, mu, median, L, out, n);
  CRA5_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ likelihoods
// Rate-estimation outputs of VAEformer.forward / encode_latent('quantized') (vaeformer.py:284-290, 314-319):
//   GaussianConditional: y_hat = round(y - mu) + mu; p = Phi((.5 - |y_hat - mu|)/s) - Phi((-.5 - |y_hat - mu|)/s),
//                        s = max(sigma, 0.11), Phi(x) = erfc(-x/sqrt 2)/2, floored at 1e-9   (entropy_models.py:645-677)
__global__ void __launch_bounds__(256)
gc_likelihood_kernel(const float* __restrict__ y, const float* __restrict__ sigma, const float* __restrict__ mu,
                     float scale_bound, float lik_bound, float* __restrict__ y_hat, float* __restrict__ lik, size_t n) {
  const float c = -0.70710678118654752440f;  // float(-(2 ** -0.5)), entropy_models.py:600
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const float m = mu[e];
    const float yh = __fadd_rn(rintf(__fsub_rn(y[e], m)), m);
    const float v = fabsf(__fsub_rn(yh, m));
    const float sc = fmaxf(sigma[e], scale_bound);
    const float upper = 0.5f * erfcf(c * __fdiv_rn(__fsub_rn(0.5f, v), sc));
    const float lower = 0.5f * erfcf(c * __fdiv_rn(__fsub_rn(-0.5f, v), sc));
    if (y_hat != nullptr) y_hat[e] = yh;
    lik[e] = fmaxf(__fsub_rn(upper, lower), lik_bound);
  }
}

void gc_likelihood(cudaStream_t st, const float* y, const float* sigma, const float* mu, float scale_bound,
                   float lik_bound, float* y_hat, float* lik, size_t n) {
  if (n == 0) return;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
  LaunchScope scope(st, "gc_likelihood", 0.0, (double)n * 20.0);
  gc_likelihood_kernel<<<blocks, 256, 0, st>>>(y, sigma, mu, scale_bound, lik_bound, y_hat, lik, n);
  CRA5_CUDA(cudaGetLastError());
}

//   EntropyBottleneck: per-channel factorised density, p = sigmoid(F(z+.5)) - sigmoid(F(z-.5)) with the 1-3-3-3-3-1
//   softplus/tanh cumulative F (entropy_models.py:434-463). `packed` holds, per channel, softplus(matrix_k), bias_k and
//   tanh(factor_k) of the five layers: m0[3] b0[3] f0[3] | (m[9] b[3] f[3]) x3 | m4[3] b4[1]  = 58 floats.
constexpr int EB_PACK = 58;
__device__ __forceinline__ float eb_logits(const float* __restrict__ p, float v) {
  float a[3], b[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float l = fmaf(p[i], v, p[3 + i]);
    a[i] = fmaf(p[6 + i], tanhf(l), l);
  }
  p += 9;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float l = fmaf(p[3 * i + 2], a[2], fmaf(p[3 * i + 1], a[1], p[3 * i] * a[0])) + p[9 + i];
      b[i] = fmaf(p[12 + i], tanhf(l), l);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) a[i] = b[i];
    p += 15;
  }
  return fmaf(p[2], a[2], fmaf(p[1], a[1], p[0] * a[0])) + p[3];
}

__global__ void __launch_bounds__(256)
eb_likelihood_kernel(const float* __restrict__ z_hat, const float* __restrict__ packed, int L, float lik_bound,
                     float* __restrict__ lik, size_t n) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const float* p = packed + (e / L) * EB_PACK;
    const float v = z_hat[e];
    const float lo = eb_logits(p, v - 0.5f), up = eb_logits(p, v + 0.5f);
    const float s_up = 1.0f / (1.0f + expf(-up)), s_lo = 1.0f / (1.0f + expf(-lo));
    lik[e] = fmaxf(s_up - s_lo, lik_bound);
  }
}

void eb_likelihood(cudaStream_t st, const float* z_hat, const float* packed, int L, float lik_bound, float* lik,
                   size_t n) {
  if (n == 0) return;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
  LaunchScope scope(st, "eb_likelihood", 0.0, (double)n * 8.0);
  eb_likelihood_kernel<<<blocks, 256, 0, st>>>(z_hat, packed, L, lik_bound, lik, n);
  CRA5_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ rANS primitives
constexpr uint32_t RANS_PRECISION = 16;      // rans_interface.cpp:49
constexpr uint32_t RANS_BYPASS_BITS = 4;     // rans_interface.cpp:51
constexpr int32_t RANS_BYPASS_MAX = 15;      // rans_interface.cpp:52
constexpr uint64_t RANS_L = 1ull << 31;      // rans64.h
constexpr int RANS_LUT = 257;                // coarse inverse-CDF entries per row (cum >> 8, plus the end sentinel)

struct RansEnc {
  uint64_t x;
  uint32_t* ptr;      // next word is written at --ptr
  uint32_t* floor_;   // lowest address we may write
  bool overflow;
  __device__ __forceinline__ void emit() {
    if (ptr > floor_) { --ptr; *ptr = (uint32_t)x; } else overflow = true;
    x >>= 32;
  }
  // x = C(s, x) = floor(x / freq) * 2^16 + x mod freq + start. The 64-bit quotient is taken from a double-precision
  // estimate (x < 2^63, quotient < 2^47, so the estimate is within 1 of the truth) and corrected exactly.
  __device__ __forceinline__ void put(uint32_t start, uint32_t freq, double rcp) {
    const uint64_t x_max = (uint64_t)freq << (31 - RANS_PRECISION + 32);  // ((L >> 16) << 32) * freq
    if (x >= x_max) emit();
    uint64_t q = __double2ull_rz(__ull2double_rn(x) * rcp);
    int64_t r = (int64_t)(x - q * freq);
    if (r < 0) { --q; r += freq; } else if (r >= (int64_t)freq) { ++q; r -= freq; }
    x = (q << RANS_PRECISION) + (uint64_t)r + start;
  }
  __device__ __forceinline__ void put_bits(uint32_t val, uint32_t nbits) {  // rans_interface.cpp:69-87
    const uint64_t x_max = ((RANS_L >> 16) << 32) * (uint64_t)(1u << (16 - nbits));
    if (x >= x_max) emit();
    x = (x << nbits) | val;
  }
};

constexpr int RANS_BATCH = 8;      // symbols gathered (independent loads) before the serial state updates
constexpr int RANS_THREADS = 32;   // one warp per CTA: the chains are latency bound, spread them over all SMs

// thread = sub-stream. Symbols are consumed last-to-first (the reference buffers all symbols and encodes the list in
// reverse, rans_interface.cpp:183-193), words are written backwards into the stream's private scratch window.
// Per batch: phase A gathers (start, freq, 1/freq, raw) for RANS_BATCH symbols with independent loads; phase B runs
// the serial state recurrence out of registers.
__global__ void __launch_bounds__(RANS_THREADS)
rans_encode_kernel(const int32_t* __restrict__ sym, const uint8_t* __restrict__ idx, int index_is_channel,
                   const int32_t* __restrict__ cdf, int cdf_stride, const int32_t* __restrict__ cdf_len,
                   const int32_t* __restrict__ offset, int n_channels, int L, int spc, int chan_len,
                   uint32_t* __restrict__ scratch, int cap_words, uint32_t* __restrict__ lengths,
                   int* __restrict__ err) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_channels * spc) return;
  const int c = s / spc, k = s - c * spc;
  const int count = (L - k + spc - 1) / spc;
  uint32_t* top = scratch + (size_t)(s + 1) * cap_words;
  RansEnc e;
  e.x = RANS_L;
  e.ptr = top;
  e.floor_ = top - cap_words + 2;  // keep room for the 2-word flush
  e.overflow = false;
  const size_t base = (size_t)c * L + k;
  for (int i0 = count; i0 > 0; i0 -= RANS_BATCH) {
    // phase A: three rounds of independent, branch-free loads (index clamped, results masked in phase B)
    size_t pos[RANS_BATCH];
    int32_t sy[RANS_BATCH], ci[RANS_BATCH], off[RANS_BATCH], mxv[RANS_BATCH], val[RANS_BATCH];
    uint32_t start[RANS_BATCH], freq[RANS_BATCH], raw[RANS_BATCH];
    double rcp[RANS_BATCH];
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) {
      const int i = max(i0 - 1 - b, 0);
      pos[b] = base + (size_t)i * spc;
      sy[b] = sym[pos[b]];
      ci[b] = index_is_channel ? (int)((pos[b] / (size_t)chan_len) % (size_t)index_is_channel) : (int)idx[pos[b]];
    }
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) {
      off[b] = offset[ci[b]];
      mxv[b] = cdf_len[ci[b]] - 2;
    }
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) {
      const int32_t v = sy[b] - off[b];
      const bool neg = v < 0, big = v >= mxv[b];
      raw[b] = neg ? (uint32_t)(-2 * v - 1) : (big ? (uint32_t)(2 * (v - mxv[b])) : 0u);
      val[b] = (neg || big) ? mxv[b] : v;   // == max_value <=> bypass (value == max_value is the sentinel bin)
      const int32_t* row = cdf + (size_t)ci[b] * cdf_stride + val[b];
      start[b] = (uint32_t)row[0];
      freq[b] = (uint32_t)row[1] - start[b];
    }
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) {
      // 1/freq to double precision without the slow-path call: float seed + two Newton steps
      const double f = (double)freq[b];
      double r = (double)__frcp_rn((float)freq[b]);
      r = r * __fma_rn(-f, r, 2.0);
      r = r * __fma_rn(-f, r, 2.0);
      rcp[b] = r;
    }
    // phase B: the serial state recurrence, out of registers
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) {
      if (i0 - 1 - b >= 0) {
        if (val[b] == mxv[b]) {
          int32_t nb = 0;
          while (nb < 8 && (raw[b] >> (nb * RANS_BYPASS_BITS)) != 0) ++nb;
          for (int32_t j = nb - 1; j >= 0; --j)
            e.put_bits((raw[b] >> (j * RANS_BYPASS_BITS)) & RANS_BYPASS_MAX, RANS_BYPASS_BITS);
          // the count is sent as [15]*q then (nb - 15q); nb <= 8 for 32-bit raw values, so q == 0
          e.put_bits((uint32_t)nb, RANS_BYPASS_BITS);
        }
        e.put(start[b], freq[b], rcp[b]);
      }
    }
  }
  // flush: stream begins with low32(x), high32(x)
  e.ptr -= 2;
  e.ptr[0] = (uint32_t)e.x;
  e.ptr[1] = (uint32_t)(e.x >> 32);
  lengths[s] = (uint32_t)(top - e.ptr) * 4u;
  if (e.overflow) atomicExch(err, 1);
}

// single-block exclusive scan of up to a few thousand stream lengths (bytes); writes offsets[n+1]
__global__ void __launch_bounds__(1024) scan_lengths_kernel(const uint32_t* __restrict__ lengths, int n,
                                                            uint32_t* __restrict__ offsets) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    uint32_t v = (i < n) ? lengths[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = warp_sums[threadIdx.x];
      uint32_t wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (threadIdx.x >= o) wi += t;
      }
      warp_sums[threadIdx.x] = wi - w;  // exclusive
    }
    __syncthreads();
    const uint32_t excl = carry + warp_sums[threadIdx.x >> 5] + incl - v;
    if (i < n) offsets[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = carry;
}

// one warp per stream: copy its words from the scratch window to the packed payload
__global__ void __launch_bounds__(256) compact_streams_kernel(const uint32_t* __restrict__ scratch, int cap_words,
                                                              const uint32_t* __restrict__ lengths,
                                                              const uint32_t* __restrict__ offsets, int n_streams,
                                                              uint8_t* __restrict__ payload) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= n_streams) return;
  const uint32_t words = lengths[s] >> 2;
  const uint32_t* src = scratch + (size_t)(s + 1) * cap_words - words;
  uint32_t* dst = reinterpret_cast<uint32_t*>(payload + offsets[s]);
  for (uint32_t w = lane; w < words; w += 32) dst[w] = src[w];
}

void rans_encode(cudaStream_t st, const int32_t* sym, const uint8_t* idx, int chan_mod, const int32_t* cdf,
                 int cdf_stride, const int32_t* cdf_len, const int32_t* offset, int n_channels, int L, int spc, int chan_len,
                 uint32_t* scratch, int cap_words, uint32_t* lengths, uint32_t* offsets, uint8_t* payload, int* err) {
  const int n_streams = n_channels * spc;
  if (n_streams == 0) return;
  {
    LaunchScope scope(st, "rans_encode", 0.0, (double)n_channels * L * (chan_mod ? 4.0 : 5.0));
    rans_encode_kernel<<<(n_streams + RANS_THREADS - 1) / RANS_THREADS, RANS_THREADS, 0, st>>>(sym, idx, chan_mod, cdf, cdf_stride,
                                                                cdf_len, offset, n_channels, L, spc, chan_len,
                                                                scratch, cap_words, lengths, err);
  }
  CRA5_CUDA(cudaGetLastError());
  {
    LaunchScope scope(st, "scan_lengths", 0.0, 8.0 * n_streams);
    scan_lengths_kernel<<<1, 1024, 0, st>>>(lengths, n_streams, offsets);
  }
  CRA5_CUDA(cudaGetLastError());
  LaunchScope scope(st, "compact_streams", 0.0, 0.0);
  compact_streams_kernel<<<(n_streams * 32 + 255) / 256, 256, 0, st>>>(scratch, cap_words, lengths, offsets, n_streams,
                                                                       payload);
  CRA5_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ decode
struct RansDec {
  uint64_t x;
  const uint32_t* ptr;
  const uint32_t* end;
  bool underflow;
  __device__ __forceinline__ uint32_t next() {
    if (ptr < end) return *ptr++;
    underflow = true;
    return 0u;
  }
  __device__ __forceinline__ uint32_t get_bits(uint32_t nbits) {  // rans_interface.cpp:89-105
    const uint32_t val = (uint32_t)(x & ((1u << nbits) - 1));
    x >>= nbits;
    if (x < RANS_L) x = (x << 32) | next();
    return val;
  }
};

// Coarse inverse-CDF table: lut[row][b] = last v with cdf[row][v] <= 256*b, b = 0..256 (one thread per entry).
__global__ void build_decode_lut_kernel(const int32_t* __restrict__ cdf, int cdf_stride,
                                        const int32_t* __restrict__ cdf_len, int rows, uint16_t* __restrict__ lut) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * RANS_LUT) return;
  const int r = e / RANS_LUT, b = e - r * RANS_LUT;
  const int32_t* row = cdf + (size_t)r * cdf_stride;
  const int32_t n_entries = cdf_len[r];
  const uint32_t target = (uint32_t)b << 8;
  int lo = 0, hi = n_entries - 1;  // row[0] = 0 <= target; row[n_entries-1] = 65536
  if (target >= 65536u) {
    lo = n_entries - 2;
  } else {
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((uint32_t)row[mid] <= target) lo = mid; else hi = mid;
    }
  }
  lut[e] = (uint16_t)lo;
}

void build_decode_lut(cudaStream_t st, const int32_t* cdf, int cdf_stride, const int32_t* cdf_len, int rows,
                      uint16_t* lut) {
  LaunchScope scope(st, "build_decode_lut", 0.0, 0.0);
  build_decode_lut_kernel<<<(rows * RANS_LUT + 255) / 256, 256, 0, st>>>(cdf, cdf_stride, cdf_len, rows, lut);
  CRA5_CUDA(cudaGetLastError());
}

// thread = sub-stream. Writes int32 symbols and/or the dequantised value sym + mean.
// The reference finds the symbol by a linear scan of the CDF row (rans_interface.cpp:246-250); here the coarse
// inverse table (staged in shared memory when it fits) narrows the range to a few entries and a short binary search
// finishes -- same result, ~3 dependent loads instead of up to 3133.
__global__ void __launch_bounds__(RANS_THREADS)
rans_decode_kernel(const uint8_t* __restrict__ payload, const uint32_t* __restrict__ offsets,
                   const uint8_t* __restrict__ idx, int index_is_channel, const int32_t* __restrict__ cdf,
                   int cdf_stride, const int32_t* __restrict__ cdf_len, const int32_t* __restrict__ offset,
                   const uint16_t* __restrict__ lut_g, int lut_rows, int n_channels, int L, int spc, int chan_len,
                   int32_t* __restrict__ sym_out, const float* __restrict__ mu, const float* __restrict__ median,
                   float* __restrict__ val_out, int* __restrict__ err, int ch_per_frame,
                   size_t mu_frame_extra) {
  extern __shared__ uint16_t lut_s[];
  const uint16_t* lut = lut_g;
  if (lut_g != nullptr && lut_rows > 0) {  // lut_rows > 0: stage the table in shared memory
    for (int e = threadIdx.x; e < lut_rows * RANS_LUT; e += blockDim.x) lut_s[e] = lut_g[e];
    __syncthreads();
    lut = lut_s;
  }
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_channels * spc) return;
  const int c = s / spc, k = s - c * spc;
  const int count = (L - k + spc - 1) / spc;
  RansDec d;
  d.ptr = reinterpret_cast<const uint32_t*>(payload + offsets[s]);
  d.end = reinterpret_cast<const uint32_t*>(payload + offsets[s + 1]);
  d.underflow = false;
  {
    const uint32_t lo = d.next(), hi = d.next();
    d.x = (uint64_t)lo | ((uint64_t)hi << 32);
  }
  const size_t mu_extra = (size_t)(c / ch_per_frame) * mu_frame_extra;   // batch: per-frame mu blocks are further apart
  const size_t base = (size_t)c * L + k;
  for (int i0 = 0; i0 < count; i0 += RANS_BATCH) {
    int ci[RANS_BATCH];
    float mean[RANS_BATCH];
    int32_t outv[RANS_BATCH];
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) {  // independent loads first
      const int i = min(i0 + b, count - 1);
      const size_t pos = base + (size_t)i * spc;
      ci[b] = index_is_channel ? (int)((pos / (size_t)chan_len) % (size_t)index_is_channel) : (int)idx[pos];
      mean[b] = (val_out == nullptr) ? 0.f : ((mu != nullptr) ? mu[pos + mu_extra] : median[ci[b]]);
    }
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) {
      outv[b] = 0;
      if (i0 + b >= count) break;
      const int32_t* row = cdf + (size_t)ci[b] * cdf_stride;
      const int32_t n_entries = cdf_len[ci[b]];
      const int32_t max_value = n_entries - 2;
      const uint32_t cum = (uint32_t)(d.x & 0xffffu);
      // last v in [0, n_entries-1) with row[v] <= cum
      int lo, hi;
      if (lut != nullptr) {
        const uint16_t* lr = lut + (size_t)ci[b] * RANS_LUT + (cum >> 8);
        lo = lr[0];
        hi = lr[1];
      } else {
        lo = 0;
        hi = n_entries - 2;
      }
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((uint32_t)row[mid] <= cum) lo = mid; else hi = mid - 1;
      }
      const uint32_t start = (uint32_t)row[lo];
      const uint32_t freq = (uint32_t)row[lo + 1] - start;
      d.x = (uint64_t)freq * (d.x >> RANS_PRECISION) + cum - start;
      if (d.x < RANS_L) d.x = (d.x << 32) | d.next();
      int32_t value = lo;
      if (value == max_value) {  // bypass, rans_interface.cpp:256-278
        int32_t val = (int32_t)d.get_bits(RANS_BYPASS_BITS);
        int32_t nb = val;
        while (val == RANS_BYPASS_MAX && !d.underflow) {
          val = (int32_t)d.get_bits(RANS_BYPASS_BITS);
          nb += val;
        }
        uint32_t raw = 0;
        for (int32_t j = 0; j < nb; ++j) {
          val = (int32_t)d.get_bits(RANS_BYPASS_BITS);
          if (j < 8) raw |= (uint32_t)val << (j * RANS_BYPASS_BITS);
          if (d.underflow) break;
        }
        value = (int32_t)(raw >> 1);
        if (raw & 1u) value = -value - 1; else value += max_value;
      }
      outv[b] = value + offset[ci[b]];
    }
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) {
      const int i = i0 + b;
      if (i < count) {
        const size_t pos = base + (size_t)i * spc;
        if (sym_out != nullptr) sym_out[pos] = outv[b];
        if (val_out != nullptr) val_out[pos] = __fadd_rn((float)outv[b], mean[b]);
      }
    }
  }
  if (d.underflow) atomicExch(err, 2);
}

void rans_decode(cudaStream_t st, const uint8_t* payload, const uint32_t* offsets, const uint8_t* idx,
                 int chan_mod, const int32_t* cdf, int cdf_stride, const int32_t* cdf_len,
                 const int32_t* offset, const uint16_t* lut, int lut_rows, int n_channels, int L, int spc,
                 int chan_len, int32_t* sym_out, const float* mu, const float* median, float* val_out, int* err, int ch_per_frame,
                 size_t mu_frame_extra) {
  const int n_streams = n_channels * spc;
  if (n_streams == 0) return;
  size_t smem = 0;
  int stage_rows = 0;
  if (lut != nullptr && (size_t)lut_rows * RANS_LUT * 2 <= 48 * 1024) {
    stage_rows = lut_rows;
    smem = (size_t)lut_rows * RANS_LUT * 2;
  }
  LaunchScope scope(st, "rans_decode", 0.0, (double)n_channels * L * (chan_mod ? 4.0 : 9.0));
  rans_decode_kernel<<<(n_streams + RANS_THREADS - 1) / RANS_THREADS, RANS_THREADS, smem, st>>>(
      payload, offsets, idx, chan_mod, cdf, cdf_stride, cdf_len, offset, lut, stage_rows, n_channels, L,
      spc, chan_len, sym_out, mu, median, val_out, err, ch_per_frame > 0 ? ch_per_frame : (1 << 30), mu_frame_extra);
  CRA5_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ shared-memory tables
// The serial chain of a sub-stream is bound by the latency of its table lookups, and the int32 CDF matrix (64 x 3133
// for the GaussianConditional) only fits L2. Packed row after row as uint16 it is 27 256 entries = 54 KB (the final
// 65536 of each row wraps to 0 and is recovered from `freq = (next - start) & 0xffff`), so every CTA keeps the whole
// table -- plus the coarse inverse table for decoding -- in shared memory and a lookup costs ~30 cycles instead of an
// L2 round trip. Stream words are prefetched one ahead, index / mean loads one batch ahead.
__global__ void __launch_bounds__(256)
pack_cdf_kernel(const int32_t* __restrict__ cdf, int cdf_stride, const int32_t* __restrict__ cdf_len, int rows,
                int32_t* __restrict__ row_off, uint16_t* __restrict__ packed, int cap, int* __restrict__ total_out) {
  __shared__ int total_s;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int r = 0; r < rows; ++r) { row_off[r] = acc; acc += cdf_len[r]; }
    row_off[rows] = acc;
    total_s = acc;
    *total_out = acc;
  }
  __syncthreads();
  if (total_s > cap) return;
  for (int r = 0; r < rows; ++r) {
    const int n = cdf_len[r], o = row_off[r];
    for (int v = threadIdx.x; v < n; v += blockDim.x) packed[o + v] = (uint16_t)cdf[(size_t)r * cdf_stride + v];
  }
}

void pack_cdf(cudaStream_t st, const int32_t* cdf, int cdf_stride, const int32_t* cdf_len, int rows, int32_t* row_off,
              uint16_t* packed, int cap, int* total_dev) {
  LaunchScope scope(st, "pack_cdf", 0.0, 0.0);
  pack_cdf_kernel<<<1, 256, 0, st>>>(cdf, cdf_stride, cdf_len, rows, row_off, packed, cap, total_dev);
  CRA5_CUDA(cudaGetLastError());
}

constexpr int RANS_STHREADS = 64;   // encoder: two warps per CTA, each with an SM sub-partition to itself
// decoder: four warps share one staged table (CDF rows + inverse table, ~87 KB for the GaussianConditional: two CTAs per
// SM). With 64 threads the 512 CTAs of an 8-frame batch were 1.7 waves over the 296 resident slots -- a half-empty
// second pass over 648-symbol chains; with 128 they are one wave.
constexpr int RANS_DTHREADS = 128;

struct SmemTables {
  const uint16_t* cdf16;   // [total] packed rows
  const int32_t* row_off;  // [rows]
  const int32_t* len;      // [rows]
  const int32_t* off;      // [rows]
  const uint16_t* lut;     // [rows][RANS_LUT] or null
};
// layout: cdf16 (total_pad uint16) | lut (rows * RANS_LUT uint16, optional, padded) | row_off | len | off (int32 each)
__host__ __device__ inline size_t smem_tables_bytes(int rows, int total, bool with_lut) {
  size_t b = (((size_t)total + 7) & ~size_t(7)) * 2;
  if (with_lut) b += (((size_t)rows * RANS_LUT + 7) & ~size_t(7)) * 2;
  return b + (size_t)rows * 12;
}
__device__ __forceinline__ SmemTables stage_tables(uint8_t* smem, const uint16_t* __restrict__ packed,
                                                   const int32_t* __restrict__ row_off, const int32_t* __restrict__ cdf_len,
                                                   const int32_t* __restrict__ offset, const uint16_t* __restrict__ lut_g,
                                                   int rows, int total) {
  const int total_pad = (total + 7) & ~7;
  uint16_t* c16 = reinterpret_cast<uint16_t*>(smem);
  uint16_t* lut = c16 + total_pad;
  const int lut_pad = (lut_g != nullptr) ? ((rows * RANS_LUT + 7) & ~7) : 0;
  int32_t* ro = reinterpret_cast<int32_t*>(lut + lut_pad);
  int32_t* ln = ro + rows;
  int32_t* of = ln + rows;
  // 16-byte copies: the packed buffers are allocated with 8-entry slack
  const uint4* src = reinterpret_cast<const uint4*>(packed);
  for (int e = threadIdx.x; e < total_pad / 8; e += blockDim.x) reinterpret_cast<uint4*>(c16)[e] = src[e];
  if (lut_g != nullptr) {
    const uint4* ls = reinterpret_cast<const uint4*>(lut_g);
    for (int e = threadIdx.x; e < lut_pad / 8; e += blockDim.x) reinterpret_cast<uint4*>(lut)[e] = ls[e];
  }
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    ro[r] = row_off[r];
    ln[r] = cdf_len[r];
    of[r] = offset[r];
  }
  __syncthreads();
  SmemTables t;
  t.cdf16 = c16; t.row_off = ro; t.len = ln; t.off = of; t.lut = (lut_g != nullptr) ? lut : nullptr;
  return t;
}

__global__ void __launch_bounds__(RANS_STHREADS)
rans_encode_smem_kernel(const int32_t* __restrict__ sym, const uint8_t* __restrict__ idx, int index_is_channel,
                        const uint16_t* __restrict__ packed, const int32_t* __restrict__ row_off,
                        const int32_t* __restrict__ cdf_len, const int32_t* __restrict__ offset, int rows, int total,
                        int n_channels, int L, int spc, int chan_len, uint32_t* __restrict__ scratch, int cap_words,
                        uint32_t* __restrict__ lengths, int* __restrict__ err) {
  extern __shared__ __align__(16) uint8_t tab_smem[];
  const SmemTables T = stage_tables(tab_smem, packed, row_off, cdf_len, offset, nullptr, rows, total);
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_channels * spc) return;
  const int c = s / spc, k = s - c * spc;
  const int count = (L - k + spc - 1) / spc;
  uint32_t* top = scratch + (size_t)(s + 1) * cap_words;
  RansEnc e;
  e.x = RANS_L;
  e.ptr = top;
  e.floor_ = top - cap_words + 2;
  e.overflow = false;
  const size_t base = (size_t)c * L + k;
  int32_t sy[RANS_BATCH], ci[RANS_BATCH];
  auto gather = [&](int i0) {   // symbols / indexes of the batch that ends at symbol i0 - 1 (consumed last to first)
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) {
      const int i = max(i0 - 1 - b, 0);
      const size_t pos = base + (size_t)i * spc;
      sy[b] = sym[pos];
      ci[b] = index_is_channel ? (c % index_is_channel) : (int)idx[pos];
    }
  };
  if (count > 0) gather(count);   // (a sub-stream beyond a short channel's end is empty: nothing of its own to read)
  for (int i0 = count; i0 > 0; i0 -= RANS_BATCH) {
    int32_t mxv[RANS_BATCH], val[RANS_BATCH];
    uint32_t start[RANS_BATCH], freq[RANS_BATCH], raw[RANS_BATCH];
    double rcp[RANS_BATCH];
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) {
      const int r = ci[b];
      mxv[b] = T.len[r] - 2;
      const int32_t v = sy[b] - T.off[r];
      const bool neg = v < 0, big = v >= mxv[b];
      raw[b] = neg ? (uint32_t)(-2 * v - 1) : (big ? (uint32_t)(2 * (v - mxv[b])) : 0u);
      val[b] = (neg || big) ? mxv[b] : v;
      const uint16_t* row = T.cdf16 + T.row_off[r] + val[b];
      start[b] = row[0];
      freq[b] = ((uint32_t)row[1] - start[b]) & 0xffffu;
      if (freq[b] == 0u) freq[b] = 65536u;
    }
    if (i0 - RANS_BATCH > 0) gather(i0 - RANS_BATCH);   // next batch's global loads fly during the serial phase
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) {
      const double f = (double)freq[b];
      double r = (double)__frcp_rn((float)freq[b]);
      r = r * __fma_rn(-f, r, 2.0);
      r = r * __fma_rn(-f, r, 2.0);
      rcp[b] = r;
    }
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) {
      if (i0 - 1 - b >= 0) {
        if (val[b] == mxv[b]) {
          int32_t nb = 0;
          while (nb < 8 && (raw[b] >> (nb * RANS_BYPASS_BITS)) != 0) ++nb;
          for (int32_t j = nb - 1; j >= 0; --j)
            e.put_bits((raw[b] >> (j * RANS_BYPASS_BITS)) & RANS_BYPASS_MAX, RANS_BYPASS_BITS);
          e.put_bits((uint32_t)nb, RANS_BYPASS_BITS);
        }
        e.put(start[b], freq[b], rcp[b]);
      }
    }
  }
  e.ptr -= 2;
  e.ptr[0] = (uint32_t)e.x;
  e.ptr[1] = (uint32_t)(e.x >> 32);
  lengths[s] = (uint32_t)(top - e.ptr) * 4u;
  if (e.overflow) atomicExch(err, 1);
}

struct RansDecP {   // decoder state; the caller peeks the next stream word at the top of each symbol
  uint64_t x;
  const uint32_t* ptr;
  const uint32_t* end;
  bool underflow;
  __device__ __forceinline__ void init(const uint32_t* p, const uint32_t* e) { ptr = p; end = e; underflow = false; }
  // the address does not depend on the coder state of this symbol, so the load overlaps the state update; it lands in
  // a fresh register (a conditional refill of a persistent look-ahead register makes the compiler wait on the load at
  // the end of the conditional block)
  __device__ __forceinline__ uint32_t peek() const { return (ptr < end) ? __ldg(ptr) : 0u; }
  __device__ __forceinline__ void renorm(uint32_t peeked) {   // x < L: take the peeked word
    if (ptr >= end) underflow = true;
    x = (x << 32) | peeked;
    ++ptr;
  }
  __device__ __forceinline__ uint32_t next() {
    if (ptr < end) return __ldg(ptr++);
    underflow = true;
    return 0u;
  }
  __device__ __forceinline__ uint32_t get_bits(uint32_t nbits) {
    const uint32_t val = (uint32_t)(x & ((1u << nbits) - 1));
    x >>= nbits;
    if (x < RANS_L) x = (x << 32) | next();
    return val;
  }
};

__global__ void __launch_bounds__(RANS_DTHREADS)
rans_decode_smem_kernel(const uint8_t* __restrict__ payload, const uint32_t* __restrict__ offsets,
                        const uint8_t* __restrict__ idx, int index_is_channel, const uint16_t* __restrict__ packed,
                        const int32_t* __restrict__ row_off, const int32_t* __restrict__ cdf_len,
                        const int32_t* __restrict__ offset, const uint16_t* __restrict__ lut_g, int rows, int total,
                        int n_channels, int L, int spc, int chan_len, int32_t* __restrict__ sym_out,
                        const float* __restrict__ mu, const float* __restrict__ median, float* __restrict__ val_out,
                        int* __restrict__ err, int ch_per_frame, size_t mu_frame_extra) {
  extern __shared__ __align__(16) uint8_t tab_smem[];
  const SmemTables T = stage_tables(tab_smem, packed, row_off, cdf_len, offset, lut_g, rows, total);
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_channels * spc) return;
  const int c = s / spc, k = s - c * spc;
  const int count = (L - k + spc - 1) / spc;
  RansDecP d;
  d.init(reinterpret_cast<const uint32_t*>(payload + offsets[s]), reinterpret_cast<const uint32_t*>(payload + offsets[s + 1]));
  {
    const uint32_t lo = d.next(), hi = d.next();
    d.x = (uint64_t)lo | ((uint64_t)hi << 32);
  }
  const size_t base = (size_t)c * L + k;
  const float med = (val_out != nullptr && mu == nullptr) ? median[index_is_channel ? c % index_is_channel : c] : 0.f;
  const size_t mu_extra = (size_t)(c / ch_per_frame) * mu_frame_extra;   // batch: per-frame mu blocks are further apart
  int ci[RANS_BATCH], ci_n[RANS_BATCH];
  float mean[RANS_BATCH], mean_n[RANS_BATCH];
  auto gather = [&](int i0, int (&cc)[RANS_BATCH], float (&mm)[RANS_BATCH]) {
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) {
      const int i = min(i0 + b, count - 1);
      const size_t pos = base + (size_t)i * spc;
      cc[b] = index_is_channel ? (c % index_is_channel) : (int)idx[pos];
      mm[b] = (val_out == nullptr) ? 0.f : ((mu != nullptr) ? mu[pos + mu_extra] : med);
    }
  };
  if (count > 0) gather(0, ci, mean);
  for (int i0 = 0; i0 < count; i0 += RANS_BATCH) {
    if (i0 + RANS_BATCH < count) gather(i0 + RANS_BATCH, ci_n, mean_n);
    int32_t outv[RANS_BATCH];
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) {
      outv[b] = 0;
      if (i0 + b >= count) break;
      const int r = ci[b];
      const uint32_t peeked = d.peek();
      const uint16_t* row = T.cdf16 + T.row_off[r];
      const int32_t n_entries = T.len[r];
      const int32_t max_value = n_entries - 2;
      const uint32_t cum = (uint32_t)(d.x & 0xffffu);
      int lo, hi;   // last v in [0, n_entries-1) with row[v] <= cum  (row[n_entries-1] = 65536 is stored as 0, never probed)
      if (T.lut != nullptr) {
        const uint16_t* lr = T.lut + r * RANS_LUT + (cum >> 8);
        lo = lr[0];
        hi = lr[1];
      } else {
        lo = 0;
        hi = n_entries - 2;
      }
      if (hi - lo <= 4) {
        // the usual case (a 256-wide bucket of the inverse table spans a handful of symbols except in the far tails):
        // four independent probes, no dependent chain, no divergence. The row is increasing, so the hits form a prefix.
        const uint16_t* pr = row + lo;
        const int c1 = (lo + 1 <= hi) & ((uint32_t)pr[1] <= cum);
        const int c2 = (lo + 2 <= hi) & ((uint32_t)pr[2] <= cum);
        const int c3 = (lo + 3 <= hi) & ((uint32_t)pr[3] <= cum);
        const int c4 = (lo + 4 <= hi) & ((uint32_t)pr[4] <= cum);
        lo += c1 + c2 + c3 + c4;
      } else {
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if ((uint32_t)row[mid] <= cum) lo = mid; else hi = mid - 1;
        }
      }
      const uint32_t start = row[lo];
      uint32_t freq = ((uint32_t)row[lo + 1] - start) & 0xffffu;
      if (freq == 0u) freq = 65536u;
      d.x = (uint64_t)freq * (d.x >> RANS_PRECISION) + cum - start;
      if (d.x < RANS_L) d.renorm(peeked);
      int32_t value = lo;
      if (value == max_value) {  // bypass, rans_interface.cpp:256-278
        int32_t val = (int32_t)d.get_bits(RANS_BYPASS_BITS);
        int32_t nb = val;
        while (val == RANS_BYPASS_MAX && !d.underflow) {
          val = (int32_t)d.get_bits(RANS_BYPASS_BITS);
          nb += val;
        }
        uint32_t raw = 0;
        for (int32_t j = 0; j < nb; ++j) {
          val = (int32_t)d.get_bits(RANS_BYPASS_BITS);
          if (j < 8) raw |= (uint32_t)val << (j * RANS_BYPASS_BITS);
          if (d.underflow) break;
        }
        value = (int32_t)(raw >> 1);
        if (raw & 1u) value = -value - 1; else value += max_value;
      }
      outv[b] = value + T.off[r];
    }
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) {
      const int i = i0 + b;
      if (i < count) {
        const size_t pos = base + (size_t)i * spc;
        if (sym_out != nullptr) sym_out[pos] = outv[b];
        if (val_out != nullptr) val_out[pos] = __fadd_rn((float)outv[b], mean[b]);
      }
    }
#pragma unroll
    for (int b = 0; b < RANS_BATCH; ++b) { ci[b] = ci_n[b]; mean[b] = mean_n[b]; }
  }
  if (d.underflow) atomicExch(err, 2);
}

// smem-table variants of rans_encode / rans_decode: `packed` / `row_off` from pack_cdf (total entries), optional coarse
// inverse table `lut` for decoding. Return false (nothing launched) when the tables do not fit shared memory.
bool rans_tables_fit(int rows, int total, bool with_lut) {
  return smem_tables_bytes(rows, total, with_lut) <= 200 * 1024;
}

void rans_encode_smem(cudaStream_t st, const int32_t* sym, const uint8_t* idx, int chan_mod,
                      const uint16_t* packed, const int32_t* row_off, const int32_t* cdf_len, const int32_t* offset,
                      int rows, int total, int n_channels, int L, int spc, int chan_len, uint32_t* scratch, int cap_words,
                      uint32_t* lengths, uint32_t* offsets, uint8_t* payload, int* err) {
  const int n_streams = n_channels * spc;
  if (n_streams == 0) return;
  const size_t smem = smem_tables_bytes(rows, total, false);
  if (smem > 48 * 1024) ensure_dynamic_smem(rans_encode_smem_kernel, 200 * 1024);
  {
    LaunchScope scope(st, "rans_encode", 0.0, (double)n_channels * L * (chan_mod ? 4.0 : 5.0));
    rans_encode_smem_kernel<<<(n_streams + RANS_STHREADS - 1) / RANS_STHREADS, RANS_STHREADS, smem, st>>>(
        sym, idx, chan_mod, packed, row_off, cdf_len, offset, rows, total, n_channels, L, spc, chan_len,
        scratch, cap_words, lengths, err);
  }
  CRA5_CUDA(cudaGetLastError());
  {
    LaunchScope scope(st, "scan_lengths", 0.0, 8.0 * n_streams);
    scan_lengths_kernel<<<1, 1024, 0, st>>>(lengths, n_streams, offsets);
  }
  CRA5_CUDA(cudaGetLastError());
  LaunchScope scope(st, "compact_streams", 0.0, 0.0);
  compact_streams_kernel<<<(n_streams * 32 + 255) / 256, 256, 0, st>>>(scratch, cap_words, lengths, offsets, n_streams,
                                                                       payload);
  CRA5_CUDA(cudaGetLastError());
}

void rans_decode_smem(cudaStream_t st, const uint8_t* payload, const uint32_t* offsets, const uint8_t* idx,
                      int chan_mod, const uint16_t* packed, const int32_t* row_off, const int32_t* cdf_len,
                      const int32_t* offset, const uint16_t* lut, int rows, int total, int n_channels, int L, int spc,
                      int chan_len, int32_t* sym_out, const float* mu, const float* median, float* val_out, int* err, int ch_per_frame,
                 size_t mu_frame_extra) {
  const int n_streams = n_channels * spc;
  if (n_streams == 0) return;
  const size_t smem = smem_tables_bytes(rows, total, lut != nullptr);
  if (smem > 48 * 1024) ensure_dynamic_smem(rans_decode_smem_kernel, 200 * 1024);
  LaunchScope scope(st, "rans_decode", 0.0, (double)n_channels * L * (chan_mod ? 4.0 : 9.0));
  rans_decode_smem_kernel<<<(n_streams + RANS_DTHREADS - 1) / RANS_DTHREADS, RANS_DTHREADS, smem, st>>>(
      payload, offsets, idx, chan_mod, packed, row_off, cdf_len, offset, lut, rows, total, n_channels, L,
      spc, chan_len, sym_out, mu, median, val_out, err, ch_per_frame > 0 ? ch_per_frame : (1 << 30), mu_frame_extra);
  CRA5_CUDA(cudaGetLastError());
}

void scan_lengths(cudaStream_t st, const uint32_t* lengths, int n, uint32_t* offsets) {
  LaunchScope scope(st, "scan_lengths", 0.0, 8.0 * n);
  scan_lengths_kernel<<<1, 1024, 0, st>>>(lengths, n, offsets);
  CRA5_CUDA(cudaGetLastError());
}

}  // namespace cra5
