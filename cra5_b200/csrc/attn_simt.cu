// attn_simt.cu -- generic fp32 attention for the small, awkward shapes of the hyperprior transformer
// (h_a / h_s: 648 tokens, 5 heads x 72, vit_nlc.py:94-112 through HyperpriorEncoder/Decoder) and for any head_dim
// other than 64. bf16 operands in the same layout the QKV epilogue writes (Q,K [head][rows][hd], Vt [head][hd][rows]),
// fp32 scores / softmax / accumulation. One warp per query row; scores are staged in shared memory.
// ~0.6 GFLOP per hyperprior block: latency-, not throughput-relevant.
#include "ptx.cuh"
#include "host_util.h"
#include "kernels.h"
#include <algorithm>
#include <cstdlib>

namespace cra5 {

constexpr int AS_WARPS = 8;

__global__ void __launch_bounds__(AS_WARPS * 32)
attn_simt_kernel(const __nv_bfloat16* __restrict__ Q, const __nv_bfloat16* __restrict__ K,
                 const __nv_bfloat16* __restrict__ Vt, __nv_bfloat16* __restrict__ out, int ldo, int hd,
                 int rows_total, int seg_len) {
  extern __shared__ float smem[];  // [AS_WARPS][seg_len] scores + [AS_WARPS][hd] query
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.z, seg = blockIdx.y;
  const int qi = blockIdx.x * AS_WARPS + warp;
  if (qi >= seg_len) return;
  float* sc = smem + (size_t)warp * seg_len;
  float* qv = smem + (size_t)AS_WARPS * seg_len + (size_t)warp * hd;
  const size_t row0 = (size_t)seg * seg_len;
  const __nv_bfloat16* q = Q + ((size_t)head * rows_total + row0 + qi) * hd;
  for (int d = lane; d < hd; d += 32) qv[d] = __bfloat162float(q[d]);
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < seg_len; j += 32) {
    const __nv_bfloat16* k = K + ((size_t)head * rows_total + row0 + j) * hd;
    float s = 0.f;
    for (int d = 0; d < hd; d += 2) {
      const __nv_bfloat162 kk = *reinterpret_cast<const __nv_bfloat162*>(k + d);
      s = fmaf(qv[d], __low2float(kk), s);
      s = fmaf(qv[d + 1], __high2float(kk), s);
    }
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < seg_len; j += 32) {
    const float p = __expf(sc[j] - mx);
    sc[j] = p;
    sum += p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncwarp();
  const float inv = 1.0f / sum;
  for (int d = lane; d < hd; d += 32) {
    const __nv_bfloat16* v = Vt + ((size_t)head * hd + d) * rows_total + row0;
    float acc = 0.f;
    for (int j = 0; j < seg_len; ++j) acc = fmaf(sc[j], __bfloat162float(v[j]), acc);
    out[(row0 + qi) * ldo + (size_t)head * hd + d] = __float2bfloat16(acc * inv);
  }
}

// ------------------------------------------------------------------------------------------------ smem-resident variant
// K of one (segment, head) is staged in shared memory for the score phase, then V^T is staged into the SAME region for
// the P V phase (648 x 72 bf16 = 93 KB each); rows are padded to an odd number of 32-bit words so both access patterns
// are bank-conflict free. 8 warps x 4 queries per CTA; every K / V word read from smem feeds 8 FMAs.
constexpr int AS2_QPW = 4;             // queries per warp
constexpr int AS2_QPB = AS_WARPS * AS2_QPW;

// two 16-bit operand elements packed in a word -> fp32 (bf16: a shift; fp16 for the split-precision levels)
template <bool F16>
__device__ __forceinline__ void unpack16x2(uint32_t w, float& lo, float& hi) {
  if constexpr (F16) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    lo = f.x; hi = f.y;
  } else {
    lo = __uint_as_float(w << 16); hi = __uint_as_float(w & 0xffff0000u);
  }
}

template <bool F16>
__global__ void __launch_bounds__(AS_WARPS * 32)
attn_small_kernel(const __nv_bfloat16* __restrict__ Q, const __nv_bfloat16* __restrict__ K,
                  const __nv_bfloat16* __restrict__ Vt, __nv_bfloat16* __restrict__ out,
                  __nv_bfloat16* __restrict__ out_lo, int ldo, int hd, int rows_total, int S, int kw, int vw,
                  int kv_words) {
  extern __shared__ uint32_t sm32[];
  uint32_t* KV = sm32;                                          // K as [S][kw] words, later V^T as [hd][vw] words
  float* sc = reinterpret_cast<float*>(sm32 + kv_words);        // [AS_WARPS][AS2_QPW][S] scores / probabilities
  float* qs = sc + (size_t)AS_WARPS * AS2_QPW * S;              // [AS_WARPS][AS2_QPW][hd]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.z, seg = blockIdx.y;
  const size_t row0 = (size_t)seg * S;
  const int hw = hd >> 1, sw = S >> 1;
  const uint32_t* Kg = reinterpret_cast<const uint32_t*>(K + ((size_t)head * rows_total + row0) * hd);
  for (int e = threadIdx.x; e < S * hw; e += blockDim.x) {
    const int j = e / hw, w = e - j * hw;
    KV[j * kw + w] = Kg[e];
  }
  const int q_base = blockIdx.x * AS2_QPB + warp * AS2_QPW;
  float* myq = qs + (size_t)warp * AS2_QPW * hd;
  float* mysc = sc + (size_t)warp * AS2_QPW * S;
  for (int t = lane; t < AS2_QPW * hd; t += 32) {
    const int qq = t / hd, d = t - qq * hd;
    const int qi = q_base + qq;
    const __nv_bfloat16 qe = Q[((size_t)head * rows_total + row0 + min(qi, S - 1)) * hd + d];
    myq[t] = (qi < S) ? (F16 ? __half2float(*reinterpret_cast<const __half*>(&qe)) : __bfloat162float(qe)) : 0.f;
  }
  __syncthreads();
  const bool active = q_base < S;
  float inv[AS2_QPW];
  if (active) {
    // ---- scores: lanes stride over keys
    float mx[AS2_QPW];
#pragma unroll
    for (int qq = 0; qq < AS2_QPW; ++qq) mx[qq] = -INFINITY;
    for (int j = lane; j < S; j += 32) {
      float acc[AS2_QPW];
#pragma unroll
      for (int qq = 0; qq < AS2_QPW; ++qq) acc[qq] = 0.f;
      const uint32_t* kr = KV + (size_t)j * kw;
      for (int w = 0; w < hw; ++w) {
        float k0, k1;
        unpack16x2<F16>(kr[w], k0, k1);
#pragma unroll
        for (int qq = 0; qq < AS2_QPW; ++qq) {
          const float2 qv = *reinterpret_cast<const float2*>(myq + qq * hd + 2 * w);
          acc[qq] = fmaf(qv.x, k0, acc[qq]);
          acc[qq] = fmaf(qv.y, k1, acc[qq]);
        }
      }
#pragma unroll
      for (int qq = 0; qq < AS2_QPW; ++qq) {
        mysc[qq * S + j] = acc[qq];
        mx[qq] = fmaxf(mx[qq], acc[qq]);
      }
    }
#pragma unroll
    for (int qq = 0; qq < AS2_QPW; ++qq) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx[qq] = fmaxf(mx[qq], __shfl_xor_sync(0xffffffffu, mx[qq], o));
      float sum = 0.f;
      for (int j = lane; j < S; j += 32) {
        const float p = __expf(mysc[qq * S + j] - mx[qq]);
        mysc[qq * S + j] = p;
        sum += p;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      inv[qq] = 1.0f / sum;
    }
  }
  __syncthreads();  // everyone is done with K
  for (int e = threadIdx.x; e < hd * sw; e += blockDim.x) {
    const int d = e / sw, w = e - d * sw;
    KV[d * vw + w] = *reinterpret_cast<const uint32_t*>(Vt + ((size_t)head * hd + d) * rows_total + row0 + 2 * w);
  }
  __syncthreads();
  if (!active) return;
  // ---- P V: lanes stride over head dims
  for (int d = lane; d < hd; d += 32) {
    const uint32_t* vr = KV + (size_t)d * vw;
    float acc[AS2_QPW];
#pragma unroll
    for (int qq = 0; qq < AS2_QPW; ++qq) acc[qq] = 0.f;
    for (int w = 0; w < sw; ++w) {
      float v0, v1;
      unpack16x2<F16>(vr[w], v0, v1);
#pragma unroll
      for (int qq = 0; qq < AS2_QPW; ++qq) {
        const float2 pv = *reinterpret_cast<const float2*>(mysc + qq * S + 2 * w);
        acc[qq] = fmaf(pv.x, v0, acc[qq]);
        acc[qq] = fmaf(pv.y, v1, acc[qq]);
      }
    }
#pragma unroll
    for (int qq = 0; qq < AS2_QPW; ++qq) {
      const int qi = q_base + qq;
      if (qi < S) {
        const float o = acc[qq] * inv[qq];
        const __nv_bfloat16 hi = __float2bfloat16(o);
        out[(row0 + qi) * ldo + (size_t)head * hd + d] = hi;
        if (out_lo != nullptr) out_lo[(row0 + qi) * ldo + (size_t)head * hd + d] = __float2bfloat16(o - __bfloat162float(hi));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ warp-MMA variant
// The hyperprior attention is 0.6 GFLOP per block -- far too small for a tcgen05/TMEM pipeline to amortise its set-up,
// and head_dim 72 is not a UMMA-friendly K -- but it runs 12 times per frame, so its latency matters. Here K and V^T of
// one (segment, head) are staged in shared memory with cp.async, and each warp runs a flash-style loop for 16 query
// rows on warp-level tensor-core MMAs (mma.sync m16n8k16 bf16, fp32 accumulate; operands through ldmatrix):
// S = Q K^T for 64 keys, online softmax in the accumulator layout, P re-used as the A operand of O += P V.
// Row strides (HD and S_pad + 8 elements) make every ldmatrix phase bank-conflict free.
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x1(uint32_t addr, uint32_t& r0) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x1.shared.b16 {%0}, [%1];" : "=r"(r0) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

constexpr int AM_WARPS = 4;
constexpr int AM_QPB = AM_WARPS * 16;   // queries per CTA
constexpr int AM_KC = 64;               // keys per step

template <int HD>
__global__ void __launch_bounds__(AM_WARPS * 32)
attn_mma_kernel(const __nv_bfloat16* __restrict__ Q, const __nv_bfloat16* __restrict__ K,
                const __nv_bfloat16* __restrict__ Vt, __nv_bfloat16* __restrict__ out, int ldo, int rows_total, int S,
                int s_pad, int sv) {
  static_assert(HD % 8 == 0 && HD <= 128, "head_dim must be a multiple of 8");
  constexpr int KS_FULL = HD / 16;          // whole 16-wide k-steps of Q K^T
  constexpr bool KS_HALF = (HD % 16) != 0;  // plus one 8-wide step (upper half zero)
  constexpr int NT_O = HD / 8;              // 8-wide output tiles
  extern __shared__ __align__(16) uint8_t am_smem[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(am_smem);   // [s_pad][HD]
  __nv_bfloat16* Vs = Ks + (size_t)s_pad * HD;                      // [HD][sv]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.z, seg = blockIdx.y;
  const size_t row0 = (size_t)seg * S;
  // ---- stage K rows [0, S) and V^T columns [0, S) of this (segment, head); zero the padding
  {
    const uint8_t* kg = reinterpret_cast<const uint8_t*>(K + ((size_t)head * rows_total + row0) * HD);
    const int k_vec = S * HD * 2 / 16;
    for (int e = threadIdx.x; e < k_vec; e += blockDim.x) cp_async16(smem_u32(am_smem + (size_t)e * 16), kg + (size_t)e * 16);
    const int kpad_vec = (s_pad - S) * HD * 2 / 16;
    for (int e = threadIdx.x; e < kpad_vec; e += blockDim.x)
      *reinterpret_cast<uint4*>(am_smem + (size_t)(k_vec + e) * 16) = make_uint4(0, 0, 0, 0);
    const int v_row_vec = S / 8, v_row_all = sv / 8;
    for (int e = threadIdx.x; e < HD * v_row_all; e += blockDim.x) {
      const int d = e / v_row_all, w = e - d * v_row_all;
      __nv_bfloat16* dst = Vs + (size_t)d * sv + w * 8;
      if (w < v_row_vec)
        cp_async16(smem_u32(dst), Vt + ((size_t)head * HD + d) * rows_total + row0 + w * 8);
      else
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  const int q0 = blockIdx.x * AM_QPB + warp * 16;
  if (q0 >= S) return;
  const int g = lane >> 2, t = lane & 3;
  // ---- Q fragments (A operand, row-major 16 x 16 per k-step), straight from global; rows >= S read row S-1
  constexpr int KS = KS_FULL + (KS_HALF ? 1 : 0);
  uint32_t qa[KS][4];
  {
    const int r_lo = min(q0 + g, S - 1), r_hi = min(q0 + g + 8, S - 1);
    const __nv_bfloat16* ql = Q + ((size_t)head * rows_total + row0 + r_lo) * HD;
    const __nv_bfloat16* qh = Q + ((size_t)head * rows_total + row0 + r_hi) * HD;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int d = ks * 16 + 2 * t;
      qa[ks][0] = *reinterpret_cast<const uint32_t*>(ql + d);
      qa[ks][1] = *reinterpret_cast<const uint32_t*>(qh + d);
      if (ks < KS_FULL) {
        qa[ks][2] = *reinterpret_cast<const uint32_t*>(ql + d + 8);
        qa[ks][3] = *reinterpret_cast<const uint32_t*>(qh + d + 8);
      } else {
        qa[ks][2] = 0u;
        qa[ks][3] = 0u;
      }
    }
  }
  float o[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  constexpr float LOG2E = 1.4426950408889634f;
  const uint32_t ks_base = smem_u32(Ks), vs_base = smem_u32(Vs);
  // ldmatrix row addresses: K tile rows = keys (lane % 8), 8-dim column block = lane / 8
  const uint32_t k_lane = (uint32_t)((lane & 7) * HD + (lane >> 3) * 8) * 2;
  // V^T tile rows = dims (lane % 8 + 8 * (lane / 16)), key block = (lane / 8) % 2
  const uint32_t v_lane = (uint32_t)(((lane & 7) + 8 * (lane >> 4)) * sv + ((lane >> 3) & 1) * 8) * 2;

  for (int kc = 0; kc < S; kc += AM_KC) {
    float sacc[AM_KC / 8][4];
#pragma unroll
    for (int nt = 0; nt < AM_KC / 8; ++nt) {
      sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
      const uint32_t kaddr = ks_base + (uint32_t)((kc + nt * 8) * HD) * 2 + k_lane;
#pragma unroll
      for (int kp = 0; kp + 1 < KS_FULL + 1 && kp * 2 + 1 < KS_FULL; ++kp) {   // pairs of whole k-steps
        uint32_t b0, b1, b2, b3;
        ldsm_x4(kaddr + kp * 64, b0, b1, b2, b3);
        mma_bf16_16816(sacc[nt], qa[2 * kp], b0, b1);
        mma_bf16_16816(sacc[nt], qa[2 * kp + 1], b2, b3);
      }
      if constexpr (KS_FULL % 2 == 1) {   // one whole k-step left over
        uint32_t b0, b1;
        ldsm_x2(kaddr + (KS_FULL - 1) * 32, b0, b1);
        mma_bf16_16816(sacc[nt], qa[KS_FULL - 1], b0, b1);
      }
      if constexpr (KS_HALF) {            // 8 remaining dims
        uint32_t b0;
        ldsm_x1(kaddr + KS_FULL * 32, b0);
        mma_bf16_16816(sacc[nt], qa[KS_FULL], b0, 0u);
      }
    }
    // ---- online softmax over these 64 keys (rows g and g + 8 of the warp's 16 queries)
    if (kc + AM_KC > S) {
#pragma unroll
      for (int nt = 0; nt < AM_KC / 8; ++nt) {
        const int key = kc + nt * 8 + 2 * t;
        if (key >= S) { sacc[nt][0] = -INFINITY; sacc[nt][2] = -INFINITY; }
        if (key + 1 >= S) { sacc[nt][1] = -INFINITY; sacc[nt][3] = -INFINITY; }
      }
    }
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < AM_KC / 8; ++nt) {
      mx_lo = fmaxf(mx_lo, fmaxf(sacc[nt][0], sacc[nt][1]));
      mx_hi = fmaxf(mx_hi, fmaxf(sacc[nt][2], sacc[nt][3]));
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
    const float a_lo = exp2f((m_lo - mn_lo) * LOG2E), a_hi = exp2f((m_hi - mn_hi) * LOG2E);
    m_lo = mn_lo; m_hi = mn_hi;
    const float nl = -mn_lo * LOG2E, nh = -mn_hi * LOG2E;
    float s_lo = 0.f, s_hi = 0.f;
    uint32_t pa[AM_KC / 16][4];
#pragma unroll
    for (int nt = 0; nt < AM_KC / 8; ++nt) {
      const float p0 = exp2f(fmaf(sacc[nt][0], LOG2E, nl)), p1 = exp2f(fmaf(sacc[nt][1], LOG2E, nl));
      const float p2 = exp2f(fmaf(sacc[nt][2], LOG2E, nh)), p3 = exp2f(fmaf(sacc[nt][3], LOG2E, nh));
      s_lo += p0 + p1;
      s_hi += p2 + p3;
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    l_lo = l_lo * a_lo + s_lo;
    l_hi = l_hi * a_hi + s_hi;
#pragma unroll
    for (int i = 0; i < NT_O; ++i) { o[i][0] *= a_lo; o[i][1] *= a_lo; o[i][2] *= a_hi; o[i][3] *= a_hi; }
    // ---- O += P V: k-steps of 16 keys, output tiles of 8 dims (two per ldmatrix.x4)
#pragma unroll
    for (int ks = 0; ks < AM_KC / 16; ++ks) {
      const uint32_t vaddr = vs_base + (uint32_t)(kc + ks * 16) * 2 + v_lane;
#pragma unroll
      for (int np = 0; np < NT_O / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(vaddr + (uint32_t)(np * 16 * sv) * 2, b0, b1, b2, b3);
        mma_bf16_16816(o[2 * np], pa[ks], b0, b1);
        mma_bf16_16816(o[2 * np + 1], pa[ks], b2, b3);
      }
      if constexpr (NT_O % 2 == 1) {
        uint32_t b0, b1;
        ldsm_x2(vaddr + (uint32_t)((NT_O - 1) * 8 * sv) * 2, b0, b1);   // lanes 16-31 supply ignored addresses
        mma_bf16_16816(o[NT_O - 1], pa[ks], b0, b1);
      }
    }
  }
  // the row sums are spread over the four lanes of a quad
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float i_lo = 1.0f / l_lo, i_hi = 1.0f / l_hi;
  const int r_lo = q0 + g, r_hi = q0 + g + 8;
#pragma unroll
  for (int i = 0; i < NT_O; ++i) {
    const int d = i * 8 + 2 * t;
    if (r_lo < S)
      *reinterpret_cast<uint32_t*>(out + (row0 + r_lo) * ldo + (size_t)head * HD + d) = pack_bf16x2(o[i][0] * i_lo, o[i][1] * i_lo);
    if (r_hi < S)
      *reinterpret_cast<uint32_t*>(out + (row0 + r_hi) * ldo + (size_t)head * HD + d) = pack_bf16x2(o[i][2] * i_hi, o[i][3] * i_hi);
  }
}

template <int HD>
static bool launch_attn_mma(cudaStream_t st, const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* Vt,
                            __nv_bfloat16* out, int ldo, int heads, int rows_total, int S) {
  if ((S & 7) != 0 || (rows_total & 7) != 0 || (ldo & 1) != 0) return false;   // 16-byte staging, 4-byte stores
  const int s_pad = (S + AM_KC - 1) / AM_KC * AM_KC;
  const int sv = s_pad + 8;   // (sv / 2) % 8 == 4: the 8 rows of an ldmatrix phase fall into distinct bank groups
  const size_t smem = ((size_t)s_pad * HD + (size_t)HD * sv) * 2;
  if (smem > 227 * 1024) return false;
  ensure_dynamic_smem(attn_mma_kernel<HD>, 227 * 1024);
  dim3 grid((S + AM_QPB - 1) / AM_QPB, rows_total / S, heads);
  LaunchScope scope(st, "attn_small", 4.0 * heads * (double)rows_total * S * HD, 4.0 * 2.0 * heads * (double)rows_total * HD);
  launch_chained(attn_mma_kernel<HD>, grid, dim3(AM_WARPS * 32), smem, st, Q, K, Vt, out, ldo, rows_total, S, s_pad, sv);
  return true;
}

void attention_simt(cudaStream_t st, const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* Vt,
                    __nv_bfloat16* out, int ldo, int heads, int hd, int rows_total, int seg_len, bool f16,
                    __nv_bfloat16* out_lo) {
  CRA5_CHECK(seg_len > 0 && rows_total % seg_len == 0, ERR_INVALID, "attention: rows must be whole segments");
  CRA5_CHECK((hd & 1) == 0, ERR_INVALID, "attention: head_dim must be even");
  static const bool no_mma = getenv("CRA5_ATTN_SMALL_SIMT") != nullptr;   // diagnostics: the fp32 SIMT kernels
  if (!no_mma && !f16 && out_lo == nullptr) {
    if (hd == 72 && launch_attn_mma<72>(st, Q, K, Vt, out, ldo, heads, rows_total, seg_len)) return;
    if (hd == 24 && launch_attn_mma<24>(st, Q, K, Vt, out, ldo, heads, rows_total, seg_len)) return;
  }
  if ((rows_total & 1) == 0 && (seg_len & 1) == 0) {
    // preferred: K, then V^T, of a (segment, head) resident in shared memory
    const int kw = (hd >> 1) | 1, vw = (seg_len >> 1) | 1;
    const int kv_words = (int)((std::max((size_t)seg_len * kw, (size_t)hd * vw) + 1) & ~size_t(1));  // keeps sc 8-byte aligned
    const size_t smem2 = (size_t)kv_words * 4 + ((size_t)AS_WARPS * AS2_QPW * seg_len + (size_t)AS_WARPS * AS2_QPW * hd) * 4;
    if (smem2 <= 227 * 1024) {
      ensure_dynamic_smem(attn_small_kernel<false>, smem2);
      ensure_dynamic_smem(attn_small_kernel<true>, smem2);
      dim3 grid((seg_len + AS2_QPB - 1) / AS2_QPB, rows_total / seg_len, heads);
      LaunchScope scope(st, "attn_small", 4.0 * heads * (double)rows_total * seg_len * hd,
                        4.0 * 2.0 * heads * (double)rows_total * hd);
      if (f16)
        attn_small_kernel<true><<<grid, AS_WARPS * 32, smem2, st>>>(Q, K, Vt, out, out_lo, ldo, hd, rows_total, seg_len, kw,
                                                                   vw, kv_words);
      else
        attn_small_kernel<false><<<grid, AS_WARPS * 32, smem2, st>>>(Q, K, Vt, out, out_lo, ldo, hd, rows_total, seg_len, kw,
                                                                    vw, kv_words);
      CRA5_CUDA(cudaGetLastError());
      return;
    }
  }
  CRA5_CHECK(!f16 && out_lo == nullptr, ERR_INVALID, "attention_simt: the fp16 / split-output form needs the shared-memory kernel");
  const size_t smem = ((size_t)AS_WARPS * seg_len + (size_t)AS_WARPS * hd) * sizeof(float);
  CRA5_CHECK(smem <= 200 * 1024, ERR_INVALID, "attention_simt: segment too long for the generic kernel");
  ensure_dynamic_smem(attn_simt_kernel, smem);
  dim3 grid((seg_len + AS_WARPS - 1) / AS_WARPS, rows_total / seg_len, heads);
  LaunchScope scope(st, "attn_simt", 4.0 * heads * (double)rows_total * seg_len * hd,
                    4.0 * 2.0 * heads * (double)rows_total * hd);
  attn_simt_kernel<<<grid, AS_WARPS * 32, smem, st>>>(Q, K, Vt, out, ldo, hd, rows_total, seg_len);
  CRA5_CUDA(cudaGetLastError());
}

}  // namespace cra5
