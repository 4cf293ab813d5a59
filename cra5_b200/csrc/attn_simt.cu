// attn_simt.cu -- generic fp32 attention for the small, awkward shapes of the hyperprior transformer
// (h_a / h_s: 648 tokens, 5 heads x 72, vit_nlc.py:94-112 through HyperpriorEncoder/Decoder) and for any head_dim
// other than 64. bf16 operands in the same layout the QKV epilogue writes (Q,K [head][rows][hd], Vt [head][hd][rows]),
// fp32 scores / softmax / accumulation. One warp per query row; scores are staged in shared memory.
// ~0.6 GFLOP per hyperprior block: latency-, not throughput-relevant.
#include "ptx.cuh"
#include "host_util.h"
#include "kernels.h"
#include <algorithm>

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

__global__ void __launch_bounds__(AS_WARPS * 32)
attn_small_kernel(const __nv_bfloat16* __restrict__ Q, const __nv_bfloat16* __restrict__ K,
                  const __nv_bfloat16* __restrict__ Vt, __nv_bfloat16* __restrict__ out, int ldo, int hd,
                  int rows_total, int S, int kw, int vw, int kv_words) {
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
    myq[t] = (qi < S) ? __bfloat162float(Q[((size_t)head * rows_total + row0 + qi) * hd + d]) : 0.f;
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
        const uint32_t kk = kr[w];
        const float k0 = __uint_as_float(kk << 16), k1 = __uint_as_float(kk & 0xffff0000u);
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
      const uint32_t vv = vr[w];
      const float v0 = __uint_as_float(vv << 16), v1 = __uint_as_float(vv & 0xffff0000u);
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
      if (qi < S) out[(row0 + qi) * ldo + (size_t)head * hd + d] = __float2bfloat16(acc[qq] * inv[qq]);
    }
  }
}

void attention_simt(cudaStream_t st, const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* Vt,
                    __nv_bfloat16* out, int ldo, int heads, int hd, int rows_total, int seg_len) {
  CRA5_CHECK(seg_len > 0 && rows_total % seg_len == 0, ERR_INVALID, "attention: rows must be whole segments");
  CRA5_CHECK((hd & 1) == 0, ERR_INVALID, "attention: head_dim must be even");
  if ((rows_total & 1) == 0 && (seg_len & 1) == 0) {
    // preferred: K, then V^T, of a (segment, head) resident in shared memory
    const int kw = (hd >> 1) | 1, vw = (seg_len >> 1) | 1;
    const int kv_words = (int)((std::max((size_t)seg_len * kw, (size_t)hd * vw) + 1) & ~size_t(1));  // keeps sc 8-byte aligned
    const size_t smem2 = (size_t)kv_words * 4 + ((size_t)AS_WARPS * AS2_QPW * seg_len + (size_t)AS_WARPS * AS2_QPW * hd) * 4;
    if (smem2 <= 227 * 1024) {
      static size_t configured2 = 0;
      if (smem2 > 48 * 1024 && smem2 > configured2) {
        CRA5_CUDA(cudaFuncSetAttribute(attn_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        configured2 = smem2;
      }
      dim3 grid((seg_len + AS2_QPB - 1) / AS2_QPB, rows_total / seg_len, heads);
      LaunchScope scope(st, "attn_small", 4.0 * heads * (double)rows_total * seg_len * hd,
                        4.0 * 2.0 * heads * (double)rows_total * hd);
      attn_small_kernel<<<grid, AS_WARPS * 32, smem2, st>>>(Q, K, Vt, out, ldo, hd, rows_total, seg_len, kw, vw,
                                                           kv_words);
      CRA5_CUDA(cudaGetLastError());
      return;
    }
  }
  const size_t smem = ((size_t)AS_WARPS * seg_len + (size_t)AS_WARPS * hd) * sizeof(float);
  CRA5_CHECK(smem <= 200 * 1024, ERR_INVALID, "attention_simt: segment too long for the generic kernel");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    CRA5_CUDA(cudaFuncSetAttribute(attn_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((seg_len + AS_WARPS - 1) / AS_WARPS, rows_total / seg_len, heads);
  LaunchScope scope(st, "attn_simt", 4.0 * heads * (double)rows_total * seg_len * hd,
                    4.0 * 2.0 * heads * (double)rows_total * hd);
  attn_simt_kernel<<<grid, AS_WARPS * 32, smem, st>>>(Q, K, Vt, out, ldo, hd, rows_total, seg_len);
  CRA5_CUDA(cudaGetLastError());
}

}  // namespace cra5
