// attn_simt.cu -- generic fp32 attention for the small, awkward shapes of the hyperprior transformer
// (h_a / h_s: 648 tokens, 5 heads x 72, vit_nlc.py:94-112 through HyperpriorEncoder/Decoder) and for any head_dim
// other than 64. bf16 operands in the same layout the QKV epilogue writes (Q,K [head][rows][hd], Vt [head][hd][rows]),
// fp32 scores / softmax / accumulation. One warp per query row; scores are staged in shared memory.
// ~0.6 GFLOP per hyperprior block: latency-, not throughput-relevant.
#include "ptx.cuh"
#include "host_util.h"
#include "kernels.h"

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

void attention_simt(cudaStream_t st, const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* Vt,
                    __nv_bfloat16* out, int ldo, int heads, int hd, int rows_total, int seg_len) {
  CRA5_CHECK(seg_len > 0 && rows_total % seg_len == 0, ERR_INVALID, "attention: rows must be whole segments");
  CRA5_CHECK((hd & 1) == 0, ERR_INVALID, "attention: head_dim must be even");
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
