// attn_tc.cu -- fused softmax(Q K^T) V on tcgen05 tensor cores for head_dim 64.
//
// Serves both attention flavours of the VAEformer trunk (reference: WindowAttention.forward
// vit_nlc.py:219-258 -- 576-token windows incl. the un-masked zero-pad tokens -- and Attention.forward
// vit_nlc.py:94-112 -- one 10 368-token segment): the token list arrives already in "attention order"
// (window-partitioned by the LayerNorm kernel), so a segment is a contiguous run of seg_len rows.
//
// One CTA = one 128-row query tile of one (segment, head); KV tiles of 128 stream through a 3-stage TMA ring.
//   warp 0      TMA producer
//   warp 1      MMA issuer: S_j = Q K_j^T into one of two TMEM score buffers (so S_{j+1} is computed while the
//               softmax of tile j runs), then O_j = P_j V_j into one of two TMEM output buffers
//   warp 2      TMEM allocator
//   warps 4-11  softmax: TWO threads per query row (warp w and w+4 share a TMEM lane quarter; the first takes score
//               columns 0-63 and output dims 0-31, the second 64-127 / 32-63). S is read from TMEM exactly once and
//               kept in registers; the row maximum is exchanged between the two threads through shared memory and a
//               64-thread named barrier; P goes to swizzled smem as bf16; O is accumulated in registers.
// Q was pre-multiplied by head_dim^-0.5 in the QKV GEMM epilogue (reference scales q before the product).
#include "ptx.cuh"
#include "host_util.h"
#include "kernels.h"

namespace cra5 {

__device__ __forceinline__ float ex2_approx(float x) {  // MUFU.EX2, no range fix-up: inputs are <= 0 (or -inf)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for x <= 0 on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax polynomial, rel. error ~1e-4, far below the
// bf16 rounding of P). The softmax is MUFU-bound at head_dim 64 (16 384 exponentials per 128x128 tile against 16
// MUFU lanes per SM and clock), so a fixed fraction of the columns takes this path to unload the XU pipe.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;            // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);      // f in [-0.5, 0.5]
  float p = fmaf(0.0555041f, f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

constexpr int AT_BM = 128;
constexpr int AT_BN = 128;
constexpr int AT_HD = 64;
constexpr int AT_STAGES = 4;
constexpr int AT_THREADS = 384;

struct AttnSmem {
  static constexpr int Q_BYTES = AT_BM * AT_HD * 2;        // 16 KB
  static constexpr int K_BYTES = AT_BN * AT_HD * 2;        // 16 KB
  static constexpr int V_BYTES = AT_HD * AT_BN * 2;        // 16 KB (two 64x64 halves)
  static constexpr int KV_BYTES = K_BYTES + V_BYTES;
  static constexpr int P_BYTES = AT_BM * AT_BN * 2;        // 32 KB (two 128x64 halves)
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = OFF_Q + Q_BYTES;
  static constexpr int OFF_P = OFF_KV + AT_STAGES * KV_BYTES;
  static constexpr int OFF_X = OFF_P + 2 * P_BYTES;        // [2][128] floats: row max / row sum exchange
  static constexpr int OFF_BAR = OFF_X + 2 * AT_BM * 4;
  static constexpr int TOTAL = OFF_BAR + 128;  // no alignment slack: the dynamic window starts 1024-aligned (checked)
};

struct AttnParams {
  int seg_len;      // tokens per segment (window size or whole sequence)
  int rows_total;   // total rows of Q/K per head (= row stride of Vt)
  __nv_bfloat16* out;  // [rows_total, ldo] bf16, head h occupies columns [h*64, h*64+64)
  int ldo;
};

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmVt, const AttnParams p) {
  using L = AttnSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* kv_full = q_full + 1;
  uint64_t* kv_empty = kv_full + AT_STAGES;
  uint64_t* s_full = kv_empty + AT_STAGES;
  uint64_t* p_full = s_full + 2;
  uint64_t* pv_full = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_full + 2);
  float* xch = reinterpret_cast<float*>(smem + L::OFF_X);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BM;          // first query row inside the segment
  const int seg = blockIdx.y;
  const int head = blockIdx.z;
  const int seg_row0 = seg * p.seg_len;       // first row of the segment
  const int n_kv = (p.seg_len + AT_BN - 1) / AT_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmVt);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < AT_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 256);
      mbar_init(&pv_full[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t TM_S = 0;     // 2 x 128 columns
  constexpr uint32_t TM_PV = 256;  // 2 x 64 columns

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_expect_tx(q_full, L::Q_BYTES);
      tma_load_2d(smem + L::OFF_Q, &tmQ, q_full, 0, head * p.rows_total + seg_row0 + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % AT_STAGES;
        const uint32_t ph = (j / AT_STAGES) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        uint8_t* sk = smem + L::OFF_KV + s * L::KV_BYTES;
        uint8_t* sv = sk + L::K_BYTES;
        const int kv0 = seg_row0 + j * AT_BN;
        mbar_expect_tx(&kv_full[s], L::KV_BYTES);
        tma_load_2d(sk, &tmK, &kv_full[s], 0, head * p.rows_total + kv0);
        tma_load_2d(sv, &tmVt, &kv_full[s], kv0, head * AT_HD);
        tma_load_2d(sv + L::V_BYTES / 2, &tmVt, &kv_full[s], kv0 + 64, head * AT_HD);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_s = umma_idesc_bf16(AT_BM, AT_BN);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(AT_BM, AT_HD);
      const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(smem + L::OFF_Q));
      const uint64_t kv_desc0 = umma_smem_desc_sw128(smem_u32(smem + L::OFF_KV));   // stage 0 K tile; +bytes/16 per step
      const uint64_t p_desc0 = umma_smem_desc_sw128(smem_u32(smem + L::OFF_P));
      auto issue_s = [&](int j) {
        const int s = j % AT_STAGES;
        mbar_wait(&kv_full[s], (j / AT_STAGES) & 1);
        tc_fence_after();
        const uint64_t kdesc = kv_desc0 + (uint64_t)((s * L::KV_BYTES) >> 4);
        const uint32_t d_s = tmem_base + TM_S + (j & 1) * AT_BN;
        umma_bf16(d_s, qdesc + 0, kdesc + 0, idesc_s, 0);
        umma_bf16(d_s, qdesc + 2, kdesc + 2, idesc_s, 1);
        umma_bf16(d_s, qdesc + 4, kdesc + 4, idesc_s, 1);
        umma_bf16(d_s, qdesc + 6, kdesc + 6, idesc_s, 1);
        umma_commit(&s_full[j & 1]);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        // S_{j+1} goes into the other score buffer; its previous content (S_{j-1}) was consumed before p_full(j-1)
        if (j + 1 < n_kv) issue_s(j + 1);
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const int s = j % AT_STAGES;
        const uint64_t vdesc = kv_desc0 + (uint64_t)((s * L::KV_BYTES + L::K_BYTES) >> 4);
        const uint64_t pdesc = p_desc0 + (uint64_t)(((j & 1) * L::P_BYTES) >> 4);
        const uint32_t d_pv = tmem_base + TM_PV + (j & 1) * AT_HD;
        constexpr uint64_t PH = (L::P_BYTES / 2) >> 4, VH = (L::V_BYTES / 2) >> 4;  // second 64-wide half of the K extent
        umma_bf16(d_pv, pdesc + 0, vdesc + 0, idesc_pv, 0);
        umma_bf16(d_pv, pdesc + 2, vdesc + 2, idesc_pv, 1);
        umma_bf16(d_pv, pdesc + 4, vdesc + 4, idesc_pv, 1);
        umma_bf16(d_pv, pdesc + 6, vdesc + 6, idesc_pv, 1);
        umma_bf16(d_pv, pdesc + PH + 0, vdesc + VH + 0, idesc_pv, 1);
        umma_bf16(d_pv, pdesc + PH + 2, vdesc + VH + 2, idesc_pv, 1);
        umma_bf16(d_pv, pdesc + PH + 4, vdesc + VH + 4, idesc_pv, 1);
        umma_bf16(d_pv, pdesc + PH + 6, vdesc + VH + 6, idesc_pv, 1);
        umma_commit(&pv_full[j & 1]);
        umma_commit(&kv_empty[s]);
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + output accumulation =====================
    const int quarter = warp & 3;            // TMEM lane quarter
    const int hsel = (warp - 4) >> 2;        // 0: score columns 0-63, output dims 0-31; 1: 64-127 / 32-63
    const int r = quarter * 32 + lane;       // row inside the query tile == TMEM lane
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f;
    constexpr int HC = AT_BN / 2;            // 64 score columns per thread
    constexpr int HD2 = AT_HD / 2;           // 32 output dims per thread
    float m = -INFINITY, l = 0.f;
    float o[HD2];
#pragma unroll
    for (int i = 0; i < HD2; ++i) o[i] = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const int valid = min(AT_BN, p.seg_len - j * AT_BN) - hsel * HC;  // valid columns in this thread's half
      const bool full_tile = (valid >= HC);  // warp-uniform: only a segment's last tile can be partial
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem_base + lane_addr + TM_S + (j & 1) * AT_BN + hsel * HC;
      uint32_t sa[32], sb[32];
      tmem_ld_32x32(s_addr, sa);
      tmem_ld_32x32(s_addr + 32, sb);
      tmem_ld_wait();
      float mloc = -INFINITY;
      if (full_tile) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          mloc = fmaxf(mloc, fmaxf(__uint_as_float(sa[i]), __uint_as_float(sa[i + 1])));
          mloc = fmaxf(mloc, fmaxf(__uint_as_float(sb[i]), __uint_as_float(sb[i + 1])));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < valid) mloc = fmaxf(mloc, __uint_as_float(sa[i]));
          if (32 + i < valid) mloc = fmaxf(mloc, __uint_as_float(sb[i]));
        }
      }
      // exchange the half-row maxima between the two threads that own this row (warps w and w+4)
      xch[hsel * AT_BM + r] = mloc;
      named_bar_sync(1 + quarter, 64);
      const float mx = fmaxf(m, fmaxf(mloc, xch[(hsel ^ 1) * AT_BM + r]));
      named_bar_sync(1 + quarter, 64);  // both reads done before the slot is overwritten by the next tile
      const float alpha = ex2_approx((m - mx) * LOG2E);  // 0 on the first tile (m = -inf)
      const float mxl = mx * LOG2E;
      float pa[32], pb[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float xa = fmaf(__uint_as_float(sa[i]), LOG2E, -mxl), xb = fmaf(__uint_as_float(sb[i]), LOG2E, -mxl);
        if ((i & 3) == 3) {  // every 4th column: polynomial on the FMA pipe
          pa[i] = ex2_poly(xa);
          pb[i] = ex2_poly(xb);
        } else {
          pa[i] = ex2_approx(xa);
          pb[i] = ex2_approx(xb);
        }
      }
      if (!full_tile) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          pa[i] = (i < valid) ? pa[i] : 0.f;
          pb[i] = (32 + i < valid) ? pb[i] : 0.f;
        }
      }
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 2) { s0 += pa[i]; s1 += pa[i + 1]; s2 += pb[i]; s3 += pb[i + 1]; }
      l = l * alpha + ((s0 + s1) + (s2 + s3));
      m = mx;
      // P (bf16) into this thread's half of the K-major SWIZZLE_128B tile: 8 chunks of 8 columns
      uint8_t* region = smem + L::OFF_P + (j & 1) * L::P_BYTES + hsel * (L::P_BYTES / 2);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        u.x = pack_bf16x2(pa[8 * q + 0], pa[8 * q + 1]);
        u.y = pack_bf16x2(pa[8 * q + 2], pa[8 * q + 3]);
        u.z = pack_bf16x2(pa[8 * q + 4], pa[8 * q + 5]);
        u.w = pack_bf16x2(pa[8 * q + 6], pa[8 * q + 7]);
        *reinterpret_cast<uint4*>(region + sw128_offset(r, q)) = u;
        u.x = pack_bf16x2(pb[8 * q + 0], pb[8 * q + 1]);
        u.y = pack_bf16x2(pb[8 * q + 2], pb[8 * q + 3]);
        u.z = pack_bf16x2(pb[8 * q + 4], pb[8 * q + 5]);
        u.w = pack_bf16x2(pb[8 * q + 6], pb[8 * q + 7]);
        *reinterpret_cast<uint4*>(region + sw128_offset(r, 4 + q)) = u;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[j & 1]);
      // fold the previous tile's P V product (this thread's 32 output dims), rescaled to the new maximum
      if (j > 0) {
        mbar_wait(&pv_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        uint32_t pv[32];
        tmem_ld_32x32(tmem_base + lane_addr + TM_PV + ((j - 1) & 1) * AT_HD + hsel * HD2, pv);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < HD2; ++i) o[i] = (o[i] + __uint_as_float(pv[i])) * alpha;
      }
    }
    {
      const int j = n_kv - 1;
      mbar_wait(&pv_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      uint32_t pv[32];
      tmem_ld_32x32(tmem_base + lane_addr + TM_PV + (j & 1) * AT_HD + hsel * HD2, pv);
      tmem_ld_wait();
      xch[hsel * AT_BM + r] = l;  // total row sum = both halves
      named_bar_sync(1 + quarter, 64);
      const float inv = 1.0f / (l + xch[(hsel ^ 1) * AT_BM + r]);
#pragma unroll
      for (int i = 0; i < HD2; ++i) o[i] = (o[i] + __uint_as_float(pv[i])) * inv;
    }
    if (q0 + r < p.seg_len) {
      __nv_bfloat16* dst = p.out + (size_t)(seg_row0 + q0 + r) * p.ldo + head * AT_HD + hsel * HD2;
#pragma unroll
      for (int i = 0; i < HD2; i += 8) {
        uint4 u;
        u.x = pack_bf16x2(o[i], o[i + 1]);
        u.y = pack_bf16x2(o[i + 2], o[i + 3]);
        u.z = pack_bf16x2(o[i + 4], o[i + 5]);
        u.w = pack_bf16x2(o[i + 6], o[i + 7]);
        *reinterpret_cast<uint4*>(dst + i) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// Q, K: [heads][rows_total][64] bf16; Vt: [heads][64][rows_total] bf16; out: [rows_total][ldo] bf16.
void attention_tc3(cudaStream_t st, const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* Vt,
                   __nv_bfloat16* out, int ldo, int heads, int rows_total, int seg_len) {
  CRA5_CHECK(seg_len > 0 && rows_total % seg_len == 0, ERR_INVALID, "attention: rows must be whole segments");
  CRA5_CHECK((rows_total & 7) == 0, ERR_INVALID, "attention: rows_total must be a multiple of 8 (TMA stride)");
  static bool configured = false;
  if (!configured) {
    CRA5_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::TOTAL));
    configured = true;
  }
  CUtensorMap tmQ = make_tmap_bf16_2d(Q, AT_HD, (uint64_t)heads * rows_total, AT_HD * 2, AT_HD, AT_BM);
  CUtensorMap tmK = make_tmap_bf16_2d(K, AT_HD, (uint64_t)heads * rows_total, AT_HD * 2, AT_HD, AT_BN);
  CUtensorMap tmVt = make_tmap_bf16_2d(Vt, (uint64_t)rows_total, (uint64_t)heads * AT_HD, (uint64_t)rows_total * 2,
                                       64, AT_HD);
  AttnParams p;
  p.seg_len = seg_len;
  p.rows_total = rows_total;
  p.out = out;
  p.ldo = ldo;
  dim3 grid((seg_len + AT_BM - 1) / AT_BM, rows_total / seg_len, heads);
  LaunchScope scope(st, "attn_tc", 4.0 * heads * (double)rows_total * seg_len * AT_HD,
                    4.0 * 2.0 * heads * (double)rows_total * AT_HD);
  attn_tc_kernel<<<grid, AT_THREADS, AttnSmem::TOTAL, st>>>(tmQ, tmK, tmVt, p);
  CRA5_CUDA(cudaGetLastError());
}

}  // namespace cra5
