// attn_tc.cu -- fused softmax(Q K^T) V on tcgen05 tensor cores for head_dim 64.
//
// Serves both attention flavours of the VAEformer trunk (reference: WindowAttention.forward
// vit_nlc.py:219-258 -- 576-token windows incl. the un-masked zero-pad tokens -- and Attention.forward
// vit_nlc.py:94-112 -- one 10 368-token segment): the token list arrives already in "attention order"
// (window-partitioned by the LayerNorm kernel), so a segment is a contiguous run of seg_len rows.
//
// One CTA = one 128-row query tile of one (segment, head). KV tiles of 128 stream through a 3-stage TMA ring.
//   warp 0   TMA producer        warp 1   MMA issuer (S = Q K^T into TMEM, then O_j = P V into TMEM)
//   warp 2   TMEM allocator      warps 4-7  softmax: one query row per thread, online max/sum in fp32,
//                                           P written to swizzled smem as bf16, O accumulated in registers.
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

constexpr int AT_BM = 128;
constexpr int AT_BN = 128;
constexpr int AT_HD = 64;
constexpr int AT_STAGES = 2;   // two CTAs share an SM: 2 x (16 KB Q + 2 x 32 KB KV + 32 KB P) = 224 KB
constexpr int AT_THREADS = 256;

struct AttnSmem {
  static constexpr int Q_BYTES = AT_BM * AT_HD * 2;        // 16 KB
  static constexpr int K_BYTES = AT_BN * AT_HD * 2;        // 16 KB
  static constexpr int V_BYTES = AT_HD * AT_BN * 2;        // 16 KB (two 64x64 halves)
  static constexpr int KV_BYTES = K_BYTES + V_BYTES;
  static constexpr int P_BYTES = AT_BM * AT_BN * 2;        // 32 KB (two 128x64 halves)
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = OFF_Q + Q_BYTES;
  static constexpr int OFF_P = OFF_KV + AT_STAGES * KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int TOTAL = OFF_BAR + 128;  // no alignment slack: the dynamic window starts 1024-aligned (checked)
};

struct AttnParams {
  int seg_len;      // tokens per segment (window size or whole sequence)
  int rows_total;   // total rows of Q/K per head (= row stride of Vt)
  __nv_bfloat16* out;  // [rows_total, ldo] bf16, head h occupies columns [h*64, h*64+64)
  int ldo;
};

__global__ void __launch_bounds__(AT_THREADS, 2)
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
  uint64_t* p_full = s_full + 1;
  uint64_t* pv_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_full + 2);

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
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    for (int s = 0; s < 2; ++s) mbar_init(&pv_full[s], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t TM_S = 0;     // 128 columns (single buffer: the second CTA on the SM fills the bubbles)
  constexpr uint32_t TM_PV = 128;  // 2 x 64 columns

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
      auto issue_s = [&](int j) {
        const int s = j % AT_STAGES;
        mbar_wait(&kv_full[s], (j / AT_STAGES) & 1);
        tc_fence_after();
        const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(smem + L::OFF_KV + s * L::KV_BYTES));
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k)
          umma_bf16(tmem_base + TM_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(p_full, j & 1);   // softmax has consumed S_j and published P_j
        tc_fence_after();
        const int s = j % AT_STAGES;
        const uint32_t sv = smem_u32(smem + L::OFF_KV + s * L::KV_BYTES + L::K_BYTES);
        const uint32_t sp = smem_u32(smem + L::OFF_P);
#pragma unroll
        for (int kk = 0; kk < AT_BN / 16; ++kk) {
          const int half = kk >> 2, k = kk & 3;
          const uint64_t pdesc = umma_smem_desc_sw128(sp + half * (L::P_BYTES / 2)) + 2 * k;
          const uint64_t vdesc = umma_smem_desc_sw128(sv + half * (L::V_BYTES / 2)) + 2 * k;
          umma_bf16(tmem_base + TM_PV + (j & 1) * AT_HD, pdesc, vdesc, idesc_pv, kk != 0);
        }
        umma_commit(&pv_full[j & 1]);
        umma_commit(&kv_empty[s]);
        if (j + 1 < n_kv) issue_s(j + 1);  // S is single-buffered: next scores only after P_j was read out of S
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + output accumulation =====================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row inside the query tile == TMEM lane
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f;
    float m = -INFINITY, l = 0.f;
    float o[AT_HD];
#pragma unroll
    for (int i = 0; i < AT_HD; ++i) o[i] = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const int valid = min(AT_BN, p.seg_len - j * AT_BN);
      const bool full_tile = (valid == AT_BN);   // warp-uniform: only a segment's last tile can be partial
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem_base + lane_addr + TM_S;
      // pass 1: row maximum
      float mx = m;
#pragma unroll 1
      for (int c = 0; c < AT_BN; c += 32) {
        uint32_t sv[32];
        tmem_ld_32x32(s_addr + c, sv);
        tmem_ld_wait();
        if (full_tile) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i < valid) mx = fmaxf(mx, __uint_as_float(sv[i]));
        }
      }
      const float alpha = ex2_approx((m - mx) * LOG2E);  // 0 on the first tile (m = -inf)
      const float mxl = mx * LOG2E;
      // pass 2: probabilities -> bf16 P tile in smem (K-major, SWIZZLE_128B)
      float rowsum = 0.f;
      uint8_t* pbase = smem + L::OFF_P;
#pragma unroll 1
      for (int c = 0; c < AT_BN; c += 32) {
        uint32_t sv[32];
        tmem_ld_32x32(s_addr + c, sv);
        tmem_ld_wait();
        float pr[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) pr[i] = ex2_approx(fmaf(__uint_as_float(sv[i]), LOG2E, -mxl));
        if (!full_tile) {
#pragma unroll
          for (int i = 0; i < 32; ++i) pr[i] = (c + i < valid) ? pr[i] : 0.f;
        }
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 4) { s0 += pr[i]; s1 += pr[i + 1]; s2 += pr[i + 2]; s3 += pr[i + 3]; }
        rowsum += (s0 + s1) + (s2 + s3);
        uint8_t* region = pbase + (c >> 6) * (L::P_BYTES / 2);
        const int chunk0 = (c & 63) >> 3;  // 0 or 4
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_bf16x2(pr[8 * q + 0], pr[8 * q + 1]);
          u.y = pack_bf16x2(pr[8 * q + 2], pr[8 * q + 3]);
          u.z = pack_bf16x2(pr[8 * q + 4], pr[8 * q + 5]);
          u.w = pack_bf16x2(pr[8 * q + 6], pr[8 * q + 7]);
          *reinterpret_cast<uint4*>(region + sw128_offset(r, chunk0 + q)) = u;
        }
      }
      l = l * alpha + rowsum;
      m = mx;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      // fold the previous tile's P V product, then rescale everything to the new maximum
      if (j > 0) {
        mbar_wait(&pv_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        const uint32_t o_addr = tmem_base + lane_addr + TM_PV + ((j - 1) & 1) * AT_HD;
#pragma unroll
        for (int c = 0; c < AT_HD; c += 32) {
          uint32_t pv[32];
          tmem_ld_32x32(o_addr + c, pv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c + i] = (o[c + i] + __uint_as_float(pv[i])) * alpha;
        }
      }
    }
    {
      const int j = n_kv - 1;
      mbar_wait(&pv_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t o_addr = tmem_base + lane_addr + TM_PV + (j & 1) * AT_HD;
      const float inv = 1.0f / l;
#pragma unroll
      for (int c = 0; c < AT_HD; c += 32) {
        uint32_t pv[32];
        tmem_ld_32x32(o_addr + c, pv);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c + i] = (o[c + i] + __uint_as_float(pv[i])) * inv;
      }
    }
    if (q0 + r < p.seg_len) {
      __nv_bfloat16* dst = p.out + (size_t)(seg_row0 + q0 + r) * p.ldo + head * AT_HD;
#pragma unroll
      for (int i = 0; i < AT_HD; i += 8) {
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
    tmem_dealloc<256>(tmem_base);
  }
}

// Q, K: [heads][rows_total][64] bf16; Vt: [heads][64][rows_total] bf16; out: [rows_total][ldo] bf16.
void attention_tc(cudaStream_t st, const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* Vt,
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
