// gemm_tc.cuh -- persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = A[M,K] * B[N,K]^T      A, B bf16 K-major in HBM, fp32 accumulation in TMEM
//
// Roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warp 3 idle, warps 4..11 = epilogue (two warps per TMEM lane quarter,
// each taking half of the tile's columns). Operand tiles are 128 x 64 (A) and BN x 64 (B) bf16,
// SWIZZLE_128B, 4-stage mbarrier ring; the accumulator is double-buffered in TMEM so the epilogue of
// tile i overlaps the main loop of tile i+1. M/N/K tails are handled by TMA zero fill + masked stores.
//
// This one main loop serves every linear layer on the VAEformer hot path (reference call sites:
// vit_nlc.py:96,242 qkv; :111,248 proj; :63-67 fc1/fc2; vaeformer.py:154-155 quant/post_quant conv;
// vit_nlc.py:302 patch-embed conv; :629 ConvTranspose2d; :741 hyperprior head) through the A-loader modes
// and the epilogue functors below.
#pragma once
#include "ptx.cuh"

namespace cra5 {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_WARP0 = 4;

// how the producer addresses the A operand
enum AMode : int {
  A_PLAIN = 0,   // 2D map (k, row)
  A_PATCH = 1,   // 3D map (cs, j, h) over the re-laid-out frame: implicit im2col for the patch-embed conv
  A_CONCAT = 2,  // 2D map; K = 2*D, second half reads rows shifted by -cc_shift (row-overlap of ConvTranspose2d)
};

struct GemmShape {
  int M, N, K;
  int a_mode;
  // A_PATCH
  int pe_kpr;       // k-blocks per kernel row r
  int pe_box_rows;  // tokens per TMA box (divides tokens-per-row and 128)
  int pe_Wp;        // tokens per patch row
  int pe_sh;        // vertical stride in pixels
  int pe_Hg;        // patch rows per frame and pixel rows per frame: token row i of a batch maps to frame i / pe_Hg,
  int pe_img_h;     //   pixel row pe_sh * (i % pe_Hg) + (i / pe_Hg) * pe_img_h   (pe_Hg <= 0: single frame)
  // A_CONCAT
  int cc_D;      // split point in K (multiple of 64)
  int cc_shift;  // row shift for the second half
  // split-bf16 precision ("bf16x3", 1-CTA kernel only): an fp32 operand v is carried as hi = bf16(v), lo = bf16(v - hi),
  // stored as two halves addressed by ONE extra (outermost) tensor-map coordinate. With both operands split every
  // k-block is issued three times -- (A_hi, B_hi), (A_lo, B_hi), (A_hi, B_lo) -- into the same fp32 accumulator, which
  // recovers ~16 mantissa bits per operand (the dropped lo*lo term is 2^-18 relative); with only B split (A is bf16 by
  // construction, e.g. the attention output) twice. The MMA issuer and the epilogues are unchanged: the producer
  // simply feeds more k-blocks.
  int a_split, b_split;
};

// Epilogue index arithmetic: the window map of a residual tile row is computed once per tile (not once per 32-column
// chunk), its four integer divisions (about 20 instructions each: I2F, MUFU.RCP, F2I, fix-up) are multiply-high + shift
// by host-precomputed constants, the QKV head split uses compares and a shift, and the un-patchify epilogue advances
// (i, j) incrementally instead of dividing per row. (Validated bit-identical against the division forms on a B200,
// 2.7 % of the frame time; round 2.)

// n / d for 0 <= n < 2^31 and a positive divisor fixed at launch time (the CUTLASS FastDivmod construction:
// m = ceil(2^(31 + ceil(log2 d)) / d); verified exhaustively for the divisors that occur here)
struct FastDiv {
  uint32_t mul, shr;
  int d;
  static FastDiv make(int d) {
    FastDiv f;
    f.d = d;
    f.mul = 0;
    f.shr = 0;
    if (d > 1) {
      int l = 0;
      while ((1u << l) < (uint32_t)d) ++l;   // ceil(log2 d)
      const int p = 31 + l;
      f.mul = (uint32_t)(((1ull << p) + (uint64_t)d - 1) / (uint64_t)d);
      f.shr = (uint32_t)(p - 32);
    }
    return f;
  }
  __device__ __forceinline__ int div(int n) const { return d == 1 ? n : (int)(__umulhi((uint32_t)n, mul) >> shr); }
};

// maps a row of the "attention order" (window-partitioned, zero-padded) token list back to the raster token
struct WinMap {
  int enabled;
  int H, W;        // token grid
  int wh, ww;      // window
  int nWr, nWc;    // windows per column / row after padding
  FastDiv f_wsz, f_ww, f_pf, f_nwc;
  void finish() {   // host: call after the integer fields are set
    f_wsz = FastDiv::make(wh * ww);
    f_ww = FastDiv::make(ww);
    f_pf = FastDiv::make(nWr * nWc);
    f_nwc = FastDiv::make(nWc);
  }
  __device__ __forceinline__ int to_token(int a) const {  // -1 for a pad row
    if (!enabled) return a;
    const int wsz = wh * ww;
    int wi = f_wsz.div(a), within = a - wi * wsz;
    int r = f_ww.div(within), c = within - r * ww;
    int per_frame = nWr * nWc;
    int b = f_pf.div(wi);
    wi -= b * per_frame;
    int wr = f_nwc.div(wi), wc = wi - wr * nWc;
    int h = wr * wh + r, w = wc * ww + c;
    if (h >= H || w >= W) return -1;
    return (b * H + h) * W + w;
  }
};

// QKV column -> (which of q|k|v, head, dim): divisions by the model width and the head width (EPI_QKV epilogues)
struct QkvSplit { int which, head, d; };
__device__ __forceinline__ QkvSplit qkv_split(int col, int D, int hd) {
  QkvSplit q;
  q.which = (col >= 2 * D) ? 2 : (col >= D ? 1 : 0);
  const int within = col - q.which * D;
  q.head = ((hd & (hd - 1)) == 0) ? (within >> (31 - __clz(hd))) : within / hd;   // warp-uniform choice
  q.d = within - q.head * hd;
  return q;
}

enum EpiKind : int {
  EPI_F32 = 0,        // out_f32 = acc + bias (+ add)
  EPI_BF16 = 1,       // out_bf16 = acc + bias
  EPI_GELU_BF16 = 2,  // out_bf16 = gelu_erf(acc + bias)
  EPI_QKV = 3,        // head split: Q (pre-scaled), K as [head][row][hd]; V transposed [head][hd][row]
  EPI_RESID = 4,      // out_f32[t] = resid[t] + acc + bias, t = winmap(row); optional bf16 copy
  EPI_T_F32 = 5,      // out_f32[col * ldo + row] = acc + bias  (channel-major / NCHW result)
  EPI_PIXSHUF = 6,    // hyperprior head: (p1 p2 c) pixel shuffle into NCHW
  EPI_CONVT = 7,      // un-patchify scatter into NCHW
  EPI_QKV_F16 = 8,    // EPI_QKV with Q, K, V written as fp16 (the attention of the split-precision levels; the
                      // reference's own GPU path runs attention on fp16 operands, vit_nlc.py:105-110). 1-CTA kernel only.
};
template <int KIND>
__device__ __forceinline__ constexpr bool is_qkv() { return KIND == EPI_QKV || KIND == EPI_QKV_F16; }

struct EpiParams {
  const float* bias;      // [N] or null
  const float* add;       // EPI_F32: extra addend [M, lda] (pos-embed), or null
  int lda;
  int add_period;         // > 0: the addend has add_period rows and repeats down the batch (row % add_period)
  // batches of frames (rows = frame * rows_per_frame + r): EPI_T_F32 / EPI_PIXSHUF / EPI_CONVT write frame b of the
  // output at out_f32 + b * fr_stride; fr_rows = rows per frame (<= 0: a single frame)
  int fr_rows;
  size_t fr_stride;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  int ldo;                // row stride of out (elements)
  // EPI_RESID
  const float* resid;
  WinMap wm;
  int bf16_col0;          // column offset of the optional bf16 copy (row stride ld_bf16)
  int ld_bf16;
  // EPI_QKV
  __nv_bfloat16 *q, *k, *vt;
  int D, hd, rows_total;  // model dim, head dim, number of rows (= row stride of vt)
  float qscale;
  // EPI_PIXSHUF
  int ps_P1, ps_P2, ps_C, ps_Wh;  // out[c][(P1*i+p1)*(P2*Wh) + P2*j+p2], row=(i*Wh+j), col=(p1*P2+p2)*C+c
  // EPI_CONVT
  int ct_r0;     // first kernel row covered by this GEMM
  int ct_CS;     // C*pw
  int ct_pw;     // patch width (== horizontal stride)
  int ct_sh;     // vertical stride
  int ct_Wp;     // tokens per row
  int ct_Himg, ct_Wimg;
  const float *ct_mean, *ct_std;   // optional per-channel de-normalisation fused into the store: x * std[c] + mean[c]
                                   // (cra5_api.de_normalization, cra5_api.py:268-271); null = normalised units
  int gelu_fast;                   // EPI_GELU_BF16: 0 = erf by A&S 7.1.26 (gelu_erf), 1 = gelu_bf16out (trunk fc1)
  // EPI_CONVT, channel-grouped column order (ct_cpg > 0): within a kernel row the N columns come in groups of 32 =
  // ct_cpg whole channels x pw (+ zero padding), ct_CS = 32 * groups per kernel row; see epilogue_convt_grouped
  int ct_cpg;
  int ct_C;                        // number of real channels (groups are padded up)
};

// exact (erf) GELU, nn.GELU default (vit_nlc.py:53). erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, i.e. fp32
// round-off level, branch-free, 2 MUFU + ~12 FMA-pipe instructions; the libm erff costs about twice that and made the
// fc1 epilogue longer than its K=1024 main loop).
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  const float erf_abs = fmaf(-p, e, 1.0f);            // erf(|x|/sqrt2)
  return 0.5f * x + 0.5f * fabsf(x) * erf_abs;          // 0.5 x (1 + sign(x) erf(|x|/sqrt2))
}

// The same function for bf16 OUTPUTS (fc1's epilogue): x * Phi(x) with Phi(x) = 0.5 (1 + tanh(x (c0 + c1 x^2 + c2 x^4))),
// coefficients fitted to the erf form (max |error| 2.5e-5 over the real line in exact arithmetic; this is NOT the
// "tanh GELU" of the literature, whose two-term fit is off by 3e-4) and one MUFU.TANH: 1 MUFU + 7 FMA-pipe instructions
// per element instead of 2 + 12. Measured against fp64 erf-GELU on a B200 (tools/micro/tanh_err.cu): see
// profiles/r2_micro_tanh.txt -- the error stays an order of magnitude below the bf16 rounding of the stored value.
// Every fp32-output and split-precision path keeps gelu_erf.
__device__ __forceinline__ float gelu_bf16out(float x) {
  const float x2 = x * x;
  const float u = x * fmaf(x2, fmaf(x2, -0.0003515175339619918f, 0.037005650955991044f), 0.797507878425557f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}

// 16-bit output element of an epilogue kind: bf16, or fp16 bits travelling in bf16-typed buffers for EPI_QKV_F16
template <int KIND>
__device__ __forceinline__ uint32_t pack_pair(float lo, float hi) {
  if constexpr (KIND == EPI_QKV_F16) {
    const __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
  } else {
    return pack_bf16x2(lo, hi);
  }
}
template <int KIND>
__device__ __forceinline__ __nv_bfloat16 cvt_elem(float v) {
  if constexpr (KIND == EPI_QKV_F16) {
    const __half h = __float2half_rn(v);
    return *reinterpret_cast<const __nv_bfloat16*>(&h);
  } else {
    return __float2bfloat16(v);
  }
}

template <int KIND>
__device__ __forceinline__ void epilogue_store(const EpiParams& p, int row, int col0, const uint32_t (&acc)[32],
                                               int M, int N) {
  if (row >= M) return;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    v[i] = __uint_as_float(acc[i]);
    if (p.bias != nullptr && col0 + i < N) v[i] += __ldg(p.bias + col0 + i);
  }
  const bool full = (col0 + 32 <= N);
  if constexpr (KIND == EPI_F32) {
    float* o = p.out_f32 + (size_t)row * p.ldo + col0;
    if (p.add != nullptr) {
      const float* a = p.add + (size_t)(p.add_period > 0 ? row % p.add_period : row) * p.lda + col0;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < N) v[i] += __ldg(a + i);
    }
    if (full && (p.ldo & 3) == 0) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
      for (int i = 0; i < 32; ++i)
        if (col0 + i < N) o[i] = v[i];
    }
  } else if constexpr (KIND == EPI_BF16 || KIND == EPI_GELU_BF16) {
    if constexpr (KIND == EPI_GELU_BF16) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
    }
    __nv_bfloat16* o = p.out_bf16 + (size_t)row * p.ldo + col0;
    if (full && (p.ldo & 7) == 0) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 u;
        u.x = pack_bf16x2(v[i], v[i + 1]);
        u.y = pack_bf16x2(v[i + 2], v[i + 3]);
        u.z = pack_bf16x2(v[i + 4], v[i + 5]);
        u.w = pack_bf16x2(v[i + 6], v[i + 7]);
        *reinterpret_cast<uint4*>(o + i) = u;
      }
    } else {
      for (int i = 0; i < 32; ++i)
        if (col0 + i < N) o[i] = __float2bfloat16(v[i]);
    }
  } else if constexpr (is_qkv<KIND>()) {
    // columns ordered [q|k|v][head][dim] (vit_nlc.py:99,242)
#pragma unroll
    for (int i0 = 0; i0 < 32; i0 += 8) {
      const int col = col0 + i0;
      if (col >= N) continue;
      const int which = col / p.D;
      const int within = col - which * p.D;
      const int head = within / p.hd;
      const int d = within - head * p.hd;
      const bool vec = ((p.hd | p.D) & 7) == 0 && (col + 8 <= N);  // 8 columns stay inside one head
      if (which == 2) {
        if (vec && (p.rows_total & 1) == 0 && (M & 1) == 0) {
          // V^T[d][row]: neighbouring lanes hold neighbouring rows; even lanes write rows (row, row+1) of column i, odd
          // lanes rows (row-1, row) of column i+1 -> one 4-byte store per lane covers two rows
          const bool odd = threadIdx.x & 1;
          const unsigned am = __activemask();  // rows >= M have returned; M is even, so lane pairs stay together
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            const float mine0 = v[i0 + i], mine1 = v[i0 + i + 1];
            const float other0 = __shfl_xor_sync(am, mine0, 1), other1 = __shfl_xor_sync(am, mine1, 1);
            const int dcol = d + i + (odd ? 1 : 0);
            const float lo = odd ? other1 : mine0, hi = odd ? mine1 : other0;
            const int r0_ = odd ? row - 1 : row;
            *reinterpret_cast<uint32_t*>(p.vt + ((size_t)head * p.hd + dcol) * p.rows_total + r0_) = pack_pair<KIND>(lo, hi);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (col + i < N) {
              int c2 = col + i - 2 * p.D;
              int h2 = c2 / p.hd, d2 = c2 - h2 * p.hd;
              p.vt[((size_t)h2 * p.hd + d2) * p.rows_total + row] = cvt_elem<KIND>(v[i0 + i]);
            }
        }
      } else {
        __nv_bfloat16* dst = (which == 0 ? p.q : p.k) + ((size_t)head * p.rows_total + row) * p.hd + d;
        const float s = (which == 0) ? p.qscale : 1.0f;
        if (vec) {
          uint4 u;
          u.x = pack_pair<KIND>(v[i0] * s, v[i0 + 1] * s);
          u.y = pack_pair<KIND>(v[i0 + 2] * s, v[i0 + 3] * s);
          u.z = pack_pair<KIND>(v[i0 + 4] * s, v[i0 + 5] * s);
          u.w = pack_pair<KIND>(v[i0 + 6] * s, v[i0 + 7] * s);
          *reinterpret_cast<uint4*>(dst) = u;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (col + i < N) {
              int c2 = col + i - which * p.D;
              int h2 = c2 / p.hd, d2 = c2 - h2 * p.hd;
              ((which == 0 ? p.q : p.k))[((size_t)h2 * p.rows_total + row) * p.hd + d2] =
                  cvt_elem<KIND>(v[i0 + i] * s);
            }
        }
      }
    }
  } else if constexpr (KIND == EPI_RESID) {
    const int t = p.wm.to_token(row);
    if (t < 0) return;
    const float* r = p.resid + (size_t)t * p.ldo + col0;
    float* o = p.out_f32 + (size_t)t * p.ldo + col0;
    if (full && (p.ldo & 3) == 0) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 rr = *reinterpret_cast<const float4*>(r + i);
        v[i] += rr.x; v[i + 1] += rr.y; v[i + 2] += rr.z; v[i + 3] += rr.w;
        *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    } else {
      for (int i = 0; i < 32; ++i)
        if (col0 + i < N) { v[i] += r[i]; o[i] = v[i]; }
    }
    if (p.out_bf16 != nullptr) {
      __nv_bfloat16* ob = p.out_bf16 + (size_t)t * p.ld_bf16 + p.bf16_col0 + col0;
      for (int i = 0; i < 32; ++i)
        if (col0 + i < N) ob[i] = __float2bfloat16(v[i]);
    }
  } else if constexpr (KIND == EPI_T_F32) {
    float* o = p.out_f32 + row;
    if (p.fr_rows > 0) {   // batch: frame b's [N][ldo] matrix starts at b * fr_stride
      const int b = row / p.fr_rows;
      o = p.out_f32 + (size_t)b * p.fr_stride + (row - b * p.fr_rows);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (col0 + i < N) o[(size_t)(col0 + i) * p.ldo] = v[i];
  } else if constexpr (KIND == EPI_PIXSHUF) {
    int rf = row;
    float* outb = p.out_f32;
    if (p.fr_rows > 0) {
      const int b = row / p.fr_rows;
      rf = row - b * p.fr_rows;
      outb += (size_t)b * p.fr_stride;
    }
    const int i_ = rf / p.ps_Wh, j_ = rf - i_ * p.ps_Wh;
    const int Wout = p.ps_P2 * p.ps_Wh;
    for (int i = 0; i < 32; ++i) {
      const int col = col0 + i;
      if (col >= N) break;
      const int pp = col / p.ps_C, c = col - pp * p.ps_C;
      const int p1 = pp / p.ps_P2, p2 = pp - p1 * p.ps_P2;
      outb[(size_t)c * p.ldo + (size_t)(p.ps_P1 * i_ + p1) * Wout + p.ps_P2 * j_ + p2] = v[i];
    }
  } else if constexpr (KIND == EPI_CONVT) {
    // row = (i, j) patch position; col = (r - r0, c, s); out[c][sh*i + r][pw*j + s]
    int rf = row;
    float* outb = p.out_f32;
    if (p.fr_rows > 0) {
      const int b = row / p.fr_rows;
      rf = row - b * p.fr_rows;
      outb += (size_t)b * p.fr_stride;
    }
    const int i_ = rf / p.ct_Wp, j_ = rf - i_ * p.ct_Wp;
    for (int i = 0; i < 32; ++i) {
      const int col = col0 + i;
      if (col >= N) break;
      const int rr = col / p.ct_CS, cs = col - rr * p.ct_CS;
      const int c = cs / p.ct_pw, s = cs - c * p.ct_pw;
      const int h = p.ct_sh * i_ + p.ct_r0 + rr;
      if (h < p.ct_Himg)
        outb[((size_t)c * p.ct_Himg + h) * p.ct_Wimg + p.ct_pw * j_ + s] =
            (p.ct_std != nullptr) ? fmaf(v[i], __ldg(p.ct_std + c), __ldg(p.ct_mean + c)) : v[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Row-wise epilogue: a warp's 32x32 accumulator chunk has been staged in shared memory (stg[r * 33 + c]); lanes now
// run along the COLUMNS of one output row at a time, so every global access of the warp is one contiguous
// 128-byte (fp32) / 64-byte (bf16) segment instead of 32 scattered ones.
constexpr int STG_LD = 33;  // padded row stride (words): conflict-free both for the row writes and the column reads

// bf16 row stores, two rows per step: lanes 0-15 take row rr, lanes 16-31 row rr+1, two adjacent columns each
template <int KIND>
__device__ __forceinline__ void store_rows_bf16x2(const float* stg, __nv_bfloat16* base, size_t ld, int rows_valid,
                                                  int lane, float b0, float b1, float scale) {
  const int half = lane >> 4, l2 = (lane & 15) * 2;
  float a0[16], a1[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {  // all shared-memory reads first: independent, so their latencies overlap
    a0[k] = stg[(2 * k + half) * STG_LD + l2];
    a1[k] = stg[(2 * k + half) * STG_LD + l2 + 1];
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int rr = 2 * k + half;
    if (rr < rows_valid)
      *reinterpret_cast<uint32_t*>(base + (size_t)rr * ld + l2) = pack_pair<KIND>((a0[k] + b0) * scale, (a1[k] + b1) * scale);
  }
}

template <int KIND>
__device__ __forceinline__ void epilogue_rows(const EpiParams& p, const float* stg, int row_base, int col0, int lane,
                                              int M, int N) {
  const int rows_valid = min(32, M - row_base);
  if (rows_valid <= 0) return;
  if constexpr (KIND == EPI_F32) {
    const int col = col0 + lane;
    const bool ok = col < N;
    const float b = (p.bias != nullptr && ok) ? __ldg(p.bias + col) : 0.f;
    const int colc = ok ? col : 0;
    if (p.add != nullptr) {  // all loads first, from always-valid (clamped) addresses so nothing waits on a select
      float av[32];
#pragma unroll
      for (int rr = 0; rr < 32; ++rr)
        av[rr] = __ldg(p.add + (size_t)((p.add_period > 0) ? (row_base + min(rr, rows_valid - 1)) % p.add_period
                                                            : (row_base + min(rr, rows_valid - 1))) * p.lda + colc);
#pragma unroll
      for (int rr = 0; rr < 32; ++rr)
        if (ok && rr < rows_valid) p.out_f32[(size_t)(row_base + rr) * p.ldo + col] = stg[rr * STG_LD + lane] + b + av[rr];
    } else {
#pragma unroll
      for (int rr = 0; rr < 32; ++rr)
        if (ok && rr < rows_valid) p.out_f32[(size_t)(row_base + rr) * p.ldo + col] = stg[rr * STG_LD + lane] + b;
    }
  } else if constexpr (KIND == EPI_BF16 || KIND == EPI_GELU_BF16) {
    // two rows per step: lanes 0-15 take row rr, lanes 16-31 row rr+1, two adjacent columns each (4-byte stores)
    const int half = lane >> 4, l2 = (lane & 15) * 2;
    const int col = col0 + l2;
    const float b0 = (p.bias != nullptr && col < N) ? __ldg(p.bias + col) : 0.f;
    const float b1 = (p.bias != nullptr && col + 1 < N) ? __ldg(p.bias + col + 1) : 0.f;
    const bool pair_ok = (col + 1 < N) && ((p.ldo & 1) == 0);
    float a0[16], a1[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {  // all shared-memory reads first: independent, so their latencies overlap
      a0[k] = stg[(2 * k + half) * STG_LD + l2];
      a1[k] = stg[(2 * k + half) * STG_LD + l2 + 1];
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int rr = 2 * k + half;
      float v0 = a0[k] + b0, v1 = a1[k] + b1;
      if constexpr (KIND == EPI_GELU_BF16) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
      if (rr < rows_valid) {
        __nv_bfloat16* o = p.out_bf16 + (size_t)(row_base + rr) * p.ldo + col;
        if (pair_ok) {
          *reinterpret_cast<uint32_t*>(o) = pack_bf16x2(v0, v1);
        } else {
          if (col < N) o[0] = __float2bfloat16(v0);
          if (col + 1 < N) o[1] = __float2bfloat16(v1);
        }
      }
    }
  } else if constexpr (is_qkv<KIND>()) {
    // columns ordered [q|k|v][head][dim] (vit_nlc.py:99,242)
    if (((p.D | p.hd) & 31) == 0 && col0 + 32 <= N) {
      // the whole 32-column chunk lies inside one head of Q or K (V chunks take the direct path): 4-byte stores
      const int which = col0 / p.D;
      const int within = col0 - which * p.D;
      const int head = within / p.hd, d0 = within - head * p.hd;
      const int l2 = (lane & 15) * 2;
      const float b0 = (p.bias != nullptr) ? __ldg(p.bias + col0 + l2) : 0.f;
      const float b1 = (p.bias != nullptr) ? __ldg(p.bias + col0 + l2 + 1) : 0.f;
      __nv_bfloat16* base = (which == 0 ? p.q : p.k) + ((size_t)head * p.rows_total + row_base) * p.hd + d0;
      store_rows_bf16x2<KIND>(stg, base, (size_t)p.hd, rows_valid, lane, b0, b1, which == 0 ? p.qscale : 1.0f);
      return;
    }
    // generic: each lane owns one column for all rows of the chunk
    const int col = col0 + lane;
    if (col >= N) return;
    const float b = (p.bias != nullptr) ? __ldg(p.bias + col) : 0.f;
    const int which = col / p.D;
    const int within = col - which * p.D;
    const int head = within / p.hd, d = within - head * p.hd;
    if (which == 2) {  // (only reached when a chunk straddles the k|v boundary; whole-V chunks take the direct path)
      __nv_bfloat16* dst = p.vt + ((size_t)head * p.hd + d) * p.rows_total + row_base;
#pragma unroll
      for (int rr = 0; rr < 32; ++rr)
        if (rr < rows_valid) dst[rr] = cvt_elem<KIND>(stg[rr * STG_LD + lane] + b);
    } else {
      const float sc = (which == 0) ? p.qscale : 1.0f;
      __nv_bfloat16* dst = (which == 0 ? p.q : p.k) + ((size_t)head * p.rows_total + row_base) * p.hd + d;
#pragma unroll
      for (int rr = 0; rr < 32; ++rr)
        if (rr < rows_valid) dst[(size_t)rr * p.hd] = cvt_elem<KIND>((stg[rr * STG_LD + lane] + b) * sc);
    }
  } else if constexpr (KIND == EPI_RESID) {
    // handled by epilogue_resid_load / epilogue_resid_finish (the residual loads are issued before the TMEM read)
  } else if constexpr (KIND == EPI_CONVT) {
    // row = (i, j) patch position; col = (r - r0, c, s); out[c][sh*i + r][pw*j + s]: a lane's column fixes (r, c, s)
    const int col = col0 + lane;
    if (col >= N) return;
    const int rk = col / p.ct_CS, cs = col - rk * p.ct_CS;
    const int c = cs / p.ct_pw, s_ = cs - c * p.ct_pw;
    float* plane = p.out_f32 + (size_t)c * p.ct_Himg * p.ct_Wimg + s_;
    const float dn_s = (p.ct_std != nullptr) ? __ldg(p.ct_std + c) : 1.0f;     // a lane's column fixes its channel
    const float dn_m = (p.ct_std != nullptr) ? __ldg(p.ct_mean + c) : 0.0f;
    int rf = row_base;
    const int i_per = (p.fr_rows > 0) ? p.fr_rows / p.ct_Wp : (1 << 30);   // patch rows per frame (batch of frames)
    if (p.fr_rows > 0) {
      const int b = row_base / p.fr_rows;
      rf = row_base - b * p.fr_rows;
      plane += (size_t)b * p.fr_stride;
    }
    int i_ = rf / p.ct_Wp, j_ = rf - i_ * p.ct_Wp;   // one division per chunk; (i, j) advance with the row
#pragma unroll 8
    for (int rr = 0; rr < rows_valid; ++rr) {
      const int h = p.ct_sh * i_ + p.ct_r0 + rk;
      if (h < p.ct_Himg)
        plane[(size_t)h * p.ct_Wimg + p.ct_pw * j_] =
            (p.ct_std != nullptr) ? fmaf(stg[rr * STG_LD + lane], dn_s, dn_m) : stg[rr * STG_LD + lane];
      if (++j_ == p.ct_Wp) {
        j_ = 0;
        if (++i_ == i_per) { i_ = 0; plane += p.fr_stride; }
      }
    }
  }
}

// EPI_CONVT with the channel-grouped column order. The plain order (r, c, s) makes a warp's 32 columns 3.2 channels, and
// a store instruction (lane = column, one token) 3-4 runs of 40 bytes in different channel planes, with ~30
// instructions of index arithmetic per stored row: ncu showed the epilogue warps issue-bound (34 % tensor pipe) and
// 12.3 GB of DRAM traffic for 8.2 GB of output (every 32-byte sector written in pieces, so L2 fetches it first). Here a
// 32-column chunk is CPG WHOLE channels (3 x 10 columns + 2 zero-weight pad columns for pw = 10; the weight is packed
// that way at upload, +7 % MMA work), the chunk is staged column-major, and for each channel the warp's 32 tokens x pw
// values -- contiguous in the output row while the tokens stay in one image row -- leave as pw full 128-byte lines
// (lane = position in that run). Everything that depends on the lane's position only -- pixel offset, staging index,
// validity -- is computed once per tile (ConvtLane); a store costs LDS + FFMA + address add + STG.
constexpr int CT_LD = 35;   // staged column stride (words): 35 = 3 mod 32 makes the (token, s) reads conflict-free
constexpr int CT_PW = 10;   // patch width the grouped path is built for (every shipped geometry)
__device__ __forceinline__ uint32_t convt_rowword(const EpiParams& p, int row, int M) {
  if (row >= M) return 0xffffffffu;
  int b = 0, rf = row;
  if (p.fr_rows > 0) { b = row / p.fr_rows; rf = row - b * p.fr_rows; }
  const int i_ = rf / p.ct_Wp, j_ = rf - i_ * p.ct_Wp;
  return (uint32_t)(p.ct_sh * i_ * p.ct_Wimg + p.ct_pw * j_) | ((uint32_t)b << 20);   // checked on the host: 20 bits
}
template <int KIND>
struct ConvtLane {};                 // nothing to carry for the other epilogue kinds
template <>
struct ConvtLane<EPI_CONVT> {
  size_t pix[CT_PW];                 // output element offset of this lane's k-th value inside a channel's row class
  uint32_t offs[CT_PW];              // the same inside its frame (validity: offs + kernel-row offset < plane size)
  int sidx[CT_PW];                   // staging index of the value inside a channel's pw staged columns
};
__device__ __forceinline__ void convt_lane_setup(const EpiParams& p, int row, int M, int lane, ConvtLane<EPI_CONVT>& cl) {
  const uint32_t own = convt_rowword(p, row, M);
  int rr = lane / CT_PW, s_ = lane - rr * CT_PW;   // store k covers positions 32 k + lane of the warp's 32 x pw run
#pragma unroll
  for (int k = 0; k < CT_PW; ++k) {
    const uint32_t w = __shfl_sync(0xffffffffu, own, rr);
    const uint32_t off = (w & 0xfffffu) + (uint32_t)s_;
    cl.offs[k] = (w == 0xffffffffu) ? 0xffffffffu : off;
    cl.pix[k] = (size_t)(w >> 20) * p.fr_stride + off;
    cl.sidx[k] = s_ * CT_LD + rr;
    rr += 32 / CT_PW; s_ += 32 % CT_PW;
    if (s_ >= CT_PW) { s_ -= CT_PW; ++rr; }
  }
}
__device__ __forceinline__ void epilogue_convt_grouped(const EpiParams& p, const uint32_t (&acc)[32], float* stg,
                                                       const ConvtLane<EPI_CONVT>& cl, int col0, int lane) {
  constexpr int CPG = 30 / CT_PW;
#pragma unroll
  for (int i = 0; i < CPG * CT_PW; ++i) stg[i * CT_LD + lane] = __uint_as_float(acc[i]);   // (pad columns are never read)
  __syncwarp();
  const int rk = col0 / p.ct_CS;
  const int g = (col0 - rk * p.ct_CS) >> 5;
  const uint32_t row_off = (uint32_t)((p.ct_r0 + rk) * p.ct_Wimg);          // kernel row inside the patch
  const uint32_t plane_elems = (uint32_t)(p.ct_Himg * p.ct_Wimg);
  const uint32_t limit = plane_elems - row_off;                              // offs < limit <=> the row is inside the image
#pragma unroll
  for (int c3 = 0; c3 < CPG; ++c3) {
    const int c = g * CPG + c3;
    if (c >= p.ct_C) break;                                                   // warp-uniform
    float* base = p.out_f32 + (size_t)c * plane_elems + row_off;
    const float dn_s = (p.ct_std != nullptr) ? __ldg(p.ct_std + c) : 1.0f;
    const float dn_m = (p.ct_std != nullptr) ? __ldg(p.ct_mean + c) : 0.0f;
    const float* sc = stg + c3 * CT_PW * CT_LD;
#pragma unroll
    for (int k = 0; k < CT_PW; ++k)
      if (cl.offs[k] < limit)   // streaming store: nothing on the GPU re-reads the reconstruction, the weight should stay in L2
        __stcs(&base[cl.pix[k]], fmaf(sc[cl.sidx[k]], dn_s, dn_m));
  }
  __syncwarp();
}

// EPI_RESID, phase 1: issue the 32 residual loads of this chunk (clamped, unconditional addresses) -- called BEFORE the
// accumulator is read out of TMEM and staged, so the DRAM/L2 latency overlaps that work. resid may alias out.
__device__ __forceinline__ void epilogue_resid_load(const EpiParams& p, int row_base, int col0, int lane, int M, int N,
                                                    float (&rv)[32], int& my_t, int tile_tok) {
  const int rows_valid = min(32, M - row_base);
  const int col = col0 + lane;
  const int colc = (col < N) ? col : 0;
  (void)rows_valid;
  my_t = tile_tok;   // lane < rows_valid <=> row_base + lane < M, which is how tile_tok was guarded
#pragma unroll
  for (int rr = 0; rr < 32; ++rr) {
    const int t = __shfl_sync(0xffffffffu, my_t, rr);
    rv[rr] = p.resid[(size_t)max(t, 0) * p.ldo + colc];
  }
}
__device__ __forceinline__ void epilogue_resid_finish(const EpiParams& p, const float* stg, int col0, int lane, int N,
                                                      const float (&rv)[32], int my_t) {
  const int col = col0 + lane;
  const bool ok = col < N;
  const float b = (p.bias != nullptr && ok) ? __ldg(p.bias + col) : 0.f;
#pragma unroll
  for (int rr = 0; rr < 32; ++rr) {
    const int t = __shfl_sync(0xffffffffu, my_t, rr);
    if (t >= 0 && ok) {
      const float v = stg[rr * STG_LD + lane] + b + rv[rr];
      p.out_f32[(size_t)t * p.ldo + col] = v;
      if (p.out_bf16 != nullptr) p.out_bf16[(size_t)t * p.ld_bf16 + p.bf16_col0 + col] = __float2bfloat16(v);
    }
  }
}

// bf16 outputs, fast path: the thread that owns a row adds bias (+ GELU / query scale), packs its 32 columns to 64 bytes
// and parks them in shared memory as four 16-byte pieces (XOR-swizzled: conflict-free for both phases); the warp then
// writes 8 whole 64-byte row segments per instruction. ~40 instructions per 32 x 32 chunk instead of ~110 for the
// fp32-staged path -- and the K = 1024 GEMMs (qkv, fc1) are epilogue-bound, not MMA-bound.
template <int KIND>
__device__ __forceinline__ void epilogue_bf16_fast(const uint32_t (&acc)[32], const float* __restrict__ bias, int col0,
                                                   float scale, uint8_t* stage, __nv_bfloat16* dst_row0, size_t ld,
                                                   int rows_valid, int lane, bool gelu_fast = false) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(bias + col0 + i));   // same address in all lanes
    v[i] = __uint_as_float(acc[i]) + b.x;
    v[i + 1] = __uint_as_float(acc[i + 1]) + b.y;
    v[i + 2] = __uint_as_float(acc[i + 2]) + b.z;
    v[i + 3] = __uint_as_float(acc[i + 3]) + b.w;
  }
  if constexpr (KIND == EPI_GELU_BF16) {
    if (gelu_fast) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu_bf16out(v[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
    }
  }
  if constexpr (is_qkv<KIND>()) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= scale;
  }
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u;
    u.x = pack_pair<KIND>(v[8 * c + 0], v[8 * c + 1]);
    u.y = pack_pair<KIND>(v[8 * c + 2], v[8 * c + 3]);
    u.z = pack_pair<KIND>(v[8 * c + 4], v[8 * c + 5]);
    u.w = pack_pair<KIND>(v[8 * c + 6], v[8 * c + 7]);
    *reinterpret_cast<uint4*>(stage + lane * 64 + ((c ^ sw) << 4)) = u;
  }
  __syncwarp();
  const int c = lane & 3;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2);
    const uint4 u = *reinterpret_cast<const uint4*>(stage + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
    if (r < rows_valid) *reinterpret_cast<uint4*>(dst_row0 + (size_t)r * ld + c * 8) = u;
  }
  __syncwarp();
}

// V^T (EPI_QKV, chunk wholly inside one head of V): the same idea transposed -- the row owner drops its 32 values into
// a [dim][row] bf16 staging tile (2-byte stores, one contiguous 64-byte line per dim), then the warp writes 16-byte
// pieces = 8 consecutive rows of one dim of V^T[head][dim][rows_total].
template <int KIND>
__device__ __forceinline__ void epilogue_vt_fast(const uint32_t (&acc)[32], const float* __restrict__ bias, int col0,
                                                 uint8_t* stage, __nv_bfloat16* vt_d0_row0, size_t rows_total,
                                                 int rows_valid, int lane) {
  __nv_bfloat16* st16 = reinterpret_cast<__nv_bfloat16*>(stage);
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(bias + col0 + i));
    st16[(i + 0) * 32 + lane] = cvt_elem<KIND>(__uint_as_float(acc[i]) + b.x);
    st16[(i + 1) * 32 + lane] = cvt_elem<KIND>(__uint_as_float(acc[i + 1]) + b.y);
    st16[(i + 2) * 32 + lane] = cvt_elem<KIND>(__uint_as_float(acc[i + 2]) + b.z);
    st16[(i + 3) * 32 + lane] = cvt_elem<KIND>(__uint_as_float(acc[i + 3]) + b.w);
  }
  __syncwarp();
  const int c = lane & 3;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int d = it * 8 + (lane >> 2);
    const uint4 u = *reinterpret_cast<const uint4*>(stage + d * 64 + c * 16);
    if (c * 8 < rows_valid) *reinterpret_cast<uint4*>(vt_d0_row0 + (size_t)d * rows_total + c * 8) = u;
  }
  __syncwarp();
}

// fp32 outputs with an fp32 addend (EPI_RESID: residual stream, rows scattered through the window map; EPI_F32: bias +
// pos-embed): same scheme at 128 bytes per row. Phase 1 (before the accumulator is read, so the DRAM / L2 latency
// overlaps it): every lane fetches the 16-byte piece of the addend it will need in phase 2. Phase 2: the row owner
// parks its 32 floats as eight swizzled 16-byte pieces; the warp then adds and writes four whole 128-byte row
// segments per instruction. The addend may alias the output: each element is read and written by the same lane.
struct F32Fast {
  float4 add[8];
  float4 bias;
  int my_t;      // output row of the accumulator row this lane owns, -1: none
  bool on;
};
template <int KIND>
__device__ __forceinline__ void epilogue_f32_prefetch(const EpiParams& p, int row_base, int col0, int lane, int M, int N,
                                                      F32Fast& f, int tile_tok) {
  const float* src = (KIND == EPI_RESID) ? p.resid : p.add;
  const int ld_src = (KIND == EPI_RESID) ? p.ldo : p.lda;
  f.on = (col0 + 32 <= N) && ((p.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out_f32) & 15) == 0) &&
         (src == nullptr || (((ld_src & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0))) &&
         (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) &&
         (KIND != EPI_RESID || p.out_bf16 == nullptr ||
          ((((p.ld_bf16 | p.bf16_col0) & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out_bf16) & 7) == 0)));
  if (!f.on) return;
  const int row = row_base + lane;
  f.my_t = (KIND == EPI_RESID) ? tile_tok : ((row < M) ? row : -1);   // tile_tok is -1 for rows >= M and for pad rows
  const int piece = lane & 7;
  f.bias = (p.bias != nullptr) ? __ldg(reinterpret_cast<const float4*>(p.bias + col0 + piece * 4))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
  // EPI_F32: a pos-embed addend repeats down a batch of frames (row % add_period); taken once per lane, not per piece
  int src_t = max(f.my_t, 0);
  if constexpr (KIND == EPI_F32) {
    if (p.add_period > 0) src_t = src_t % p.add_period;
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int t = __shfl_sync(0xffffffffu, src_t, it * 4 + (lane >> 3));
    f.add[it] = (src != nullptr)
                    ? *reinterpret_cast<const float4*>(src + (size_t)t * ld_src + col0 + piece * 4)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
template <int KIND>
__device__ __forceinline__ void epilogue_f32_finish(const EpiParams& p, const uint32_t (&acc)[32], uint8_t* stage,
                                                    int col0, int lane, const F32Fast& f) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<uint4*>(stage + lane * 128 + ((q ^ (lane & 7)) << 4)) =
        make_uint4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
  __syncwarp();
  const int piece = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + (lane >> 3);
    const int t = __shfl_sync(0xffffffffu, f.my_t, r);
    const float4 u = *reinterpret_cast<const float4*>(stage + r * 128 + ((piece ^ (r & 7)) << 4));
    if (t >= 0) {
      const float4 v = make_float4(u.x + f.bias.x + f.add[it].x, u.y + f.bias.y + f.add[it].y,
                                   u.z + f.bias.z + f.add[it].z, u.w + f.bias.w + f.add[it].w);
      *reinterpret_cast<float4*>(p.out_f32 + (size_t)t * p.ldo + col0 + piece * 4) = v;
      if constexpr (KIND == EPI_RESID) {
        if (p.out_bf16 != nullptr)
          *reinterpret_cast<uint2*>(p.out_bf16 + (size_t)t * p.ld_bf16 + p.bf16_col0 + col0 + piece * 4) =
              make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      }
    }
  }
  __syncwarp();
}

// kinds whose natural store direction is along the ROWS (thread = row): written straight from registers
template <int KIND>
__device__ __forceinline__ constexpr bool epi_is_direct() { return KIND == EPI_T_F32 || KIND == EPI_PIXSHUF; }

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_OFFSET = GEMM_STAGES * STAGE_BYTES;          // 8 epilogue warps x [32][33] fp32 staging
  static constexpr int STG_BYTES = 8 * 32 * STG_LD * 4;
  static constexpr int BAR_OFFSET = STG_OFFSET + STG_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256;  // no alignment slack: the dynamic window starts 1024-aligned (checked)
};

template <int BN, int KIND>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmShape shp, const EpiParams epi) {
  using L = GemmSmem<BN>;
  constexpr uint32_t TMEM_COLS = 2 * BN;  // 256 or 512: power of two >= 32
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B operand tiles need 1024-byte alignment
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + GEMM_STAGES;
  uint64_t* tfull_bar = empty_bar + GEMM_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (shp.M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (shp.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (shp.K + GEMM_BK - 1) / GEMM_BK;
  const int split_terms = 1 + (shp.a_split ? 1 : 0) + (shp.b_split ? 1 : 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * GEMM_BM;   // n fastest: CTAs running together share the A tile
        const int n0 = (tile % n_tiles) * BN;
        int pe_j0[8], pe_h0[8], pe_nbox = 0, pe_r = 0, pe_kc = 0, pe_cs0 = 0;
        if (shp.a_mode == A_PATCH) {
          pe_nbox = GEMM_BM / shp.pe_box_rows;  // <= 8 (box rows >= 16 for the supported geometries, checked on the host)
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int t = m0 + g * shp.pe_box_rows;
            const int i = t / shp.pe_Wp;
            pe_j0[g] = t - i * shp.pe_Wp;
            const int fb = (shp.pe_Hg > 0) ? i / shp.pe_Hg : 0;     // frame of a batch
            pe_h0[g] = shp.pe_sh * (i - fb * shp.pe_Hg) + fb * shp.pe_img_h;
          }
        }
        if (split_terms > 1) {
          // split-bf16 operands (see GemmShape): every k-block is fed once per term, halves selected by the outermost
          // tensor-map coordinate
          for (int kb = 0; kb < k_blocks; ++kb) {
            for (int term = 0; term < split_terms; ++term, ++it) {
              const int a_half = (shp.a_split && term == 1) ? 1 : 0;
              const int b_half = (shp.b_split && term == split_terms - 1) ? 1 : 0;
              const int s = it % GEMM_STAGES;
              const uint32_t ph = (it / GEMM_STAGES) & 1;
              mbar_wait(&empty_bar[s], ph ^ 1);
              uint8_t* sa = smem + s * L::STAGE_BYTES;
              uint8_t* sb = sa + L::A_BYTES;
              mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
              const int k0 = kb * GEMM_BK;
              if (shp.a_mode == A_PLAIN) {
                if (shp.a_split) tma_load_3d(sa, &tmA, &full_bar[s], k0, m0, a_half);
                else tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);
              } else if (shp.a_mode == A_PATCH) {
#pragma unroll
                for (int g = 0; g < 8; ++g)
                  if (g < pe_nbox) {
                    if (shp.a_split)
                      tma_load_4d(sa + g * shp.pe_box_rows * 128, &tmA, &full_bar[s], pe_cs0, pe_j0[g], pe_h0[g] + pe_r, a_half);
                    else
                      tma_load_3d(sa + g * shp.pe_box_rows * 128, &tmA, &full_bar[s], pe_cs0, pe_j0[g], pe_h0[g] + pe_r);
                  }
              } else {  // A_CONCAT
                const int kk = (k0 < shp.cc_D) ? k0 : k0 - shp.cc_D;
                const int mm = (k0 < shp.cc_D) ? m0 : m0 - shp.cc_shift;
                if (shp.a_split) tma_load_3d(sa, &tmA, &full_bar[s], kk, mm, a_half);
                else tma_load_2d(sa, &tmA, &full_bar[s], kk, mm);
              }
              if (shp.b_split) tma_load_3d(sb, &tmB, &full_bar[s], k0, n0, b_half);
              else tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);
            }
            if (shp.a_mode == A_PATCH) {
              if (++pe_kc == shp.pe_kpr) { pe_kc = 0; pe_cs0 = 0; ++pe_r; } else { pe_cs0 += GEMM_BK; }
            }
          }
          continue;
        }
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % GEMM_STAGES;
          const uint32_t ph = (it / GEMM_STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
          if (shp.a_mode == A_PLAIN) {
            tma_load_2d(sa, &tmA, &full_bar[s], kb * GEMM_BK, m0);
          } else if (shp.a_mode == A_PATCH) {
            // (the kernel row r and column block advance incrementally; box coordinates were hoisted per tile)
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if (g < pe_nbox)
                tma_load_3d(sa + g * shp.pe_box_rows * 128, &tmA, &full_bar[s], pe_cs0, pe_j0[g], pe_h0[g] + pe_r);
            if (++pe_kc == shp.pe_kpr) { pe_kc = 0; pe_cs0 = 0; ++pe_r; } else { pe_cs0 += GEMM_BK; }
          } else {  // A_CONCAT
            const int k0 = kb * GEMM_BK;
            if (k0 < shp.cc_D)
              tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);
            else
              tma_load_2d(sa, &tmA, &full_bar[s], k0 - shp.cc_D, m0 - shp.cc_shift);
          }
          tma_load_2d(sb, &tmB, &full_bar[s], kb * GEMM_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
      uint32_t it = 0, tl = 0;
      const int k_iters = k_blocks * split_terms;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
        const uint32_t as = tl & 1;
        const uint32_t aph = (tl >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < k_iters; ++kb, ++it) {
          const int s = it % GEMM_STAGES;
          const uint32_t ph = (it / GEMM_STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
          const uint64_t adesc = umma_smem_desc_sw128(sa);
          const uint64_t bdesc = umma_smem_desc_sw128(sa + L::A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 elements (32 bytes) along K inside the 128-byte swizzle atom: +2 in 16-byte units
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete
      }
    }
  } else if (warp >= GEMM_EPI_WARP0) {
    // ===================== epilogue =====================
    const int ew = warp - GEMM_EPI_WARP0;      // 0..7
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int half = ew >> 2;                  // which half of the columns
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
      const int m0 = (tile / n_tiles) * GEMM_BM;
      const int n0 = (tile % n_tiles) * BN;
      const uint32_t as = tl & 1;
      const uint32_t aph = (tl >> 1) & 1;
      int tile_tok = -1;   // the window map of this lane's accumulator row, computed once per tile
      if constexpr (KIND == EPI_RESID) {
        // pull this warp's share of the residual tile (32 rows x BN/2 fp32) towards L2 while the main loop of the tile
        // is still running: the epilogue is otherwise bound by four serial DRAM round trips per tile
        const int prow = m0 + quarter * 32 + lane;
        const int pt = (prow < shp.M) ? epi.wm.to_token(prow) : -1;
        tile_tok = pt;
        if (pt >= 0) {
          const float* pr = epi.resid + (size_t)pt * epi.ldo + n0 + half * (BN / 2);
#pragma unroll
          for (int q = 0; q < BN / 2; q += 32)
            if (n0 + half * (BN / 2) + q < shp.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + q));
        }
      }
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const int row_base = m0 + quarter * 32;
      const int row = row_base + lane;
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + as * BN + half * (BN / 2);
      float* stg = reinterpret_cast<float*>(smem + L::STG_OFFSET) + ew * (32 * STG_LD);
      ConvtLane<KIND> ctl;
      if constexpr (KIND == EPI_CONVT) {
        if (epi.ct_cpg > 0) convt_lane_setup(epi, row, shp.M, lane, ctl);
      }
#pragma unroll 1
      for (int c = 0; c < BN / 2; c += 32) {
        const int col0 = n0 + half * (BN / 2) + c;
        if (col0 >= shp.N) break;  // warp-uniform
        float rv[32];
        int my_t = -1;
        F32Fast ff;
        ff.on = false;
        if constexpr (KIND == EPI_RESID || KIND == EPI_F32)
          epilogue_f32_prefetch<KIND>(epi, row_base, col0, lane, shp.M, shp.N, ff, tile_tok);
        if constexpr (KIND == EPI_RESID) {
          if (!ff.on) epilogue_resid_load(epi, row_base, col0, lane, shp.M, shp.N, rv, my_t, tile_tok);
        }
        uint32_t acc[32];
        tmem_ld_32x32(taddr + c, acc);
        tmem_ld_wait();
        if constexpr (KIND == EPI_RESID || KIND == EPI_F32) {
          if (ff.on) {   // warp-uniform
            epilogue_f32_finish<KIND>(epi, acc, reinterpret_cast<uint8_t*>(stg), col0, lane, ff);
            continue;
          }
        }
        bool direct = epi_is_direct<KIND>();
        if constexpr (is_qkv<KIND>())  // a chunk that lies wholly inside V is written transposed, thread = row
          direct = (col0 >= 2 * epi.D) && (col0 + 32 <= shp.N);
        bool fast = false;
        if constexpr (KIND == EPI_BF16 || KIND == EPI_GELU_BF16) {
          fast = (col0 + 32 <= shp.N) && ((epi.ldo & 7) == 0) && ((reinterpret_cast<uintptr_t>(epi.out_bf16) & 15) == 0) &&
                 (epi.bias == nullptr || (reinterpret_cast<uintptr_t>(epi.bias) & 15) == 0);
          if (fast)
            epilogue_bf16_fast<KIND>(acc, epi.bias, col0, 1.0f, reinterpret_cast<uint8_t*>(stg),
                                     epi.out_bf16 + (size_t)row_base * epi.ldo + col0, (size_t)epi.ldo,
                                     min(32, shp.M - row_base), lane, epi.gelu_fast != 0);
        }
        if constexpr (is_qkv<KIND>()) {   // a chunk wholly inside one head of Q or K
          fast = !direct && (col0 + 32 <= shp.N) && (((epi.D | epi.hd) & 31) == 0) && (col0 < 2 * epi.D) &&
                 (((reinterpret_cast<uintptr_t>(epi.q) | reinterpret_cast<uintptr_t>(epi.k)) & 15) == 0) &&
                 (epi.bias == nullptr || (reinterpret_cast<uintptr_t>(epi.bias) & 15) == 0);
          if (fast) {
            const QkvSplit qs = qkv_split(col0, epi.D, epi.hd);
            const int which = qs.which, head = qs.head, d0 = qs.d;
            __nv_bfloat16* base = (which == 0 ? epi.q : epi.k) + ((size_t)head * epi.rows_total + row_base) * epi.hd + d0;
            epilogue_bf16_fast<KIND>(acc, epi.bias, col0, which == 0 ? epi.qscale : 1.0f, reinterpret_cast<uint8_t*>(stg),
                                     base, (size_t)epi.hd, min(32, shp.M - row_base), lane);
          }
        }
        if constexpr (is_qkv<KIND>()) {   // a chunk wholly inside one head of V: transposed store
          if (!fast && direct && (((epi.D | epi.hd) & 31) == 0) && ((epi.rows_total & 7) == 0) && ((shp.M & 7) == 0) &&
              ((reinterpret_cast<uintptr_t>(epi.vt) & 15) == 0) &&
              (epi.bias == nullptr || (reinterpret_cast<uintptr_t>(epi.bias) & 15) == 0)) {
            const QkvSplit qs = qkv_split(col0, epi.D, epi.hd);
            const int head = qs.head, d0 = qs.d;
            epilogue_vt_fast<KIND>(acc, epi.bias, col0, reinterpret_cast<uint8_t*>(stg),
                             epi.vt + ((size_t)head * epi.hd + d0) * epi.rows_total + row_base, (size_t)epi.rows_total,
                             min(32, shp.M - row_base), lane);
            fast = true;
          }
        }
        if constexpr (KIND == EPI_CONVT) {
          if (epi.ct_cpg > 0) {   // warp-uniform
            epilogue_convt_grouped(epi, acc, stg, ctl, col0, lane);
            continue;
          }
        }
        if (fast) {
        } else if (direct) {
          epilogue_store<KIND>(epi, row, col0, acc, shp.M, shp.N);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) stg[lane * STG_LD + i] = __uint_as_float(acc[i]);
          __syncwarp();
          if constexpr (KIND == EPI_RESID)
            epilogue_resid_finish(epi, stg, col0, lane, shp.N, rv, my_t);
          else
            epilogue_rows<KIND>(epi, stg, row_base, col0, lane, shp.M, shp.N);
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}


// =================================================================================================================
// CTA-pair variant: a cluster of two CTAs (two SMs of one TPC) computes one 256 x 256 output tile with
// tcgen05.mma.cta_group::2. Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 N rows), so
// the L2 -> SM traffic per output element drops by a third against the single-CTA 128 x 256 tile -- and the
// single-CTA kernel is L2-bandwidth bound (48 KB per 512 MMA cycles per SM, chip-wide above the LTS throughput cap).
// The leader CTA (cluster rank 0) issues the MMAs; both CTAs run a TMA producer and eight epilogue warps over their own
// half of the accumulator (TMEM lanes = their 128 rows).
struct GemmSmem2 {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;        // 16 KB: this CTA's 128 rows
  static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;       // 16 KB: this CTA's half of the N rows
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 6;
  static constexpr int STG_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int STG_BYTES = 8 * 32 * STG_LD * 4;
  static constexpr int BAR_OFFSET = STG_OFFSET + STG_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256;
};

template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmShape shp,
                const EpiParams epi) {
  using L = GemmSmem2;
  constexpr int BN = L::BN;
  constexpr int ST = L::STAGES;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + ST;
  uint64_t* tfull_bar = empty_bar + ST;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int m_tiles = (shp.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const int n_tiles = (shp.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (shp.K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < ST; ++s) {
      mbar_init(&full_bar[s], 1);    // leader's producer arrives (with the byte count of BOTH CTAs)
      mbar_init(&empty_bar[s], 1);   // multicast MMA commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);   // multicast MMA commit
      mbar_init(&tempty_bar[s], 16); // 8 epilogue warps of each CTA arrive on the LEADER's barrier
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer (both CTAs) =====================
      uint32_t it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        const int m0 = (tile / n_tiles) * (2 * GEMM_BM) + (int)rank * GEMM_BM;
        const int n0 = (tile % n_tiles) * BN + (int)rank * (BN / 2);
        int pe_j0[8], pe_h0[8], pe_nbox = 0, pe_r = 0, pe_kc = 0, pe_cs0 = 0;
        if (shp.a_mode == A_PATCH) {
          pe_nbox = GEMM_BM / shp.pe_box_rows;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int t = m0 + g * shp.pe_box_rows;
            const int i = t / shp.pe_Wp;
            pe_j0[g] = t - i * shp.pe_Wp;
            const int fb = (shp.pe_Hg > 0) ? i / shp.pe_Hg : 0;     // frame of a batch
            pe_h0[g] = shp.pe_sh * (i - fb * shp.pe_Hg) + fb * shp.pe_img_h;
          }
        }
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % ST;
          const uint32_t ph = (it / ST) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          const uint32_t fb = map_to_cta(smem_u32(&full_bar[s]), 0);  // the leader's barrier
          if (leader) mbar_expect_tx(&full_bar[s], 2 * L::STAGE_BYTES);
          if (shp.a_mode == A_PLAIN) {
            tma_load_2d_pair(sa, &tmA, fb, kb * GEMM_BK, m0);
          } else if (shp.a_mode == A_PATCH) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if (g < pe_nbox)
                tma_load_3d_pair(sa + g * shp.pe_box_rows * 128, &tmA, fb, pe_cs0, pe_j0[g], pe_h0[g] + pe_r);
            if (++pe_kc == shp.pe_kpr) { pe_kc = 0; pe_cs0 = 0; ++pe_r; } else { pe_cs0 += GEMM_BK; }
          } else {  // A_CONCAT
            const int k0 = kb * GEMM_BK;
            if (k0 < shp.cc_D)
              tma_load_2d_pair(sa, &tmA, fb, k0, m0);
            else
              tma_load_2d_pair(sa, &tmA, fb, k0 - shp.cc_D, m0 - shp.cc_shift);
          }
          tma_load_2d_pair(sb, &tmB, fb, kb * GEMM_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===================== MMA issuer (leader CTA only) =====================
      constexpr uint32_t idesc = umma_idesc_bf16(2 * GEMM_BM, BN);
      uint32_t it = 0, tl = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++tl) {
        const uint32_t as = tl & 1;
        const uint32_t aph = (tl >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % ST;
          const uint32_t ph = (it / ST) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
          const uint64_t adesc = umma_smem_desc_sw128(sa);
          const uint64_t bdesc = umma_smem_desc_sw128(sa + L::A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_pair(&empty_bar[s]);   // frees this stage in BOTH CTAs
        }
        umma_commit_pair(&tfull_bar[as]);    // accumulator complete, both CTAs
      }
    }
  } else if (warp >= GEMM_EPI_WARP0) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int ew = warp - GEMM_EPI_WARP0;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    uint32_t tl = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++tl) {
      const int m0 = (tile / n_tiles) * (2 * GEMM_BM) + (int)rank * GEMM_BM;
      const int n0 = (tile % n_tiles) * BN;
      const uint32_t as = tl & 1;
      const uint32_t aph = (tl >> 1) & 1;
      int tile_tok = -1;
      if constexpr (KIND == EPI_RESID) {
        const int prow = m0 + quarter * 32 + lane;
        const int pt = (prow < shp.M) ? epi.wm.to_token(prow) : -1;
        tile_tok = pt;
        if (pt >= 0) {
          const float* pr = epi.resid + (size_t)pt * epi.ldo + n0 + half * (BN / 2);
#pragma unroll
          for (int q = 0; q < BN / 2; q += 32)
            if (n0 + half * (BN / 2) + q < shp.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + q));
        }
      }
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const int row_base = m0 + quarter * 32;
      const int row = row_base + lane;
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + as * BN + half * (BN / 2);
      float* stg = reinterpret_cast<float*>(smem + L::STG_OFFSET) + ew * (32 * STG_LD);
      ConvtLane<KIND> ctl;
      if constexpr (KIND == EPI_CONVT) {
        if (epi.ct_cpg > 0) convt_lane_setup(epi, row, shp.M, lane, ctl);
      }
#pragma unroll 1
      for (int c = 0; c < BN / 2; c += 32) {
        const int col0 = n0 + half * (BN / 2) + c;
        if (col0 >= shp.N) break;  // warp-uniform
        float rv[32];
        int my_t = -1;
        F32Fast ff;
        ff.on = false;
        if constexpr (KIND == EPI_RESID || KIND == EPI_F32)
          epilogue_f32_prefetch<KIND>(epi, row_base, col0, lane, shp.M, shp.N, ff, tile_tok);
        if constexpr (KIND == EPI_RESID) {
          if (!ff.on) epilogue_resid_load(epi, row_base, col0, lane, shp.M, shp.N, rv, my_t, tile_tok);
        }
        uint32_t acc[32];
        tmem_ld_32x32(taddr + c, acc);
        tmem_ld_wait();
        if constexpr (KIND == EPI_RESID || KIND == EPI_F32) {
          if (ff.on) {   // warp-uniform
            epilogue_f32_finish<KIND>(epi, acc, reinterpret_cast<uint8_t*>(stg), col0, lane, ff);
            continue;
          }
        }
        bool direct = epi_is_direct<KIND>();
        if constexpr (is_qkv<KIND>()) direct = (col0 >= 2 * epi.D) && (col0 + 32 <= shp.N);
        bool fast = false;
        if constexpr (KIND == EPI_BF16 || KIND == EPI_GELU_BF16) {
          fast = (col0 + 32 <= shp.N) && ((epi.ldo & 7) == 0) && ((reinterpret_cast<uintptr_t>(epi.out_bf16) & 15) == 0) &&
                 (epi.bias == nullptr || (reinterpret_cast<uintptr_t>(epi.bias) & 15) == 0);
          if (fast)
            epilogue_bf16_fast<KIND>(acc, epi.bias, col0, 1.0f, reinterpret_cast<uint8_t*>(stg),
                                     epi.out_bf16 + (size_t)row_base * epi.ldo + col0, (size_t)epi.ldo,
                                     min(32, shp.M - row_base), lane, epi.gelu_fast != 0);
        }
        if constexpr (is_qkv<KIND>()) {   // a chunk wholly inside one head of Q or K
          fast = !direct && (col0 + 32 <= shp.N) && (((epi.D | epi.hd) & 31) == 0) && (col0 < 2 * epi.D) &&
                 (((reinterpret_cast<uintptr_t>(epi.q) | reinterpret_cast<uintptr_t>(epi.k)) & 15) == 0) &&
                 (epi.bias == nullptr || (reinterpret_cast<uintptr_t>(epi.bias) & 15) == 0);
          if (fast) {
            const QkvSplit qs = qkv_split(col0, epi.D, epi.hd);
            const int which = qs.which, head = qs.head, d0 = qs.d;
            __nv_bfloat16* base = (which == 0 ? epi.q : epi.k) + ((size_t)head * epi.rows_total + row_base) * epi.hd + d0;
            epilogue_bf16_fast<KIND>(acc, epi.bias, col0, which == 0 ? epi.qscale : 1.0f, reinterpret_cast<uint8_t*>(stg),
                                     base, (size_t)epi.hd, min(32, shp.M - row_base), lane);
          }
        }
        if constexpr (is_qkv<KIND>()) {   // a chunk wholly inside one head of V: transposed store
          if (!fast && direct && (((epi.D | epi.hd) & 31) == 0) && ((epi.rows_total & 7) == 0) && ((shp.M & 7) == 0) &&
              ((reinterpret_cast<uintptr_t>(epi.vt) & 15) == 0) &&
              (epi.bias == nullptr || (reinterpret_cast<uintptr_t>(epi.bias) & 15) == 0)) {
            const QkvSplit qs = qkv_split(col0, epi.D, epi.hd);
            const int head = qs.head, d0 = qs.d;
            epilogue_vt_fast<KIND>(acc, epi.bias, col0, reinterpret_cast<uint8_t*>(stg),
                             epi.vt + ((size_t)head * epi.hd + d0) * epi.rows_total + row_base, (size_t)epi.rows_total,
                             min(32, shp.M - row_base), lane);
            fast = true;
          }
        }
        if constexpr (KIND == EPI_CONVT) {
          if (epi.ct_cpg > 0) {   // warp-uniform
            epilogue_convt_grouped(epi, acc, stg, ctl, col0, lane);
            continue;
          }
        }
        if (fast) {
        } else if (direct) {
          epilogue_store<KIND>(epi, row, col0, acc, shp.M, shp.N);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) stg[lane * STG_LD + i] = __uint_as_float(acc[i]);
          __syncwarp();
          if constexpr (KIND == EPI_RESID)
            epilogue_resid_finish(epi, stg, col0, lane, shp.N, rv, my_t);
          else
            epilogue_rows<KIND>(epi, stg, row_base, col0, lane, shp.M, shp.N);
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tempty_bar[as]), 0));
    }
  }
  tc_fence_before();
  cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer may still touch this CTA's smem / barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<TMEM_COLS>(tmem_base);
  }
}

}  // namespace cra5
