// attn_tc4.cu -- persistent, ping-pong fused softmax(Q K^T) V on tcgen05 tensor cores for head_dim 64.
//
// Contract (reference: WindowAttention.forward vit_nlc.py:219-258 incl. the un-masked zero-pad
// tokens, and Attention.forward vit_nlc.py:94-112): token list in "attention order", a segment = seg_len contiguous
// rows, Q pre-multiplied by head_dim^-0.5. Organised around what bounds head_dim 64 on Blackwell: the softmax, not
// the MMAs (512 tensor cycles per 128x128 tile against 16 384 exponentials on 16 MUFU lanes per SM and clock).
//
//   * one CTA per SM, persistent over work items; an item = TWO 128-row query tiles (A, B) of one (segment, head) that
//     share every K/V tile streamed through a 4-stage TMA ring; the next item's Q and K/V are prefetched while the
//     current one drains, TMEM is allocated once;
//   * warp 0 TMA producer, warps 1 / 3 MMA issuers of tile A / B, warp 2 TMEM allocator, warps 4-7 softmax of tile A,
//     warps 8-11 softmax of tile B (thread = query row, all 128 scores of the row in registers; setmaxnreg moves registers from warps
//     0-3 to the softmax warps);
//   * S = Q K^T lands in TMEM (one buffer per tile; it is free again as soon as the row is in registers, so S of step
//     j+1 is computed during the softmax of step j); P goes back to TMEM as bf16 and feeds the PV product straight
//     from TMEM (tcgen05.mma with the A operand in tensor memory) -- no shared-memory round trip, no proxy fence;
//   * O accumulates in TMEM over the whole KV loop. The running maximum is only moved when a row's tile maximum
//     exceeds it by more than 2^8 (in the exp2 domain); then, and only then, O and the row sum are rescaled -- the
//     result is exact because numerator and denominator always share the same reference;
//   * exponentials: packed fp32x2 arithmetic (fma.rn.f32x2 / add.f32x2) for the scale-subtract and the row sums; a
//     fixed fraction of the column pairs evaluates 2^x with a Cody-Waite split + cubic on the FMA pipe instead of
//     MUFU.EX2, which balances the two pipes.
#include "ptx.cuh"
#include "host_util.h"
#include "kernels.h"
#include <cstdlib>

// Development aid (-DCRA5_ATTN_TRACE, tools/attn_trace.py; never in the shipped library): CTA 0's lane-quarter-0 softmax
// warps of tile A and B stamp clock64 at the phase boundaries of their first 64 KV steps.
#ifdef CRA5_ATTN_TRACE
__device__ long long g_attn_trace[2 * 64 * 8];
#define A4_TRACE(tile, step, ph)                                                                      \
  do {                                                                                                \
    if (blockIdx.x == 0 && (warp & 3) == 0 && lane == 0 && (step) < 64u) {                            \
      long long t_;                                                                                   \
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_)::"memory");                                    \
      g_attn_trace[((tile) * 64 + (step)) * 8 + (ph)] = t_;                                           \
    }                                                                                                 \
  } while (0)
extern "C" __attribute__((visibility("default"))) int cra5_debug_attn_trace(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_attn_trace, sizeof(g_attn_trace));
}
#else
#define A4_TRACE(tile, step, ph) do {} while (0)
#endif

namespace cra5 {

namespace {

constexpr int A4_BM = 128;          // rows per query tile
constexpr int A4_BN = 128;          // keys per KV tile
constexpr int A4_HD = 64;
constexpr int A4_STAGES = 4;
constexpr int A4_THREADS = 384;

struct A4Smem {
  static constexpr int Q_TILE = A4_BM * A4_HD * 2;         // 16 KB
  static constexpr int Q_BUF = 2 * Q_TILE;                 // tiles A and B
  static constexpr int K_BYTES = A4_BN * A4_HD * 2;        // 16 KB
  static constexpr int V_BYTES = A4_HD * A4_BN * 2;        // 16 KB (two 64-key halves, each [64 dims][64 keys])
  static constexpr int KV_BYTES = K_BYTES + V_BYTES;
  static constexpr int OFF_Q = 0;                          // two Q buffers (items alternate)
  static constexpr int OFF_KV = OFF_Q + 2 * Q_BUF;
  static constexpr int OFF_BAR = OFF_KV + A4_STAGES * KV_BYTES;
  static constexpr int TOTAL = OFF_BAR + 256;
};

// TMEM columns (512 allocated)
constexpr uint32_t TM_S = 0;      // S_A at 0, S_B at 128 (fp32, 128 columns each)
constexpr uint32_t TM_P = 256;    // P_A at 256, P_B at 320 (bf16 pairs, 64 columns each)
constexpr uint32_t TM_O = 384;    // O_A at 384, O_B at 448 (fp32, 64 columns each)

struct A4Params {
  int seg_len;        // tokens per segment
  int rows_total;     // rows of Q/K per head (= row stride of Vt)
  int n_seg;
  int heads;
  int n_qp;           // query-tile pairs per segment
  int q_part_from;    // segments whose index within their frame (seg % seg_period) is >= this only need their first
  int q_part_rows;    //   q_part_rows query rows (padded windows)
  int seg_period;     // segments per frame of a batch
  __nv_bfloat16* out; // [rows_total, ldo]
  __nv_bfloat16* out_lo;  // optional: bf16(O - bf16(O)), so that the projection can run in split-bf16 form
  int ldo;
};

__device__ __forceinline__ float ex2_mufu(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 2^x for a pair, x <= ~8, on the FMA/ALU pipes: n = round(x), f = x - n in [-0.5, 0.5], cubic minimax for 2^f
// (relative error ~1e-4, an order of magnitude below the bf16 rounding of P), exponent added as an integer.
__device__ __forceinline__ void ex2_poly2(uint64_t x2, float& r0, float& r1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  x0 = fmaxf(x0, -125.0f);
  x1 = fmaxf(x1, -125.0f);
  const uint64_t xc = pack2(x0, x1);
  const uint64_t magic = pack2(12582912.0f, 12582912.0f);      // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const uint64_t nmagic = pack2(-12582912.0f, -12582912.0f);
  const uint64_t t = add2(xc, magic);
  const uint64_t n = add2(t, nmagic);
  const uint64_t f = fma2(n, pack2(-1.0f, -1.0f), xc);
  uint64_t p = fma2(pack2(0.0555041f, 0.0555041f), f, pack2(0.2402265f, 0.2402265f));
  p = fma2(p, f, pack2(0.6931472f, 0.6931472f));
  p = fma2(p, f, pack2(1.0f, 1.0f));
  float p0, p1, t0, t1;
  unpack2(p, p0, p1);
  unpack2(t, t0, t1);
  r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// D[tmem] (+)= A[tmem] * B[smem desc]: A is a 128-lane x K bf16 operand in tensor memory (two K elements per column)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// exponentials of one 128-score row against the reference maximum folded into nm = -m_ref * log2(e): P as packed
// 16-bit pairs (bf16, or fp16 for the split-precision levels) and the fp32 row sum
template <int POLY, bool F16>
__device__ __forceinline__ float softmax_row(const uint32_t (&s)[A4_BN], float nm, uint32_t (&pk)[A4_BN / 2]) {
  constexpr float LOG2E = 1.4426950408889634f;
  const uint64_t l2e2 = pack2(LOG2E, LOG2E);
  const uint64_t nm2 = pack2(nm, nm);
  uint64_t sum_a = pack2(0.f, 0.f), sum_b = pack2(0.f, 0.f);
#pragma unroll
  for (int q = 0; q < A4_BN / 2; ++q) {
    const int i = 2 * q;
    const uint64_t xs = fma2(pack2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), l2e2, nm2);
    float e0, e1;
    if ((q & 7) < POLY) {
      ex2_poly2(xs, e0, e1);
    } else {
      float x0, x1;
      unpack2(xs, x0, x1);
      e0 = ex2_mufu(x0);
      e1 = ex2_mufu(x1);
    }
    if (q & 1) sum_b = add2(sum_b, pack2(e0, e1)); else sum_a = add2(sum_a, pack2(e0, e1));
    if constexpr (F16) {
      const __half2 hp = __floats2half2_rn(e0, e1);
      pk[q] = *reinterpret_cast<const uint32_t*>(&hp);
    } else {
      pk[q] = pack_bf16x2(e0, e1);
    }
  }
  float a0, a1, b0, b1;
  unpack2(sum_a, a0, a1);
  unpack2(sum_b, b0, b1);
  return (a0 + a1) + (b0 + b1);
}

struct Item {
  int head, seg, q0;
  bool act_a, act_b;
};
__device__ __forceinline__ Item decode_item(const A4Params& p, int item) {
  Item it;
  const int qp = item % p.n_qp;
  const int rest = item / p.n_qp;
  it.seg = rest % p.n_seg;
  it.head = rest / p.n_seg;
  it.q0 = qp * (2 * A4_BM);
  const int need = ((it.seg % p.seg_period) >= p.q_part_from) ? p.q_part_rows : p.seg_len;  // rows whose output is consumed
  it.act_a = it.q0 < need;
  it.act_b = it.q0 + A4_BM < need;
  return it;
}

// POLY: of every 8 column pairs, this many take the polynomial exp2. F16: Q, K, V arrive as fp16 (EPI_QKV_F16) and P is
// written as fp16 -- 11 instead of 8 significand bits on every attention operand, the format the reference's own GPU
// path uses (flash-attn on .half() tensors, vit_nlc.py:105-110); same tcgen05 kind::f16 pipeline, same speed.
// Where a KV step's cycles go (tools/attn_trace.py, clock64 stamps of one softmax warp, global shape, B200): 2740 per
// step and tile = wait S 140 + TMEM load and maximum 540 + exponentials 1700 + wait PV 105 + store P 97, tiles A and B
// of a scheduler nearly in phase. tools/micro/softmax_ops.cu gives the pipe rates behind it (cycles per warp
// instruction and scheduler): MUFU.EX2 8 (ex2.f16x2 / bf16x2 are two MUFUs: no gain), FFMA2 2.24, FADD2 / FMNMX3 / F2FP 2,
// FFMA / IADD 1. Per tile and warp the exponentials occupy the MUFU for 640 cycles and the FMA pipe for ~680 (the
// cubic costs the FMA pipe 8.5 cycles per exponential it takes off the MUFU's 8: POLY = 3 is the balance point), so
// two warps cannot finish them in less than ~1360; measured 1700.
// (Tried and dropped in round 2, tools/perf_attn.py, global shape at 8 frames per launch, all bit-identical:
//   * two threads per query row -- 16 softmax warps with 64 scores each, tile maxima swapped through shared memory
//     behind 64-thread named barriers, half tiles for a segment's 64-key tail: 715 TFLOP/s against 901 (the partner
//     warps run in lock step, so the phases still collide, and 112 registers per thread leave ptxas no room);
//   * speculative maximum -- exponentials against the current reference while the tile maximum is reduced alongside,
//     redo from TMEM when the reference has to move: 767, because S can then only be released to the MMA issuer after
//     the exponentials and the next S arrives 700 cycles late;
//   * ping-pong -- the two tiles' exponential phases exclude each other through named barriers, FA3-style: 775 alone,
//     845 with the speculative maximum; a single warp needs 1340 cycles for its exponentials, so exclusion loses more
//     than the overlap of the other tile's load / store phases wins;
//   * S fetched one step ahead (TMEM load issued after the exponentials, landing during the P store) with P stored in
//     two halves to make room in the register file: 810 -- the barrier wait in the middle of the exponentials splits
//     ptxas' scheduling region;
//   * more softmax warps per scheduler with 64-key steps (thread = row, 64 scores, no exchange between warps): four
//     tiles with P stored over S (TMEM 4 x 64 + 4 x 64 = 512 columns): 630 -- S(j+1) then depends on PV(j) through TMEM
//     and the tensor pipe drains between them (8 MMAs took 1200 cycles to issue); three tiles with the protocol of this
//     kernel (TMEM 3 x 64 S + 3 x 32 P + 3 x 64 O, control warps at the highest warp ids): 860-890 on the global shape,
//     560-575 against 520 on the 576-token windows, 800-855 against 765-790 at one frame per launch -- and a step of 64
//     keys still takes 1900 cycles: left alone the three tiles fall in step, and when a ring of named barriers keeps
//     them a third of a step apart each warp's exponentials take as long as before (~1000 cycles for 64 columns). One
//     in-order warp issues this instruction mix at an IPC of ~0.4; the FMA pipe is 60 % busy, the MUFU 50 %, with
//     two or with three warps per scheduler. Not kept: no gain on the shape that dominates the step;
//   * two threads per row once more, with what the variants above had taught (control warps at the highest ids, lean
//     issuer loops) and the tiles' exponential phases alternating through mbarriers: 815 (860 with POLY = 2) against
//     905. The trace shows the alternation working and a tile's two warps taking 1250 cycles for 2 x 64 columns -- what
//     ONE warp takes for 128. The exponentials are bound by the scheduler's FMA pipe (~700 cycles per tile and warp:
//     profiles/r2_ncu_attn_stalls.txt) plus MUFU (640), not by how many warps share them.)
// (Also tried: four instead of two running maxima per row -- no change, the
// FMNMX3 chain already hides behind the second TMEM load; computing the exponentials speculatively against the previous
// reference maximum while the tile maximum is still being reduced -- the scores then have to stay live for a possible
// redo and the kernel spills. POLY = 3 remains the best split: 907 / 881 / 726 TFLOP/s for POLY 3 / 2 / 4 on the global
// shape at 8 frames per launch.)
template <int POLY, bool F16>
__global__ void __launch_bounds__(A4_THREADS, 1)
attn_tc4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmVt, const A4Params p) {
  using L = A4Smem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);   // [2]
  uint64_t* q_empty = q_full + 2;                                       // [2]
  uint64_t* kv_full = q_empty + 2;                                      // [STAGES]
  uint64_t* kv_empty = kv_full + A4_STAGES;                             // [STAGES]
  uint64_t* s_full = kv_empty + A4_STAGES;                              // [2] per query tile
  uint64_t* s_empty = s_full + 2;
  uint64_t* p_full = s_empty + 2;
  uint64_t* pv_done = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_items = p.heads * p.n_seg * p.n_qp;
  const int n_kv = (p.seg_len + A4_BN - 1) / A4_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmVt);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 2);       // one arrival per query tile's MMA thread (an absent tile B still arrives)
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 128);
      mbar_init(&p_full[s], 128);
      mbar_init(&pv_done[s], 1);
    }
    for (int s = 0; s < A4_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 2);      // one arrival per query tile's MMA thread
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    reg_dec<64>();    // the pool setmaxnreg draws from is what the launch allocated: 384 x 168 >= 128 x 64 + 256 x 216
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer =====================
      uint32_t it_q = 0, it_kv = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        if (!it.act_a) continue;
        const int seg_row0 = it.seg * p.seg_len;
        const int qb = it_q & 1;
        mbar_wait(&q_empty[qb], ((it_q >> 1) & 1) ^ 1);
        uint8_t* sq = smem + L::OFF_Q + qb * L::Q_BUF;
        mbar_expect_tx(&q_full[qb], it.act_b ? L::Q_BUF : L::Q_TILE);
        tma_load_2d(sq, &tmQ, &q_full[qb], 0, it.head * p.rows_total + seg_row0 + it.q0);
        if (it.act_b) tma_load_2d(sq + L::Q_TILE, &tmQ, &q_full[qb], 0, it.head * p.rows_total + seg_row0 + it.q0 + A4_BM);
        ++it_q;
        for (int j = 0; j < n_kv; ++j, ++it_kv) {
          const int s = it_kv % A4_STAGES;
          mbar_wait(&kv_empty[s], ((it_kv / A4_STAGES) & 1) ^ 1);
          uint8_t* sk = smem + L::OFF_KV + s * L::KV_BYTES;
          uint8_t* sv = sk + L::K_BYTES;
          const int kv0 = seg_row0 + j * A4_BN;
          mbar_expect_tx(&kv_full[s], L::KV_BYTES);
          tma_load_2d(sk, &tmK, &kv_full[s], 0, it.head * p.rows_total + kv0);
          tma_load_2d(sv, &tmVt, &kv_full[s], kv0, it.head * A4_HD);
          tma_load_2d(sv + L::V_BYTES / 2, &tmVt, &kv_full[s], kv0 + 64, it.head * A4_HD);
        }
      }
    } else if ((warp == 1 || warp == 3) && lane == 0) {
      // ===================== MMA issuers: warp 1 drives query tile A, warp 3 tile B =====================
      // (two independent in-order instruction streams: neither tile's PV product ever queues behind a wait that
      // belongs to the other tile)
      const int x = warp >> 1;
      constexpr uint32_t idesc_s = umma_idesc_f16kind(A4_BM, A4_BN, F16 ? 0u : 1u);
      constexpr uint32_t idesc_pv = umma_idesc_f16kind(A4_BM, A4_HD, F16 ? 0u : 1u);
      const uint64_t q_desc0 = umma_smem_desc_sw128(smem_u32(smem + L::OFF_Q));
      const uint64_t kv_desc0 = umma_smem_desc_sw128(smem_u32(smem + L::OFF_KV));
      const uint32_t d_s = tmem_base + TM_S + x * A4_BN;
      const uint32_t d_o = tmem_base + TM_O + x * A4_HD;
      const uint32_t a_p = tmem_base + TM_P + x * 64;   // 8 columns (16 keys) per MMA
      uint32_t it_q = 0, it_kv = 0;
      uint32_t cnt_s = 0, cnt_p = 0;   // S / PV products issued for this tile (phases of s_full,s_empty / p_full,pv_done)
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        if (!it.act_a) continue;
        const int qb = it_q & 1;
        const uint32_t q_phase = (it_q >> 1) & 1;
        const uint32_t kv_first = it_kv;
        it_kv += n_kv;
        ++it_q;
        if (x == 1 && !it.act_b) {
          // Tile B is absent from this item. Its issuer still CONSUMES the item's Q buffer and K/V stages -- waits for
          // each to fill, releases it unused -- so that its barrier phases advance in lock step with the producer. (It
          // used to skip the item while tile A's issuer signed the releases for it. A parity wait only tells "phase
          // n done" from "phase n+1 done": after skipping two items in a row -- padded-window segments -- this thread
          // could be two phases behind q_full, take a Q buffer that had not landed yet as ready and release buffers
          // that were still in use. Timing dependent: seen only with a second stream competing for the memory system.)
          mbar_wait(&q_full[qb], q_phase);
          for (int j = 0; j < n_kv; ++j) {
            const int s = (kv_first + j) % A4_STAGES;
            mbar_wait(&kv_full[s], ((kv_first + j) / A4_STAGES) & 1);
            mbar_arrive(&kv_empty[s]);
          }
          mbar_arrive(&q_empty[qb]);
          continue;
        }
        mbar_wait(&q_full[qb], q_phase);
        tc_fence_after();
        const uint64_t qdesc = q_desc0 + (uint64_t)((qb * L::Q_BUF + x * L::Q_TILE) >> 4);
        for (int j = 0; j <= n_kv; ++j) {
          if (j < n_kv) {
            const int s = (kv_first + j) % A4_STAGES;
            mbar_wait(&kv_full[s], ((kv_first + j) / A4_STAGES) & 1);
            mbar_wait(&s_empty[x], (cnt_s & 1) ^ 1);   // the softmax warps hold the previous S of this tile in registers
            tc_fence_after();
            const uint64_t kdesc = kv_desc0 + (uint64_t)((s * L::KV_BYTES) >> 4);
            umma_bf16(d_s, qdesc + 0, kdesc + 0, idesc_s, 0);
            umma_bf16(d_s, qdesc + 2, kdesc + 2, idesc_s, 1);
            umma_bf16(d_s, qdesc + 4, kdesc + 4, idesc_s, 1);
            umma_bf16(d_s, qdesc + 6, kdesc + 6, idesc_s, 1);
            umma_commit(&s_full[x]);
            ++cnt_s;
            if (j == n_kv - 1)                         // Q buffer reusable once the last S products retire
              umma_commit(&q_empty[qb]);
          }
          if (j > 0) {
            const int s = (kv_first + j - 1) % A4_STAGES;
            const uint64_t vdesc = kv_desc0 + (uint64_t)((s * L::KV_BYTES + L::K_BYTES) >> 4);
            constexpr uint64_t VH = (L::V_BYTES / 2) >> 4;    // second 64-key half
            mbar_wait(&p_full[x], cnt_p & 1);
            tc_fence_after();
            const uint32_t acc = (j - 1) > 0;
            umma_bf16_ts(d_o, a_p + 0, vdesc + 0, idesc_pv, acc);
            umma_bf16_ts(d_o, a_p + 8, vdesc + 2, idesc_pv, 1);
            umma_bf16_ts(d_o, a_p + 16, vdesc + 4, idesc_pv, 1);
            umma_bf16_ts(d_o, a_p + 24, vdesc + 6, idesc_pv, 1);
            umma_bf16_ts(d_o, a_p + 32, vdesc + VH + 0, idesc_pv, 1);
            umma_bf16_ts(d_o, a_p + 40, vdesc + VH + 2, idesc_pv, 1);
            umma_bf16_ts(d_o, a_p + 48, vdesc + VH + 4, idesc_pv, 1);
            umma_bf16_ts(d_o, a_p + 56, vdesc + VH + 6, idesc_pv, 1);
            umma_commit(&pv_done[x]);
            ++cnt_p;
            umma_commit(&kv_empty[s]);
          }
        }
      }
    }
  } else {
    reg_inc<216>();
    // ===================== softmax: thread = query row, all 128 scores of the row in registers =====================
    constexpr int NC = A4_BN;                   // score columns per thread
    constexpr int ND = A4_HD;                   // output dims per thread
    const int sw = warp - 4;
    const int x = sw >> 2;                      // query tile: 0 = A, 1 = B
    const int quarter = warp & 3;               // TMEM lane quarter
    const int r = quarter * 32 + lane;          // row inside the tile == TMEM lane
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + lane_addr + TM_S + x * A4_BN;
    const uint32_t t_p = tmem_base + lane_addr + TM_P + x * 64;
    const uint32_t t_o = tmem_base + lane_addr + TM_O + x * A4_HD;
    constexpr float LOG2E = 1.4426950408889634f;
    constexpr float RESCALE_TH = 8.0f / LOG2E;  // move the reference maximum only for growth beyond 2^8
    uint32_t cnt = 0;                           // steps processed by this tile (phases of all four barriers)

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item it = decode_item(p, item);
      if (!(x == 0 ? it.act_a : it.act_b)) continue;
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < n_kv; ++j, ++cnt) {
        const int valid = p.seg_len - j * A4_BN;   // >= 128 for all but a segment's last tile
        A4_TRACE(x, cnt, 0);
        mbar_wait(&s_full[x], cnt & 1);
        tc_fence_after();
        A4_TRACE(x, cnt, 1);
        uint32_t s[NC];
        float mt0 = -INFINITY, mt1 = -INFINITY;
        // full tiles read S in two halves: the row maximum of columns 0..63 is taken while the TMEM load of columns
        // 64..127 is in flight. Same values, same maximum.
        if (valid >= NC) {
          tmem_ld_32x32(t_s + 0, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
          tmem_ld_32x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
          tmem_ld_wait();
          tmem_ld_32x32(t_s + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
          tmem_ld_32x32(t_s + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
#pragma unroll
          for (int i = 0; i < NC / 2; i += 4) {
            mt0 = fmaxf(mt0, fmaxf(__uint_as_float(s[i]), __uint_as_float(s[i + 1])));
            mt1 = fmaxf(mt1, fmaxf(__uint_as_float(s[i + 2]), __uint_as_float(s[i + 3])));
          }
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&s_empty[x]);                // S of the next step may overwrite the buffer now
#pragma unroll
          for (int i = NC / 2; i < NC; i += 4) {
            mt0 = fmaxf(mt0, fmaxf(__uint_as_float(s[i]), __uint_as_float(s[i + 1])));
            mt1 = fmaxf(mt1, fmaxf(__uint_as_float(s[i + 2]), __uint_as_float(s[i + 3])));
          }
        } else
        {
#pragma unroll
          for (int c = 0; c < NC; c += 32) tmem_ld_32x32(t_s + c, *reinterpret_cast<uint32_t(*)[32]>(&s[c]));
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&s_empty[x]);                // S of the next step may overwrite the buffer now
          if (valid < NC) {                        // warp-uniform: keys beyond the segment do not exist
#pragma unroll
            for (int i = 0; i < NC; ++i)
              if (i >= valid) s[i] = 0xff800000u;  // -inf
          }
#pragma unroll
          for (int i = 0; i < NC; i += 4) {
            mt0 = fmaxf(mt0, fmaxf(__uint_as_float(s[i]), __uint_as_float(s[i + 1])));
            mt1 = fmaxf(mt1, fmaxf(__uint_as_float(s[i + 2]), __uint_as_float(s[i + 3])));
          }
        }
        const float mt = fmaxf(mt0, mt1);
        uint32_t pk[NC / 2];                      // P as packed 16-bit pairs
        float lsum = 0.f;
        if (j == 0) {
          m_ref = mt;
        } else if (__any_sync(0xffffffffu, mt > m_ref + RESCALE_TH)) {
          // rare: move the reference maximum and rescale O / l. Every earlier PV product of this tile must have retired.
          mbar_wait(&pv_done[x], (cnt - 1) & 1);
          tc_fence_after();
          const float m_new = fmaxf(m_ref, mt);
          const float alpha = ex2_mufu((m_ref - m_new) * LOG2E);
          l *= alpha;
          m_ref = m_new;
          const uint64_t a2 = pack2(alpha, alpha);
#pragma unroll
          for (int h = 0; h < ND; h += 32) {
            uint32_t o[32];
            tmem_ld_32x32(t_o + h, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const uint64_t v = mul2(pack2(__uint_as_float(o[i]), __uint_as_float(o[i + 1])), a2);
              float v0, v1;
              unpack2(v, v0, v1);
              o[i] = __float_as_uint(v0);
              o[i + 1] = __float_as_uint(v1);
            }
            tmem_st_32x32(t_o + h, o);
          }
          tmem_st_wait();
        }
        A4_TRACE(x, cnt, 2);
        lsum = softmax_row<POLY, F16>(s, -m_ref * LOG2E, pk);
        l += lsum;
        // the previous PV product of this tile reads P: it must have retired before P is overwritten (it was issued a
        // whole softmax step ago, so this wait does not stall in steady state)
        A4_TRACE(x, cnt, 4);
        if (j > 0) {
          mbar_wait(&pv_done[x], (cnt - 1) & 1);
          tc_fence_after();
        }
        A4_TRACE(x, cnt, 5);
#pragma unroll
        for (int c = 0; c < NC / 2; c += 32) tmem_st_32x32(t_p + c, *reinterpret_cast<uint32_t(*)[32]>(&pk[c]));
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[x]);
        A4_TRACE(x, cnt, 6);
      }
      // ---- item epilogue: O / l -> bf16 -> global
      mbar_wait(&pv_done[x], (cnt - 1) & 1);
      tc_fence_after();
      const float inv = 1.0f / l;
      const int qrow = it.q0 + x * A4_BM + r;     // row inside the segment
      const bool store = qrow < p.seg_len;
      const size_t out_off = (size_t)(it.seg * p.seg_len + qrow) * p.ldo + it.head * A4_HD;
      __nv_bfloat16* dst = p.out + out_off;
#pragma unroll
      for (int h = 0; h < ND; h += 32) {
        uint32_t o[32];
        tmem_ld_32x32(t_o + h, o);
        tmem_ld_wait();
        if (store) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + h + i) = u;
          }
          if (p.out_lo != nullptr) {   // split-precision projection: the rounding error of the bf16 store, as bf16
            auto lo2 = [&](int i) {
              const float a = __uint_as_float(o[i]) * inv, b = __uint_as_float(o[i + 1]) * inv;
              return pack_bf16x2(a - __bfloat162float(__float2bfloat16(a)), b - __bfloat162float(__float2bfloat16(b)));
            };
#pragma unroll
            for (int i = 0; i < 32; i += 8)
              *reinterpret_cast<uint4*>(p.out_lo + out_off + h + i) = make_uint4(lo2(i), lo2(i + 2), lo2(i + 4), lo2(i + 6));
          }
        }
      }
      tc_fence_before();   // the O reads are ordered before this thread's next p_full arrive, which gates the next PV
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

// Q, K: [heads][rows_total][64] bf16; Vt: [heads][64][rows_total] bf16; out: [rows_total][ldo] bf16.
// Segments with index (within their frame: seg % seg_period) >= q_part_from only need their first q_part_rows query rows
// (the rest are window-pad rows whose output the caller drops); pass q_part_from >= seg_period for "all rows".
// seg_period = segments per frame when rows_total holds a batch of frames (0: one frame).
void attention_tc(cudaStream_t st, const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* Vt,
                  __nv_bfloat16* out, int ldo, int heads, int rows_total, int seg_len, int q_part_from,
                  int q_part_rows, int seg_period, bool f16, __nv_bfloat16* out_lo) {
  CRA5_CHECK(seg_len > 0 && rows_total % seg_len == 0, ERR_INVALID, "attention: rows must be whole segments");
  CRA5_CHECK((rows_total & 7) == 0, ERR_INVALID, "attention: rows_total must be a multiple of 8 (TMA stride)");
  static const int poly = [] { const char* e = getenv("CRA5_ATTN_POLY"); return e ? atoi(e) : 3; }();
  auto kern = f16 ? attn_tc4_kernel<3, true>
                  : poly == 0 ? attn_tc4_kernel<0, false>
                  : poly == 2 ? attn_tc4_kernel<2, false>
                  : poly == 4 ? attn_tc4_kernel<4, false>
                              : attn_tc4_kernel<3, false>;
  ensure_dynamic_smem(kern, A4Smem::TOTAL);
  CUtensorMap tmQ = make_tmap_bf16_2d(Q, A4_HD, (uint64_t)heads * rows_total, A4_HD * 2, A4_HD, A4_BM);
  CUtensorMap tmK = make_tmap_bf16_2d(K, A4_HD, (uint64_t)heads * rows_total, A4_HD * 2, A4_HD, A4_BN);
  CUtensorMap tmVt = make_tmap_bf16_2d(Vt, (uint64_t)rows_total, (uint64_t)heads * A4_HD, (uint64_t)rows_total * 2,
                                       64, A4_HD);
  A4Params p;
  p.seg_len = seg_len;
  p.rows_total = rows_total;
  p.n_seg = rows_total / seg_len;
  p.heads = heads;
  p.n_qp = (seg_len + 2 * A4_BM - 1) / (2 * A4_BM);
  p.seg_period = (seg_period > 0 && seg_period <= p.n_seg) ? seg_period : p.n_seg;   // segments per frame
  p.q_part_from = (q_part_rows > 0 && q_part_rows < seg_len) ? q_part_from : p.seg_period;
  p.q_part_rows = q_part_rows;
  p.out = out;
  p.out_lo = out_lo;
  p.ldo = ldo;
  const int n_items = heads * p.n_seg * p.n_qp;
  const int grid = n_items < device_sm_count() ? n_items : device_sm_count();
  LaunchScope scope(st, "attn_tc", 4.0 * heads * (double)rows_total * seg_len * A4_HD,
                    4.0 * 2.0 * heads * (double)rows_total * A4_HD);
  launch_chained(kern, dim3(grid), dim3(A4_THREADS), A4Smem::TOTAL, st, tmQ, tmK, tmVt, p);
}

}  // namespace cra5
