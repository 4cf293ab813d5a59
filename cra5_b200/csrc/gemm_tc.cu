// gemm_tc.cu -- host launcher for the tcgen05 GEMM (gemm_tc.cuh) + a plain SIMT GEMM used only by the
// GPU self-checks (tests compare the tensor-core kernel against it and against the CPU oracle).
#include "gemm_tc.cuh"
#include "host_util.h"
#include "kernels.h"
#include <stdlib.h>

namespace cra5 {

template <int BN, int KIND>
static void launch_one(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& shp,
                       const EpiParams& epi) {
  auto kern = gemm_tc_kernel<BN, KIND>;
  ensure_dynamic_smem(kern, GemmSmem<BN>::TOTAL);
  const int m_tiles = (shp.M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (shp.N + BN - 1) / BN;
  int grid = m_tiles * n_tiles;
  const int sms = device_sm_count();
  if (grid > sms) grid = sms;
  // (split-bf16 launches execute 2-3x the tensor work of their algorithmic flops, the grouped un-patchify order 32/30 of
  // them for its zero-weight pad columns; the profiler counts the algorithmic ones)
  double n_alg = shp.N;
  if (KIND == EPI_CONVT && epi.ct_cpg > 0) n_alg = (double)(shp.N / epi.ct_CS) * epi.ct_C * epi.ct_pw;
  LaunchScope scope(st, (shp.a_split || shp.b_split) ? "gemm_tc_split" : "gemm_tc", 2.0 * shp.M * n_alg * shp.K,
                    2.0 * ((double)shp.M * shp.K + (double)shp.N * shp.K) +
                        (double)shp.M * n_alg * ((KIND == EPI_BF16 || KIND == EPI_GELU_BF16 || KIND == EPI_QKV || KIND == EPI_QKV_F16) ? 2.0 : 4.0));
  launch_chained(kern, dim3(grid), dim3(GEMM_THREADS), GemmSmem<BN>::TOTAL, st, tmA, tmB, shp, epi);
}

template <int KIND>
static void launch_pair(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& shp,
                        const EpiParams& epi) {
  auto kern = gemm_tc2_kernel<KIND>;
  ensure_dynamic_smem(kern, GemmSmem2::TOTAL);
  const int m_tiles = (shp.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const int n_tiles = (shp.N + GemmSmem2::BN - 1) / GemmSmem2::BN;
  int clusters = m_tiles * n_tiles;
  const int max_clusters = device_sm_count() / 2;
  if (clusters > max_clusters) clusters = max_clusters;
  double n_alg = shp.N;   // the grouped un-patchify order carries zero-weight pad columns: count the algorithmic ones
  if (KIND == EPI_CONVT && epi.ct_cpg > 0) n_alg = (double)(shp.N / epi.ct_CS) * epi.ct_C * epi.ct_pw;
  LaunchScope scope(st, "gemm_tc", 2.0 * shp.M * n_alg * shp.K,
                    2.0 * ((double)shp.M * shp.K + (double)shp.N * shp.K) +
                        (double)shp.M * n_alg * ((KIND == EPI_BF16 || KIND == EPI_GELU_BF16 || KIND == EPI_QKV) ? 2.0 : 4.0));
  launch_chained(kern, dim3(2 * clusters), dim3(GEMM_THREADS), GemmSmem2::TOTAL, st, tmA, tmB, shp, epi);  // cluster dims are compiled in
}

void launch_gemm_pair(cudaStream_t st, int kind, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& shp,
                      const EpiParams& epi) {
  switch (kind) {
    case EPI_F32: launch_pair<EPI_F32>(st, tmA, tmB, shp, epi); break;
    case EPI_BF16: launch_pair<EPI_BF16>(st, tmA, tmB, shp, epi); break;
    case EPI_GELU_BF16: launch_pair<EPI_GELU_BF16>(st, tmA, tmB, shp, epi); break;
    case EPI_QKV: launch_pair<EPI_QKV>(st, tmA, tmB, shp, epi); break;
    case EPI_RESID: launch_pair<EPI_RESID>(st, tmA, tmB, shp, epi); break;
    case EPI_T_F32: launch_pair<EPI_T_F32>(st, tmA, tmB, shp, epi); break;
    case EPI_PIXSHUF: launch_pair<EPI_PIXSHUF>(st, tmA, tmB, shp, epi); break;
    case EPI_CONVT: launch_pair<EPI_CONVT>(st, tmA, tmB, shp, epi); break;
    default: throw Error(ERR_INTERNAL, "unknown epilogue kind");
  }
}

// CTA-pair tiles (256 x 256, each SM stages half of B) where the shape is large enough for them. The choice depends on
// N, K and the epilogue only, never on M beyond a floor: a batch of frames must take the same kernel as one frame so
// that results stay bit-identical across batch sizes.
bool gemm_use_pair(int M, int N, int K, int kind) {
  static const char* env = getenv("CRA5_GEMM_PAIR");  // diagnostics: 0 = never, 1 = whenever legal
  if (env != nullptr) return atoi(env) != 0 && N >= 256;
  // measured on B200 (bench.py kernel_sites, 8 frames per call): qkv 1381 TFLOP/s, fc2 1368, proj 822 with the pair
  // kernel against 1343 / 1281 / 711-773 with 1-CTA tiles. Since the GELU and un-patchify epilogues got cheaper
  // (end of round 2) the main loop shows there too: fc1 1234 -> 1247, un-patchify class A 0.571 -> 0.542 ms per frame
  // (the 1-CTA 128 x 256 tile pulls 48 KB per k-block from L2, the pair 32 KB per CTA; ~56 B per clock and SM is what
  // the GEMM captures show L2 delivering). The small hyperprior GEMMs stay 1-CTA.
  if (kind == EPI_QKV && N >= 3072 && M >= 4096) return true;
  if (kind == EPI_GELU_BF16 && N >= 3072 && K >= 1024 && M >= 4096) return true;
  if (kind == EPI_CONVT && N >= 4096 && K >= 1024 && M >= 4096) return true;
  return kind == EPI_RESID && N >= 512 && N <= 1024 && K >= 1024 && M >= 4096;
}

template <int BN>
static void launch_kind(cudaStream_t st, int kind, const CUtensorMap& tmA, const CUtensorMap& tmB,
                        const GemmShape& shp, const EpiParams& epi) {
  switch (kind) {
    case EPI_F32: launch_one<BN, EPI_F32>(st, tmA, tmB, shp, epi); break;
    case EPI_BF16: launch_one<BN, EPI_BF16>(st, tmA, tmB, shp, epi); break;
    case EPI_GELU_BF16: launch_one<BN, EPI_GELU_BF16>(st, tmA, tmB, shp, epi); break;
    case EPI_QKV: launch_one<BN, EPI_QKV>(st, tmA, tmB, shp, epi); break;
    case EPI_QKV_F16: launch_one<BN, EPI_QKV_F16>(st, tmA, tmB, shp, epi); break;
    case EPI_RESID: launch_one<BN, EPI_RESID>(st, tmA, tmB, shp, epi); break;
    case EPI_T_F32: launch_one<BN, EPI_T_F32>(st, tmA, tmB, shp, epi); break;
    case EPI_PIXSHUF: launch_one<BN, EPI_PIXSHUF>(st, tmA, tmB, shp, epi); break;
    case EPI_CONVT: launch_one<BN, EPI_CONVT>(st, tmA, tmB, shp, epi); break;
    default: throw Error(ERR_INTERNAL, "unknown epilogue kind");
  }
}

int gemm_pick_bn(int N) {
  static const char* env = getenv("CRA5_GEMM_BN");  // diagnostics only
  if (env != nullptr && N > 128) return atoi(env) == 128 ? 128 : 256;
  return N > 128 ? 256 : 128;
}

void launch_gemm(cudaStream_t st, int bn, int kind, const CUtensorMap& tmA, const CUtensorMap& tmB,
                 const GemmShape& shp, const EpiParams& epi) {
  CRA5_CHECK(shp.M > 0 && shp.N > 0 && shp.K > 0, ERR_INVALID, "gemm: empty problem");
  if (bn == 256)
    launch_kind<256>(st, kind, tmA, tmB, shp, epi);
  else if (bn == 128)
    launch_kind<128>(st, kind, tmA, tmB, shp, epi);
  else
    throw Error(ERR_INTERNAL, "gemm: unsupported BN");
}

// 2D operand map (k, row), or 3D (k, row, half) when the operand is carried as split bf16 halves `half_elems` apart
static CUtensorMap operand_map(const __nv_bfloat16* p, int K, int rows, int ld, int box_rows, size_t half_elems) {
  if (half_elems == 0) return make_tmap_bf16_2d(p, (uint64_t)K, (uint64_t)rows, (uint64_t)ld * 2, GEMM_BK, box_rows);
  uint64_t dims[3] = {(uint64_t)K, (uint64_t)rows, 2};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)half_elems * 2};
  uint32_t box[3] = {GEMM_BK, (uint32_t)box_rows, 1};
  return make_tmap_bf16(p, 3, dims, strides, box, true);
}

// Plain row-major GEMM convenience: A [M,K] (row stride lda elements), B [N,K] (row stride ldb), both bf16.
void gemm_plain(cudaStream_t st, int kind, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M,
                int N, int K, const EpiParams& epi, const GemmSplit& split) {
  if (!split.any() && gemm_use_pair(M, N, K, kind)) {
    CUtensorMap tmA = make_tmap_bf16_2d(A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, GEMM_BK, GEMM_BM);
    CUtensorMap tmB = make_tmap_bf16_2d(B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, GEMM_BK, GemmSmem2::BN / 2);
    GemmShape shp{};
    shp.M = M; shp.N = N; shp.K = K; shp.a_mode = A_PLAIN;
    launch_gemm_pair(st, kind, tmA, tmB, shp, epi);
    return;
  }
  int bn = gemm_pick_bn(N);
  CUtensorMap tmA = operand_map(A, K, M, lda, GEMM_BM, split.a_half);
  CUtensorMap tmB = operand_map(B, K, N, ldb, bn, split.b_half);
  GemmShape shp{};
  shp.M = M; shp.N = N; shp.K = K; shp.a_mode = A_PLAIN;
  shp.a_split = split.a_half != 0; shp.b_split = split.b_half != 0;
  launch_gemm(st, bn, kind, tmA, tmB, shp, epi);
}

// ---------------------------------------------------------------- SIMT check kernel (tests only)
__global__ void gemm_simt_check_kernel(const __nv_bfloat16* __restrict__ A, int lda,
                                       const __nv_bfloat16* __restrict__ B, int ldb, const float* __restrict__ bias,
                                       float* __restrict__ C, int ldc, int M, int N, int K) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k)
    acc = fmaf(__bfloat162float(A[(size_t)m * lda + k]), __bfloat162float(B[(size_t)n * ldb + k]), acc);
  if (bias) acc += bias[n];
  C[(size_t)m * ldc + n] = acc;
}

void gemm_simt_check(cudaStream_t st, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb,
                     const float* bias, float* C, int ldc, int M, int N, int K) {
  dim3 block(128), grid((N + 127) / 128, M);
  LaunchScope scope(st, "gemm_simt_check", 2.0 * M * N * K, 0.0);
  gemm_simt_check_kernel<<<grid, block, 0, st>>>(A, lda, B, ldb, bias, C, ldc, M, N, K);
  CRA5_CUDA(cudaGetLastError());
}

}  // namespace cra5
