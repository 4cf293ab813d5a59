// kernels.h -- host-callable launchers of the sm_100a kernels (internal C++ interface).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace cra5 {

struct GemmShape;
struct EpiParams;

int gemm_pick_bn(int N);
void launch_gemm(cudaStream_t st, int bn, int kind, const CUtensorMap& tmA, const CUtensorMap& tmB,
                 const GemmShape& shp, const EpiParams& epi);
void gemm_plain(cudaStream_t st, int kind, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M,
                int N, int K, const EpiParams& epi);
void gemm_simt_check(cudaStream_t st, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb,
                     const float* bias, float* C, int ldc, int M, int N, int K);

// fused softmax(QK^T)V, head_dim 64 (attn_tc.cu)
void attention_tc(cudaStream_t st, const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* Vt,
                  __nv_bfloat16* out, int ldo, int heads, int rows_total, int seg_len);

}  // namespace cra5
