/* cra5_b200.h -- C ABI of libcra5b200.so, the B200-native (sm_100a) implementation of the CRA5 / VAEformer
 * encode -> quantize -> entropy-code -> decode hot path.
 *
 * Conventions:
 *   - every function returns an int status (CRA5_OK == 0); the message of the last failure on the calling
 *     thread is returned by cra5_last_error();
 *   - pointers named *_dev are CUDA device pointers owned by the caller (e.g. torch `data_ptr()`); `stream` is a
 *     cudaStream_t passed as void* (NULL = default stream);
 *   - no torch / C++ types cross this boundary; the library owns its workspaces inside the model handle.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the CRA5 repository).
 */
#ifndef CRA5_B200_H
#define CRA5_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CRA5_API __attribute__((visibility("default")))
#else
#define CRA5_API
#endif

#define CRA5_OK 0
#define CRA5_ERR_INVALID 1   /* bad argument / unsupported geometry (reference raises ValueError) */
#define CRA5_ERR_CUDA 2      /* CUDA failure or no sm_100 device */
#define CRA5_ERR_STATE 3     /* e.g. "Uninitialized CDFs. Run update() first" (entropy_models.py:218-237) */
#define CRA5_ERR_BITSTREAM 4 /* malformed container / truncated stream */
#define CRA5_ERR_INTERNAL 5

CRA5_API const char* cra5_last_error(void);
CRA5_API int cra5_abi_version(void);

/* ------------------------------------------------------------------------------------------------------------
 * Kernel-level operators (used by the parity tests; the model entry points below are built from them)
 * ---------------------------------------------------------------------------------------------------------- */

/* epilogue kinds of cra5_op_gemm */
#define CRA5_EPI_F32 0       /* out f32  = A*B^T + bias                           (nn.Linear, vit_nlc.py:57-59) */
#define CRA5_EPI_BF16 1      /* out bf16 = A*B^T + bias                                                          */
#define CRA5_EPI_GELU_BF16 2 /* out bf16 = gelu_erf(A*B^T + bias)                 (Mlp.fc1+act, vit_nlc.py:63-64) */
#define CRA5_EPI_RESID 4     /* out f32  = resid + A*B^T + bias                   (Block residual, vit_nlc.py:284) */
#define CRA5_EPI_T_F32 5     /* out f32 [N][ldo] = (A*B^T + bias)^T               (1x1 conv -> NCHW, vaeformer.py:154) */

/* C[M,N] = A[M,K] * B[N,K]^T with A, B bf16 row-major (lda/ldb in elements, multiples of 8), fp32 accumulate on
 * tcgen05 tensor cores. Replaces the cuBLAS GEMM behind torch.nn.Linear / 1x1 Conv2d in the reference. */
CRA5_API int cra5_op_gemm(const void* A_dev, int lda, const void* B_dev, int ldb, int M, int N, int K,
                 const float* bias_dev, int epilogue, void* out_dev, int ldo, const float* resid_dev, void* stream);

/* Same contract, plain SIMT fp32-FMA kernel. Self-check only (never on the product path). */
CRA5_API int cra5_op_gemm_check(const void* A_dev, int lda, const void* B_dev, int ldb, int M, int N, int K,
                       const float* bias_dev, float* out_dev, int ldo, void* stream);

/* out[row, head*64 + d] = softmax(Q K^T) V per (segment, head); Q pre-scaled. Q,K: [heads][rows][64] bf16,
 * Vt: [heads][64][rows] bf16, rows = n_segments * seg_len. Replaces Attention.forward (vit_nlc.py:94-112) and the
 * per-window attention of WindowAttention.forward (vit_nlc.py:242-246). */
CRA5_API int cra5_op_attention(const void* Q_dev, const void* K_dev, const void* Vt_dev, void* out_dev, int ldo, int heads,
                      int rows_total, int seg_len, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRA5_B200_H */
