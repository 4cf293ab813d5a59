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
#define CRA5_EPI_GELU_BF16_TRUNK 9 /* the same function as 2, evaluated the way the trunk's fc1 epilogue does at the
                              * default precision: x * Phi(x) with Phi through one MUFU.TANH of an odd quintic fitted to
                              * the erf form (|error| <= 3e-5, below the bf16 rounding of the output); kinds 6-8 are
                              * internal (scatter epilogues of the model entry points) */
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

/* Same contract for any even head_dim (fp32 SIMT kernel; the hyperprior transformer's 5 heads x 72, vit_nlc.py:94-112
 * through HyperpriorEncoder/Decoder). Q,K: [heads][rows][head_dim], Vt: [heads][head_dim][rows]. */
CRA5_API int cra5_op_attention_generic(const void* Q_dev, const void* K_dev, const void* Vt_dev, void* out_dev, int ldo,
                                       int heads, int head_dim, int rows_total, int seg_len, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Host utilities (run once per model, like the reference's C++ helpers)
 * ---------------------------------------------------------------------------------------------------------- */

/* float pmf -> strictly increasing uint32 CDF with cdf[n] == 1 << precision. Replaces
 * compressai._CXX.pmf_to_quantized_cdf (cpp_exts/ops/ops.cpp:40-109, bound at :111-118; caller
 * entropy_models.py:89-92). `cdf_out` holds n + 1 entries. CRA5_ERR_INVALID for negative / non-finite / all-zero
 * pmf (the reference raises ValueError through std::domain_error). */
CRA5_API int cra5_pmf_to_quantized_cdf(const float* pmf, int n, int precision, uint32_t* cdf_out);

/* ------------------------------------------------------------------------------------------------------------
 * Model handle
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct cra5_config {
  int32_t in_chans;          /* vaeformer.py:106 */
  int32_t img_h, img_w;      /* vaeformer.py:119 */
  int32_t patch_h, patch_w;  /* vaeformer.py:104 */
  int32_t stride_h, stride_w;/* vaeformer.py:105 */
  int32_t dim, depth, num_heads, mlp_ratio; /* vit_nlc.py:1009-1015 */
  int32_t n_windows;         /* vaeformer.py:112 */
  int32_t window_h[4], window_w[4];
  int32_t interval;          /* vaeformer.py:113 */
  int32_t latent_chans;      /* vaeformer.py:94 */
  int32_t z_chans;           /* vaeformer.py:95 */
  int32_t hyper_dim, hyper_depth, hyper_heads; /* vaeformer.py:129-131 */
  int32_t hyper_patch_h, hyper_patch_w;        /* vaeformer.py:124 */
  float ln_eps;              /* vit_nlc.py:381 */
  int32_t streams_per_channel_y; /* chunk-parallel coder: interleaved rANS sub-streams per latent channel */
  int32_t streams_per_channel_z;
  int32_t max_batch;         /* frames per call the workspace is sized for (0 or 1: one frame). The reference API is
                              * batched throughout (vaeformer.py:350-376; entropy_models.py:263-272 loops batch items). */
} cra5_config;

typedef struct cra5_model cra5_model;

#define CRA5_DT_F32 0
#define CRA5_DT_BF16 1
#define CRA5_DT_I32 2
#define CRA5_DT_U8 3

/* Allocates the workspace on the current CUDA device. Replaces VAEformer.__init__ (vaeformer.py:78-166). */
CRA5_API int cra5_model_create(const cra5_config* cfg, cra5_model** out);
CRA5_API int cra5_model_destroy(cra5_model* m);
CRA5_API int cra5_model_workspace_bytes(cra5_model* m, uint64_t* bytes);

/* Hand a parameter to the model by its state-dict name (device pointer, caller keeps it alive). Matmul weights are
 * bf16, everything else fp32; repacked conv weights use the names documented in cra5_b200/vaeformer.py.
 * Replaces nn.Module.load_state_dict (models/base.py:69-89). */
CRA5_API int cra5_model_set_tensor(cra5_model* m, const char* name, const void* dev_ptr, int dtype, int64_t numel);

/* CDF tables as built by update() / shipped in a checkpoint: which = 0 EntropyBottleneck, 1 GaussianConditional.
 * Device int32 pointers: cdf [rows][cols], cdf_length [rows], offset [rows]
 * (entropy_models.py:127-129 buffers `_quantized_cdf`, `_cdf_length`, `_offset`). */
CRA5_API int cra5_model_set_cdf(cra5_model* m, int which, const int32_t* cdf_dev, const int32_t* cdf_length_dev,
                                const int32_t* offset_dev, int rows, int cols);

/* Interleaved rANS sub-streams per latent channel for y and z (parallelism / rate knob of the CR5B container). */
CRA5_API int cra5_model_set_coder(cra5_model* m, int streams_per_channel_y, int streams_per_channel_z);

/* Arithmetic of the linear / conv layers. The reference computes them in fp32 (nn.Linear / Conv2d on torch CPU,
 * vit_nlc.py:63-67, 96, 111, 302, 629; vaeformer.py:154-155); the default here (level 0) feeds bf16 operands to the
 * tensor cores with fp32 accumulation. Levels > 0 switch groups of GEMM sites to the split-bf16 form (every fp32
 * operand carried as bf16 hi + bf16 lo, three tcgen05 products per k-block into one fp32 accumulator: ~16 mantissa
 * bits per operand), which is what makes the quantised SYMBOLS agree with the fp32 reference:
 *   1  encoder tail (the last two g_a blocks, quant_conv) + the whole hyperprior (h_a, h_s)
 *   2  1 + patch-embed conv and every g_a block: every layer the bitstream depends on
 *   3  2 + the decoder (post_quant_conv, g_s, reconstruction head)
 * A level needs the split copies "<name>.x3" (bf16 [2][N][K]: hi, lo) of the weights it covers, handed over with
 * cra5_model_set_tensor beforehand; ERR_STATE otherwise. Attention inside a covered block runs on fp16 Q/K/V/P (the
 * format of the reference's own GPU path, flash-attn on .half() tensors, vit_nlc.py:105-110) with an fp32 softmax and
 * passes its output to the projection as a bf16 hi + lo pair. */
CRA5_API int cra5_model_set_precision(cra5_model* m, int level);

/* x (C,H,W) fp32 -> y (latent, Hg, Wg) fp32. mean/std: optional per-channel (C) device arrays; when given the input
 * is physical-unit data and (x-mean)/std (cra5_api.normalization, cra5_api.py:264-266) is fused into the first kernel.
 * Replaces VAEformer.encode_latent(type='float') (vaeformer.py:272-292) = cra5_api.encode_to_latent (:53-71). */
CRA5_API int cra5_encode_to_latent(cra5_model* m, const float* x_dev, float* y_dev, const float* mean_dev,
                                   const float* std_dev, void* stream);

/* y -> y_hat = round(y - mu) + mu with (sigma, mu) = h_s(round(h_a(y) - median) + median): the `type='quantized'`
 * tail of VAEformer.encode_latent (vaeformer.py:284-290). */
CRA5_API int cra5_latent_quantized(cra5_model* m, const float* y_dev, float* y_hat_dev, void* stream);

/* Eval-mode forward up to the latent with the rate-estimation outputs: y_hat as above, y_lik (latent, Hg, Wg) =
 * GaussianConditional likelihoods and z_lik (z_chans, Hh, Wh) = EntropyBottleneck likelihoods, both floored at 1e-9
 * (vaeformer.py:314-319; entropy_models.py:465-510, 645-677). Any output may be NULL. */
CRA5_API int cra5_latent_likelihoods(cra5_model* m, const float* y_dev, float* y_hat_dev, float* y_lik_dev,
                                     float* z_lik_dev, void* stream);

/* y -> {y string, z string}. The returned pointers are pinned host buffers owned by the handle, valid until the next
 * call on it. Replaces VAEformer.compress_from_latent (vaeformer.py:334-348) = cra5_api.latent_to_bin (:73-79), i.e.
 * h_a, EntropyBottleneck.compress, h_s, build_indexes, GaussianConditional.compress and the C++ coder behind them
 * (entropy_models.py:239-272 -> compressai.ans RansEncoder.encode_with_indexes, rans_interface.cpp:202-213). */
CRA5_API int cra5_latent_to_bin(cra5_model* m, const float* y_dev, const uint8_t** y_bytes, uint64_t* y_len,
                                const uint8_t** z_bytes, uint64_t* z_len, void* stream);

/* {y string, z string} (host memory) -> y_hat (latent, Hg, Wg) fp32 on the device. Replaces
 * VAEformer.decompress(return_format='latent') (vaeformer.py:378-391) = cra5_api.bin_to_latent (:127-144), i.e.
 * EntropyBottleneck.decompress, h_s, build_indexes, GaussianConditional.decompress and RansDecoder.decode_with_indexes
 * (rans_interface.cpp:215-284). */
CRA5_API int cra5_bin_to_latent(cra5_model* m, const uint8_t* y_bytes, uint64_t y_len, const uint8_t* z_bytes,
                                uint64_t z_len, int z_h, int z_w, float* y_hat_dev, void* stream);

/* y_hat -> x_hat (C,H,W) fp32, normalised units. Replaces VAEformer.decode_latent (vaeformer.py:294-300) =
 * cra5_api.latent_to_reconstruction (:146-151). */
CRA5_API int cra5_latent_to_reconstruction(cra5_model* m, const float* y_hat_dev, float* x_hat_dev, void* stream);

/* The same four calls on a batch of `batch` <= cfg.max_batch frames, contiguous in every tensor: x (B,C,H,W),
 * y / y_hat (B,latent,Hg,Wg), x_hat (B,C,H,W). The batch runs as ONE launch of every kernel -- B x tokens rows through
 * each GEMM / LayerNorm / attention, B x channels sub-streams through the entropy kernels -- and the results are
 * bit-identical to B single-frame calls (fixed tile order, no split-K). The reference is batched the same way:
 * VAEformer.compress / decompress take (B, C, H, W) tensors (vaeformer.py:350-400) and EntropyModel.compress loops the
 * batch items (entropy_models.py:263-272). y_bytes[b] / z_bytes[b] are per-frame containers (one pinned slot per frame,
 * valid until the next call on the handle); all four arrays hold `batch` entries. */
CRA5_API int cra5_encode_to_latent_batch(cra5_model* m, const float* x_dev, float* y_dev, const float* mean_dev,
                                         const float* std_dev, int batch, void* stream);
CRA5_API int cra5_latent_to_bin_batch(cra5_model* m, const float* y_dev, int batch, const uint8_t** y_bytes,
                                      uint64_t* y_len, const uint8_t** z_bytes, uint64_t* z_len, void* stream);
CRA5_API int cra5_bin_to_latent_batch(cra5_model* m, const uint8_t* const* y_bytes, const uint64_t* y_len,
                                      const uint8_t* const* z_bytes, const uint64_t* z_len, int batch, int z_h, int z_w,
                                      float* y_hat_dev, void* stream);
CRA5_API int cra5_latent_to_reconstruction_batch(cra5_model* m, const float* y_hat_dev, float* x_hat_dev, int batch,
                                                 void* stream);

/* cra5_latent_to_reconstruction_batch with the de-normalisation of cra5_api.decode_from_bin(return_format=
 * 'de_normalized') fused into the last kernel's stores: x_hat[c] = x_hat_normalised[c] * std[c] + mean[c]
 * (cra5_api.de_normalization, cra5_api.py:268-271) -- no separate read-modify-write pass over the 1.1 GB frame.
 * mean / std: per-channel (C) device arrays. */
CRA5_API int cra5_latent_to_reconstruction_denorm(cra5_model* m, const float* y_hat_dev, float* x_hat_dev,
                                                  const float* mean_dev, const float* std_dev, int batch, void* stream);

/* per-channel (x - mean)/std (forward != 0) or x*std + mean (forward == 0); in == out allowed.
 * Replaces cra5_api.normalization / de_normalization (cra5_api.py:264-271). */
CRA5_API int cra5_normalize(const float* in_dev, float* out_dev, const float* mean_dev, const float* std_dev,
                            int channels, uint64_t hw, int forward, void* stream);

/* Debug tap for the parity tests: device pointer of an intermediate of the LAST call ("y", "z", "z_hat", "scales",
 * "means", "y_symbols", "y_indexes", "z_symbols", "tokens"). */
CRA5_API int cra5_model_tap(cra5_model* m, const char* name, const void** dev_ptr, int64_t* numel, int* dtype);

CRA5_API int cra5_model_tap_read(cra5_model* m, const char* name, void* dst_dev, uint64_t dst_bytes, void* stream);

/* Measurement hooks (bench.py): number of kernels this library has launched on the calling thread; an optional
 * CUDA-event profiler that brackets every launch on its own stream and reports, per kernel and call site, launches,
 * summed device milliseconds and algorithmic flops / bytes as a JSON object. */
CRA5_API int cra5_launch_count(uint64_t* count);
CRA5_API int cra5_profile_enable(int on);
CRA5_API int cra5_profile_report(char* buf, uint64_t cap, uint64_t* needed);

/* ------------------------------------------------------------------------------------------------------------
 * Entropy-stage operators on caller-provided device buffers (parity tests drive these directly)
 * ---------------------------------------------------------------------------------------------------------- */

/* symbols = round_half_even(y - mu), index = scale-table bucket of max(sigma, bound), y_hat = symbols + mu. Any of
 * y/sym/idx/y_hat may be NULL to skip that part. Replaces EntropyModel.quantize (entropy_models.py:155-184) +
 * GaussianConditional.build_indexes (:679-685). */
CRA5_API int cra5_op_gc_quantize(const float* y_dev, const float* sigma_dev, const float* mu_dev,
                                 const float* scale_table_dev, int levels, float bound, int32_t* sym_dev,
                                 uint8_t* idx_dev, float* y_hat_dev, uint64_t n, void* stream);

/* Chunk-parallel rANS over a (n_channels, L) int32 symbol tensor. idx_dev NULL => index == channel
 * (EntropyBottleneck). Writes the container (header, stream lengths, payload) to out_host (capacity out_cap) and its
 * size to out_len. Each sub-stream is bit-identical to RansEncoder.encode_with_indexes (rans_interface.cpp:202-213)
 * applied to that sub-stream's symbols. */
CRA5_API int cra5_op_rans_encode(const int32_t* sym_dev, const uint8_t* idx_dev, const int32_t* cdf_dev, int cdf_cols,
                                 const int32_t* cdf_length_dev, const int32_t* offset_dev, int n_channels, int L,
                                 int spc, uint8_t* out_host, uint64_t out_cap, uint64_t* out_len, void* stream);

/* Inverse of cra5_op_rans_encode: container (host) -> int32 symbols (device). */
CRA5_API int cra5_op_rans_decode(const uint8_t* bytes_host, uint64_t len, const uint8_t* idx_dev,
                                 const int32_t* cdf_dev, int cdf_cols, const int32_t* cdf_length_dev,
                                 const int32_t* offset_dev, int n_channels, int L, int32_t* sym_dev, void* stream);

/* The same two operations with the number of rows of the CDF table stated (cdf_dev is [cdf_rows][cdf_cols]): the coder
 * then packs the table into shared memory (uint16 rows + coarse inverse table), the path cra5_latent_to_bin /
 * cra5_bin_to_latent take. Same bytes, same symbols. */
CRA5_API int cra5_op_rans_encode_table(const int32_t* sym_dev, const uint8_t* idx_dev, const int32_t* cdf_dev,
                                       int cdf_rows, int cdf_cols, const int32_t* cdf_length_dev,
                                       const int32_t* offset_dev, int n_channels, int L, int spc, uint8_t* out_host,
                                       uint64_t out_cap, uint64_t* out_len, void* stream);
CRA5_API int cra5_op_rans_decode_table(const uint8_t* bytes_host, uint64_t len, const uint8_t* idx_dev,
                                       const int32_t* cdf_dev, int cdf_rows, int cdf_cols,
                                       const int32_t* cdf_length_dev, const int32_t* offset_dev, int n_channels, int L,
                                       int32_t* sym_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRA5_B200_H */
