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
// split-bf16 operands: element distance from the hi half to the lo half of A / B (0 = that operand is plain bf16)
struct GemmSplit {
  size_t a_half = 0, b_half = 0;
  bool any() const { return a_half != 0 || b_half != 0; }
};
void gemm_plain(cudaStream_t st, int kind, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M,
                int N, int K, const EpiParams& epi, const GemmSplit& split = GemmSplit());
void gemm_simt_check(cudaStream_t st, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb,
                     const float* bias, float* C, int ldc, int M, int N, int K);

// fused softmax(QK^T)V, head_dim 64. attention_tc (attn_tc4.cu) is the persistent ping-pong kernel; segments with
// index >= q_part_from only need their first q_part_rows query rows (window-pad rows).
void attention_tc(cudaStream_t st, const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* Vt,
                  __nv_bfloat16* out, int ldo, int heads, int rows_total, int seg_len, int q_part_from = 1 << 30,
                  int q_part_rows = 0, int seg_period = 0, bool f16 = false, __nv_bfloat16* out_lo = nullptr);

// f16: Q, K, V hold fp16 bits (EPI_QKV_F16) -- routed to the fp32-softmax shared-memory kernel; out_lo: optional
// bf16(O - bf16(O)) for a split-precision projection
void attention_simt(cudaStream_t st, const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* Vt,
                    __nv_bfloat16* out, int ldo, int heads, int hd, int rows_total, int seg_len, bool f16 = false,
                    __nv_bfloat16* out_lo = nullptr);

// elementwise.cu
struct WinMap;
// (every A-operand producer takes an optional `lo` output for the split-bf16 precision mode: v ~ bf16 hi + bf16 lo)
void frame_to_patches(cudaStream_t st, const float* x, __nv_bfloat16* out, const float* mean, const float* std_,
                      int C, int H, int W, int Wp, int pw, int cs_pad, __nv_bfloat16* out_lo = nullptr);
void layernorm_bf16(cudaStream_t st, const float* x, const float* gamma, const float* beta, float eps,
                    __nv_bfloat16* out, int rows_out, int D, const WinMap& wm, __nv_bfloat16* out_lo = nullptr);
void im2col_latent(cudaStream_t st, const float* y, __nv_bfloat16* A, int C, int Hy, int Wy, int p1, int p2, int lda,
                   __nv_bfloat16* A_lo = nullptr);
void transpose_cast(cudaStream_t st, const float* in, __nv_bfloat16* out, int C, int T, int ldo,
                    __nv_bfloat16* out_lo = nullptr);
void cast_bf16(cudaStream_t st, const float* in, __nv_bfloat16* out, size_t n);
void split_rows(cudaStream_t st, const float* in, int ld_in, int rows, int cols, __nv_bfloat16* hi, __nv_bfloat16* lo,
                int ld_out, bool gelu);
void affine_channels(cudaStream_t st, const float* in, float* out, const float* a, const float* b, size_t hw, int C,
                     int forward);

// entropy.cu
// One launch covers `frames` frames of n elements each: y / sym / idx / y_hat advance by n per frame, sigma and mu by
// param_stride (the hyperprior writes sigma | mu per frame, so consecutive frames' sigma are 2n apart).
void gc_quantize_index(cudaStream_t st, const float* y, const float* sigma, const float* mu, const float* scale_table,
                       int levels, float bound, int32_t* sym, uint8_t* idx, float* y_hat, size_t n, int frames = 1,
                       size_t param_stride = 0);
// (n_ch = channels per frame: the median index is (e / L) % n_ch, so a batch of frames is one launch)
void eb_quantize(cudaStream_t st, const float* z, const float* median, int L, int n_ch, int32_t* sym, float* z_hat,
                 size_t n);
void dequantize(cudaStream_t st, const int32_t* sym, const float* mu, const float* median, int L, float* out,
                size_t n);
void gc_likelihood(cudaStream_t st, const float* y, const float* sigma, const float* mu, float scale_bound,
                   float lik_bound, float* y_hat, float* lik, size_t n);
void eb_likelihood(cudaStream_t st, const float* z_hat, const float* packed, int L, float lik_bound, float* lik,
                   size_t n);
// chan_mod: 0 = the CDF row of a symbol is idx[pos]; > 0 = it is (channel % chan_mod) -- EntropyBottleneck, where a
// batch of frames is coded as frames * chan_mod channels. Decoders: ch_per_frame / mu_frame_extra place the per-symbol
// means of frame f at mu[pos + f * mu_frame_extra] (the hyperprior writes sigma | mu per frame).
void rans_encode(cudaStream_t st, const int32_t* sym, const uint8_t* idx, int chan_mod, const int32_t* cdf,
                 int cdf_stride, const int32_t* cdf_len, const int32_t* offset, int n_channels, int L, int spc,
                 int chan_len, uint32_t* scratch, int cap_words, uint32_t* lengths, uint32_t* offsets, uint8_t* payload, int* err);
void rans_decode(cudaStream_t st, const uint8_t* payload, const uint32_t* offsets, const uint8_t* idx,
                 int chan_mod, const int32_t* cdf, int cdf_stride, const int32_t* cdf_len,
                 const int32_t* offset, const uint16_t* lut, int lut_rows, int n_channels, int L, int spc,
                 int chan_len, int32_t* sym_out, const float* mu, const float* median, float* val_out, int* err,
                 int ch_per_frame = 0, size_t mu_frame_extra = 0);
void build_decode_lut(cudaStream_t st, const int32_t* cdf, int cdf_stride, const int32_t* cdf_len, int rows,
                      uint16_t* lut);
void scan_lengths(cudaStream_t st, const uint32_t* lengths, int n, uint32_t* offsets);
// shared-memory-table variants (CDF rows packed as uint16, see entropy.cu)
void pack_cdf(cudaStream_t st, const int32_t* cdf, int cdf_stride, const int32_t* cdf_len, int rows, int32_t* row_off,
              uint16_t* packed, int cap, int* total_dev);
bool rans_tables_fit(int rows, int total, bool with_lut);
void rans_encode_smem(cudaStream_t st, const int32_t* sym, const uint8_t* idx, int chan_mod,
                      const uint16_t* packed, const int32_t* row_off, const int32_t* cdf_len, const int32_t* offset,
                      int rows, int total, int n_channels, int L, int spc, int chan_len, uint32_t* scratch, int cap_words,
                      uint32_t* lengths, uint32_t* offsets, uint8_t* payload, int* err);
void rans_decode_smem(cudaStream_t st, const uint8_t* payload, const uint32_t* offsets, const uint8_t* idx,
                      int chan_mod, const uint16_t* packed, const int32_t* row_off, const int32_t* cdf_len,
                      const int32_t* offset, const uint16_t* lut, int rows, int total, int n_channels, int L, int spc,
                      int chan_len, int32_t* sym_out, const float* mu, const float* median, float* val_out, int* err,
                      int ch_per_frame = 0, size_t mu_frame_extra = 0);

}  // namespace cra5
