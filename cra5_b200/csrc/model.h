// model.h -- the VAEformer codec as a sequence of sm_100a kernel launches over library-owned workspaces.
// Host-side mirror of cra5/models/vaeformer/vaeformer.py:272-400 (encode_latent / decode_latent /
// compress_from_latent / decompress). One Model per (GPU, stream); no global state.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/cra5_b200.h"
#include "gemm_tc.cuh"

namespace cra5 {

struct TensorRef {
  const void* ptr = nullptr;
  int dtype = 0;  // CRA5_DT_*
  int64_t numel = 0;
};

struct BlockWeights {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *qkv_b, *proj_b, *fc1_b, *fc2_b;
  const __nv_bfloat16 *qkv_w, *proj_w, *fc1_w, *fc2_w;
  // split-bf16 copies [2][N][K] (hi, lo) of the same weights; null unless uploaded ("<name>.x3", precision levels > 0)
  const __nv_bfloat16 *qkv_w3 = nullptr, *proj_w3 = nullptr, *fc1_w3 = nullptr, *fc2_w3 = nullptr;
  bool has_split() const { return qkv_w3 && proj_w3 && fc1_w3 && fc2_w3; }
};

struct CdfTable {
  const int32_t* cdf = nullptr;     // [rows][cols] device
  const int32_t* length = nullptr;  // [rows] device
  const int32_t* offset = nullptr;  // [rows] device
  int rows = 0, cols = 0;
  bool ready() const { return cdf != nullptr && rows > 0; }
};

// Scratch for one transformer trunk (main or hyper)
class RansCoder;

struct TrunkBuffers {
  float* x;                 // [T][D] residual stream
  __nv_bfloat16* a;         // [Tpad][D] LayerNorm output (attention order) / generic bf16 A operand
  __nv_bfloat16 *q, *k, *vt;  // [heads][Tpad][hd], [heads][hd][Tpad]
  __nv_bfloat16* o;         // [Tpad][D] attention output
  __nv_bfloat16* h;         // [T][mlp*D]
  // split-bf16 precision mode (allocated by Model::set_precision): A operands carried as [2][...] hi | lo halves
  __nv_bfloat16* a2 = nullptr;   // [2][Tpad][D]
  size_t a_half = 0;
  __nv_bfloat16* h2 = nullptr;   // [2][T][mlp*D]
  size_t h_half = 0;
  float* f32 = nullptr;          // [T][mlp*D] pre-activation of fc1 (GELU + split run as their own kernel)
  __nv_bfloat16* o2 = nullptr;   // [2][Tpad][D] attention output hi | lo
  size_t o_half = 0;
};

class Model {
 public:
  explicit Model(const cra5_config& cfg);
  ~Model();
  Model(const Model&) = delete;

  void set_tensor(const std::string& name, const void* ptr, int dtype, int64_t numel);
  void set_cdf(int which, const int32_t* cdf, const int32_t* len, const int32_t* off, int rows, int cols);
  void set_coder(int spc_y, int spc_z);
  // 0 = bf16 operands everywhere (default); 1 = split-bf16 ("bf16x3") GEMMs on the encoder tail (last two g_a blocks,
  // quant_conv) and the whole hyperprior; 2 = + patch-embed and every g_a block (everything that decides the symbols);
  // 3 = + the decoder. Needs the "<name>.x3" weight copies of the sites it covers.
  void set_precision(int level);
  int precision() const { return precision_; }

  // hot path (all pointers device, fp32). B frames (<= max_batch, contiguous in every tensor) run as ONE launch per
  // kernel: B * tokens rows through every GEMM / LayerNorm / attention, B * channels through the entropy kernels.
  void encode_to_latent(const float* x, float* y, const float* mean, const float* std_, int B, cudaStream_t st);
  // mean / std (per channel, device) optionally fuse x * std + mean into the un-patchify store
  void latent_to_reconstruction(const float* y_hat, float* x_hat, int B, cudaStream_t st, const float* mean = nullptr,
                                const float* std_ = nullptr);
  void latent_quantized(const float* y, float* y_hat, cudaStream_t st);  // encode_latent(type='quantized') tail
  // same, plus the likelihood tensors of the eval-mode forward (rate estimation); any output may be null
  void latent_likelihoods(const float* y, float* y_hat, float* y_lik, float* z_lik, cudaStream_t st);
  // entropy stage; returns (per frame) pinned host buffers owned by the model, valid until the next call
  void latent_to_bin(const float* y, int B, const uint8_t** y_bytes, size_t* y_len, const uint8_t** z_bytes,
                     size_t* z_len, cudaStream_t st);
  void bin_to_latent(const uint8_t* const* y_bytes, const size_t* y_len, const uint8_t* const* z_bytes,
                     const size_t* z_len, int B, int z_h, int z_w, float* y_hat, cudaStream_t st);
  int max_batch() const { return Bm; }
  // debug taps for the parity tests (device pointers into the workspace, valid until the next call)
  const void* tap(const std::string& name, int64_t* numel, int* dtype) const;

  const cra5_config& config() const { return cfg_; }
  size_t workspace_bytes() const { return ws_bytes_; }
  // geometry
  int Hg, Wg, T;        // token grid of the main trunk
  int Hh, Wh, Th;       // hyper grid
  int Tpad;             // max rows of any window-partitioned order
  int hd, hdh;          // head dims
  int kpr, cs_pad, box_rows;  // patch-embed implicit-GEMM geometry
  int nA, nB;           // ConvTranspose kernel-row classes (non-overlapping rows, overlapping rows)
  int Bm = 1;           // frames per call the workspace is sized for
  int M2 = 0;           // rows per frame of the reconstruction head's operand buffer

 private:
  void finalize();  // resolve tensor pointers (first use)
  const void* need(const std::string& name, int dtype, int64_t numel) const;
  BlockWeights block_weights(const std::string& prefix, int D, int mlp) const;
  WinMap make_winmap(int block_window_h, int block_window_w) const;
  void run_block(cudaStream_t st, const BlockWeights& w, const TrunkBuffers& tb, const float* x_in, float* x_out, int T_,
                 int D, int heads, int mlp, int win_h, int win_w, __nv_bfloat16* cat_out, int cat_col0, int cat_ld,
                 bool precise, int frames, bool main);
  const __nv_bfloat16* need_x3(const std::string& name, int64_t numel) const;  // "<name>.x3": [2][numel] bf16
  void* alloc2(size_t bytes);
  void run_h_a(cudaStream_t st, const float* y, int B);     // -> z_ [B][zc][Th]
  void run_h_s(cudaStream_t st, const float* z_hat, int B); // -> params_ [B][sigma | mu][T]
  void* alloc(size_t bytes);

  cra5_config cfg_;
  std::map<std::string, TensorRef> tensors_;
  bool finalized_ = false;
  std::vector<BlockWeights> ga_, gs_, ha_, hs_;
  CdfTable eb_, gc_;
  int spc_y_, spc_z_;

  uint8_t* ws_ = nullptr;
  size_t ws_bytes_ = 0, ws_used_ = 0;
  int precision_ = 0;
  uint8_t* ws2_ = nullptr;          // split-precision workspace (set_precision)
  size_t ws2_bytes_ = 0, ws2_used_ = 0;
  __nv_bfloat16 *cat2_ = nullptr, *patches2_ = nullptr, *ytok2_ = nullptr, *ztok2_ = nullptr, *ah2_ = nullptr,
                *fin2_ = nullptr;
  size_t cat_half_ = 0, patches_half_ = 0, ytok_half_ = 0, ztok_half_ = 0, ah_half_ = 0, fin_half_ = 0;
  TrunkBuffers main_, hyper_;
  float *x1_, *x2_;                 // outputs of the two parallel head blocks
  __nv_bfloat16* cat_;              // [T][2D] bf16 (mean || logvar tokens)
  __nv_bfloat16* patches_;          // [Himg][Wg][cs_pad] bf16
  float *y_, *yhat_, *params_;      // [latent][T], [latent][T], [2*latent][T]
  __nv_bfloat16* ytok_;             // [T][latent]
  float *z_, *zhat_;                // [zc][Th]
  __nv_bfloat16* ztok_;             // [Th][zc]
  __nv_bfloat16* ah_;               // hyper im2col / generic A operand [Th][max K]
  __nv_bfloat16* fin_;              // final LayerNorm of g_s, [frame][M2][D] with zero rows T..M2-1 per frame
  int32_t *ysym_, *zsym_;
  uint8_t* yidx_;
  RansCoder* coder_ = nullptr;
  uint8_t *host_y_ = nullptr, *host_z_ = nullptr;  // pinned output containers
  size_t host_y_cap_ = 0, host_z_cap_ = 0, frame_cap_y_ = 0, frame_cap_z_ = 0;   // one container slot per frame
  std::map<std::string, TensorRef> taps_;
};

}  // namespace cra5
