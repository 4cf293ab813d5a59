// model.cu -- VAEformer encode / entropy-code / decode as kernel launches (see model.h).
#include "model.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "coder.h"
#include "host_util.h"
#include "kernels.h"

namespace cra5 {

static int ceil_div(int a, int b) { return (a + b - 1) / b; }
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Model::Model(const cra5_config& c) : cfg_(c) {
  require_sm100();
  CRA5_CHECK(c.in_chans > 0 && c.dim > 0 && c.depth >= 2 && c.num_heads > 0, ERR_INVALID, "config: sizes");
  CRA5_CHECK(c.dim % c.num_heads == 0 && c.hyper_dim % c.hyper_heads == 0, ERR_INVALID,
             "config: width must be divisible by heads");
  CRA5_CHECK(c.dim % 8 == 0 && c.hyper_dim % 8 == 0 && c.latent_chans % 8 == 0 && c.z_chans % 8 == 0, ERR_INVALID,
             "unsupported geometry: widths must be multiples of 8");
  CRA5_CHECK(c.patch_w == c.stride_w, ERR_INVALID, "unsupported geometry: horizontal patch overlap");
  CRA5_CHECK(c.stride_h <= c.patch_h && c.patch_h <= 2 * c.stride_h, ERR_INVALID,
             "unsupported geometry: vertical patch size must be in [stride, 2*stride]");
  CRA5_CHECK(c.n_windows >= 1 && c.n_windows <= 4 && c.interval >= 1, ERR_INVALID, "config: windows");
  Hg = (c.img_h - c.patch_h) / c.stride_h + 1;
  Wg = (c.img_w - c.patch_w) / c.stride_w + 1;
  T = Hg * Wg;
  CRA5_CHECK(Hg % c.hyper_patch_h == 0 && Wg % c.hyper_patch_w == 0, ERR_INVALID,
             "unsupported geometry: token grid not divisible by the hyperprior patch");
  Hh = Hg / c.hyper_patch_h;
  Wh = Wg / c.hyper_patch_w;
  Th = Hh * Wh;
  hd = c.dim / c.num_heads;
  hdh = c.hyper_dim / c.hyper_heads;
  CRA5_CHECK(hd == 64, ERR_INVALID, "unsupported geometry: trunk head_dim must be 64 (tcgen05 attention tile)");
  // window-partitioned row counts
  Tpad = T;
  for (int i = 0; i < c.n_windows; ++i) {
    const int rows = ceil_div(Hg, c.window_h[i]) * c.window_h[i] * ceil_div(Wg, c.window_w[i]) * c.window_w[i];
    CRA5_CHECK(rows % 8 == 0, ERR_INVALID, "unsupported geometry: padded window rows must be a multiple of 8");
    Tpad = std::max(Tpad, rows);
  }
  CRA5_CHECK(T % 8 == 0, ERR_INVALID, "unsupported geometry: token count must be a multiple of 8");
  // patch-embed implicit GEMM
  const int CS = c.in_chans * c.patch_w;
  kpr = ceil_div(CS, GEMM_BK);
  // row pitch of the re-laid-out frame: a multiple of 64 elements, so that every 128-byte row of a TMA box is ONE aligned
  // L2 line (with the minimal 16-byte padding the 5360-byte pitch made each of them straddle two)
  cs_pad = (int)align_up((size_t)CS, 64);
  box_rows = 0;
  for (int b = 128; b >= 16; b >>= 1)   // >= 16 rows per TMA box keeps a 128-row tile within 8 boxes
    if (Wg % b == 0) { box_rows = b; break; }
  CRA5_CHECK(box_rows != 0, ERR_INVALID, "unsupported geometry: patches per row must be a multiple of 16");
  nB = c.patch_h - c.stride_h;
  nA = c.stride_h - nB;
  spc_y_ = c.streams_per_channel_y > 0 ? c.streams_per_channel_y : 16;
  spc_z_ = c.streams_per_channel_z > 0 ? c.streams_per_channel_z : 4;
  CRA5_CHECK(spc_y_ <= CR5B_MAX_SPC && spc_z_ <= CR5B_MAX_SPC, ERR_INVALID, "streams per channel must be <= 64");

  // ---------------- workspace (every activation buffer holds max_batch frames: a batch is ONE launch per kernel)
  Bm = c.max_batch > 0 ? c.max_batch : 1;
  CRA5_CHECK(Bm <= 64, ERR_INVALID, "config: max_batch must be <= 64");
  const int D = c.dim, Dh = c.hyper_dim, mlp = c.mlp_ratio, lat = c.latent_chans, zc = c.z_chans;
  const int hidden = std::max(1, (int)sqrt((double)(Dh / zc))) * zc;
  const size_t kh_max = std::max<size_t>({(size_t)lat * c.hyper_patch_h * c.hyper_patch_w, (size_t)Dh, (size_t)hidden,
                                          (size_t)zc});
  const size_t nb = (size_t)Bm;
  M2 = (Hg + 1) * Wg;   // rows per frame of the reconstruction head's operand: Hg patch rows + one zero row block
  auto trunk_bytes = [&](size_t T_, size_t Tp, int D_, int mlp_) {
    return align_up(T_ * D_ * 4, 256) + 5 * align_up(Tp * D_ * 2, 256) + align_up(T_ * mlp_ * D_ * 2, 256);
  };
  const int hmlp = std::max(mlp, ceil_div(hidden, Dh));
  size_t total = 0;
  total += trunk_bytes(nb * T, nb * Tpad, D, mlp) + trunk_bytes(nb * Th, nb * Th, Dh, hmlp);
  total += 2 * align_up(nb * T * D * 4, 256) + align_up(nb * T * 2 * D * 2, 256);
  total += align_up(nb * c.img_h * Wg * cs_pad * 2, 256);
  total += 2 * align_up(nb * lat * T * 4, 256) + align_up(nb * 2 * lat * T * 4, 256) + align_up(nb * T * lat * 2, 256);
  total += 2 * align_up(nb * zc * Th * 4, 256) + align_up(nb * Th * zc * 2, 256) + align_up(nb * Th * kh_max * 2, 256);
  total += align_up(nb * lat * T * 4, 256) + align_up(nb * zc * Th * 4, 256) + align_up(nb * lat * T, 256);
  total += align_up(nb * M2 * D * 2, 256);
  total += 4096;
  ws_bytes_ = total;
  CRA5_CUDA(cudaMalloc(&ws_, ws_bytes_));
  CRA5_CUDA(cudaMemset(ws_, 0, ws_bytes_));
  CRA5_CUDA(cudaDeviceSynchronize());   // the memset runs on the legacy stream; callers use their own streams
  auto trunk_alloc = [&](TrunkBuffers& tb, size_t T_, size_t Tp, int D_, int mlp_) {
    tb.x = (float*)alloc(T_ * D_ * 4);
    tb.a = (__nv_bfloat16*)alloc(Tp * D_ * 2);
    tb.q = (__nv_bfloat16*)alloc(Tp * D_ * 2);
    tb.k = (__nv_bfloat16*)alloc(Tp * D_ * 2);
    tb.vt = (__nv_bfloat16*)alloc(Tp * D_ * 2);
    tb.o = (__nv_bfloat16*)alloc(Tp * D_ * 2);
    tb.h = (__nv_bfloat16*)alloc(T_ * mlp_ * D_ * 2);
  };
  trunk_alloc(main_, nb * T, nb * Tpad, D, mlp);
  trunk_alloc(hyper_, nb * Th, nb * Th, Dh, hmlp);
  x1_ = (float*)alloc(nb * T * D * 4);
  x2_ = (float*)alloc(nb * T * D * 4);
  cat_ = (__nv_bfloat16*)alloc(nb * T * 2 * D * 2);
  patches_ = (__nv_bfloat16*)alloc(nb * c.img_h * Wg * cs_pad * 2);
  y_ = (float*)alloc(nb * lat * T * 4);
  yhat_ = (float*)alloc(nb * lat * T * 4);
  params_ = (float*)alloc(nb * 2 * lat * T * 4);
  ytok_ = (__nv_bfloat16*)alloc(nb * T * lat * 2);
  z_ = (float*)alloc(nb * zc * Th * 4);
  zhat_ = (float*)alloc(nb * zc * Th * 4);
  ztok_ = (__nv_bfloat16*)alloc(nb * Th * zc * 2);
  ah_ = (__nv_bfloat16*)alloc(nb * Th * kh_max * 2);
  ysym_ = (int32_t*)alloc(nb * lat * T * 4);
  zsym_ = (int32_t*)alloc(nb * zc * Th * 4);
  yidx_ = (uint8_t*)alloc(nb * lat * T);
  fin_ = (__nv_bfloat16*)alloc(nb * M2 * D * 2);   // final LayerNorm of g_s, [frame][M2][D]; rows T..M2-1 stay zero
  coder_ = new RansCoder(nb * lat * T, Bm * std::max(lat, zc));
  frame_cap_y_ = align_up(RansCoder::max_container_bytes((size_t)lat * T, lat * CR5B_MAX_SPC), 256);
  frame_cap_z_ = align_up(RansCoder::max_container_bytes((size_t)zc * Th, zc * CR5B_MAX_SPC), 256);
  host_y_cap_ = nb * frame_cap_y_;
  host_z_cap_ = nb * frame_cap_z_;
  CRA5_CUDA(cudaHostAlloc(&host_y_, host_y_cap_, cudaHostAllocMapped));   // the coder's kernels write into these
  CRA5_CUDA(cudaHostAlloc(&host_z_, host_z_cap_, cudaHostAllocMapped));
}

Model::~Model() {
  delete coder_;
  cudaFree(ws_);
  cudaFree(ws2_);
  cudaFreeHost(host_y_);
  cudaFreeHost(host_z_);
}

void* Model::alloc(size_t bytes) {
  const size_t off = ws_used_;
  ws_used_ += align_up(bytes, 256);
  CRA5_CHECK(ws_used_ <= ws_bytes_, ERR_INTERNAL, "workspace sizing");
  return ws_ + off;
}

void* Model::alloc2(size_t bytes) {
  const size_t off = ws2_used_;
  ws2_used_ += align_up(bytes, 256);
  CRA5_CHECK(ws2_used_ <= ws2_bytes_, ERR_INTERNAL, "split workspace sizing");
  return ws2_ + off;
}

// Precision levels (see model.h). The split-bf16 mode carries every fp32 GEMM operand as hi = bf16(v), lo = bf16(v - hi)
// and accumulates A_hi B_hi + A_lo B_hi + A_hi B_lo in the same fp32 TMEM accumulator (gemm_tc.cuh, GemmShape::a_split):
// ~16 mantissa bits per operand instead of 8, on the same tcgen05 pipeline, at 3x the tensor work of the covered sites.
// Attention in a split-precision block runs on fp16 Q, K, V and P -- 11 significand bits instead of 8, the format the
// reference's own GPU path uses (flash-attn on .half() tensors, vit_nlc.py:105-110) -- and hands its output to the
// projection as a bf16 hi | lo pair; blocks outside the level keep bf16 operands.
void Model::set_precision(int level) {
  CRA5_CHECK(level >= 0 && level <= 3, ERR_INVALID, "precision level must be 0 (bf16), 1 (tail), 2 (encoder) or 3 (all)");
  if (level > 0 && ws2_ == nullptr) {
    const cra5_config& c = cfg_;
    const int D = c.dim, Dh = c.hyper_dim, mlp = c.mlp_ratio, lat = c.latent_chans, zc = c.z_chans;
    const int hidden = std::max(1, (int)sqrt((double)(Dh / zc))) * zc;
    const size_t kh_max = std::max<size_t>({(size_t)lat * c.hyper_patch_h * c.hyper_patch_w, (size_t)Dh, (size_t)hidden,
                                            (size_t)zc});
    const size_t hw = std::max<size_t>((size_t)mlp * Dh, (size_t)hidden);   // widest hyper MLP activation
    const size_t nb = (size_t)Bm;
    main_.a_half = nb * Tpad * D;
    main_.h_half = nb * T * mlp * D;
    main_.o_half = nb * Tpad * D;
    hyper_.a_half = nb * Th * std::max<size_t>(Dh, kh_max);
    hyper_.h_half = nb * Th * hw;
    hyper_.o_half = nb * Th * Dh;
    cat_half_ = nb * T * 2 * D;
    patches_half_ = nb * c.img_h * Wg * cs_pad;
    ytok_half_ = nb * T * lat;
    ztok_half_ = nb * Th * zc;
    ah_half_ = nb * Th * kh_max;
    fin_half_ = nb * M2 * D;
    size_t total = 4096;
    for (size_t half : {main_.a_half, main_.h_half, main_.o_half, hyper_.a_half, hyper_.h_half, hyper_.o_half, cat_half_,
                        patches_half_, ytok_half_, ztok_half_, ah_half_, fin_half_})
      total += align_up(2 * half * 2, 256);
    total += align_up(nb * T * mlp * D * 4, 256) + align_up(nb * Th * hw * 4, 256);
    CRA5_CUDA(cudaMalloc(&ws2_, total));
    ws2_bytes_ = total;
    CRA5_CUDA(cudaMemset(ws2_, 0, total));
    CRA5_CUDA(cudaDeviceSynchronize());   // the memset runs on the legacy stream; callers use their own streams
    main_.a2 = (__nv_bfloat16*)alloc2(2 * main_.a_half * 2);
    main_.h2 = (__nv_bfloat16*)alloc2(2 * main_.h_half * 2);
    main_.f32 = (float*)alloc2(nb * T * mlp * D * 4);
    main_.o2 = (__nv_bfloat16*)alloc2(2 * main_.o_half * 2);
    hyper_.a2 = (__nv_bfloat16*)alloc2(2 * hyper_.a_half * 2);
    hyper_.h2 = (__nv_bfloat16*)alloc2(2 * hyper_.h_half * 2);
    hyper_.f32 = (float*)alloc2(nb * Th * hw * 4);
    hyper_.o2 = (__nv_bfloat16*)alloc2(2 * hyper_.o_half * 2);
    cat2_ = (__nv_bfloat16*)alloc2(2 * cat_half_ * 2);
    patches2_ = (__nv_bfloat16*)alloc2(2 * patches_half_ * 2);
    ytok2_ = (__nv_bfloat16*)alloc2(2 * ytok_half_ * 2);
    ztok2_ = (__nv_bfloat16*)alloc2(2 * ztok_half_ * 2);
    ah2_ = (__nv_bfloat16*)alloc2(2 * ah_half_ * 2);
    fin2_ = (__nv_bfloat16*)alloc2(2 * fin_half_ * 2);
  }
  precision_ = level;
  finalized_ = false;   // block weights pick up their ".x3" copies
}

const __nv_bfloat16* Model::need_x3(const std::string& name, int64_t numel) const {
  auto it = tensors_.find(name + ".x3");
  CRA5_CHECK(it != tensors_.end(), ERR_STATE,
             "precision level " + std::to_string(precision_) + " needs the split weight copy '" + name +
                 ".x3' (VAEformer.set_precision uploads it)");
  CRA5_CHECK(it->second.dtype == CRA5_DT_BF16 && it->second.numel == 2 * numel, ERR_INVALID,
             "split weight '" + name + ".x3' must be bf16 [2][N][K]");
  return (const __nv_bfloat16*)it->second.ptr;
}

void Model::set_tensor(const std::string& name, const void* ptr, int dtype, int64_t numel) {
  CRA5_CHECK(ptr != nullptr && numel > 0, ERR_INVALID, "set_tensor: null tensor '" + name + "'");
  CRA5_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, ERR_INVALID, "set_tensor: '" + name + "' must be 16-byte aligned");
  tensors_[name] = TensorRef{ptr, dtype, numel};
  finalized_ = false;
}

void Model::set_cdf(int which, const int32_t* cdf, const int32_t* len, const int32_t* off, int rows, int cols) {
  CRA5_CHECK(which == 0 || which == 1, ERR_INVALID, "set_cdf: which must be 0 (EntropyBottleneck) or 1 (GaussianConditional)");
  CRA5_CHECK(cdf && len && off && rows > 0 && cols >= 3, ERR_INVALID, "set_cdf: invalid CDF size");
  CdfTable& t = which == 0 ? eb_ : gc_;
  if (which == 0) CRA5_CHECK(rows == cfg_.z_chans, ERR_INVALID, "set_cdf: EntropyBottleneck needs one row per z channel");
  if (which == 1) CRA5_CHECK(rows <= 256, ERR_INVALID, "set_cdf: at most 256 scale levels");
  t.cdf = cdf; t.length = len; t.offset = off; t.rows = rows; t.cols = cols;
  if (coder_ != nullptr) coder_->invalidate_lut();
}

void Model::set_coder(int spc_y, int spc_z) {
  // 0 selects the reference's single sequential stream (interop with archives / the PyTorch reference)
  CRA5_CHECK(spc_y >= 0 && spc_y <= CR5B_MAX_SPC && spc_z >= 0 && spc_z <= CR5B_MAX_SPC, ERR_INVALID,
             "streams per channel must be in [0, 64]");
  spc_y_ = spc_y;
  spc_z_ = spc_z;
}

const void* Model::need(const std::string& name, int dtype, int64_t numel) const {
  auto it = tensors_.find(name);
  CRA5_CHECK(it != tensors_.end(), ERR_STATE, "missing parameter '" + name + "' (load_state_dict incomplete)");
  CRA5_CHECK(it->second.dtype == dtype, ERR_INVALID, "parameter '" + name + "' has the wrong dtype");
  CRA5_CHECK(it->second.numel == numel, ERR_INVALID,
             "size mismatch for '" + name + "': expected " + std::to_string(numel) + " elements, got " +
                 std::to_string(it->second.numel));
  return it->second.ptr;
}

BlockWeights Model::block_weights(const std::string& p, int D, int mlp) const {
  BlockWeights w;
  auto f = [&](const std::string& n, int64_t ne) { return (const float*)need(p + n, CRA5_DT_F32, ne); };
  auto b = [&](const std::string& n, int64_t ne) { return (const __nv_bfloat16*)need(p + n, CRA5_DT_BF16, ne); };
  w.ln1_g = f(".norm1.weight", D); w.ln1_b = f(".norm1.bias", D);
  w.ln2_g = f(".norm2.weight", D); w.ln2_b = f(".norm2.bias", D);
  w.qkv_w = b(".attn.qkv.weight", (int64_t)3 * D * D); w.qkv_b = f(".attn.qkv.bias", 3 * D);
  w.proj_w = b(".attn.proj.weight", (int64_t)D * D); w.proj_b = f(".attn.proj.bias", D);
  w.fc1_w = b(".mlp.fc1.weight", (int64_t)mlp * D * D); w.fc1_b = f(".mlp.fc1.bias", mlp * D);
  w.fc2_w = b(".mlp.fc2.weight", (int64_t)mlp * D * D); w.fc2_b = f(".mlp.fc2.bias", D);
  auto x3 = [&](const std::string& n, int64_t ne) -> const __nv_bfloat16* {
    auto it = tensors_.find(p + n + ".x3");
    if (it == tensors_.end()) return nullptr;
    CRA5_CHECK(it->second.dtype == CRA5_DT_BF16 && it->second.numel == 2 * ne, ERR_INVALID,
               "split weight '" + p + n + ".x3' must be bf16 [2][N][K]");
    return (const __nv_bfloat16*)it->second.ptr;
  };
  w.qkv_w3 = x3(".attn.qkv.weight", (int64_t)3 * D * D);
  w.proj_w3 = x3(".attn.proj.weight", (int64_t)D * D);
  w.fc1_w3 = x3(".mlp.fc1.weight", (int64_t)mlp * D * D);
  w.fc2_w3 = x3(".mlp.fc2.weight", (int64_t)mlp * D * D);
  return w;
}

void Model::finalize() {
  if (finalized_) return;
  const cra5_config& c = cfg_;
  const int n_enc = c.depth / 2 + 1, n_dec = c.depth - c.depth / 2;
  ga_.clear(); gs_.clear(); ha_.clear(); hs_.clear();
  for (int i = 0; i < n_enc; ++i) ga_.push_back(block_weights("g_a.blocks." + std::to_string(i), c.dim, c.mlp_ratio));
  for (int i = 0; i < n_dec; ++i) gs_.push_back(block_weights("g_s.blocks." + std::to_string(i), c.dim, c.mlp_ratio));
  for (int i = 0; i < c.hyper_depth / 2; ++i)
    ha_.push_back(block_weights("h_a.blocks." + std::to_string(i), c.hyper_dim, c.mlp_ratio));
  for (int i = 0; i < c.hyper_depth - c.hyper_depth / 2; ++i)
    hs_.push_back(block_weights("h_s.blocks." + std::to_string(i), c.hyper_dim, c.mlp_ratio));
  finalized_ = true;
}

// window (wh, ww) partition of the Hg x Wg grid, zero padded to whole windows (vit_nlc.py:229-237)
WinMap Model::make_winmap(int wh, int ww) const {
  WinMap m{};
  if (wh <= 0) return m;  // global block
  m.enabled = 1;
  m.H = Hg; m.W = Wg; m.wh = wh; m.ww = ww;
  m.nWr = ceil_div(Hg, wh);
  m.nWc = ceil_div(Wg, ww);
  m.finish();   // multiply-shift constants for the four divisions of to_token()
  return m;
}

// Block.forward (vit_nlc.py:282-287): x_out = x_in + attn(LN1(x_in)); x_out += mlp(LN2(x_out))
// T_ = frames * (tokens per frame): the rows of a batch of frames run through every kernel as one problem
void Model::run_block(cudaStream_t st, const BlockWeights& w, const TrunkBuffers& tb, const float* x_in, float* x_out,
                      int T_, int D, int heads, int mlp, int win_h, int win_w, __nv_bfloat16* cat_out, int cat_col0,
                      int cat_ld, bool precise, int frames, bool main) {
  const int hd_ = D / heads;
  WinMap wm = main ? make_winmap(win_h, win_w) : WinMap{};
  const int rows = wm.enabled ? frames * wm.nWr * wm.nWc * wm.wh * wm.ww : T_;
  const int seg = wm.enabled ? wm.wh * wm.ww : T_ / frames;
  if (precise) CRA5_CHECK(w.has_split() && tb.a2 != nullptr, ERR_STATE, "split weights / workspace missing for a precise block");
  // split-bf16 mode: LayerNorm emits hi | lo halves, the GEMMs run three (proj: two) terms per k-block; the GELU and the
  // bf16 casts that the default path fuses into GEMM epilogues run as one extra pass (split_rows) over fp32 outputs
  __nv_bfloat16* ln_out = precise ? tb.a2 : tb.a;
  layernorm_bf16(st, x_in, w.ln1_g, w.ln1_b, cfg_.ln_eps, ln_out, rows, D, wm, precise ? tb.a2 + tb.a_half : nullptr);
  {
    EpiParams e{};
    e.bias = w.qkv_b;
    e.q = tb.q; e.k = tb.k; e.vt = tb.vt;
    e.D = D; e.hd = hd_; e.rows_total = rows;
    e.qscale = 1.0f / sqrtf((float)hd_);
    TagScope tag_("qkv");
    if (precise)   // Q, K, V leave the epilogue as fp16
      gemm_plain(st, EPI_QKV_F16, ln_out, D, w.qkv_w3, D, rows, 3 * D, D, e, GemmSplit{tb.a_half, (size_t)3 * D * D});
    else
      gemm_plain(st, EPI_QKV, tb.a, D, w.qkv_w, D, rows, 3 * D, D, e);
  }
  if (hd_ == 64) {
    // bottom-row windows of a vertically padded grid: their trailing rows are pad tokens (vit_nlc.py:229-237). They
    // still act as keys / values, but their own outputs are cropped away (:252-256), so those query tiles are skipped.
    int part_from = 1 << 30, part_rows = 0;
    if (wm.enabled && Hg % wm.wh != 0 && Wg % wm.ww == 0) {
      part_from = (wm.nWr - 1) * wm.nWc;
      part_rows = (Hg - (wm.nWr - 1) * wm.wh) * wm.ww;
    }
    attention_tc(st, tb.q, tb.k, tb.vt, precise ? tb.o2 : tb.o, D, heads, rows, seg, part_from, part_rows,
                 wm.enabled ? wm.nWr * wm.nWc : 1, precise, precise ? tb.o2 + tb.o_half : nullptr);
  }
  else
    attention_simt(st, tb.q, tb.k, tb.vt, precise ? tb.o2 : tb.o, D, heads, hd_, rows, seg, precise,
                   precise ? tb.o2 + tb.o_half : nullptr);
  {
    EpiParams e{};
    e.bias = w.proj_b;
    e.resid = x_in; e.out_f32 = x_out; e.ldo = D; e.wm = wm;
    TagScope tag_("proj");
    if (precise)
      gemm_plain(st, EPI_RESID, tb.o2, D, w.proj_w3, D, rows, D, D, e, GemmSplit{tb.o_half, (size_t)D * D});
    else
      gemm_plain(st, EPI_RESID, tb.o, D, w.proj_w, D, rows, D, D, e);
  }
  layernorm_bf16(st, x_out, w.ln2_g, w.ln2_b, cfg_.ln_eps, ln_out, T_, D, WinMap{}, precise ? tb.a2 + tb.a_half : nullptr);
  if (precise) {
    {
      EpiParams e{};
      e.bias = w.fc1_b;
      e.out_f32 = tb.f32; e.ldo = mlp * D;
      TagScope tag_("fc1");
      gemm_plain(st, EPI_F32, ln_out, D, w.fc1_w3, D, T_, mlp * D, D, e, GemmSplit{tb.a_half, (size_t)mlp * D * D});
    }
    split_rows(st, tb.f32, mlp * D, T_, mlp * D, tb.h2, tb.h2 + tb.h_half, mlp * D, true);
    {
      EpiParams e{};
      e.bias = w.fc2_b;
      e.resid = x_out; e.out_f32 = x_out; e.ldo = D;
      TagScope tag_("fc2");
      gemm_plain(st, EPI_RESID, tb.h2, mlp * D, w.fc2_w3, mlp * D, T_, D, mlp * D, e,
                 GemmSplit{tb.h_half, (size_t)mlp * D * D});
    }
    if (cat_out != nullptr)   // the encoder's mean || logvar concat feeds quant_conv: cat_out = hi half of cat2_
      split_rows(st, x_out, D, T_, D, cat_out + cat_col0, cat_out + cat_half_ + cat_col0, cat_ld, false);
  } else {
    {
      EpiParams e{};
      e.bias = w.fc1_b;
      e.out_bf16 = tb.h; e.ldo = mlp * D;
      static const bool exact = [] { const char* v = getenv("CRA5_GELU_EXACT"); return v != nullptr && atoi(v) != 0; }();
      e.gelu_fast = exact ? 0 : 1;
      TagScope tag_("fc1");
      gemm_plain(st, EPI_GELU_BF16, tb.a, D, w.fc1_w, D, T_, mlp * D, D, e);
    }
    {
      EpiParams e{};
      e.bias = w.fc2_b;
      e.resid = x_out; e.out_f32 = x_out; e.ldo = D;
      e.out_bf16 = cat_out; e.bf16_col0 = cat_col0; e.ld_bf16 = cat_ld;
      TagScope tag_("fc2");
      gemm_plain(st, EPI_RESID, tb.h, mlp * D, w.fc2_w, mlp * D, T_, D, mlp * D, e);
    }
  }
  if (main) {   // diagnostics: operand buffers of the last trunk block that ran
    taps_["blk.q"] = TensorRef{tb.q, CRA5_DT_BF16, (int64_t)rows * D};
    taps_["blk.k"] = TensorRef{tb.k, CRA5_DT_BF16, (int64_t)rows * D};
    taps_["blk.vt"] = TensorRef{tb.vt, CRA5_DT_BF16, (int64_t)rows * D};
    taps_["blk.o"] = TensorRef{precise ? tb.o2 : tb.o, CRA5_DT_BF16, (int64_t)rows * D};
    taps_["blk.a"] = TensorRef{tb.a, CRA5_DT_BF16, (int64_t)T_ * D};
    taps_["blk.h"] = TensorRef{tb.h, CRA5_DT_BF16, (int64_t)T_ * mlp * D};
  }
}

static void window_of(const cra5_config& c, int abs_block, int* wh, int* ww) {
  // vit_nlc.py:402-407 / :614-619: global every `interval`-th block, else window_size[min(i % interval, n-1)]
  if ((abs_block + 1) % c.interval == 0) { *wh = 0; *ww = 0; return; }
  const int which = std::min(abs_block % c.interval, c.n_windows - 1);
  *wh = c.window_h[which];
  *ww = c.window_w[which];
}

// ViT_Encoder.forward + quant_conv + mode(): x -> y  (vit_nlc.py:458-486, vaeformer.py:272-283)
void Model::encode_to_latent(const float* x, float* y, const float* mean, const float* std_, int B, cudaStream_t st) {
  finalize();
  NvtxRange nvtx_("encode_to_latent");
  const cra5_config& c = cfg_;
  const int D = c.dim, CS = c.in_chans * c.patch_w;
  CRA5_CHECK((mean == nullptr) == (std_ == nullptr), ERR_INVALID, "mean and std must be given together");
  CRA5_CHECK(B >= 1 && B <= Bm, ERR_INVALID, "batch larger than the model's max_batch");
  const int TB = B * T;
  const bool p_embed = precision_ >= 2;
  __nv_bfloat16* patches = p_embed ? patches2_ : patches_;
  const size_t frame_in = (size_t)c.in_chans * c.img_h * c.img_w, frame_p = (size_t)c.img_h * Wg * cs_pad;
  for (int b = 0; b < B; ++b)
    frame_to_patches(st, x + b * frame_in, patches + b * frame_p, mean, std_, c.in_chans, c.img_h, c.img_w, Wg, c.patch_w,
                     cs_pad, p_embed ? patches2_ + patches_half_ + b * frame_p : nullptr);
  {
    // implicit-GEMM patch embedding: K ordered (kernel row r, channel, column s), zero padded per r to kpr*64
    const int Kp = c.patch_h * kpr * GEMM_BK;
    const int bn = gemm_pick_bn(D);
    CUtensorMap tmA, tmB;
    if (p_embed) {   // split-bf16: one more (outermost) coordinate selects the hi / lo half of both operands
      const __nv_bfloat16* Wpe = need_x3("g_a.patch_embed.proj.weight", (int64_t)D * Kp);
      uint64_t dims[4] = {(uint64_t)CS, (uint64_t)Wg, (uint64_t)B * c.img_h, 2};
      uint64_t strides[3] = {(uint64_t)cs_pad * 2, (uint64_t)Wg * cs_pad * 2, (uint64_t)patches_half_ * 2};
      uint32_t box[4] = {GEMM_BK, (uint32_t)box_rows, 1, 1};
      tmA = make_tmap_bf16(patches, 4, dims, strides, box, true);
      uint64_t bdims[3] = {(uint64_t)Kp, (uint64_t)D, 2};
      uint64_t bstrides[2] = {(uint64_t)Kp * 2, (uint64_t)D * Kp * 2};
      uint32_t bbox[3] = {GEMM_BK, (uint32_t)bn, 1};
      tmB = make_tmap_bf16(Wpe, 3, bdims, bstrides, bbox, true);
    } else {
      const __nv_bfloat16* Wpe = (const __nv_bfloat16*)need("g_a.patch_embed.proj.weight", CRA5_DT_BF16, (int64_t)D * Kp);
      uint64_t dims[3] = {(uint64_t)CS, (uint64_t)Wg, (uint64_t)B * c.img_h};
      uint64_t strides[2] = {(uint64_t)cs_pad * 2, (uint64_t)Wg * cs_pad * 2};
      uint32_t box[3] = {GEMM_BK, (uint32_t)box_rows, 1};
      tmA = make_tmap_bf16(patches, 3, dims, strides, box, true);
      tmB = make_tmap_bf16_2d(Wpe, (uint64_t)Kp, (uint64_t)D, (uint64_t)Kp * 2, GEMM_BK, bn);
    }
    GemmShape shp{};
    shp.M = TB; shp.N = D; shp.K = Kp; shp.a_mode = A_PATCH;
    shp.pe_kpr = kpr; shp.pe_box_rows = box_rows; shp.pe_Wp = Wg; shp.pe_sh = c.stride_h;
    shp.pe_Hg = Hg; shp.pe_img_h = c.img_h;   // frames of a batch are img_h pixel rows apart, not stride_h * Hg
    shp.a_split = shp.b_split = p_embed ? 1 : 0;
    EpiParams e{};
    e.bias = (const float*)need("g_a.patch_embed.proj.bias", CRA5_DT_F32, D);
    e.add = (const float*)need("g_a.pos_embed", CRA5_DT_F32, (int64_t)T * D);
    e.lda = D; e.add_period = T;
    e.out_f32 = main_.x; e.ldo = D;
    TagScope tag_("patch_embed");
    launch_gemm(st, bn, EPI_F32, tmA, tmB, shp, e);
  }
  taps_["tokens"] = TensorRef{main_.x, CRA5_DT_F32, (int64_t)TB * D};
  const int n = (int)ga_.size();
  int wh, ww;
  static const int debug_stop = [] { const char* e = getenv("CRA5_DEBUG_STOP_AFTER"); return e ? atoi(e) : -1; }();
  for (int i = 0; i < n - 2; ++i) {
    if (debug_stop >= 0 && i >= debug_stop) return;   // diagnostics: leave the residual stream of block i-1 in "tokens"
    window_of(c, i, &wh, &ww);
    run_block(st, ga_[i], main_, main_.x, main_.x, TB, D, c.num_heads, c.mlp_ratio, wh, ww, nullptr, 0, 0, precision_ >= 2,
              B, true);
  }
  // the last two blocks share their input and window setting; outputs are concatenated (vit_nlc.py:467-472)
  const bool tail = precision_ >= 1;
  __nv_bfloat16* cat = tail ? cat2_ : cat_;
  window_of(c, n - 2, &wh, &ww);
  run_block(st, ga_[n - 2], main_, main_.x, x1_, TB, D, c.num_heads, c.mlp_ratio, wh, ww, cat, 0, 2 * D, tail, B, true);
  run_block(st, ga_[n - 1], main_, main_.x, x2_, TB, D, c.num_heads, c.mlp_ratio, wh, ww, cat, D, 2 * D, tail, B, true);
  {
    // quant_conv 1x1, only the `mean` half of the moments is ever used (distributions.py:32,71-72)
    const int lat = c.latent_chans;
    EpiParams e{};
    e.bias = (const float*)need("quant_conv.bias", CRA5_DT_F32, lat);
    e.out_f32 = y; e.ldo = T;
    e.fr_rows = T; e.fr_stride = (size_t)lat * T;   // y is [frame][latent][T]
    TagScope tag_("quant_conv");
    if (tail)
      gemm_plain(st, EPI_T_F32, cat, 2 * D, need_x3("quant_conv.weight", (int64_t)lat * 2 * D), 2 * D, TB, lat, 2 * D, e,
                 GemmSplit{cat_half_, (size_t)lat * 2 * D});
    else
      gemm_plain(st, EPI_T_F32, cat_, 2 * D, (const __nv_bfloat16*)need("quant_conv.weight", CRA5_DT_BF16, (int64_t)lat * 2 * D),
                 2 * D, TB, lat, 2 * D, e);
  }
  taps_["y"] = TensorRef{y, CRA5_DT_F32, (int64_t)B * c.latent_chans * T};
}

// HyperpriorEncoder: y (latent, Hg, Wg) -> z_ (zc, Hh, Wh)   (vit_nlc.py:488-551 via :477-486)
void Model::run_h_a(cudaStream_t st, const float* y, int B) {
  finalize();
  NvtxRange nvtx_("h_a");
  const cra5_config& c = cfg_;
  const int Dh = c.hyper_dim, lat = c.latent_chans, zc = c.z_chans;
  const int Kc = lat * c.hyper_patch_h * c.hyper_patch_w;
  const int hidden = std::max(1, (int)sqrt((double)(Dh / zc))) * zc;
  const bool pr = precision_ >= 1;   // the hyperprior decides the scale indexes and means: split-bf16 from level 1 on
  const int TB = B * Th;
  __nv_bfloat16* ah = pr ? ah2_ : ah_;
  for (int b = 0; b < B; ++b)
    im2col_latent(st, y + (size_t)b * lat * T, ah + (size_t)b * Th * Kc, lat, Hg, Wg, c.hyper_patch_h, c.hyper_patch_w, Kc,
                  pr ? ah2_ + ah_half_ + (size_t)b * Th * Kc : nullptr);
  {
    EpiParams e{};
    e.bias = (const float*)need("h_a.patch_embed.proj.bias", CRA5_DT_F32, Dh);
    e.add = (const float*)need("h_a.pos_embed", CRA5_DT_F32, (int64_t)Th * Dh);
    e.lda = Dh; e.add_period = Th;
    e.out_f32 = hyper_.x; e.ldo = Dh;
    TagScope tag_("h_a.patch_embed");
    if (pr)
      gemm_plain(st, EPI_F32, ah, Kc, need_x3("h_a.patch_embed.proj.weight", (int64_t)Dh * Kc), Kc, TB, Dh, Kc, e,
                 GemmSplit{ah_half_, (size_t)Dh * Kc});
    else
      gemm_plain(st, EPI_F32, ah_, Kc, (const __nv_bfloat16*)need("h_a.patch_embed.proj.weight", CRA5_DT_BF16, (int64_t)Dh * Kc),
                 Kc, TB, Dh, Kc, e);
  }
  for (size_t i = 0; i < ha_.size(); ++i)
    run_block(st, ha_[i], hyper_, hyper_.x, hyper_.x, TB, Dh, c.hyper_heads, c.mlp_ratio, 0, 0, nullptr, 0, 0, pr, B, false);
  // quan_mlp (vit_nlc.py:544-546): fc1 -> GELU -> fc2, no norm in front
  if (pr) {
    split_rows(st, hyper_.x, Dh, TB, Dh, hyper_.a2, hyper_.a2 + hyper_.a_half, Dh, false);
    {
      EpiParams e{};
      e.bias = (const float*)need("h_a.quan_mlp.fc1.bias", CRA5_DT_F32, hidden);
      e.out_f32 = hyper_.f32; e.ldo = hidden;
      TagScope tag_("h_a.quan_fc1");
      gemm_plain(st, EPI_F32, hyper_.a2, Dh, need_x3("h_a.quan_mlp.fc1.weight", (int64_t)hidden * Dh), Dh, TB, hidden, Dh, e,
                 GemmSplit{hyper_.a_half, (size_t)hidden * Dh});
    }
    split_rows(st, hyper_.f32, hidden, TB, hidden, hyper_.h2, hyper_.h2 + hyper_.h_half, hidden, true);
    {
      EpiParams e{};
      e.bias = (const float*)need("h_a.quan_mlp.fc2.bias", CRA5_DT_F32, zc);
      e.out_f32 = z_; e.ldo = Th;
      e.fr_rows = Th; e.fr_stride = (size_t)zc * Th;
      TagScope tag_("h_a.quan_fc2");
      gemm_plain(st, EPI_T_F32, hyper_.h2, hidden, need_x3("h_a.quan_mlp.fc2.weight", (int64_t)zc * hidden), hidden, TB, zc,
                 hidden, e, GemmSplit{hyper_.h_half, (size_t)zc * hidden});
    }
  } else {
    cast_bf16(st, hyper_.x, hyper_.a, (size_t)TB * Dh);
    {
      EpiParams e{};
      e.bias = (const float*)need("h_a.quan_mlp.fc1.bias", CRA5_DT_F32, hidden);
      e.out_bf16 = hyper_.h; e.ldo = hidden;
      TagScope tag_("h_a.quan_fc1");
      gemm_plain(st, EPI_GELU_BF16, hyper_.a, Dh, (const __nv_bfloat16*)need("h_a.quan_mlp.fc1.weight", CRA5_DT_BF16, (int64_t)hidden * Dh),
                 Dh, TB, hidden, Dh, e);
    }
    {
      EpiParams e{};
      e.bias = (const float*)need("h_a.quan_mlp.fc2.bias", CRA5_DT_F32, zc);
      e.out_f32 = z_; e.ldo = Th;
      e.fr_rows = Th; e.fr_stride = (size_t)zc * Th;
      TagScope tag_("h_a.quan_fc2");
      gemm_plain(st, EPI_T_F32, hyper_.h, hidden, (const __nv_bfloat16*)need("h_a.quan_mlp.fc2.weight", CRA5_DT_BF16, (int64_t)zc * hidden),
                 hidden, TB, zc, hidden, e);
    }
  }
  taps_["z"] = TensorRef{z_, CRA5_DT_F32, (int64_t)B * zc * Th};
}

// HyperpriorDecoder: z_hat (zc, Hh, Wh) -> params_ = [sigma (latent) | mu (latent)] x (Hg, Wg)  (vit_nlc.py:696-748)
void Model::run_h_s(cudaStream_t st, const float* z_hat, int B) {
  finalize();
  NvtxRange nvtx_("h_s");
  const cra5_config& c = cfg_;
  const int Dh = c.hyper_dim, lat = c.latent_chans, zc = c.z_chans;
  const int hidden = std::max(1, (int)sqrt((double)(Dh / zc))) * zc;
  const bool pr = precision_ >= 1;
  const int Nf = 2 * lat * c.hyper_patch_h * c.hyper_patch_w;
  const int TB = B * Th;
  if (pr) {
    for (int b = 0; b < B; ++b)
      transpose_cast(st, z_hat + (size_t)b * zc * Th, ztok2_ + (size_t)b * Th * zc, zc, Th, zc,
                     ztok2_ + ztok_half_ + (size_t)b * Th * zc);
    {
      EpiParams e{};
      e.bias = (const float*)need("h_s.post_quan_mlp.fc1.bias", CRA5_DT_F32, hidden);
      e.out_f32 = hyper_.f32; e.ldo = hidden;
      TagScope tag_("h_s.post_fc1");
      gemm_plain(st, EPI_F32, ztok2_, zc, need_x3("h_s.post_quan_mlp.fc1.weight", (int64_t)hidden * zc), zc, TB, hidden, zc, e,
                 GemmSplit{ztok_half_, (size_t)hidden * zc});
    }
    split_rows(st, hyper_.f32, hidden, TB, hidden, hyper_.h2, hyper_.h2 + hyper_.h_half, hidden, true);
    {
      EpiParams e{};
      e.bias = (const float*)need("h_s.post_quan_mlp.fc2.bias", CRA5_DT_F32, Dh);
      e.out_f32 = hyper_.x; e.ldo = Dh;
      TagScope tag_("h_s.post_fc2");
      gemm_plain(st, EPI_F32, hyper_.h2, hidden, need_x3("h_s.post_quan_mlp.fc2.weight", (int64_t)Dh * hidden), hidden, TB, Dh,
                 hidden, e, GemmSplit{hyper_.h_half, (size_t)Dh * hidden});
    }
  } else {
    for (int b = 0; b < B; ++b)
      transpose_cast(st, z_hat + (size_t)b * zc * Th, ztok_ + (size_t)b * Th * zc, zc, Th, zc);
    {
      EpiParams e{};
      e.bias = (const float*)need("h_s.post_quan_mlp.fc1.bias", CRA5_DT_F32, hidden);
      e.out_bf16 = hyper_.h; e.ldo = hidden;
      TagScope tag_("h_s.post_fc1");
      gemm_plain(st, EPI_GELU_BF16, ztok_, zc, (const __nv_bfloat16*)need("h_s.post_quan_mlp.fc1.weight", CRA5_DT_BF16, (int64_t)hidden * zc),
                 zc, TB, hidden, zc, e);
    }
    {
      EpiParams e{};
      e.bias = (const float*)need("h_s.post_quan_mlp.fc2.bias", CRA5_DT_F32, Dh);
      e.out_f32 = hyper_.x; e.ldo = Dh;
      TagScope tag_("h_s.post_fc2");
      gemm_plain(st, EPI_F32, hyper_.h, hidden, (const __nv_bfloat16*)need("h_s.post_quan_mlp.fc2.weight", CRA5_DT_BF16, (int64_t)Dh * hidden),
                 hidden, TB, Dh, hidden, e);
    }
  }
  for (size_t i = 0; i < hs_.size(); ++i)
    run_block(st, hs_[i], hyper_, hyper_.x, hyper_.x, TB, Dh, c.hyper_heads, c.mlp_ratio, 0, 0, nullptr, 0, 0, pr, B, false);
  __nv_bfloat16* fin = pr ? hyper_.a2 : hyper_.a;
  layernorm_bf16(st, hyper_.x, (const float*)need("h_s.norm.weight", CRA5_DT_F32, Dh),
                 (const float*)need("h_s.norm.bias", CRA5_DT_F32, Dh), c.ln_eps, fin, TB, Dh, WinMap{},
                 pr ? hyper_.a2 + hyper_.a_half : nullptr);
  {
    // Linear(Dh, 2*latent*p1*p2, bias=False) + rearrange '(p1 p2 c)' (vit_nlc.py:741, 671-680)
    EpiParams e{};
    e.out_f32 = params_; e.ldo = T;
    e.ps_P1 = c.hyper_patch_h; e.ps_P2 = c.hyper_patch_w; e.ps_C = 2 * lat; e.ps_Wh = Wh;
    e.fr_rows = Th; e.fr_stride = (size_t)2 * lat * T;   // params_ is [frame][sigma (latent) | mu (latent)][T]
    TagScope tag_("h_s.final");
    if (pr)
      gemm_plain(st, EPI_PIXSHUF, fin, Dh, need_x3("h_s.final.weight", (int64_t)Nf * Dh), Dh, TB, Nf, Dh, e,
                 GemmSplit{hyper_.a_half, (size_t)Nf * Dh});
    else
      gemm_plain(st, EPI_PIXSHUF, hyper_.a, Dh, (const __nv_bfloat16*)need("h_s.final.weight", CRA5_DT_BF16, (int64_t)Nf * Dh), Dh, TB,
                 Nf, Dh, e);
  }
  // (frame 0 of the batch; the parity tests read these after single-frame calls)
  taps_["scales"] = TensorRef{params_, CRA5_DT_F32, (int64_t)lat * T};
  taps_["means"] = TensorRef{params_ + (size_t)lat * T, CRA5_DT_F32, (int64_t)lat * T};
  taps_["params"] = TensorRef{params_, CRA5_DT_F32, (int64_t)B * 2 * lat * T};
}

static const float* scale_table_of(const Model* m, const std::map<std::string, TensorRef>& t, int rows) {
  auto it = t.find("gaussian_conditional.scale_table");
  CRA5_CHECK(it != t.end(), ERR_STATE, "Uninitialized CDFs. Run update() first");
  CRA5_CHECK(it->second.dtype == CRA5_DT_F32 && it->second.numel == rows, ERR_INVALID,
             "scale_table does not match the GaussianConditional CDF rows");
  (void)m;
  return (const float*)it->second.ptr;
}

constexpr float SCALE_BOUND = 0.11f;  // entropy_models.py:560

// tail of VAEformer.encode_latent(type='quantized') (vaeformer.py:284-290), eval mode = "dequantize"
void Model::latent_quantized(const float* y, float* y_hat, cudaStream_t st) {
  const cra5_config& c = cfg_;
  CRA5_CHECK(gc_.ready(), ERR_STATE, "Uninitialized CDFs. Run update() first");
  const float* med = (const float*)need("entropy_bottleneck.medians", CRA5_DT_F32, c.z_chans);
  run_h_a(st, y, 1);
  eb_quantize(st, z_, med, Th, c.z_chans, nullptr, zhat_, (size_t)c.z_chans * Th);
  run_h_s(st, zhat_, 1);
  gc_quantize_index(st, y, nullptr, params_ + (size_t)c.latent_chans * T, nullptr, 1, SCALE_BOUND, nullptr, nullptr, y_hat,
                    (size_t)c.latent_chans * T);
}

// eval-mode forward up to the latent (vaeformer.py:314-319): z_hat / z likelihoods from the EntropyBottleneck, then
// y_hat / y likelihoods from the GaussianConditional fed by h_s(z_hat)
void Model::latent_likelihoods(const float* y, float* y_hat, float* y_lik, float* z_lik, cudaStream_t st) {
  const cra5_config& c = cfg_;
  constexpr float LIK_BOUND = 1e-9f;  // entropy_models.py:111
  const float* med = (const float*)need("entropy_bottleneck.medians", CRA5_DT_F32, c.z_chans);
  run_h_a(st, y, 1);
  eb_quantize(st, z_, med, Th, c.z_chans, nullptr, zhat_, (size_t)c.z_chans * Th);
  if (z_lik != nullptr)
    eb_likelihood(st, zhat_, (const float*)need("entropy_bottleneck.packed", CRA5_DT_F32, (int64_t)c.z_chans * 58), Th,
                  LIK_BOUND, z_lik, (size_t)c.z_chans * Th);
  run_h_s(st, zhat_, 1);
  const size_t n = (size_t)c.latent_chans * T;
  if (y_lik != nullptr)
    gc_likelihood(st, y, params_, params_ + n, SCALE_BOUND, LIK_BOUND, y_hat, y_lik, n);
  else if (y_hat != nullptr)
    gc_quantize_index(st, y, nullptr, params_ + n, nullptr, 1, SCALE_BOUND, nullptr, nullptr, y_hat, n);
  taps_["z_hat"] = TensorRef{zhat_, CRA5_DT_F32, (int64_t)c.z_chans * Th};
}

// VAEformer.compress_from_latent (vaeformer.py:334-348) for a batch of B frames: every kernel -- hyperprior GEMMs,
// quantise + scale index, rANS encode, length scan, compaction, container write -- is launched ONCE for the batch; the
// reference loops over the batch items inside EntropyModel.compress (entropy_models.py:263-272).
void Model::latent_to_bin(const float* y, int B, const uint8_t** y_bytes, size_t* y_len, const uint8_t** z_bytes,
                          size_t* z_len, cudaStream_t st) {
  NvtxRange nvtx_("latent_to_bin");
  const cra5_config& c = cfg_;
  CRA5_CHECK(eb_.ready() && gc_.ready(), ERR_STATE, "Uninitialized CDFs. Run update() first");
  CRA5_CHECK(B >= 1 && B <= Bm, ERR_INVALID, "batch larger than the model's max_batch");
  const int lat = c.latent_chans, zc = c.z_chans;
  const size_t n = (size_t)lat * T;
  const float* med = (const float*)need("entropy_bottleneck.medians", CRA5_DT_F32, zc);
  const float* table = scale_table_of(this, tensors_, gc_.rows);
  run_h_a(st, y, B);
  // z: symbols + the z_hat the decoder will see (the reference decodes its own z string, vaeformer.py:340)
  eb_quantize(st, z_, med, Th, zc, zsym_, zhat_, (size_t)B * zc * Th);
  taps_["z_symbols"] = TensorRef{zsym_, CRA5_DT_I32, (int64_t)B * zc * Th};
  taps_["z_hat"] = TensorRef{zhat_, CRA5_DT_F32, (int64_t)B * zc * Th};
  run_h_s(st, zhat_, B);
  {
    TagScope tag_("encode");   // symbols + indexes: the 17 B/element launch the HBM roofline target is stated for
    gc_quantize_index(st, y, params_, params_ + n, table, gc_.rows, SCALE_BOUND, ysym_, yidx_, nullptr, n, B, 2 * n);
  }
  taps_["y_symbols"] = TensorRef{ysym_, CRA5_DT_I32, (int64_t)(B * n)};
  taps_["y_indexes"] = TensorRef{yidx_, CRA5_DT_U8, (int64_t)(B * n)};
  if (spc_y_ > 0 && spc_z_ > 0) {
    // all containers are written into host memory by the kernels themselves: one synchronisation for the batch
    coder_->encode_begin(st, 0, zsym_, nullptr, eb_, zc, Th, spc_z_, host_z_, host_z_cap_, B, frame_cap_z_);
    coder_->encode_begin(st, 1, ysym_, yidx_, gc_, lat, T, spc_y_, host_y_, host_y_cap_, B, frame_cap_y_);
    CRA5_CUDA(cudaStreamSynchronize(st));
    coder_->encode_end(st, 0, zc, Th, spc_z_, host_z_, host_z_cap_, B, frame_cap_z_, z_len);
    coder_->encode_end(st, 1, lat, T, spc_y_, host_y_, host_y_cap_, B, frame_cap_y_, y_len);
    for (int b = 0; b < B; ++b) {
      y_bytes[b] = host_y_ + (size_t)b * frame_cap_y_;
      z_bytes[b] = host_z_ + (size_t)b * frame_cap_z_;
    }
  } else {  // reference-format single streams (interop path): one sequential stream per tensor and frame
    for (int b = 0; b < B; ++b) {
      z_len[b] = coder_->encode(st, zsym_ + (size_t)b * zc * Th, nullptr, eb_, zc, Th, spc_z_, host_z_ + (size_t)b * frame_cap_z_,
                                frame_cap_z_);
      y_len[b] = coder_->encode(st, ysym_ + b * n, yidx_ + b * n, gc_, lat, T, spc_y_, host_y_ + (size_t)b * frame_cap_y_,
                                frame_cap_y_);
      y_bytes[b] = host_y_ + (size_t)b * frame_cap_y_;
      z_bytes[b] = host_z_ + (size_t)b * frame_cap_z_;
    }
  }
}

// VAEformer.decompress(return_format='latent') (vaeformer.py:378-391) for a batch of B frames
void Model::bin_to_latent(const uint8_t* const* y_bytes, const size_t* y_len, const uint8_t* const* z_bytes,
                          const size_t* z_len, int B, int z_h, int z_w, float* y_hat, cudaStream_t st) {
  NvtxRange nvtx_("bin_to_latent");
  const cra5_config& c = cfg_;
  CRA5_CHECK(eb_.ready() && gc_.ready(), ERR_STATE, "Uninitialized CDFs. Run update() first");
  CRA5_CHECK(z_h == Hh && z_w == Wh, ERR_INVALID, "z shape does not match the model geometry");
  CRA5_CHECK(B >= 1 && B <= Bm, ERR_INVALID, "batch larger than the model's max_batch");
  const int lat = c.latent_chans, zc = c.z_chans;
  const size_t n = (size_t)lat * T;
  const float* med = (const float*)need("entropy_bottleneck.medians", CRA5_DT_F32, zc);
  const float* table = scale_table_of(this, tensors_, gc_.rows);
  // The CR5B containers of the whole batch are staged up front in disjoint parts of the pinned buffer and the chain --
  // z decode, h_s, scale indexes, y decode, each ONE launch for the batch -- is enqueued without a host synchronisation
  // in between; one synchronisation at the end fetches the error word of both decodes.
  bool chunked = true;
  size_t z_total = 0, y_total = 0;
  for (int b = 0; b < B; ++b) {
    chunked = chunked && y_bytes[b] != nullptr && z_bytes[b] != nullptr && y_len[b] >= 4 && z_len[b] >= 4 &&
              memcmp(y_bytes[b], "CR5B", 4) == 0 && memcmp(z_bytes[b], "CR5B", 4) == 0;
    z_total += z_len[b];
    y_total += y_len[b];
  }
  const size_t y_off = (z_total + 255) & ~size_t(255);
  if (chunked && y_off + y_total <= coder_->stage_capacity()) {
    bool any = false;
    try {
      any = coder_->decode_cr5b(st, z_bytes, z_len, B, nullptr, eb_, zc, Th, zsym_, nullptr, med, zhat_, 0, true, 0);
      taps_["z_symbols"] = TensorRef{zsym_, CRA5_DT_I32, (int64_t)B * zc * Th};
      taps_["z_hat"] = TensorRef{zhat_, CRA5_DT_F32, (int64_t)B * zc * Th};
      run_h_s(st, zhat_, B);
      {
        TagScope tag_("decode_idx");   // decode side: indexes only (4 B in, 1 B out per element)
        gc_quantize_index(st, nullptr, params_, nullptr, table, gc_.rows, SCALE_BOUND, nullptr, yidx_, nullptr, n, B, 2 * n);
      }
      // frame f's means start at params_ + n + f * 2n while its symbols start at f * n: n extra elements per frame
      any = coder_->decode_cr5b(st, y_bytes, y_len, B, yidx_, gc_, lat, T, ysym_, params_ + n, nullptr, y_hat, y_off, false,
                                n) || any;
    } catch (...) {
      cudaStreamSynchronize(st);   // what was enqueued may still read the staging buffer
      coder_->reset_error(st);
      throw;
    }
    if (any) coder_->decode_finish(st);
    taps_["y_symbols"] = TensorRef{ysym_, CRA5_DT_I32, (int64_t)(B * n)};
    taps_["y_indexes"] = TensorRef{yidx_, CRA5_DT_U8, (int64_t)(B * n)};
    return;
  }
  // general path: reference-format streams (one GPU thread per tensor) or containers larger than the staging buffer;
  // one frame per call, four host synchronisations
  CRA5_CHECK(B == 1, ERR_INVALID, "reference-format streams are decoded one frame per call");
  coder_->decode(st, z_bytes[0], z_len[0], nullptr, eb_, zc, Th, zsym_, nullptr, med, zhat_);
  taps_["z_symbols"] = TensorRef{zsym_, CRA5_DT_I32, (int64_t)zc * Th};
  taps_["z_hat"] = TensorRef{zhat_, CRA5_DT_F32, (int64_t)zc * Th};
  run_h_s(st, zhat_, 1);
  gc_quantize_index(st, nullptr, params_, nullptr, table, gc_.rows, SCALE_BOUND, nullptr, yidx_, nullptr, n);
  coder_->decode(st, y_bytes[0], y_len[0], yidx_, gc_, lat, T, ysym_, params_ + n, nullptr, y_hat);
  taps_["y_symbols"] = TensorRef{ysym_, CRA5_DT_I32, (int64_t)n};
  taps_["y_indexes"] = TensorRef{yidx_, CRA5_DT_U8, (int64_t)n};
}

// VAEformer.decode_latent (vaeformer.py:294-300): post_quant_conv + ViT_Decoder.forward (vit_nlc.py:682-693)
void Model::latent_to_reconstruction(const float* y_hat, float* x_hat, int B, cudaStream_t st, const float* mean,
                                      const float* std_) {
  finalize();
  NvtxRange nvtx_("latent_to_reconstruction");
  CRA5_CHECK((mean == nullptr) == (std_ == nullptr), ERR_INVALID, "mean and std must be given together");
  const cra5_config& c = cfg_;
  const int D = c.dim, lat = c.latent_chans, CS = c.in_chans * c.patch_w;
  CRA5_CHECK(B >= 1 && B <= Bm, ERR_INVALID, "batch larger than the model's max_batch");
  const int TB = B * T;
  const bool pr = precision_ >= 3;   // the decoder does not influence the bitstream; split-bf16 only at the top level
  __nv_bfloat16* ytok = pr ? ytok2_ : ytok_;
  for (int b = 0; b < B; ++b)
    transpose_cast(st, y_hat + (size_t)b * lat * T, ytok + (size_t)b * T * lat, lat, T, lat,
                   pr ? ytok2_ + ytok_half_ + (size_t)b * T * lat : nullptr);
  {
    EpiParams e{};
    e.bias = (const float*)need("post_quant_conv.bias", CRA5_DT_F32, D);
    e.out_f32 = main_.x; e.ldo = D;
    TagScope tag_("post_quant_conv");
    if (pr)
      gemm_plain(st, EPI_F32, ytok2_, lat, need_x3("post_quant_conv.weight", (int64_t)D * lat), lat, TB, D, lat, e,
                 GemmSplit{ytok_half_, (size_t)D * lat});
    else
      gemm_plain(st, EPI_F32, ytok_, lat, (const __nv_bfloat16*)need("post_quant_conv.weight", CRA5_DT_BF16, (int64_t)D * lat), lat,
                 TB, D, lat, e);
  }
  const int n = (int)gs_.size();
  for (int i = 0; i < n; ++i) {
    int wh, ww;
    window_of(c, c.depth / 2 + i, &wh, &ww);
    run_block(st, gs_[i], main_, main_.x, main_.x, TB, D, c.num_heads, c.mlp_ratio, wh, ww, nullptr, 0, 0, pr, B, true);
  }
  const float* ng = (const float*)need("g_s.norm.weight", CRA5_DT_F32, D);
  const float* nbias = (const float*)need("g_s.norm.bias", CRA5_DT_F32, D);
  const bool conv_head = (c.img_h == 721 && c.img_w == 1440);  // vit_nlc.py:628
  const size_t frame_out = (size_t)c.in_chans * c.img_h * c.img_w;
  if (conv_head) {
    // ConvTranspose2d(k=(ph,pw), s=(sh,pw)) as two GEMMs with a scatter epilogue, no atomics:
    //  class A: kernel rows r in [nB, sh) touch exactly one patch row  -> K = D
    //  class B: output rows sh*i' + r', r' < nB receive kernel row r' of patch i' and row r'+sh of patch i'-1 -> K = 2D
    // Operand: the final LayerNorm, laid out [frame][M2 = (Hg+1)*Wg rows][D] with the last Wg rows of every frame zero
    // (never written), so that class B's "patch row i-1" of the first row block of a frame and "patch row Hg" of its
    // last read zeros -- within a batch exactly as at the edges of a single frame.
    __nv_bfloat16* fin = pr ? fin2_ : fin_;
    const size_t fin_half = pr ? fin_half_ : 0;
    for (int b = 0; b < B; ++b)
      layernorm_bf16(st, main_.x + (size_t)b * T * D, ng, nbias, c.ln_eps, fin + (size_t)b * M2 * D, T, D, WinMap{},
                     pr ? fin2_ + fin_half_ + (size_t)b * M2 * D : nullptr);
    const int MB = B * M2;
    EpiParams e{};
    e.out_f32 = x_hat;
    e.ct_CS = CS; e.ct_pw = c.patch_w; e.ct_sh = c.stride_h; e.ct_Wp = Wg; e.ct_Himg = c.img_h; e.ct_Wimg = c.img_w;
    e.ct_mean = mean; e.ct_std = std_;   // de-normalisation fused into the un-patchify store
    e.fr_rows = M2; e.fr_stride = frame_out;
    // Both kernel-row classes run on the channel-grouped column order when the patch width is the shipped one
    // (gemm_tc.cuh: epilogue_convt_grouped): per kernel row, groups of 32 columns = 3 whole channels x pw + 2 zero-weight
    // pad columns; the weights are packed that way at upload (cra5_b200/vaeformer.py::_upload)
    const bool grouped = c.patch_w == CT_PW;
    const int cpg = 30 / CT_PW, groups = (c.in_chans + cpg - 1) / cpg;
    const int CSg = grouped ? 32 * groups : CS;     // columns per kernel row
    if (grouped) {
      CRA5_CHECK((int64_t)Hg * c.stride_h * c.img_w + c.img_w <= (1 << 20) && B < 4096, ERR_INVALID,
                 "unsupported geometry for the grouped un-patchify epilogue");
      e.ct_cpg = cpg; e.ct_C = c.in_chans;
    }
    e.ct_CS = CSg;
    if (nA > 0) {
      e.ct_r0 = nB;
      TagScope tag_("convT_A");
      if (pr)
        gemm_plain(st, EPI_CONVT, fin, D, need_x3("g_s.final.A", (int64_t)nA * CSg * D), D, MB, nA * CSg, D, e,
                   GemmSplit{fin_half, (size_t)nA * CSg * D});
      else
        gemm_plain(st, EPI_CONVT, fin, D, (const __nv_bfloat16*)need("g_s.final.A", CRA5_DT_BF16, (int64_t)nA * CSg * D), D, MB,
                   nA * CSg, D, e);
    }
    if (nB > 0) {
      e.ct_r0 = 0;
      const int N2 = nB * CSg, K2 = 2 * D;
      CRA5_CHECK(D % GEMM_BK == 0, ERR_INVALID, "unsupported geometry: width must be a multiple of 64 for the conv head");
      const int bn = gemm_pick_bn(N2);
      CUtensorMap tmA, tmB;
      if (pr) {
        const __nv_bfloat16* Wb = need_x3("g_s.final.B", (int64_t)N2 * K2);
        uint64_t ad[3] = {(uint64_t)D, (uint64_t)MB, 2}, as[2] = {(uint64_t)D * 2, (uint64_t)fin_half * 2};
        uint32_t ab[3] = {GEMM_BK, GEMM_BM, 1};
        tmA = make_tmap_bf16(fin, 3, ad, as, ab, true);
        uint64_t bd[3] = {(uint64_t)K2, (uint64_t)N2, 2}, bs[2] = {(uint64_t)K2 * 2, (uint64_t)N2 * K2 * 2};
        uint32_t bb[3] = {GEMM_BK, (uint32_t)bn, 1};
        tmB = make_tmap_bf16(Wb, 3, bd, bs, bb, true);
      } else {
        const __nv_bfloat16* Wb = (const __nv_bfloat16*)need("g_s.final.B", CRA5_DT_BF16, (int64_t)N2 * K2);
        tmA = make_tmap_bf16_2d(fin, (uint64_t)D, (uint64_t)MB, (uint64_t)D * 2, GEMM_BK, GEMM_BM);
        tmB = make_tmap_bf16_2d(Wb, (uint64_t)K2, (uint64_t)N2, (uint64_t)K2 * 2, GEMM_BK, bn);
      }
      GemmShape shp{};
      shp.M = MB; shp.N = N2; shp.K = K2; shp.a_mode = A_CONCAT; shp.cc_D = D; shp.cc_shift = Wg;
      shp.a_split = shp.b_split = pr ? 1 : 0;
      TagScope tag_("convT_B");
      launch_gemm(st, bn, EPI_CONVT, tmA, tmB, shp, e);
    }
  } else {
    // Linear(D, C*p1*p2, bias=False) + rearrange 'b h w (p1 p2 c) -> b c (h p1) (w p2)' (vit_nlc.py:632, 671-680)
    __nv_bfloat16* fin = pr ? main_.a2 : main_.a;
    const size_t fin_half = pr ? main_.a_half : 0;
    layernorm_bf16(st, main_.x, ng, nbias, c.ln_eps, fin, TB, D, WinMap{}, pr ? main_.a2 + main_.a_half : nullptr);
    const int Nf = c.in_chans * c.patch_h * c.patch_w;
    EpiParams e{};
    e.out_f32 = x_hat; e.ldo = (Hg * c.patch_h) * (Wg * c.patch_w);
    e.ps_P1 = c.patch_h; e.ps_P2 = c.patch_w; e.ps_C = c.in_chans; e.ps_Wh = Wg;
    e.fr_rows = T; e.fr_stride = frame_out;
    TagScope tag_("linear_head");
    if (pr)
      gemm_plain(st, EPI_PIXSHUF, fin, D, need_x3("g_s.final.weight", (int64_t)Nf * D), D, TB, Nf, D, e,
                 GemmSplit{fin_half, (size_t)Nf * D});
    else
      gemm_plain(st, EPI_PIXSHUF, main_.a, D, (const __nv_bfloat16*)need("g_s.final.weight", CRA5_DT_BF16, (int64_t)Nf * D), D, TB,
                 Nf, D, e);
    if (mean != nullptr)   // the Linear head (non-ERA5 geometries) has no fused form: one extra pass per frame
      for (int b = 0; b < B; ++b)
        affine_channels(st, x_hat + b * frame_out, x_hat + b * frame_out, mean, std_, (size_t)c.img_h * c.img_w, c.in_chans, 0);
  }
}

const void* Model::tap(const std::string& name, int64_t* numel, int* dtype) const {
  auto it = taps_.find(name);
  CRA5_CHECK(it != taps_.end(), ERR_INVALID, "no such tap: " + name);
  *numel = it->second.numel;
  *dtype = it->second.dtype;
  return it->second.ptr;
}

}  // namespace cra5
