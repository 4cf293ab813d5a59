// capi.cpp -- extern "C" surface of libcra5b200.so (declared in include/cra5_b200.h).
#include "../../include/cra5_b200.h"

#include "capi_util.h"
#include "gemm_tc.cuh"
#include "kernels.h"
#include "model.h"
#include "coder.h"

#include <memory>
#include <vector>
#include <algorithm>
#include <string.h>

namespace cra5 {
std::string& last_error_slot() {
  static thread_local std::string s;
  return s;
}
}  // namespace cra5

namespace cra5 {
void pmf_to_quantized_cdf(const float* pmf, int n, int precision, uint32_t* cdf);
}

struct cra5_model {
  std::unique_ptr<cra5::Model> impl;
};

using namespace cra5;

extern "C" {

const char* cra5_last_error(void) { return last_error_slot().c_str(); }
int cra5_abi_version(void) { return 1; }

int cra5_op_gemm(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias,
                 int epilogue, void* out, int ldo, const float* resid, void* stream) {
  return guarded([&] {
    require_sm100();
    EpiParams e{};
    e.bias = bias;
    e.ldo = ldo;
    if (epilogue == 9) {   // CRA5_EPI_GELU_BF16_TRUNK
      epilogue = EPI_GELU_BF16;
      e.gelu_fast = 1;
    }
    switch (epilogue) {
      case EPI_F32:
      case EPI_T_F32: e.out_f32 = static_cast<float*>(out); break;
      case EPI_BF16:
      case EPI_GELU_BF16: e.out_bf16 = static_cast<__nv_bfloat16*>(out); break;
      case EPI_RESID:
        e.out_f32 = static_cast<float*>(out);
        e.resid = resid;
        CRA5_CHECK(resid != nullptr, ERR_INVALID, "EPI_RESID needs resid");
        break;
      default: throw Error(ERR_INVALID, "cra5_op_gemm: unsupported epilogue");
    }
    gemm_plain(static_cast<cudaStream_t>(stream), epilogue, static_cast<const __nv_bfloat16*>(A), lda,
               static_cast<const __nv_bfloat16*>(B), ldb, M, N, K, e);
  });
}

int cra5_op_gemm_check(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias,
                       float* out, int ldo, void* stream) {
  return guarded([&] {
    gemm_simt_check(static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(A), lda,
                    static_cast<const __nv_bfloat16*>(B), ldb, bias, out, ldo, M, N, K);
  });
}

int cra5_op_attention(const void* Q, const void* K, const void* Vt, void* out, int ldo, int heads, int rows_total,
                      int seg_len, void* stream) {
  return guarded([&] {
    require_sm100();
    attention_tc(static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(Q),
                 static_cast<const __nv_bfloat16*>(K), static_cast<const __nv_bfloat16*>(Vt),
                 static_cast<__nv_bfloat16*>(out), ldo, heads, rows_total, seg_len);
  });
}

int cra5_op_attention_generic(const void* Q, const void* K, const void* Vt, void* out, int ldo, int heads, int head_dim,
                              int rows_total, int seg_len, void* stream) {
  return guarded([&] {
    require_sm100();
    attention_simt(static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(Q),
                   static_cast<const __nv_bfloat16*>(K), static_cast<const __nv_bfloat16*>(Vt),
                   static_cast<__nv_bfloat16*>(out), ldo, heads, head_dim, rows_total, seg_len);
  });
}

int cra5_pmf_to_quantized_cdf(const float* pmf, int n, int precision, uint32_t* cdf_out) {
  return guarded([&] {
    CRA5_CHECK(pmf != nullptr && cdf_out != nullptr, ERR_INVALID, "null argument");
    pmf_to_quantized_cdf(pmf, n, precision, cdf_out);
  });
}

int cra5_model_create(const cra5_config* cfg, cra5_model** out) {
  return guarded([&] {
    CRA5_CHECK(cfg != nullptr && out != nullptr, ERR_INVALID, "null argument");
    auto m = std::make_unique<cra5_model>();
    m->impl = std::make_unique<Model>(*cfg);
    *out = m.release();
  });
}

int cra5_model_destroy(cra5_model* m) {
  return guarded([&] { delete m; });
}

#define MODEL_GUARD(m) CRA5_CHECK((m) != nullptr && (m)->impl, ERR_INVALID, "null model handle")

int cra5_model_workspace_bytes(cra5_model* m, uint64_t* bytes) {
  return guarded([&] {
    MODEL_GUARD(m);
    *bytes = m->impl->workspace_bytes();
  });
}

int cra5_model_set_tensor(cra5_model* m, const char* name, const void* dev_ptr, int dtype, int64_t numel) {
  return guarded([&] {
    MODEL_GUARD(m);
    CRA5_CHECK(name != nullptr, ERR_INVALID, "null name");
    m->impl->set_tensor(name, dev_ptr, dtype, numel);
  });
}

int cra5_model_set_cdf(cra5_model* m, int which, const int32_t* cdf, const int32_t* len, const int32_t* off, int rows,
                       int cols) {
  return guarded([&] {
    MODEL_GUARD(m);
    m->impl->set_cdf(which, cdf, len, off, rows, cols);
  });
}

int cra5_model_set_coder(cra5_model* m, int spc_y, int spc_z) {
  return guarded([&] {
    MODEL_GUARD(m);
    m->impl->set_coder(spc_y, spc_z);
  });
}

int cra5_model_set_precision(cra5_model* m, int level) {
  return guarded([&] {
    MODEL_GUARD(m);
    m->impl->set_precision(level);
  });
}

int cra5_encode_to_latent_batch(cra5_model* m, const float* x, float* y, const float* mean, const float* std_, int batch,
                                void* stream) {
  return guarded([&] {
    MODEL_GUARD(m);
    CRA5_CHECK(x && y, ERR_INVALID, "null tensor");
    m->impl->encode_to_latent(x, y, mean, std_, batch, static_cast<cudaStream_t>(stream));
  });
}
int cra5_encode_to_latent(cra5_model* m, const float* x, float* y, const float* mean, const float* std_, void* stream) {
  return cra5_encode_to_latent_batch(m, x, y, mean, std_, 1, stream);
}

int cra5_latent_quantized(cra5_model* m, const float* y, float* y_hat, void* stream) {
  return guarded([&] {
    MODEL_GUARD(m);
    CRA5_CHECK(y && y_hat, ERR_INVALID, "null tensor");
    m->impl->latent_quantized(y, y_hat, static_cast<cudaStream_t>(stream));
  });
}

int cra5_latent_likelihoods(cra5_model* m, const float* y, float* y_hat, float* y_lik, float* z_lik, void* stream) {
  return guarded([&] {
    MODEL_GUARD(m);
    CRA5_CHECK(y != nullptr, ERR_INVALID, "null tensor");
    m->impl->latent_likelihoods(y, y_hat, y_lik, z_lik, static_cast<cudaStream_t>(stream));
  });
}

int cra5_latent_to_bin_batch(cra5_model* m, const float* y, int batch, const uint8_t** y_bytes, uint64_t* y_len,
                             const uint8_t** z_bytes, uint64_t* z_len, void* stream) {
  return guarded([&] {
    MODEL_GUARD(m);
    CRA5_CHECK(y && y_bytes && y_len && z_bytes && z_len, ERR_INVALID, "null argument");
    CRA5_CHECK(batch >= 1 && batch <= m->impl->max_batch(), ERR_INVALID, "batch larger than the model's max_batch");
    std::vector<size_t> yl(batch), zl(batch);
    m->impl->latent_to_bin(y, batch, y_bytes, yl.data(), z_bytes, zl.data(), static_cast<cudaStream_t>(stream));
    for (int b = 0; b < batch; ++b) {
      y_len[b] = yl[b];
      z_len[b] = zl[b];
    }
  });
}
int cra5_latent_to_bin(cra5_model* m, const float* y, const uint8_t** y_bytes, uint64_t* y_len, const uint8_t** z_bytes,
                       uint64_t* z_len, void* stream) {
  return cra5_latent_to_bin_batch(m, y, 1, y_bytes, y_len, z_bytes, z_len, stream);
}

int cra5_bin_to_latent_batch(cra5_model* m, const uint8_t* const* y_bytes, const uint64_t* y_len,
                             const uint8_t* const* z_bytes, const uint64_t* z_len, int batch, int z_h, int z_w, float* y_hat,
                             void* stream) {
  return guarded([&] {
    MODEL_GUARD(m);
    CRA5_CHECK(y_hat && y_bytes && y_len && z_bytes && z_len, ERR_INVALID, "null argument");
    CRA5_CHECK(batch >= 1 && batch <= m->impl->max_batch(), ERR_INVALID, "batch larger than the model's max_batch");
    std::vector<size_t> yl(y_len, y_len + batch), zl(z_len, z_len + batch);
    m->impl->bin_to_latent(y_bytes, yl.data(), z_bytes, zl.data(), batch, z_h, z_w, y_hat, static_cast<cudaStream_t>(stream));
  });
}
int cra5_bin_to_latent(cra5_model* m, const uint8_t* y_bytes, uint64_t y_len, const uint8_t* z_bytes, uint64_t z_len,
                       int z_h, int z_w, float* y_hat, void* stream) {
  return cra5_bin_to_latent_batch(m, &y_bytes, &y_len, &z_bytes, &z_len, 1, z_h, z_w, y_hat, stream);
}

int cra5_latent_to_reconstruction_batch(cra5_model* m, const float* y_hat, float* x_hat, int batch, void* stream) {
  return guarded([&] {
    MODEL_GUARD(m);
    CRA5_CHECK(y_hat && x_hat, ERR_INVALID, "null tensor");
    m->impl->latent_to_reconstruction(y_hat, x_hat, batch, static_cast<cudaStream_t>(stream));
  });
}
int cra5_latent_to_reconstruction_denorm(cra5_model* m, const float* y_hat, float* x_hat, const float* mean,
                                         const float* std_, int batch, void* stream) {
  return guarded([&] {
    MODEL_GUARD(m);
    CRA5_CHECK(y_hat && x_hat && mean && std_, ERR_INVALID, "null tensor");
    m->impl->latent_to_reconstruction(y_hat, x_hat, batch, static_cast<cudaStream_t>(stream), mean, std_);
  });
}
int cra5_latent_to_reconstruction(cra5_model* m, const float* y_hat, float* x_hat, void* stream) {
  return cra5_latent_to_reconstruction_batch(m, y_hat, x_hat, 1, stream);
}

int cra5_normalize(const float* in, float* out, const float* mean, const float* std_, int channels, uint64_t hw,
                   int forward, void* stream) {
  return guarded([&] {
    require_sm100();
    CRA5_CHECK(in && out && mean && std_ && channels > 0, ERR_INVALID, "null argument");
    affine_channels(static_cast<cudaStream_t>(stream), in, out, mean, std_, hw, channels, forward);
  });
}

int cra5_model_tap(cra5_model* m, const char* name, const void** dev_ptr, int64_t* numel, int* dtype) {
  return guarded([&] {
    MODEL_GUARD(m);
    *dev_ptr = m->impl->tap(name, numel, dtype);
  });
}

int cra5_launch_count(uint64_t* count) {
  return guarded([&] { *count = launch_count(); });
}
int cra5_profile_enable(int on) {
  return guarded([&] {
    prof_reset();
    prof_enable(on != 0);
  });
}
int cra5_profile_report(char* buf, uint64_t cap, uint64_t* needed) {
  return guarded([&] {
    const std::string js = prof_report_json();
    if (needed) *needed = js.size() + 1;
    if (buf != nullptr && cap > 0) {
      const size_t n = std::min<size_t>(js.size(), cap - 1);
      memcpy(buf, js.data(), n);
      buf[n] = 0;
    }
    prof_reset();
  });
}

int cra5_model_tap_read(cra5_model* m, const char* name, void* dst_dev, uint64_t dst_bytes, void* stream) {
  return guarded([&] {
    MODEL_GUARD(m);
    int64_t numel = 0;
    int dtype = 0;
    const void* src = m->impl->tap(name, &numel, &dtype);
    const size_t esz = (dtype == CRA5_DT_F32 || dtype == CRA5_DT_I32) ? 4 : (dtype == CRA5_DT_BF16 ? 2 : 1);
    CRA5_CHECK((uint64_t)numel * esz <= dst_bytes, ERR_INVALID, "tap_read: destination too small");
    CRA5_CUDA(cudaMemcpyAsync(dst_dev, src, (size_t)numel * esz, cudaMemcpyDeviceToDevice,
                              static_cast<cudaStream_t>(stream)));
  });
}

int cra5_op_gc_quantize(const float* y, const float* sigma, const float* mu, const float* scale_table, int levels,
                        float bound, int32_t* sym, uint8_t* idx, float* y_hat, uint64_t n, void* stream) {
  return guarded([&] {
    require_sm100();
    CRA5_CHECK(scale_table != nullptr || idx == nullptr, ERR_INVALID, "scale table required for indexes");
    CRA5_CHECK(y == nullptr || mu != nullptr, ERR_INVALID, "means required for symbols");
    CRA5_CHECK(idx == nullptr || sigma != nullptr, ERR_INVALID, "scales required for indexes");
    gc_quantize_index(static_cast<cudaStream_t>(stream), y, sigma, mu, scale_table, levels, bound, sym, idx, y_hat, n);
  });
}

namespace {
// the op-level coder entry points keep one lazily grown coder per thread
RansCoder* op_coder(size_t n_symbols, int n_channels) {
  static thread_local std::unique_ptr<RansCoder> coder;
  static thread_local size_t cap_sym = 0;
  static thread_local int cap_ch = 0;
  if (!coder || n_symbols > cap_sym || n_channels > cap_ch) {
    coder.reset();
    cap_sym = std::max<size_t>(n_symbols, cap_sym);
    cap_ch = std::max(n_channels, cap_ch);
    coder = std::make_unique<RansCoder>(std::max<size_t>(cap_sym, 1), std::max(cap_ch, 1));
  }
  return coder.get();
}
}  // namespace

static int op_rans_encode(const int32_t* sym, const uint8_t* idx, const int32_t* cdf, int cdf_rows, int cdf_cols,
                          const int32_t* cdf_len, const int32_t* offset, int n_channels, int L, int spc,
                          uint8_t* out_host, uint64_t out_cap, uint64_t* out_len, void* stream) {
  return guarded([&] {
    require_sm100();
    CRA5_CHECK(out_host && out_len, ERR_INVALID, "null argument");
    CRA5_CHECK(n_channels >= 0 && L >= 0, ERR_INVALID, "negative size");
    CdfTable t;
    t.cdf = cdf; t.length = cdf_len; t.offset = offset; t.rows = cdf_rows; t.cols = cdf_cols;
    if (cdf == nullptr) t.rows = 0;
    RansCoder* c = op_coder((size_t)n_channels * L, n_channels);
    c->invalidate_lut();   // op-level callers may reuse a device pointer for a different table
    *out_len = c->encode(static_cast<cudaStream_t>(stream), sym, idx, t, n_channels, L, spc, out_host, out_cap);
  });
}

static int op_rans_decode(const uint8_t* bytes, uint64_t len, const uint8_t* idx, const int32_t* cdf, int cdf_rows,
                          int cdf_cols, const int32_t* cdf_len, const int32_t* offset, int n_channels, int L,
                          int32_t* sym, void* stream) {
  return guarded([&] {
    require_sm100();
    CdfTable t;
    t.cdf = cdf; t.length = cdf_len; t.offset = offset; t.rows = cdf_rows; t.cols = cdf_cols;
    if (cdf == nullptr) t.rows = 0;
    RansCoder* c = op_coder((size_t)n_channels * L, n_channels);
    c->invalidate_lut();
    c->decode(static_cast<cudaStream_t>(stream), bytes, len, idx, t, n_channels, L, sym, nullptr, nullptr, nullptr);
  });
}

int cra5_op_rans_encode(const int32_t* sym, const uint8_t* idx, const int32_t* cdf, int cdf_cols, const int32_t* cdf_len,
                        const int32_t* offset, int n_channels, int L, int spc, uint8_t* out_host, uint64_t out_cap,
                        uint64_t* out_len, void* stream) {
  // row count unknown: the table stays in global memory
  return op_rans_encode(sym, idx, cdf, 1 << 30, cdf_cols, cdf_len, offset, n_channels, L, spc, out_host, out_cap, out_len,
                        stream);
}

int cra5_op_rans_decode(const uint8_t* bytes, uint64_t len, const uint8_t* idx, const int32_t* cdf, int cdf_cols,
                        const int32_t* cdf_len, const int32_t* offset, int n_channels, int L, int32_t* sym,
                        void* stream) {
  return op_rans_decode(bytes, len, idx, cdf, 1 << 30, cdf_cols, cdf_len, offset, n_channels, L, sym, stream);
}

int cra5_op_rans_encode_table(const int32_t* sym, const uint8_t* idx, const int32_t* cdf, int cdf_rows, int cdf_cols,
                              const int32_t* cdf_len, const int32_t* offset, int n_channels, int L, int spc,
                              uint8_t* out_host, uint64_t out_cap, uint64_t* out_len, void* stream) {
  if (cdf_rows <= 0) return guarded([&] { throw Error(ERR_INVALID, "cdf_rows must be positive"); });
  return op_rans_encode(sym, idx, cdf, cdf_rows, cdf_cols, cdf_len, offset, n_channels, L, spc, out_host, out_cap,
                        out_len, stream);
}

int cra5_op_rans_decode_table(const uint8_t* bytes, uint64_t len, const uint8_t* idx, const int32_t* cdf, int cdf_rows,
                              int cdf_cols, const int32_t* cdf_len, const int32_t* offset, int n_channels, int L,
                              int32_t* sym, void* stream) {
  if (cdf_rows <= 0) return guarded([&] { throw Error(ERR_INVALID, "cdf_rows must be positive"); });
  return op_rans_decode(bytes, len, idx, cdf, cdf_rows, cdf_cols, cdf_len, offset, n_channels, L, sym, stream);
}

}  // extern "C"
