// capi.cpp -- extern "C" surface of libcra5b200.so (declared in include/cra5_b200.h).
#include "../../include/cra5_b200.h"

#include "capi_util.h"
#include "gemm_tc.cuh"
#include "kernels.h"

namespace cra5 {
std::string& last_error_slot() {
  static thread_local std::string s;
  return s;
}
}  // namespace cra5

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

}  // extern "C"
