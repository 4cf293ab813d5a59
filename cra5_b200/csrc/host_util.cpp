#include "host_util.h"

#include <mutex>

namespace cra5 {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  if (fn == nullptr) throw Error(ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver / GPU?)");
  return fn;
}

CUtensorMap make_tmap_bf16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, bool swizzle128) {
  CRA5_CHECK(rank >= 1 && rank <= 5, ERR_INTERNAL, "tensor map rank");
  CRA5_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, ERR_INVALID, "tensor map base must be 16-byte aligned");
  CUtensorMap m;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) {
      gstr[i] = strides_bytes[i];
      CRA5_CHECK((gstr[i] & 15) == 0, ERR_INVALID, "tensor map stride must be a multiple of 16 bytes");
    }
  }
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim,
                               gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return m;
}

int device_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    CRA5_CUDA(cudaGetDevice(&dev));
    CRA5_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

void require_sm100() {
  int dev = 0, major = 0;
  CRA5_CUDA(cudaGetDevice(&dev));
  CRA5_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CRA5_CHECK(major == 10, ERR_CUDA, "cra5_b200 kernels are built for sm_100a only; no such device is current");
}

}  // namespace cra5
