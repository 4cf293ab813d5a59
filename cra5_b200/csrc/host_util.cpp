#include "host_util.h"

#include <nvtx3/nvToolsExt.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <utility>
#include <map>
#include <sstream>
#include <vector>

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
  constexpr int MAX_DEV = 64;
  static std::atomic<int> cache[MAX_DEV];   // zero-initialised; 0 = not queried yet
  int dev = 0;
  CRA5_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < MAX_DEV) {
    const int c = cache[dev].load(std::memory_order_relaxed);
    if (c != 0) return c;
  }
  int n = 0;
  CRA5_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  if (dev >= 0 && dev < MAX_DEV) cache[dev].store(n, std::memory_order_relaxed);
  return n;
}

void ensure_dynamic_smem(const void* func, size_t bytes) {
  if (bytes <= 48 * 1024) return;   // the default limit needs no attribute
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> done;
  int dev = 0;
  CRA5_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = done[std::make_pair(dev, func)];
  if (bytes <= have) return;
  CRA5_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  have = bytes;
}

bool nvtx_enabled() {
  static const bool on = [] { const char* e = getenv("CRA5_NVTX"); return e != nullptr && atoi(e) != 0; }();
  return on;
}
void nvtx_push(const char* name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }

void require_sm100() {
  int dev = 0, major = 0;
  CRA5_CUDA(cudaGetDevice(&dev));
  CRA5_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CRA5_CHECK(major == 10, ERR_CUDA, "cra5_b200 kernels are built for sm_100a only; no such device is current");
}

// ---------------------------------------------------------------------------------------------- profiler
namespace {
struct Rec {
  std::string name;
  cudaEvent_t a, b;
  double flops, bytes;
};
struct Prof {
  bool enabled = false;
  std::string tag;
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  uint64_t launches = 0;
};
Prof& prof() {
  static thread_local Prof p;
  return p;
}
cudaEvent_t get_event(Prof& p) {
  if (!p.pool.empty()) {
    cudaEvent_t e = p.pool.back();
    p.pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void count_launch(int n) { prof().launches += (uint64_t)n; }
uint64_t launch_count() { return prof().launches; }
void prof_enable(bool on) { prof().enabled = on; }
bool prof_enabled() { return prof().enabled; }
void prof_set_tag(const char* tag) { prof().tag = tag ? tag : ""; }
void prof_reset() {
  Prof& p = prof();
  for (auto& r : p.recs) {
    p.pool.push_back(r.a);
    p.pool.push_back(r.b);
  }
  p.recs.clear();
}

LaunchScope::LaunchScope(cudaStream_t st, const char* kernel, double flops, double bytes) : st_(st), slot_(-1) {
  Prof& p = prof();
  p.launches += 1;
  if (!p.enabled) return;
  Rec r;
  r.name = p.tag.empty() ? std::string(kernel) : std::string(kernel) + ":" + p.tag;
  r.a = get_event(p);
  r.b = get_event(p);
  r.flops = flops;
  r.bytes = bytes;
  cudaEventRecord(r.a, st);
  slot_ = (int)p.recs.size();
  p.recs.push_back(r);
}
LaunchScope::~LaunchScope() {
  if (slot_ >= 0) cudaEventRecord(prof().recs[slot_].b, st_);
}

std::string prof_report_json() {
  Prof& p = prof();
  cudaDeviceSynchronize();
  struct Agg { uint64_t n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : p.recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    Agg& a = agg[r.name];
    a.n += 1; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
  }
  std::ostringstream os;
  os.precision(9);
  os << "{";
  bool first = true;
  for (auto& kv : agg) {
    if (!first) os << ", ";
    first = false;
    os << "\"" << kv.first << "\": {\"launches\": " << kv.second.n << ", \"ms\": " << kv.second.ms
       << ", \"flops\": " << kv.second.flops << ", \"bytes\": " << kv.second.bytes << "}";
  }
  os << "}";
  return os.str();
}

}  // namespace cra5
