// host_util.h -- host-side helpers shared by the launchers and the C-ABI: error reporting,
// TMA tensor-map encoding (driver entry point fetched at run time, no -lcuda link dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdexcept>
#include <string>

namespace cra5 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// status codes returned through the C-ABI (include/cra5_b200.h)
enum Status : int {
  OK = 0,
  ERR_INVALID = 1,      // bad argument / unsupported geometry   (reference: ValueError)
  ERR_CUDA = 2,         // CUDA runtime/driver failure, or no usable sm_100 device
  ERR_STATE = 3,        // e.g. "Uninitialized CDFs. Run update() first" (entropy_models.py:218-237)
  ERR_BITSTREAM = 4,    // malformed container / truncated stream
  ERR_INTERNAL = 5,
};

#define CRA5_CUDA(expr)                                                                              \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      throw ::cra5::Error(::cra5::ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
  } while (0)

#define CRA5_CHECK(cond, code, msg)                                   \
  do {                                                                \
    if (!(cond)) throw ::cra5::Error((code), std::string(msg));       \
  } while (0)

// bf16 tensor map, up to 4 dims. dims[0] is the contiguous dimension; strides_bytes[i] is the stride of dims[i+1].
CUtensorMap make_tmap_bf16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, bool swizzle128);

inline CUtensorMap make_tmap_bf16_2d(const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                                     uint32_t box_inner, uint32_t box_rows) {
  uint64_t dims[2] = {inner, rows};
  uint64_t strides[1] = {row_stride_bytes};
  uint32_t box[2] = {box_inner, box_rows};
  return make_tmap_bf16(base, 2, dims, strides, box, true);
}

int device_sm_count();   // of the CURRENT device (cached per device ordinal)

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (device, function): one process may drive several GPUs (the Python
// surface takes device=), so the "already configured" state is kept per device ordinal, under a mutex (codec lanes call
// the launchers from several host threads). Raises the limit only when `bytes` exceeds what was set before.
void ensure_dynamic_smem(const void* func, size_t bytes);
template <typename F>
inline void ensure_dynamic_smem(F* func, size_t bytes) { ensure_dynamic_smem(reinterpret_cast<const void*>(func), bytes); }

// ---- launch accounting + optional per-kernel CUDA-event profiler (bench.py's roofline numbers come from here)
void count_launch(int n = 1);
uint64_t launch_count();
void prof_enable(bool on);
bool prof_enabled();
void prof_reset();
void prof_set_tag(const char* tag);  // call-site label appended to the kernel name ("qkv", "fc1", ...)
std::string prof_report_json();      // synchronises, then {"name": {"launches": n, "ms": t, "flops": f, "bytes": b}, ...}

// RAII: counts one launch and, when profiling is on, brackets it with CUDA events on its stream
struct LaunchScope {
  LaunchScope(cudaStream_t st, const char* kernel, double flops, double bytes);
  ~LaunchScope();
  cudaStream_t st_;
  int slot_;
};
// NVTX ranges (header-only nvtx3: a no-op unless a tool such as Nsight Systems / ncu --nvtx is attached), enabled with
// CRA5_NVTX=1: one range per public entry point (encode_to_latent, latent_to_bin, ...) and one per tagged call site
// (qkv, proj, fc1, ...), so a timeline reads like the reference's module tree.
bool nvtx_enabled();
void nvtx_push(const char* name);
void nvtx_pop();
struct NvtxRange {
  explicit NvtxRange(const char* name) : on_(nvtx_enabled()) { if (on_) nvtx_push(name); }
  ~NvtxRange() { if (on_) nvtx_pop(); }
  bool on_;
};
struct TagScope {
  explicit TagScope(const char* tag) : nv_(tag) { prof_set_tag(tag); }
  ~TagScope() { prof_set_tag(""); }
  NvtxRange nv_;
};
void require_sm100();

// Launch of a kernel on the per-frame chain (LayerNorm -> GEMM -> attention -> GEMM ...): a plain stream launch.
// (Programmatic dependent launch was tried here in round 2: 0.7 % faster and NOT bit-identical on a B200 -- removed.)
template <typename... KArgs, typename... Args>
inline void launch_chained(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  kern<<<grid, block, smem, st>>>(args...);
  CRA5_CUDA(cudaGetLastError());
}

}  // namespace cra5
