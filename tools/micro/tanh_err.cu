// Accuracy of tanh.approx.f32 (MUFU.TANH) and of the tanh-form GELU built on it against fp64 references, and the
// throughput of MUFU.TANH next to MUFU.EX2 / MUFU.RCP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tanh_err tanh_err.cu && ./tanh_err
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float gelu_tanh3(float x) {
  const float x2 = x * x;
  const float u = x * fmaf(x2, fmaf(x2, -0.0003515175339619918f, 0.037005650955991044f), 0.797507878425557f);
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(u), h);
}
__global__ void eval(const float* x, float* t, float* g, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { t[i] = tanh_approx(x[i]); g[i] = gelu_tanh3(x[i]); }
}
template <int OP>
__global__ void rate(int iters, long long* out, float* sink) {
  float v[32];
  for (int i = 0; i < 32; ++i) v[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
    }
  }
  long long t1 = clock64();
  float a = 0; for (int i = 0; i < 32; ++i) a += v[i];
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
int main() {
  const int n = 1 << 20;
  std::vector<float> hx(n), ht(n), hg(n);
  for (int i = 0; i < n; ++i) hx[i] = -10.0f + 20.0f * i / (n - 1);
  float *dx, *dt, *dg; cudaMalloc(&dx, 4 * n); cudaMalloc(&dt, 4 * n); cudaMalloc(&dg, 4 * n);
  cudaMemcpy(dx, hx.data(), 4 * n, cudaMemcpyHostToDevice);
  eval<<<n / 256, 256>>>(dx, dt, dg, n);
  cudaMemcpy(ht.data(), dt, 4 * n, cudaMemcpyDeviceToHost); cudaMemcpy(hg.data(), dg, 4 * n, cudaMemcpyDeviceToHost);
  double max_abs_t = 0, max_rel_t = 0, max_abs_g = 0, max_abs_g_neg = 0, max_rel_g_pos = 0; double xa = 0, xg = 0;
  for (int i = 0; i < n; ++i) {
    const double x = hx[i], t = std::tanh(x), g = 0.5 * x * (1.0 + std::erf(x / std::sqrt(2.0)));
    const double et = std::fabs(ht[i] - t), eg = std::fabs(hg[i] - g);
    if (et > max_abs_t) { max_abs_t = et; xa = x; }
    if (std::fabs(t) > 1e-3 && et / std::fabs(t) > max_rel_t) max_rel_t = et / std::fabs(t);
    if (eg > max_abs_g) { max_abs_g = eg; xg = x; }
    if (x < 0 && eg > max_abs_g_neg) max_abs_g_neg = eg;
    if (x > 0.01 && eg / g > max_rel_g_pos) max_rel_g_pos = eg / g;
  }
  printf("tanh.approx.f32: max abs err %.3e (x=%.3f), max rel err %.3e\n", max_abs_t, xa, max_rel_t);
  printf("gelu_tanh3 vs erf-GELU (fp64): max abs err %.3e (x=%.3f); x<0: %.3e; x>0.01 max rel err %.3e (bf16 half ulp = 3.9e-3)\n",
         max_abs_g, xg, max_abs_g_neg, max_rel_g_pos);
  long long* d; float* s; cudaMalloc(&d, 8 * 148); cudaMalloc(&s, 4 * 148 * 256);
  const char* names[3] = {"tanh.approx.f32", "ex2.approx.ftz.f32", "rcp.approx.ftz.f32"};
  for (int op = 0; op < 3; ++op) {
    if (op == 0) rate<0><<<148, 256>>>(2000, d, s); else if (op == 1) rate<1><<<148, 256>>>(2000, d, s); else rate<2><<<148, 256>>>(2000, d, s);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    printf("%-20s %.2f cycles per warp instruction and scheduler (2 warps each)\n", names[op], c / (2000.0 * 32 * 2));
  }
  return 0;
}
