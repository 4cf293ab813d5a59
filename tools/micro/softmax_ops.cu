// Reciprocal throughput (cycles per warp instruction and SM sub-partition) of the instructions a softmax inner loop is
// made of, measured with W warps per CTA on every SM: tells which pipe bounds attn_tc4's softmax warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_ops softmax_ops.cu && ./softmax_ops
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

constexpr int CH = 32;   // independent chains per thread

#define OP_LOOP(NAME, DECL, BODY, FOLD)                                                     \
  __global__ void NAME(int iters, long long* out, uint32_t* sink) {                        \
    DECL;                                                                                  \
    __syncthreads();                                                                       \
    long long t0 = clock64();                                                              \
    for (int it = 0; it < iters; ++it) {                                                   \
      _Pragma("unroll") for (int i = 0; i < CH; ++i) { BODY; }                             \
    }                                                                                      \
    long long t1 = clock64();                                                              \
    uint32_t acc = 0;                                                                      \
    _Pragma("unroll") for (int i = 0; i < CH; ++i) { FOLD; }                               \
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;                                       \
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;                                     \
  }

#define F32_DECL float v[CH]; for (int i = 0; i < CH; ++i) v[i] = -1.0f - 0.001f * (threadIdx.x + i)
#define U32_DECL uint32_t v[CH]; for (int i = 0; i < CH; ++i) v[i] = 0xb800b800u + threadIdx.x + i
#define U64_DECL uint64_t v[CH]; for (int i = 0; i < CH; ++i) v[i] = 0x3f8000003f800000ull + threadIdx.x + i; \
                 uint64_t ca = 0x3f8000013f800001ull + threadIdx.x, cb = 0x3a8000013a800001ull + threadIdx.x

OP_LOOP(k_ex2_f32, F32_DECL, asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i])), acc += __float_as_uint(v[i]))
OP_LOOP(k_ex2_f16x2, U32_DECL, asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i])), acc += v[i])
OP_LOOP(k_ex2_bf16x2, U32_DECL, asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v[i])), acc += v[i])
OP_LOOP(k_fma_f32, F32_DECL; float a = 1.0001f + threadIdx.x * 1e-9f; float b = 1e-3f + threadIdx.x * 1e-9f,
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(a), "f"(b)), acc += __float_as_uint(v[i]))
OP_LOOP(k_fma_f32x2, U64_DECL, asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(ca), "l"(cb)), acc += (uint32_t)v[i])
OP_LOOP(k_add_f32x2, U64_DECL, asm volatile("add.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(cb)), acc += (uint32_t)v[i])
OP_LOOP(k_add_f32, F32_DECL; float b = 1e-3f + threadIdx.x * 1e-9f, asm volatile("add.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(b)),
        acc += __float_as_uint(v[i]))
OP_LOOP(k_max3_f32, F32_DECL; float a = 1.0001f + threadIdx.x * 1e-9f; float b = 1e-3f + threadIdx.x * 1e-9f,
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(a), "f"(b)), acc += __float_as_uint(v[i]))
OP_LOOP(k_max_f32, F32_DECL; float a = 1.0001f + threadIdx.x * 1e-9f, asm volatile("max.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(a)),
        acc += __float_as_uint(v[i]))
// pack two fp32 into bf16x2 / f16x2 (F2FP), unpack f16 -> f32, hfma2 / hadd2
OP_LOOP(k_cvt_bf16x2, F32_DECL; float a = 1.0001f + threadIdx.x * 1e-9f,
        { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[i]), "f"(a)); v[i] = __uint_as_float(r); },
        acc += __float_as_uint(v[i]))
OP_LOOP(k_cvt_f16x2, F32_DECL; float a = 1.0001f + threadIdx.x * 1e-9f,
        { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[i]), "f"(a)); v[i] = __uint_as_float(r); },
        acc += __float_as_uint(v[i]))
OP_LOOP(k_cvt_f32_f16, U32_DECL,
        { float r; asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, hi;}" : "=f"(r) : "r"(v[i])); v[i] = __float_as_uint(r); },
        acc += v[i])
OP_LOOP(k_hadd2, U32_DECL; uint32_t b = 0x14001400u + threadIdx.x, asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(v[i]) : "r"(b)), acc += v[i])
OP_LOOP(k_hfma2, U32_DECL; uint32_t a = 0x3c003c00u; uint32_t b = 0x14001400u + threadIdx.x,
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(a), "r"(b)), acc += v[i])
OP_LOOP(k_iadd, U32_DECL; uint32_t b = threadIdx.x, asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(b)), acc += v[i])
// mixes: ex2.f32 and fma2 / ex2.f16x2 and cvt interleaved (do they overlap?)
OP_LOOP(k_mix_ex2f32_fma2, F32_DECL; uint64_t w[CH]; for (int i = 0; i < CH; ++i) w[i] = 0x3f8000003f800000ull + i;
        uint64_t ca = 0x3f8000013f800001ull + threadIdx.x; uint64_t cb = 0x3a8000013a800001ull,
        { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i])); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[i]) : "l"(ca), "l"(cb)); },
        acc += __float_as_uint(v[i]) + (uint32_t)w[i])
OP_LOOP(k_mix_ex2h2_cvt, U32_DECL; float f[CH]; for (int i = 0; i < CH; ++i) f[i] = -1.0f - 0.01f * i; float a = 1.0001f + threadIdx.x * 1e-9f,
        { asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i])); uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(f[i]), "f"(a)); f[i] = __uint_as_float(r); },
        acc += v[i] + __float_as_uint(f[i]))

template <typename K>
void run(const char* name, K kern, int ops_per_body = 1) {
  long long* d; uint32_t* s;
  cudaMalloc(&d, 8 * 148); cudaMalloc(&s, 4 * 148 * 1024);
  for (int warps : {4, 8, 12}) {
    const int iters = 4000;
    kern<<<148, warps * 32>>>(iters, d, s);
    kern<<<148, warps * 32>>>(iters, d, s);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i]; cyc /= 148;
    const double per_smsp = warps / 4.0;
    printf("%-22s warps/SMSP %.0f: %.2f cycles per warp-instruction-group per SMSP (%d op%s per group)\n", name, per_smsp,
           cyc / (iters * (double)CH * per_smsp), ops_per_body, ops_per_body > 1 ? "s" : "");
  }
  cudaFree(d); cudaFree(s);
}

int main() {
  run("ex2.f32", k_ex2_f32);
  run("ex2.f16x2", k_ex2_f16x2);
  run("ex2.bf16x2", k_ex2_bf16x2);
  run("fma.f32", k_fma_f32);
  run("fma.f32x2", k_fma_f32x2);
  run("add.f32x2", k_add_f32x2);
  run("add.f32", k_add_f32);
  run("max3.f32", k_max3_f32);
  run("max.f32", k_max_f32);
  run("cvt.bf16x2.f32", k_cvt_bf16x2);
  run("cvt.f16x2.f32", k_cvt_f16x2);
  run("cvt.f32.f16", k_cvt_f32_f16);
  run("add.f16x2", k_hadd2);
  run("fma.f16x2", k_hfma2);
  run("add.u32", k_iadd);
  run("ex2.f32 + fma.f32x2", k_mix_ex2f32_fma2, 2);
  run("ex2.f16x2 + cvt.f16x2", k_mix_ex2h2_cvt, 2);
  return 0;
}
