// Microbenchmark: ex2.approx.ftz.f32 / FFMA2 / FFMA issue rates per SM on this GPU (run on the GPU box).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[8];
  uint64_t p[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; p[i] = ((uint64_t)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f); }
  const uint64_t c = ((uint64_t)__float_as_uint(0.999f) << 32) | __float_as_uint(1.001f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2f(a[i]) * 0.0f + a[i];           // 1 MUFU + 1 FFMA
      else if (MODE == 1) p[i] = fma2(p[i], c, c);               // 1 FFMA2
      else if (MODE == 2) a[i] = fmaf(a[i], 0.999f, 0.001f);     // 1 FFMA
      else a[i] = ex2f(a[i]);                                    // MUFU only (values collapse, rate unaffected)
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((uint32_t)p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int warps_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sms * 1024 * 4);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms, warps_per_sm * 32>>>(out, 100, 1.0f);
  cudaEventRecord(e0);
  k<MODE><<<sms, warps_per_sm * 32>>>(out, iters, 1.0f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)sms * warps_per_sm * 32 * iters * 8;
  printf("%-10s warps/SM %2d: %.1f Gop/s total, %.2f thread-ops/ns/SM  (%.3f ms)\n", name, warps_per_sm, ops / ms / 1e6, ops / ms / 1e6 / sms, ms);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 16}) { run<3>("ex2", w); run<0>("ex2+ffma", w); run<1>("ffma2", w); run<2>("ffma", w); }
  return 0;
}
