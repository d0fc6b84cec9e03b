// MUFU.EX2 / FFMA throughput microbenchmark (developer tool)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2(a[i]) * 0.5f;          // 1 MUFU + 1 FMUL
      else if (MODE == 1) a[i] = fmaf(a[i], 0.999f, 0.001f);  // 1 FFMA
      else { float t = fmaf(a[i], 0.5f, 0.1f); t = fmaf(t, a[i], 0.2f); t = fmaf(t, a[i], 0.3f); a[i] = fmaf(t, a[i], 1e-3f); }
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  int iters = 20000;
  for (int mode = 0; mode < 3; ++mode) {
    for (int warps : {4, 8, 16, 32}) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      auto launch = [&]() { if (mode == 0) k<0><<<148, warps * 32>>>(out, iters); else if (mode == 1) k<1><<<148, warps * 32>>>(out, iters); else k<2><<<148, warps * 32>>>(out, iters); };
      launch(); cudaDeviceSynchronize();
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double ops = 148.0 * warps * 32 * (double)iters * 8 * (mode == 2 ? 4 : 1);
      printf("mode %d warps/SM %2d: %.3f ms  %.2f Tops/s  = %.2f ops/clk/SM @1.965GHz\n", mode, warps, ms, ops / ms / 1e9, ops / ms / 1e3 / 148 / 1.965e6);
    }
  }
  return 0;
}
