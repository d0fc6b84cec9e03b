// MUFU throughput of packed half-precision exponentials vs fp32 (developer tool)
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2h2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned ex2b2(unsigned x) { unsigned y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float tanhf_(float x) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(unsigned* out, int iters) {
  unsigned a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0x3c003c00u + threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = __float_as_uint(ex2f(__uint_as_float(a[i])));
      else if (MODE == 1) a[i] = ex2h2(a[i]);
      else if (MODE == 2) a[i] = ex2b2(a[i]);
      else a[i] = __float_as_uint(tanhf_(__uint_as_float(a[i])));
    }
  }
  unsigned s = 0; for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  unsigned* out; cudaMalloc(&out, 148 * 1024 * 4);
  int iters = 20000;
  const char* names[4] = {"ex2.f32", "ex2.f16x2", "ex2.bf16x2", "tanh.f32"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int warps : {4, 8, 16}) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      auto launch = [&]() { if (mode == 0) k<0><<<148, warps * 32>>>(out, iters); else if (mode == 1) k<1><<<148, warps * 32>>>(out, iters); else if (mode == 2) k<2><<<148, warps * 32>>>(out, iters); else k<3><<<148, warps * 32>>>(out, iters); };
      launch(); cudaDeviceSynchronize();
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double instr = 148.0 * warps * 32 * (double)iters * 8;
      printf("%-10s warps/SM %2d: %.3f ms  %.2f T lane-instr/s (x2 values for the packed forms)\n", names[mode], warps, ms, instr / ms / 1e9);
    }
  }
  return 0;
}
