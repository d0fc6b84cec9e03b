// Microbenchmark: cost of grouping the lanes of a warp by a 9-bit digit, (a) with BITS ballots (what CUB's radix rank
// does) and (b) with match.any, in the setting of the radix pass: 16 warps per CTA, 2 CTAs per SM, random digits,
// followed by the warp-private histogram update.  Prints ns per warp-item (32 keys) per SM-resident warp set.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o match_bench match_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int BITS>
__device__ __forceinline__ unsigned peers_ballot(uint32_t d) {
  unsigned peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < BITS; ++b) {
    const bool bit = (d >> b) & 1u;
    const unsigned m = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? m : ~m;
  }
  return peers;
}

template <int MODE>
__global__ void __launch_bounds__(512) k(uint32_t* out, int iters, uint32_t seed) {
  __shared__ uint32_t wh[16][512];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 16 * 512; i += 512) (&wh[0][0])[i] = 0;
  __syncthreads();
  uint32_t x = seed ^ (blockIdx.x * 512 + threadIdx.x) * 2654435761u, acc = 0;
  for (int it = 0; it < iters; ++it) {
    x = x * 1664525u + 1013904223u;
    const uint32_t d = (x >> 13) & 511u;
    unsigned peers;
    if (MODE == 0) peers = peers_ballot<9>(d);
    else peers = __match_any_sync(0xffffffffu, d);
    if (MODE < 2) {
      const int leader = __ffs(peers) - 1;
      uint32_t b = 0;
      if (lane == leader) {
        b = wh[warp][d];
        wh[warp][d] = b + __popc(peers);
      }
      b = __shfl_sync(0xffffffffu, b, leader);
      __syncwarp();
      acc += b + __popc(peers & ((1u << lane) - 1u));
    } else {
      acc += atomicAdd(&wh[warp][d], 1u);  // MODE 2: shared-memory atomic per key (unordered ranks: counting only)
    }
  }
  out[blockIdx.x * 512 + threadIdx.x] = acc;
}

int main() {
  uint32_t* out;
  cudaMalloc(&out, 296 * 512 * 4);
  const int iters = 4096;
  const char* names[3] = {"9 ballots + warp-private RMW", "match.any + warp-private RMW", "shared-memory atomicAdd"};
  for (int mode = 0; mode < 3; ++mode) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a);
      if (mode == 0) k<0><<<296, 512>>>(out, iters, 1);
      else if (mode == 1) k<1><<<296, 512>>>(out, iters, 1);
      else k<2><<<296, 512>>>(out, iters, 1);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (rep == 2)
        printf("%-32s %8.3f ms for %d warp-items per warp (32 warps / SM): %6.1f ns per warp-item per SM -> a 4096-key "
               "tile (128 warp-items, 2 tiles / SM): %5.2f us\n",
               names[mode], ms, iters, ms * 1e6 / (iters * 32.0), ms * 1e3 / iters * 128 / 16);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
