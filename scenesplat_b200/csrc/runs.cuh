// Run-length primitive over a sorted sequence: given a per-position head flag, one single-pass
// kernel (decoupled look-back) yields the exclusive rank of every position (= id of its run).
#pragma once
#include "common.cuh"

namespace ss {

constexpr int kRunThreads = 256;
constexpr int kRunItems = 8;
constexpr int kRunTile = kRunThreads * kRunItems;

// workspace: [counter u32 * rows (padded to 16B)][status u32 * rows * tiles]
inline size_t runs_counter_bytes(int rows) { return align_up((size_t)rows * 4, 16); }
inline size_t runs_workspace_bytes(int64_t n, int rows = 1) {
  return runs_counter_bytes(rows) + (size_t)rows * ceil_div64(n > 0 ? n : 1, kRunTile) * 4;
}

// Functor contract (row = blockIdx.y, independent sequences of the same length):
//   __device__ bool head(int row, int64_t j)   -> true when position j starts a new run (j == 0 must be true)
//   __device__ void emit(int row, int64_t j, uint32_t run_id, bool is_head)
// After the kernel, *total_out (if non-null) = number of runs of row 0.
template <typename F>
__global__ void __launch_bounds__(kRunThreads) runs_kernel(F f, int64_t n, int tiles, uint32_t* counter,
                                                            uint32_t* status, int64_t* total_out) {
  __shared__ uint32_t s_scan[33];
  __shared__ int s_tile;
  __shared__ uint32_t s_excl;
  const int row = blockIdx.y;
  status += (size_t)row * tiles;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(counter + row, 1u);
  __syncthreads();
  const int tile = s_tile;
  const int64_t base = (int64_t)tile * kRunTile + (int64_t)threadIdx.x * kRunItems;
  uint32_t flags = 0, cnt = 0;
#pragma unroll
  for (int i = 0; i < kRunItems; ++i) {
    const int64_t j = base + i;
    if (j < n && f.head(row, j)) {
      flags |= 1u << i;
      ++cnt;
    }
  }
  uint32_t total;
  uint32_t excl = block_exclusive_scan(cnt, s_scan, total);
  if (threadIdx.x == 0) s_excl = lookback_exclusive(status, 1, tile, total);
  __syncthreads();
  uint32_t run = s_excl + excl;  // number of heads strictly before this thread's first item
#pragma unroll
  for (int i = 0; i < kRunItems; ++i) {
    const int64_t j = base + i;
    if (j < n) {
      const bool h = (flags >> i) & 1u;
      if (h) ++run;
      f.emit(row, j, run - 1u, h);
    }
  }
  if (total_out && row == 0 && (int64_t)(tile + 1) * kRunTile >= n && threadIdx.x == 0) *total_out = (int64_t)(s_excl + total);
}

template <typename F>
inline int runs_launch(F f, int64_t n, void* workspace, int64_t* total_out, cudaStream_t stream, int rows = 1) {
  if (n <= 0) {
    if (total_out) SS_CUDA(cudaMemsetAsync(total_out, 0, 8, stream));
    return SS_OK;
  }
  const int tiles = (int)ceil_div64(n, kRunTile);
  SS_CUDA(cudaMemsetAsync(workspace, 0, runs_workspace_bytes(n, rows), stream));
  uint32_t* counter = (uint32_t*)workspace;
  uint32_t* status = (uint32_t*)((char*)workspace + runs_counter_bytes(rows));
  runs_kernel<F><<<dim3(tiles, rows), kRunThreads, 0, stream>>>(f, n, tiles, counter, status, total_out);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // namespace ss
