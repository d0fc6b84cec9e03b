// Run-length primitive over a sorted sequence: given a per-position head flag, one single-pass
// kernel (decoupled look-back) yields the exclusive rank of every position (= id of its run).
//
// Tiles are taken in ticket order, row 0's tiles first; the prefix over earlier tiles is read in parallel
// (tiles_exclusive_all).  A functor may declare `kAfterRow0 = true`: its rows >= 1 then run their head flags, scan and look-back
// immediately but wait with `emit` until every tile of row 0 has emitted (row 0's tickets are all smaller, so its tiles
// are resident or done: no deadlock) -- this lets one launch build results that rows >= 1 derive from row 0's.
#pragma once
#include "common.cuh"

namespace ss {

#ifdef SS_RUNS_TRACE  // tools/micro/runs_trace.cu: per-CTA %globaltimer at the phase boundaries
__device__ unsigned long long* g_runs_trace;
__device__ __forceinline__ void runs_trace(int slot, unsigned long long extra = 0) {
  if (threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_runs_trace[(size_t)blockIdx.x * 8 + slot] = slot == 7 ? extra : t;
  }
}
#define SS_RUNS_TRACE_AT(slot, ...) runs_trace(slot, ##__VA_ARGS__)
#else
#define SS_RUNS_TRACE_AT(slot, ...)
#endif

constexpr int kRunThreads = 256;
constexpr int kRunItems = 4;
constexpr int kRunTile = kRunThreads * kRunItems;

// workspace: [ticket u32 | pad to 128 B][row-0 done counter u32 | pad to 128 B][status u32 * rows * tiles]
// (the ticket, the polled done counter and the look-back status words live in separate 128-byte lines: pollers of one
// must not queue in front of the others at the L2 slice)
inline size_t runs_counter_bytes(int) { return 256; }
inline size_t runs_workspace_bytes(int64_t n, int rows = 1) {
  return runs_counter_bytes(rows) + (size_t)rows * ceil_div64(n > 0 ? n : 1, kRunTile) * 4;
}

template <typename F, typename = void>
struct runs_after_row0 { static constexpr bool value = false; };
template <typename F>
struct runs_after_row0<F, decltype((void)F::kAfterRow0)> { static constexpr bool value = F::kAfterRow0; };

// Exclusive prefix over the tiles before `tile` WITHOUT a chain: every tile publishes its own count, then the whole CTA
// reads the counts of all earlier tiles in parallel (one round trip once they are published; tickets are handed out in
// order, so every earlier tile is resident or done).  tiles^2 / 2 words of L2 reads in total: ~43 k words for a
// 300 k-point sequence, against a chained look-back whose inclusive prefixes advance 32 tiles per round trip.
// Returns the prefix to all threads.  `smem` needs 9 words; contains __syncthreads.
__device__ __forceinline__ uint32_t tiles_exclusive_all(uint32_t* status, int tile, uint32_t aggregate, uint32_t* smem) {
  if (threadIdx.x == 0) st_volatile_u32(status + tile, kFlagAggregate | aggregate);
  uint32_t acc = 0u;
  for (int t = (int)threadIdx.x; t < tile; t += (int)blockDim.x) {
    uint32_t v;
    while (((v = ld_volatile_u32(status + t)) & kFlagAggregate) == 0u) {}
    acc += v & kValueMask;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane_id() == 0) smem[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t sum = 0u;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum += smem[w];
    smem[8] = sum;
  }
  __syncthreads();
  return smem[8];
}

// Functor contract (rows = independent sequences of the same length):
//   __device__ bool head(int row, int64_t j)   -> true when position j starts a new run (j == 0 must be true)
//   __device__ void emit(int row, int64_t j, uint32_t run_id, bool is_head)
//   optional: __device__ void finish(int row, uint32_t total_runs)   (called once per row by the row's last tile)
// After the kernel, *total_out (if non-null) = number of runs of row 0.
template <typename F>
__global__ void __launch_bounds__(kRunThreads) runs_kernel(F f, int64_t n, int tiles, uint32_t* counter,
                                                            uint32_t* status, int64_t* total_out) {
  __shared__ uint32_t s_scan[33];
  __shared__ int s_ticket;
  __shared__ uint32_t s_excl;
  SS_RUNS_TRACE_AT(0);
  if (threadIdx.x == 0) s_ticket = (int)atomicAdd(counter, 1u);
  __syncthreads();
  const int row = s_ticket / tiles, tile = s_ticket - row * tiles;
  SS_RUNS_TRACE_AT(1);
  SS_RUNS_TRACE_AT(7, (unsigned long long)row);
  status += (size_t)row * tiles;
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
  SS_RUNS_TRACE_AT(2);
  uint32_t total;
  uint32_t excl = block_exclusive_scan(cnt, s_scan, total);
  SS_RUNS_TRACE_AT(3);
  const uint32_t before = tiles_exclusive_all(status, tile, total, s_scan);
  SS_RUNS_TRACE_AT(4);
  if (threadIdx.x == 0) {
    s_excl = before;
    if (runs_after_row0<F>::value && row > 0) {  // row 0's emits must be visible before ours start
      unsigned ns = 1000;  // row 0 needs microseconds: poll rarely, then faster
      while (ld_volatile_u32(counter + 32) < (uint32_t)tiles) {
        __nanosleep(ns);
        ns = ns > 400 ? ns / 2 : 200;
      }
      __threadfence();
    }
  }
  __syncthreads();
  SS_RUNS_TRACE_AT(5);
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
  const bool last_tile = (int64_t)(tile + 1) * kRunTile >= n;
  if (last_tile && threadIdx.x == 0) {
    if (total_out && row == 0) *total_out = (int64_t)(s_excl + total);
    f.finish(row, s_excl + total);
  }
  if (runs_after_row0<F>::value && row == 0) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(counter + 32, 1u);
    }
  }
#ifdef SS_RUNS_TRACE
  __syncthreads();
  SS_RUNS_TRACE_AT(6);
#endif
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
  runs_kernel<F><<<tiles * rows, kRunThreads, 0, stream>>>(f, n, tiles, counter, status, total_out);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // namespace ss
