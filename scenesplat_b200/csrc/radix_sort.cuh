// Stable LSD radix sort of (u64 key, u32 value) rows, 8-bit digits, one launch per pass over ALL
// rows (blockIdx.y = row).  Single-pass-per-digit design: per-tile warp-level multi-split ranking
// (match.any) + decoupled look-back across tiles (chained scan), so a pass is one read and one
// write of the data.  Digit histograms for all passes are produced up front by the producer of the
// keys (e.g. the fused encode kernel) and exclusive-scanned inside every CTA's prologue.
//
// Workspace layout (all u32 unless noted), see RadixPlan.
#pragma once
#include "common.cuh"

namespace ss {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048 keys / tile

struct RadixPlan {
  int rows, n, passes, tiles;
  size_t off_hist;     // [rows][passes][256] u32
  size_t off_counter;  // [passes][rows] u32
  size_t off_status;   // [passes][rows][tiles][256] u32
  size_t zero_bytes;   // bytes from off_hist that must be zeroed before use
  size_t off_keys[2];  // [rows][n] u64
  size_t off_vals[2];  // [rows][n] u32
  size_t total;
};

inline RadixPlan make_radix_plan(int rows, int n, int key_bits) {
  RadixPlan p;
  p.rows = rows;
  p.n = n;
  p.passes = key_bits <= 0 ? 1 : (key_bits + kRadixBits - 1) / kRadixBits;
  p.tiles = n > 0 ? (n + kSortTile - 1) / kSortTile : 1;
  size_t o = 0;
  p.off_hist = o;
  o += (size_t)rows * p.passes * kRadix * 4;
  p.off_counter = o;
  o += align_up((size_t)p.passes * rows * 4, 16);
  p.off_status = o;
  o += (size_t)p.passes * rows * p.tiles * kRadix * 4;
  p.zero_bytes = o;
  o = align_up(o, 256);
  for (int i = 0; i < 2; ++i) {
    p.off_keys[i] = o;
    o += align_up((size_t)rows * n * 8, 256);
  }
  for (int i = 0; i < 2; ++i) {
    p.off_vals[i] = o;
    o += align_up((size_t)rows * n * 4, 256);
  }
  p.total = o;
  return p;
}

// Output modes of a pass.
//   kPairs : write (key, value) to the ping-pong buffers
//   kFinal : write order[row][dst] = value and inverse[row][value] = dst as int64 (serialization)
//   kFinalPairs: write sorted keys (u64) and values as int64 `order` (GridSample needs the sorted keys)
enum SortOut { kPairs = 0, kFinal = 1, kFinalPairs = 2 };

template <int OUT>
__global__ void __launch_bounds__(kSortThreads)
radix_pass_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                  int64_t* __restrict__ order_out, int64_t* __restrict__ inverse_out,
                  const uint32_t* __restrict__ ghist,  // [rows][passes][256]
                  uint32_t* __restrict__ counters,     // [rows] for this pass
                  uint32_t* __restrict__ status,       // [rows][tiles][256] for this pass
                  int n, int tiles, int passes, int pass, int shift, size_t row_stride) {
  __shared__ uint32_t s_whist[kSortThreads / 32][kRadix + 1];
  __shared__ uint32_t s_digit_start[kRadix];  // start of digit run inside the staged tile
  __shared__ uint32_t s_global_base[kRadix];  // global destination of the first key of the run
  __shared__ uint32_t s_scan[33];
  __shared__ uint64_t s_keys[kSortTile];
  __shared__ uint32_t s_vals[kSortTile];
  __shared__ int s_tile;

  const int row = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  keys_in += (size_t)row * row_stride;
  if (vals_in) vals_in += (size_t)row * row_stride;
  if (OUT == kPairs) {
    keys_out += (size_t)row * row_stride;
    vals_out += (size_t)row * row_stride;
  } else {
    order_out += (size_t)row * row_stride;
    if (OUT == kFinal) inverse_out += (size_t)row * row_stride;
    if (OUT == kFinalPairs) keys_out += (size_t)row * row_stride;
  }

  if (tid == 0) s_tile = (int)atomicAdd(&counters[row], 1u);
  for (int i = tid; i < (kSortThreads / 32) * (kRadix + 1); i += kSortThreads) (&s_whist[0][0])[i] = 0u;
  __syncthreads();
  const int tile = s_tile;
  const int base = tile * kSortTile;
  const int tile_n = min(kSortTile, n - base);

  // ---- load (warp-contiguous chunks keep the stable order simple) and rank inside the warp
  uint64_t key[kSortItems];
  uint32_t val[kSortItems];
  uint32_t rank[kSortItems];
  uint32_t dig[kSortItems];
  const int wbase = warp * (32 * kSortItems);
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    const int li = wbase + it * 32 + lane;
    const bool ok = li < tile_n;
    key[it] = ok ? keys_in[base + li] : 0ull;
    val[it] = ok ? (vals_in ? vals_in[base + li] : (uint32_t)(base + li)) : 0u;
    dig[it] = ok ? (uint32_t)((key[it] >> shift) & (kRadix - 1)) : (uint32_t)kRadix;
  }
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    const uint32_t d = dig[it];
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    const uint32_t cnt = __popc(peers);
    const uint32_t before = __popc(peers & lanemask_lt());
    uint32_t b = 0;
    if (lane == leader) {
      b = s_whist[warp][d];
      s_whist[warp][d] = b + cnt;
    }
    b = __shfl_sync(0xffffffffu, b, leader);
    rank[it] = b + before;
    __syncwarp();
  }
  __syncthreads();

  // ---- per digit (thread d): prefix over warps, tile aggregate, look-back, global base
  {
    const int d = tid;  // kSortThreads == kRadix
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) {
      uint32_t c = s_whist[w][d];
      s_whist[w][d] = run;
      run += c;
    }
    // global start of this digit's bin = exclusive scan of the pass histogram
    const uint32_t gcount = ghist[((size_t)row * passes + pass) * kRadix + d];
    uint32_t gtotal;
    const uint32_t bin_start = block_exclusive_scan(gcount, s_scan, gtotal);
    uint32_t ttotal;
    const uint32_t dstart = block_exclusive_scan(run, s_scan, ttotal);
    s_digit_start[d] = dstart;
    const uint32_t excl = lookback_exclusive(status + ((size_t)row * tiles) * kRadix + d, kRadix, tile, run);
    s_global_base[d] = bin_start + excl;
  }
  __syncthreads();

  // ---- stage the tile sorted by digit
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    const uint32_t d = dig[it];
    if (d < (uint32_t)kRadix) {
      const uint32_t pos = s_digit_start[d] + s_whist[warp][d] + rank[it];
      s_keys[pos] = key[it];
      s_vals[pos] = val[it];
    }
  }
  __syncthreads();

  // ---- coalesced scatter
  for (int i = tid; i < tile_n; i += kSortThreads) {
    const uint64_t k = s_keys[i];
    const uint32_t v = s_vals[i];
    const uint32_t d = (uint32_t)((k >> shift) & (kRadix - 1));
    const uint32_t dst = s_global_base[d] + ((uint32_t)i - s_digit_start[d]);
    if (OUT == kPairs) {
      keys_out[dst] = k;
      vals_out[dst] = v;
    } else if (OUT == kFinal) {
      order_out[dst] = (int64_t)v;
      inverse_out[v] = (int64_t)dst;
    } else {
      keys_out[dst] = k;
      order_out[dst] = (int64_t)v;
    }
  }
}

// Runs all passes.  `keys0` = initial keys [rows][n] (values implicit = index).  The workspace region
// [off_hist, off_hist + zero_bytes) must already hold the histograms (and zeros elsewhere).
// mode kFinal: writes order/inverse; kFinalPairs: writes sorted keys to `sorted_keys` + order.
inline int radix_sort_run(const RadixPlan& p, char* ws, const uint64_t* keys0, int final_mode,
                          int64_t* order_out, int64_t* inverse_out, uint64_t* sorted_keys, cudaStream_t stream) {
  if (p.n <= 0) return SS_OK;
  const uint32_t* ghist = (const uint32_t*)(ws + p.off_hist);
  uint32_t* counters = (uint32_t*)(ws + p.off_counter);
  uint32_t* status = (uint32_t*)(ws + p.off_status);
  uint64_t* kbuf[2] = {(uint64_t*)(ws + p.off_keys[0]), (uint64_t*)(ws + p.off_keys[1])};
  uint32_t* vbuf[2] = {(uint32_t*)(ws + p.off_vals[0]), (uint32_t*)(ws + p.off_vals[1])};
  dim3 grid(p.tiles, p.rows);
  const uint64_t* kin = keys0;
  const uint32_t* vin = nullptr;
  for (int pass = 0; pass < p.passes; ++pass) {
    uint32_t* cnt = counters + (size_t)pass * p.rows;
    uint32_t* st = status + (size_t)pass * p.rows * p.tiles * kRadix;
    const int shift = pass * kRadixBits;
    if (pass == p.passes - 1) {
      if (final_mode == kFinal)
        radix_pass_kernel<kFinal><<<grid, kSortThreads, 0, stream>>>(kin, vin, nullptr, nullptr, order_out, inverse_out,
                                                                     ghist, cnt, st, p.n, p.tiles, p.passes, pass, shift,
                                                                     (size_t)p.n);
      else
        radix_pass_kernel<kFinalPairs><<<grid, kSortThreads, 0, stream>>>(kin, vin, sorted_keys, nullptr, order_out,
                                                                          nullptr, ghist, cnt, st, p.n, p.tiles, p.passes,
                                                                          pass, shift, (size_t)p.n);
    } else {
      const int o = pass & 1;
      radix_pass_kernel<kPairs><<<grid, kSortThreads, 0, stream>>>(kin, vin, kbuf[o], vbuf[o], nullptr, nullptr, ghist,
                                                                   cnt, st, p.n, p.tiles, p.passes, pass, shift,
                                                                   (size_t)p.n);
      kin = kbuf[o];
      vin = vbuf[o];
    }
    SS_CHECK_LAUNCH();
  }
  return SS_OK;
}

}  // namespace ss
