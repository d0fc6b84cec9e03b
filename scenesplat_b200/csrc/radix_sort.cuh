// Stable LSD radix sort of (key, u32 value) rows, one launch per digit pass over ALL rows (blockIdx.y = row).
//
// Shape of a pass (one read and one write of the data, no separate histogram or scan launch):
//   * 4096-key tiles (512 threads x 8 keys), taken in ticket order;
//   * stable ranking inside the warp through warp-private shared-memory tables (warp_rank below: one atomicOr per key
//     collects the lanes with the same digit; measured ~10x cheaper than matching lanes with BITS ballots);
//   * per digit: prefix over the warps, tile aggregate, decoupled look-back over earlier tiles with EIGHT status words
//     probed per round trip, global base from the pass histogram;
//   * the tile is staged sorted by digit in shared memory (aliasing the warp histograms) and written out in runs;
//   * while the keys are still in registers the tile also counts the NEXT pass's digits (shared-memory atomics) and
//     adds them to that pass's global histogram, so the producer of the keys only has to histogram digit 0.
// Digit width adapts to the key width: passes = ceil(key_bits / 10), BITS = ceil(key_bits / passes) in {8, 9, 10}
// (27-bit serialization codes of a depth-9 chunk: 3 passes of 9 bits); keys of <= 32 bits travel as u32 between passes.
//
// Workspace layout: see RadixPlan.  [off_hist, off_hist + small_zero_bytes) must be zero before the PRODUCER kernel
// runs (one small cudaMemsetAsync); [off_status, + status_bytes) must be zero before the first pass: the producer
// kernels clear it themselves (radix_zero_status), which keeps a 1-2 MB memset off the stream.
#pragma once
#include "common.cuh"

namespace ss {

constexpr int kSortThreads = 512;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys / tile
constexpr int kMaxRadixBits = 10;

struct RadixPlan {
  int rows, n, key_bits, bits, passes, tiles, key32;
  size_t off_hist;          // [rows][passes][1 << bits] u32
  size_t off_counter;       // [passes][rows] u32 tile tickets
  size_t small_zero_bytes;  // hist + counters (contiguous from off_hist)
  size_t off_status;        // [passes][rows][tiles][1 << bits] u32
  size_t status_bytes;
  size_t off_keys[2];       // [rows][n] u32 (key32) or u64
  size_t off_vals[2];       // [rows][n] u32
  size_t total;
  __host__ __device__ int bins() const { return 1 << bits; }
  __host__ __device__ size_t hist_row_stride() const { return (size_t)passes << bits; }
};

inline RadixPlan make_radix_plan(int rows, int n, int key_bits) {
  RadixPlan p;
  p.rows = rows;
  p.n = n;
  p.key_bits = key_bits < 1 ? 1 : key_bits;
  p.passes = (p.key_bits + kMaxRadixBits - 1) / kMaxRadixBits;
  p.bits = (p.key_bits + p.passes - 1) / p.passes;
  if (p.bits < 8) p.bits = 8;
  p.key32 = p.key_bits <= 32;
  p.tiles = n > 0 ? (n + kSortTile - 1) / kSortTile : 1;
  size_t o = 0;
  p.off_hist = o;
  o += (size_t)rows * p.passes * p.bins() * 4;
  p.off_counter = o;
  o += align_up((size_t)p.passes * rows * 4, 16);
  p.small_zero_bytes = o;
  o = align_up(o, 256);
  p.off_status = o;
  p.status_bytes = align_up((size_t)p.passes * rows * p.tiles * p.bins() * 4, 16);
  o += p.status_bytes;
  o = align_up(o, 256);
  const size_t ksz = p.key32 ? 4 : 8;
  for (int i = 0; i < 2; ++i) {
    p.off_keys[i] = o;
    o += align_up((size_t)rows * n * ksz, 256);
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
//   kFinal : write order[row][dst] = value and inverse[row][value] = dst as int64 (serialization); the sorted keys too
//            when `sorted_keys` is given
//   kFinalPairs: write sorted keys (u64) and values as int64 `order` (GridSample needs the sorted keys)
enum SortOut { kPairs = 0, kFinal = 1, kFinalPairs = 2 };

// ---------------------------------------------------------------------------------------------------------------
// Producer-side helpers.

// Grid-stride clear of the look-back status words (called by every producer kernel; 16-byte stores).
__device__ __forceinline__ void radix_zero_status(void* status, size_t bytes) {
  uint4* p = reinterpret_cast<uint4*>(status);
  const size_t n16 = bytes >> 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_uint4(0u, 0u, 0u, 0u);
}

// Stable rank of a lane's key among the equal digits this warp has seen so far, from two WARP-PRIVATE shared-memory
// tables: `mask[d]` collects the lanes that hold digit d in this step (atomicOr), `hist[d]` is the running count.
// Blackwell's shared-memory atomics make this ~10x cheaper than matching the lanes with BITS ballots (and 20x cheaper
// than match.any): tools/micro/match_bench.cu, profiles/r2_index.md.  All 32 lanes must call it.
__device__ __forceinline__ uint32_t warp_rank(uint16_t* hist, uint32_t* mask, uint32_t d, bool valid) {
  const unsigned lt = lanemask_lt();
  if (valid) atomicOr(&mask[d], 1u << lane_id());
  __syncwarp();
  uint32_t peers = 0, b = 0;
  if (valid) {
    peers = mask[d];
    b = hist[d];
  }
  __syncwarp();
  if (valid && (peers & lt) == 0u) {  // lowest lane of the group
    hist[d] = (uint16_t)(b + __popc(peers));
    mask[d] = 0u;
  }
  __syncwarp();
  return b + __popc(peers & lt);
}

// Exclusive scans of two values per thread at once over the block (shared barriers).  smem: 2 x 33 words.
__device__ __forceinline__ void block_exclusive_scan2(uint32_t a, uint32_t b, uint32_t* smem, uint32_t& ea, uint32_t& eb,
                                                      uint32_t& ta, uint32_t& tb) {
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const uint32_t ia = warp_inclusive_scan(a), ib = warp_inclusive_scan(b);
  if (lane == 31) {
    smem[warp] = ia;
    smem[33 + warp] = ib;
  }
  __syncthreads();
  if (warp == 0) {
    const uint32_t wa = lane < nwarp ? smem[lane] : 0u, wb = lane < nwarp ? smem[33 + lane] : 0u;
    const uint32_t sa = warp_inclusive_scan(wa), sb = warp_inclusive_scan(wb);
    smem[lane] = sa - wa;
    smem[33 + lane] = sb - wb;
    if (lane == 31) {
      smem[32] = sa;
      smem[65] = sb;
    }
  }
  __syncthreads();
  ea = smem[warp] + ia - a;
  eb = smem[33 + warp] + ib - b;
  ta = smem[32];
  tb = smem[65];
  __syncthreads();
}

// Decoupled look-back with kProbe predecessors probed per round trip (one status word per (tile, slot)).
constexpr int kProbe = 8;
__device__ __forceinline__ uint32_t lookback_exclusive_batched(uint32_t* status, int stride, int tile, uint32_t aggregate) {
  uint32_t* mine = status + (size_t)tile * stride;
  if (tile == 0) {
    st_volatile_u32(mine, kFlagInclusive | aggregate);
    return 0u;
  }
  st_volatile_u32(mine, kFlagAggregate | aggregate);
  uint32_t excl = 0u;
  int t = tile - 1;
  while (true) {
    uint32_t v[kProbe];
#pragma unroll
    for (int i = 0; i < kProbe; ++i)
      v[i] = t - i >= 0 ? ld_volatile_u32(status + (size_t)(t - i) * stride) : kFlagInclusive;  // tile -1: inclusive 0
    int used = 0;
    bool done = false;
#pragma unroll
    for (int i = 0; i < kProbe; ++i) {
      if (!done && used == i && (v[i] & (kFlagInclusive | kFlagAggregate)) != 0u) {
        excl += v[i] & kValueMask;
        used = i + 1;
        if (v[i] & kFlagInclusive) done = true;
      }
    }
    if (done) break;
    t -= used;
  }
  st_volatile_u32(mine, kFlagInclusive | (excl + aggregate));
  return excl;
}

// Runs all passes (radix_sort.cu).  `keys0` = initial keys [rows][n] as u64 (values implicit = index).  The producer
// has cleared the status words and left the digit-0 histogram of every row at ghist[row * hist_row_stride + d].
// mode kFinal: writes order / inverse; kFinalPairs: writes the sorted keys (u64) to `sorted_keys` + order.
int radix_sort_run(const RadixPlan& p, char* ws, const uint64_t* keys0, int final_mode, int64_t* order_out,
                   int64_t* inverse_out, uint64_t* sorted_keys, cudaStream_t stream);

}  // namespace ss
