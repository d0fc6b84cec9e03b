// Kernel map (neighbour table) of a submanifold k^3 convolution, built from the serialization codes the Point
// already has: the codes go into an open-addressing table in the workspace (one CAS per voxel), every voxel then
// looks its neighbours' codes up (1 - 2 probes of 16 bytes each; round 1 - 2 searched the sorted row: 19 dependent
// loads per lookup).  Only half of the taps are searched; the mirrored tap is filled through the
// symmetry nbr[t][p] = q  <=>  nbr[k^3-1-t][q] = p.
//
// Replaces (reference): the indice-pair build inside spconv.SubMConv3d
// (call sites point_transformer_v3m1_base.py:277-284, :499-506; tensor built structure.py:131-138).
// tap t = (i*k + j)*k + l  <->  offset (i-r, j-r, l-r) on (x, y, z);  nbr is tap-major [k^3][n] int32.
#include "runs.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

// Open-addressing table of the active voxels: 16-byte entries {key, voxel}, linear probing, at most half full.
// Equal keys (not produced by GridSample, but legal input) keep the SMALLEST voxel index: what a lower-bound search
// in a stably sorted row returns.
struct __align__(16) KmapSlot {
  unsigned long long key;
  unsigned int voxel;
  unsigned int pad;
};
constexpr unsigned long long kKmapEmpty = ~0ull;  // codes use at most 48 + batch bits

inline int64_t kmap_table_slots(int64_t n) {
  int64_t s = 1024;
  while (s < 2 * n) s <<= 1;
  return s;
}
inline size_t kmap_table_bytes(int64_t n) { return align_up((size_t)kmap_table_slots(n > 0 ? n : 1) * sizeof(KmapSlot), 256); }

__device__ __forceinline__ uint32_t kmap_hash(uint64_t key, int shift) {
  return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> shift);
}

__global__ void __launch_bounds__(256)
kmap_insert_kernel(const int64_t* __restrict__ code, int64_t n, KmapSlot* __restrict__ table, int shift, uint32_t mask) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const unsigned long long key = (unsigned long long)code[p];
  uint32_t slot = kmap_hash(key, shift);
  while (true) {
    const unsigned long long old = atomicCAS(&table[slot].key, kKmapEmpty, key);
    if (old == kKmapEmpty || old == key) {
      atomicMin(&table[slot].voxel, (unsigned int)p);
      return;
    }
    slot = (slot + 1u) & mask;
  }
}

// One thread per voxel p (coalesced coordinate reads and direct-tap writes); the lower half of the taps is looked up,
// the mirrored tap is written at the neighbour.  K is a template parameter (tap offsets are constants of the unrolled
// loop: no integer divisions), and for the z orders the key of a neighbour is the OR of three per-axis terms that are
// spread ONCE per voxel and offset (3 K spreads instead of 3 per tap: ncu put the first version of this kernel at 245
// warp instructions per lookup, issue-bound, most of them the 64-bit bit interleave).
template <typename CoordT, int K, bool ZORDER>
__global__ void __launch_bounds__(256)
kmap_search_kernel(const CoordT* __restrict__ grid_coord, const int64_t* __restrict__ batch,
                   const KmapSlot* __restrict__ table, int shift, uint32_t mask, int64_t n, int depth,
                   int order_id, int32_t* __restrict__ nbr, unsigned long long* __restrict__ tap_count) {
  constexpr int k3 = K * K * K, half = k3 / 2, r = K / 2;
  __shared__ unsigned int s_cnt[half];
  for (int i = threadIdx.x; i < half; i += blockDim.x) s_cnt[i] = 0u;
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = p < n;
  int x = 0, y = 0, z = 0;
  uint64_t bpart = 0;
  if (active) {
    x = (int)grid_coord[p * 3 + 0];
    y = (int)grid_coord[p * 3 + 1];
    z = (int)grid_coord[p * 3 + 2];
    bpart = (uint64_t)batch[p] << (3 * depth);
    nbr[(size_t)half * n + p] = (int32_t)p;  // centre tap
  }
  const int lim = 1 << depth;
  // per-axis key terms of the K offsets (z orders): bit i of the first curve axis -> 3 i + 2, second -> 3 i + 1, z -> 3 i;
  // z-trans swaps the roles of x and y.  ~0ull marks an offset outside the grid.
  uint64_t kx[K], ky[K], kz[K];
#pragma unroll
  for (int d = 0; ZORDER && d < K; ++d) {
    const int qx = x + d - r, qy = y + d - r, qz = z + d - r;
    kx[d] = (qx >= 0 && qx < lim) ? spread3((uint64_t)qx) << (order_id == 1 ? 1 : 2) : ~0ull;
    ky[d] = (qy >= 0 && qy < lim) ? spread3((uint64_t)qy) << (order_id == 1 ? 2 : 1) : ~0ull;
    kz[d] = (qz >= 0 && qz < lim) ? spread3((uint64_t)qz) : ~0ull;
  }
  const uint4* slots = reinterpret_cast<const uint4*>(table);
#pragma unroll(ZORDER ? half : 1)  // (the Hilbert variant keeps a rolled loop: its key is ~200 instructions per tap)
  for (int t = 0; t < half; ++t) {
    constexpr int KK = K * K;
    const int i = t / KK, j = (t / K) % K, l = t % K;  // constants after unrolling
    int32_t found = -1;
    if (active) {
      uint64_t key;
      bool inside;
      if (ZORDER) {
        inside = (kx[i] != ~0ull) && (ky[j] != ~0ull) && (kz[l] != ~0ull);
        key = bpart | kx[i] | ky[j] | kz[l];
      } else {
        const int qx = x + i - r, qy = y + j - r, qz = z + l - r;
        inside = qx >= 0 && qy >= 0 && qz >= 0 && qx < lim && qy < lim && qz < lim;
        key = inside ? (bpart | sfc_key(order_id, (uint32_t)qx, (uint32_t)qy, (uint32_t)qz, depth)) : 0ull;
      }
      if (inside) {
        uint32_t slot = kmap_hash(key, shift);
        while (true) {
          const uint4 e = __ldg(slots + slot);
          const uint64_t ek = ((uint64_t)e.y << 32) | e.x;
          if (ek == key) {
            found = (int32_t)e.z;
            break;
          }
          if (ek == kKmapEmpty) break;
          slot = (slot + 1u) & mask;
        }
      }
      nbr[(size_t)t * n + p] = found;
      if (found >= 0) nbr[(size_t)(k3 - 1 - t) * n + found] = (int32_t)p;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, found >= 0);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(&s_cnt[t], __popc(ballot));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const unsigned c = s_cnt[i];
    if (c) {
      atomicAdd(&tap_count[i], (unsigned long long)c);
      atomicAdd(&tap_count[k3 - 1 - i], (unsigned long long)c);
    }
  }
}

__global__ void kmap_center_count(unsigned long long* tap_count, int half, int64_t n) {
  if (threadIdx.x == 0 && blockIdx.x == 0) tap_count[half] = (unsigned long long)n;
}

// Pair lists for the gather-GEMM-scatter conv: along the sorted order j, for every tap t the active
// (in, out) pairs get consecutive rows starting at tap_base[t] (host-provided, padded to the GEMM tile).
struct PairRuns {
  const int32_t* nbr;
  const int64_t* order;
  const int64_t* tap_base;  // [k3] device copy
  int64_t n;
  int32_t* pair_in;   // [P_pad]  input row of every Y row
  int32_t* ypos;      // [k3][n]  Y row of (tap, output point) or -1
  int32_t* ypos_rank; // [n][32] (nullable, k3 <= 32) the same by RANK j along `order`: one 128-byte line per output
  int32_t* tile_first_rank;  // [P_pad / 256] (nullable) rank j of the first output of every 256-row product tile
  // "head" = pair is active: rank among active pairs of the tap = exclusive count of heads before j
  __device__ bool head(int t, int64_t j) const { return nbr[(size_t)t * n + order[j]] >= 0; }
  __device__ void emit(int t, int64_t j, uint32_t run, bool is_head) const {
    const int64_t p = order[j];
    if (is_head) {
      const int64_t pos = tap_base[t] + (int64_t)run;
      pair_in[pos] = nbr[(size_t)t * n + p];
      ypos[(size_t)t * n + p] = (int32_t)pos;
      if (ypos_rank) ypos_rank[j * 32 + t] = (int32_t)pos;
      if (tile_first_rank && (run & 255u) == 0u) tile_first_rank[pos >> 8] = (int32_t)j;  // tap bases are multiples of 256
    } else {
      ypos[(size_t)t * n + p] = -1;
      if (ypos_rank) ypos_rank[j * 32 + t] = -1;
    }
  }
  __device__ void finish(int, uint32_t) const {}
};

// The same pair lists for k = 3 in ONE pass over the ranks: a thread takes four ranks j and all 27 taps of each (PairRuns
// above walks one (tap, rank) per thread: 27 reads of order[j], 27 four-byte writes into every 128-byte line of ypos_rank and a
// memset of that table before).  Tiles of 1024 ranks are taken in ticket order; a tile publishes its 27 tap counts in one
// 128-byte status line and reads the lines of all earlier tiles in parallel (warp = tile, lane = tap): no chained look-back.
// ypos_rank rows leave through a per-warp shared-memory transposition as full 128-byte lines (columns 27..31 = -1).
constexpr int kPairThreads = 256;
// ranks per thread: 4 at the benchmark chunk's level 0 (293 CTAs = one wave at two CTAs per SM), fewer at the coarser levels,
// where a level is a few dozen CTAs and the per-thread serial work (27 taps x items) sets the time (ncu: 39 us for 9 CTAs)
inline int pair27_items(int64_t n) { return n > 150000 ? 4 : (n > 40000 ? 2 : 1); }
inline size_t pair27_workspace_bytes(int64_t n) { return 256 + (size_t)ceil_div64(n > 0 ? n : 1, kPairThreads) * 128; }

template <int kPairItems>
__global__ void __launch_bounds__(kPairThreads, 2)
pair27_kernel(const int32_t* __restrict__ nbr, const int64_t* __restrict__ order, const int64_t* __restrict__ tap_base,
              int64_t n, uint32_t* counter, uint32_t* status, int32_t* __restrict__ pair_in, int32_t* __restrict__ ypos,
              int32_t* __restrict__ ypos_rank, int32_t* __restrict__ tile_first_rank) {
  constexpr int K3 = 27, W = kPairThreads / 32, kPairTile = kPairThreads * kPairItems;
  __shared__ uint32_t s_cnt[kPairItems * W][K3];  // ballot totals per (item, warp), then their exclusive scan per tap
  __shared__ uint32_t s_part[W][32];
  __shared__ uint32_t s_before[32];
  __shared__ int32_t s_base[K3];
  __shared__ int32_t s_tr[W][32][33];
  __shared__ int s_ticket;
  if (threadIdx.x == 0) s_ticket = (int)atomicAdd(counter, 1u);
  if (threadIdx.x < K3) s_base[threadIdx.x] = (int32_t)tap_base[threadIdx.x];
  __syncthreads();
  const int tile = s_ticket, lane = (int)lane_id(), w = (int)(threadIdx.x >> 5);
  const int64_t base = (int64_t)tile * kPairTile;
  int32_t p[kPairItems];
  uint32_t m[kPairItems];
#pragma unroll
  for (int i = 0; i < kPairItems; ++i) {
    const int64_t j = base + i * kPairThreads + threadIdx.x;
    p[i] = j < n ? (int32_t)order[j] : -1;
    m[i] = 0u;
  }
#pragma unroll
  for (int t = 0; t < K3; ++t) {
#pragma unroll
    for (int i = 0; i < kPairItems; ++i)
      if (p[i] >= 0 && nbr[(size_t)t * n + p[i]] >= 0) m[i] |= 1u << t;
  }
#pragma unroll
  for (int t = 0; t < K3; ++t) {
#pragma unroll
    for (int i = 0; i < kPairItems; ++i) {
      const uint32_t b = __ballot_sync(0xffffffffu, (m[i] >> t) & 1u);
      if (lane == 0) s_cnt[i * W + w][t] = (uint32_t)__popc(b);
    }
  }
  __syncthreads();
  if (w == 0) {  // lane = tap: exclusive scan over the tile's (item, warp) groups in rank order, publish the tile's count
    uint32_t run = 0u;
    if (lane < K3) {
      for (int e = 0; e < kPairItems * W; ++e) {
        const uint32_t c = s_cnt[e][lane];
        s_cnt[e][lane] = run;
        run += c;
      }
    }
    st_volatile_u32(status + (size_t)tile * 32 + lane, kFlagAggregate | run);
  }
  uint32_t acc = 0u;
  for (int tt = w; tt < tile; tt += W) {
    uint32_t v;
    while (((v = ld_volatile_u32(status + (size_t)tt * 32 + lane)) & kFlagAggregate) == 0u) {}
    acc += v & kValueMask;
  }
  s_part[w][lane] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t sum = 0u;
#pragma unroll
    for (int ww = 0; ww < W; ++ww) sum += s_part[ww][threadIdx.x];
    s_before[threadIdx.x] = sum;
  }
  __syncthreads();
  const uint32_t lt = lanemask_lt();
#pragma unroll
  for (int i = 0; i < kPairItems; ++i) {
    const int64_t j = base + i * kPairThreads + threadIdx.x;
    const int32_t pi = p[i];
    const uint32_t mi = m[i];
#pragma unroll
    for (int t = 0; t < K3; ++t) {
      const bool bit = (mi >> t) & 1u;
      const uint32_t b = __ballot_sync(0xffffffffu, bit);
      const uint32_t run = s_before[t] + s_cnt[i * W + w][t] + (uint32_t)__popc(b & lt);
      const int32_t pos = bit ? s_base[t] + (int32_t)run : -1;
      if (pi >= 0) {
        ypos[(size_t)t * n + pi] = pos;
        if (bit) {
          pair_in[pos] = nbr[(size_t)t * n + pi];
          if (tile_first_rank && (run & 255u) == 0u) tile_first_rank[pos >> 8] = (int32_t)j;  // tap bases: multiples of 256
        }
      }
      s_tr[w][lane][t] = pos;
    }
    if (ypos_rank) {
#pragma unroll
      for (int t = K3; t < 32; ++t) s_tr[w][lane][t] = -1;
      __syncwarp();
      const int64_t j0 = base + i * kPairThreads + w * 32;
      for (int r = 0; r < 32 && j0 + r < n; ++r) ypos_rank[(j0 + r) * 32 + lane] = s_tr[w][r][lane];
      __syncwarp();
    }
  }
}

// The k_to^3 kernel map of a voxel set is a subset of the rows of its k_from^3 map (same voxels, k_to < k_from): tap
// (i, j, l) of the small window is tap (i + d, j + d, l + d) of the large one, d = (k_from - k_to) / 2.  Row copy instead of a
// second search (the level-0 3^3 map of the xCPE convs from the stem's 5^3 map).
__global__ void __launch_bounds__(256)
kmap_subset_kernel(const int32_t* __restrict__ from, const int64_t* __restrict__ count_from, int64_t n, int k_from, int k_to,
                   int32_t* __restrict__ to, int64_t* __restrict__ count_to) {
  const int t = blockIdx.y, d = (k_from - k_to) / 2;
  const int i = t / (k_to * k_to), j = (t / k_to) % k_to, l = t % k_to;
  const int tf = ((i + d) * k_from + (j + d)) * k_from + (l + d);
  const int32_t* src = from + (size_t)tf * n;
  int32_t* dst = to + (size_t)t * n;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) dst[p] = src[p];
  if (blockIdx.x == 0 && threadIdx.x == 0) count_to[t] = count_from[tf];
}

// Order of the product tiles for the fused conv: position of tile t = number of tiles with a smaller (first rank, index)
// key.  T is a few thousand (7.6 k at the benchmark chunk): T^2 comparisons from shared-memory chunks, no sort.  Eight
// lanes share one tile (each takes every eighth candidate; a warp reads eight distinct words per step, broadcast to its
// four tiles), so T / 32 CTAs are in flight instead of T / 256 (30 CTAs at the benchmark chunk: 111 us for 58 M compares).
constexpr int kRankLanes = 8;
constexpr int kRankTilesPerCta = 256 / kRankLanes;
__global__ void __launch_bounds__(256) tile_rank_kernel(const int32_t* __restrict__ first_rank, int tiles,
                                                        int32_t* __restrict__ tile_order, int32_t* __restrict__ tile_pos) {
  __shared__ int32_t s_fr[1024];
  const int t = blockIdx.x * kRankTilesPerCta + (int)(threadIdx.x / kRankLanes), part = (int)(threadIdx.x % kRankLanes);
  const int32_t mine = t < tiles ? first_rank[t] : 0;
  int pos = 0;
  for (int base = 0; base < tiles; base += 1024) {
    const int m = min(1024, tiles - base);
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += 256) s_fr[i] = first_rank[base + i];
    __syncthreads();
#pragma unroll 8
    for (int i = part; i < m; i += kRankLanes) {
      const int32_t v = s_fr[i];
      pos += (v < mine || (v == mine && base + i < t)) ? 1 : 0;
    }
  }
#pragma unroll
  for (int o = kRankLanes / 2; o; o >>= 1) pos += __shfl_xor_sync(0xffffffffu, pos, o);
  if (part == 0 && t < tiles) {
    tile_pos[t] = pos;
    tile_order[pos] = t;
  }
}

}  // namespace ss

extern "C" {

size_t ss_kmap_workspace_bytes(int64_t n, int k) {
  if (n < 0 || k < 1) return 0;
  const size_t scan = ss::runs_workspace_bytes(n, k * k * k), pair27 = ss::pair27_workspace_bytes(n);
  return ss::kmap_table_bytes(n) + ss::align_up(scan > pair27 ? scan : pair27, 256) + 512;
}

int ss_kmap_build(const void* grid_coord, int coord_is_int32, const int64_t* batch, const int64_t* code_row,
                  const int64_t* order_row, int64_t n, int depth, int order_id, int k, int32_t* nbr,
                  int64_t* tap_count_dev, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || n > 0x7fffffff || depth < 1 || depth > 16 || order_id < 0 || order_id > 3 || (k != 3 && k != 5) ||
      !tap_count_dev)
    return SS_BAD_ARGS;
  const int k3 = k * k * k;
  SS_CUDA(cudaMemsetAsync(tap_count_dev, 0, (size_t)k3 * 8, stream));
  if (n == 0) return SS_OK;
  if (!grid_coord || !batch || !code_row || !order_row || !nbr || !workspace) return SS_BAD_ARGS;
  if (workspace_bytes < ss_kmap_workspace_bytes(n, k) - 512) return SS_BAD_ARGS;
  char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  ss::KmapSlot* table = (ss::KmapSlot*)ws;
  const int64_t slots = ss::kmap_table_slots(n);
  int log2s = 0;
  while ((1ll << log2s) < slots) ++log2s;
  const int shift = 64 - log2s;
  const uint32_t mask = (uint32_t)(slots - 1);
  const int blocks = ss::ceil_div((int)n, 256);
  SS_CUDA(cudaMemsetAsync(table, 0xff, (size_t)slots * sizeof(ss::KmapSlot), stream));
  ss::kmap_insert_kernel<<<blocks, 256, 0, stream>>>(code_row, n, table, shift, mask);
  SS_CHECK_LAUNCH();
  // mirrored half (taps > centre) is only written where a neighbour exists
  SS_CUDA(cudaMemsetAsync(nbr + (size_t)(k3 / 2 + 1) * n, 0xff, (size_t)(k3 / 2) * n * 4, stream));
#define SS_KSEARCH_(T, KK, Z)                                                                                           \
  ss::kmap_search_kernel<T, KK, Z><<<blocks, 256, 0, stream>>>((const T*)grid_coord, batch, table, shift, mask, n, depth,  \
                                                               order_id, nbr, (unsigned long long*)tap_count_dev)
#define SS_KSEARCH_K_(T, Z)        \
  do {                             \
    if (k == 3) SS_KSEARCH_(T, 3, Z); \
    else SS_KSEARCH_(T, 5, Z);     \
  } while (0)
  if (order_id < 2) {
    if (coord_is_int32) SS_KSEARCH_K_(int, true);
    else SS_KSEARCH_K_(long long, true);
  } else {
    if (coord_is_int32) SS_KSEARCH_K_(int, false);
    else SS_KSEARCH_K_(long long, false);
  }
#undef SS_KSEARCH_K_
#undef SS_KSEARCH_
  SS_CHECK_LAUNCH();
  ss::kmap_center_count<<<1, 32, 0, stream>>>((unsigned long long*)tap_count_dev, k3 / 2, n);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

int ss_kmap_subset(const int32_t* nbr_from, const int64_t* count_from, int64_t n, int k_from, int k_to, int32_t* nbr_to,
                   int64_t* count_to, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || k_to < 1 || k_from <= k_to || (k_from - k_to) % 2 != 0 || k_from > 9 || !count_from || !count_to)
    return SS_BAD_ARGS;
  if (n > 0 && (!nbr_from || !nbr_to)) return SS_BAD_ARGS;
  const int blocks = n > 0 ? (int)ss::imin64(ss::ceil_div64(n, 256), 2 * ss::kNumSMs) : 1;
  ss::kmap_subset_kernel<<<dim3(blocks, k_to * k_to * k_to), 256, 0, stream>>>(nbr_from, count_from, n, k_from, k_to, nbr_to,
                                                                            count_to);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

int ss_kmap_pairs(const int32_t* nbr, const int64_t* order_row, int64_t n, int k, const int64_t* tap_base_dev,
                  int64_t p_pad, int32_t* pair_in, int32_t* ypos, int32_t* ypos_rank, int32_t* tile_first_rank,
                  int32_t* tile_order, int32_t* tile_pos, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || (k != 3 && k != 5) || p_pad < 0) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!nbr || !order_row || !tap_base_dev || !pair_in || !ypos || !workspace) return SS_BAD_ARGS;
  if ((ypos_rank && k != 3) || (tile_first_rank && p_pad % 256 != 0)) return SS_BAD_ARGS;
  if ((tile_order || tile_pos) && !(tile_first_rank && tile_order && tile_pos)) return SS_BAD_ARGS;
  // padding rows of every tap segment gather row 0 (their products are never read back)
  SS_CUDA(cudaMemsetAsync(pair_in, 0, (size_t)p_pad * 4, stream));
  if (workspace_bytes < ss_kmap_workspace_bytes(n, k) - 512) return SS_BAD_ARGS;
  char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  ws += ss::kmap_table_bytes(n);
  int rc = SS_OK;
  if (k == 3) {
    SS_CUDA(cudaMemsetAsync(ws, 0, ss::pair27_workspace_bytes(n), stream));
    const int items = ss::pair27_items(n);
    const int ptiles = (int)ss::ceil_div64(n, (int64_t)ss::kPairThreads * items);
#define SS_PAIR27_(I)                                                                                                  \
  ss::pair27_kernel<I><<<ptiles, ss::kPairThreads, 0, stream>>>(nbr, order_row, tap_base_dev, n, (uint32_t*)ws,          \
                                                               (uint32_t*)(ws + 256), pair_in, ypos, ypos_rank,         \
                                                               tile_first_rank)
    if (items == 4) SS_PAIR27_(4);
    else if (items == 2) SS_PAIR27_(2);
    else SS_PAIR27_(1);
#undef SS_PAIR27_
    SS_CHECK_LAUNCH();
  } else {
    ss::PairRuns f{nbr, order_row, tap_base_dev, n, pair_in, ypos, ypos_rank, tile_first_rank};
    rc = ss::runs_launch(f, n, ws, nullptr, stream, k * k * k);
  }
  if (rc != SS_OK || !tile_order || p_pad == 0) return rc;
  const int tiles = (int)(p_pad / 256);
  ss::tile_rank_kernel<<<ss::ceil_div(tiles, ss::kRankTilesPerCta), 256, 0, stream>>>(tile_first_rank, tiles, tile_order, tile_pos);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // extern "C"
