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
// the mirrored tap is written at the neighbour.
template <typename CoordT>
__global__ void __launch_bounds__(256)
kmap_search_kernel(const CoordT* __restrict__ grid_coord, const int64_t* __restrict__ batch,
                   const KmapSlot* __restrict__ table, int shift, uint32_t mask, int64_t n, int depth,
                   int order_id, int k, int32_t* __restrict__ nbr, unsigned long long* __restrict__ tap_count) {
  extern __shared__ unsigned int s_cnt[];  // [k^3 / 2]
  const int k3 = k * k * k, half = k3 / 2, r = k / 2;
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
  const uint4* slots = reinterpret_cast<const uint4*>(table);
  for (int t = 0; t < half; ++t) {
    const int dx = t / (k * k) - r, dy = (t / k) % k - r, dz = t % k - r;
    int32_t found = -1;
    if (active) {
      const int qx = x + dx, qy = y + dy, qz = z + dz;
      if (qx >= 0 && qy >= 0 && qz >= 0 && qx < lim && qy < lim && qz < lim) {
        const uint64_t key = bpart | sfc_key(order_id, (uint32_t)qx, (uint32_t)qy, (uint32_t)qz, depth);
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
  return ss::kmap_table_bytes(n) + ss::align_up(ss::runs_workspace_bytes(n, k * k * k), 256) + 512;
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
  const size_t smem = (size_t)(k3 / 2) * 4;
  if (coord_is_int32)
    ss::kmap_search_kernel<int><<<blocks, 256, smem, stream>>>((const int*)grid_coord, batch, table, shift, mask, n, depth,
                                                              order_id, k, nbr, (unsigned long long*)tap_count_dev);
  else
    ss::kmap_search_kernel<long long><<<blocks, 256, smem, stream>>>((const long long*)grid_coord, batch, table, shift,
                                                                    mask, n, depth, order_id, k, nbr,
                                                                    (unsigned long long*)tap_count_dev);
  SS_CHECK_LAUNCH();
  ss::kmap_center_count<<<1, 32, 0, stream>>>((unsigned long long*)tap_count_dev, k3 / 2, n);
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
  // columns 27..31 of the rank-major table stay -1
  if (ypos_rank) SS_CUDA(cudaMemsetAsync(ypos_rank, 0xff, (size_t)n * 32 * 4, stream));
  // padding rows of every tap segment gather row 0 (their products are never read back)
  SS_CUDA(cudaMemsetAsync(pair_in, 0, (size_t)p_pad * 4, stream));
  if (workspace_bytes < ss_kmap_workspace_bytes(n, k) - 512) return SS_BAD_ARGS;
  char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  ws += ss::kmap_table_bytes(n);
  ss::PairRuns f{nbr, order_row, tap_base_dev, n, pair_in, ypos, ypos_rank, tile_first_rank};
  int rc = ss::runs_launch(f, n, ws, nullptr, stream, k * k * k);
  if (rc != SS_OK || !tile_order || p_pad == 0) return rc;
  const int tiles = (int)(p_pad / 256);
  ss::tile_rank_kernel<<<ss::ceil_div(tiles, ss::kRankTilesPerCta), 256, 0, stream>>>(tile_first_rank, tiles, tile_order, tile_pos);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // extern "C"
