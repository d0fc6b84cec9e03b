// Submanifold convolution, stage 1 (gather-GEMM), third generation: CTA PAIRS (tcgen05 cta_group::2).
//
// Replaces (reference): spconv.SubMConv3d forward for the xCPE 3^3 convs
// (point_transformer_v3m1_base.py:277-284; fp32 in the reference, bf16 x bf16 -> fp32 here).
//
//   prod[r, :] = X[pair_in[r], :] @ W_tap(r)^T        r < p_pad, taps contiguous and padded to 256 rows
//
// Same work items as conv_gemm2.cu (256 product rows x 256 output columns, slab-fastest), but a CTA pair shares one:
// each CTA gathers its own 128 rows of A (cp.async, published by the copies themselves) and stages HALF of the 256
// weight rows (TMA); the leader issues tcgen05.mma.cta_group::2 with M = 256.  One 128 x 256 fp32 accumulator per CTA
// leaves the other 256 TMEM columns for a second buffer, so draining tile i (TMEM -> bf16 -> coalesced stores) runs
// under the MMAs of tile i + 1 instead of stalling the tensor pipe (10 % of the time at C = 768, more at C = 256 / 512
// where a tile has fewer K chunks).  Operand bytes staged per FLOP are unchanged (32 KB per 4.2 MFLOP per CTA).
//
// A and W live in SEPARATE rings: the gathers have ~3 us of latency under load and the kernel's rate is (gathered bytes in
// flight) / latency, so A gets 7 slots of 16 KB and the TMA-fed W 4 (the pair halves W's footprint per CTA, which is what
// makes room).  The leader's a_full barrier of a slot collects its own 128 gather threads
// (cp.async.mbarrier.arrive.noinc) and one remote arrive from the peer's relay warp, which waits for the peer's own 128
// gather arrivals; w_full counts the TMA bytes of BOTH weight halves.  Slot release and accumulator hand-over are
// multicast commits.
//
// Measured at the dec0 shape (1.94 M pairs, C = 768): 1.91 ms = 1202 TFLOP/s useful against 2.06 ms (1115) for
// conv_gemm2.cu, at C = 512: 0.219 against 0.252 ms; bit-identical products.  What made the difference: the peer's relay
// must not be one thread issuing one release-arrive per slot in sequence (a remote arrive has microseconds of latency:
// 670 TFLOP/s, and 740 even with the gathers switched off); with one relay lane per slot and relaxed arrives the
// hand-overs overlap.
//
// 15 warps per CTA: 0-7 epilogue (row quarter x column half), 8-11 A gather, 12 W producer (TMA, one lane),
// 13 MMA issuer (leader) + TMEM allocator, 14 relay (peer).
#include "pair_common.cuh"
#include "conv_reduce.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

constexpr int kG3Threads = 480;
#ifndef SS_G3_RED_WARPS
#define SS_G3_RED_WARPS 12
#endif
constexpr int kG3RedWarps = SS_G3_RED_WARPS;      // reducer warps of the fused kernel (warps 16..27; warp 15 idles)
constexpr int kG3FusedThreads = 512 + 32 * kG3RedWarps;  // 896 threads x 72 registers (re-dividing the registers between the
// roles with setmaxnreg -- 88 / 48 / 40 / 80 -- measured slower: the gather warps spill; profiles/r2_conv.md)
constexpr int kG3BK = 64;
constexpr int kG3SA = 7;  // A ring (gathered rows: ~3 us of latency under load, so as many bytes in flight as fit)
constexpr int kG3SW = 4;  // W ring (TMA: a quarter of that latency)
constexpr int kG3TileM = 256;
constexpr int kG3BN = 256;

struct Gemm3Smem {
  static constexpr int kA = 128 * kG3BK * 2;  // this CTA's 128 gathered rows
  static constexpr int kB = 128 * kG3BK * 2;  // this CTA's half of the weight rows
  static constexpr int kOffW = kG3SA * kA;
  static constexpr int kOffEpi = kOffW + kG3SW * kB;
  static constexpr int kOffBar = kOffEpi + 8 * 2048;
  static constexpr int kTotal = kOffBar + 512 + 1024;
};

// Arguments of the fused gather-sum stage (J > 0).
struct ConvFuse {
  const int32_t* tile_order;  // [tiles] processing order of the 256-row product tiles: by the rank of their first output
  const int32_t* tile_pos;    // [tiles] inverse of tile_order
  int* tile_flag;             // [tiles + 2] zeroed by the caller; += 1 per epilogue warp and item of the tile; word
                              // `tiles + 1` is the reducers' FRONT: the furthest tile (in processing order, + 1) a reducer
                              // has reached (posted every 16th voxel) or has had to wait for
  int n_tiles;
  int flag_target;            // 16 * n_slabs: all of the tile's products are in memory
  int lag_tiles;              // the GEMM starts tile k (in processing order) only while k < front + lag_tiles
  const int64_t* order_row;   // [n] the serialized order the pair lists were built along
  const int32_t* ypos_rank;   // [n][32] product rows of the voxel at rank r (ss_kmap_pairs)
  const float* bias;
  const float* res;
  const float* g0;
  const float* b0;
  const float* g1;
  const float* b1;
  float eps;
  int64_t n;
  int k3;
  float* res_out;
  __nv_bfloat16* norm_out;
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add1(int* p) {
  asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
#ifndef SS_G3_LAG_MB
#define SS_G3_LAG_MB 32
#endif
// J = 0: products only.  J = 1..4 (= ceil(cout / 256)): FUSED gather-sum + LN + residual + LN.  Twelve more warps per CTA
// walk the output voxels along the serialized order the pair rows of every tap follow (ranks w, w + W, ..: W = all reducer
// warps of the grid), wait until the <= 27 product tiles the voxel needs have been stored (per-tile counters: release
// reductions by the epilogue warps, acquire loads here, prefetched one voxel ahead), sum the rows and finish the Block's
// LN(cpe) + residual + LN(norm1) (conv_reduce.cuh).  The GEMM takes the tiles in the order of their first output's rank
// (tile_order, exact: ss_kmap_pairs) instead of tap by tap and stays at most lag_tiles (~32 MB of products) ahead of the
// reducers' front, so the bf16 products are read back from L2 instead of from HBM (dec0: DRAM reads 7.0 -> 3.8 GB) and
// most of the gather-sum pass (1.03 ms at dec0) runs under the MMAs (2.90 -> 2.62 ms).
// Deadlock freedom: reducers wait only for tiles; a tile a reducer waits for is below the front that reducer has
// posted, so the GEMM may start it, and a CTA pair's earlier items are earlier in the order; every CTA of the
// persistent grid is resident (1 per SM) or becomes resident without help from this kernel.  All spin waits are bounded
// (4 s, trap).
template <int J>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(J > 0 ? kG3FusedThreads : kG3Threads, 1)
gather_gemm_pair_kernel(const __nv_bfloat16* __restrict__ X, const int32_t* __restrict__ pair_in,
                        const __grid_constant__ CUtensorMap tmap_w, const int32_t* __restrict__ tile_tap, int cin, int cout,
                        int n_slabs, int64_t n_items, __nv_bfloat16* __restrict__ prod, const ConvFuse fz) {
  constexpr bool FUSE = J > 0;
  using S = Gemm3Smem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* a_full = (uint64_t*)(smem + S::kOffBar);  // [SA] leader: its 128 gathers + the peer's relay
  uint64_t* a_done = a_full + kG3SA;                  // [SA] peer: its 128 gathers
  uint64_t* a_empty = a_done + kG3SA;                 // [SA] one multicast commit
  uint64_t* w_full = a_empty + kG3SA;                 // [SW] leader: expect_tx arrive, bytes of both halves
  uint64_t* w_empty = w_full + kG3SW;                 // [SW] one multicast commit
  uint64_t* acc_full = w_empty + kG3SW;               // [2]
  uint64_t* acc_empty = acc_full + 2;                   // [2] leader's copy: 8 warps x 2 CTAs
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = pair::cta_rank();
  const bool leader = rank == 0;
  const int64_t pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int nk = (cin + kG3BK - 1) / kG3BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kG3SA; ++s) {
      tc::mbar_init(&a_full[s], 128 + 1);
      tc::mbar_init(&a_done[s], 128);
      tc::mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kG3SW; ++s) {
      tc::mbar_init(&w_full[s], 1);
      tc::mbar_init(&w_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&acc_full[b], 1);
      tc::mbar_init(&acc_empty[b], 16);
    }
    tc::mbar_fence_init();
  }
  if (warp == 12 && lane == 0) tc::tma_prefetch_desc(&tmap_w);
  if (warp == 13) pair::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  pair::cluster_sync();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8 && warp < 12) {
    // ------------------------------------------------------------------ A gather: 128 threads, this CTA's 128 rows
    const int tid = threadIdx.x - 256;      // 0..127
    const int sub = tid >> 3, c = tid & 7;  // 8 lanes cover one 128-byte row segment
    int64_t g = 0;
    for (int64_t item = pair_id; item < n_items; item += n_pairs) {
      const int64_t tile = FUSE ? fz.tile_order[item / n_slabs] : item / n_slabs;
      if constexpr (FUSE) {
        // stay within lag_tiles of what the reducers have asked for: the products are read back while still in L2.
        // (Every tile a reducer waits for is below its own request, so this wait cannot starve the reducers.)
        const int k = (int)(item / n_slabs);
        const int* front = fz.tile_flag + fz.n_tiles + 1;
        if (lane == 0 && k >= ld_relaxed_gpu(front) + fz.lag_tiles) {
          const uint64_t t0 = tc::global_timer_ns();
          while (k >= ld_relaxed_gpu(front) + fz.lag_tiles) {
            __nanosleep(500);
            if (tc::global_timer_ns() - t0 > 4000000000ull) __trap();  // 4 s: a protocol bug must trap, never hang
          }
        }
        __syncwarp();
      }
      int32_t rows[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rows[i] = pair_in[tile * kG3TileM + rank * 128 + i * 16 + sub];
      for (int kc = 0; kc < nk; ++kc, ++g) {
        const int s = (int)(g % kG3SA);
        // one poller per warp: 128 threads spinning on the barrier delay the (remote, multicast) arrival they wait for
        if (lane == 0) tc::mbar_wait(&a_empty[s], (uint32_t)((g / kG3SA) & 1) ^ 1);
        __syncwarp();
        const uint32_t a_base = tc::smem_u32(smem + s * S::kA);
        const int k0 = kc * kG3BK;
        if (c < (min(kG3BK, cin - k0) >> 3)) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 16 + sub;
            tc::cp_async16(a_base + tc::sw128_offset(r, c), X + (size_t)rows[i] * cin + k0 + c * 8);
          }
        }
        tc::cp_async_mbar_arrive_noinc(leader ? &a_full[s] : &a_done[s]);
      }
    }
  } else if (warp >= 12 && warp < 16) {
    // W producer, MMA issuer, relay (and an idle warp in the fused kernel): one warpgroup, one register budget
    if (warp == 12) {
    // ------------------------------------------------------------------ W producer (TMA, one lane, both CTAs)
    if (lane == 0) {
      int64_t g = 0;
      for (int64_t item = pair_id; item < n_items; item += n_pairs) {
        const int64_t tidx = item / n_slabs;
        const int64_t tile = FUSE ? fz.tile_order[tidx] : tidx;
        const int n0 = (int)(item - tidx * n_slabs) * kG3BN;
        const int tap = tile_tap[tile];
        for (int kc = 0; kc < nk; ++kc, ++g) {
          const int s = (int)(g % kG3SW);
          tc::mbar_wait(&w_empty[s], (uint32_t)((g / kG3SW) & 1) ^ 1);
          if (leader) tc::mbar_arrive_expect_tx(&w_full[s], 2 * S::kB);  // both weight halves
          pair::tma_load_2d(tc::smem_u32(smem + S::kOffW + s * S::kB), &tmap_w, kc * kG3BK,
                            tap * cout + n0 + (int)rank * 128, &w_full[s]);
        }
      }
    }
    } else if (warp == 14) {
    // ------------------------------------------------------------------ relay (peer): "my 128 rows of the stage have landed"
    // one lane per A slot: a remote arrive has microseconds of latency, so the slots' hand-overs must overlap
    if (!leader && lane < kG3SA) {
      int64_t my_items = 0;
      for (int64_t item = pair_id; item < n_items; item += n_pairs) ++my_items;
      const int64_t total = my_items * nk;
      for (int64_t g = lane; g < total; g += kG3SA) {
        tc::mbar_wait(&a_done[lane], (uint32_t)((g / kG3SA) & 1));
        pair::mbar_arrive_cta_relaxed(&a_full[lane], 0);
      }
    }
    } else if (warp == 13) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(kG3TileM, kG3BN);
      const uint64_t d_base = tc::umma_desc_sw128(0);
      const uint32_t s0 = tc::smem_u32(smem);
      int64_t g = 0;
      int it = 0;
      for (int64_t item = pair_id; item < n_items; item += n_pairs, ++it) {
        const int b = it & 1;
        tc::mbar_wait(&acc_empty[b], (uint32_t)((it >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        for (int kc = 0; kc < nk; ++kc, ++g) {
          const int sa = (int)(g % kG3SA), sw = (int)(g % kG3SW);
          tc::mbar_wait(&w_full[sw], (uint32_t)((g / kG3SW) & 1));
          tc::mbar_wait(&a_full[sa], (uint32_t)((g / kG3SA) & 1));
          tc::tc_fence_after();
          const uint32_t a0 = (s0 + sa * S::kA) >> 4;
          const uint32_t b0 = (s0 + S::kOffW + sw * S::kB) >> 4;
          const int ksteps = min(kG3BK, cin - kc * kG3BK) >> 4;
          for (int k = 0; k < ksteps; ++k)
            pair::umma_bf16_elect(tmem_base + b * kG3BN, d_base | (uint64_t)((a0 + 2 * k) & 0x3fff),
                                  d_base | (uint64_t)((b0 + 2 * k) & 0x3fff), idesc, (kc | k) ? 1u : 0u);
          pair::umma_commit_elect(&a_empty[sa]);
          pair::umma_commit_elect(&w_empty[sw]);
        }
        pair::umma_commit_elect(&acc_full[b]);
      }
    }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ epilogue warps 0..7: (row quarter, column half)
    uint8_t* stg = smem + S::kOffEpi + warp * 2048;
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int it = 0;
    for (int64_t item = pair_id; item < n_items; item += n_pairs, ++it) {
      const int64_t tidx = item / n_slabs;
      const int64_t tile = FUSE ? fz.tile_order[tidx] : tidx;
      const int n0 = (int)(item - tidx * n_slabs) * kG3BN;
      const int b = it & 1;
      tc::mbar_wait(&acc_full[b], (uint32_t)((it >> 1) & 1));
      tc::tc_fence_after();
      __nv_bfloat16* obase = prod + ((size_t)tile * kG3TileM + rank * 128 + quarter * 32) * cout + n0 + half * 128;
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        if (n0 + half * 128 + j * 32 >= cout) break;
        uint32_t v[32];
        tc::tmem_ld32(t_lane + b * kG3BN + half * 128 + j * 32, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 o;
          o.x = tc::pack_bf16(__uint_as_float(v[u * 8 + 0]), __uint_as_float(v[u * 8 + 1]));
          o.y = tc::pack_bf16(__uint_as_float(v[u * 8 + 2]), __uint_as_float(v[u * 8 + 3]));
          o.z = tc::pack_bf16(__uint_as_float(v[u * 8 + 4]), __uint_as_float(v[u * 8 + 5]));
          o.w = tc::pack_bf16(__uint_as_float(v[u * 8 + 6]), __uint_as_float(v[u * 8 + 7]));
          *reinterpret_cast<uint4*>(stg + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4)) = o;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + (lane >> 2), cc = lane & 3;
          const uint4 o = *reinterpret_cast<const uint4*>(stg + r * 64 + ((cc ^ ((r >> 1) & 3)) << 4));
          *reinterpret_cast<uint4*>(obase + (size_t)r * cout + j * 32 + cc * 8) = o;
        }
        __syncwarp();
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        pair::mbar_arrive_cta(&acc_empty[b], 0);
        if (FUSE) red_release_gpu_add1(fz.tile_flag + tile);  // this warp's share of the tile is in memory
      }
    }
  } else if (FUSE && warp >= 16) {
    // ------------------------------------------------------------------ reducers: one warp per output voxel, in rank order
    constexpr int JJ = FUSE ? J : 1;
    constexpr int RR = JJ >= 3 ? 2 : 4;  // product rows in flight per lane (80 registers per thread)
    const int64_t n = fz.n;
    const float invC = 1.f / (float)cout;
    int* front = fz.tile_flag + fz.n_tiles + 1;
    // Ranks w, w + W, w + 2 W, ... (W = all reducer warps of the grid): the warps move through the ranks together (the
    // leaders are held back by the tiles the GEMM has not produced yet), so the live part of the product buffer is the GEMM's
    // lead (lag_tiles) plus ~W outputs.  The index chain rank -> (voxel id, row of product
    // positions) -> tile flags is fetched two / one voxel ahead, so a voxel costs the product loads, the residual load
    // and the stores.
    const int64_t W = (int64_t)gridDim.x * kG3RedWarps;
    int64_t r = (int64_t)blockIdx.x * kG3RedWarps + (warp - 16);
    int64_t p_a = r < n ? fz.order_row[r] : 0, p_b = r + W < n ? fz.order_row[r + W] : 0;
    int32_t pos_a = r < n ? fz.ypos_rank[(size_t)r * 32 + lane] : -1;
    int32_t pos_b = r + W < n ? fz.ypos_rank[(size_t)(r + W) * 32 + lane] : -1;
    int flag_a = pos_a >= 0 ? ld_acquire_gpu(fz.tile_flag + (pos_a >> 8)) : 0x7fffffff;
    for (int it = 0; r < n; r += W, ++it) {
      const int64_t p = p_a;
      const int32_t mypos = pos_a;
      int flag = flag_a;
      // prefetch: flags of the next voxel, positions / voxel id of the one after it
      p_a = p_b;
      pos_a = pos_b;
      flag_a = pos_a >= 0 ? ld_acquire_gpu(fz.tile_flag + (pos_a >> 8)) : 0x7fffffff;
      p_b = r + 2 * W < n ? fz.order_row[r + 2 * W] : 0;
      pos_b = r + 2 * W < n ? fz.ypos_rank[(size_t)(r + 2 * W) * 32 + lane] : -1;
      // every 16th voxel: tell the GEMM how far this warp has come (the centre tap's tile: every voxel has one), so that
      // it keeps lag_tiles ahead of the reducers without waiting for one of them to run dry.  (Rare on purpose: same-
      // address atomics serialise in one L2 slice at ~5 ns each; one per voxel made them the kernel's bottleneck.)
      if ((it & 15) == 0 && lane == fz.k3 / 2 && mypos >= 0) atomicMax(front, fz.tile_pos[mypos >> 8] + 1);
      if (flag < fz.flag_target) {  // (lanes without a pair carry INT_MAX)
        const int t = mypos >> 8;
        const int* f = fz.tile_flag + t;
        if (ld_relaxed_gpu(f) < fz.flag_target) {
          atomicMax(front, fz.tile_pos[t] + 1);  // the GEMM may be holding this tile back: let it run up to it
          const uint64_t t0 = tc::global_timer_ns();
          while (ld_acquire_gpu(f) < fz.flag_target) {
            __nanosleep(200);
            if (tc::global_timer_ns() - t0 > 4000000000ull) __trap();  // 4 s
          }
        }
      }
      // every lane has seen its tile's release counter complete with an acquire load; the warp barrier extends that to
      // the rows of the other lanes' taps
      __syncwarp();
      float acc[JJ][8];
      conv_acc_init<JJ>(fz.bias, lane, cout, acc);
      conv_gather_sum<JJ, true, RR>(prod, mypos, lane, cout, acc);
      conv_ln_res_ln_store<JJ>(acc, p, lane, cout, invC, fz.eps, fz.res, fz.g0, fz.b0, fz.g1, fz.b1, fz.res_out,
                               fz.norm_out);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  pair::cluster_sync();
  if (warp == 13) {
    tc::tc_fence_after();
    pair::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace ss

namespace ss {

template <int J>
static int launch_gather_gemm_pair(const void* in_bf16, const int32_t* pair_in, const void* w_bf16, const int32_t* tile_tap,
                                   int64_t p_pad, int k3, int cin, int cout, void* prod_bf16, const ConvFuse& fz,
                                   cudaStream_t stream) {
  const int64_t tiles = p_pad / kG3TileM;
  // W viewed as one [k3 * cout, cin] K-major matrix, boxes of 128 rows (one CTA's half of a 256-column slab); a box that
  // runs past a tap's last column reads the next tap's rows (or zeros past the end): those columns are never stored
  CUtensorMap tmap;
  int rc = make_tmap_bf16_2d(&tmap, w_bf16, (uint64_t)k3 * cout, (uint64_t)cin, 128, kG3BK);
  if (rc) return rc;
  auto kern = gather_gemm_pair_kernel<J>;
  SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm3Smem::kTotal));
  const int n_slabs = (cout + kG3BN - 1) / kG3BN;
  const int64_t n_items = tiles * n_slabs;
  const int pairs = (int)imin64(n_items, kNumSMs / 2);
  ConvFuse f = fz;
  f.flag_target = 16 * n_slabs;
  f.n_tiles = (int)tiles;
  // ~SS_G3_LAG_MB MB of products between the GEMM front and the furthest tile asked for; never fewer tiles than keep every CTA pair
  // two items deep
  const int64_t lag_bytes = (int64_t)SS_G3_LAG_MB * 1024 * 1024 / (512ll * cout), lag_min = 2 * (kNumSMs / 2) / n_slabs + 8;
  f.lag_tiles = (int)(lag_bytes > lag_min ? lag_bytes : lag_min);
  if (J > 0) SS_CUDA(cudaMemsetAsync(f.tile_flag, 0, (size_t)(tiles + 2) * sizeof(int), stream));
  kern<<<2 * pairs, J > 0 ? kG3FusedThreads : kG3Threads, Gemm3Smem::kTotal, stream>>>(
      (const __nv_bfloat16*)in_bf16, pair_in, tmap, tile_tap, cin, cout, n_slabs, n_items, (__nv_bfloat16*)prod_bf16, f);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // namespace ss

extern "C" int ss_subm_conv_gemm_pair(const void* in_bf16, const int32_t* pair_in, const void* w_bf16,
                                      const int32_t* tile_tap, int64_t p_pad, int k3, int cin, int cout, void* prod_bf16,
                                      void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (k3 < 1 || p_pad < 0 || p_pad % ss::kG3TileM != 0 || cin < 16 || cin % 16 != 0 || cout < 256 || cout % 32 != 0)
    return SS_BAD_ARGS;
  if (p_pad == 0) return SS_OK;
  if (!in_bf16 || !pair_in || !w_bf16 || !tile_tap || !prod_bf16) return SS_BAD_ARGS;
  if (((uintptr_t)in_bf16 | (uintptr_t)w_bf16 | (uintptr_t)prod_bf16) % 16 != 0) return SS_BAD_ARGS;
  return ss::launch_gather_gemm_pair<0>(in_bf16, pair_in, w_bf16, tile_tap, p_pad, k3, cin, cout, prod_bf16, ss::ConvFuse{},
                                        stream);
}

// The xCPE conv of a Block in ONE launch: gather-GEMM + gather-sum + LN(cpe) + residual + LN(norm1)
// (ss_subm_conv_gemm_pair followed by ss_subm_conv_reduce_add_ln, bit-identical to that pair of calls).
// tile_order: a permutation of the p_pad / 256 product tiles (ops.kmap_pairs: by the rank of the tile's first output);
// tile_flags: p_pad / 256 ints of scratch; order_row: the serialized order the pair lists were built along.
extern "C" int ss_subm_conv_fused_add_ln(const void* in_bf16, const int32_t* pair_in, const void* w_bf16,
                                         const int32_t* tile_tap, const int32_t* tile_order, const int32_t* tile_pos,
                                         int64_t p_pad, int k3, int cin, int cout, void* prod_bf16, int32_t* tile_flags,
                                         const int64_t* order_row, const int32_t* ypos_rank, const float* bias,
                                         const float* res, const float* g0,
                                         const float* b0, const float* g1, const float* b1, float eps, int64_t n,
                                         float* res_out, void* norm_out_bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (k3 < 1 || k3 > 32 || p_pad < 0 || p_pad % ss::kG3TileM != 0 || cin < 16 || cin % 16 != 0 || cout < 256 ||
      cout % 32 != 0 || cout > 1024 || n < 0)
    return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (p_pad == 0) return SS_BAD_ARGS;  // every voxel has its centre tap
  if (!in_bf16 || !pair_in || !w_bf16 || !tile_tap || !tile_order || !tile_pos || !prod_bf16 || !tile_flags || !order_row ||
      !ypos_rank ||
      !res || !g0 || !b0 || !g1 || !b1 || !res_out || !norm_out_bf16)
    return SS_BAD_ARGS;
  if (((uintptr_t)in_bf16 | (uintptr_t)w_bf16 | (uintptr_t)prod_bf16 | (uintptr_t)res | (uintptr_t)res_out |
       (uintptr_t)norm_out_bf16 | (uintptr_t)bias | (uintptr_t)g0 | (uintptr_t)b0 | (uintptr_t)g1 | (uintptr_t)b1) % 16 != 0)
    return SS_BAD_ARGS;
  ss::ConvFuse f{tile_order, tile_pos, (int*)tile_flags, 0, 0, 0, order_row, ypos_rank, bias, res, g0, b0, g1, b1, eps, n, k3, res_out,
                 (__nv_bfloat16*)norm_out_bf16};
  switch ((cout + 255) / 256) {
    case 1: return ss::launch_gather_gemm_pair<1>(in_bf16, pair_in, w_bf16, tile_tap, p_pad, k3, cin, cout, prod_bf16, f, stream);
    case 2: return ss::launch_gather_gemm_pair<2>(in_bf16, pair_in, w_bf16, tile_tap, p_pad, k3, cin, cout, prod_bf16, f, stream);
    case 3: return ss::launch_gather_gemm_pair<3>(in_bf16, pair_in, w_bf16, tile_tap, p_pad, k3, cin, cout, prod_bf16, f, stream);
    default: return ss::launch_gather_gemm_pair<4>(in_bf16, pair_in, w_bf16, tile_tap, p_pad, k3, cin, cout, prod_bf16, f, stream);
  }
}
