// Submanifold convolution, stage 1 (gather-GEMM), second generation: persistent CTAs, 256 x BN tiles.
//
// Replaces (reference): spconv.SubMConv3d forward for the xCPE 3^3 convs
// (point_transformer_v3m1_base.py:277-284; fp32 in the reference, bf16 x bf16 -> fp32 here).
//
//   prod[r, :] = X[pair_in[r], :] @ W_tap(r)^T        r < p_pad, taps contiguous and padded to 256 rows
//
// Why 256 rows per CTA: the first-generation kernel (conv_gemm.cu, 128 x 256 tiles) reads 48 KB of operands
// from L2 per 4.2 MFLOP (85 FLOP/B) and is L2-bandwidth bound at ~0.6 PFLOP/s; two 128-row M tiles sharing
// every W stage read 64 KB per 8.4 MFLOP (128 FLOP/B).  TMEM holds both fp32 accumulators (2 x BN columns).
// Work items (256-row tile, BN-column slab) are walked slab-fastest, so the slabs of one tile run on
// neighbouring SMs at the same time and their identical A gathers hit L2.
//
// 18 warps:  0-7  epilogue, one warp per (row quarter, accumulator): TMEM -> bf16 -> shared-memory transpose ->
//                 64-byte coalesced global stores
//            8-15 A producers: 16-byte cp.async gathers into 128B-swizzled K-major tiles, published per thread with
//                 cp.async.mbarrier.arrive.noinc (the arrival fires when the thread's copies have landed; a
//                 wait_group + fence + arrive chain was measured to serialise MMA k behind MMA k-1)
//            16   W producer (TMA, one lane)       17   MMA issuer (tcgen05.mma M128 x N=BN x K16, 8 per stage)
// The producers run ahead into the next work item while the epilogue drains the accumulators.
#include "tc_common.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

constexpr int kG2Threads = 576;
constexpr int kG2TileM = 256;
constexpr int kG2BK = 64;  // bf16 elements per K chunk = one 128-byte swizzle row

template <int BN, int STAGES>
struct Gemm2Smem {
  static constexpr int kABytes = 128 * kG2BK * 2;  // one 128-row A tile: 16 KB
  static constexpr int kBBytes = BN * kG2BK * 2;
  static constexpr int kStageBytes = 2 * kABytes + kBBytes;
  static constexpr int kOffStage = 0;
  static constexpr int kOffEpi = STAGES * kStageBytes;  // 8 warps x 2 KB transpose buffers
  static constexpr int kOffBar = kOffEpi + 8 * 2048;
  static constexpr int kTotal = kOffBar + 256 + 1024 /*alignment slack*/;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kG2Threads, 1)
gather_gemm256_kernel(const __nv_bfloat16* __restrict__ X, const int32_t* __restrict__ pair_in,
                      const __grid_constant__ CUtensorMap tmap_w, const int32_t* __restrict__ tile_tap, int cin, int cout,
                      int n_slabs, int64_t n_items, __nv_bfloat16* __restrict__ prod) {
  using S = Gemm2Smem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + S::kOffBar);  // [STAGES] 256 gather threads + 1 TMA expect_tx arrive
  uint64_t* empty_bar = full_bar + STAGES;              // [STAGES] one tcgen05.commit
  uint64_t* acc_full = empty_bar + STAGES;              // accumulators of the item complete
  uint64_t* acc_empty = acc_full + 1;                   // accumulators drained (256 epilogue threads)
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably uniform
  const int nk = (cin + kG2BK - 1) / kG2BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&full_bar[s], 256 + 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    tc::mbar_init(acc_full, 1);
    tc::mbar_init(acc_empty, 256);
    tc::mbar_fence_init();
  }
  if (warp == 16 && lane == 0) tc::tma_prefetch_desc(&tmap_w);
  if (warp == 17) tc::tmem_alloc<(2 * BN < 32 ? 32 : 2 * BN)>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8 && warp < 16) {
    // ------------------------------------------------------------------ A producers (gather), 256 threads
    const int tid = threadIdx.x - 256;  // 0..255
    const int sub = tid >> 3, c = tid & 7;  // 8 lanes cover one 128-byte row segment
    int64_t g = 0;                          // chunks issued so far (all items)
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int64_t tile = item / n_slabs;
      int32_t rows[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rows[i] = pair_in[tile * kG2TileM + i * 32 + sub];
      for (int kc = 0; kc < nk; ++kc, ++g) {
        const int s = (int)(g % STAGES);
        tc::mbar_wait(&empty_bar[s], (uint32_t)((g / STAGES) & 1) ^ 1);
        const uint32_t a_base = tc::smem_u32(smem + S::kOffStage + s * S::kStageBytes);
        const int k0 = kc * kG2BK;
        if (c < (min(kG2BK, cin - k0) >> 3)) {  // valid 16-byte chunks of this K chunk
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 32 + sub;  // rows 0..127: tile A0, 128..255: tile A1 (each 16 KB, own swizzle atom grid)
            tc::cp_async16(a_base + (r >> 7) * S::kABytes + tc::sw128_offset(r & 127, c),
                           X + (size_t)rows[i] * cin + k0 + c * 8);
          }
        }
        // the stage is published by the copies themselves (arrive when this thread's cp.asyncs have landed):
        // the thread never blocks on its own data, so STAGES stages of gathers are in flight
        tc::cp_async_mbar_arrive_noinc(&full_bar[s]);
      }
    }
  } else if (warp == 16) {
    // ------------------------------------------------------------------ W producer (TMA, one lane)
    if (lane == 0) {
      int64_t g = 0;
      for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int64_t tile = item / n_slabs;
        const int n0 = (int)(item - tile * n_slabs) * BN;
        const int tap = tile_tap[tile];
        for (int kc = 0; kc < nk; ++kc, ++g) {
          const int s = (int)(g % STAGES);
          tc::mbar_wait(&empty_bar[s], (uint32_t)((g / STAGES) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(&full_bar[s], S::kBBytes);
          tc::tma_load_2d(tc::smem_u32(smem + S::kOffStage + s * S::kStageBytes + 2 * S::kABytes), &tmap_w, kc * kG2BK,
                          tap * cout + n0, &full_bar[s]);
        }
      }
    }
  } else if (warp == 17) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, elected lane per op)
    constexpr uint32_t idesc = tc::umma_idesc_bf16(128, BN);
    const uint64_t d_base = tc::umma_desc_sw128(0);
    const uint32_t s0 = tc::smem_u32(smem + S::kOffStage);
    int64_t g = 0;
    int it = 0;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      tc::mbar_wait(acc_empty, (uint32_t)(it & 1) ^ 1);  // previous item's accumulators drained
      tc::tc_fence_after();
      for (int kc = 0; kc < nk; ++kc, ++g) {
        const int s = (int)(g % STAGES);
        tc::mbar_wait(&full_bar[s], (uint32_t)((g / STAGES) & 1));
        tc::tc_fence_after();
        const uint32_t a0 = (s0 + s * S::kStageBytes) >> 4;
        const uint32_t a1 = a0 + (S::kABytes >> 4);
        const uint32_t b0 = a0 + (2 * S::kABytes >> 4);
        const int ksteps = min(kG2BK, cin - kc * kG2BK) >> 4;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t db = d_base | (uint64_t)((b0 + 2 * k) & 0x3fff);
          tc::umma_bf16_elect(tmem_base, d_base | (uint64_t)((a0 + 2 * k) & 0x3fff), db, idesc, (kc | k) ? 1u : 0u);
          tc::umma_bf16_elect(tmem_base + BN, d_base | (uint64_t)((a1 + 2 * k) & 0x3fff), db, idesc, (kc | k) ? 1u : 0u);
        }
        tc::umma_commit_elect(&empty_bar[s]);  // frees the stage when these MMAs have read it
      }
      tc::umma_commit_elect(acc_full);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 0..7: (quarter, accumulator)
    uint8_t* stg = smem + S::kOffEpi + warp * 2048;  // [32 rows][64 B], 16-byte chunks XOR-swizzled by (row >> 1) & 3
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int it = 0;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int64_t tile = item / n_slabs;
      const int n0 = (int)(item - tile * n_slabs) * BN;
      tc::mbar_wait(acc_full, (uint32_t)(it & 1));
      tc::tc_fence_after();
      {
        __nv_bfloat16* obase = prod + ((size_t)tile * kG2TileM + half * 128 + quarter * 32) * cout + n0;
#pragma unroll 1
        for (int j = 0; j < BN / 32; ++j) {
          if (n0 + j * 32 >= cout) break;
          uint32_t v[32];
          tc::tmem_ld32(t_lane + half * BN + j * 32, v);
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
            *reinterpret_cast<uint4*>(obase + (size_t)r * cout + j * 32 + cc * 8) = o;  // 4 lanes = 64 contiguous bytes
          }
          __syncwarp();
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(acc_empty);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc::tc_fence_after();
    tc::tmem_dealloc<(2 * BN < 32 ? 32 : 2 * BN)>(tmem_base);
  }
}

template <int BN, int STAGES>
static int launch_gather_gemm256(const void* X, const int32_t* pair_in, const CUtensorMap& tmap, const int32_t* tile_tap,
                                 int64_t tiles, int cin, int cout, void* prod, cudaStream_t stream) {
  using S = Gemm2Smem<BN, STAGES>;
  auto kern = gather_gemm256_kernel<BN, STAGES>;
  SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
  const int n_slabs = (cout + BN - 1) / BN;
  const int64_t n_items = tiles * n_slabs;
  const int grid = (int)imin64(n_items, kNumSMs);
  kern<<<grid, kG2Threads, S::kTotal, stream>>>((const __nv_bfloat16*)X, pair_in, tmap, tile_tap, cin, cout, n_slabs,
                                                n_items, (__nv_bfloat16*)prod);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // namespace ss

extern "C" int ss_subm_conv_gemm256(const void* in_bf16, const int32_t* pair_in, const void* w_bf16,
                                    const int32_t* tile_tap, int64_t p_pad, int k3, int cin, int cout, void* prod_bf16,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (k3 < 1 || p_pad < 0 || p_pad % ss::kG2TileM != 0 || cin < 16 || cin % 16 != 0 || cout < 32 || cout % 32 != 0)
    return SS_BAD_ARGS;
  if (p_pad == 0) return SS_OK;
  if (!in_bf16 || !pair_in || !w_bf16 || !tile_tap || !prod_bf16) return SS_BAD_ARGS;
  if (((uintptr_t)in_bf16 | (uintptr_t)w_bf16 | (uintptr_t)prod_bf16) % 16 != 0) return SS_BAD_ARGS;
  const int64_t tiles = p_pad / ss::kG2TileM;
  const int bn = cout >= 256 ? 256 : (cout > 64 ? 128 : (cout > 32 ? 64 : 32));
  // W viewed as one [k3 * cout, cin] K-major matrix; rows past the last tap read as zero (TMA OOB fill)
  CUtensorMap tmap;
  int rc = ss::make_tmap_bf16_2d(&tmap, w_bf16, (uint64_t)k3 * cout, (uint64_t)cin, (uint32_t)bn, ss::kG2BK);
  if (rc) return rc;
  switch (bn) {
    case 256: return ss::launch_gather_gemm256<256, 3>(in_bf16, pair_in, tmap, tile_tap, tiles, cin, cout, prod_bf16, stream);
    case 128: return ss::launch_gather_gemm256<128, 4>(in_bf16, pair_in, tmap, tile_tap, tiles, cin, cout, prod_bf16, stream);
    case 64: return ss::launch_gather_gemm256<64, 4>(in_bf16, pair_in, tmap, tile_tap, tiles, cin, cout, prod_bf16, stream);
    default: return ss::launch_gather_gemm256<32, 4>(in_bf16, pair_in, tmap, tile_tap, tiles, cin, cout, prod_bf16, stream);
  }
}
