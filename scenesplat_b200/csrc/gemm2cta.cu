// Dense Linear layer on CTA PAIRS (tcgen05 cta_group::2): out = act(x W^T + b), bf16 in / bf16 out, fp32 accumulate.
//
// Replaces (reference): every nn.Linear of the PTv3 forward (point_transformer_v3m1_base.py: qkv / proj :181-221, MLP
// :225-248, pooling / unpooling projections :416,471-482) with the elementwise work that follows it fused into the TMEM
// epilogue:
//   ACT = 1  fc1 + bias + exact GELU: the N x 4C hidden tensor is written once instead of being written, read and
//            written again by a separate GELU pass;
//   RES = 1  fc2 + bias + residual add (Block, ref :334-336): the fp32 residual stream is updated in the epilogue and
//            its bf16 shadow (the next Block's conv operand) is written beside it; the bf16 round trip of the MLP
//            output and the separate add pass are gone.
//
// Why a CTA pair: with one CTA per tile, a 256 x 256 output tile needs both of its 128-row accumulators in that CTA's
// TMEM (512 columns, conv_gemm2.cu), which leaves no second buffer, so the epilogue cannot overlap the next tile's MMAs
// and a 15-instruction activation epilogue would be fully exposed.  Here the two CTAs of a cluster split the tile: each
// stages its own 128 rows of x and HALF of the 256 weight rows, the leader issues tcgen05.mma.cta_group::2 with M = 256
// (each CTA's tensor core reads both halves of W through the pair's shared-memory window), and each CTA accumulates
// its own 128 x 256 half in 256 TMEM columns.  The other 256 columns are a second accumulator: the epilogue of tile i
// runs under the MMAs of tile i + 1.  Operand bytes staged per FLOP are the same as in the 256-row single-CTA kernel
// (32 KB per 4.2 MFLOP per CTA).
//
// Per CTA, 10 warps: 0-7 epilogue (row quarter x column half: TMEM -> bias -> activation -> bf16 -> shared-memory
// transpose -> 64-byte coalesced stores), 8 TMA producer (one lane; both CTAs' copies signal the LEADER's full
// barrier), 9 MMA issuer (leader only) and TMEM allocator.  Stage release and accumulator hand-over go to both CTAs
// with multicast commits; the peer's epilogue warps release an accumulator with a remote mbarrier arrive.
#include "pair_common.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

constexpr int kPBK = 64;     // bf16 elements per K chunk = one 128-byte swizzle row
constexpr int kPStages = 5;
constexpr int kPTileM = 256; // rows per pair tile (128 per CTA)
constexpr int kPTileN = 256; // columns per pair tile (each CTA stages 128 weight rows)

struct PairSmem {
  static constexpr int kA = 128 * kPBK * 2;
  static constexpr int kB = 128 * kPBK * 2;
  static constexpr int kStage = kA + kB;
  static constexpr int kOffEpi = kPStages * kStage;  // 8 warps x (2 KB bf16 + 4 KB fp32) transpose buffers
  static constexpr int kOffBar = kOffEpi + 8 * 6144;
  static constexpr int kTotal = kOffBar + 256 + 1024 /*alignment slack*/;
};


// ACT: 0 = none, 1 = exact GELU.  RES: 1 = out_f32 = res + (x W^T + b) (res may be out_f32 itself), bf16 copy in `out`
// (optional).  EW: epilogue warps, 8 (row quarter x column half) or 16 (row quarter x column quarter; RES = 0 only: the
// GELU epilogue is 2 MUFU + 13 other instructions per element, ~5 k issue cycles per 128 x 256 tile against 6.1 k cycles
// of MMA, and with two warps per SM sub-partition its dependent chains are not hidden under the next tile's MMAs).
template <int ACT, int RES, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((EW + 2) * 32, 1)
linear_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                   const float* __restrict__ bias, int64_t n_rows, int cin, int cout, __nv_bfloat16* __restrict__ out,
                   const float* res, float* out_f32) {
  using S = PairSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + S::kOffBar);  // [stages] leader's copy is the one that counts
  uint64_t* empty_bar = full_bar + kPStages;            // [stages] one multicast commit
  uint64_t* acc_full = empty_bar + kPStages;            // [2] one multicast commit
  uint64_t* acc_empty = acc_full + 2;                   // [2] leader's copy: 8 warps x 2 CTAs
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = pair::cta_rank();
  const bool leader = rank == 0;
  const int64_t pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int nk = (cin + kPBK - 1) / kPBK;
  const int n_slabs = (cout + kPTileN - 1) / kPTileN;
  const int64_t n_items = ((n_rows + kPTileM - 1) / kPTileM) * n_slabs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPStages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&acc_full[b], 1);
      tc::mbar_init(&acc_empty[b], 2 * EW);
    }
    tc::mbar_fence_init();
  }
  if (warp == EW && lane == 0) {
    tc::tma_prefetch_desc(&tmap_x);
    tc::tma_prefetch_desc(&tmap_w);
  }
  if (warp == EW + 1) pair::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  pair::cluster_sync();  // both CTAs' barriers are initialised before anyone signals across the pair
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == EW) {
    // ------------------------------------------------------------------ TMA producer (one lane, both CTAs)
    if (lane == 0) {
      int64_t g = 0;
      for (int64_t item = pair_id; item < n_items; item += n_pairs) {
        const int64_t tile = item / n_slabs;
        const int n0 = (int)(item - tile * n_slabs) * kPTileN;
        for (int kc = 0; kc < nk; ++kc, ++g) {
          const int s = (int)(g % kPStages);
          tc::mbar_wait(&empty_bar[s], (uint32_t)((g / kPStages) & 1) ^ 1);
          if (leader) tc::mbar_arrive_expect_tx(&full_bar[s], 2 * S::kStage);  // both CTAs' tiles
          const uint32_t st = tc::smem_u32(smem + s * S::kStage);
          pair::tma_load_2d(st, &tmap_x, kc * kPBK, (int)(tile * kPTileM + rank * 128), &full_bar[s]);
          pair::tma_load_2d(st + S::kA, &tmap_w, kc * kPBK, n0 + (int)rank * 128, &full_bar[s]);
        }
      }
    }
  } else if (warp == EW + 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(kPTileM, kPTileN);
      const uint64_t d_base = tc::umma_desc_sw128(0);
      const uint32_t s0 = tc::smem_u32(smem);
      int64_t g = 0;
      int it = 0;
      for (int64_t item = pair_id; item < n_items; item += n_pairs, ++it) {
        const int b = it & 1;
        tc::mbar_wait(&acc_empty[b], (uint32_t)((it >> 1) & 1) ^ 1);  // both CTAs have drained this accumulator
        tc::tc_fence_after();
        for (int kc = 0; kc < nk; ++kc, ++g) {
          const int s = (int)(g % kPStages);
          tc::mbar_wait(&full_bar[s], (uint32_t)((g / kPStages) & 1));
          tc::tc_fence_after();
          const uint32_t a0 = (s0 + s * S::kStage) >> 4;
          const uint32_t b0 = a0 + (S::kA >> 4);
          const int ksteps = (min(kPBK, cin - kc * kPBK) + 15) >> 4;
          for (int k = 0; k < ksteps; ++k)
            pair::umma_bf16_elect(tmem_base + b * kPTileN, d_base | (uint64_t)((a0 + 2 * k) & 0x3fff),
                                  d_base | (uint64_t)((b0 + 2 * k) & 0x3fff), idesc, (kc | k) ? 1u : 0u);
          pair::umma_commit_elect(&empty_bar[s]);  // frees the stage in both CTAs
        }
        pair::umma_commit_elect(&acc_full[b]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 0..7: (row quarter, column half)
    // bf16 tile [32 rows][64 B], 16-byte chunks XOR-swizzled by (row >> 1) & 3; RES: fp32 tile [32 rows][128 B] behind it,
    // 16-byte chunks XOR-swizzled by row & 7
    static_assert(EW == 8 || (EW == 16 && RES == 0), "16 epilogue warps: no fp32 staging tiles");
    constexpr int kCols = 256 / (EW / 4);  // columns per epilogue warp
    constexpr int kGroups = kCols / 32;
    uint8_t* stg = smem + S::kOffEpi + warp * (EW == 8 ? 6144 : 2048);
    uint8_t* stf = stg + 2048;
    const int quarter = warp & 3, half = warp >> 2;  // (half = column part: 0..1 or 0..3)
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int it = 0;
    for (int64_t item = pair_id; item < n_items; item += n_pairs, ++it) {
      const int64_t tile = item / n_slabs;
      const int n0 = (int)(item - tile * n_slabs) * kPTileN;
      const int b = it & 1;
      tc::mbar_wait(&acc_full[b], (uint32_t)((it >> 1) & 1));
      tc::tc_fence_after();
      const int64_t row0 = tile * kPTileM + rank * 128 + quarter * 32;
      // 32-column groups of this warp inside the matrix; the TMEM load of group j + 1 is in flight while group j is
      // converted and stored (two register buffers, loop fully unrolled)
      const int nj = max(0, min(kGroups, (cout - (n0 + half * kCols) + 31) / 32));
      constexpr bool kDB = RES == 0 && (EW == 8 || ACT == 1);  // TMEM load of group j + 1 under the conversion of group j
      uint32_t vbuf[kDB ? 2 : 1][32];
      if (kDB && nj > 0) tc::tmem_ld32(t_lane + b * kPTileN + half * kCols, vbuf[0]);
#pragma unroll
      for (int j = 0; j < kGroups; ++j) {
        if (j >= nj) break;
        const int c0 = n0 + half * kCols + j * 32;
        uint32_t (&v)[32] = vbuf[kDB ? (j & 1) : 0];
        if constexpr (RES == 1) {
          tc::tmem_ld32(t_lane + b * kPTileN + half * kCols + j * 32, v);
          // residual tile -> shared memory with coalesced 128-byte row segments (8 lanes per row, 4 rows per access)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3), cc = lane & 7;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < n_rows) x = *reinterpret_cast<const float4*>(res + (size_t)(row0 + r) * cout + c0 + cc * 4);
            *reinterpret_cast<float4*>(stf + r * 128 + ((cc ^ (r & 7)) << 4)) = x;
          }
          __syncwarp();
          tc::tmem_ld_wait();
          // this thread's row: accumulator + bias + residual -> fp32 tile (in place) and bf16 tile
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float y[8];
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
              const int cc = 2 * u + hlf;
              float4* ptr = reinterpret_cast<float4*>(stf + lane * 128 + ((cc ^ (lane & 7)) << 4));
              const float4 x = *ptr;
              float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
              if (bias) bb = __ldg(reinterpret_cast<const float4*>(bias + c0 + cc * 4));
              y[hlf * 4 + 0] = x.x + (__uint_as_float(v[cc * 4 + 0]) + bb.x);
              y[hlf * 4 + 1] = x.y + (__uint_as_float(v[cc * 4 + 1]) + bb.y);
              y[hlf * 4 + 2] = x.z + (__uint_as_float(v[cc * 4 + 2]) + bb.z);
              y[hlf * 4 + 3] = x.w + (__uint_as_float(v[cc * 4 + 3]) + bb.w);
              *ptr = make_float4(y[hlf * 4 + 0], y[hlf * 4 + 1], y[hlf * 4 + 2], y[hlf * 4 + 3]);
            }
            uint4 o;
            o.x = tc::pack_bf16(y[0], y[1]);
            o.y = tc::pack_bf16(y[2], y[3]);
            o.z = tc::pack_bf16(y[4], y[5]);
            o.w = tc::pack_bf16(y[6], y[7]);
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4)) = o;
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3), cc = lane & 7;
            const float4 y = *reinterpret_cast<const float4*>(stf + r * 128 + ((cc ^ (r & 7)) << 4));
            if (row0 + r < n_rows) *reinterpret_cast<float4*>(out_f32 + (size_t)(row0 + r) * cout + c0 + cc * 4) = y;
          }
        } else {
        if (!kDB) tc::tmem_ld32(t_lane + b * kPTileN + half * kCols + j * 32, v);
        tc::tmem_ld_wait();
        if (kDB && j + 1 < nj) tc::tmem_ld32(t_lane + b * kPTileN + half * kCols + (j + 1) * 32, vbuf[kDB ? ((j + 1) & 1) : 0]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[u * 8 + e]);
          if (bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0 + u * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + u * 8 + 4));
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
          }
          if (ACT == 1) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = gelu_fast(f[e]);
          }
          uint4 o;
          o.x = tc::pack_bf16(f[0], f[1]);
          o.y = tc::pack_bf16(f[2], f[3]);
          o.z = tc::pack_bf16(f[4], f[5]);
          o.w = tc::pack_bf16(f[6], f[7]);
          *reinterpret_cast<uint4*>(stg + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4)) = o;
        }
        }
        __syncwarp();
        if (RES == 0 || out != nullptr) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = i * 8 + (lane >> 2), cc = lane & 3;
            const uint4 o = *reinterpret_cast<const uint4*>(stg + r * 64 + ((cc ^ ((r >> 1) & 3)) << 4));
            if (row0 + r < n_rows)
              *reinterpret_cast<uint4*>(out + (size_t)(row0 + r) * cout + c0 + cc * 8) = o;  // 4 lanes = 64 contiguous bytes
          }
        }
        __syncwarp();
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) pair::mbar_arrive_cta(&acc_empty[b], 0);  // the leader's MMA warp owns the accumulators
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  pair::cluster_sync();  // no CTA of the pair frees TMEM or exits while the other still computes or signals
  if (warp == EW + 1) {
    tc::tc_fence_after();
    pair::tmem_dealloc<512>(tmem_base);
  }
}

#ifndef SS_NARROW_MAX_CIN
#define SS_NARROW_MAX_CIN 32  // (64 / 128: measured slower than the pair kernel, profiles/r2_gemm.md)
#endif
#ifndef SS_GEMM_EW_PLAIN
#define SS_GEMM_EW_PLAIN 8
#endif
#ifndef SS_GEMM_EW_ACT
#define SS_GEMM_EW_ACT 16  // epilogue warps of the fc1 + GELU variant (tools/run_linear.py A/B: -DSS_GEMM_EW_ACT=8)
#endif
template <int ACT, int RES, int EW = 8>
static int launch_linear_pair(const CUtensorMap& tx, const CUtensorMap& tw, const float* bias, int64_t n, int cin, int cout,
                              void* out, const float* res, float* out_f32, cudaStream_t stream) {
  auto kern = linear_pair_kernel<ACT, RES, EW>;
  SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PairSmem::kTotal));
  const int64_t n_items = ((n + kPTileM - 1) / kPTileM) * ((cout + kPTileN - 1) / kPTileN);
  const int pairs = (int)imin64(n_items, kNumSMs / 2);
  kern<<<2 * pairs, (EW + 2) * 32, PairSmem::kTotal, stream>>>(tx, tw, bias, n, cin, cout, (__nv_bfloat16*)out, res, out_f32);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // namespace ss

namespace ss {
int launch_linear_narrow(const void* x, const void* w, const float* bias, int64_t n, int cin, int cout, int act, void* out,
                         cudaStream_t stream);  // gemm_narrow.cu
}

static int linear_pair_entry(const void* x_bf16, const void* w_bf16, const float* bias, int64_t n, int cin, int cout, int act,
                             void* out_bf16, const float* res, float* out_f32, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || cin < 16 || cin % 16 != 0 || cout < 32 || cout % 32 != 0 || act < 0 || act > 1) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!x_bf16 || !w_bf16) return SS_BAD_ARGS;
  if (((uintptr_t)x_bf16 | (uintptr_t)w_bf16 | (uintptr_t)out_bf16 | (uintptr_t)bias | (uintptr_t)res | (uintptr_t)out_f32) % 16 != 0)
    return SS_BAD_ARGS;
  // the C = 32 stage (enc0 qkv / proj / fc1, the 32 -> 64 pooling projection) is a row streamer: W in shared memory, warp-level
  // MMA (gemm_narrow.cu; measured faster than the pair kernel only up to cin = 32: profiles/r2_gemm.md)
  if (!res && cin % 32 == 0 && cin <= SS_NARROW_MAX_CIN && cout <= 256) return ss::launch_linear_narrow(x_bf16, w_bf16, bias, n, cin, cout, act, out_bf16, stream);
  CUtensorMap tx, tw;
  int rc = ss::make_tmap_bf16_2d(&tx, x_bf16, (uint64_t)n, (uint64_t)cin, 128, ss::kPBK);
  if (rc) return rc;
  rc = ss::make_tmap_bf16_2d(&tw, w_bf16, (uint64_t)cout, (uint64_t)cin, 128, ss::kPBK);
  if (rc) return rc;
  if (res) return ss::launch_linear_pair<0, 1>(tx, tw, bias, n, cin, cout, out_bf16, res, out_f32, stream);
  return act == 1 ? ss::launch_linear_pair<1, 0, SS_GEMM_EW_ACT>(tx, tw, bias, n, cin, cout, out_bf16, nullptr, nullptr, stream)
                  : ss::launch_linear_pair<0, 0, SS_GEMM_EW_PLAIN>(tx, tw, bias, n, cin, cout, out_bf16, nullptr, nullptr, stream);
}

extern "C" int ss_linear_act_bf16(const void* x_bf16, const void* w_bf16, const float* bias, int64_t n, int cin, int cout,
                                  int act, void* out_bf16, void* stream_) {
  if (!out_bf16) return SS_BAD_ARGS;
  return linear_pair_entry(x_bf16, w_bf16, bias, n, cin, cout, act, out_bf16, nullptr, nullptr, stream_);
}

extern "C" int ss_linear_residual_bf16(const void* x_bf16, const void* w_bf16, const float* bias, const float* res, int64_t n,
                                       int cin, int cout, float* out_f32, void* out_bf16, void* stream_) {
  if (!res || !out_f32) return SS_BAD_ARGS;
  return linear_pair_entry(x_bf16, w_bf16, bias, n, cin, cout, 0, out_bf16, res, out_f32, stream_);
}
