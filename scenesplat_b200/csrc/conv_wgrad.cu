// Submanifold convolution, weight gradient on the tensor cores (training path).
//
// Replaces (reference): the wgrad of spconv.SubMConv3d (autograd of point_transformer_v3m1_base.py:277-284).
//
//   dW[t][co][ci] = sum over the active pairs r of tap t of  dY[pair_out[r]][co] * X[pair_in[r]][ci]
//
// A GEMM per tap whose reduction dimension is the PAIR index: M = cout, N = cin, K = pairs of the tap, with BOTH
// operands gathered row-wise (a gathered row is contiguous along M resp. N: MN-major operands, UMMA no-swizzle
// core-matrix layout, the same layout the attention kernel uses for V).  The pair list of a tap is cut into K
// chunks (split-K: the centre tap alone holds 1/6.5 of all pairs), a work item is (K chunk, 256 x 256 output
// tile); the two M = 128 halves of a tile share every B stage like in the forward kernel (conv_gemm2.cu) and
// the partial tile is added to the fp32 dW with vector reductions (red.global.add.v4.f32; the summation order
// over the chunks is not fixed, as in every split-K weight gradient).
//
// 18 warps: 0-7 epilogue (TMEM -> red.global), 8-11 / 12-15 gather producers of dY / X (cp.async, 16 per thread and
// stage, whole 32-byte sectors per instruction, published with cp.async.mbarrier.arrive.noinc), 16 idle, 17 MMA issuer.
#include "tc_common.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

constexpr int kWgThreads = 576;
constexpr int kWgKS = 64;     // pair rows per stage
constexpr int kWgTile = 256;  // output tile edge (two M = 128 halves x N = 256)
constexpr int kWgStages = 3;

struct WgSmem {
  static constexpr int kABytes = 2 * 16 * kWgKS * 16;  // two halves x 16 chunks of 8 channels: 32 KB
  static constexpr int kBBytes = 32 * kWgKS * 16;       // 32 chunks: 32 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOffBar = kWgStages * kStageBytes;
  static constexpr int kTotal = kOffBar + 256 + 1024;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// chunks: [n_chunks] int4 (tap, k_begin, k_end, 0) with k_* absolute rows of the pair lists
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ dY,
                  const int32_t* __restrict__ pair_in, const int64_t* __restrict__ pair_out,
                  const int4* __restrict__ chunks, int n_chunks, int cin, int cout, float* __restrict__ dW) {
  using S = WgSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + S::kOffBar);
  uint64_t* empty_bar = full_bar + kWgStages;
  uint64_t* acc_full = empty_bar + kWgStages;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int mt = (cout + kWgTile - 1) / kWgTile, nt = (cin + kWgTile - 1) / kWgTile;
  const int64_t n_items = (int64_t)n_chunks * mt * nt;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      tc::mbar_init(&full_bar[s], 256);
      tc::mbar_init(&empty_bar[s], 1);
    }
    tc::mbar_init(acc_full, 1);
    tc::mbar_init(acc_empty, 256);
    tc::mbar_fence_init();
  }
  if (warp == 17) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8 && warp < 16) {
    // ------------------------------------------------------------------ gather producers
    // warps 8-11 gather the dY rows (A), warps 12-15 the X rows (B).  One instruction of a warp moves 16 pair rows x
    // 2 adjacent 16-byte pieces (lane = (row, piece parity)): every 32-byte L2 sector is fetched whole by one
    // instruction, and the 16 rows land in 16 distinct 16-byte slots of the core-matrix layout.
    const int pw = warp - 8;                 // 0..7
    const bool side_b = pw >= 4;
    const int j = (pw & 3) * 16 + (lane >> 1);  // pair row inside the stage
    const int hpar = lane & 1;
    // element (row j, 16-byte chunk c) -> c * (KS * 16) + (j / 8) * 128 + (j % 8) * 16   (per operand / half)
    const uint32_t row_off = (uint32_t)((j >> 3) * 128 + (j & 7) * 16);
    int64_t g = 0;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int64_t ch = item / (mt * nt);
      const int tile = (int)(item - ch * (mt * nt));
      const int m0 = (tile / nt) * kWgTile, n0 = (tile % nt) * kWgTile;
      const int4 c4 = chunks[ch];
      const int k_beg = c4.y, k_end = c4.z;
      const __nv_bfloat16* base = side_b ? X : dY;
      const int width = side_b ? cin : cout, col0 = side_b ? n0 : m0;
      for (int k0 = k_beg; k0 < k_end; k0 += kWgKS, ++g) {
        const int s = (int)(g % kWgStages);
        const int r = k0 + j;
        const bool ok = r < k_end;
        const int64_t row = ok ? (side_b ? (int64_t)pair_in[r] : pair_out[r]) : 0;
        tc::mbar_wait(&empty_bar[s], (uint32_t)((g / kWgStages) & 1) ^ 1);
        const uint32_t st = tc::smem_u32(smem + s * S::kStageBytes) + (side_b ? (uint32_t)S::kABytes : 0u) + row_off;
        const __nv_bfloat16* src = base + (size_t)row * width;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int c = 2 * u + hpar;  // chunk 0..31 (A: half = c / 16)
          const int col = col0 + c * 8;
          const bool v = ok && col < width;
          tc::cp_async16(st + (uint32_t)(c * (kWgKS * 16)), src + (v ? col : 0), v ? 16u : 0u);
        }
        tc::cp_async_mbar_arrive_noinc(&full_bar[s]);
      }
    }
  } else if (warp == 17) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = tc::umma_idesc_bf16(128, 256, 1, 1);  // A and B MN-major
    const uint32_t s0 = tc::smem_u32(smem);
    const uint64_t d_base = tc::umma_desc_nosw(0, 128, kWgKS * 16);
    int64_t g = 0;
    int it = 0;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int4 c4 = chunks[item / (mt * nt)];
      tc::mbar_wait(acc_empty, (uint32_t)(it & 1) ^ 1);
      tc::tc_fence_after();
      bool first = true;
      for (int k0 = c4.y; k0 < c4.z; k0 += kWgKS, ++g) {
        const int s = (int)(g % kWgStages);
        tc::mbar_wait(&full_bar[s], (uint32_t)((g / kWgStages) & 1));
        tc::tc_fence_after();
        const uint32_t a0 = (s0 + s * S::kStageBytes) >> 4;
        const uint32_t a1 = a0 + (S::kABytes / 2 >> 4);
        const uint32_t b0 = a0 + (S::kABytes >> 4);
#pragma unroll
        for (int k = 0; k < kWgKS / 16; ++k) {  // 16 pair rows = 2 groups of 8 rows = 256 bytes
          const uint64_t db = d_base | (uint64_t)((b0 + 16 * k) & 0x3fff);
          const uint32_t acc = (first && k == 0) ? 0u : 1u;
          tc::umma_bf16_elect(tmem_base, d_base | (uint64_t)((a0 + 16 * k) & 0x3fff), db, idesc, acc);
          tc::umma_bf16_elect(tmem_base + 256, d_base | (uint64_t)((a1 + 16 * k) & 0x3fff), db, idesc, acc);
        }
        first = false;
        tc::umma_commit_elect(&empty_bar[s]);
      }
      tc::umma_commit_elect(acc_full);
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ epilogue: (row quarter, M half)
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int it = 0;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int64_t ch = item / (mt * nt);
      const int tile = (int)(item - ch * (mt * nt));
      const int m0 = (tile / nt) * kWgTile, n0 = (tile % nt) * kWgTile;
      const int tap = chunks[ch].x;
      tc::mbar_wait(acc_full, (uint32_t)(it & 1));
      tc::tc_fence_after();
      const int m = m0 + half * 128 + quarter * 32 + lane;
      float* orow = dW + ((size_t)tap * cout + (m < cout ? m : 0)) * cin + n0;
#pragma unroll 1
      for (int jj = 0; jj < 256 / 32; ++jj) {
        if (n0 + jj * 32 >= cin) break;
        uint32_t v[32];
        tc::tmem_ld32(t_lane + half * 256 + jj * 32, v);
        tc::tmem_ld_wait();
        if (m < cout) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            red_add_v4(orow + jj * 32 + u * 4, __uint_as_float(v[4 * u]), __uint_as_float(v[4 * u + 1]),
                       __uint_as_float(v[4 * u + 2]), __uint_as_float(v[4 * u + 3]));
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
    tc::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace ss

extern "C" int ss_subm_conv_wgrad(const void* x_bf16, const void* dy_bf16, const int32_t* pair_in, const int64_t* pair_out,
                                  const int32_t* chunks, int n_chunks, int k3, int cin, int cout, float* dw,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_chunks < 0 || k3 < 1 || cin < 32 || cin % 32 != 0 || cout < 32 || cout % 32 != 0) return SS_BAD_ARGS;
  if (n_chunks == 0) return SS_OK;
  if (!x_bf16 || !dy_bf16 || !pair_in || !pair_out || !chunks || !dw) return SS_BAD_ARGS;
  if (((uintptr_t)x_bf16 | (uintptr_t)dy_bf16 | (uintptr_t)dw | (uintptr_t)chunks) % 16 != 0) return SS_BAD_ARGS;
  auto kern = ss::conv_wgrad_kernel;
  SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ss::WgSmem::kTotal));
  const int mt = (cout + ss::kWgTile - 1) / ss::kWgTile, nt = (cin + ss::kWgTile - 1) / ss::kWgTile;
  const int64_t n_items = (int64_t)n_chunks * mt * nt;
  const int grid = (int)ss::imin64(n_items, ss::kNumSMs);
  kern<<<grid, ss::kWgThreads, ss::WgSmem::kTotal, stream>>>((const __nv_bfloat16*)x_bf16, (const __nv_bfloat16*)dy_bf16,
                                                             pair_in, pair_out, (const int4*)chunks, n_chunks, cin, cout,
                                                             dw);
  SS_CHECK_LAUNCH();
  return SS_OK;
}
