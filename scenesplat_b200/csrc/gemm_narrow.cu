// Narrow Linear layers (cin in {32, 64, 96, 128}, cout <= 256): out = act(x W^T + b), bf16 in / bf16 out, fp32 accumulate.
//
// Replaces (reference): the nn.Linear layers of the first two encoder stages (C = 32 / 64: qkv, proj, fc1 + GELU, fc2;
// point_transformer_v3m1_base.py:181-248) and the narrow pooling / unpooling projections (:416,471-482).
//
// These layers move 2 (cin + cout) bytes per row for 2 cin cout FLOPs: at C = 32 that is 12 FLOP / B, HBM-bound by a factor
// of 20 on B200.  The CTA-pair tcgen05 GEMM (gemm2cta.cu) has one 64-wide K chunk per 256 x 256 tile there and is bound by
// its per-tile TMA -> MMA -> epilogue hand-overs (0.050 ms for 299 k x 32 -> 96 where the bytes need 0.012 ms).  Here W
// (<= 64 KB) is staged once per CTA in shared memory and every warp streams 16-row strips: A fragments straight from
// global memory (each 64..256-byte row is read by 4 lanes x cin / 16 steps, the sectors stay in L1 in between), warp-level
// mma.sync m16n8k16 against B fragments from shared memory, bias / GELU on the accumulator registers, the strip's output
// staged per warp in shared memory and written with 16-byte row-contiguous stores.  No tensor-memory / TMA machinery: the
// kernel is a row streamer.  Used for cin = 32 only (ss_linear_act_bf16): its time grows by ~8 us per 32 output columns at
// 299 k rows (the warp-level MMA path and the epilogue, not the bytes), so from cin = 64 / cout = 192 on the pair kernel is
// faster again (profiles/r2_gemm.md).
#include <cuda_bf16.h>
#include "common.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

constexpr int kNarrowThreads = 256;
constexpr int kNarrowChunk = 128;  // output columns per accumulator pass (16 n-tiles x 4 fp32 registers)

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// KB = cin / 32 (1..4).  Shared memory: W [cout][ldw] bf16, then one [16][kNarrowChunk + 8] bf16 staging tile per warp.
//
// The MMA's k index is a summation index: any assignment of physical columns to it is valid as long as A and B use the
// same one.  Per 32-column block a lane therefore takes the 16 CONTIGUOUS bytes [tig * 8, tig * 8 + 8) of its rows (one
// LDG.128 per row: 4 lanes cover a 64-byte row segment) and of W's row (one LDS.128 per n-tile), and feeds element pairs
// 0 / 1 to the first k-step's (a0 a1 | a2 a3, b0 | b1) and pairs 2 / 3 to the second: 4x fewer load instructions than
// the canonical fragment layout (which takes 4-byte pieces 16 bytes apart), full sectors.
template <int KB, int ACT>
__global__ void __launch_bounds__(kNarrowThreads, 2)
linear_narrow_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias,
                     int64_t n, int cout, __nv_bfloat16* __restrict__ out) {
  constexpr int cin = KB * 32;
  // row stride = 64 (mod 128) bytes: the 8 lanes of an LDS.128 phase (2 rows x 4 pieces) hit 8 distinct 16-byte bank groups
  constexpr int ldw = cin + ((cin * 2) % 128 == 64 ? 0 : 32);
  constexpr int lds = kNarrowChunk + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sW = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __nv_bfloat16* sO = sW + (size_t)cout * ldw + (size_t)warp * 16 * lds;
  for (int i = threadIdx.x; i < cout * (cin / 8); i += kNarrowThreads) {
    const int r = i / (cin / 8), c = i - r * (cin / 8);
    *reinterpret_cast<uint4*>(sW + (size_t)r * ldw + c * 8) = *reinterpret_cast<const uint4*>(w + (size_t)r * cin + c * 8);
  }
  __syncthreads();
  const int gid = lane >> 2, tig = lane & 3;
  const int64_t strips = (n + 15) >> 4;
  const int64_t nw = (int64_t)gridDim.x * (kNarrowThreads / 32);
  auto load_strip = [&](int64_t st, uint4 (&q0)[KB], uint4 (&q1)[KB]) {
    const int64_t r0 = st * 16 + gid, r1 = r0 + 8;
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      q0[kb] = r0 < n ? *reinterpret_cast<const uint4*>(x + r0 * cin + kb * 32 + tig * 8) : make_uint4(0u, 0u, 0u, 0u);
      q1[kb] = r1 < n ? *reinterpret_cast<const uint4*>(x + r1 * cin + kb * 32 + tig * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  int64_t s = (int64_t)blockIdx.x * (kNarrowThreads / 32) + warp;
  uint4 q0[KB], q1[KB], p0[KB], p1[KB];
  if (s < strips) load_strip(s, q0, q1);
  for (; s < strips; s += nw) {
    // the next strip's rows are in flight while this one is multiplied, converted and stored (a row streamer lives on
    // bytes in flight: 16 warps per SM x one strip each did not cover the DRAM latency)
    if (s + nw < strips) load_strip(s + nw, p0, p1);
    for (int n0 = 0; n0 < cout; n0 += kNarrowChunk) {
      const int nt_n = min(kNarrowChunk, cout - n0) >> 3;  // n-tiles of 8 columns in this chunk
      float acc[kNarrowChunk / 8][4];
#pragma unroll
      for (int t = 0; t < kNarrowChunk / 8; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const uint32_t a_lo[4] = {q0[kb].x, q1[kb].x, q0[kb].y, q1[kb].y};
        const uint32_t a_hi[4] = {q0[kb].z, q1[kb].z, q0[kb].w, q1[kb].w};
#pragma unroll
        for (int t = 0; t < kNarrowChunk / 8; ++t) {
          if (t < nt_n) {
            const uint4 wv = *reinterpret_cast<const uint4*>(sW + (size_t)(n0 + t * 8 + gid) * ldw + kb * 32 + tig * 8);
            mma_bf16_16816(acc[t], a_lo, wv.x, wv.y);
            mma_bf16_16816(acc[t], a_hi, wv.z, wv.w);
          }
        }
      }
      // bias / activation on the accumulators, bf16 pairs into the warp's staging tile
#pragma unroll
      for (int t = 0; t < kNarrowChunk / 8; ++t) {
        if (t < nt_n) {
          const int c = t * 8 + tig * 2;
          float b0 = 0.f, b1 = 0.f;
          if (bias) {
            b0 = bias[n0 + c];
            b1 = bias[n0 + c + 1];
          }
          float v0 = acc[t][0] + b0, v1 = acc[t][1] + b1, v2 = acc[t][2] + b0, v3 = acc[t][3] + b1;
          if (ACT == 1) {
            v0 = gelu_fast(v0);
            v1 = gelu_fast(v1);
            v2 = gelu_fast(v2);
            v3 = gelu_fast(v3);
          }
          *reinterpret_cast<uint32_t*>(sO + gid * lds + c) = pack2_bf16(v0, v1);
          *reinterpret_cast<uint32_t*>(sO + (gid + 8) * lds + c) = pack2_bf16(v2, v3);
        }
      }
      __syncwarp();
      const int pieces = nt_n;  // 16-byte pieces per row of the chunk
      for (int i = lane; i < 16 * pieces; i += 32) {
        const int r = i / pieces, p = i - r * pieces;
        const int64_t row = s * 16 + r;
        if (row < n)
          *reinterpret_cast<uint4*>(out + row * cout + n0 + p * 8) = *reinterpret_cast<const uint4*>(sO + r * lds + p * 8);
      }
      __syncwarp();
    }
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      q0[kb] = p0[kb];
      q1[kb] = p1[kb];
    }
  }
}

template <int KB>
static int launch_narrow_kb(const void* x, const void* w, const float* bias, int64_t n, int cout, int act, void* out,
                            cudaStream_t stream) {
  constexpr int cin = KB * 32;
  constexpr int ldw = cin + ((cin * 2) % 128 == 64 ? 0 : 32);
  const size_t smem = ((size_t)cout * ldw + (size_t)(kNarrowThreads / 32) * 16 * (kNarrowChunk + 8)) * 2;
  const int64_t strips = (n + 15) >> 4;
  const int blocks = (int)imin64(ceil_div64(strips, kNarrowThreads / 32), 2 * kNumSMs);
#define SS_NARROW_(A)                                                                                              \
  do {                                                                                                             \
    auto kern = linear_narrow_kernel<KB, A>;                                                                       \
    SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
    kern<<<blocks, kNarrowThreads, smem, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias, n, cout, \
                                                   (__nv_bfloat16*)out);                                           \
  } while (0)
  if (act == 1) SS_NARROW_(1);
  else SS_NARROW_(0);
#undef SS_NARROW_
  SS_CHECK_LAUNCH();
  return SS_OK;
}

// Called by ss_linear_act_bf16 (gemm2cta.cu) for cin in {32, 64, 96, 128}, cout % 8 == 0, cout <= 256.
int launch_linear_narrow(const void* x, const void* w, const float* bias, int64_t n, int cin, int cout, int act, void* out,
                         cudaStream_t stream) {
  switch (cin / 32) {
    case 1: return launch_narrow_kb<1>(x, w, bias, n, cout, act, out, stream);
    case 2: return launch_narrow_kb<2>(x, w, bias, n, cout, act, out, stream);
    case 3: return launch_narrow_kb<3>(x, w, bias, n, cout, act, out, stream);
    case 4: return launch_narrow_kb<4>(x, w, bias, n, cout, act, out, stream);
    default: return SS_BAD_ARGS;
  }
}

}  // namespace ss
