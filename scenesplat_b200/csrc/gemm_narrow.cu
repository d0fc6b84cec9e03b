// Narrow Linear layers (cin <= 128, cout <= 256): out = act(x W^T + b), bf16 in / bf16 out, fp32 accumulate.
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
// kernel is a row streamer, the tensor pipe is idle either way.  Used for cin <= 32 only (ss_linear_act_bf16): from cin = 64
// on the 4-byte A-fragment loads (32 per strip and lane at cin = 128) bind it and the pair kernel is faster again.
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

// KS = cin / 16 (1..8).  Shared memory: W [cout][cin + 8] bf16, then one [16][kNarrowChunk + 8] bf16 staging tile per warp.
template <int KS, int ACT>
__global__ void __launch_bounds__(kNarrowThreads, 2)
linear_narrow_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias,
                     int64_t n, int cout, __nv_bfloat16* __restrict__ out) {
  constexpr int cin = KS * 16;
  constexpr int ldw = cin + 8;             // +16 bytes per row: the 8 rows of a B fragment fall into distinct banks
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
  for (int64_t s = (int64_t)blockIdx.x * (kNarrowThreads / 32) + warp; s < strips; s += nw) {
    const int64_t r0 = s * 16 + gid, r1 = r0 + 8;
    uint32_t a[KS][4];
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      const int c = kk * 16 + tig * 2;
      a[kk][0] = r0 < n ? *reinterpret_cast<const uint32_t*>(x + r0 * cin + c) : 0u;
      a[kk][1] = r1 < n ? *reinterpret_cast<const uint32_t*>(x + r1 * cin + c) : 0u;
      a[kk][2] = r0 < n ? *reinterpret_cast<const uint32_t*>(x + r0 * cin + c + 8) : 0u;
      a[kk][3] = r1 < n ? *reinterpret_cast<const uint32_t*>(x + r1 * cin + c + 8) : 0u;
    }
    for (int n0 = 0; n0 < cout; n0 += kNarrowChunk) {
      const int nt_n = min(kNarrowChunk, cout - n0) >> 3;  // n-tiles of 8 columns in this chunk
      float acc[kNarrowChunk / 8][4];
#pragma unroll
      for (int t = 0; t < kNarrowChunk / 8; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
        for (int t = 0; t < kNarrowChunk / 8; ++t) {
          if (t < nt_n) {
            const __nv_bfloat16* bp = sW + (size_t)(n0 + t * 8 + gid) * ldw + kk * 16 + tig * 2;
            mma_bf16_16816(acc[t], a[kk], *reinterpret_cast<const uint32_t*>(bp), *reinterpret_cast<const uint32_t*>(bp + 8));
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
  }
}

template <int KS>
static int launch_narrow_ks(const void* x, const void* w, const float* bias, int64_t n, int cout, int act, void* out,
                            cudaStream_t stream) {
  const size_t smem = ((size_t)cout * (KS * 16 + 8) + (size_t)(kNarrowThreads / 32) * 16 * (kNarrowChunk + 8)) * 2;
  const int64_t strips = (n + 15) >> 4;
  const int blocks = (int)imin64(ceil_div64(strips, kNarrowThreads / 32), 2 * kNumSMs);
#define SS_NARROW_(A)                                                                                              \
  do {                                                                                                             \
    auto kern = linear_narrow_kernel<KS, A>;                                                                       \
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

// Called by ss_linear_act_bf16 (gemm2cta.cu) for cin in {16, 32, .., 128}, cout % 8 == 0, cout <= 256.
int launch_linear_narrow(const void* x, const void* w, const float* bias, int64_t n, int cin, int cout, int act, void* out,
                         cudaStream_t stream) {
  switch (cin / 16) {
    case 1: return launch_narrow_ks<1>(x, w, bias, n, cout, act, out, stream);
    case 2: return launch_narrow_ks<2>(x, w, bias, n, cout, act, out, stream);
    case 3: return launch_narrow_ks<3>(x, w, bias, n, cout, act, out, stream);
    case 4: return launch_narrow_ks<4>(x, w, bias, n, cout, act, out, stream);
    case 5: return launch_narrow_ks<5>(x, w, bias, n, cout, act, out, stream);
    case 6: return launch_narrow_ks<6>(x, w, bias, n, cout, act, out, stream);
    case 7: return launch_narrow_ks<7>(x, w, bias, n, cout, act, out, stream);
    case 8: return launch_narrow_ks<8>(x, w, bias, n, cout, act, out, stream);
    default: return SS_BAD_ARGS;
  }
}

}  // namespace ss
