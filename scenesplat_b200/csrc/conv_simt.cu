// Submanifold convolution, SIMT fp32-accumulate path: used for the stem (5^3, Cin = 11 -> 32, a K too
// small for the tensor pipe) and for channel counts the tcgen05 path does not cover.
//
// Replaces (reference): spconv.SubMConv3d forward (call sites point_transformer_v3m1_base.py:277-284,
// :499-506; fp32 forced by pointcept/models/modules.py:68-74).
// out[p, co] = bias[co] + sum_t sum_ci Wt[t][ci][co] * in[nbr[t][p]][ci]   (cross-correlation).
// Optional fused epilogue: per-channel affine (eval BatchNorm) + GELU (the stem's BN -> GELU).
#include "common.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

template <typename T> __device__ __forceinline__ float cvt_in(T v);
template <> __device__ __forceinline__ float cvt_in<float>(float v) { return v; }
template <> __device__ __forceinline__ float cvt_in<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T cvt_out(float v);
template <> __device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

__device__ __forceinline__ float gelu_erf_(float x) { return gelu_fast(x); }

// One warp per output voxel; lane owns output channels lane + 32u (u < UMAX => Cout <= 32*UMAX).
template <typename TI, typename TO, int UMAX>
__global__ void __launch_bounds__(256)
subm_conv_simt_kernel(const TI* __restrict__ in, const int32_t* __restrict__ nbr, const float* __restrict__ wt,
                      const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                      int act, int64_t n, int k3, int cin, int cout, TO* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp0; p < n; p += nwarp) {
    float acc[UMAX];
#pragma unroll
    for (int u = 0; u < UMAX; ++u) acc[u] = 0.f;
    for (int t = 0; t < k3; ++t) {
      const int32_t q = nbr[(size_t)t * n + p];  // warp-uniform
      if (q < 0) continue;
      const TI* row = in + (size_t)q * cin;
      const float* w = wt + (size_t)t * cin * cout;
      for (int c0 = 0; c0 < cin; c0 += 32) {
        const float xv = (c0 + lane < cin) ? cvt_in<TI>(row[c0 + lane]) : 0.f;
        const int cm = min(32, cin - c0);
        for (int cc = 0; cc < cm; ++cc) {
          const float x = __shfl_sync(0xffffffffu, xv, cc);
          const float* wr = w + (size_t)(c0 + cc) * cout;
#pragma unroll
          for (int u = 0; u < UMAX; ++u) {
            const int co = lane + 32 * u;
            if (co < cout) acc[u] = fmaf(x, __ldg(wr + co), acc[u]);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UMAX; ++u) {
      const int co = lane + 32 * u;
      if (co < cout) {
        float v = acc[u] + (bias ? bias[co] : 0.f);
        if (scale) v = v * scale[co] + shift[co];
        if (act == 1) v = gelu_erf_(v);
        out[(size_t)p * cout + co] = cvt_out<TO>(v);
      }
    }
  }
}

// Small-channel variant (the stem: 5^3 taps, Cin = 11 -> Cout = 32): one warp per 32 CONSECUTIVE voxels.
// lane = voxel for the loads (the 125 neighbour-table reads are coalesced 128-byte rows, the 11-float input rows of
// a tap are fetched by all lanes at once: one memory latency per tap instead of one per (tap, voxel) pair),
// lane = output channel for the arithmetic (input values are broadcast with shuffles, the tap's weights sit in
// registers and are reused by the 32 voxels).  acc[i] = output row of voxel i, channel `lane`.
template <typename TI, typename TO, int CIN, bool EXACT>  // EXACT: cin == CIN (no per-channel predicates)
__global__ void __launch_bounds__(256)
subm_conv_small_kernel(const TI* __restrict__ in, const int32_t* __restrict__ nbr, const float* __restrict__ wt,
                       const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                       int act, int64_t n, int k3, int cin_rt, int cout, TO* __restrict__ out) {
  constexpr int CINP = (CIN + 3) / 4 * 4;
  __shared__ float4 s_rows[8][32 * (CINP / 4)];  // per warp: the 32 voxels' input rows of the current tap
  const int cin = EXACT ? CIN : cin_rt;
  const int lane = threadIdx.x & 31;
  const float4* warp_rows = s_rows[threadIdx.x >> 5];
  float4* my_row = s_rows[threadIdx.x >> 5] + lane * (CINP / 4);
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p0 = warp0 * 32; p0 < n; p0 += nwarp * 32) {
    const int64_t p = p0 + lane;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    int32_t q_next = p < n ? nbr[p] : -1;  // tap 0; the table read of tap t + 1 is issued before tap t's arithmetic
    for (int t = 0; t < k3; ++t) {
      const int32_t q = q_next;
      if (t + 1 < k3) q_next = p < n ? nbr[(size_t)(t + 1) * n + p] : -1;
      const uint32_t mask = __ballot_sync(0xffffffffu, q >= 0);
      if (mask == 0u) continue;
      float x[CINP], w[CIN];
#pragma unroll
      for (int ci = 0; ci < CINP; ++ci)
        x[ci] = (ci < CIN && q >= 0 && (EXACT || ci < cin)) ? cvt_in<TI>(in[(size_t)q * cin + ci]) : 0.f;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
        w[ci] = ((EXACT || ci < cin) && lane < cout) ? __ldg(wt + ((size_t)t * cin + ci) * cout + lane) : 0.f;
      // the lane's input row goes to shared memory; every lane then reads voxel i's row as broadcast 16-byte loads
      // (3 LDS.128 per active (tap, voxel) pair instead of 11 SHFL: the shuffles bound the round-1 kernel)
      if (q >= 0) {
#pragma unroll
        for (int c4 = 0; c4 < CINP / 4; ++c4)
          my_row[c4] = make_float4(x[4 * c4], x[4 * c4 + 1], x[4 * c4 + 2], x[4 * c4 + 3]);
      }
      __syncwarp();
#pragma unroll
      for (int g4 = 0; g4 < 8; ++g4) {
        if (((mask >> (4 * g4)) & 0xfu) == 0u) continue;  // warp-uniform
#pragma unroll
        for (int i = 4 * g4; i < 4 * g4 + 4; ++i) {
          if ((mask >> i) & 1u) {
            float xi[CINP];
#pragma unroll
            for (int c4 = 0; c4 < CINP / 4; ++c4) {
              const float4 v = warp_rows[i * (CINP / 4) + c4];
              xi[4 * c4] = v.x, xi[4 * c4 + 1] = v.y, xi[4 * c4 + 2] = v.z, xi[4 * c4 + 3] = v.w;
            }
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci)
              if (EXACT || ci < cin) acc[i] = fmaf(xi[ci], w[ci], acc[i]);
          }
        }
      }
      __syncwarp();
    }
    if (lane < cout) {
      const float b = bias ? bias[lane] : 0.f;
      const float sc = scale ? scale[lane] : 1.f, sh = scale ? shift[lane] : 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (p0 + i < n) {
          float v = acc[i] + b;
          if (scale) v = v * sc + sh;
          if (act == 1) v = gelu_erf_(v);
          out[(size_t)(p0 + i) * cout + lane] = cvt_out<TO>(v);
        }
      }
    }
  }
}

}  // namespace ss

extern "C" int ss_subm_conv_simt(const void* in, int in_is_bf16, const int32_t* nbr, const float* wt, const float* bias,
                                 const float* scale, const float* shift, int act, int64_t n, int k3, int cin, int cout,
                                 void* out, int out_is_bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || k3 < 1 || cin < 1 || cout < 1 || cout > 1024 || (scale && !shift)) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!in || !nbr || !wt || !out) return SS_BAD_ARGS;
  if (cin <= 16 && cout <= 32) {  // stem-like shapes: warp per 32 voxels
    const int blocks = (int)ss::imin64(ss::ceil_div64(n, 8 * 32), 8 * ss::kNumSMs);
#define SS_SMALL_(TI, TO)                                                                                            \
  do {                                                                                                               \
    if (cin == 11)                                                                                                   \
      ss::subm_conv_small_kernel<TI, TO, 11, true><<<blocks, 256, 0, stream>>>((const TI*)in, nbr, wt, bias, scale,   \
                                                                               shift, act, n, k3, cin, cout, (TO*)out); \
    else                                                                                                             \
      ss::subm_conv_small_kernel<TI, TO, 16, false><<<blocks, 256, 0, stream>>>((const TI*)in, nbr, wt, bias, scale,  \
                                                                                shift, act, n, k3, cin, cout, (TO*)out); \
  } while (0)
    if (in_is_bf16 && out_is_bf16) SS_SMALL_(__nv_bfloat16, __nv_bfloat16);
    else if (in_is_bf16) SS_SMALL_(__nv_bfloat16, float);
    else if (out_is_bf16) SS_SMALL_(float, __nv_bfloat16);
    else SS_SMALL_(float, float);
#undef SS_SMALL_
    SS_CHECK_LAUNCH();
    return SS_OK;
  }
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, 8), 32 * ss::kNumSMs);
#define SS_LAUNCH_(TI, TO, U)                                                                                     \
  ss::subm_conv_simt_kernel<TI, TO, U><<<blocks, 256, 0, stream>>>((const TI*)in, nbr, wt, bias, scale, shift, act, n, \
                                                                   k3, cin, cout, (TO*)out)
#define SS_BY_U_(TI, TO)                                   \
  do {                                                     \
    if (cout <= 32) SS_LAUNCH_(TI, TO, 1);                 \
    else if (cout <= 64) SS_LAUNCH_(TI, TO, 2);            \
    else if (cout <= 128) SS_LAUNCH_(TI, TO, 4);           \
    else if (cout <= 256) SS_LAUNCH_(TI, TO, 8);           \
    else if (cout <= 512) SS_LAUNCH_(TI, TO, 16);          \
    else SS_LAUNCH_(TI, TO, 32);                           \
  } while (0)
  if (in_is_bf16 && out_is_bf16) SS_BY_U_(__nv_bfloat16, __nv_bfloat16);
  else if (in_is_bf16) SS_BY_U_(__nv_bfloat16, float);
  else if (out_is_bf16) SS_BY_U_(float, __nv_bfloat16);
  else SS_BY_U_(float, float);
#undef SS_BY_U_
#undef SS_LAUNCH_
  SS_CHECK_LAUNCH();
  return SS_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Weight gradient of the stem-like convs (tiny Cin, Cout = 32; training):
//     dw[t][ci][co] = sum_p in[nbr[t][p]][ci] * dy[p][co]
// Replaces the autograd of spconv.SubMConv3d(11 -> 32, k = 5) w.r.t. its weight (point_transformer_v3m1_base.py:499-506).
// Persistent CTAs walk blocks of 16 voxels: the block's gathered inputs XN[t][p][ci] (zero where the tap has no
// neighbour) and dy rows are staged in shared memory, then every thread owns 4 (tap, ci) pairs with their 32 output
// channels in registers (128 accumulators) and adds the block in.  All k^3 taps are evaluated densely: 7 GFMA at
// 164 k voxels, which is nothing next to the latency of the gathers.  Deterministic: per-CTA partial sums, then one
// thread per weight adds the partials in CTA order.
namespace ss {
constexpr int kSwP = 16;         // voxels per block
constexpr int kSwThreads = 384;  // x 4 pairs = 1536 >= 125 * 11
constexpr int kSwPairs = 4;

__global__ void __launch_bounds__(kSwThreads, 1)
stem_wgrad_partial_kernel(const float* __restrict__ in, const float* __restrict__ dy, const int32_t* __restrict__ nbr,
                          int64_t n, int k3, int cin, float* __restrict__ partial) {
  extern __shared__ float sw_smem[];
  float* xn = sw_smem;                     // [k3][kSwP][cin]
  float* dys = sw_smem + k3 * kSwP * cin;  // [kSwP][32]
  const int npairs = k3 * cin;
  float acc[kSwPairs][32];
#pragma unroll
  for (int a = 0; a < kSwPairs; ++a)
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[a][c] = 0.f;
  const int64_t nblk = (n + kSwP - 1) / kSwP;
  for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int64_t p0 = blk * kSwP;
    for (int idx = threadIdx.x; idx < k3 * kSwP; idx += kSwThreads) {
      const int t = idx / kSwP, pl = idx - t * kSwP;
      const int64_t p = p0 + pl;
      const int q = p < n ? nbr[(size_t)t * n + p] : -1;
      float* dst = xn + (size_t)idx * cin;
      if (q >= 0) {
        const float* src = in + (size_t)q * cin;
        for (int c = 0; c < cin; ++c) dst[c] = src[c];
      } else {
        for (int c = 0; c < cin; ++c) dst[c] = 0.f;
      }
    }
    for (int idx = threadIdx.x; idx < kSwP * 32; idx += kSwThreads) {
      const int64_t p = p0 + idx / 32;
      dys[idx] = p < n ? dy[(size_t)p * 32 + (idx & 31)] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < kSwPairs; ++a) {
      const int pair = threadIdx.x + a * kSwThreads;
      if (pair < npairs) {
        const int t = pair / cin, ci = pair - t * cin;
        const float* xr = xn + (size_t)t * kSwP * cin + ci;
#pragma unroll 4
        for (int pl = 0; pl < kSwP; ++pl) {
          const float v = xr[pl * cin];
          if (v != 0.f) {
            const float4* d4 = reinterpret_cast<const float4*>(dys + pl * 32);
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const float4 d = d4[c4];
              acc[a][4 * c4] = fmaf(v, d.x, acc[a][4 * c4]);
              acc[a][4 * c4 + 1] = fmaf(v, d.y, acc[a][4 * c4 + 1]);
              acc[a][4 * c4 + 2] = fmaf(v, d.z, acc[a][4 * c4 + 2]);
              acc[a][4 * c4 + 3] = fmaf(v, d.w, acc[a][4 * c4 + 3]);
            }
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < kSwPairs; ++a) {
    const int pair = threadIdx.x + a * kSwThreads;
    if (pair < npairs) {
      float4* dst = reinterpret_cast<float4*>(partial + ((size_t)blockIdx.x * npairs + pair) * 32);
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4)
        dst[c4] = make_float4(acc[a][4 * c4], acc[a][4 * c4 + 1], acc[a][4 * c4 + 2], acc[a][4 * c4 + 3]);
    }
  }
}

__global__ void __launch_bounds__(256)
stem_wgrad_final_kernel(const float* __restrict__ partial, int blocks, int64_t nw, float* __restrict__ dw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nw) return;
  float t = 0.f;
  for (int b = 0; b < blocks; ++b) t += partial[(size_t)b * nw + i];
  dw[i] = t;
}
}  // namespace ss

extern "C" size_t ss_stem_conv_wgrad_workspace_bytes(int k3, int cin) {
  return (k3 < 1 || cin < 1) ? 0 : (size_t)ss::kNumSMs * k3 * cin * 32 * sizeof(float);
}

extern "C" int ss_stem_conv_wgrad(const float* in, const float* dy, const int32_t* nbr, int64_t n, int k3, int cin, int cout,
                                  float* dw, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || k3 < 1 || cin < 1 || cout != 32 || k3 * cin > ss::kSwThreads * ss::kSwPairs || !dw) return SS_BAD_ARGS;
  const size_t smem = ((size_t)k3 * ss::kSwP * cin + ss::kSwP * 32) * sizeof(float);
  if (smem > 200 * 1024) return SS_BAD_ARGS;
  const int64_t nw = (int64_t)k3 * cin * 32;
  if (n == 0) {
    SS_CUDA(cudaMemsetAsync(dw, 0, (size_t)nw * sizeof(float), stream));
    return SS_OK;
  }
  if (!in || !dy || !nbr || !workspace || (uintptr_t)workspace % 16 != 0) return SS_BAD_ARGS;
  if (workspace_bytes < ss_stem_conv_wgrad_workspace_bytes(k3, cin)) return SS_BAD_ARGS;
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, ss::kSwP), ss::kNumSMs);
  SS_CUDA(cudaFuncSetAttribute(ss::stem_wgrad_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ss::stem_wgrad_partial_kernel<<<blocks, ss::kSwThreads, smem, stream>>>(in, dy, nbr, n, k3, cin, (float*)workspace);
  SS_CHECK_LAUNCH();
  ss::stem_wgrad_final_kernel<<<(unsigned)ss::ceil_div64(nw, 256), 256, 0, stream>>>((const float*)workspace, blocks, nw, dw);
  SS_CHECK_LAUNCH();
  return SS_OK;
}
