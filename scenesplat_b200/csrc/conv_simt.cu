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
  const int cin = EXACT ? CIN : cin_rt;
  const int lane = threadIdx.x & 31;
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
      float x[CIN], w[CIN];
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        x[ci] = (q >= 0 && (EXACT || ci < cin)) ? cvt_in<TI>(in[(size_t)q * cin + ci]) : 0.f;
        w[ci] = ((EXACT || ci < cin) && lane < cout) ? __ldg(wt + ((size_t)t * cin + ci) * cout + lane) : 0.f;
      }
#pragma unroll
      for (int g4 = 0; g4 < 8; ++g4) {
        if (((mask >> (4 * g4)) & 0xfu) == 0u) continue;  // warp-uniform
#pragma unroll
        for (int i = 4 * g4; i < 4 * g4 + 4; ++i) {
          if ((mask >> i) & 1u) {
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci)
              if (EXACT || ci < cin) acc[i] = fmaf(__shfl_sync(0xffffffffu, x[ci], i), w[ci], acc[i]);
          }
        }
      }
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
