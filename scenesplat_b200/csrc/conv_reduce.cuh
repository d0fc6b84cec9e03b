// Per-voxel bodies of the xCPE conv's gather-sum stage, shared by the stand-alone kernels (conv_gemm.cu) and the reducer
// warps of the fused gather-GEMM (conv_gemm3.cu).  One warp per output voxel, 16-byte (8 x bf16) lanes, J x 256 channels.
#pragma once
#include "tc_common.cuh"

namespace ss {

// acc[j][u] = bias + sum over the voxel's active taps of prod[pos, :]: `mypos` is the lane's (= tap's) product row or -1;
// the active rows are walked in ascending tap order (fixed summation order) R at a time, so R row loads are in
// flight per lane.  CG: load the rows with ld.global.cs (the fused kernel reads rows another SM has just written, once).
template <int J, bool CG, int R = 4>
__device__ __forceinline__ void conv_gather_sum(const __nv_bfloat16* __restrict__ prod, int32_t mypos, int lane, int C,
                                                float (&acc)[J][8]) {
  uint32_t m = __ballot_sync(0xffffffffu, mypos >= 0);
  while (m) {
    int32_t pos[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int t = m ? __ffs(m) - 1 : 0;
      pos[q] = m ? __shfl_sync(0xffffffffu, mypos, t) : -1;
      m &= m - 1;  // (0 stays 0)
    }
    uint4 v[R][J];
#pragma unroll
    for (int q = 0; q < R; ++q)
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c0 = j * 256 + lane * 8;
        v[q][j] = make_uint4(0u, 0u, 0u, 0u);
        if (pos[q] >= 0 && c0 < C) {
          const uint4* src = reinterpret_cast<const uint4*>(prod + (size_t)pos[q] * C + c0);
          v[q][j] = CG ? __ldcs(src) : *src;  // streaming: L2 only, first to be evicted (the row is dead after this read)
        }
      }
#pragma unroll
    for (int q = 0; q < R; ++q) {
      if (pos[q] < 0) continue;  // warp-uniform; (adding the zero vector would turn a -0 sum into +0)
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[q][j]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 f = __bfloat1622float2(h[u]);
          acc[j][2 * u] += f.x;
          acc[j][2 * u + 1] += f.y;
        }
      }
    }
  }
}

template <int J>
__device__ __forceinline__ void conv_acc_init(const float* __restrict__ bias, int lane, int C, float (&acc)[J][8]) {
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c0 = j * 256 + lane * 8;
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[j][u] = (bias && c0 < C) ? bias[c0 + u] : 0.f;
  }
}

__device__ __forceinline__ float conv_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// z = acc (fp32, never rounded to bf16): y = res[p] + LN0(z) -> res_out[p] (fp32), LN1(y) -> norm_out[p] (bf16)
// (point_transformer_v3m1_base.py:318-326 with the conv's Linear folded into the taps).
template <int J>
__device__ __forceinline__ void conv_ln_res_ln_store(float (&acc)[J][8], int64_t p, int lane, int C, float invC, float eps,
                                                     const float* res, const float* __restrict__ g0,
                                                     const float* __restrict__ b0, const float* __restrict__ g1,
                                                     const float* __restrict__ b1, float* res_out,
                                                     __nv_bfloat16* __restrict__ norm_out) {
  // ---- LN0(z)
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < J; ++j)
#pragma unroll
    for (int u = 0; u < 8; ++u) s += acc[j][u];  // lanes past C hold zeros
  float mean = conv_warp_sum(s) * invC;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const bool ok = j * 256 + lane * 8 < C;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float d = ok ? acc[j][u] - mean : 0.f;
      q += d * d;
    }
  }
  float rstd = rsqrtf(conv_warp_sum(q) * invC + eps);
  // ---- y = res + LN0(z)
  s = 0.f;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c0 = j * 256 + lane * 8;
    if (c0 < C) {
      const float4 ga = *reinterpret_cast<const float4*>(g0 + c0), gb = *reinterpret_cast<const float4*>(g0 + c0 + 4);
      const float4 ba = *reinterpret_cast<const float4*>(b0 + c0), bb = *reinterpret_cast<const float4*>(b0 + c0 + 4);
      // streaming loads / stores for the residual and the outputs: they must not push the (re-read) products and
      // gathered input rows of the fused kernel out of L2
      const float4 xa = __ldcs(reinterpret_cast<const float4*>(res + (size_t)p * C + c0));
      const float4 xb = __ldcs(reinterpret_cast<const float4*>(res + (size_t)p * C + c0 + 4));
      const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
      const float bt[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
      const float xx[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc[j][u] = xx[u] + ((acc[j][u] - mean) * rstd * gg[u] + bt[u]);
        s += acc[j][u];
      }
      float4* o = reinterpret_cast<float4*>(res_out + (size_t)p * C + c0);
      __stcs(o, make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]));
      __stcs(o + 1, make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]));
    }
  }
  // ---- LN1(y) -> bf16
  mean = conv_warp_sum(s) * invC;
  q = 0.f;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const bool ok = j * 256 + lane * 8 < C;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float d = ok ? acc[j][u] - mean : 0.f;
      q += d * d;
    }
  }
  rstd = rsqrtf(conv_warp_sum(q) * invC + eps);
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c0 = j * 256 + lane * 8;
    if (c0 < C) {
      const float4 ga = *reinterpret_cast<const float4*>(g1 + c0), gb = *reinterpret_cast<const float4*>(g1 + c0 + 4);
      const float4 ba = *reinterpret_cast<const float4*>(b1 + c0), bb = *reinterpret_cast<const float4*>(b1 + c0 + 4);
      const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
      const float bt[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
      float o[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) o[u] = (acc[j][u] - mean) * rstd * gg[u] + bt[u];
      uint4 w;
      w.x = tc::pack_bf16(o[0], o[1]);
      w.y = tc::pack_bf16(o[2], o[3]);
      w.z = tc::pack_bf16(o[4], o[5]);
      w.w = tc::pack_bf16(o[6], o[7]);
      __stcs(reinterpret_cast<uint4*>(norm_out + (size_t)p * C + c0), w);
    }
  }
}

}  // namespace ss
