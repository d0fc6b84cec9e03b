// Adjoint kernels of the training path.
//
// LayerNorm backward (the reference trains through torch's nn.LayerNorm: Block.norm1 / norm2 / cpe[2],
// point_transformer_v3m1_base.py:277-338):  y = (x - mean) * rstd * gamma + beta over the C channels of a row;
//   g = dy * gamma,  dx = rstd * (g - mean_C(g) - xhat * mean_C(g * xhat)),  dgamma += dy * xhat,  dbeta += dy.
// One warp per row, 16-byte accesses, mean / rstd recomputed from x (nothing saved by the forward), the
// per-channel parameter gradients accumulated in registers over the rows of a warp, combined per block in shared
// memory and added to global memory with one atomic per channel per block: one pass over x and dy.
#include "common.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

template <typename T> __device__ __forceinline__ void b_ld8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void b_ld8<float>(const float* p, float (&v)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void b_ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
template <typename T> __device__ __forceinline__ void b_st8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void b_st8<float>(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void b_st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ float b_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename TX, typename TDY, int NCH>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const TX* __restrict__ x, const TDY* __restrict__ dy, const float* __restrict__ gamma, float eps,
                     int64_t n, int C, TX* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float s_acc[];  // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float invC = 1.f / (float)C;
  float gm[NCH][8], adg[NCH][8], adb[NCH][8];
#pragma unroll
  for (int u = 0; u < NCH; ++u) {
    const int c = u * 256 + lane * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) gm[u][e] = adg[u][e] = adb[u][e] = 0.f;
    if (c < C) b_ld8<float>(gamma + c, gm[u]);
  }
  for (int64_t r = warp0; r < n; r += nwarp) {
    float xv[NCH][8], dv[NCH][8];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < NCH; ++u) {
      const int c = u * 256 + lane * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) xv[u][e] = dv[u][e] = 0.f;
      if (c < C) {
        b_ld8<TX>(x + (size_t)r * C + c, xv[u]);
        b_ld8<TDY>(dy + (size_t)r * C + c, dv[u]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s += xv[u][e];
    }
    const float mean = b_warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int u = 0; u < NCH; ++u) {
      const bool ok = u * 256 + lane * 8 < C;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = ok ? xv[u][e] - mean : 0.f;
        xv[u][e] = d;
        q += d * d;
      }
    }
    const float rstd = rsqrtf(b_warp_sum(q) * invC + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int u = 0; u < NCH; ++u)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = xv[u][e] * rstd, g = dv[u][e] * gm[u][e];
        xv[u][e] = xh;
        s1 += g;
        s2 += g * xh;
        adg[u][e] += dv[u][e] * xh;
        adb[u][e] += dv[u][e];
      }
    s1 = b_warp_sum(s1) * invC;
    s2 = b_warp_sum(s2) * invC;
#pragma unroll
    for (int u = 0; u < NCH; ++u) {
      const int c = u * 256 + lane * 8;
      if (c < C) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = rstd * (dv[u][e] * gm[u][e] - s1 - xv[u][e] * s2);
        b_st8<TX>(dx + (size_t)r * C + c, o);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < NCH; ++u) {
    const int c = u * 256 + lane * 8;
    if (c < C) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(&s_acc[c + e], adg[u][e]);
        atomicAdd(&s_acc[C + c + e], adb[u][e]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&dgamma[i], s_acc[i]);
    atomicAdd(&dbeta[i], s_acc[C + i]);
  }
}

}  // namespace ss

extern "C" int ss_layernorm_backward(const void* x, int x_is_bf16, const void* dy, int dy_is_bf16, const float* gamma,
                                     float eps, int64_t n, int channels, void* dx, float* dgamma, float* dbeta,
                                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || channels < 8 || channels % 8 != 0 || channels > 1024) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!x || !dy || !gamma || !dx || !dgamma || !dbeta) return SS_BAD_ARGS;
  if (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)gamma) % 16 != 0) return SS_BAD_ARGS;
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, 8), 2 * ss::kNumSMs);
  const size_t smem = (size_t)2 * channels * sizeof(float);
  const int nch = (channels + 255) / 256;
#define SS_LNB_(TX, TDY, N)                                                                                           \
  ss::layernorm_bwd_kernel<TX, TDY, N><<<blocks, 256, smem, stream>>>((const TX*)x, (const TDY*)dy, gamma, eps, n, channels, \
                                                                      (TX*)dx, dgamma, dbeta)
#define SS_LNB_N_(TX, TDY)              \
  do {                                  \
    if (nch == 1) SS_LNB_(TX, TDY, 1);  \
    else if (nch == 2) SS_LNB_(TX, TDY, 2); \
    else if (nch == 3) SS_LNB_(TX, TDY, 3); \
    else SS_LNB_(TX, TDY, 4);           \
  } while (0)
  if (x_is_bf16 && dy_is_bf16) SS_LNB_N_(__nv_bfloat16, __nv_bfloat16);
  else if (x_is_bf16) SS_LNB_N_(__nv_bfloat16, float);
  else if (dy_is_bf16) SS_LNB_N_(float, __nv_bfloat16);
  else SS_LNB_N_(float, float);
#undef SS_LNB_N_
#undef SS_LNB_
  SS_CHECK_LAUNCH();
  return SS_OK;
}

// GELU (erf form) backward: dx = dy * (Phi(x) + x * phi(x)), bf16 in / bf16 out, 8 elements per thread.
namespace ss {
__global__ void __launch_bounds__(256)
gelu_bwd_bf16x8_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ dx, int64_t nvec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    uint4 xv = x[i], gv = dy[i], ov;
    const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&xv);
    const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gv);
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 a = __bfloat1622float2(xh[u]), g = __bfloat1622float2(gh[u]);
      float2 o;
      {
        const float cdf = 0.5f * (1.f + erff(a.x * 0.70710678118654752440f));
        o.x = g.x * (cdf + a.x * 0.3989422804014327f * __expf(-0.5f * a.x * a.x));
      }
      {
        const float cdf = 0.5f * (1.f + erff(a.y * 0.70710678118654752440f));
        o.y = g.y * (cdf + a.y * 0.3989422804014327f * __expf(-0.5f * a.y * a.y));
      }
      oh[u] = __floats2bfloat162_rn(o.x, o.y);
    }
    dx[i] = ov;
  }
}
}  // namespace ss

extern "C" int ss_gelu_backward_bf16(const void* x, const void* dy, int64_t n_elements, void* dx, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_elements < 0 || n_elements % 8 != 0) return SS_BAD_ARGS;
  if (n_elements == 0) return SS_OK;
  if (!x || !dy || !dx || ((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx) % 16 != 0) return SS_BAD_ARGS;
  const int64_t nvec = n_elements / 8;
  const int blocks = (int)ss::imin64(ss::ceil_div64(nvec, 256), 32 * ss::kNumSMs);
  ss::gelu_bwd_bf16x8_kernel<<<blocks, 256, 0, stream>>>((const uint4*)x, (const uint4*)dy, (uint4*)dx, nvec);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

// Column sums of a bf16 matrix (bias gradients of the Linear layers: db = sum_rows dy), fp32 accumulation, deterministic
// (two stages: per-block partial rows, then one thread per column adds the partials in block order).
namespace ss {
constexpr int kColsumCols = 1024;  // columns per blockIdx.y
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const __nv_bfloat16* __restrict__ x, int64_t n, int c, float* __restrict__ partial) {
  __shared__ float red[8][kColsumCols];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * kColsumCols;
  const int cw = min(kColsumCols, c - c0);  // columns of this slab (multiple of 8)
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[i][u] = 0.f;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + warp, wn = (int64_t)gridDim.x * 8;
  // two rows per iteration: 8 independent 16-byte loads in flight per lane
  for (int64_t r = w0; r < n; r += 2 * wn) {
    const __nv_bfloat16* row0 = x + (size_t)r * c + c0;
    const bool two = r + wn < n;
    const __nv_bfloat16* row1 = two ? row0 + (size_t)wn * c : row0;
    uint4 q0[4], q1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int v = (lane + 32 * i) * 8;
      q0[i] = q1[i] = make_uint4(0u, 0u, 0u, 0u);  // bf16 zeros
      if (v < cw) {
        q0[i] = *reinterpret_cast<const uint4*>(row0 + v);
        if (two) q1[i] = *reinterpret_cast<const uint4*>(row1 + v);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&q0[i]);
      const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&q1[i]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f0 = __bfloat1622float2(h0[u]), f1 = __bfloat1622float2(h1[u]);
        acc[i][2 * u] += f0.x + f1.x;
        acc[i][2 * u + 1] += f0.y + f1.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int u = 0; u < 8; ++u) red[warp][(lane + 32 * i) * 8 + u] = acc[i][u];
  __syncthreads();
  for (int col = threadIdx.x; col < cw; col += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][col];
    partial[(size_t)blockIdx.x * c + c0 + col] = t;
  }
}
__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ partial, int blocks, int c, float* __restrict__ out) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= c) return;
  float t = 0.f;
  for (int b = 0; b < blocks; ++b) t += partial[(size_t)b * c + col];
  out[col] = t;
}
}  // namespace ss

extern "C" size_t ss_colsum_workspace_bytes(int channels) {
  return channels < 1 ? 0 : (size_t)2 * ss::kNumSMs * channels * sizeof(float);
}

extern "C" int ss_colsum_bf16(const void* x, int64_t n, int channels, float* out, void* workspace, size_t workspace_bytes,
                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || channels < 8 || channels % 8 != 0 || !out) return SS_BAD_ARGS;
  if (n == 0) {
    SS_CUDA(cudaMemsetAsync(out, 0, (size_t)channels * sizeof(float), stream));
    return SS_OK;
  }
  if (!x || !workspace || ((uintptr_t)x | (uintptr_t)workspace) % 16 != 0) return SS_BAD_ARGS;
  if (workspace_bytes < ss_colsum_workspace_bytes(channels)) return SS_BAD_ARGS;
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, 8), 2 * ss::kNumSMs);
  dim3 grid((unsigned)blocks, (unsigned)((channels + ss::kColsumCols - 1) / ss::kColsumCols));
  ss::colsum_partial_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, n, channels, (float*)workspace);
  SS_CHECK_LAUNCH();
  ss::colsum_final_kernel<<<(channels + 255) / 256, 256, 0, stream>>>((const float*)workspace, blocks, channels, out);
  SS_CHECK_LAUNCH();
  return SS_OK;
}
