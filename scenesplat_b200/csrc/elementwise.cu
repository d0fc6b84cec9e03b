// Fused row-wise kernels around the GEMMs of a PTv3 Block (reference
// point_transformer_v3m1_base.py:318-338): residual add + (inner) LayerNorm + (outer) LayerNorm in
// one pass over the row, eval-mode BatchNorm folded to scale/shift + GELU, exact-erf GELU, and
// F.normalize (pointcept/models/default.py:98).  One warp per row; the row lives in registers.
#include "common.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

template <typename T> __device__ __forceinline__ float e_in(T v);
template <> __device__ __forceinline__ float e_in<float>(float v) { return v; }
template <> __device__ __forceinline__ float e_in<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T e_out(float v);
template <> __device__ __forceinline__ float e_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 e_out<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

__device__ __forceinline__ float e_gelu(float x) { return gelu_fast(x); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kMaxPerLane = 32;  // C <= 1024

// y = res + f(delta),  f = LayerNorm(g0, b0) when g0 != null else identity;  res may be null (y = f(delta)).
// Writes res_out (fp32, may alias res) and/or norm_out = LayerNorm(y; g1, b1) (or a plain cast of y when g1 == null).
template <typename TD, typename TN, int VPL>
__global__ void __launch_bounds__(256)
add_layernorm_kernel(const float* res, const TD* __restrict__ delta, const float* __restrict__ g0,
                     const float* __restrict__ b0, const float* __restrict__ g1, const float* __restrict__ b1, float eps,
                     int64_t n, int C, float* res_out, TN* __restrict__ norm_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float invC = 1.f / (float)C;
  for (int64_t r = warp0; r < n; r += nwarp) {
    float v[VPL];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < VPL; ++u) {
      const int c = lane + 32 * u;
      v[u] = (c < C && delta) ? e_in<TD>(delta[(size_t)r * C + c]) : 0.f;
      s += v[u];
    }
    if (g0) {
      const float mean = warp_sum(s) * invC;
      float q = 0.f;
#pragma unroll
      for (int u = 0; u < VPL; ++u) {
        const int c = lane + 32 * u;
        const float d = c < C ? v[u] - mean : 0.f;
        q += d * d;
      }
      const float rstd = rsqrtf(warp_sum(q) * invC + eps);
#pragma unroll
      for (int u = 0; u < VPL; ++u) {
        const int c = lane + 32 * u;
        if (c < C) v[u] = (v[u] - mean) * rstd * g0[c] + b0[c];
      }
    }
    s = 0.f;
#pragma unroll
    for (int u = 0; u < VPL; ++u) {
      const int c = lane + 32 * u;
      if (c < C) {
        if (res) v[u] += res[(size_t)r * C + c];
        if (res_out) res_out[(size_t)r * C + c] = v[u];
        s += v[u];
      }
    }
    if (norm_out) {
      if (g1) {
        const float mean = warp_sum(s) * invC;
        float q = 0.f;
#pragma unroll
        for (int u = 0; u < VPL; ++u) {
          const int c = lane + 32 * u;
          const float d = c < C ? v[u] - mean : 0.f;
          q += d * d;
        }
        const float rstd = rsqrtf(warp_sum(q) * invC + eps);
#pragma unroll
        for (int u = 0; u < VPL; ++u) {
          const int c = lane + 32 * u;
          if (c < C) norm_out[(size_t)r * C + c] = e_out<TN>((v[u] - mean) * rstd * g1[c] + b1[c]);
        }
      } else {
#pragma unroll
        for (int u = 0; u < VPL; ++u) {
          const int c = lane + 32 * u;
          if (c < C) norm_out[(size_t)r * C + c] = e_out<TN>(v[u]);
        }
      }
    }
  }
}


// ---- vectorised variant (C % 8 == 0): TPR threads cooperate on a row, each owning NCH chunks of 8
// contiguous channels (16-byte bf16 / 2 x 16-byte fp32 accesses).  All loads of a row are issued before
// any store (res and res_out may alias, so the compiler must not be left to interleave them).
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float2 f = __bfloat1622float2(h[u]);
    v[2 * u] = f.x;
    v[2 * u + 1] = f.y;
  }
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 o;
  __nv_bfloat162 h;
  h = __floats2bfloat162_rn(v[0], v[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(v[2], v[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(v[4], v[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(v[6], v[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
  *reinterpret_cast<uint4*>(p) = o;
}
template <int TPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = TPR / 2; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename TD, typename TN, int TPR, int NCH>
__global__ void __launch_bounds__(256)
add_layernorm_vec_kernel(const float* res, const TD* __restrict__ delta, const float* __restrict__ g0,
                         const float* __restrict__ b0, const float* __restrict__ g1, const float* __restrict__ b1,
                         float eps, int64_t n, int C, float* res_out, TN* __restrict__ norm_out, float l2_eps = 0.f,
                         __nv_bfloat16* __restrict__ shadow_out = nullptr) {
  constexpr int RPW = 32 / TPR;  // rows per warp
  const int lane = threadIdx.x & 31, t = lane % TPR, sub = lane / TPR;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float invC = 1.f / (float)C;
  const int nchunk = C >> 3;
  for (int64_t r0 = warp0 * RPW; r0 < n; r0 += nwarp * RPW) {
    const int64_t r = r0 + sub;
    const bool rok = r < n;
    float v[NCH][8], x[NCH][8];
#pragma unroll
    for (int u = 0; u < NCH; ++u) {
      const int c = t + TPR * u;
      const bool ok = rok && c < nchunk;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[u][e] = x[u][e] = 0.f;
      if (ok && delta) ld8(delta + (size_t)r * C + c * 8, v[u]);
      if (ok && res) ld8(res + (size_t)r * C + c * 8, x[u]);
    }
    if (g0) {
      float s = 0.f;
#pragma unroll
      for (int u = 0; u < NCH; ++u)
#pragma unroll
        for (int e = 0; e < 8; ++e) s += v[u][e];
      const float mean = group_sum<TPR>(s) * invC;
      float q = 0.f;
#pragma unroll
      for (int u = 0; u < NCH; ++u) {
        const bool ok = (t + TPR * u) < nchunk;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = ok ? v[u][e] - mean : 0.f;
          q += d * d;
        }
      }
      const float rstd = rsqrtf(group_sum<TPR>(q) * invC + eps);
#pragma unroll
      for (int u = 0; u < NCH; ++u) {
        const int c = t + TPR * u;
        if (c < nchunk) {
          float gg[8], bb[8];
          ld8(g0 + c * 8, gg);
          ld8(b0 + c * 8, bb);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[u][e] = (v[u][e] - mean) * rstd * gg[e] + bb[e];
        }
      }
    }
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < NCH; ++u) {
      const int c = t + TPR * u;
      const bool ok = rok && c < nchunk;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[u][e] += x[u][e];
        s += v[u][e];
      }
      if (ok && res_out) st8(res_out + (size_t)r * C + c * 8, v[u]);
    }
    if (norm_out) {
      float mean = 0.f, rstd = 1.f;
      if (g1) {
        mean = group_sum<TPR>(s) * invC;
        float q = 0.f;
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
          const bool ok = (t + TPR * u) < nchunk;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float d = ok ? v[u][e] - mean : 0.f;
            q += d * d;
          }
        }
        rstd = rsqrtf(group_sum<TPR>(q) * invC + eps);
      } else if (l2_eps > 0.f) {  // F.normalize(y, p=2, dim=1, eps): y / max(|y|_2, eps)
        float q = 0.f;
#pragma unroll
        for (int u = 0; u < NCH; ++u)
#pragma unroll
          for (int e = 0; e < 8; ++e) q += v[u][e] * v[u][e];  // (lanes past C hold zeros)
        rstd = 1.f / fmaxf(sqrtf(group_sum<TPR>(q)), l2_eps);
      }
#pragma unroll
      for (int u = 0; u < NCH; ++u) {
        const int c = t + TPR * u;
        if (rok && c < nchunk) {
          float o[8];
          if (g1) {
            float gg[8], bb[8];
            ld8(g1 + c * 8, gg);
            ld8(b1 + c * 8, bb);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = (v[u][e] - mean) * rstd * gg[e] + bb[e];
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = v[u][e] * rstd;  // (rstd = 1 unless L2-normalising)
          }
          st8(norm_out + (size_t)r * C + c * 8, o);
          if (shadow_out) st8(shadow_out + (size_t)r * C + c * 8, o);  // bf16 copy of the same row (operand of the head GEMM)
        }
      }
    }
  }
}

// out = act(x * scale[c] + shift[c]);  act: 0 none, 1 GELU(erf).  scale/shift nullable (pure activation).
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
affine_act_kernel(const TI* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                  int64_t total, int C, TO* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float v = e_in<TI>(x[i]);
    if (scale) {
      const int c = (int)(i % C);
      v = v * scale[c] + shift[c];
    }
    if (act == 1) v = e_gelu(v);
    out[i] = e_out<TO>(v);
  }
}

// bf16 x 8 vectorised GELU (the MLP hidden activation, N x 4C elements)
__global__ void __launch_bounds__(256) gelu_bf16x8_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int64_t nvec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    uint4 v = x[i];
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float2 f = __bfloat1622float2(h[u]);
      f.x = e_gelu(f.x);
      f.y = e_gelu(f.y);
      h[u] = __floats2bfloat162_rn(f.x, f.y);
    }
    out[i] = v;
  }
}

// out[r, :] = x[r, :] / max(||x[r, :]||_2, eps)   (F.normalize, p = 2, dim = 1)
template <typename TI, typename TO, int VPL>
__global__ void __launch_bounds__(256)
l2_normalize_kernel(const TI* __restrict__ x, int64_t n, int C, float eps, TO* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n; r += nwarp) {
    float v[VPL];
    float q = 0.f;
#pragma unroll
    for (int u = 0; u < VPL; ++u) {
      const int c = lane + 32 * u;
      v[u] = c < C ? e_in<TI>(x[(size_t)r * C + c]) : 0.f;
      q += v[u] * v[u];
    }
    const float inv = 1.f / fmaxf(sqrtf(warp_sum(q)), eps);
#pragma unroll
    for (int u = 0; u < VPL; ++u) {
      const int c = lane + 32 * u;
      if (c < C) out[(size_t)r * C + c] = e_out<TO>(v[u] * inv);
    }
  }
}

}  // namespace ss

extern "C" {

int ss_add_layernorm(const float* res, const void* delta, int delta_is_bf16, const float* g0, const float* b0,
                     const float* g1, const float* b1, float eps, int64_t n, int channels, float* res_out,
                     void* norm_out, int norm_is_bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || channels < 1 || channels > 32 * ss::kMaxPerLane || (g0 && !b0) || (g1 && !b1)) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if ((!res && !delta) || (!res_out && !norm_out)) return SS_BAD_ARGS;
  const bool aligned = (((uintptr_t)res | (uintptr_t)delta | (uintptr_t)res_out | (uintptr_t)norm_out | (uintptr_t)g0 |
                         (uintptr_t)b0 | (uintptr_t)g1 | (uintptr_t)b1) % 16) == 0;
  if (channels % 8 == 0 && aligned) {
    const int nchunk = channels / 8;
    int tpr = 1;
    while (tpr < 32 && tpr < nchunk) tpr <<= 1;
    const int nch = (nchunk + tpr - 1) / tpr;
    const int rpw = 32 / tpr;
    const int vblocks = (int)ss::imin64(ss::ceil_div64(n, 8 * rpw), 16 * ss::kNumSMs);
#define SS_LNV_(TD, TN, T, N)                                                                                        \
  ss::add_layernorm_vec_kernel<TD, TN, T, N><<<vblocks, 256, 0, stream>>>(res, (const TD*)delta, g0, b0, g1, b1, eps, \
                                                                          n, channels, res_out, (TN*)norm_out)
#define SS_LNV_T_(TD, TN)                                        \
  do {                                                           \
    if (tpr == 1) SS_LNV_(TD, TN, 1, 1);                         \
    else if (tpr == 2) SS_LNV_(TD, TN, 2, 1);                    \
    else if (tpr == 4) SS_LNV_(TD, TN, 4, 1);                    \
    else if (tpr == 8) SS_LNV_(TD, TN, 8, 1);                    \
    else if (tpr == 16) SS_LNV_(TD, TN, 16, 1);                  \
    else if (nch == 1) SS_LNV_(TD, TN, 32, 1);                   \
    else if (nch == 2) SS_LNV_(TD, TN, 32, 2);                   \
    else if (nch == 3) SS_LNV_(TD, TN, 32, 3);                   \
    else SS_LNV_(TD, TN, 32, 4);                                 \
  } while (0)
    if (delta_is_bf16 && norm_is_bf16) SS_LNV_T_(__nv_bfloat16, __nv_bfloat16);
    else if (delta_is_bf16) SS_LNV_T_(__nv_bfloat16, float);
    else if (norm_is_bf16) SS_LNV_T_(float, __nv_bfloat16);
    else SS_LNV_T_(float, float);
#undef SS_LNV_T_
#undef SS_LNV_
    SS_CHECK_LAUNCH();
    return SS_OK;
  }
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, 8), 32 * ss::kNumSMs);
  const int vpl = (channels + 31) / 32;
#define SS_LN_(TD, TN, V)                                                                                         \
  ss::add_layernorm_kernel<TD, TN, V><<<blocks, 256, 0, stream>>>(res, (const TD*)delta, g0, b0, g1, b1, eps, n, \
                                                                  channels, res_out, (TN*)norm_out)
#define SS_LN_V_(TD, TN)                        \
  do {                                          \
    if (vpl <= 1) SS_LN_(TD, TN, 1);            \
    else if (vpl <= 2) SS_LN_(TD, TN, 2);       \
    else if (vpl <= 4) SS_LN_(TD, TN, 4);       \
    else if (vpl <= 8) SS_LN_(TD, TN, 8);       \
    else if (vpl <= 16) SS_LN_(TD, TN, 16);     \
    else if (vpl <= 24) SS_LN_(TD, TN, 24);     \
    else SS_LN_(TD, TN, 32);                    \
  } while (0)
  if (delta_is_bf16 && norm_is_bf16) SS_LN_V_(__nv_bfloat16, __nv_bfloat16);
  else if (delta_is_bf16) SS_LN_V_(__nv_bfloat16, float);
  else if (norm_is_bf16) SS_LN_V_(float, __nv_bfloat16);
  else SS_LN_V_(float, float);
#undef SS_LN_V_
#undef SS_LN_
  SS_CHECK_LAUNCH();
  return SS_OK;
}

// out = F.normalize(res + delta, p=2, dim=1, eps) in one pass: the last Block's residual add fused with the L2 normalisation
// LangPretrainer applies to the backbone output (models/default.py:98); the un-normalised sum is not written.  out_bf16
// (nullable) receives the same rows rounded to bf16: the operand of the zero-shot head's GEMM (saves the separate
// fp32 -> bf16 pass over [n, channels]).  channels % 8 == 0, <= 1024, 16-byte aligned pointers.
int ss_add_l2_normalize(const float* res, const void* delta, int delta_is_bf16, float eps, int64_t n, int channels,
                        float* out, void* out_bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || channels < 8 || channels % 8 != 0 || channels > 1024 || !(eps > 0.f)) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!res || !delta || !out) return SS_BAD_ARGS;
  if ((((uintptr_t)res | (uintptr_t)delta | (uintptr_t)out | (uintptr_t)out_bf16) % 16) != 0) return SS_BAD_ARGS;
  const int nchunk = channels / 8;
  int tpr = 1;
  while (tpr < 32 && tpr < nchunk) tpr <<= 1;
  const int nch = (nchunk + tpr - 1) / tpr;
  const int rpw = 32 / tpr;
  const int vblocks = (int)ss::imin64(ss::ceil_div64(n, 8 * rpw), 16 * ss::kNumSMs);
#define SS_L2_(TD, T, N)                                                                                            \
  ss::add_layernorm_vec_kernel<TD, float, T, N><<<vblocks, 256, 0, stream>>>(res, (const TD*)delta, nullptr, nullptr,   \
                                                                             nullptr, nullptr, 0.f, n, channels, nullptr, \
                                                                             out, eps, (__nv_bfloat16*)out_bf16)
#define SS_L2_T_(TD)                       \
  do {                                     \
    if (tpr == 1) SS_L2_(TD, 1, 1);        \
    else if (tpr == 2) SS_L2_(TD, 2, 1);   \
    else if (tpr == 4) SS_L2_(TD, 4, 1);   \
    else if (tpr == 8) SS_L2_(TD, 8, 1);   \
    else if (tpr == 16) SS_L2_(TD, 16, 1); \
    else if (nch == 1) SS_L2_(TD, 32, 1);  \
    else if (nch == 2) SS_L2_(TD, 32, 2);  \
    else if (nch == 3) SS_L2_(TD, 32, 3);  \
    else SS_L2_(TD, 32, 4);                \
  } while (0)
  if (delta_is_bf16) SS_L2_T_(__nv_bfloat16);
  else SS_L2_T_(float);
#undef SS_L2_T_
#undef SS_L2_
  SS_CHECK_LAUNCH();
  return SS_OK;
}

int ss_affine_act(const void* x, int in_is_bf16, const float* scale, const float* shift, int act, int64_t n,
                  int channels, void* out, int out_is_bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || channels < 1 || (scale && !shift)) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!x || !out) return SS_BAD_ARGS;
  const int64_t total = n * channels;
  if (in_is_bf16 && out_is_bf16 && !scale && act == 1 && total % 8 == 0 && ((uintptr_t)x | (uintptr_t)out) % 16 == 0) {
    const int64_t nvec = total / 8;
    const int blocks = (int)ss::imin64(ss::ceil_div64(nvec, 256), 32 * ss::kNumSMs);
    ss::gelu_bf16x8_kernel<<<blocks, 256, 0, stream>>>((const uint4*)x, (uint4*)out, nvec);
    SS_CHECK_LAUNCH();
    return SS_OK;
  }
  const int blocks = (int)ss::imin64(ss::ceil_div64(total, 256), 32 * ss::kNumSMs);
  if (in_is_bf16 && out_is_bf16)
    ss::affine_act_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, stream>>>(
        (const __nv_bfloat16*)x, scale, shift, act, total, channels, (__nv_bfloat16*)out);
  else if (in_is_bf16)
    ss::affine_act_kernel<__nv_bfloat16, float><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)x, scale, shift, act,
                                                                            total, channels, (float*)out);
  else if (out_is_bf16)
    ss::affine_act_kernel<float, __nv_bfloat16><<<blocks, 256, 0, stream>>>((const float*)x, scale, shift, act, total,
                                                                            channels, (__nv_bfloat16*)out);
  else
    ss::affine_act_kernel<float, float><<<blocks, 256, 0, stream>>>((const float*)x, scale, shift, act, total, channels,
                                                                    (float*)out);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

int ss_l2_normalize(const void* x, int in_is_bf16, int64_t n, int channels, float eps, void* out, int out_is_bf16,
                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || channels < 1 || channels > 1024) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!x || !out) return SS_BAD_ARGS;
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, 8), 32 * ss::kNumSMs);
  const int vpl = (channels + 31) / 32;
#define SS_NM_(TI, TO, V) \
  ss::l2_normalize_kernel<TI, TO, V><<<blocks, 256, 0, stream>>>((const TI*)x, n, channels, eps, (TO*)out)
#define SS_NM_V_(TI, TO)                     \
  do {                                       \
    if (vpl <= 4) SS_NM_(TI, TO, 4);         \
    else if (vpl <= 8) SS_NM_(TI, TO, 8);    \
    else if (vpl <= 16) SS_NM_(TI, TO, 16);  \
    else if (vpl <= 24) SS_NM_(TI, TO, 24);  \
    else SS_NM_(TI, TO, 32);                 \
  } while (0)
  if (in_is_bf16 && out_is_bf16) SS_NM_V_(__nv_bfloat16, __nv_bfloat16);
  else if (in_is_bf16) SS_NM_V_(__nv_bfloat16, float);
  else if (out_is_bf16) SS_NM_V_(float, __nv_bfloat16);
  else SS_NM_V_(float, float);
#undef SS_NM_V_
#undef SS_NM_
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // extern "C"
