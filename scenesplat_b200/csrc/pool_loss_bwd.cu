// Adjoints of the pooling / unpooling reductions and of the language losses (training path, SURVEY.md 8b export list).
//
//   ss_segment_mean_bwd          d src of SerializedPooling's segment mean / sum (reference: autograd through
//                                torch_scatter.segment_csr, point_transformer_v3m1_base.py:416-418):
//                                dsrc[p, :] = dout[cluster[p], :] * (1 / count[cluster[p]] for "mean")
//   ss_unpool_gather_add_bwd     d child of SerializedUnpooling's gather (ref :478, autograd of point.feat[inverse]):
//                                dchild[m, :] = sum over the members p of cluster m of dout[p, :]  (a segment sum along
//                                the parent's first serialized order: deterministic, no atomics)
//   ss_cos_l2_loss_bwd           d pred of  w_c * mean_valid(1 - cos(pred, target)) + w_l * mean_valid ||pred - target||^2
//                                (pointcept/models/losses/misc.py:254-295), one pass, the valid count and the upstream
//                                gradient read from device memory (no host sync)
//   ss_class_half_sums_bwd       d pred of the per-(class, half) sums of AggregatedContrastiveLoss (misc.py:384-385)
// All HBM-bound row kernels with 16-byte accesses.
#include "common.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

template <typename T> __device__ __forceinline__ void p_ld8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void p_ld8<float>(const float* p, float (&v)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void p_ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
template <> __device__ __forceinline__ void p_ld8<__half>(const __half* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
template <typename T> __device__ __forceinline__ void p_st8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void p_st8<float>(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void p_st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// One thread per (row, 8-channel piece); consecutive threads walk the pieces of a row.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
segment_mean_bwd_kernel(const TI* __restrict__ dout, const int64_t* __restrict__ cluster,
                        const int64_t* __restrict__ seg_start, int64_t n, int C, int mean, TO* __restrict__ dsrc) {
  const int pieces = C >> 3;
  const int64_t total = n * pieces;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / pieces;
    const int c = (int)(i - p * pieces) * 8;
    const int64_t m = cluster[p];
    float v[8];
    p_ld8(dout + (size_t)m * C + c, v);
    if (mean) {
      const float inv = 1.f / (float)max((long long)(seg_start[m + 1] - seg_start[m]), 1ll);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= inv;
    }
    p_st8(dsrc + (size_t)p * C + c, v);
  }
}

// One warp per row: dot / norms by shuffle, then the gradient row.
template <typename TP, typename TT>
__global__ void __launch_bounds__(256)
cos_l2_bwd_kernel(const TP* __restrict__ pred, const TT* __restrict__ target, const uint8_t* __restrict__ mask,
                  const double* __restrict__ acc3, const float* __restrict__ gout, float w_cos, float w_l2, int64_t n, int C,
                  float* __restrict__ dpred) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float g = gout ? *gout : 1.f;
  const float inv_n = 1.f / fmaxf((float)acc3[2], 1.f);
  const float sc = g * w_cos * inv_n, sl = g * w_l2 * inv_n * 2.f;
  const int pieces = C >> 3;
  for (int64_t r = warp0; r < n; r += nwarp) {
    float* drow = dpred + (size_t)r * C;
    if (!mask[r]) {
      for (int c = lane; c < pieces; c += 32) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        p_st8(drow + c * 8, z);
      }
      continue;
    }
    float dot = 0.f, pp = 0.f, tt = 0.f;
    for (int c = lane; c < pieces; c += 32) {
      float a[8], b[8];
      p_ld8(pred + (size_t)r * C + c * 8, a);
      p_ld8(target + (size_t)r * C + c * 8, b);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        dot = fmaf(a[e], b[e], dot);
        pp = fmaf(a[e], a[e], pp);
        tt = fmaf(b[e], b[e], tt);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dot += __shfl_xor_sync(0xffffffffu, dot, o);
      pp += __shfl_xor_sync(0xffffffffu, pp, o);
      tt += __shfl_xor_sync(0xffffffffu, tt, o);
    }
    // cos = dot / (max(|p|, eps) max(|t|, eps));  d(1 - cos)/dp = -(t / (|p||t|) - cos p / |p|^2)
    const float np = fmaxf(sqrtf(pp), 1e-8f), nt = fmaxf(sqrtf(tt), 1e-8f);
    const float inv_pt = 1.f / (np * nt);
    const float ct = -sc * inv_pt;                    // coefficient of t
    const float cp = sc * dot * inv_pt / (np * np);   // coefficient of p (from the cosine term)
    for (int c = lane; c < pieces; c += 32) {
      float a[8], b[8], d[8];
      p_ld8(pred + (size_t)r * C + c * 8, a);
      p_ld8(target + (size_t)r * C + c * 8, b);
#pragma unroll
      for (int e = 0; e < 8; ++e) d[e] = ct * b[e] + cp * a[e] + sl * (a[e] - b[e]);
      p_st8(drow + c * 8, d);
    }
  }
}

__global__ void __launch_bounds__(256)
class_half_sums_bwd_kernel(const float* __restrict__ dsums, const uint8_t* __restrict__ mask,
                           const int64_t* __restrict__ segment, const int64_t* __restrict__ half, int64_t n, int C,
                           int n_classes, float* __restrict__ dpred) {
  const int pieces = C >> 3;
  const int64_t total = n * pieces;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / pieces;
    const int c = (int)(i - p * pieces) * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int64_t s = segment[p];
    if (mask[p] && s >= 0 && s < n_classes) p_ld8(dsums + (size_t)(s * 2 + (half[p] != 0)) * C + c, v);
    p_st8(dpred + (size_t)p * C + c, v);
  }
}

}  // namespace ss

extern "C" {

int ss_segment_mean_bwd(const void* dout, int dout_is_bf16, const int64_t* cluster, const int64_t* seg_start, int64_t n,
                        int channels, int reduce, void* dsrc, int dsrc_is_bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || channels < 8 || channels % 8 != 0 || (reduce != 0 && reduce != 1)) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!dout || !cluster || !seg_start || !dsrc) return SS_BAD_ARGS;
  if (((uintptr_t)dout | (uintptr_t)dsrc) % 16 != 0) return SS_BAD_ARGS;
  const int blocks = (int)ss::imin64(ss::ceil_div64(n * (channels / 8), 256), 32 * ss::kNumSMs);
#define SS_SMB_(TI, TO) \
  ss::segment_mean_bwd_kernel<TI, TO><<<blocks, 256, 0, stream>>>((const TI*)dout, cluster, seg_start, n, channels, reduce, (TO*)dsrc)
  if (dout_is_bf16 && dsrc_is_bf16) SS_SMB_(__nv_bfloat16, __nv_bfloat16);
  else if (dout_is_bf16) SS_SMB_(__nv_bfloat16, float);
  else if (dsrc_is_bf16) SS_SMB_(float, __nv_bfloat16);
  else SS_SMB_(float, float);
#undef SS_SMB_
  SS_CHECK_LAUNCH();
  return SS_OK;
}

int ss_unpool_gather_add_bwd(const void* dout, int dout_is_bf16, const int64_t* order0, const int64_t* seg_start, int64_t m,
                             int channels, void* dchild, int dchild_is_bf16, void* stream) {
  // the adjoint of a gather by cluster id is the segment SUM over each cluster's members
  return ss_segment_reduce(dout, dout_is_bf16, order0, seg_start, nullptr, m, channels, 0, nullptr, nullptr, 0, dchild,
                           dchild_is_bf16, stream);
}

int ss_cos_l2_loss_bwd(const void* pred, int pred_is_bf16, const void* target, int target_dtype, const uint8_t* mask,
                       int64_t n, int channels, const double* acc3, const float* grad_out, float w_cos, float w_l2,
                       float* dpred, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || channels < 8 || channels % 8 != 0 || target_dtype < 0 || target_dtype > 2) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!pred || !target || !mask || !acc3 || !dpred) return SS_BAD_ARGS;
  if (((uintptr_t)pred | (uintptr_t)target | (uintptr_t)dpred) % 16 != 0) return SS_BAD_ARGS;
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, 8), 32 * ss::kNumSMs);
#define SS_CLB_(TP, TT)                                                                                                   \
  ss::cos_l2_bwd_kernel<TP, TT><<<blocks, 256, 0, stream>>>((const TP*)pred, (const TT*)target, mask, acc3, grad_out, w_cos, \
                                                            w_l2, n, channels, dpred)
#define SS_CLB_T_(TP)                                   \
  do {                                                  \
    if (target_dtype == 0) SS_CLB_(TP, float);          \
    else if (target_dtype == 1) SS_CLB_(TP, __nv_bfloat16); \
    else SS_CLB_(TP, __half);                           \
  } while (0)
  if (pred_is_bf16) SS_CLB_T_(__nv_bfloat16);
  else SS_CLB_T_(float);
#undef SS_CLB_T_
#undef SS_CLB_
  SS_CHECK_LAUNCH();
  return SS_OK;
}

int ss_class_half_sums_bwd(const float* dsums, const uint8_t* mask, const int64_t* segment, const int64_t* half, int64_t n,
                           int channels, int n_classes, float* dpred, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || channels < 8 || channels % 8 != 0 || n_classes < 1) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!dsums || !mask || !segment || !half || !dpred) return SS_BAD_ARGS;
  if (((uintptr_t)dsums | (uintptr_t)dpred) % 16 != 0) return SS_BAD_ARGS;
  const int blocks = (int)ss::imin64(ss::ceil_div64(n * (channels / 8), 256), 32 * ss::kNumSMs);
  ss::class_half_sums_bwd_kernel<<<blocks, 256, 0, stream>>>(dsums, mask, segment, half, n, channels, n_classes, dpred);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // extern "C"
