// 768-d language head and language-pretraining losses.
//
// Replaces (reference):
//   zero-shot head  pointcept/engines/hooks/evaluator.py:793-800, pointcept/engines/test.py:335-349
//                   (torch.mm -> sigmoid -> max / accumulate; N x K logits never touch HBM here)
//   CosineSimilarity + L2Loss  pointcept/models/losses/misc.py:247-295 (one fused pass, no masked copies)
//   AggregatedContrastiveLoss's per-class half sums  pointcept/models/losses/misc.py:355-389
#include "common.cuh"
#include <cuda_fp16.h>
#include "../../include/scenesplat_b200.h"

namespace ss {

template <typename T> __device__ __forceinline__ float l_in(T v);
template <> __device__ __forceinline__ float l_in<float>(float v) { return v; }
template <> __device__ __forceinline__ float l_in<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float l_in<__half>(__half v) { return __half2float(v); }

__device__ __forceinline__ float l_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kLangVPL = 24;  // 768 / 32

// One warp per point: the (optionally L2-normalised) feature row lives in registers, the K text
// embeddings stream through L1/L2.  mode 0: write max prob + label (-1 below threshold);
// mode 1: probs_accum[idx[p], :] += sigmoid(logits) (the tester's fragment accumulation).
template <typename TF>
__global__ void __launch_bounds__(256)
lang_head_kernel(const TF* __restrict__ feat, const float* __restrict__ text, int64_t n, int C, int K, int normalize,
                 float threshold, int mode, const int64_t* __restrict__ idx, float* __restrict__ max_prob,
                 int64_t* __restrict__ label, float* __restrict__ probs_accum) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp0; p < n; p += nwarp) {
    float v[kLangVPL];
    float q = 0.f;
#pragma unroll
    for (int u = 0; u < kLangVPL; ++u) {
      const int c = lane + 32 * u;
      v[u] = c < C ? l_in<TF>(feat[(size_t)p * C + c]) : 0.f;
      q += v[u] * v[u];
    }
    if (normalize) {
      const float inv = 1.f / fmaxf(sqrtf(l_warp_sum(q)), 1e-12f);
#pragma unroll
      for (int u = 0; u < kLangVPL; ++u) v[u] *= inv;
    }
    float best = -1.f;
    int arg = 0;
    for (int k = 0; k < K; ++k) {
      const float* t = text + (size_t)k * C;
      float d = 0.f;
#pragma unroll
      for (int u = 0; u < kLangVPL; ++u) {
        const int c = lane + 32 * u;
        if (c < C) d = fmaf(v[u], __ldg(t + c), d);
      }
      d = l_warp_sum(d);
      const float pr = 1.f / (1.f + __expf(-d));
      if (mode == 1) {
        if (lane == 0) probs_accum[(size_t)(idx ? idx[p] : p) * K + k] += pr;
      } else if (pr > best) {
        best = pr;
        arg = k;
      }
    }
    if (mode == 0 && lane == 0) {
      max_prob[p] = best;
      label[p] = best < threshold ? -1 : arg;
    }
  }
}

// acc[0] += sum_valid (1 - cos(pred, target)); acc[1] += sum_valid ||pred - target||^2; acc[2] += n_valid
template <typename TP, typename TT>
__global__ void __launch_bounds__(256)
cos_l2_kernel(const TP* __restrict__ pred, const TT* __restrict__ target, const uint8_t* __restrict__ mask, int64_t n,
              int C, double* __restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float s_cos = 0.f, s_l2 = 0.f, s_n = 0.f;
  for (int64_t p = warp0; p < n; p += nwarp) {
    if (!mask[p]) continue;
    float pp = 0.f, tt = 0.f, pt = 0.f, dd = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float a = l_in<TP>(pred[(size_t)p * C + c]);
      const float b = l_in<TT>(target[(size_t)p * C + c]);
      pp = fmaf(a, a, pp);
      tt = fmaf(b, b, tt);
      pt = fmaf(a, b, pt);
      const float d = a - b;
      dd = fmaf(d, d, dd);
    }
    pp = l_warp_sum(pp);
    tt = l_warp_sum(tt);
    pt = l_warp_sum(pt);
    dd = l_warp_sum(dd);
    // torch.cosine_similarity: x.y / max(||x|| * ||y||, eps), eps = 1e-8
    const float cosv = pt / fmaxf(sqrtf(pp) * sqrtf(tt), 1e-8f);
    s_cos += 1.f - cosv;
    s_l2 += dd;
    s_n += 1.f;
  }
  if (lane == 0 && s_n > 0.f) {
    atomicAdd(&acc[0], (double)s_cos);
    atomicAdd(&acc[1], (double)s_l2);
    atomicAdd(&acc[2], (double)s_n);
  }
}

// sums[(label * 2 + half), :] += pred[p, :] for valid points (mask & label >= 0 & half in {0,1});
// counts[label * 2 + half] += 1.
template <typename TP>
__global__ void __launch_bounds__(256)
class_half_sums_kernel(const TP* __restrict__ pred, const uint8_t* __restrict__ mask, const int64_t* __restrict__ segment,
                       const int64_t* __restrict__ half, int64_t n, int C, int n_classes, float* __restrict__ sums,
                       int* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp0; p < n; p += nwarp) {
    if (!mask[p]) continue;
    const int64_t lab = segment[p];
    const int64_t hf = half[p];
    if (lab < 0 || lab >= n_classes || hf < 0 || hf > 1) continue;
    float* dst = sums + (size_t)(lab * 2 + hf) * C;
    for (int c = lane; c < C; c += 32) atomicAdd(dst + c, l_in<TP>(pred[(size_t)p * C + c]));
    if (lane == 0) atomicAdd(&counts[lab * 2 + hf], 1);
  }
}

}  // namespace ss

extern "C" {

int ss_lang_head(const void* feat, int feat_is_bf16, const float* text, int64_t n, int channels, int n_classes,
                 int normalize, float threshold, int mode, const int64_t* idx, float* max_prob, int64_t* label,
                 float* probs_accum, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || channels < 1 || channels > 32 * ss::kLangVPL || n_classes < 1 || (mode != 0 && mode != 1)) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!feat || !text) return SS_BAD_ARGS;
  if (mode == 0 && (!max_prob || !label)) return SS_BAD_ARGS;
  if (mode == 1 && !probs_accum) return SS_BAD_ARGS;
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, 8), 32 * ss::kNumSMs);
  if (feat_is_bf16)
    ss::lang_head_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)feat, text, n, channels,
                                                                    n_classes, normalize, threshold, mode, idx, max_prob,
                                                                    label, probs_accum);
  else
    ss::lang_head_kernel<float><<<blocks, 256, 0, stream>>>((const float*)feat, text, n, channels, n_classes, normalize,
                                                            threshold, mode, idx, max_prob, label, probs_accum);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

int ss_cos_l2_loss(const void* pred, int pred_is_bf16, const void* target, int target_dtype, const uint8_t* mask,
                   int64_t n, int channels, double* acc3, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || channels < 1 || !acc3 || target_dtype < 0 || target_dtype > 2) return SS_BAD_ARGS;
  SS_CUDA(cudaMemsetAsync(acc3, 0, 24, stream));
  if (n == 0) return SS_OK;
  if (!pred || !target || !mask) return SS_BAD_ARGS;
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, 8), 16 * ss::kNumSMs);
#define SS_CL_(TP, TT) \
  ss::cos_l2_kernel<TP, TT><<<blocks, 256, 0, stream>>>((const TP*)pred, (const TT*)target, mask, n, channels, acc3)
  // target_dtype: 0 = fp32, 1 = bf16, 2 = fp16 (lang_feat is stored as fp16 on disk)
  if (pred_is_bf16) {
    if (target_dtype == 0) SS_CL_(__nv_bfloat16, float);
    else if (target_dtype == 1) SS_CL_(__nv_bfloat16, __nv_bfloat16);
    else SS_CL_(__nv_bfloat16, __half);
  } else {
    if (target_dtype == 0) SS_CL_(float, float);
    else if (target_dtype == 1) SS_CL_(float, __nv_bfloat16);
    else SS_CL_(float, __half);
  }
#undef SS_CL_
  SS_CHECK_LAUNCH();
  return SS_OK;
}

int ss_class_half_sums(const void* pred, int pred_is_bf16, const uint8_t* mask, const int64_t* segment,
                       const int64_t* half, int64_t n, int channels, int n_classes, float* sums, int32_t* counts,
                       void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || channels < 1 || n_classes < 1 || !sums || !counts) return SS_BAD_ARGS;
  SS_CUDA(cudaMemsetAsync(sums, 0, (size_t)n_classes * 2 * channels * 4, stream));
  SS_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_classes * 2 * 4, stream));
  if (n == 0) return SS_OK;
  if (!pred || !mask || !segment || !half) return SS_BAD_ARGS;
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, 8), 16 * ss::kNumSMs);
  if (pred_is_bf16)
    ss::class_half_sums_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)pred, mask, segment, half,
                                                                          n, channels, n_classes, sums, counts);
  else
    ss::class_half_sums_kernel<float><<<blocks, 256, 0, stream>>>((const float*)pred, mask, segment, half, n, channels,
                                                                  n_classes, sums, counts);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // extern "C"
