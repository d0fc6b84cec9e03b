// SphereCrop on the GPU: rank all points by squared distance to a centre and keep the nearest `point_max`.
//
// Replaces (reference): pointcept/datasets/transform.py:1419-1535 (SphereCrop, modes "random" / "center"):
//   idx_crop = np.argsort(np.sum(np.square(coord - center), 1))[:point_max]
// The distance is evaluated with numpy's fp32 arithmetic ((dx*dx + dy*dy) + dz*dz, every operation rounded, no
// FMA), its bit pattern (non-negative float -> monotone uint32) is the 32-bit radix-sort key, and the stable LSD
// sort breaks ties by ascending index (numpy's own tie order is unspecified: introsort).
#include "radix_sort.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

constexpr int kCropBits = 8;  // 32-bit keys: 4 passes of 8 bits (make_radix_plan)
constexpr int kCropThreads = 256;
__global__ void __launch_bounds__(kCropThreads)
sphere_key_kernel(const float* __restrict__ coord, int64_t n, float cx, float cy, float cz, uint64_t* __restrict__ key,
                  uint32_t* __restrict__ ghist, void* status, size_t status_bytes) {
  constexpr int BINS = 1 << kCropBits;
  __shared__ uint32_t s_hist[BINS];  // first-digit histogram (shared-memory atomics)
  radix_zero_status(status, status_bytes);
  for (int i = threadIdx.x; i < BINS; i += blockDim.x) s_hist[i] = 0u;
  __syncthreads();
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + threadIdx.x;
    const bool ok = i < n;
    uint64_t k = 0;
    if (ok) {
      const float dx = __fsub_rn(coord[3 * i], cx), dy = __fsub_rn(coord[3 * i + 1], cy), dz = __fsub_rn(coord[3 * i + 2], cz);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      k = (uint64_t)__float_as_uint(d2);
      key[i] = k;
    }
    if (ok) atomicAdd(&s_hist[(uint32_t)k & (BINS - 1)], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < BINS; d += blockDim.x) {
    const uint32_t c = s_hist[d];
    if (c) atomicAdd(&ghist[d], c);
  }
}

struct CropPlan {
  RadixPlan radix;
  size_t off_key, total;
};

inline CropPlan make_crop_plan(int64_t n) {
  CropPlan p;
  p.radix = make_radix_plan(1, (int)n, 32);
  size_t o = align_up(p.radix.total, 256);
  p.off_key = o;
  o += align_up((size_t)(n > 0 ? n : 1) * 8, 256);
  p.total = o;
  return p;
}

}  // namespace ss

extern "C" {

size_t ss_sphere_crop_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  return ss::make_crop_plan(n).total + 256;
}

int ss_sphere_crop_order(const float* coord, int64_t n, const float* center3, int64_t* order, uint64_t* dist_bits_sorted,
                         void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || n > 0x3fffffff || !center3) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!coord || !order || !dist_bits_sorted || !workspace) return SS_BAD_ARGS;
  ss::CropPlan p = ss::make_crop_plan(n);
  if (workspace_bytes < p.total) return SS_BAD_ARGS;
  char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  uint64_t* key = (uint64_t*)(ws + p.off_key);
  if (p.radix.bits != ss::kCropBits) return SS_BAD_ARGS;
  SS_CUDA(cudaMemsetAsync(ws + p.radix.off_hist, 0, p.radix.small_zero_bytes, stream));
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, ss::kCropThreads), 2 * ss::kNumSMs);
  ss::sphere_key_kernel<<<blocks, ss::kCropThreads, 0, stream>>>(coord, n, center3[0], center3[1], center3[2], key,
                                                                 (uint32_t*)(ws + p.radix.off_hist),
                                                                 ws + p.radix.off_status, p.radix.status_bytes);
  SS_CHECK_LAUNCH();
  return ss::radix_sort_run(p.radix, ws, key, ss::kFinalPairs, order, nullptr, dist_bits_sorted, stream);
}

}  // extern "C"
