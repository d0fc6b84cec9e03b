// GridSample voxel hash / unique / select on the GPU.
//
// Replaces (reference): pointcept/datasets/transform.py:1211-1300 (GridSample.__call__ train mode),
// :1302-1330 (test-mode fragments), :1384-1416 (ravel_hash_vec / fnv_hash_vec).
//
// Bit-exact outputs: voxel order (ascending uint64 hash), `inverse`, `count`, `grid_coord`.
// The representative of a voxel is made explicit: members ordered by original index (stable
// sort), member `rand[v] % count[v]` (train) or `frag % count[v]` (test) is taken.
#include "radix_sort.cuh"
#include "runs.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

__device__ __forceinline__ long long voxel_of(float c, double grid_size) {
  // float64 divide, as numpy >= 2 does for float32_array / float64_0d (transform.py:1213)
  return (long long)floor((double)c / grid_size);
}

__global__ void __launch_bounds__(256) gs_minmax_kernel(const float* __restrict__ coord, int64_t n, double grid_size,
                                                        long long* __restrict__ mn, long long* __restrict__ mx) {
  long long lo[3] = {LLONG_MAX, LLONG_MAX, LLONG_MAX}, hi[3] = {LLONG_MIN, LLONG_MIN, LLONG_MIN};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const long long g = voxel_of(coord[i * 3 + a], grid_size);
      lo[a] = min(lo[a], g);
      hi[a] = max(hi[a], g);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&mn[a], lo[a]);
      atomicMax(&mx[a], hi[a]);
    }
  }
}

__global__ void gs_init_minmax(long long* mn, long long* mx) {
  if (threadIdx.x < 3) {
    mn[threadIdx.x] = LLONG_MAX;
    mx[threadIdx.x] = LLONG_MIN;
  }
}

// key = hash(voxel - min); also the histogram of the first radix digit (warp-private counters filled by ballot matching;
// the sort passes count the later digits themselves) and the clear of the sort's look-back status words.
constexpr int kGsBits = 10;  // 64-bit keys: 7 passes of 10 bits (make_radix_plan)
constexpr int kGsThreads = 256;
__global__ void __launch_bounds__(kGsThreads)
gs_hash_kernel(const float* __restrict__ coord, int64_t n, double grid_size, const long long* __restrict__ mn,
               const long long* __restrict__ mx, int hash_type, uint64_t* __restrict__ key, uint32_t* __restrict__ ghist,
               void* status, size_t status_bytes) {
  constexpr int BINS = 1 << kGsBits;
  __shared__ uint32_t s_hist[BINS];  // first-digit histogram (shared-memory atomics)
  radix_zero_status(status, status_bytes);
  for (int i = threadIdx.x; i < BINS; i += blockDim.x) s_hist[i] = 0u;
  __syncthreads();
  const long long m0 = mn[0], m1 = mn[1], m2 = mn[2];
  const unsigned long long e1 = (unsigned long long)(mx[1] - m1) + 1ull, e2 = (unsigned long long)(mx[2] - m2) + 1ull;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + threadIdx.x;
    const bool ok = i < n;
    unsigned long long h = 0;
    if (ok) {
      const unsigned long long g0 = (unsigned long long)(voxel_of(coord[i * 3 + 0], grid_size) - m0);
      const unsigned long long g1 = (unsigned long long)(voxel_of(coord[i * 3 + 1], grid_size) - m1);
      const unsigned long long g2 = (unsigned long long)(voxel_of(coord[i * 3 + 2], grid_size) - m2);
      if (hash_type == 0) {  // "fnv" as written in the reference: multiply, then xor the whole coordinate
        h = 14695981039346656037ull;
        h *= 1099511628211ull; h ^= g0;
        h *= 1099511628211ull; h ^= g1;
        h *= 1099511628211ull; h ^= g2;
      } else {  // ravel
        h = (g0 * e1 + g1) * e2 + g2;
      }
      key[i] = h;
    }
    if (ok) atomicAdd(&s_hist[(uint32_t)h & (BINS - 1)], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < BINS; d += blockDim.x) {
    const uint32_t c = s_hist[d];
    if (c) atomicAdd(&ghist[d], c);
  }
}

struct GsRuns {
  const uint64_t* key_sorted;
  const int64_t* idx_sort;
  int64_t* inverse;  // [n] per raw point -> voxel rank
  int64_t* start;    // [n+1] first sorted position of every voxel; start[M] = n written by fix-up
  int64_t n;
  __device__ bool head(int, int64_t j) const { return j == 0 || key_sorted[j] != key_sorted[j - 1]; }
  __device__ void emit(int, int64_t j, uint32_t run, bool is_head) const {
    inverse[idx_sort[j]] = (int64_t)run;
    if (is_head) start[run] = j;
  }
  __device__ void finish(int, uint32_t total) const { start[total] = n; }
};

__global__ void __launch_bounds__(256)
gs_select_kernel(const float* __restrict__ coord, double grid_size, const long long* __restrict__ mn,
                 const int64_t* __restrict__ idx_sort, const int64_t* __restrict__ start, const int64_t* __restrict__ m_dev,
                 const int64_t* __restrict__ rnd, int64_t frag, int64_t* __restrict__ idx_unique,
                 int64_t* __restrict__ grid_coord_out, int64_t* __restrict__ count_out) {
  const int64_t m = *m_dev;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < m; v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = start[v], c = start[v + 1] - s;
    const int64_t r = rnd ? rnd[v] : frag;
    const int64_t src = idx_sort[s + r % c];
    idx_unique[v] = src;
    if (count_out) count_out[v] = c;
    if (grid_coord_out) {
#pragma unroll
      for (int a = 0; a < 3; ++a) grid_coord_out[v * 3 + a] = voxel_of(coord[src * 3 + a], grid_size) - mn[a];
    }
  }
}

// dst[v, :] = src[idx[v], :] for rows of `row_bytes` (multiple of 4); 16-byte vectors when aligned.
template <typename VecT>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const VecT* __restrict__ src, VecT* __restrict__ dst, const int64_t* __restrict__ idx,
                   const int64_t* __restrict__ count_dev, int64_t count, int vec_per_row) {
  const int64_t m = count_dev ? *count_dev : count;
  const int64_t total = m * vec_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / vec_per_row;
    const int c = (int)(i - v * vec_per_row);
    dst[i] = src[idx[v] * vec_per_row + c];
  }
}

struct GsPlan {
  RadixPlan radix;
  size_t off_minmax, off_key, off_keysorted, off_runs, total;
};

inline GsPlan make_gs_plan(int64_t n) {
  GsPlan p;
  p.radix = make_radix_plan(1, (int)n, 64);
  size_t o = align_up(p.radix.total, 256);
  p.off_minmax = o; o += 256;
  p.off_key = o; o += align_up((size_t)n * 8, 256);
  p.off_keysorted = o; o += align_up((size_t)n * 8, 256);
  p.off_runs = o; o += align_up(runs_workspace_bytes(n), 256);
  p.total = o;
  return p;
}

}  // namespace ss

extern "C" {

size_t ss_gridsample_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  return ss::make_gs_plan(n).total + 256;
}

int ss_gridsample_index(const float* coord, int64_t n, double grid_size, int hash_type, int64_t* idx_sort,
                        int64_t* inverse, int64_t* start, int64_t* m_dev, int64_t* min_coord_dev, void* workspace,
                        size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || n > 0x3fffffff || !(grid_size > 0) || (hash_type != 0 && hash_type != 1)) return SS_BAD_ARGS;
  if (!m_dev) return SS_BAD_ARGS;
  if (n == 0) {
    SS_CUDA(cudaMemsetAsync(m_dev, 0, 8, stream));
    return SS_OK;
  }
  if (!coord || !idx_sort || !inverse || !start || !min_coord_dev || !workspace) return SS_BAD_ARGS;
  ss::GsPlan p = ss::make_gs_plan(n);
  if (workspace_bytes < p.total) return SS_BAD_ARGS;
  char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  long long* mn = (long long*)(ws + p.off_minmax);
  long long* mx = mn + 4;
  uint64_t* key = (uint64_t*)(ws + p.off_key);
  uint64_t* key_sorted = (uint64_t*)(ws + p.off_keysorted);
  if (p.radix.bits != ss::kGsBits) return SS_BAD_ARGS;
  SS_CUDA(cudaMemsetAsync(ws + p.radix.off_hist, 0, p.radix.small_zero_bytes, stream));
  ss::gs_init_minmax<<<1, 32, 0, stream>>>(mn, mx);
  SS_CHECK_LAUNCH();
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, 256), 8 * ss::kNumSMs);
  ss::gs_minmax_kernel<<<blocks, 256, 0, stream>>>(coord, n, grid_size, mn, mx);
  SS_CHECK_LAUNCH();
  ss::gs_hash_kernel<<<(int)ss::imin64(blocks, 2 * ss::kNumSMs), ss::kGsThreads, 0, stream>>>(
      coord, n, grid_size, mn, mx, hash_type, key, (uint32_t*)(ws + p.radix.off_hist), ws + p.radix.off_status,
      p.radix.status_bytes);
  SS_CHECK_LAUNCH();
  int rc = ss::radix_sort_run(p.radix, ws, key, ss::kFinalPairs, idx_sort, nullptr, key_sorted, stream);
  if (rc) return rc;
  ss::GsRuns f{key_sorted, idx_sort, inverse, start, n};
  rc = ss::runs_launch(f, n, ws + p.off_runs, m_dev, stream);
  if (rc) return rc;
  SS_CUDA(cudaMemcpyAsync(min_coord_dev, mn, 24, cudaMemcpyDeviceToDevice, stream));
  return SS_OK;
}

int ss_gridsample_select(const float* coord, int64_t n, double grid_size, const int64_t* min_coord_dev,
                         const int64_t* idx_sort, const int64_t* start, const int64_t* m_dev, const int64_t* rnd,
                         int64_t frag, int64_t* idx_unique, int64_t* grid_coord_out, int64_t* count_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || frag < 0) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!coord || !min_coord_dev || !idx_sort || !start || !m_dev || !idx_unique) return SS_BAD_ARGS;
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, 256), 8 * ss::kNumSMs);
  ss::gs_select_kernel<<<blocks, 256, 0, stream>>>(coord, grid_size, (const long long*)min_coord_dev, idx_sort, start,
                                                   m_dev, rnd, frag, idx_unique, grid_coord_out, count_out);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

int ss_gather_rows(const void* src, int64_t row_bytes, const int64_t* idx, const int64_t* count_dev, int64_t count,
                   void* dst, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (row_bytes <= 0 || count < 0) return SS_BAD_ARGS;
  if (count == 0) return SS_OK;
  if (!src || !idx || !dst) return SS_BAD_ARGS;
  const bool v16 = (row_bytes % 16 == 0) && (((uintptr_t)src | (uintptr_t)dst) % 16 == 0);
  const bool v4 = (row_bytes % 4 == 0) && (((uintptr_t)src | (uintptr_t)dst) % 4 == 0);
  const int64_t vec = v16 ? row_bytes / 16 : (v4 ? row_bytes / 4 : row_bytes);
  const int blocks = (int)ss::imin64(ss::ceil_div64(count * vec, 256), 16 * ss::kNumSMs);
  if (v16)
    ss::gather_rows_kernel<uint4><<<blocks, 256, 0, stream>>>((const uint4*)src, (uint4*)dst, idx, count_dev, count, (int)vec);
  else if (v4)
    ss::gather_rows_kernel<uint32_t><<<blocks, 256, 0, stream>>>((const uint32_t*)src, (uint32_t*)dst, idx, count_dev, count, (int)vec);
  else
    ss::gather_rows_kernel<uint8_t><<<blocks, 256, 0, stream>>>((const uint8_t*)src, (uint8_t*)dst, idx, count_dev, count, (int)vec);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // extern "C"
