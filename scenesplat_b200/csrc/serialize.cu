// Point.serialization on the GPU: fused space-filling-curve encoding of up to 4 orders + digit
// histograms, then a stable 64-bit-key radix sort of all rows at once, the last pass writing
// `order` and the inverse permutation directly.
//
// Replaces (reference): pointcept/models/utils/structure.py:47-102 (Point.serialization),
// pointcept/models/utils/serialization/default.py:8-24 (encode), z_order.py:66-101, hilbert.py:91-198,
// pointcept/models/utils/misc.py:19-24 (offset2batch).
#include "radix_sort.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

constexpr int kEncThreads = 256;

// Per-CTA: 256 points.  Coordinates are staged through shared memory with fully coalesced 8-byte
// (int64) or 4-byte (int32) loads; every thread then encodes its point for all rows.
template <typename CoordT>
__global__ void __launch_bounds__(kEncThreads)
encode_hist_kernel(const CoordT* __restrict__ grid_coord, const int64_t* __restrict__ offset, int n_batch,
                   int64_t* __restrict__ batch_out, int64_t* __restrict__ code, uint32_t* __restrict__ ghist, int n,
                   int depth, int rows, int4 order_ids, int passes) {
  extern __shared__ uint32_t s_dyn[];  // [rows][passes][256] histogram
  __shared__ CoordT s_coord[kEncThreads * 3];
  const int tid = threadIdx.x;
  const int nh = rows * passes * kRadix;
  for (int i = tid; i < nh; i += kEncThreads) s_dyn[i] = 0u;
  const int64_t first = (int64_t)blockIdx.x * kEncThreads;
  const int cnt = (int)min((int64_t)kEncThreads, (int64_t)n - first);
  for (int i = tid; i < cnt * 3; i += kEncThreads) s_coord[i] = grid_coord[first * 3 + i];
  __syncthreads();
  if (tid < cnt) {
    const int64_t p = first + tid;
    const uint32_t x = (uint32_t)s_coord[tid * 3 + 0];
    const uint32_t y = (uint32_t)s_coord[tid * 3 + 1];
    const uint32_t z = (uint32_t)s_coord[tid * 3 + 2];
    // batch id = number of offsets <= p  (offset is the cumulative count, misc.py:19-24)
    int lo = 0, hi = n_batch;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (offset[mid] <= p) lo = mid + 1; else hi = mid;
    }
    const uint64_t b = (uint64_t)lo;
    if (batch_out) batch_out[p] = (int64_t)b;
    const int ids[4] = {order_ids.x, order_ids.y, order_ids.z, order_ids.w};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r < rows) {
        const uint64_t c = (b << (3 * depth)) | sfc_key(ids[r], x, y, z, depth);
        code[(size_t)r * n + p] = (int64_t)c;
        for (int ps = 0; ps < passes; ++ps)
          atomicAdd(&s_dyn[(r * passes + ps) * kRadix + (int)((c >> (ps * kRadixBits)) & (kRadix - 1))], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < nh; i += kEncThreads) {
    const uint32_t c = s_dyn[i];
    if (c) atomicAdd(&ghist[i], c);
  }
}

// max over all grid coordinates (Point.serialization's depth = bit_length(max), structure.py:66)
template <typename CoordT>
__global__ void __launch_bounds__(256) coord_max_kernel(const CoordT* __restrict__ g, int64_t n3, long long* out) {
  long long m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (int64_t)gridDim.x * blockDim.x)
    m = max(m, (long long)g[i]);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

static inline int key_bits_for(int depth, int n_batch) {
  int bb = 0;
  while ((1ll << bb) < (long long)n_batch) ++bb;  // bits needed for batch ids 0..n_batch-1
  return 3 * depth + bb;
}

}  // namespace ss

extern "C" {

int ss_coord_max(const void* grid_coord, int coord_is_int32, int64_t n, int64_t* out_max_dev, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!grid_coord || !out_max_dev || n < 0) return SS_BAD_ARGS;
  SS_CUDA(cudaMemsetAsync(out_max_dev, 0, 8, stream));
  if (n == 0) return SS_OK;
  const int64_t n3 = n * 3;
  const int blocks = (int)ss::imin64(ss::ceil_div64(n3, 256 * 8), 4 * ss::kNumSMs);
  if (coord_is_int32)
    ss::coord_max_kernel<int><<<blocks, 256, 0, stream>>>((const int*)grid_coord, n3, (long long*)out_max_dev);
  else
    ss::coord_max_kernel<long long><<<blocks, 256, 0, stream>>>((const long long*)grid_coord, n3, (long long*)out_max_dev);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

size_t ss_serialize_workspace_bytes(int64_t n, int rows, int depth, int n_batch) {
  if (n < 0 || rows < 1 || rows > 4) return 0;
  ss::RadixPlan p = ss::make_radix_plan(rows, (int)n, ss::key_bits_for(depth, n_batch));
  return p.total + 256;
}

int ss_serialize(const void* grid_coord, int coord_is_int32, const int64_t* offset, int n_batch, int64_t n, int depth,
                 int rows, const int* order_ids, int64_t* batch_out, int64_t* code, int64_t* order, int64_t* inverse,
                 void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || n > 0x3fffffff || rows < 1 || rows > 4 || depth < 1 || depth > 16 || n_batch < 1 || !order_ids)
    return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!grid_coord || !offset || !code || !order || !inverse || !workspace) return SS_BAD_ARGS;
  const int kb = ss::key_bits_for(depth, n_batch);
  {  // structure.py:69: depth * 3 + len(offset).bit_length() <= 63
    int bl = 0;
    while ((n_batch >> bl) != 0) ++bl;
    if (3 * depth + bl > 63) return SS_BAD_ARGS;
  }
  for (int r = 0; r < rows; ++r)
    if (order_ids[r] < 0 || order_ids[r] > 3) return SS_BAD_ARGS;
  ss::RadixPlan p = ss::make_radix_plan(rows, (int)n, kb);
  if (workspace_bytes < p.total) return SS_BAD_ARGS;
  char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  SS_CUDA(cudaMemsetAsync(ws + p.off_hist, 0, p.zero_bytes, stream));
  int4 ids = make_int4(order_ids[0], rows > 1 ? order_ids[1] : 0, rows > 2 ? order_ids[2] : 0, rows > 3 ? order_ids[3] : 0);
  const int blocks = ss::ceil_div((int)n, ss::kEncThreads);
  const size_t smem = (size_t)rows * p.passes * ss::kRadix * 4;
  uint32_t* ghist = (uint32_t*)(ws + p.off_hist);
  if (coord_is_int32)
    ss::encode_hist_kernel<int><<<blocks, ss::kEncThreads, smem, stream>>>((const int*)grid_coord, offset, n_batch,
                                                                           batch_out, code, ghist, (int)n, depth, rows,
                                                                           ids, p.passes);
  else
    ss::encode_hist_kernel<long long><<<blocks, ss::kEncThreads, smem, stream>>>(
        (const long long*)grid_coord, (const int64_t*)offset, n_batch, batch_out, code, ghist, (int)n, depth, rows, ids,
        p.passes);
  SS_CHECK_LAUNCH();
  return ss::radix_sort_run(p, ws, (const uint64_t*)code, ss::kFinal, order, inverse, nullptr, stream);
}

}  // extern "C"
