// Point.serialization on the GPU: fused space-filling-curve encoding of up to 4 orders + first-digit histogram, then a
// stable radix sort of all rows at once (3 passes of 9 bits on 32-bit keys for a depth-9 chunk), the last pass writing
// `order` and the inverse permutation directly.  One small memset + 1 + passes launches.
//
// Replaces (reference): pointcept/models/utils/structure.py:47-102 (Point.serialization),
// pointcept/models/utils/serialization/default.py:8-24 (encode), z_order.py:66-101, hilbert.py:91-198,
// pointcept/models/utils/misc.py:19-24 (offset2batch).
#include "radix_sort.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

constexpr int kEncThreads = 1024;

// One wave of fat CTAs (<= one per SM, 32 warps), every thread encodes points blockIdx.x * 1024 + tid, + gridDim.x * 1024,
// ...: 24-byte (int64) / 12-byte (int32) coordinate rows read directly (a warp covers 768 / 384 contiguous bytes), the
// codes of all rows written coalesced, and ONLY the first radix digit histogrammed (later digits are counted by the sort
// passes themselves) with shared-memory atomics into one CTA-wide table, flushed with one global atomic per (row, digit)
// and CTA.  The Hilbert codes come from the z-order keys through the finite-state machine of common.cuh; depth <= 10
// uses 32-bit arithmetic.  Also clears the sort's look-back status words.
template <typename CoordT, int BITS>
__global__ void __launch_bounds__(kEncThreads)
encode_hist_kernel(const CoordT* __restrict__ grid_coord, const int64_t* __restrict__ offset, int n_batch,
                   int64_t* __restrict__ batch_out, int64_t* __restrict__ code, uint32_t* __restrict__ ghist,
                   size_t hist_row_stride, int n, int depth, int rows, int4 order_ids, void* status, size_t status_bytes) {
  constexpr int BINS = 1 << BITS;
  __shared__ uint32_t s_hist[4 * BINS];  // [rows][BINS]
  __shared__ uint16_t s_fsm[kHilbertStates * 8];
  const int tid = threadIdx.x;
  hilbert_fsm_load(s_fsm);
  radix_zero_status(status, status_bytes);
  for (int i = tid; i < rows * BINS; i += kEncThreads) s_hist[i] = 0u;
  __syncthreads();
  const int ids[4] = {order_ids.x, order_ids.y, order_ids.z, order_ids.w};
  const bool small = depth <= 10;
  for (int64_t b0 = (int64_t)blockIdx.x * kEncThreads; b0 < n; b0 += (int64_t)gridDim.x * kEncThreads) {
    const int64_t p = b0 + tid;
    const bool ok = p < n;
    if (ok) {
      const uint32_t x = (uint32_t)grid_coord[p * 3 + 0];
      const uint32_t y = (uint32_t)grid_coord[p * 3 + 1];
      const uint32_t z = (uint32_t)grid_coord[p * 3 + 2];
      // batch id = number of offsets <= p  (offset is the cumulative count, misc.py:19-24)
      int lo = 0, hi = n_batch;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (offset[mid] <= p) lo = mid + 1; else hi = mid;
      }
      const uint64_t bpart = (uint64_t)lo << (3 * depth);
      if (batch_out) batch_out[p] = (int64_t)lo;
      // z keys of (x, y, z) and (y, x, z) from three bit spreads; the Hilbert keys from the z keys (common.cuh)
      uint64_t zk, ztk;
      if (small) {
        const uint32_t m = (1u << depth) - 1u;
        const uint32_t sx = spread3_10(x & m), sy = spread3_10(y & m), sz = spread3_10(z & m);
        zk = (sx << 2) | (sy << 1) | sz;
        ztk = (sy << 2) | (sx << 1) | sz;
      } else {
        const uint32_t m = (1u << depth) - 1u;
        const uint64_t sx = spread3(x & m), sy = spread3(y & m), sz = spread3(z & m);
        zk = (sx << 2) | (sy << 1) | sz;
        ztk = (sy << 2) | (sx << 1) | sz;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (r < rows) {
          const uint64_t zz = (ids[r] & 1) ? ztk : zk;
          uint64_t k = zz;
          if (ids[r] >= 2)
            k = small ? (uint64_t)hilbert_from_zkey<uint32_t>((uint32_t)zz, depth, s_fsm)
                      : hilbert_from_zkey<uint64_t>(zz, depth, s_fsm);
          const uint64_t c = bpart | k;
          code[(size_t)r * n + p] = (int64_t)c;
          atomicAdd(&s_hist[r * BINS + ((uint32_t)c & (BINS - 1))], 1u);
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < rows * BINS; i += kEncThreads) {
    const uint32_t s = s_hist[i];
    if (s) atomicAdd(&ghist[(size_t)(i >> BITS) * hist_row_stride + (i & (BINS - 1))], s);
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
  SS_CUDA(cudaMemsetAsync(ws + p.off_hist, 0, p.small_zero_bytes, stream));
  int4 ids = make_int4(order_ids[0], rows > 1 ? order_ids[1] : 0, rows > 2 ? order_ids[2] : 0, rows > 3 ? order_ids[3] : 0);
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, ss::kEncThreads), ss::kNumSMs);  // <= one CTA per SM
  uint32_t* ghist = (uint32_t*)(ws + p.off_hist);
  void* status = ws + p.off_status;
#define SS_ENC_(T, B)                                                                                               \
  ss::encode_hist_kernel<T, B><<<blocks, ss::kEncThreads, 0, stream>>>((const T*)grid_coord, offset, n_batch, batch_out, \
                                                                       code, ghist, p.hist_row_stride(), (int)n, depth,  \
                                                                       rows, ids, status, p.status_bytes)
#define SS_ENC_B_(T)                  \
  do {                                \
    if (p.bits == 8) SS_ENC_(T, 8);   \
    else if (p.bits == 9) SS_ENC_(T, 9); \
    else SS_ENC_(T, 10);              \
  } while (0)
  if (coord_is_int32) SS_ENC_B_(int);
  else SS_ENC_B_(long long);
#undef SS_ENC_B_
#undef SS_ENC_
  SS_CHECK_LAUNCH();
  return ss::radix_sort_run(p, ws, (const uint64_t*)code, ss::kFinal, order, inverse, nullptr, stream);
}

}  // extern "C"
