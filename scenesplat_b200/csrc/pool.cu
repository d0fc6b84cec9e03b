// SerializedPooling / SerializedUnpooling on the GPU.
//
// Replaces (reference): pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py:384-428
// (torch.unique + torch.sort + argsort x4 + scatter_ + torch_scatter.segment_csr) and :471-482.
//
// Key fact: the parent's `order[0]` already sorts by `code[0]`, and all members of a
// cluster share `code[r] >> 3` in every row r, so clusters are contiguous runs along EVERY parent
// order row.  Cluster ids, counts, the pooled codes AND the pooled order/inverse of all rows
// therefore come out of run-length scans -- no sort, no unique.
#include "runs.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

// One launch for a whole pooling level.  Sequence y = 0 walks the parent's order row 0: cluster ids, segment starts,
// heads, the pooled codes of every row, pooled grid_coord / batch.  Sequences y >= 1 walk the other parent rows: their
// run ids ARE the pooled order (rank of the cluster along that curve); they need row 0's cluster ids, so their emit
// phase waits for row 0 (runs.cuh, kAfterRow0) while their flags, scans and look-backs proceed.
struct PoolRuns {
  static constexpr bool kAfterRow0 = true;
  const int64_t* code;   // parent [k][n]
  const int64_t* order;  // parent [k][n]
  const int64_t* grid_coord;  // parent [n][3] (nullable)
  const int64_t* batch;       // parent [n] (nullable)
  int64_t n;
  int k;
  int shift;  // 3 * pooling_depth
  int pool_depth;
  int4 src_row;        // child row r' <- parent row src_row[r']
  int rows_parent[5];  // sequence y -> parent row (y = 0: parent row 0)
  int rows_child[5];   // sequence y -> child row
  int64_t m_cap;       // row stride of the child arrays
  int64_t* cluster;    // [n]
  int64_t* seg_start;  // [n+1]
  int64_t* head_idx;   // [m_cap]
  int64_t* ccode;      // child [k][m_cap]
  int64_t* corder;
  int64_t* cinverse;
  int64_t* cgrid;      // [m_cap][3]
  int64_t* cbatch;     // [m_cap]
  __device__ bool head(int y, int64_t j) const {
    if (j == 0) return true;
    const int64_t* c = code + (size_t)rows_parent[y] * n;
    const int64_t* o = order + (size_t)rows_parent[y] * n;
    return (c[o[j]] >> shift) != (c[o[j - 1]] >> shift);
  }
  __device__ void emit(int y, int64_t j, uint32_t run, bool is_head) const {
    if (y == 0) {
      const int64_t p = order[j];
      cluster[p] = (int64_t)run;
      if (is_head) {
        seg_start[run] = j;
        head_idx[run] = p;
        const int rows[4] = {src_row.x, src_row.y, src_row.z, src_row.w};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (r < k) {
            ccode[(size_t)r * m_cap + run] = code[(size_t)rows[r] * n + p] >> shift;
            if (rows[r] == 0) {
              corder[(size_t)r * m_cap + run] = (int64_t)run;
              cinverse[(size_t)r * m_cap + run] = (int64_t)run;
            }
          }
        }
        if (cgrid) {
#pragma unroll
          for (int a = 0; a < 3; ++a) cgrid[(size_t)run * 3 + a] = grid_coord[p * 3 + a] >> pool_depth;
        }
        if (cbatch) cbatch[run] = batch[p];
      }
    } else if (is_head) {
      const int64_t c = __ldcg(cluster + order[(size_t)rows_parent[y] * n + j]);  // written by other CTAs of this launch
      corder[(size_t)rows_child[y] * m_cap + run] = c;
      cinverse[(size_t)rows_child[y] * m_cap + c] = (int64_t)run;
    }
  }
  __device__ void finish(int y, uint32_t total) const {
    if (y == 0) seg_start[total] = n;
  }
};

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return gelu_fast(x); }

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// One warp per segment (grid-stride).  out[s, :] = reduce_{j in [start[s], start[s+1])} src[order[j], :]
// (order == nullptr: rows in place, the plain torch_scatter.segment_csr; empty segments give 0 like torch_scatter)
// then optional per-channel affine (eval-mode BatchNorm folded to scale/shift) and GELU.
// reduce: 0 = sum, 1 = mean, 2 = max, 3 = min
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
segment_reduce_kernel(const TI* __restrict__ src, const int64_t* __restrict__ order, const int64_t* __restrict__ start,
                      const int64_t* __restrict__ m_dev, int64_t m, int C, int reduce, const float* __restrict__ scale,
                      const float* __restrict__ shift, int act, TO* __restrict__ out) {
  const int64_t M = m_dev ? *m_dev : m;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = warp0; s < M; s += nwarp) {
    const int64_t a = start[s], b = start[s + 1];
    for (int c0 = 0; c0 < C; c0 += 128) {  // each lane owns up to 4 channels per sweep (c0 + lane + 32*u)
      float acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = reduce == 2 ? -INFINITY : (reduce == 3 ? INFINITY : 0.f);
      for (int64_t j = a; j < b; ++j) {
        const TI* row = src + (size_t)(order ? order[j] : j) * C;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + lane + 32 * u;
          if (c < C) {
            const float v = to_f<TI>(row[c]);
            acc[u] = reduce == 2 ? fmaxf(acc[u], v) : (reduce == 3 ? fminf(acc[u], v) : acc[u] + v);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + lane + 32 * u;
        if (c < C) {
          float v = b > a ? acc[u] : 0.f;
          if (reduce == 1 && b > a) v = v / (float)(b - a);
          if (scale) v = v * scale[c] + shift[c];
          if (act == 1) v = gelu_erf(v);
          out[(size_t)s * C + c] = from_f<TO>(v);
        }
      }
    }
  }
}

// out[i, :] = f(a[i, :]) + f(b[cluster[i], :]),  f = optional affine (folded BN) + GELU per branch.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
unpool_kernel(const TI* __restrict__ a, const TI* __restrict__ b, const int64_t* __restrict__ cluster, int64_t n, int C,
              const float* __restrict__ sa, const float* __restrict__ ta, const float* __restrict__ sb,
              const float* __restrict__ tb, int act, TO* __restrict__ out, TO* __restrict__ out_a) {
  const int64_t total = n * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    float va = to_f<TI>(a[i]);
    float vb = to_f<TI>(b[(size_t)cluster[r] * C + c]);
    if (sa) va = va * sa[c] + ta[c];
    if (sb) vb = vb * sb[c] + tb[c];
    if (act == 1) {
      va = gelu_erf(va);
      vb = gelu_erf(vb);
    }
    out[i] = from_f<TO>(va + vb);
    if (out_a) out_a[i] = from_f<TO>(va);
  }
}

// 8 channels per thread (16-byte bf16 / 2 x 16-byte fp32 accesses), one row index per 8 elements.
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

template <typename TI, typename TO, typename TA>
__global__ void __launch_bounds__(256)
unpool_vec8_kernel(const TI* __restrict__ a, const TI* __restrict__ b, const int64_t* __restrict__ cluster, int64_t n,
                   int C, const float* __restrict__ sa, const float* __restrict__ ta, const float* __restrict__ sb,
                   const float* __restrict__ tb, int act, TO* __restrict__ out, TA* __restrict__ out_a) {
  const int c8 = C >> 3;
  const int64_t total = n * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c8;
    const int c = (int)(i - r * c8) << 3;
    float va[8], vb[8], o[8];
    load8<TI>(a + (size_t)r * C + c, va);
    load8<TI>(b + (size_t)cluster[r] * C + c, vb);
    if (sa) {
      float s[8], t[8];
      load8<float>(sa + c, s);
      load8<float>(ta + c, t);
#pragma unroll
      for (int u = 0; u < 8; ++u) va[u] = va[u] * s[u] + t[u];
    }
    if (sb) {
      float s[8], t[8];
      load8<float>(sb + c, s);
      load8<float>(tb + c, t);
#pragma unroll
      for (int u = 0; u < 8; ++u) vb[u] = vb[u] * s[u] + t[u];
    }
    if (act == 1) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        va[u] = gelu_erf(va[u]);
        vb[u] = gelu_erf(vb[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) o[u] = va[u] + vb[u];
    store8<TO>(out + (size_t)r * C + c, o);
    if (out_a) store8<TA>(out_a + (size_t)r * C + c, va);
  }
}

// Vectorised segment reduction (rows of 8 x k channels, 16-byte aligned): 2^lps_log2 lanes share one segment, every lane
// owns 8 consecutive channels per sweep (one 16-byte bf16 / two 16-byte fp32 loads per member row), so a warp reduces
// 32 >> lps_log2 segments at once and keeps two member rows in flight.  Same summation order as the scalar kernel
// (members in ascending j), so the results are bit-identical.  Optionally the same launch reduces a second, 3-channel
// fp32 source over the same segments with "mean" (the pooled coordinates of SerializedPooling, ptv3:400-404).
template <int REDUCE> __device__ __forceinline__ float seg_op(float a, float v) {
  return REDUCE == 2 ? fmaxf(a, v) : (REDUCE == 3 ? fminf(a, v) : a + v);
}

template <typename TI, typename TO, int REDUCE, int LPS_LOG2>
__global__ void __launch_bounds__(256)
segment_reduce_vec_kernel(const TI* __restrict__ src, const int64_t* __restrict__ order, const int64_t* __restrict__ start,
                          const int64_t* __restrict__ m_dev, int64_t m, int C, const float* __restrict__ scale,
                          const float* __restrict__ shift, int act, TO* __restrict__ out, const float* __restrict__ coord,
                          float* __restrict__ coord_out) {
  const int64_t M = m_dev ? *m_dev : m;
  const int lane = threadIdx.x & 31;
  constexpr int lps = 1 << LPS_LOG2, spw = 32 >> LPS_LOG2;
  const int g = lane & (lps - 1), sub = lane >> LPS_LOG2;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr float init = REDUCE == 2 ? -INFINITY : (REDUCE == 3 ? INFINITY : 0.f);
  for (int64_t s0 = warp0 * spw; s0 < M; s0 += nwarp * spw) {
    const int64_t s = s0 + sub;
    if (s >= M) continue;
    const int a = (int)start[s], b = (int)start[s + 1];
    const bool do_coord = coord != nullptr && g < 3;  // lanes 0..2 of the group also average one coordinate axis
    float cacc = 0.f;
    for (int c0 = g * 8; c0 < C; c0 += lps * 8) {
      float acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = init;
      const bool first = c0 == g * 8;
      int j = a;
      for (; j + 1 < b; j += 2) {
        const uint32_t r0 = (uint32_t)(order ? order[j] : j), r1 = (uint32_t)(order ? order[j + 1] : j + 1);
        float v0[8], v1[8];
        load8<TI>(src + (size_t)r0 * (uint32_t)C + c0, v0);
        load8<TI>(src + (size_t)r1 * (uint32_t)C + c0, v1);
        if (do_coord && first) cacc = (cacc + coord[(size_t)r0 * 3u + g]) + coord[(size_t)r1 * 3u + g];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = seg_op<REDUCE>(seg_op<REDUCE>(acc[u], v0[u]), v1[u]);
      }
      if (j < b) {
        const uint32_t r0 = (uint32_t)(order ? order[j] : j);
        float v0[8];
        load8<TI>(src + (size_t)r0 * (uint32_t)C + c0, v0);
        if (do_coord && first) cacc += coord[(size_t)r0 * 3u + g];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = seg_op<REDUCE>(acc[u], v0[u]);
      }
      if (b <= a) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = 0.f;
      } else if (REDUCE == 1) {
        const float cnt = (float)(b - a);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = acc[u] / cnt;
      }
      if (scale) {
        float sc[8], sh[8];
        load8<float>(scale + c0, sc);
        load8<float>(shift + c0, sh);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = acc[u] * sc[u] + sh[u];
      }
      if (act == 1) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = gelu_erf(acc[u]);
      }
      store8<TO>(out + (size_t)s * C + c0, acc);
    }
    if (do_coord) coord_out[(size_t)s * 3 + g] = b > a ? cacc / (float)(b - a) : 0.f;
  }
}

}  // namespace ss

extern "C" {

size_t ss_pool_workspace_bytes(int64_t n) { return ss::align_up(ss::runs_workspace_bytes(n, 5), 256) + 256; }

int ss_pool_index(const int64_t* code, const int64_t* order, const int64_t* grid_coord, const int64_t* batch, int64_t n,
                  int k, int pooling_depth, const int* src_row, int64_t m_cap, int64_t* cluster, int64_t* seg_start,
                  int64_t* head, int64_t* m_dev, int64_t* child_code, int64_t* child_order, int64_t* child_inverse,
                  int64_t* child_grid_coord, int64_t* child_batch, void* workspace, size_t workspace_bytes,
                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || k < 1 || k > 4 || pooling_depth < 0 || pooling_depth > 16 || m_cap < n || !src_row || !m_dev)
    return SS_BAD_ARGS;
  if (n == 0) {
    SS_CUDA(cudaMemsetAsync(m_dev, 0, 8, stream));
    return SS_OK;
  }
  if (!code || !order || !cluster || !seg_start || !head || !child_code || !child_order || !child_inverse || !workspace)
    return SS_BAD_ARGS;
  if (workspace_bytes < ss_pool_workspace_bytes(n) - 256) return SS_BAD_ARGS;
  int n0 = 0;
  for (int r = 0; r < k; ++r) {
    if (src_row[r] < 0 || src_row[r] >= k) return SS_BAD_ARGS;
    n0 += src_row[r] == 0;
  }
  if (n0 > 1) return SS_BAD_ARGS;  // src_row is a selection of distinct parent rows
  char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  ss::PoolRuns f;
  f.code = code; f.order = order; f.grid_coord = grid_coord; f.batch = batch; f.n = n; f.k = k;
  f.shift = 3 * pooling_depth; f.pool_depth = pooling_depth;
  f.src_row = make_int4(src_row[0], k > 1 ? src_row[1] : 0, k > 2 ? src_row[2] : 0, k > 3 ? src_row[3] : 0);
  f.m_cap = m_cap; f.cluster = cluster; f.seg_start = seg_start; f.head_idx = head; f.ccode = child_code;
  f.corder = child_order; f.cinverse = child_inverse; f.cgrid = child_grid_coord; f.cbatch = child_batch;
  int rows = 1;
  for (int y = 0; y < 5; ++y) f.rows_parent[y] = f.rows_child[y] = 0;
  for (int r = 0; r < k; ++r) {
    if (src_row[r] == 0) {
      f.rows_child[0] = r;
    } else {
      f.rows_parent[rows] = src_row[r];
      f.rows_child[rows] = r;
      ++rows;
    }
  }
  return ss::runs_launch(f, n, ws, m_dev, stream, rows);
}

static int segment_reduce_impl(const void* src, int src_is_bf16, const float* coord, const int64_t* order,
                               const int64_t* seg_start, const int64_t* m_dev, int64_t m, int channels, int reduce,
                               const float* scale, const float* shift, int act, void* out, int out_is_bf16,
                               float* coord_out, cudaStream_t stream) {
  if (m < 0 || channels < 1 || reduce < 0 || reduce > 3 || (scale && !shift) || (coord && !coord_out)) return SS_BAD_ARGS;
  if (m == 0) return SS_OK;
  if (!src || !seg_start || !out) return SS_BAD_ARGS;
  const bool vec = channels % 8 == 0 && (!coord || channels >= 32) &&  // the fused coordinates use 3 lanes of a group
                   (((uintptr_t)src | (uintptr_t)out | (uintptr_t)scale | (uintptr_t)shift) % 16 == 0);
  if (vec) {
    int lps_log2 = 0;
    while ((8 << lps_log2) < channels && lps_log2 < 5) ++lps_log2;
    const int64_t warps = ss::ceil_div64(m, 32 >> lps_log2);
    const int blocks = (int)ss::imin64(ss::ceil_div64(warps, 8), 16 * ss::kNumSMs);
#define SS_SEGV3_(TI, TO, R, L)                                                                                       \
  ss::segment_reduce_vec_kernel<TI, TO, R, L><<<blocks, 256, 0, stream>>>((const TI*)src, order, seg_start, m_dev, m,     \
                                                                          channels, scale, shift, act, (TO*)out, coord, \
                                                                          coord_out)
#define SS_SEGV2_(TI, TO, R)                    \
  do {                                          \
    switch (lps_log2) {                         \
      case 0: SS_SEGV3_(TI, TO, R, 0); break;   \
      case 1: SS_SEGV3_(TI, TO, R, 1); break;   \
      case 2: SS_SEGV3_(TI, TO, R, 2); break;   \
      case 3: SS_SEGV3_(TI, TO, R, 3); break;   \
      case 4: SS_SEGV3_(TI, TO, R, 4); break;   \
      default: SS_SEGV3_(TI, TO, R, 5); break;  \
    }                                           \
  } while (0)
#define SS_SEGV_(TI, TO)                        \
  do {                                          \
    switch (reduce) {                           \
      case 0: SS_SEGV2_(TI, TO, 0); break;      \
      case 1: SS_SEGV2_(TI, TO, 1); break;      \
      case 2: SS_SEGV2_(TI, TO, 2); break;      \
      default: SS_SEGV2_(TI, TO, 3); break;     \
    }                                           \
  } while (0)
    if (src_is_bf16 && out_is_bf16) SS_SEGV_(__nv_bfloat16, __nv_bfloat16);
    else if (src_is_bf16) SS_SEGV_(__nv_bfloat16, float);
    else if (out_is_bf16) SS_SEGV_(float, __nv_bfloat16);
    else SS_SEGV_(float, float);
#undef SS_SEGV3_
#undef SS_SEGV2_
#undef SS_SEGV_
    SS_CHECK_LAUNCH();
    return SS_OK;
  }
  const int blocks = (int)ss::imin64(ss::ceil_div64(m, 8), 16 * ss::kNumSMs);
  if (src_is_bf16 && out_is_bf16)
    ss::segment_reduce_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, stream>>>(
        (const __nv_bfloat16*)src, order, seg_start, m_dev, m, channels, reduce, scale, shift, act, (__nv_bfloat16*)out);
  else if (src_is_bf16)
    ss::segment_reduce_kernel<__nv_bfloat16, float><<<blocks, 256, 0, stream>>>(
        (const __nv_bfloat16*)src, order, seg_start, m_dev, m, channels, reduce, scale, shift, act, (float*)out);
  else if (out_is_bf16)
    ss::segment_reduce_kernel<float, __nv_bfloat16><<<blocks, 256, 0, stream>>>(
        (const float*)src, order, seg_start, m_dev, m, channels, reduce, scale, shift, act, (__nv_bfloat16*)out);
  else
    ss::segment_reduce_kernel<float, float><<<blocks, 256, 0, stream>>>(
        (const float*)src, order, seg_start, m_dev, m, channels, reduce, scale, shift, act, (float*)out);
  SS_CHECK_LAUNCH();
  if (coord) {  // scalar fallback: the coordinates in a second launch of the generic kernel
    ss::segment_reduce_kernel<float, float><<<blocks, 256, 0, stream>>>(coord, order, seg_start, m_dev, m, 3, 1, nullptr,
                                                                        nullptr, 0, coord_out);
    SS_CHECK_LAUNCH();
  }
  return SS_OK;
}

int ss_segment_reduce(const void* src, int src_is_bf16, const int64_t* order, const int64_t* seg_start,
                      const int64_t* m_dev, int64_t m, int channels, int reduce, const float* scale, const float* shift,
                      int act, void* out, int out_is_bf16, void* stream_) {
  return segment_reduce_impl(src, src_is_bf16, nullptr, order, seg_start, m_dev, m, channels, reduce, scale, shift, act,
                             out, out_is_bf16, nullptr, (cudaStream_t)stream_);
}

int ss_pool_reduce(const void* src, int src_is_bf16, const float* coord, const int64_t* order, const int64_t* seg_start,
                   int64_t m, int channels, int reduce, const float* scale, const float* shift, int act, void* out,
                   int out_is_bf16, float* coord_out, void* stream_) {
  if (!coord || !coord_out) return SS_BAD_ARGS;
  return segment_reduce_impl(src, src_is_bf16, coord, order, seg_start, nullptr, m, channels, reduce, scale, shift, act,
                             out, out_is_bf16, coord_out, (cudaStream_t)stream_);
}

int ss_unpool_gather_add(const void* a, const void* b, int in_is_bf16, const int64_t* cluster, int64_t n, int channels,
                         const float* scale_a, const float* shift_a, const float* scale_b, const float* shift_b, int act,
                         void* out, void* out_a, int out_flags, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int out_is_bf16 = out_flags & 1;
  if (n < 0 || channels < 1 || (scale_a && !shift_a) || (scale_b && !shift_b)) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!a || !b || !cluster || !out) return SS_BAD_ARGS;
  const bool vec = channels % 8 == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)out | (uintptr_t)out_a |
                                           (uintptr_t)scale_a | (uintptr_t)shift_a | (uintptr_t)scale_b |
                                           (uintptr_t)shift_b) % 16 == 0);
  const bool a_bf16 = (out_flags & 2) != 0 || out_is_bf16;  // bit 1: out_a in bf16 even when out is fp32
  if (vec) {
    const int vblocks = (int)ss::imin64(ss::ceil_div64(n * (channels / 8), 256), 16 * ss::kNumSMs);
#define SS_UNPOOL_V_(TI, TO, TA)                                                                                          \
  ss::unpool_vec8_kernel<TI, TO, TA><<<vblocks, 256, 0, stream>>>((const TI*)a, (const TI*)b, cluster, n, channels, scale_a, \
                                                                  shift_a, scale_b, shift_b, act, (TO*)out, (TA*)out_a)
#define SS_UNPOOL_VA_(TI, TO)                                   \
  do {                                                          \
    if (a_bf16) SS_UNPOOL_V_(TI, TO, __nv_bfloat16);            \
    else SS_UNPOOL_V_(TI, TO, float);                           \
  } while (0)
    if (in_is_bf16 && out_is_bf16) SS_UNPOOL_VA_(__nv_bfloat16, __nv_bfloat16);
    else if (in_is_bf16) SS_UNPOOL_VA_(__nv_bfloat16, float);
    else if (out_is_bf16) SS_UNPOOL_VA_(float, __nv_bfloat16);
    else SS_UNPOOL_VA_(float, float);
#undef SS_UNPOOL_VA_
#undef SS_UNPOOL_V_
    SS_CHECK_LAUNCH();
    return SS_OK;
  }
  if (a_bf16 != (out_is_bf16 != 0)) return SS_BAD_ARGS;  // the scalar fallback writes out and out_a in one dtype
  const int blocks = (int)ss::imin64(ss::ceil_div64(n * channels, 256), 16 * ss::kNumSMs);
  if (in_is_bf16 && out_is_bf16)
    ss::unpool_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, stream>>>(
        (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, cluster, n, channels, scale_a, shift_a, scale_b, shift_b, act,
        (__nv_bfloat16*)out, (__nv_bfloat16*)out_a);
  else if (in_is_bf16)
    ss::unpool_kernel<__nv_bfloat16, float><<<blocks, 256, 0, stream>>>(
        (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, cluster, n, channels, scale_a, shift_a, scale_b, shift_b, act,
        (float*)out, (float*)out_a);
  else if (out_is_bf16)
    ss::unpool_kernel<float, __nv_bfloat16><<<blocks, 256, 0, stream>>>(
        (const float*)a, (const float*)b, cluster, n, channels, scale_a, shift_a, scale_b, shift_b, act,
        (__nv_bfloat16*)out, (__nv_bfloat16*)out_a);
  else
    ss::unpool_kernel<float, float><<<blocks, 256, 0, stream>>>((const float*)a, (const float*)b, cluster, n, channels,
                                                                scale_a, shift_a, scale_b, shift_b, act, (float*)out,
                                                                (float*)out_a);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // extern "C"
