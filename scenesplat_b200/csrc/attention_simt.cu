// Patch-wise serialized attention, SIMT fp32 path (head dim 8/16/32/48/64, any patch size whose K/V fit
// shared memory).  It is the generic fall-back of the tcgen05 kernel (attention_tc.cu) and the
// second, independent GPU implementation the parity tests cross-check that kernel with.
//
// Replaces (reference): SerializedAttention.forward's gather -> flash_attn_varlen_qkvpacked_func ->
// gather (point_transformer_v3m1_base.py:181-216) and get_padding_and_inverse (:114-170).
// The `order[pad]` / `unpad[inverse]` gathers are fused: patch s reads its tokens through the
// serialized order row and writes every query's result straight to the point's own row.
#include "common.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

// Patch table (one int4 per patch: q_begin, q_end, kv_begin, kv_end in SORTED positions):
//   batch item with n <= K tokens -> one patch of n tokens
//   n > K -> ceil(n/K) patches; the last one owns rows [(P-1)K, n) and attends to the window [n-K, n)
// (reference ptv3:114-170).  Upper bound on patches: n_total / K + n_batch.
__global__ void patch_table_kernel(const int64_t* __restrict__ offset, int n_batch, int K, int max_patches,
                                   int4* __restrict__ table, int* __restrict__ n_patches_out) {
  // single CTA; batch counts are tiny
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int b = 0; b < n_batch; ++b) {
    const int64_t beg = b ? offset[b - 1] : 0, end = offset[b];
    const int n = (int)(end - beg);
    const int P = n <= K ? (n > 0 ? 1 : 0) : (n + K - 1) / K;
    const int base = s_base;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      int4 e;
      if (n <= K) {
        e = make_int4((int)beg, (int)end, (int)beg, (int)end);
      } else if (p < P - 1) {
        e = make_int4((int)beg + p * K, (int)beg + (p + 1) * K, (int)beg + p * K, (int)beg + (p + 1) * K);
      } else {
        e = make_int4((int)beg + (P - 1) * K, (int)end, (int)end - K, (int)end);
      }
      if (base + p < max_patches) table[base + p] = e;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + P;
    __syncthreads();
  }
  for (int p = s_base + threadIdx.x; p < max_patches; p += blockDim.x) table[p] = make_int4(0, 0, 0, 0);
  if (threadIdx.x == 0 && n_patches_out) *n_patches_out = s_base;
}

template <typename T> __device__ __forceinline__ float a_in(T v);
template <> __device__ __forceinline__ float a_in<float>(float v) { return v; }
template <> __device__ __forceinline__ float a_in<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T a_out(float v);
template <> __device__ __forceinline__ float a_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 a_out<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// grid = (max_patches, H); one CTA per (patch, head).  K/V of the head are staged in shared memory
// (fp32), every thread owns queries tid, tid + blockDim, ... and runs an online softmax over the keys.
template <typename TI, typename TO, int D>
__global__ void __launch_bounds__(256)
patch_attention_simt_kernel(const TI* __restrict__ qkv, const int64_t* __restrict__ order_row,
                            const int4* __restrict__ table, int H, float scale_log2e, TO* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char s_raw[];  // [kv_len][D] K then [kv_len][D] V, stored as TI
  TI* s_kv = reinterpret_cast<TI*>(s_raw);
  const int4 e = table[blockIdx.x];
  const int q_beg = e.x, q_end = e.y, kv_beg = e.z, kv_len = e.w - e.z;
  if (q_end <= q_beg) return;
  const int h = blockIdx.y;
  const int C = H * D;
  TI* sK = s_kv;
  TI* sV = s_kv + (size_t)kv_len * D;
  for (int i = threadIdx.x; i < kv_len * D; i += blockDim.x) {
    const int j = i / D, c = i - j * D;
    const TI* row = qkv + (size_t)order_row[kv_beg + j] * (3 * C);
    sK[i] = row[C + h * D + c];
    sV[i] = row[2 * C + h * D + c];
  }
  __syncthreads();
  for (int qi = q_beg + threadIdx.x; qi < q_end; qi += blockDim.x) {
    const int64_t prow = order_row[qi];
    const TI* row = qkv + (size_t)prow * (3 * C) + h * D;
    float q[D], acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
      q[c] = a_in<TI>(row[c]) * scale_log2e;
      acc[c] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int j0 = 0; j0 < kv_len; j0 += 8) {
      float s[8];
      float cm = m;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u;
        float d = -INFINITY;
        if (j < kv_len) {
          d = 0.f;
          const TI* kr = sK + (size_t)j * D;
#pragma unroll
          for (int c = 0; c < D; ++c) d = fmaf(q[c], a_in<TI>(kr[c]), d);
        }
        s[u] = d;
        cm = fmaxf(cm, d);
      }
      const float corr = exp2f(m - cm);
      l *= corr;
#pragma unroll
      for (int c = 0; c < D; ++c) acc[c] *= corr;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u;
        if (j < kv_len) {
          const float pj = exp2f(s[u] - cm);
          l += pj;
          const TI* vr = sV + (size_t)j * D;
#pragma unroll
          for (int c = 0; c < D; ++c) acc[c] = fmaf(pj, a_in<TI>(vr[c]), acc[c]);
        }
      }
      m = cm;
    }
    const float inv = 1.f / l;
    TO* orow = out + (size_t)prow * C + h * D;
#pragma unroll
    for (int c = 0; c < D; ++c) orow[c] = a_out<TO>(acc[c] * inv);
  }
}

}  // namespace ss

extern "C" {

int ss_patch_table(const int64_t* offset, int n_batch, int patch_size, int max_patches, int32_t* table,
                   int32_t* n_patches_dev, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!offset || n_batch < 1 || patch_size < 1 || max_patches < 1 || !table) return SS_BAD_ARGS;
  ss::patch_table_kernel<<<1, 256, 0, stream>>>(offset, n_batch, patch_size, max_patches, (int4*)table, n_patches_dev);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

int ss_patch_attention_simt(const void* qkv, int in_is_bf16, const int64_t* order_row, const int32_t* table,
                            int max_patches, int patch_size, int heads, int head_dim, float scale, void* out,
                            int out_is_bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (max_patches < 0 || heads < 1 || head_dim < 1 || patch_size < 1) return SS_BAD_ARGS;
  if (max_patches == 0) return SS_OK;
  if (!qkv || !order_row || !table || !out) return SS_BAD_ARGS;
  const size_t smem = (size_t)patch_size * head_dim * 2 * (in_is_bf16 ? 2 : 4);
  if (smem > 227 * 1024) return SS_BAD_ARGS;
  const float sl2 = scale * 1.4426950408889634f;
  dim3 grid(max_patches, heads);
#define SS_ATT_(TI, TO, D)                                                                                        \
  do {                                                                                                            \
    auto kern = ss::patch_attention_simt_kernel<TI, TO, D>;                                                       \
    SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                  \
    kern<<<grid, 256, smem, stream>>>((const TI*)qkv, order_row, (const int4*)table, heads, sl2, (TO*)out);       \
  } while (0)
#define SS_ATT_D_(TI, TO)                                   \
  do {                                                      \
    switch (head_dim) {                                     \
      case 8: SS_ATT_(TI, TO, 8); break;                    \
      case 16: SS_ATT_(TI, TO, 16); break;                  \
      case 32: SS_ATT_(TI, TO, 32); break;                  \
      case 48: SS_ATT_(TI, TO, 48); break;                  \
      case 64: SS_ATT_(TI, TO, 64); break;                  \
      default: return SS_BAD_ARGS;                          \
    }                                                       \
  } while (0)
  if (in_is_bf16 && out_is_bf16) SS_ATT_D_(__nv_bfloat16, __nv_bfloat16);
  else if (in_is_bf16) SS_ATT_D_(__nv_bfloat16, float);
  else if (out_is_bf16) SS_ATT_D_(float, __nv_bfloat16);
  else SS_ATT_D_(float, float);
#undef SS_ATT_D_
#undef SS_ATT_
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // extern "C"
