// Patch-wise serialized attention on the 5th-gen tensor cores (tcgen05 + TMEM), bf16 in / bf16 out.
//
// Replaces (reference): SerializedAttention.forward's `qkv[order]` gather, flash_attn_varlen_qkvpacked_func
// and `feat[inverse]` gather (point_transformer_v3m1_base.py:181-216); patch rule of :114-170 comes in as
// the device patch table (attention_simt.cu: patch_table_kernel).
//
// One CTA per (patch, head).  K and V of the head (<= 1024 tokens) stay resident in shared memory in the
// UMMA no-swizzle core-matrix layout (K: K-major, V: MN-major), gathered once through the serialized
// order with 16-byte cp.async.  Two softmax groups (4 warps each, one thread per query row) ping-pong
// over 128-row query blocks; each owns S [128 x 128] fp32 + O [128 x d] fp32 in TMEM:
//     pass 1   S = Q K_c^T for the 8 key chunks -> exact row max (no online rescaling, O is never touched
//              by CUDA cores until the end)
//     pass 2   S again -> p = exp2((s - m) * scale * log2e), row sum, P (bf16) written over S in TMEM,
//              O += P V_c with P as the TMEM A operand
//     end      O / l -> bf16 -> written to the point's own row (the [inverse] gather is fused)
// Warp 8 gathers (cp.async -> mbarrier), warp 9 issues every tcgen05.mma.  The kernel is exp-bound by
// design for head dims 16..48 (N*K*H exponentials >> MMA time), see DESIGN.md.
#include "tc_common.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

constexpr int kAttThreads = 320;  // 8 softmax warps + loader warp + MMA warp
constexpr int kQB = 128;          // query rows per block
constexpr int kKC = 128;          // keys per chunk

template <int D, int KMAX>
struct AttSmem {
  static constexpr int kK = KMAX * D * 2;
  static constexpr int kV = KMAX * D * 2;
  static constexpr int kQ = kQB * D * 2;  // per group
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kK;
  static constexpr int kOffQ = kK + kV;
  static constexpr int kOffBar = kOffQ + 2 * kQ;
  static constexpr int kTotal = kOffBar + 256 + 128;
};

template <int D, int KMAX>
__global__ void __launch_bounds__(kAttThreads, 1)
patch_attention_tc_kernel(const __nv_bfloat16* __restrict__ qkv, const int64_t* __restrict__ order_row,
                          const int4* __restrict__ table, int H, float scale_log2e, __nv_bfloat16* __restrict__ out) {
  using S = AttSmem<D, KMAX>;
  const int4 e = table[blockIdx.x];
  const int q_beg = e.x, n_q = e.y - e.x, kv_beg = e.z, kv_len = e.w - e.z;
  if (n_q <= 0) return;  // block-uniform: unused table entry
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  uint64_t* bars = (uint64_t*)(smem + S::kOffBar);
  uint64_t* kv_full = bars;        // [1]
  uint64_t* q_full = bars + 1;     // [2]
  uint64_t* q_free = bars + 3;     // [2]
  uint64_t* s_full = bars + 5;     // [2]
  uint64_t* s_done = bars + 7;     // [2]
  uint64_t* o_full = bars + 9;     // [2]
  uint64_t* o_free = bars + 11;    // [2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int C = H * D;
  const int nch = (kv_len + kKC - 1) / kKC;  // key chunks
  const int nqb = (n_q + kQB - 1) / kQB;     // query blocks
  const int n_it = (nqb + 1) / 2;            // group g handles query block 2*it + g

  if (threadIdx.x == 0) {
    tc::mbar_init(kv_full, 32);
    for (int g = 0; g < 2; ++g) {
      tc::mbar_init(&q_full[g], 32);
      tc::mbar_init(&q_free[g], 1);
      tc::mbar_init(&s_full[g], 1);
      tc::mbar_init(&s_done[g], 128);
      tc::mbar_init(&o_full[g], 1);
      tc::mbar_init(&o_free[g], 128);
    }
    tc::mbar_fence_init();
  }
  __syncthreads();
  // The loader starts gathering K/V right away; everybody else meets on named barrier 1 once TMEM is allocated
  // (a co-resident CTA may have to wait for the previous CTA's TMEM, its gathers overlap that wait).
  uint32_t tmem_base = 0;
  if (warp != 8) {
    if (warp == 9) tc::tmem_alloc<512>(tmem_slot);
    tc::tc_fence_before();
    asm volatile("bar.sync 1, 288;" ::: "memory");
    tc::tc_fence_after();
    tmem_base = *tmem_slot;
  }
  constexpr int kChunksPerRow = D / 8;  // 16-byte chunks per row

  if (warp < 8) {
    // =========================================================== softmax groups
    const int g = warp >> 2;
    const int row = (warp & 3) * 32 + lane;  // row inside the query block == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t tS = t_lane + g * 256;
    const uint32_t tO = t_lane + g * 256 + 128;
    uint32_t ph_s = 0, ph_o = 0;
    for (int it = 0; it < n_it; ++it) {
      const int qb = 2 * it + g;
      if (qb >= nqb) break;
      // ---- pass 1: exact row max
      float m = -INFINITY;
      for (int c = 0; c < nch; ++c) {
        tc::mbar_wait(&s_full[g], ph_s);
        ph_s ^= 1;
        tc::tc_fence_after();
        const int valid = min(kKC, kv_len - c * kKC);
#pragma unroll 1
        for (int j = 0; j < kKC / 32; ++j) {
          if (j * 32 >= valid) break;
          uint32_t v[32];
          tc::tmem_ld32(tS + j * 32, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 32; ++u)
            if (j * 32 + u < valid) m = fmaxf(m, __uint_as_float(v[u]));
        }
        tc::tc_fence_before();
        tc::mbar_arrive(&s_done[g]);
      }
      const float msc = m * scale_log2e;
      // ---- pass 2: probabilities, row sum, P -> TMEM (over S)
      float l = 0.f;
      for (int c = 0; c < nch; ++c) {
        tc::mbar_wait(&s_full[g], ph_s);
        ph_s ^= 1;
        tc::tc_fence_after();
        const int valid = min(kKC, kv_len - c * kKC);
#pragma unroll 1
        for (int j = 0; j < kKC / 32; ++j) {
          uint32_t v[32];
          tc::tmem_ld32(tS + j * 32, v);
          tc::tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            float p0 = exp2f(fmaf(__uint_as_float(v[2 * u]), scale_log2e, -msc));
            float p1 = exp2f(fmaf(__uint_as_float(v[2 * u + 1]), scale_log2e, -msc));
            if (j * 32 + 2 * u >= valid) p0 = 0.f;
            if (j * 32 + 2 * u + 1 >= valid) p1 = 0.f;
            l += p0 + p1;
            pk[u] = tc::pack_bf16(p0, p1);
          }
          tc::tmem_st16(tS + j * 16, pk);
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(&s_done[g]);
      }
      // ---- epilogue: O / l -> bf16 -> the point's own row
      tc::mbar_wait(&o_full[g], ph_o);
      ph_o ^= 1;
      tc::tc_fence_after();
      const int qi = qb * kQB + row;
      const float inv = 1.f / l;
      __nv_bfloat16* orow = nullptr;
      if (qi < n_q) orow = out + (size_t)order_row[q_beg + qi] * C + h * D;
#pragma unroll
      for (int j = 0; j < D / 16; ++j) {
        uint32_t v[16];
        tc::tmem_ld16(tO + j * 16, v);
        tc::tmem_ld_wait();
        if (orow) {
          uint4 o0, o1;
          o0.x = tc::pack_bf16(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
          o0.y = tc::pack_bf16(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
          o0.z = tc::pack_bf16(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
          o0.w = tc::pack_bf16(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
          o1.x = tc::pack_bf16(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv);
          o1.y = tc::pack_bf16(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv);
          o1.z = tc::pack_bf16(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv);
          o1.w = tc::pack_bf16(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv);
          uint4* dst = reinterpret_cast<uint4*>(orow + j * 16);
          dst[0] = o0;
          dst[1] = o1;
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&o_free[g]);
    }
  } else if (warp == 8) {
    // =========================================================== loader (cp.async gathers)
    // K / V: element (key j, 16-byte chunk c) -> K: c * (KMAX*16) + (j/8)*128 + (j%8)*16
    //                                            V: same formula (MN-major core matrices, see header)
    const int nkeys = nch * kKC;
    const uint32_t sK = tc::smem_u32(smem + S::kOffK), sV = tc::smem_u32(smem + S::kOffV);
    for (int idx = lane; idx < nkeys * kChunksPerRow; idx += 32) {
      const int j = idx / kChunksPerRow, c = idx - j * kChunksPerRow;
      const uint32_t off = (uint32_t)(c * (KMAX * 16) + (j >> 3) * 128 + (j & 7) * 16);
      const bool ok = j < kv_len;
      const __nv_bfloat16* src = qkv + (ok ? (size_t)order_row[kv_beg + j] * (3 * C) : 0) + h * D + c * 8;
      tc::cp_async16(sK + off, src + C, ok ? 16u : 0u);
      tc::cp_async16(sV + off, src + 2 * C, ok ? 16u : 0u);
    }
    tc::cp_async_mbar_arrive_noinc(kv_full);
    uint32_t ph_qfree[2] = {1, 1};  // a fresh barrier passes a parity-1 wait
    for (int it = 0; it < n_it; ++it) {
      for (int g = 0; g < 2; ++g) {
        const int qb = 2 * it + g;
        if (qb >= nqb) break;
        tc::mbar_wait(&q_free[g], ph_qfree[g]);
        ph_qfree[g] ^= 1;
        const uint32_t sQ = tc::smem_u32(smem + S::kOffQ + g * S::kQ);
        for (int idx = lane; idx < kQB * kChunksPerRow; idx += 32) {
          const int r = idx / kChunksPerRow, c = idx - r * kChunksPerRow;
          const int qi = qb * kQB + r;
          const bool ok = qi < n_q;
          const __nv_bfloat16* src = qkv + (ok ? (size_t)order_row[q_beg + qi] * (3 * C) : 0) + h * D + c * 8;
          tc::cp_async16(sQ + (uint32_t)(c * (kQB * 16) + (r >> 3) * 128 + (r & 7) * 16), src, ok ? 16u : 0u);
        }
        tc::cp_async_mbar_arrive_noinc(&q_full[g]);
      }
    }
  } else {
    // =========================================================== MMA issuer (one lane)
    if (lane == 0) {
      constexpr uint32_t idesc_s = tc::umma_idesc_bf16(kQB, kKC, 0, 0);  // S = Q K^T : M=128, N=128
      constexpr uint32_t idesc_o = tc::umma_idesc_bf16(kQB, D, 0, 1);    // O += P V : M=128, N=D, B MN-major
      const uint32_t sK = tc::smem_u32(smem + S::kOffK), sV = tc::smem_u32(smem + S::kOffV);
      tc::mbar_wait(kv_full, 0);
      uint32_t ph_qfull[2] = {0, 0}, ph_sdone[2] = {0, 0}, ph_ofree[2] = {0, 0};
      for (int it = 0; it < n_it; ++it) {
        const int ng = (2 * it + 1 < nqb) ? 2 : 1;
        for (int g = 0; g < ng; ++g) {
          tc::mbar_wait(&q_full[g], ph_qfull[g]);
          ph_qfull[g] ^= 1;
        }
        tc::tc_fence_after();
        for (int step = 0; step <= 2 * nch; ++step) {
          for (int g = 0; g < ng; ++g) {
            const uint32_t tS = tmem_base + g * 256, tO = tmem_base + g * 256 + 128;
            if (step > 0) {  // previous step's S consumed (pass 1) / P written (pass 2)
              tc::mbar_wait(&s_done[g], ph_sdone[g]);
              ph_sdone[g] ^= 1;
              tc::tc_fence_after();
            }
            if (step > nch) {  // O += P V for the previous pass-2 chunk
              const int pc = step - nch - 1;
              if (pc == 0 && it > 0) {  // the previous query block's O must have been read out
                tc::mbar_wait(&o_free[g], ph_ofree[g]);
                ph_ofree[g] ^= 1;
                tc::tc_fence_after();
              }
#pragma unroll
              for (int t = 0; t < kKC / 16; ++t) {
                const uint64_t dv = tc::umma_desc_nosw(sV + pc * (kKC / 8) * 128 + t * 256, 128, KMAX * 16);
                tc::umma_bf16_ts(tO, tS + t * 8, dv, idesc_o, (pc | t) ? 1u : 0u);
              }
            }
            if (step < 2 * nch) {
              const int c = step % nch;
              const uint32_t sQ = tc::smem_u32(smem + S::kOffQ + g * S::kQ);
#pragma unroll
              for (int t = 0; t < D / 16; ++t) {
                const uint64_t da = tc::umma_desc_nosw(sQ + 2 * t * (kQB * 16), kQB * 16, 128);
                const uint64_t db = tc::umma_desc_nosw(sK + c * (kKC / 8) * 128 + 2 * t * (KMAX * 16), KMAX * 16, 128);
                tc::umma_bf16(tS, da, db, idesc_s, t ? 1u : 0u);
              }
              tc::umma_commit(&s_full[g]);
              if (step == 2 * nch - 1) tc::umma_commit(&q_free[g]);  // last read of this Q block
            } else {
              tc::umma_commit(&o_full[g]);
            }
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

template <int D>
static int launch_attention(const void* qkv, const int64_t* order_row, const int32_t* table, int max_patches, int heads,
                            float scale, void* out, cudaStream_t stream) {
  constexpr int KMAX = 1024;
  using S = AttSmem<D, KMAX>;
  auto kern = patch_attention_tc_kernel<D, KMAX>;
  SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
  dim3 grid((unsigned)max_patches, (unsigned)heads);
  kern<<<grid, kAttThreads, S::kTotal, stream>>>((const __nv_bfloat16*)qkv, order_row, (const int4*)table, heads,
                                                 scale * 1.4426950408889634f, (__nv_bfloat16*)out);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // namespace ss

extern "C" int ss_patch_attention(const void* qkv_bf16, const int64_t* order_row, const int32_t* table, int max_patches,
                                  int patch_size, int heads, int head_dim, float scale, void* out_bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (max_patches < 0 || heads < 1 || patch_size < 1 || patch_size > 1024) return SS_BAD_ARGS;
  if (max_patches == 0) return SS_OK;
  if (!qkv_bf16 || !order_row || !table || !out_bf16) return SS_BAD_ARGS;
  if (((uintptr_t)qkv_bf16 | (uintptr_t)out_bf16) % 16 != 0) return SS_BAD_ARGS;
  switch (head_dim) {
    case 16: return ss::launch_attention<16>(qkv_bf16, order_row, table, max_patches, heads, scale, out_bf16, stream);
    case 32: return ss::launch_attention<32>(qkv_bf16, order_row, table, max_patches, heads, scale, out_bf16, stream);
    case 48: return ss::launch_attention<48>(qkv_bf16, order_row, table, max_patches, heads, scale, out_bf16, stream);
    default: return SS_BAD_ARGS;
  }
}
