// Patch-wise serialized attention on the 5th-gen tensor cores (tcgen05 + TMEM), bf16 in / bf16 out.
//
// Replaces (reference): SerializedAttention.forward's `qkv[order]` gather, flash_attn_varlen_qkvpacked_func
// and `feat[inverse]` gather (point_transformer_v3m1_base.py:181-216); patch rule of :114-170 comes in as
// the device patch table (attention_simt.cu: patch_table_kernel).
//
// One CTA per (patch, head).  K and V of the head (<= 1024 tokens) stay resident in shared memory in the
// UMMA no-swizzle core-matrix layout (K: K-major, V: MN-major), gathered once through the serialized
// order with 16-byte cp.async.  Query blocks of 128 rows are processed one after the other by 8 softmax
// warps: warp w owns TMEM lanes 32*(w%4).. (rows) and the column half (w/4) of every 128-key chunk.
// S chunks [128 x 128] fp32 live in a ring of 3 TMEM slots, so the MMA warp runs two chunks ahead of the
// softmax warps (no MMA round trip on the critical path):
//     pass 1   S = Q K_c^T for every key chunk -> exact row max (no online rescaling: O is never touched
//              by CUDA cores until the end)
//     pass 2   S again -> p = exp2((s - m) * scale * log2e), row sum, P (bf16) written over the consumed
//              half of the S slot, O += P V_c with P as the TMEM A operand
//     end      O / l -> bf16 -> written to the point's own row (the [inverse] gather is fused)
// Warp 8 gathers (cp.async -> mbarrier), warp 9 issues every tcgen05.mma.  The kernel is exp-bound by
// design for head dims 16..48 (N*K*H exponentials >> MMA time), see DESIGN.md.
#include "tc_common.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

constexpr int kAttThreads = 320;  // 8 softmax warps + loader warp + MMA warp
constexpr int kQB = 128;          // query rows per block
constexpr int kKC = 128;          // keys per chunk
constexpr int kSlots = 3;         // S ring
constexpr int kOCol = kSlots * kKC;

template <int D, int KMAX>
struct AttSmem {
  static constexpr int kK = KMAX * D * 2;
  static constexpr int kV = KMAX * D * 2;
  static constexpr int kQ = kQB * D * 2;  // per buffer (2 buffers)
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kK;
  static constexpr int kOffQ = kK + kV;
  static constexpr int kOffX = kOffQ + 2 * kQ;       // row max exchange [2][128] + row sum [128] floats
  static constexpr int kOffBar = kOffX + 3 * 128 * 4;
  static constexpr int kTotal = kOffBar + 256 + 128;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

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
  float* s_max = (float*)(smem + S::kOffX);  // [2][128]
  float* s_sum = s_max + 256;                // [128]
  uint64_t* bars = (uint64_t*)(smem + S::kOffBar);
  uint64_t* kv_full = bars;      // [1]
  uint64_t* q_full = bars + 1;   // [2]
  uint64_t* q_free = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;   // [3]
  uint64_t* s_done = bars + 8;   // [3]
  uint64_t* o_full = bars + 11;  // [1]
  uint64_t* o_free = bars + 12;  // [1]
  uint32_t* tmem_slot = (uint32_t*)(bars + 13);

  // warp index through a shuffle: the compiler then KNOWS it is warp-uniform (role branches stay convergent and
  // the MMA warp's descriptors can live in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int C = H * D;
  const int nch = (kv_len + kKC - 1) / kKC;  // key chunks
  const int nqb = (n_q + kQB - 1) / kQB;     // query blocks
  const int T = 2 * nch;                     // S productions per query block (pass 1 + pass 2)

  if (threadIdx.x == 0) {
    tc::mbar_init(kv_full, 32);
    for (int g = 0; g < 2; ++g) {
      tc::mbar_init(&q_full[g], 32);
      tc::mbar_init(&q_free[g], 1);
    }
    for (int s = 0; s < kSlots; ++s) {
      tc::mbar_init(&s_full[s], 1);
      tc::mbar_init(&s_done[s], 256);
    }
    tc::mbar_init(o_full, 1);
    tc::mbar_init(o_free, 128);
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
    tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  }
  constexpr int kChunksPerRow = D / 8;  // 16-byte chunks per row

  if (warp < 8) {
    // =========================================================== softmax warps
    // Software pipeline over the global step sequence G = qb * T + j (j < nch: pass 1, else pass 2): the
    // TMEM load of step G + 1 is in flight while step G is being computed (TMEM read bandwidth and the MUFU
    // work of a step are of the same order, so they must overlap).
    const int half = warp >> 2;
    const int row = (warp & 3) * 32 + lane;  // row inside the query block == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int col_h = 64 * half;  // this warp's columns of every S slot
    const int total = nqb * T;
    float m = -INFINITY, msc = 0.f, l = 0.f;

    auto issue_ld = [&](int G, uint32_t (&lo)[32], uint32_t (&hi)[32]) {
      const int slot = G % kSlots;
      tc::mbar_wait(&s_full[slot], (G / kSlots) & 1);
      tc::tc_fence_after();
      const uint32_t tS = t_lane + slot * kKC + col_h;
      tc::tmem_ld32(tS, lo);
      tc::tmem_ld32(tS + 32, hi);
    };

    auto process = [&](int G, uint32_t (&lo)[32], uint32_t (&hi)[32]) {
      const int qb = G / T, j = G - qb * T, slot = G % kSlots;
      const int c = j < nch ? j : j - nch;
      const int valid = min(kKC, kv_len - c * kKC);
      if (j < nch) {
        // ---- pass 1: exact row max over this warp's column half
        if (col_h + 64 <= valid) {
#pragma unroll
          for (int u = 0; u < 32; ++u) m = fmaxf(m, fmaxf(__uint_as_float(lo[u]), __uint_as_float(hi[u])));
        } else {
#pragma unroll
          for (int u = 0; u < 32; ++u) {
            if (col_h + u < valid) m = fmaxf(m, __uint_as_float(lo[u]));
            if (col_h + 32 + u < valid) m = fmaxf(m, __uint_as_float(hi[u]));
          }
        }
        tc::tc_fence_before();
        tc::mbar_arrive(&s_done[slot]);
        if (j == nch - 1) {  // combine the two column halves
          s_max[half * 128 + row] = m;
          asm volatile("bar.sync 2, 256;" ::: "memory");
          m = fmaxf(s_max[row], s_max[128 + row]);
          msc = m * scale_log2e;
          l = 0.f;
        }
      } else {
        // ---- pass 2: probabilities, row sum, P -> TMEM (over the consumed half of the S slot)
        const uint32_t tS = t_lane + slot * kKC + col_h;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int c0 = col_h + 32 * jj;
          uint32_t (&v)[32] = jj == 0 ? lo : hi;
          uint32_t pk[16];
          if (c0 + 32 <= valid) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              // P is rounded to bf16 on the integer pipe and the row sum is taken over the ROUNDED weights, so
              // O / l stays an exactly normalised convex combination
              const uint32_t b0 = tc::bf16_round_bits(ex2_approx(fmaf(__uint_as_float(v[2 * u]), scale_log2e, -msc)));
              const uint32_t b1 = tc::bf16_round_bits(ex2_approx(fmaf(__uint_as_float(v[2 * u + 1]), scale_log2e, -msc)));
              l += __uint_as_float(b0) + __uint_as_float(b1);
              pk[u] = tc::pack_bf16_bits(b0, b1);
            }
          } else {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              uint32_t b0 = tc::bf16_round_bits(ex2_approx(fmaf(__uint_as_float(v[2 * u]), scale_log2e, -msc)));
              uint32_t b1 = tc::bf16_round_bits(ex2_approx(fmaf(__uint_as_float(v[2 * u + 1]), scale_log2e, -msc)));
              if (c0 + 2 * u >= valid) b0 = 0u;
              if (c0 + 2 * u + 1 >= valid) b1 = 0u;
              l += __uint_as_float(b0) + __uint_as_float(b1);
              pk[u] = tc::pack_bf16_bits(b0, b1);
            }
          }
          tc::tmem_st16(tS + 16 * jj, pk);  // keys 64*half + 32*jj .. +31 -> 16 packed columns
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(&s_done[slot]);
        if (j == T - 1) {
          if (half == 1) s_sum[row] = l;
          asm volatile("bar.sync 2, 256;" ::: "memory");
          if (half == 0) {
            // ---- epilogue: O / l -> bf16 -> the point's own row
            l += s_sum[row];
            tc::mbar_wait(o_full, qb & 1);
            tc::tc_fence_after();
            const int qi = qb * kQB + row;
            const float inv = 1.f / l;
            __nv_bfloat16* orow = nullptr;
            if (qi < n_q) orow = out + (size_t)order_row[q_beg + qi] * C + h * D;
            const uint32_t tO = t_lane + kOCol;
#pragma unroll
            for (int jo = 0; jo < D / 16; ++jo) {
              uint32_t v[16];
              tc::tmem_ld16(tO + jo * 16, v);
              tc::tmem_ld_wait();  // (also drains the prefetched S load of the next step: harmless)
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
                uint4* dst = reinterpret_cast<uint4*>(orow + jo * 16);
                dst[0] = o0;
                dst[1] = o1;
              }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(o_free);
          }
          m = -INFINITY;  // next query block
        }
      }
    };

    uint32_t a_lo[32], a_hi[32], b_lo[32], b_hi[32];
    issue_ld(0, a_lo, a_hi);
    for (int G = 0; G < total; G += 2) {
      tc::tmem_ld_wait();
      if (G + 1 < total) issue_ld(G + 1, b_lo, b_hi);
      process(G, a_lo, a_hi);
      if (G + 1 < total) {
        tc::tmem_ld_wait();
        if (G + 2 < total) issue_ld(G + 2, a_lo, a_hi);
        process(G + 1, b_lo, b_hi);
      }
    }
  } else if (warp == 8) {
    // =========================================================== loader (cp.async gathers)
    // element (row j, 16-byte chunk c) -> c * (ROWS*16) + (j/8)*128 + (j%8)*16  (UMMA no-swizzle core matrices;
    // K and Q are K-major operands, V is an MN-major operand with the same byte layout)
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
    for (int qb = 0; qb < nqb; ++qb) {
      const int b = qb & 1;
      tc::mbar_wait(&q_free[b], ((qb >> 1) & 1) ^ 1);  // a fresh barrier passes a parity-1 wait
      const uint32_t sQ = tc::smem_u32(smem + S::kOffQ + b * S::kQ);
      for (int idx = lane; idx < kQB * kChunksPerRow; idx += 32) {
        const int r = idx / kChunksPerRow, c = idx - r * kChunksPerRow;
        const int qi = qb * kQB + r;
        const bool ok = qi < n_q;
        const __nv_bfloat16* src = qkv + (ok ? (size_t)order_row[q_beg + qi] * (3 * C) : 0) + h * D + c * 8;
        tc::cp_async16(sQ + (uint32_t)(c * (kQB * 16) + (r >> 3) * 128 + (r & 7) * 16), src, ok ? 16u : 0u);
      }
      tc::cp_async_mbar_arrive_noinc(&q_full[b]);
    }
  } else {
    // =========================================================== MMA issuer (whole warp, one elected lane per op)
    {
      constexpr uint32_t idesc_s = tc::umma_idesc_bf16(kQB, kKC, 0, 0);  // S = Q K^T : M=128, N=128
      constexpr uint32_t idesc_o = tc::umma_idesc_bf16(kQB, D, 0, 1);    // O += P V : M=128, N=D, B MN-major
      const uint32_t sK = tc::smem_u32(smem + S::kOffK), sV = tc::smem_u32(smem + S::kOffV);
      const uint32_t sQ0 = tc::smem_u32(smem + S::kOffQ);
      const uint32_t tO = tmem_base + kOCol;
      // descriptor bases (only the 14-bit start-address field changes per MMA)
      const uint64_t dq_base = tc::umma_desc_nosw(0, kQB * 16, 128);
      const uint64_t dk_base = tc::umma_desc_nosw(0, KMAX * 16, 128);
      const uint64_t dv_base = tc::umma_desc_nosw(0, 128, KMAX * 16);
      tc::mbar_wait(kv_full, 0);
      tc::tc_fence_after();
      const int total = nqb * T;
      int qb = 0, j = 0;        // position of step G
      int qb_r = 0, j_r = 0;    // position of step R = G - kSlots
      for (int G = 0; G < total + kSlots; ++G) {
        if (G >= kSlots) {
          const int R = G - kSlots, rslot = R % kSlots;
          tc::mbar_wait(&s_done[rslot], (R / kSlots) & 1);  // S consumed (pass 1) / P written (pass 2)
          tc::tc_fence_after();
          if (j_r >= nch) {  // O += P V for that pass-2 chunk
            const int pc = j_r - nch;
            if (pc == 0 && qb_r > 0) {  // the previous query block's O must have been read out
              tc::mbar_wait(o_free, (qb_r - 1) & 1);
              tc::tc_fence_after();
            }
            const uint32_t tP = tmem_base + rslot * kKC;
            const uint32_t v0 = (sV + pc * (kKC / 8) * 128) >> 4;
#pragma unroll
            for (int t = 0; t < kKC / 16; ++t) {
              const uint64_t dv = dv_base | (uint64_t)((v0 + t * 16) & 0x3fff);
              const uint32_t a = tP + (t < 4 ? 8 * t : 64 + 8 * (t - 4));  // P of column half 0 / half 1
              tc::umma_bf16_ts_elect(tO, a, dv, idesc_o, (pc | t) ? 1u : 0u);
            }
            if (pc == nch - 1) tc::umma_commit_elect(o_full);
          }
          if (++j_r == T) { j_r = 0; ++qb_r; }
        }
        if (G < total) {
          const int slot = G % kSlots, b = qb & 1;
          if (j == 0) {
            tc::mbar_wait(&q_full[b], (qb >> 1) & 1);
            tc::tc_fence_after();
          }
          const int c = j < nch ? j : j - nch;
          const uint32_t q0 = (sQ0 + b * S::kQ) >> 4;
          const uint32_t k0 = (sK + c * (kKC / 8) * 128) >> 4;
          const uint32_t tS = tmem_base + slot * kKC;
#pragma unroll
          for (int t = 0; t < D / 16; ++t) {
            const uint64_t da = dq_base | (uint64_t)((q0 + 2 * t * kQB) & 0x3fff);
            const uint64_t db = dk_base | (uint64_t)((k0 + 2 * t * KMAX) & 0x3fff);
            tc::umma_bf16_elect(tS, da, db, idesc_s, t ? 1u : 0u);
          }
          tc::umma_commit_elect(&s_full[slot]);
          if (j == T - 1) tc::umma_commit_elect(&q_free[b]);  // last read of this Q buffer
          if (++j == T) { j = 0; ++qb; }
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
