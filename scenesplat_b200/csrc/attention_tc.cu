// Patch-wise serialized attention on the 5th-gen tensor cores (tcgen05 + TMEM), bf16 in / bf16 out.
//
// Replaces (reference): SerializedAttention.forward's `qkv[order]` gather, flash_attn_varlen_qkvpacked_func
// and `feat[inverse]` gather (point_transformer_v3m1_base.py:181-216); patch rule of :114-170 comes in as
// the device patch table (attention_simt.cu: patch_table_kernel).
//
// One CTA per (head, patch).  K and V of the head (<= 1024 tokens) stay resident in shared memory in the
// UMMA no-swizzle core-matrix layout (K: K-major, V: MN-major), gathered once through the serialized
// order with 16-byte cp.async, chunk by chunk behind per-chunk mbarriers so the first MMAs start while
// the rest of K/V is still in flight.
//
// Single pass, online softmax, two query tiles of 128 rows in flight (ping-pong):
//   warps 0-3  softmax group 0 (tiles 0, 2, 4, ..), warps 4-7 softmax group 1 (tiles 1, 3, ..): one thread
//              owns one query row: the 128 scores of a key chunk come out of TMEM into registers, the
//              row max / rescale decision / row sum are thread-local (no shuffles, no shared memory)
//   warp 8     loader (cp.async gathers of Q tiles and K/V chunks -> mbarriers)
//   warp 9     issues every tcgen05.mma: S_g = Q_g K_c^T, then O_g += P_g V_c with P as the TMEM A operand
// While group 0 runs its exponentials the tensor pipe produces the next S of group 1 and vice versa, so
// neither the MMA round trip nor the TMEM traffic is on the critical path; the kernel is bound by the
// N*K*H exponentials (MUFU) for head dims 16..48, see DESIGN.md.
// TMEM columns: S_0/P_0 [0,128)  S_1/P_1 [128,256)  O_0 [256,320)  O_1 [320,384).  P (bf16) overwrites the
// first 64 columns of its own S tile; tcgen05.mma ops execute in issue order, so the next Q K^T of the
// group may be queued right behind the P V that reads those columns.
// Online softmax with lazy rescaling: the running reference max only moves when the chunk max exceeds it
// by more than 2^8 (any shift cancels in O / l; bf16 / fp32 have the exponent range), so O is touched by
// CUDA cores almost only on the first chunk(s) of a tile.
#include <cstdlib>
#include "tc_common.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

constexpr int kAttThreads = 384;  // 8 softmax warps + 2 Q loader warps + 2 MMA warps
constexpr int kQB = 128;          // query rows per tile
constexpr int kKC = 128;          // keys per chunk
constexpr int kPCol = 256;        // first P column (bf16 pairs: 64 columns per tile)
constexpr int kPStride = 64;
constexpr int kOCol = 384;        // first O column
constexpr int kOStride = 64;      // columns reserved per O tile (D of O, then the row-sum columns)
constexpr float kLazy = 8.f;      // log2 units the running max may lag behind

#ifdef SS_ATT_TRACE  // developer instrumentation (tools/micro/att_bench.cu): clock64 stamps per CTA
constexpr int kTraceSlots = 112, kTraceCtas = 2048;
__device__ long long g_att_trace[kTraceCtas * kTraceSlots];
#define ATT_TRACE(slot)                                                                             \
  do {                                                                                              \
    if (lane == 0 && blockIdx.x < kTraceCtas) g_att_trace[blockIdx.x * kTraceSlots + (slot)] = clock64(); \
  } while (0)
#else
#define ATT_TRACE(slot) do {} while (0)
#endif

template <int D, int KMAX>
struct AttSmem {
  static constexpr int kK = KMAX * D * 2;
  static constexpr int kV = KMAX * D * 2;
  static constexpr int kQ = kQB * D * 2;  // per buffer (one per softmax group)
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kK;
  static constexpr int kOffQ = kK + kV;
  static constexpr int kOffBar = kOffQ + 2 * kQ;
  static constexpr int kOffOnes = kOffBar + 256;  // 16 keys x 16 dims of bf16 1.0
  static constexpr int kTotal = kOffOnes + 512 + 128;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}

// VAR 0: P by cvt.rn.bf16x2 (F2FP), row sum by FADD
// VAR 2: P truncated to bf16 (PRMT only), row sum on the tensor core (P times a ones tile -> l columns behind O):
//        O / l is then an exactly normalised convex combination of the V rows with the weights actually used
template <int D, int KMAX, int VAR>
__global__ void __launch_bounds__(kAttThreads, 1)
patch_attention_tc_kernel(const __nv_bfloat16* __restrict__ qkv, const int64_t* __restrict__ order_row,
                          const int4* __restrict__ table, int H, float scale_log2e, __nv_bfloat16* __restrict__ out) {
  using S = AttSmem<D, KMAX>;
  const int4 e = table[blockIdx.x / H];
  const int q_beg = e.x, n_q = e.y - e.x, kv_beg = e.z, kv_len = e.w - e.z;
  if (n_q <= 0) return;  // block-uniform: unused table entry
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  uint64_t* bars = (uint64_t*)(smem + S::kOffBar);
  uint64_t* kv_full = bars;       // [8]  K/V chunk c landed (32 lane arrivals of the warp that gathered it)
  uint64_t* q_full = bars + 8;    // [2]  Q tile of group g landed
  uint64_t* q_free = bars + 10;   // [2]  last Q K^T of the tile done: buffer may be refilled
  uint64_t* s_full = bars + 12;   // [2]  S_g ready
  uint64_t* s_free = bars + 14;   // [2]  S_g is in registers (128 rows): the next Q K^T may overwrite it
  uint64_t* p_ready = bars + 16;  // [2]  P_g written by all 128 rows
  uint64_t* pv_done = bars + 18;  // [2]  P_g V done: P_g may be rewritten, O_g is quiescent
  uint32_t* tmem_slot = (uint32_t*)(bars + 20);

  // warp index through a shuffle: the compiler then KNOWS it is warp-uniform (role branches stay convergent and
  // the MMA warps' descriptors can live in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int h = blockIdx.x % H;
  const int C = H * D;
  const int nch = (kv_len + kKC - 1) / kKC;  // key chunks
  const int nqb = (n_q + kQB - 1) / kQB;     // query tiles
  constexpr int kChunksPerRow = D / 8;       // 16-byte pieces per row
  constexpr int kItems = 4 * kChunksPerRow;  // pieces per lane per unit of 128 rows

  if (threadIdx.x == 0) {
    for (int c = 0; c < 8; ++c) tc::mbar_init(&kv_full[c], 32);
    for (int g = 0; g < 2; ++g) {
      tc::mbar_init(&q_full[g], 32);
      tc::mbar_init(&q_free[g], 1);
      tc::mbar_init(&s_full[g], 1);
      tc::mbar_init(&s_free[g], 128);
      tc::mbar_init(&p_ready[g], 128);
      tc::mbar_init(&pv_done[g], 1);
    }
    tc::mbar_fence_init();
  }
  if (VAR == 2 && threadIdx.x >= 256 && threadIdx.x < 384) {
    reinterpret_cast<uint32_t*>(smem + S::kOffOnes)[threadIdx.x - 256] = 0x3f803f80u;
    tc::fence_proxy_async();
  }
  if (warp == 9) tc::tmem_alloc<512>(tmem_slot);

  // ---- gathers.  Lanes walk the 16-byte pieces of a row first (item = row * kChunksPerRow + c), so one warp
  // instruction touches 32 / kChunksPerRow rows.  Element (row j, piece c) -> c * (ROWS*16) + (j/8)*128 + (j%8)*16
  // (UMMA no-swizzle core matrices; K and Q are K-major operands, V is an MN-major operand, same byte layout).
  auto gather_kv = [&](int ch) {
    const uint32_t sK = tc::smem_u32(smem + S::kOffK), sV = tc::smem_u32(smem + S::kOffV);
#pragma unroll
    for (int i0 = 0; i0 < kItems; i0 += 4) {
      const __nv_bfloat16* src[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int item = lane + 32 * (i0 + u);
        const int r = item / kChunksPerRow, c = item - r * kChunksPerRow;
        const int j = ch * kKC + r;
        ok[u] = j < kv_len;
        src[u] = qkv + (ok[u] ? (size_t)order_row[kv_beg + j] * (3 * C) : 0) + h * D + c * 8;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int item = lane + 32 * (i0 + u);
        const int r = item / kChunksPerRow, c = item - r * kChunksPerRow;
        const int j = ch * kKC + r;
        const uint32_t off = (uint32_t)(c * (KMAX * 16) + (j >> 3) * 128 + (j & 7) * 16);
        tc::cp_async16(sK + off, src[u] + C, ok[u] ? 16u : 0u);
        tc::cp_async16(sV + off, src[u] + 2 * C, ok[u] ? 16u : 0u);
      }
    }
  };
  auto gather_q = [&](int t) {
    const uint32_t sQ = tc::smem_u32(smem + S::kOffQ + (t & 1) * S::kQ);
#pragma unroll
    for (int i0 = 0; i0 < kItems; i0 += 4) {
      const __nv_bfloat16* src[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int item = lane + 32 * (i0 + u);
        const int r = item / kChunksPerRow, c = item - r * kChunksPerRow;
        const int qi = t * kQB + r;
        ok[u] = qi < n_q;
        src[u] = qkv + (ok[u] ? (size_t)order_row[q_beg + qi] * (3 * C) : 0) + h * D + c * 8;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int item = lane + 32 * (i0 + u);
        const int r = item / kChunksPerRow, c = item - r * kChunksPerRow;
        tc::cp_async16(sQ + (uint32_t)(c * (kQB * 16) + (r >> 3) * 128 + (r & 7) * 16), src[u], ok[u] ? 16u : 0u);
      }
    }
  };
  // Prologue: the 8 softmax warps have nothing to do until the first S tile exists, so each of them gathers one
  // K/V chunk (all 8 chunks and both Q tiles are in flight at once); issued before the block-wide barrier so the
  // gathers overlap the TMEM allocation.
  if (warp < 8 && warp < nch) gather_kv(warp);
  if (warp == 8) gather_q(0);
  if (warp == 11 && nqb > 1) gather_q(1);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) ATT_TRACE(0);

  if (warp < 8) {
    // =========================================================== softmax warps
    if (warp < nch) {  // publish the chunk this warp gathered
      tc::cp_async_wait_all();
      tc::fence_proxy_async();
      tc::mbar_arrive(&kv_full[warp]);
    }
    const int g = warp >> 2;
    const int row = (warp & 3) * 32 + lane;  // row inside the query tile == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t tS = t_lane + g * kKC;
    const uint32_t tP = t_lane + kPCol + g * kPStride;
    const uint32_t tO = t_lane + kOCol + g * kOStride;
    const int ntiles = g == 0 ? (nqb + 1) / 2 : nqb / 2;
    int s = 0;  // step counter of this group (phase of s_full / s_free / p_ready / pv_done)
    for (int i = 0; i < ntiles; ++i) {
      float msc = -INFINITY;  // running reference max (log2 domain, already scaled)
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      for (int j = 0; j < nch; ++j, ++s) {
        uint32_t v[4][32];
        tc::mbar_wait(&s_full[g], s & 1);
        tc::tc_fence_after();
        if (warp == 0 && s == 0) ATT_TRACE(10);
        const bool tr_on = (warp & 3) == 0 && s >= 8 && s < 16;
        [[maybe_unused]] const int tr_base = 16 + g * 32 + (s - 8) * 4;
        if (tr_on) ATT_TRACE(tr_base);
        tc::tmem_ld32(tS, v[0]);
        tc::tmem_ld32(tS + 32, v[1]);
        tc::tmem_ld32(tS + 64, v[2]);
        tc::tmem_ld32(tS + 96, v[3]);
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(&s_free[g]);  // the next Q K^T of the group runs under this step's exponentials
        if (tr_on) ATT_TRACE(tr_base + 1);
        const int valid = kv_len - j * kKC;
        if (valid < kKC) {  // warp-uniform: ragged last chunk of a short sequence
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int u = 0; u < 32; ++u)
              if (q * 32 + u >= valid) v[q][u] = 0xff800000u;  // -inf: exp2 -> 0, ignored by the max
        }
        float mx[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          mx[q] = fmax3(__uint_as_float(v[q][0]), __uint_as_float(v[q][1]), __uint_as_float(v[q][2]));
#pragma unroll
          for (int u = 3; u < 31; u += 2) mx[q] = fmax3(mx[q], __uint_as_float(v[q][u]), __uint_as_float(v[q][u + 1]));
          mx[q] = fmaxf(mx[q], __uint_as_float(v[q][31]));
        }
        const float nm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * scale_log2e;  // scale > 0
        const bool need = nm > msc + kLazy;
        if (__any_sync(0xffffffffu, need)) {
          const float newm = need ? nm : msc;
          const float f = ex2_approx(msc - newm);  // first chunk: exp2(-inf) = 0
          msc = newm;
          l0 *= f; l1 *= f; l2 *= f; l3 *= f;
          if (j > 0) {
            tc::mbar_wait(&pv_done[g], (s - 1) & 1);  // O_g must be quiescent
            tc::tc_fence_after();
#pragma unroll
            for (int jo = 0; jo < (VAR == 2 ? D / 16 + 1 : D / 16); ++jo) {
              uint32_t o[16];
              tc::tmem_ld16(tO + jo * 16, o);
              tc::tmem_ld_wait();
#pragma unroll
              for (int u = 0; u < 16; ++u) o[u] = __float_as_uint(__uint_as_float(o[u]) * f);
              tc::tmem_st16(tO + jo * 16, o);
            }
          }
        }
        if (tr_on) ATT_TRACE(tr_base + 2);
        const float nmsc = -msc;
        uint32_t pk[4][16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(v[q][2 * u]), scale_log2e, nmsc));
            const float p1 = ex2_approx(fmaf(__uint_as_float(v[q][2 * u + 1]), scale_log2e, nmsc));
            if (VAR == 0) {
              if (u & 1) { l2 += p0; l3 += p1; } else { l0 += p0; l1 += p1; }
              pk[q][u] = tc::pack_bf16(p0, p1);
            } else {
              pk[q][u] = tc::pack_bf16_bits(__float_as_uint(p0), __float_as_uint(p1));
            }
          }
        }
        if (j > 0) {  // P_g is still being read by the previous P V of the group (j == 0: the epilogue waited)
          tc::mbar_wait(&pv_done[g], (s - 1) & 1);
          tc::tc_fence_after();
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) tc::tmem_st16(tP + q * 16, pk[q]);  // keys 32q .. 32q+31 -> 16 packed columns
        tc::tmem_st_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(&p_ready[g]);
        if (tr_on) ATT_TRACE(tr_base + 3);
      }
      // ---- epilogue of the tile: O / l -> bf16 -> the point's own row (the [inverse] gather is fused)
      tc::mbar_wait(&pv_done[g], (s - 1) & 1);
      tc::tc_fence_after();
      const int qi = (2 * i + g) * kQB + row;
      float lsum = (l0 + l1) + (l2 + l3);
      if (VAR == 2) {
        lsum = __uint_as_float(tc::tmem_ld1(tO + D));
        tc::tmem_ld_wait();
      }
      const float inv = 1.f / lsum;
      __nv_bfloat16* orow = nullptr;
      if (qi < n_q) orow = out + (size_t)order_row[q_beg + qi] * C + h * D;
#pragma unroll
      for (int jo = 0; jo < D / 16; ++jo) {
        uint32_t o[16];
        tc::tmem_ld16(tO + jo * 16, o);
        tc::tmem_ld_wait();
        if (orow) {
          uint4 o0, o1;
          o0.x = tc::pack_bf16(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
          o0.y = tc::pack_bf16(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
          o0.z = tc::pack_bf16(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
          o0.w = tc::pack_bf16(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
          o1.x = tc::pack_bf16(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
          o1.y = tc::pack_bf16(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
          o1.z = tc::pack_bf16(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
          o1.w = tc::pack_bf16(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
          uint4* dst = reinterpret_cast<uint4*>(orow + jo * 16);
          dst[0] = o0;
          dst[1] = o1;
        }
      }
      tc::tc_fence_before();  // ordered before the next tile's first P V by the next p_ready arrival
      if (warp == 0 && i == 0) ATT_TRACE(11);
    }
    if (warp == 0) ATT_TRACE(12);
  } else if (warp == 8 || warp == 11) {
    // =========================================================== Q loaders (warp 8: group 0, warp 11: group 1)
    const int g = warp == 8 ? 0 : 1;
    const int ntiles = g == 0 ? (nqb + 1) / 2 : nqb / 2;
    for (int i = 0; i < ntiles; ++i) {
      if (i > 0) {
        tc::mbar_wait_sleep(&q_free[g], (i - 1) & 1);
        gather_q(2 * i + g);
      }
      tc::cp_async_wait_all();
      tc::fence_proxy_async();
      tc::mbar_arrive(&q_full[g]);
    }
    if (warp == 8) ATT_TRACE(15);
  } else {
    // =========================================================== MMA issuers (warp 9: group 0, warp 10: group 1;
    // whole warp in uniform control flow, one elected lane per op)
    const int g = warp - 9;
    const int total = (g == 0 ? (nqb + 1) / 2 : nqb / 2) * nch;
    constexpr uint32_t idesc_s = tc::umma_idesc_bf16(kQB, kKC, 0, 0);  // S = Q K^T : M=128, N=128
    constexpr uint32_t idesc_o = tc::umma_idesc_bf16(kQB, D, 0, 1);    // O += P V : M=128, N=D, B MN-major
    constexpr uint32_t idesc_l = tc::umma_idesc_bf16(kQB, 16, 0, 1);   // l += P 1 : N = 16 replicated columns
    const uint32_t sK = tc::smem_u32(smem + S::kOffK), sV = tc::smem_u32(smem + S::kOffV);
    const uint32_t q0 = tc::smem_u32(smem + S::kOffQ + g * S::kQ) >> 4;
    const uint32_t tSg = tmem_base + g * kKC;
    const uint32_t tPg = tmem_base + kPCol + g * kPStride;
    const uint32_t tOg = tmem_base + kOCol + g * kOStride;
    // descriptor bases (only the 14-bit start-address field changes per MMA)
    const uint64_t dq_base = tc::umma_desc_nosw(0, kQB * 16, 128);
    const uint64_t dk_base = tc::umma_desc_nosw(0, KMAX * 16, 128);
    const uint64_t dv_base = tc::umma_desc_nosw(0, 128, KMAX * 16);
    const uint64_t d_ones = tc::umma_desc_nosw(tc::smem_u32(smem + S::kOffOnes), 128, 256);
    int kv_ready = 0;    // chunks known to have landed
    int jn = 0, tn = 0;  // chunk / tile index of the NEXT Q K^T

    auto issue_qk = [&]() {
      const int j = jn;
      if (j == 0) {
        tc::mbar_wait(&q_full[g], tn & 1);
        tc::tc_fence_after();
      }
      if (kv_ready <= j) {
        tc::mbar_wait(&kv_full[j], 0);
        tc::tc_fence_after();
        kv_ready = j + 1;
        if (g == 0) ATT_TRACE(2 + j);
      }
      const uint32_t k0 = (sK + j * (kKC / 8) * 128) >> 4;
#pragma unroll
      for (int t = 0; t < D / 16; ++t) {
        const uint64_t da = dq_base | (uint64_t)((q0 + 2 * t * kQB) & 0x3fff);
        const uint64_t db = dk_base | (uint64_t)((k0 + 2 * t * KMAX) & 0x3fff);
        tc::umma_bf16_elect(tSg, da, db, idesc_s, t ? 1u : 0u);
      }
      tc::umma_commit_elect(&s_full[g]);
      if (j == nch - 1) {
        tc::umma_commit_elect(&q_free[g]);  // last read of this Q buffer
        jn = 0;
        ++tn;
      } else {
        jn = j + 1;
      }
    };

    if (total > 0) issue_qk();
    int pc = 0;  // chunk index of the next P V
    for (int s = 0; s < total; ++s) {
      if (s + 1 < total) {
        tc::mbar_wait(&s_free[g], s & 1);
        tc::tc_fence_after();
        issue_qk();
      }
      tc::mbar_wait(&p_ready[g], s & 1);
      tc::tc_fence_after();
      if (s >= 8 && s < 16) ATT_TRACE(80 + (s - 8) * 4 + g * 2);
      const uint32_t v0 = (sV + pc * (kKC / 8) * 128) >> 4;
#pragma unroll
      for (int t = 0; t < kKC / 16; ++t) {
        const uint64_t dv = dv_base | (uint64_t)((v0 + t * 16) & 0x3fff);
        tc::umma_bf16_ts_elect(tOg, tPg + 8 * t, dv, idesc_o, (pc | t) ? 1u : 0u);
      }
      if (VAR == 2) {
#pragma unroll
        for (int t = 0; t < kKC / 16; ++t)
          tc::umma_bf16_ts_elect(tOg + D, tPg + 8 * t, d_ones, idesc_l, (pc | t) ? 1u : 0u);
      }
      tc::umma_commit_elect(&pv_done[g]);
      pc = pc == nch - 1 ? 0 : pc + 1;
      if (s >= 8 && s < 16) ATT_TRACE(80 + (s - 8) * 4 + g * 2 + 1);
    }
    if (warp == 9) ATT_TRACE(14);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ATT_TRACE(13);
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

template <int D, int VAR>
static int launch_attention_var(const void* qkv, const int64_t* order_row, const int32_t* table, int max_patches,
                                int heads, float scale, void* out, cudaStream_t stream) {
  constexpr int KMAX = 1024;
  using S = AttSmem<D, KMAX>;
  auto kern = patch_attention_tc_kernel<D, KMAX, VAR>;
  SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
  // heads fastest: the H CTAs of a patch run together and share the gathered rows' DRAM sectors through L2
  dim3 grid((unsigned)((size_t)heads * max_patches));
  kern<<<grid, kAttThreads, S::kTotal, stream>>>((const __nv_bfloat16*)qkv, order_row, (const int4*)table, heads,
                                                 scale * 1.4426950408889634f, (__nv_bfloat16*)out);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

template <int D>
static int launch_attention(const void* qkv, const int64_t* order_row, const int32_t* table, int max_patches, int heads,
                            float scale, void* out, cudaStream_t stream) {
  static const int var = getenv("SS_ATT_VAR") ? atoi(getenv("SS_ATT_VAR")) : 2;  // tuning hook
  switch (var) {
    case 0: return launch_attention_var<D, 0>(qkv, order_row, table, max_patches, heads, scale, out, stream);
    default: return launch_attention_var<D, 2>(qkv, order_row, table, max_patches, heads, scale, out, stream);
  }
}

}  // namespace ss

extern "C" int ss_patch_attention(const void* qkv_bf16, const int64_t* order_row, const int32_t* table, int max_patches,
                                  int patch_size, int heads, int head_dim, float scale, void* out_bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (max_patches < 0 || heads < 1 || patch_size < 1 || patch_size > 1024 || !(scale > 0.f)) return SS_BAD_ARGS;
  if (max_patches == 0) return SS_OK;
  if ((long long)max_patches * heads > 0x7fffffffLL) return SS_BAD_ARGS;
  if (!qkv_bf16 || !order_row || !table || !out_bf16) return SS_BAD_ARGS;
  if (((uintptr_t)qkv_bf16 | (uintptr_t)out_bf16) % 16 != 0) return SS_BAD_ARGS;
  switch (head_dim) {
    case 16: return ss::launch_attention<16>(qkv_bf16, order_row, table, max_patches, heads, scale, out_bf16, stream);
    case 32: return ss::launch_attention<32>(qkv_bf16, order_row, table, max_patches, heads, scale, out_bf16, stream);
    case 48: return ss::launch_attention<48>(qkv_bf16, order_row, table, max_patches, heads, scale, out_bf16, stream);
    default: return SS_BAD_ARGS;
  }
}
