// Patch-wise serialized attention on the 5th-gen tensor cores (tcgen05 + TMEM), bf16 in / bf16 out.
//
// Replaces (reference): SerializedAttention.forward's `qkv[order]` gather, flash_attn_varlen_qkvpacked_func
// and `feat[inverse]` gather (point_transformer_v3m1_base.py:181-216); patch rule of :114-170 comes in as
// the device patch table (attention_simt.cu: patch_table_kernel).
//
// One CTA per (head, patch), 12 warps.  K and V of the head (<= 1024 tokens) stay resident in shared memory in the
// UMMA no-swizzle core-matrix layout (K: K-major, V: MN-major).  They are gathered through the serialized order
// with 16-byte cp.async by the 8 softmax warps themselves (one 128-key chunk each, all in flight at once: those
// warps have nothing else to do before the first S tile exists) behind per-chunk mbarriers.
//
// Single pass, online softmax, two query tiles of 128 rows in flight (one per softmax group), 64 keys per step, and
// TWO score / weight buffers per group, so that a group's tensor-core work for step s + 1 and s + 2 is already done
// (or under way) while its softmax warps work on step s: no softmax warp ever waits for an MMA round trip in steady
// state, and the MMA warp never waits for the softmax either (round 1's kernel had one S buffer per group: every
// step exposed ~1000 clk of mbarrier / TMEM / MMA latency, profiles/r1_attention_ncu.md).
//   warps 0-3   softmax group 0 (tiles 0, 2, 4, ..), warps 4-7 softmax group 1 (tiles 1, 3, ..): one thread owns
//               one query row; the 64 scores of a step come out of TMEM into registers ONCE, the row max and the
//               rescale decision are thread-local (no shuffles, no shared memory)
//   warps 8,11  Q loaders of group 0 / 1: the NEXT tile's rows are prefetched into registers and stored the moment
//               the last Q K^T of the current tile has released the (single) Q buffer of the group
//   warps 9,10  MMA issuers of group 0 / 1: S_g[b] = Q_g K_c^T two steps ahead of the softmax, O_g += P_g[b] V_c and
//               l_g += P_g[b] 1 once P_g[b] is written
// TMEM columns: S_g[b] (g*2+b)*64 in [0,256) | P_g[b] 256 + (g*2+b)*32 in [256,384) (bf16 pairs) | O_g, l_g
// 384 + g*64 in [384,512).
// P is CUT to bf16 (no rounding instruction) and the row sum l is accumulated by the tensor core from the same
// bf16 weights (P times a 16x16 tile of ones): O / l is an exactly normalised convex combination of V rows.
// Online softmax with lazy rescaling: the running reference max only moves when the step max exceeds it by more
// than 2^8 (any shift cancels in O / l; bf16 / fp32 have the exponent range), so O is touched by CUDA cores almost
// only on the first step(s) of a tile.  The kernel is bound by the N*K*H exponentials, not by the tensor pipe
// (see DESIGN.md); a fraction of them is evaluated on the FMA pipe (exp2_poly).
#include "tc_common.cuh"
#include "attention_math.cuh"
#include "../../include/scenesplat_b200.h"

#ifndef SS_ATT_POLY
#define SS_ATT_POLY 2
#endif

namespace ss {
namespace att {

constexpr int kThreads = 384;  // 8 softmax warps + 2 Q loader warps + 2 MMA warps
constexpr int kQB = 128;       // query rows per tile
constexpr int kKS = 64;        // keys per step
constexpr int kKG = 128;       // keys per prologue gather chunk (one softmax warp each)
constexpr int kSCol = 0;       // S_g[b] at kSCol + (g * 2 + b) * 64
constexpr int kPCol = 256;     // P_g[b] at kPCol + (g * 2 + b) * 32
constexpr int kOCol = 384;     // O_g at kOCol + g * 64 (D columns of O, then 16 replicated row-sum columns)
constexpr float kLazy = 8.f;   // log2 units the running max may lag behind

template <int D, int KMAX>
struct Smem {
  static constexpr int kK = KMAX * D * 2;
  static constexpr int kV = KMAX * D * 2;
  static constexpr int kQ = kQB * D * 2;  // per buffer (one per softmax group)
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kK;
  static constexpr int kOffQ = kK + kV;
  static constexpr int kOffBar = kOffQ + 2 * kQ;
  static constexpr int kOffOnes = kOffBar + 384;  // 16 keys x 16 dims of bf16 1.0
  static constexpr int kTotal = kOffOnes + 512 + 128;
};

// POLY: of every 8 exponentials, POLY are evaluated by exp2_poly on the FMA pipe and 8 - POLY by MUFU.EX2
template <int D, int KMAX, int POLY>
__global__ void __launch_bounds__(kThreads, 1)
patch_attention_kernel(const __nv_bfloat16* __restrict__ qkv, const int64_t* __restrict__ order_row,
                       const int4* __restrict__ table, int H, float scale_log2e, __nv_bfloat16* __restrict__ out,
                       float* __restrict__ lse2, int64_t lse_stride) {
  using S = Smem<D, KMAX>;
  const int4 e = table[blockIdx.x / H];
  const int q_beg = e.x, n_q = e.y - e.x, kv_beg = e.z, kv_len = e.w - e.z;
  if (n_q <= 0) return;  // block-uniform: unused table entry
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  uint64_t* bars = (uint64_t*)(smem + S::kOffBar);
  uint64_t* kv_full = bars;       // [8]    K/V gather chunk c landed (32 lane arrivals of the warp that gathered it)
  uint64_t* q_full = bars + 8;    // [2]    Q tile of group g landed
  uint64_t* q_free = bars + 10;   // [2]    last Q K^T of the tile done: buffer may be refilled
  uint64_t* s_full = bars + 12;   // [2][2] S_g[b] ready
  uint64_t* s_free = bars + 16;   // [2][2] S_g[b] is in registers (128 rows): a later Q K^T may overwrite it
  uint64_t* p_ready = bars + 20;  // [2][2] P_g[b] written by all 128 rows
  uint64_t* pv_done = bars + 24;  // [2][2] P_g[b] V done: P_g[b] may be rewritten, O_g is quiescent up to that step
  uint32_t* tmem_slot = (uint32_t*)(bars + 28);

  // warp index through a shuffle: the compiler then KNOWS it is warp-uniform (role branches stay convergent and
  // the MMA warps' descriptors can live in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int h = blockIdx.x % H;
  const int C = H * D;
  const int nst = (kv_len + kKS - 1) / kKS;  // steps (64-key chunks) per query tile
  const int ngc = (kv_len + kKG - 1) / kKG;  // prologue gather chunks
  const int nqb = (n_q + kQB - 1) / kQB;     // query tiles
  constexpr int kChunksPerRow = D / 8;       // 16-byte pieces per row
  constexpr int kItems = 4 * kChunksPerRow;  // pieces per lane per unit of 128 rows

  if (threadIdx.x == 0) {
    for (int c = 0; c < 8; ++c) tc::mbar_init(&kv_full[c], 32);
    for (int g = 0; g < 2; ++g) {
      tc::mbar_init(&q_full[g], 32);
      tc::mbar_init(&q_free[g], 1);
    }
    for (int i = 0; i < 4; ++i) {
      tc::mbar_init(&s_full[i], 1);
      tc::mbar_init(&s_free[i], 128);
      tc::mbar_init(&p_ready[i], 128);
      tc::mbar_init(&pv_done[i], 1);
    }
    tc::mbar_fence_init();
  }
  if (threadIdx.x >= 256 && threadIdx.x < 384) {
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
        const int j = ch * kKG + r;
        ok[u] = j < kv_len;
        src[u] = qkv + (ok[u] ? (size_t)order_row[kv_beg + j] * (3 * C) : 0) + h * D + c * 8;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int item = lane + 32 * (i0 + u);
        const int r = item / kChunksPerRow, c = item - r * kChunksPerRow;
        const int j = ch * kKG + r;
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
  if (warp < 8 && warp < ngc) gather_kv(warp);
  if (warp == 8) gather_q(0);
  if (warp == 11 && nqb > 1) gather_q(1);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // =========================================================== softmax warps
    if (warp < ngc) {  // publish the chunk this warp gathered
      tc::cp_async_wait_all();
      tc::fence_proxy_async();
      tc::mbar_arrive(&kv_full[warp]);
    }
    const int g = warp >> 2;
    const int row = (warp & 3) * 32 + lane;  // row inside the query tile == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t tS = t_lane + kSCol + g * 128;  // + b * 64
    const uint32_t tP = t_lane + kPCol + g * 64;   // + b * 32
    const uint32_t tO = t_lane + kOCol + g * 64;
    const int ntiles = g == 0 ? (nqb + 1) / 2 : nqb / 2;
    int s = 0;  // step counter of this group: buffer b = s & 1, use number (phase) = (s >> 1) & 1
    for (int i = 0; i < ntiles; ++i) {
      float msc = -INFINITY;  // running reference max (log2 domain, already scaled)
      const int qi = (2 * i + g) * kQB + row;
      const int64_t out_row = qi < n_q ? order_row[q_beg + qi] : -1;  // fetched now, needed by the tile's epilogue
      for (int j = 0; j < nst; ++j, ++s) {
        const int b = s & 1;
        const int gb = g * 2 + b;
        uint32_t v[2][32];
        tc::mbar_wait(&s_full[gb], (s >> 1) & 1);
        tc::tc_fence_after();
        tc::tmem_ld32(tS + b * 64, v[0]);
        tc::tmem_ld32(tS + b * 64 + 32, v[1]);
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(&s_free[gb]);  // the Q K^T of step s + 2 may overwrite S_g[b]
        const int valid = kv_len - j * kKS;
        if (valid < kKS) {  // warp-uniform: ragged last step of a short sequence
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int u = 0; u < 32; ++u)
              if (q * 32 + u >= valid) v[q][u] = 0xff800000u;  // -inf: exp2 -> 0, ignored by the max
        }
        float mx[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float m0 = fmax3(__uint_as_float(v[q][0]), __uint_as_float(v[q][1]), __uint_as_float(v[q][2]));
          float m1 = fmax3(__uint_as_float(v[q][3]), __uint_as_float(v[q][4]), __uint_as_float(v[q][5]));
#pragma unroll
          for (int u = 6; u < 30; u += 4) {
            m0 = fmax3(m0, __uint_as_float(v[q][u]), __uint_as_float(v[q][u + 1]));
            m1 = fmax3(m1, __uint_as_float(v[q][u + 2]), __uint_as_float(v[q][u + 3]));
          }
          mx[q] = fmax3(m0, m1, fmaxf(__uint_as_float(v[q][30]), __uint_as_float(v[q][31])));
        }
        const float nm = fmaxf(mx[0], mx[1]) * scale_log2e;  // scale > 0
        const bool need = nm > msc + kLazy;
        if (__any_sync(0xffffffffu, need)) {
          const float newm = need ? nm : msc;
          const float f = ex2_approx(msc - newm);  // first step of a tile: exp2(-inf) = 0
          msc = newm;
          if (j > 0) {
            tc::mbar_wait(&pv_done[g * 2 + ((s - 1) & 1)], ((s - 1) >> 1) & 1);  // O_g quiescent: P V of step s - 1 done
            tc::tc_fence_after();
#pragma unroll
            for (int jo = 0; jo < D / 16 + 1; ++jo) {  // O columns and the row-sum columns behind them
              uint32_t o[16];
              tc::tmem_ld16(tO + jo * 16, o);
              tc::tmem_ld_wait();
#pragma unroll
              for (int u = 0; u < 16; ++u) o[u] = __float_as_uint(__uint_as_float(o[u]) * f);
              tc::tmem_st16(tO + jo * 16, o);
            }
          }
        }
        const float nmsc = -msc;
        uint32_t pk[2][16];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float x0 = fmaf(__uint_as_float(v[q][2 * u]), scale_log2e, nmsc);
            const float x1 = fmaf(__uint_as_float(v[q][2 * u + 1]), scale_log2e, nmsc);
            const float p0 = ((2 * u) & 7) < POLY ? exp2_poly(x0) : ex2_approx(x0);
            const float p1 = ((2 * u + 1) & 7) < POLY ? exp2_poly(x1) : ex2_approx(x1);
            pk[q][u] = tc::pack_bf16_bits(__float_as_uint(p0), __float_as_uint(p1));  // truncation: see header
          }
        }
        if (s >= 2) {  // P_g[b] was read by the P V of step s - 2 (long done in steady state)
          tc::mbar_wait(&pv_done[gb], ((s - 2) >> 1) & 1);
          tc::tc_fence_after();
        }
        tc::tmem_st16(tP + b * 32, pk[0]);       // keys 0..31 of the step -> 16 packed columns
        tc::tmem_st16(tP + b * 32 + 16, pk[1]);  // keys 32..63
        tc::tmem_st_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(&p_ready[gb]);
      }
      // ---- epilogue of the tile: O / l -> bf16 -> the point's own row (the [inverse] gather is fused)
      tc::mbar_wait(&pv_done[g * 2 + ((s - 1) & 1)], ((s - 1) >> 1) & 1);
      tc::tc_fence_after();
      const float lsum = __uint_as_float(tc::tmem_ld1(tO + D));  // sum of the bf16 weights, from the tensor core
      tc::tmem_ld_wait();
      const float inv = 1.f / lsum;
      // training: log2-domain log-sum-exp of the row (scores already scaled), by sorted position, for the backward
      if (lse2 && out_row >= 0) lse2[(size_t)h * lse_stride + q_beg + qi] = msc + log2f(lsum);
      __nv_bfloat16* orow = out_row >= 0 ? out + (size_t)out_row * C + h * D : nullptr;
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
    }
  } else if (warp == 8 || warp == 11) {
    // =========================================================== Q loaders (warp 8: group 0, warp 11: group 1)
    const int g = warp == 8 ? 0 : 1;
    const int ntiles = g == 0 ? (nqb + 1) / 2 : nqb / 2;
    // tile i = 0 was gathered straight into shared memory in the prologue
    tc::cp_async_wait_all();
    tc::fence_proxy_async();
    tc::mbar_arrive(&q_full[g]);
    // later tiles: the rows are fetched into REGISTERS while the previous tile of the group is still being
    // processed, and stored the moment its last Q K^T has released the buffer (a third Q buffer does not fit
    // beside K/V at head dim 48; a gather issued only then would expose ~3k cycles of latency per tile)
    for (int i = 1; i < ntiles; ++i) {
      const int t = 2 * i + g;
      uint4 r[kItems];
#pragma unroll
      for (int u = 0; u < kItems; ++u) {
        const int item = lane + 32 * u;
        const int rr = item / kChunksPerRow, c = item - rr * kChunksPerRow;
        const int qi = t * kQB + rr;
        r[u] = make_uint4(0u, 0u, 0u, 0u);
        if (qi < n_q)
          r[u] = __ldg(reinterpret_cast<const uint4*>(qkv + (size_t)order_row[q_beg + qi] * (3 * C) + h * D + c * 8));
      }
      tc::mbar_wait_sleep(&q_free[g], (i - 1) & 1);
      uint8_t* sQ = smem + S::kOffQ + g * S::kQ;
#pragma unroll
      for (int u = 0; u < kItems; ++u) {
        const int item = lane + 32 * u;
        const int rr = item / kChunksPerRow, c = item - rr * kChunksPerRow;
        *reinterpret_cast<uint4*>(sQ + c * (kQB * 16) + (rr >> 3) * 128 + (rr & 7) * 16) = r[u];
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&q_full[g]);
    }
  } else {
    // =========================================================== MMA issuers (warp 9: group 0, warp 10: group 1;
    // whole warp in uniform control flow, one elected lane per op)
    const int g = warp - 9;
    const int total = (g == 0 ? (nqb + 1) / 2 : nqb / 2) * nst;
    constexpr uint32_t idesc_s = tc::umma_idesc_bf16(kQB, kKS, 0, 0);  // S = Q K^T : M=128, N=64
    constexpr uint32_t idesc_o = tc::umma_idesc_bf16(kQB, D, 0, 1);    // O += P V : M=128, N=D, B MN-major
    constexpr uint32_t idesc_l = tc::umma_idesc_bf16(kQB, 16, 0, 1);   // l += P 1 : N = 16 replicated columns
    const uint32_t sK = tc::smem_u32(smem + S::kOffK), sV = tc::smem_u32(smem + S::kOffV);
    const uint32_t q0 = tc::smem_u32(smem + S::kOffQ + g * S::kQ) >> 4;
    const uint32_t tSg = tmem_base + kSCol + g * 128;
    const uint32_t tPg = tmem_base + kPCol + g * 64;
    const uint32_t tOg = tmem_base + kOCol + g * 64;
    // descriptor bases (only the 14-bit start-address field changes per MMA)
    const uint64_t dq_base = tc::umma_desc_nosw(0, kQB * 16, 128);
    const uint64_t dk_base = tc::umma_desc_nosw(0, KMAX * 16, 128);
    const uint64_t dv_base = tc::umma_desc_nosw(0, 128, KMAX * 16);
    const uint64_t d_ones = tc::umma_desc_nosw(tc::smem_u32(smem + S::kOffOnes), 128, 256);
    int kv_ready = 0;            // gather chunks known to have landed
    int sq = 0, jq = 0, tq = 0;  // step / chunk-in-tile / tile index of the NEXT Q K^T

    auto issue_qk = [&]() {
      const int b = sq & 1;
      if (sq >= 2) {  // S_g[b] of step sq - 2 must be in the softmax warps' registers
        tc::mbar_wait(&s_free[g * 2 + b], ((sq - 2) >> 1) & 1);
        tc::tc_fence_after();
      }
      if (jq == 0) {
        tc::mbar_wait(&q_full[g], tq & 1);
        tc::tc_fence_after();
      }
      const int gc = (jq * kKS) / kKG;
      if (kv_ready <= gc) {
        tc::mbar_wait(&kv_full[gc], 0);
        tc::tc_fence_after();
        kv_ready = gc + 1;
      }
      const uint32_t k0 = (sK + jq * (kKS / 8) * 128) >> 4;
#pragma unroll
      for (int t = 0; t < D / 16; ++t) {
        const uint64_t da = dq_base | (uint64_t)((q0 + 2 * t * kQB) & 0x3fff);
        const uint64_t db = dk_base | (uint64_t)((k0 + 2 * t * KMAX) & 0x3fff);
        tc::umma_bf16_elect(tSg + b * 64, da, db, idesc_s, t ? 1u : 0u);
      }
      tc::umma_commit_elect(&s_full[g * 2 + b]);
      ++sq;
      if (jq == nst - 1) {
        tc::umma_commit_elect(&q_free[g]);  // last read of this Q buffer
        jq = 0;
        ++tq;
      } else {
        ++jq;
      }
    };

    if (total > 0) issue_qk();
    if (total > 1) issue_qk();
    int pj = 0;  // chunk-in-tile index of the next P V
    for (int s = 0; s < total; ++s) {
      const int b = s & 1;
      // Q K^T of step s + 2 first (it only needs S_g[b] of step s to be in registers, which happens early in the
      // softmax's step s), unless it opens a new tile: then this step's P V must not queue behind the wait for the
      // new Q rows
      const bool qk_first = sq < total && jq != 0;
      if (qk_first) issue_qk();
      tc::mbar_wait(&p_ready[g * 2 + b], (s >> 1) & 1);
      tc::tc_fence_after();
      const uint32_t v0 = (sV + pj * (kKS / 8) * 128) >> 4;
#pragma unroll
      for (int t = 0; t < kKS / 16; ++t) {
        const uint64_t dv = dv_base | (uint64_t)((v0 + t * 16) & 0x3fff);
        tc::umma_bf16_ts_elect(tOg, tPg + b * 32 + 8 * t, dv, idesc_o, (pj | t) ? 1u : 0u);
      }
#pragma unroll
      for (int t = 0; t < kKS / 16; ++t)
        tc::umma_bf16_ts_elect(tOg + D, tPg + b * 32 + 8 * t, d_ones, idesc_l, (pj | t) ? 1u : 0u);
      tc::umma_commit_elect(&pv_done[g * 2 + b]);
      pj = pj == nst - 1 ? 0 : pj + 1;
      if (!qk_first && sq < total) issue_qk();
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
static int launch(const void* qkv, const int64_t* order_row, const int32_t* table, int max_patches, int heads,
                  float scale, void* out, float* lse2, int64_t lse_stride, cudaStream_t stream) {
  constexpr int KMAX = 1024;
  using S = Smem<D, KMAX>;
  auto kern = patch_attention_kernel<D, KMAX, SS_ATT_POLY>;
  SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
  // heads fastest: the H CTAs of a patch run together and share the gathered rows' DRAM sectors through L2
  dim3 grid((unsigned)((size_t)heads * max_patches));
  kern<<<grid, kThreads, S::kTotal, stream>>>((const __nv_bfloat16*)qkv, order_row, (const int4*)table, heads,
                                              scale * 1.4426950408889634f, (__nv_bfloat16*)out, lse2, lse_stride);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // namespace att
}  // namespace ss

#ifndef SS_ATT_ENTRY
#define SS_ATT_ENTRY(name) name
#endif

static int patch_attention_entry(const void* qkv_bf16, const int64_t* order_row, const int32_t* table, int max_patches,
                                 int patch_size, int heads, int head_dim, float scale, void* out_bf16, float* lse2,
                                 int64_t lse_stride, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (max_patches < 0 || heads < 1 || patch_size < 1 || patch_size > 1024 || !(scale > 0.f)) return SS_BAD_ARGS;
  if (max_patches == 0) return SS_OK;
  if ((long long)max_patches * heads > 0x7fffffffLL) return SS_BAD_ARGS;
  if (!qkv_bf16 || !order_row || !table || !out_bf16) return SS_BAD_ARGS;
  if (((uintptr_t)qkv_bf16 | (uintptr_t)out_bf16) % 16 != 0) return SS_BAD_ARGS;
  switch (head_dim) {
    case 16: return ss::att::launch<16>(qkv_bf16, order_row, table, max_patches, heads, scale, out_bf16, lse2, lse_stride, stream);
    case 32: return ss::att::launch<32>(qkv_bf16, order_row, table, max_patches, heads, scale, out_bf16, lse2, lse_stride, stream);
    case 48: return ss::att::launch<48>(qkv_bf16, order_row, table, max_patches, heads, scale, out_bf16, lse2, lse_stride, stream);
    default: return SS_BAD_ARGS;
  }
}

extern "C" int SS_ATT_ENTRY(ss_patch_attention)(const void* qkv_bf16, const int64_t* order_row, const int32_t* table,
                                                int max_patches, int patch_size, int heads, int head_dim, float scale,
                                                void* out_bf16, void* stream_) {
  return patch_attention_entry(qkv_bf16, order_row, table, max_patches, patch_size, heads, head_dim, scale, out_bf16,
                               nullptr, 0, stream_);
}

extern "C" int SS_ATT_ENTRY(ss_patch_attention_lse)(const void* qkv_bf16, const int64_t* order_row, const int32_t* table,
                                                    int max_patches, int patch_size, int heads, int head_dim, float scale,
                                                    void* out_bf16, float* lse2, int64_t n, void* stream_) {
  if (!lse2 || n < 0) return SS_BAD_ARGS;
  return patch_attention_entry(qkv_bf16, order_row, table, max_patches, patch_size, heads, head_dim, scale, out_bf16,
                               lse2, n, stream_);
}
