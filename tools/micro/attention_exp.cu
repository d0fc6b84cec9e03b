// EXPERIMENTAL (round 2, not part of the library build): the attention forward restructured on the evidence of the
// clock64 traces in profiles/r2_attention.md.  Status: numerically correct only in its first form (git history:
// "Attention forward restructured"); this last form (pre-scaled keys + max-free bounded path + group ping-pong) still
// fails the ragged / short-sequence cases of tests/test_gpu_float.py and is 2 % SLOWER than the shipped kernel at head
// dim 48 (3-6 % faster at 16 / 32).  Kept as the record of what was tried; build it with tools/micro/att_bench.cu.
//
// Patch-wise serialized attention on the 5th-gen tensor cores (tcgen05 + TMEM), bf16 in / bf16 out.
//
// Replaces (reference): SerializedAttention.forward's `qkv[order]` gather, flash_attn_varlen_qkvpacked_func
// and `feat[inverse]` gather (point_transformer_v3m1_base.py:181-216); patch rule of :114-170 comes in as
// the device patch table (attention_simt.cu: patch_table_kernel).
//
// One CTA per (head, patch), 14 warps.  K and V of the head (<= 1024 tokens) stay resident in shared memory in the
// UMMA no-swizzle core-matrix layout (K: K-major, V: MN-major).  They are gathered through the serialized order
// with 16-byte cp.async by the 8 softmax warps themselves (two 64-key units each, all in flight at once: those
// warps have nothing else to do before the first S tile exists) behind per-unit mbarriers.  When a unit has landed
// its K rows are multiplied IN PLACE by scale * log2(e) (rounded back to bf16: 2^-9 relative per element, the size of
// the input quantisation itself), so that S = Q K'^T is already the base-2 exponent and the softmax warps spend no
// FFMA per score on the scaling.
//
// Single pass, two query tiles of 128 rows in flight (one per softmax group), 64 keys per step, and TWO score /
// weight buffers per group: the tensor-core work of a group for step s + 1 and s + 2 is done (or under way) while its
// softmax warps work on step s, so in steady state neither side waits for the other's round trip (round 1's kernel
// had one S buffer per group and exposed ~1000 clk of mbarrier / TMEM / MMA latency per step).
//   warps 0-7   softmax: group g = w / 4 (tiles g, g + 2, ..), row quarter = w % 4: one thread owns one query row
//               (= one TMEM lane); the 64 scores of a step come out of TMEM into registers ONCE, the row max and the
//               rescale decision (general steps only) are thread-local (no shuffles, no shared memory)
//   warps 8-11  Q loaders (group, row half): the NEXT tile's rows are prefetched into registers and stored the moment
//               the last Q K^T of the current tile has released the (single) Q buffer of the group
//   warps 12,13 MMA issuers of group 0 / 1: S_g[b] = Q_g K_c^T two steps ahead of the softmax, O_g += P_g[b] V_c and
//               l_g += P_g[b] 1 once P_g[b] is written
// TMEM columns: S_g[b] (g*2+b)*64 in [0,256) | P_g[b] 256 + (g*2+b)*32 in [256,384) (bf16 pairs) | O_g, l_g
// 384 + g*64 in [384,512).
// P is CUT to bf16 (no rounding instruction) and the row sum l is accumulated by the tensor core from the same
// bf16 weights (P times a 16x16 tile of ones): O / l is an exactly normalised convex combination of V rows.
//
// Softmax reference.  The weights are 2^(S - m_row) for ANY per-row reference m_row (it cancels in O / l); what m_row
// has to guarantee is that nothing overflows or underflows in fp32 / bf16.  Two paths per step, chosen per warp:
//   * bounded step (the common case): |S| <= |q_row| * max_j |k'_j| (Cauchy-Schwarz; the norms of the scaled keys of
//     every 64-key unit are taken once per CTA, the query norms once per tile).  When that bound is <= 96 for all 32
//     rows of the warp and the row's reference is still 0 (its value at tile start), every exponent lies in [-96, 96]
//     and sums of 1024 weights stay below 2^107: the step needs NO row max, no subtraction, no rescale of O:
//     TMEM -> EX2 -> bf16 pack -> TMEM, 16 columns at a time with the next block's TMEM load in flight.
//   * general step (huge logits, or the ragged last step of a short sequence): online softmax with lazy rescaling
//     over the 64 scores of the step; the running reference only moves when the step max exceeds it by more than
//     2^8, so O is touched by CUDA cores almost only once per tile.
// Both write the same P / arrive on the same barriers, so warps and steps may mix freely.  The kernel is bound by the
// N*K*H exponentials, not by the tensor pipe; a fraction of them is evaluated on the FMA pipe
// (exp2_poly).
#include "../../scenesplat_b200/csrc/tc_common.cuh"
#include "../../scenesplat_b200/csrc/attention_math.cuh"
#include "../../include/scenesplat_b200.h"

#ifndef SS_ATT_POLY
#define SS_ATT_POLY 2
#endif

namespace ss {
namespace att {

#ifdef SS_ATT_TRACE2  // developer instrumentation (tools/micro/att_bench.cu): clock64 stamps of the first CTAs
constexpr int kTraceSlots = 64, kTraceCtas = 2048;
__device__ long long g_att_trace[kTraceCtas * kTraceSlots];
#define ATT_TRACE(slot)                                                                                    \
  do {                                                                                                     \
    if (lane == 0 && blockIdx.x < kTraceCtas) g_att_trace[blockIdx.x * kTraceSlots + (slot)] = clock64(); \
  } while (0)
#else
#define ATT_TRACE(slot) do {} while (0)
#endif

constexpr int kThreads = 448;  // 8 softmax warps + 4 Q loader warps + 2 MMA warps
constexpr int kQB = 128;       // query rows per tile
constexpr int kKS = 64;        // keys per step = keys per prologue gather unit
constexpr int kSCol = 0;       // S_g[b] at kSCol + (g * 2 + b) * 64
constexpr int kPCol = 256;     // P_g[b] at kPCol + (g * 2 + b) * 32
constexpr int kOCol = 384;     // O_g at kOCol + g * 64 (D columns of O, then 16 replicated row-sum columns)
constexpr float kLazy = 8.f;   // log2 units the running max may lag behind
constexpr float kBound = 96.f;  // bounded step: |S| <= kBound for every row of the warp and reference 0: exponents in
                                // [-96, 96], sums of 1024 terms < 2^107

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
  static constexpr int kOffKn = kOffOnes + 512;   // [16] float: max |k_j| of key unit c
  static constexpr int kOffIdx = kOffKn + 64;     // [12 warps][64] int32: gathered row numbers of the warp's unit
  static constexpr int kTotal = kOffIdx + 12 * 64 * 4 + 128;
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
  uint64_t* kv_full = bars;       // [16]   K/V unit c landed (32 lane arrivals of the warp that gathered it)
  uint64_t* q_full = bars + 16;   // [2]    Q tile of group g landed (two row halves: 64 lane arrivals)
  uint64_t* q_free = bars + 18;   // [2]    last Q K^T of the tile done: buffer may be refilled
  uint64_t* s_full = bars + 20;   // [2][2] S_g[b] ready
  uint64_t* s_free = bars + 24;   // [2][2] S_g[b] is in registers (128 rows): a later Q K^T may overwrite it
  uint64_t* p_ready = bars + 28;  // [2][2] P_g[b] written by all 128 rows
  uint64_t* pv_done = bars + 32;  // [2][2] P_g[b] V done: P_g[b] may be rewritten, O_g is quiescent up to that step
  uint32_t* tmem_slot = (uint32_t*)(bars + 36);
  float* kn = reinterpret_cast<float*>(smem + S::kOffKn);

  // warp index through a shuffle: the compiler then KNOWS it is warp-uniform (role branches stay convergent and
  // the MMA warps' descriptors can live in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int h = blockIdx.x % H;
  const int C = H * D;
  if (warp == 0) ATT_TRACE(0);
  const int nst = (kv_len + kKS - 1) / kKS;  // steps (= key units) per query tile
  const int nqb = (n_q + kQB - 1) / kQB;     // query tiles
  constexpr int kChunksPerRow = D / 8;       // 16-byte pieces per row
  constexpr int kItems = 2 * kChunksPerRow;  // pieces per lane per unit of 64 rows

  if (threadIdx.x == 0) {
    for (int c = 0; c < 16; ++c) tc::mbar_init(&kv_full[c], 32);
    for (int g = 0; g < 2; ++g) {
      tc::mbar_init(&q_full[g], 64);
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
  if (warp == 12) tc::tmem_alloc<512>(tmem_slot);

  // ---- gathers of 64 rows.  The row numbers of a unit are fetched with ONE coalesced load per lane (2 each) and
  // staged in shared memory, so the 16-byte row pieces are all requested back to back (a per-piece index load
  // serialises dependent DRAM round trips).  Lanes then walk the pieces of a row first (item = row * kChunksPerRow + c),
  // so one warp instruction touches 32 / kChunksPerRow rows.
  // Element (row j, piece c) -> c * (ROWS*16) + (j/8)*128 + (j%8)*16 (UMMA no-swizzle core matrices; K and Q are
  // K-major operands, V is an MN-major operand, same byte layout).
  int32_t* my_idx = reinterpret_cast<int32_t*>(smem + S::kOffIdx) + (warp < 12 ? warp : 0) * 64;
  auto stage_idx = [&](int first, int count) {  // rows [first, first + 64) of the sorted order; -1 past `count`
    __syncwarp();  // (the previous unit's indices have been read by every lane)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = lane + 32 * u;
      my_idx[r] = r < count ? (int32_t)order_row[first + r] : -1;
    }
    __syncwarp();
  };
  auto gather_kv = [&](int unit) {
    const uint32_t sK = tc::smem_u32(smem + S::kOffK), sV = tc::smem_u32(smem + S::kOffV);
    stage_idx(kv_beg + unit * kKS, kv_len - unit * kKS);
#pragma unroll
    for (int u = 0; u < kItems; ++u) {
      const int item = lane + 32 * u;
      const int r = item / kChunksPerRow, c = item - r * kChunksPerRow;
      const int j = unit * kKS + r;
      const int32_t row = my_idx[r];
      const __nv_bfloat16* src = qkv + (size_t)(row < 0 ? 0 : row) * (3 * C) + h * D + c * 8;
      const uint32_t off = (uint32_t)(c * (KMAX * 16) + (j >> 3) * 128 + (j & 7) * 16);
      tc::cp_async16(sK + off, src + C, row < 0 ? 0u : 16u);
      tc::cp_async16(sV + off, src + 2 * C, row < 0 ? 0u : 16u);
    }
  };
  auto gather_q_half = [&](int t, int hh) {
    const uint32_t sQ = tc::smem_u32(smem + S::kOffQ + (t & 1) * S::kQ);
    stage_idx(q_beg + t * kQB + hh * 64, n_q - t * kQB - hh * 64);
#pragma unroll
    for (int u = 0; u < kItems; ++u) {
      const int item = lane + 32 * u;
      const int rr = item / kChunksPerRow, c = item - rr * kChunksPerRow;
      const int r = hh * 64 + rr;
      const int32_t row = my_idx[rr];
      const __nv_bfloat16* src = qkv + (size_t)(row < 0 ? 0 : row) * (3 * C) + h * D + c * 8;
      tc::cp_async16(sQ + (uint32_t)(c * (kQB * 16) + (r >> 3) * 128 + (r & 7) * 16), src, row < 0 ? 0u : 16u);
    }
  };
  // Prologue: the 8 softmax warps have nothing to do until the first S tile exists, so each of them gathers two
  // K/V units (w first, then w + 8: the early units land first; all 16 units and both Q tiles are in flight at once);
  // issued before the block-wide barrier so the gathers overlap the TMEM allocation.
  if (warp < 8) {
    if (warp < nst) gather_kv(warp);
    tc::cp_async_commit();
    if (warp + 8 < nst) gather_kv(warp + 8);
    tc::cp_async_commit();
  }
  if (warp >= 8 && warp < 12) {
    const int g = (warp - 8) >> 1, hh = (warp - 8) & 1;
    if (g < nqb) gather_q_half(g, hh);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) ATT_TRACE(1);

  if (warp < 8) {
    // =========================================================== softmax warps
    // publish the units this warp gathered: K rows scaled in place by scale * log2e, the largest scaled key norm of
    // the unit recorded for the bounded-step test
    auto publish = [&](int unit) {
      float k2 = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = unit * kKS + lane + 32 * u;  // rows past kv_len were zero-filled
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < kChunksPerRow; ++c) {
          uint4* ptr = reinterpret_cast<uint4*>(smem + S::kOffK + c * (KMAX * 16) + (j >> 3) * 128 + (j & 7) * 16);
          uint4 w = *ptr;
          __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float2 f = __bfloat1622float2(hp[q]);
            hp[q] = __floats2bfloat162_rn(f.x * scale_log2e, f.y * scale_log2e);
            f = __bfloat1622float2(hp[q]);
            a = fmaf(f.x, f.x, fmaf(f.y, f.y, a));
          }
          *ptr = w;
        }
        k2 = fmaxf(k2, a);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) k2 = fmaxf(k2, __shfl_xor_sync(0xffffffffu, k2, o));
      if (lane == 0) kn[unit] = sqrtf(k2);
      tc::fence_proxy_async();
      tc::mbar_arrive(&kv_full[unit]);  // (release: kn[unit] is visible to whoever saw an S tile of this unit)
    };
    if (warp < nst) {
      tc::cp_async_wait<1>();
      __syncwarp();  // every lane's pieces of the unit have landed (a lane rewrites rows other lanes gathered)
      ATT_TRACE(2 + warp);
      publish(warp);
    }
    if (warp + 8 < nst) {
      tc::cp_async_wait<0>();
      __syncwarp();
      publish(warp + 8);
    }
    const int g = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;  // row inside the query tile == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t tS = t_lane + kSCol + g * 128;  // + b * 64
    const uint32_t tP = t_lane + kPCol + g * 64;   // + b * 32
    const uint32_t tO = t_lane + kOCol + g * 64;
    const int ntiles = g == 0 ? (nqb + 1) / 2 : nqb / 2;
    // bounded steps need the query norms, read from the Q tile in shared memory at the start of a tile; with <= 2 steps
    // per tile the Q buffer may already belong to the next tile by then: such short sequences take the general path
    const bool use_bound = nst > 2;
    // Group ping-pong.  All softmax warps of an SM sub-partition drift into lock step (nothing breaks the symmetry), and
    // then the exponential bursts of the two groups collide on the MUFU / FMA pipes while their TMEM / barrier parts
    // leave the pipes idle together (measured: 1245 clk per step pair, of which ~770 is pipe time; profiles/).  The
    // groups therefore take turns in the arithmetic part of a step: named barrier 1 = "group 0 may compute", 2 = "group
    // 1 may compute" (128 waiting + 128 arriving threads each); one group's TMEM loads / stores / mbarrier traffic runs
    // under the other's exponentials.  Step k of group 0 follows step k - 1 of group 1 and precedes its step k.
    const int n_other = (g == 0 ? nqb / 2 : (nqb + 1) / 2) * nst;  // steps of the other group
    const bool pp = nqb > 1;
    if (pp && g == 1) asm volatile("bar.arrive 1, 256;" ::: "memory");  // group 0 takes the first turn
    auto turn_begin = [&](int k) {
      if (!pp) return;
      if (g == 0) {
        if (k <= n_other) asm volatile("bar.sync 1, 256;" ::: "memory");
      } else {
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
    };
    auto turn_end = [&](int k) {
      if (!pp) return;
      if (g == 0) {
        if (k < n_other) asm volatile("bar.arrive 2, 256;" ::: "memory");
      } else {
        if (k + 1 < n_other) asm volatile("bar.arrive 1, 256;" ::: "memory");
      }
    };
    int s = 0;  // step counter of this group: buffer b = s & 1, use number (phase) = (s >> 1) & 1
    // barrier probes are issued one step early (a try_wait costs ~100 clk of latency even when the phase is complete)
    bool s_ok = ntiles > 0 && __all_sync(0xffffffffu, tc::mbar_try_wait(&s_full[g * 2], 0));
    for (int i = 0; i < ntiles; ++i) {
      float msc = 0.f;  // reference of the row (log2 domain, already scaled); the first general step replaces it
      const int qi = (2 * i + g) * kQB + row;
      const int64_t out_row = qi < n_q ? order_row[q_beg + qi] : -1;  // fetched now, needed by the tile's epilogue
      float qn = INFINITY;  // |q_row|
      if (use_bound) {
        tc::mbar_wait(&q_full[g], i & 1);
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < kChunksPerRow; ++c) {
          const uint4 w = *reinterpret_cast<const uint4*>(smem + S::kOffQ + g * S::kQ + c * (kQB * 16) + (row >> 3) * 128 +
                                                          (row & 7) * 16);
          const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __bfloat1622float2(hp[q]);
            a = fmaf(f.x, f.x, fmaf(f.y, f.y, a));
          }
        }
        qn = sqrtf(a);
      }
      if (warp == 0 && i == 0) ATT_TRACE(12);
      for (int j = 0; j < nst; ++j, ++s) {
        const int b = s & 1;
        const int gb = g * 2 + b;
        if (!s_ok) tc::mbar_wait(&s_full[gb], (s >> 1) & 1);
        tc::tc_fence_after();
        if (warp == 0 && s == 0) ATT_TRACE(11);
        if (warp == 0 && i == 1) ATT_TRACE(20 + j);
        // probe now what the end of the step needs: P_g[b] was read by the P V of step s - 2 (long done in steady state)
        const bool p_ok = s < 2 || __all_sync(0xffffffffu, tc::mbar_try_wait(&pv_done[gb], ((s - 2) >> 1) & 1));
        const int valid = kv_len - j * kKS;  // valid keys of the step
        const bool bounded = valid >= kKS && __all_sync(0xffffffffu, qn * kn[j] <= kBound && msc == 0.f);
        if (bounded) {
          // ---- bounded step, reference 0: the weight is 2^S.  Scores -> registers, then (in this group's turn) the
          // 64 exponentials, then the weights -> TMEM
          uint32_t v[4][16], pk[4][8];
#pragma unroll
          for (int q = 0; q < 4; ++q) tc::tmem_ld16(tS + b * 64 + 16 * q, v[q]);
          tc::tmem_ld_wait();
          tc::tc_fence_before();
          tc::mbar_arrive(&s_free[gb]);  // the Q K^T of step s + 2 may overwrite S_g[b]
          s_ok = __all_sync(0xffffffffu, tc::mbar_try_wait(&s_full[gb ^ 1], ((s + 1) >> 1) & 1));  // the next step's S
          turn_begin(s);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float x0 = __uint_as_float(v[q][2 * u]), x1 = __uint_as_float(v[q][2 * u + 1]);
              const float p0 = ((2 * u) & 7) < POLY ? exp2_poly_bounded(x0) : ex2_approx(x0);
              const float p1 = ((2 * u + 1) & 7) < POLY ? exp2_poly_bounded(x1) : ex2_approx(x1);
              pk[q][u] = tc::pack_bf16_bits(__float_as_uint(p0), __float_as_uint(p1));  // truncation: see header
            }
          }
          turn_end(s);
          if (!p_ok) {
            tc::mbar_wait(&pv_done[gb], ((s - 2) >> 1) & 1);
            tc::tc_fence_after();
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) tc::tmem_st8(tP + b * 32 + 8 * q, pk[q]);
        } else {
          // ---- general step: online softmax with lazy rescaling over the 64 scores
          uint32_t v[2][32];
          tc::tmem_ld32(tS + b * 64, v[0]);
          tc::tmem_ld32(tS + b * 64 + 32, v[1]);
          tc::tmem_ld_wait();
          tc::tc_fence_before();
          tc::mbar_arrive(&s_free[gb]);
          s_ok = __all_sync(0xffffffffu, tc::mbar_try_wait(&s_full[gb ^ 1], ((s + 1) >> 1) & 1));
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
          const float nm = fmaxf(mx[0], mx[1]);
          // first step of a tile: nothing accumulated yet, the reference simply becomes the step max (it may go DOWN)
          const bool need = j == 0 || nm > msc + kLazy;
          if (__any_sync(0xffffffffu, need)) {
            const float newm = need ? nm : msc;
            const float f = ex2_approx(msc - newm);
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
          uint32_t pk[2][16];
          turn_begin(s);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const float x0 = __uint_as_float(v[q][2 * u]) - msc;
              const float x1 = __uint_as_float(v[q][2 * u + 1]) - msc;
              const float p0 = ((2 * u) & 7) < POLY ? exp2_poly(x0) : ex2_approx(x0);
              const float p1 = ((2 * u + 1) & 7) < POLY ? exp2_poly(x1) : ex2_approx(x1);
              pk[q][u] = tc::pack_bf16_bits(__float_as_uint(p0), __float_as_uint(p1));  // truncation: see header
            }
          }
          turn_end(s);
          if (!p_ok) {
            tc::mbar_wait(&pv_done[gb], ((s - 2) >> 1) & 1);
            tc::tc_fence_after();
          }
          tc::tmem_st16(tP + b * 32, pk[0]);       // keys 0..31 of the step -> 16 packed columns
          tc::tmem_st16(tP + b * 32 + 16, pk[1]);  // keys 32..63
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(&p_ready[gb]);
        if (warp == 0 && i == 1) ATT_TRACE(36 + j);
      }
      if (warp == 0 && i == 0) ATT_TRACE(13);
      // ---- epilogue of the tile: O / l -> bf16 -> the point's own row (the [inverse] gather is fused)
      tc::mbar_wait(&pv_done[g * 2 + ((s - 1) & 1)], ((s - 1) >> 1) & 1);
      tc::tc_fence_after();
      uint32_t o[D / 16][16];
      const uint32_t lbits = tc::tmem_ld1(tO + D);  // sum of the bf16 weights, from the tensor core
#pragma unroll
      for (int jo = 0; jo < D / 16; ++jo) tc::tmem_ld16(tO + jo * 16, o[jo]);
      tc::tmem_ld_wait();
      const float lsum = __uint_as_float(lbits);
      const float inv = 1.f / lsum;
      // training: log2-domain log-sum-exp of the row (scores already scaled), by sorted position, for the backward
      if (lse2 && out_row >= 0) lse2[(size_t)h * lse_stride + q_beg + qi] = msc + log2f(lsum);
      if (out_row >= 0) {
        __nv_bfloat16* orow = out + (size_t)out_row * C + h * D;
#pragma unroll
        for (int jo = 0; jo < D / 16; ++jo) {
          uint4 o0, o1;
          o0.x = tc::pack_bf16(__uint_as_float(o[jo][0]) * inv, __uint_as_float(o[jo][1]) * inv);
          o0.y = tc::pack_bf16(__uint_as_float(o[jo][2]) * inv, __uint_as_float(o[jo][3]) * inv);
          o0.z = tc::pack_bf16(__uint_as_float(o[jo][4]) * inv, __uint_as_float(o[jo][5]) * inv);
          o0.w = tc::pack_bf16(__uint_as_float(o[jo][6]) * inv, __uint_as_float(o[jo][7]) * inv);
          o1.x = tc::pack_bf16(__uint_as_float(o[jo][8]) * inv, __uint_as_float(o[jo][9]) * inv);
          o1.y = tc::pack_bf16(__uint_as_float(o[jo][10]) * inv, __uint_as_float(o[jo][11]) * inv);
          o1.z = tc::pack_bf16(__uint_as_float(o[jo][12]) * inv, __uint_as_float(o[jo][13]) * inv);
          o1.w = tc::pack_bf16(__uint_as_float(o[jo][14]) * inv, __uint_as_float(o[jo][15]) * inv);
          uint4* dst = reinterpret_cast<uint4*>(orow + jo * 16);
          dst[0] = o0;
          dst[1] = o1;
        }
      }
      tc::tc_fence_before();  // ordered before the next tile's first P V by the next p_ready arrival
      if (warp == 0 && i == 0) ATT_TRACE(14);
    }
    if (warp == 0) ATT_TRACE(15);
    if (warp == 4) ATT_TRACE(16);
  } else if (warp < 12) {
    // =========================================================== Q loaders: (group, row half)
    const int g = (warp - 8) >> 1, hh = (warp - 8) & 1;
    const int ntiles = g == 0 ? (nqb + 1) / 2 : nqb / 2;
    // tile i = 0 was gathered straight into shared memory in the prologue
    if (ntiles > 0) {
      tc::cp_async_wait_all();
      tc::fence_proxy_async();
      tc::mbar_arrive(&q_full[g]);
    }
    // later tiles: the rows are fetched into REGISTERS while the previous tile of the group is still being
    // processed, and stored the moment its last Q K^T has released the buffer (a third Q buffer does not fit
    // beside K/V at head dim 48; a gather issued only then would expose ~3k cycles of latency per tile)
    for (int i = 1; i < ntiles; ++i) {
      const int t = 2 * i + g;
      uint4 r[kItems];
#pragma unroll
      for (int u = 0; u < kItems; ++u) {
        const int item = lane + 32 * u;
        const int rr = hh * 64 + item / kChunksPerRow, c = item % kChunksPerRow;
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
        const int rr = hh * 64 + item / kChunksPerRow, c = item % kChunksPerRow;
        *reinterpret_cast<uint4*>(sQ + c * (kQB * 16) + (rr >> 3) * 128 + (rr & 7) * 16) = r[u];
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&q_full[g]);
    }
  } else {
    // =========================================================== MMA issuers (warp 12: group 0, warp 13: group 1;
    // whole warp in uniform control flow, one elected lane per op)
    const int g = warp - 12;
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
    int kv_ready = 0;            // key units known to have landed
    int sq = 0, jq = 0, tq = 0;  // step / unit-in-tile / tile index of the NEXT Q K^T

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
      if (kv_ready <= jq) {
        tc::mbar_wait(&kv_full[jq], 0);
        tc::tc_fence_after();
        kv_ready = jq + 1;
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
    if (warp == 12) ATT_TRACE(10);
    if (total > 1) issue_qk();
    int pj = 0;  // unit-in-tile index of the next P V
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
  if (warp == 12) ATT_TRACE(17);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) ATT_TRACE(18);
  if (warp == 12) {
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

#ifdef SS_ATT_TRACE2
extern "C" void ss_att_trace_read(long long* host, int* slots, int* ctas) {
  *slots = ss::att::kTraceSlots;
  *ctas = ss::att::kTraceCtas;
  if (host) cudaMemcpyFromSymbol(host, ss::att::g_att_trace, sizeof(long long) * ss::att::kTraceSlots * ss::att::kTraceCtas);
}
#endif

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
