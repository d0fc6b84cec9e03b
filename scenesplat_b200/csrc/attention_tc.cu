// Patch-wise serialized attention on the 5th-gen tensor cores (tcgen05 + TMEM), bf16 in / bf16 out.
//
// Replaces (reference): SerializedAttention.forward's `qkv[order]` gather, flash_attn_varlen_qkvpacked_func
// and `feat[inverse]` gather (point_transformer_v3m1_base.py:181-216); patch rule of :114-170 comes in as
// the device patch table (attention_simt.cu: patch_table_kernel).
//
// PERSISTENT CTAs (one per SM, 22 warps), each walking the (head, patch) work items blockIdx.x, + gridDim.x, ...  K and V
// of the item (<= 1024 tokens) stay resident in shared memory in the UMMA no-swizzle core-matrix layout (K: K-major,
// V: MN-major), gathered through the serialized order with 16-byte cp.async by four loader warps that run AHEAD of the
// item being computed: chunk c (128 keys) of the next item is gathered as soon as both softmax groups have retired
// chunk c in their last query tile (kv_free, committed by the MMA issuers), its arrival published by the copies
// themselves (cp.async.mbarrier.arrive.noinc).  The K/V gather of an item (196 KB of scattered 96-byte row slices at
// d = 48: ~20 k cycles of one SM's miss path, 15 % of an item when it sat in front of every CTA) hides under the last
// tile pair of the item before it; barrier / TMEM setup is paid once per SM.  Round 2: dec0 2.18 -> 1.84 ms.
//
// Single pass, online softmax, two query tiles of 128 rows in flight (one per softmax group), 128 keys per step:
//   warps 0-15  softmax: group g = w / 8 (tiles g, g + 2, ..), column half = (w / 4) % 2 (64 of the step's keys),
//               row quarter = w % 4 (one thread = one query row = one TMEM lane).  Four softmax warps per SM
//               sub-partition keep the MUFU pipe fed while their neighbours sit in the TMEM / mbarrier part of a
//               step.  The row max needs both halves: the two warps of a row exchange their local maxima through
//               shared memory (double-buffered by step parity) behind a 64-thread named barrier and then take the
//               same lazy-rescale decision.  Scores are read from TMEM twice (max sweep, exponential sweep) 32
//               columns at a time, so a thread holds 32 scores and the kernel fits the 80 registers / thread that
//               22 warps leave.  The two groups take turns in the exponential sweep (named barriers 9 / 10).
//   warps 16-19 loaders: K/V chunks lw and lw + 4 of every item, and the Q tiles of (group lw / 2, row half lw % 2):
//               the NEXT tile's rows (or the next item's first tile) are prefetched into registers and stored the
//               moment the last Q K^T of the current tile has released the (single) Q buffer of the group
//   warps 20,21 MMA issuers of group 0 / 1: S_g = Q_g K_c^T as soon as S_g of the previous step has been read into
//               registers (s_free), O_g += P_g V_c and l_g += P_g 1 once P_g is written (p_ready)
// Barrier parities run over the CTA's lifetime (steps / tiles / loads per chunk so far), never per item: ragged items
// (short last patch, fewer chunks or tiles, a group without tiles) only change how often a barrier is used.
// TMEM columns: S_0 [0,128) S_1 [128,256) | P_0 [256,320) P_1 [320,384) (bf16 pairs) | O_0,l_0 [384,448) O_1,l_1
// [448,512).  S and P are separate so the next Q K^T runs under the exponentials of the current step.
// P is CUT to bf16 (no rounding instruction) and the row sum l is accumulated by the tensor core from the same
// bf16 weights (P times a 16x16 tile of ones): O / l is an exactly normalised convex combination of V rows.
// Online softmax with lazy rescaling: the running reference max only moves when the chunk max exceeds it by more
// than 2^8 (any shift cancels in O / l; bf16 / fp32 have the exponent range), so O is touched by CUDA cores almost
// only on the first chunk(s) of a tile.  The kernel is bound by the N*K*H exponentials, not by the tensor pipe
// (see DESIGN.md); a fraction of them is evaluated on the FMA pipe (exp2_poly).
//
// Measured and NOT adopted (profiles/r2_attention.md): 64-key steps with double-buffered S / P, a max-free softmax under
// a Cauchy-Schwarz bound, pre-scaled keys, deeper TMEM prefetch, L2 prefetch of the next CTA's rows (tools/micro/
// attention_exp.cu), and tile epilogues moved to the loader warps (O and l to registers behind an o_full / o_free
// hand-over: 1.84 -> 2.01 ms).
#include "tc_common.cuh"
#include "attention_math.cuh"
#include "../../include/scenesplat_b200.h"

#ifndef SS_ATT_POLY
#define SS_ATT_POLY 2
#endif
#ifndef SS_ATT_PP
#define SS_ATT_PP 1
#endif

namespace ss {

constexpr int kQB = 128;          // query rows per tile
constexpr int kKC = 128;          // keys per chunk
constexpr int kPCol = 256;        // first P column (bf16 pairs: 64 columns per tile)
constexpr int kPStride = 64;
constexpr int kOCol = 384;        // first O column
constexpr int kOStride = 64;      // columns reserved per O tile (D of O, then the row-sum columns)
constexpr float kLazy = 8.f;      // log2 units the running max may lag behind

template <int D, int KMAX>
struct AttSmem {
  static constexpr int kK = KMAX * D * 2;
  static constexpr int kV = KMAX * D * 2;
  static constexpr int kQ = kQB * D * 2;  // per buffer (one per softmax group)
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kK;
  static constexpr int kOffQ = kK + kV;
  static constexpr int kOffBar = kOffQ + 2 * kQ;
  static constexpr int kOffOnes = kOffBar + 512;  // 16 keys x 16 dims of bf16 1.0
  static constexpr int kTotal = kOffOnes + 512 + 128;
};

constexpr int kAtt16Threads = 704;  // 16 softmax warps + 4 Q loader warps + 2 MMA warps

template <int D, int KMAX>
struct Att16Smem {
  using B = AttSmem<D, KMAX>;
  static constexpr int kOffXmax = B::kOffOnes + 512;  // [2 parities][2 groups][2 halves][128 rows] floats
  static constexpr int kTotal = kOffXmax + 4096 + 128;
};

template <int D, int KMAX, int POLY>
__global__ void __launch_bounds__(kAtt16Threads, 1)
patch_attention_tc16_kernel(const __nv_bfloat16* __restrict__ qkv, const int64_t* __restrict__ order_row,
                            const int4* __restrict__ table, int H, int n_items, float scale_log2e,
                            __nv_bfloat16* __restrict__ out, int pingpong, float* __restrict__ lse2, int64_t lse_stride) {
  using S = AttSmem<D, KMAX>;
  using S16 = Att16Smem<D, KMAX>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  uint64_t* bars = (uint64_t*)(smem + S::kOffBar);
  uint64_t* kv_full = bars;       // [8]  32 lane arrivals (one loader warp per chunk, fired by its copies)
  uint64_t* kv_free = bars + 8;   // [8]  2 commits: the last use of the chunk by either group in this item
  uint64_t* q_full = bars + 16;   // [2]  64 lane arrivals (two row halves)
  uint64_t* q_free = bars + 18;   // [2]
  uint64_t* s_full = bars + 20;   // [2]
  uint64_t* s_free = bars + 22;   // [2]  256 arrivals
  uint64_t* p_ready = bars + 24;  // [2]  256 arrivals
  uint64_t* pv_done = bars + 26;  // [2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 28);
  float* xmax = (float*)(smem + S16::kOffXmax);


  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int C = H * D;
  constexpr int kChunksPerRow = D / 8;
  constexpr int kItemsHalf = 2 * kChunksPerRow;  // 16-byte pieces per lane per 64 rows

  if (threadIdx.x == 0) {
    for (int c = 0; c < 8; ++c) {
      tc::mbar_init(&kv_full[c], 32);
      tc::mbar_init(&kv_free[c], 2);
    }
    for (int g = 0; g < 2; ++g) {
      tc::mbar_init(&q_full[g], 64);
      tc::mbar_init(&q_free[g], 1);
      tc::mbar_init(&s_full[g], 1);
      tc::mbar_init(&s_free[g], 256);
      tc::mbar_init(&p_ready[g], 256);
      tc::mbar_init(&pv_done[g], 1);
    }
    tc::mbar_fence_init();
  }
  if (threadIdx.x >= 512 && threadIdx.x < 640) {
    reinterpret_cast<uint32_t*>(smem + S::kOffOnes)[threadIdx.x - 512] = 0x3f803f80u;
    tc::fence_proxy_async();
  }
  if (warp == 20) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 16) {
    // =========================================================== softmax warps
    const int g = warp >> 3, half = (warp >> 2) & 1, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t tS = t_lane + g * kKC + half * 64;
    const uint32_t tP = t_lane + kPCol + g * kPStride + half * 32;
    const uint32_t tO = t_lane + kOCol + g * kOStride;
    const int pair_bar = 1 + g * 4 + quarter;  // named barrier of the two warps sharing these rows
    int s = 0;                                 // steps of this group since the CTA started (barrier parities)
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int4 e = table[w / H];
      const int h = w % H;
      const int q_beg = e.x, n_q = e.y - e.x, kv_len = e.w - e.z;
      if (n_q <= 0) continue;  // unused table entry (all roles skip it)
      const int nch = (kv_len + kKC - 1) / kKC;
      const int nqb = (n_q + kQB - 1) / kQB;
      const int ntiles = g == 0 ? (nqb + 1) / 2 : nqb / 2;
      // ping-pong (named barriers 9 / 10, 256 waiting + 256 arriving threads): the groups take turns in the
      // exponential sweep, so one group's MUFU burst overlaps the other group's TMEM / handshake part.  Group 0 waits
      // (barrier 9) before its steps 0 .. T1 - 1 (T1 = group 1's steps in this item): step 0 for group 1's entry into
      // the item, step k for the end of group 1's step k - 1; group 1 waits (barrier 10) for the end of group 0's step k.
      // Every arrival is consumed before its sender can arrive on the same barrier again, ALSO across items: with an
      // odd number of query tiles group 1 leaves the item first, and an arrival after its last step (nobody needs it)
      // followed by its entry arrival for the next item would complete barrier 9 without group 0 (a deadlock that only
      // a persistent CTA can reach; found with 13 batch elements x 5 tiles at d = 32).
      const int total_g1 = (nqb / 2) * nch;
      const bool pp = pingpong && nqb > 1;
      if (pp && g == 1) asm volatile("bar.arrive 9, 512;" ::: "memory");  // group 0 takes the first turn
      int si = 0;  // step within the item
      for (int i = 0; i < ntiles; ++i) {
        float msc = -INFINITY;
        const int qi = (2 * i + g) * kQB + row;
        const int64_t out_row = qi < n_q ? order_row[q_beg + qi] : -1;
        for (int j = 0; j < nch; ++j, ++s, ++si) {
          const int valid = kv_len - j * kKC - half * 64;  // valid keys among this warp's 64 columns (may be <= 0)
          tc::mbar_wait(&s_full[g], s & 1);
          tc::tc_fence_after();
          // ---- sweep 1: local row max over the warp's 64 columns
          float mx = -INFINITY;
#pragma unroll
          for (int hq = 0; hq < 2; ++hq) {
            uint32_t v[32];
            tc::tmem_ld32(tS + 32 * hq, v);
            tc::tmem_ld_wait();
            if (valid - 32 * hq < 32) {
#pragma unroll
              for (int u = 0; u < 32; ++u)
                if (32 * hq + u >= valid) v[u] = 0xff800000u;
            }
            float m0 = fmax3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
            float m1 = fmax3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
#pragma unroll
            for (int u = 6; u < 30; u += 4) {
              m0 = fmax3(m0, __uint_as_float(v[u]), __uint_as_float(v[u + 1]));
              m1 = fmax3(m1, __uint_as_float(v[u + 2]), __uint_as_float(v[u + 3]));
            }
            mx = fmaxf(mx, fmax3(m0, m1, fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]))));
          }
          float* xm = xmax + ((s & 1) * 4 + g * 2) * 128;
          xm[half * 128 + row] = mx;
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          mx = fmaxf(mx, xm[(half ^ 1) * 128 + row]);
          const float nm = mx * scale_log2e;  // scale > 0
          const bool need = nm > msc + kLazy;
          if (__any_sync(0xffffffffu, need)) {  // identical in both warps of the pair
            const float newm = need ? nm : msc;
            const float f = ex2_approx(msc - newm);
            msc = newm;
            if (j > 0 && half == 0) {
              tc::mbar_wait(&pv_done[g], (s - 1) & 1);
              tc::tc_fence_after();
#pragma unroll
              for (int jo = 0; jo < D / 16 + 1; ++jo) {
                uint32_t o[16];
                tc::tmem_ld16(tO + jo * 16, o);
                tc::tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 16; ++u) o[u] = __float_as_uint(__uint_as_float(o[u]) * f);
                tc::tmem_st16(tO + jo * 16, o);
              }
            }
          }
          if (j > 0) {  // P_g is read by the previous P V of the group until pv_done
            tc::mbar_wait(&pv_done[g], (s - 1) & 1);
            tc::tc_fence_after();
          }
          // ---- sweep 2: exponentials, 32 columns at a time
          if (pp) {
            if (g == 0) {
              if (si < total_g1) asm volatile("bar.sync 9, 512;" ::: "memory");
            } else {
              asm volatile("bar.sync 10, 512;" ::: "memory");
            }
          }
          const float nmsc = -msc;
#pragma unroll
          for (int hq = 0; hq < 2; ++hq) {
            uint32_t v[32];
            tc::tmem_ld32(tS + 32 * hq, v);
            tc::tmem_ld_wait();
            if (hq == 1) {  // S_g is in registers for the last time: the next Q K^T may overwrite it
              tc::tc_fence_before();
              tc::mbar_arrive(&s_free[g]);
            }
            if (valid - 32 * hq < 32) {
#pragma unroll
              for (int u = 0; u < 32; ++u)
                if (32 * hq + u >= valid) v[u] = 0xff800000u;
            }
            uint32_t pk[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const float x0 = fmaf(__uint_as_float(v[2 * u]), scale_log2e, nmsc);
              const float x1 = fmaf(__uint_as_float(v[2 * u + 1]), scale_log2e, nmsc);
              const float p0 = ((2 * u) & 7) < POLY ? exp2_poly(x0) : ex2_approx(x0);
              const float p1 = ((2 * u + 1) & 7) < POLY ? exp2_poly(x1) : ex2_approx(x1);
              pk[u] = tc::pack_bf16_bits(__float_as_uint(p0), __float_as_uint(p1));
            }
            tc::tmem_st16(tP + 16 * hq, pk);
          }
          if (pp) {
            if (g == 0) {
              if (si < total_g1) asm volatile("bar.arrive 10, 512;" ::: "memory");
            } else {
              if (si + 1 < total_g1) asm volatile("bar.arrive 9, 512;" ::: "memory");
            }
          }
          tc::tmem_st_wait();
          tc::tc_fence_before();
          tc::mbar_arrive(&p_ready[g]);
        }
        // ---- epilogue: this warp writes the column half [half * D/2, (half + 1) * D/2) of its rows
        tc::mbar_wait(&pv_done[g], (s - 1) & 1);
        tc::tc_fence_after();
        const float lsum = __uint_as_float(tc::tmem_ld1(tO + D));
        tc::tmem_ld_wait();
        const float inv = 1.f / lsum;
        if (lse2 && half == 0 && out_row >= 0) lse2[(size_t)h * lse_stride + q_beg + qi] = msc + log2f(lsum);
        __nv_bfloat16* orow = out_row >= 0 ? out + (size_t)out_row * C + h * D + half * (D / 2) : nullptr;
#pragma unroll
        for (int jo = 0; jo < D / 16; ++jo) {
          uint32_t o[8];
          tc::tmem_ld8(tO + half * (D / 2) + jo * 8, o);
          tc::tmem_ld_wait();
          if (orow) {
            uint4 o0;
            o0.x = tc::pack_bf16(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
            o0.y = tc::pack_bf16(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
            o0.z = tc::pack_bf16(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
            o0.w = tc::pack_bf16(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
            *reinterpret_cast<uint4*>(orow + jo * 8) = o0;
          }
        }
        tc::tc_fence_before();
      }
    }
  } else if (warp < 20) {
    // =========================================================== loaders: K/V chunks lw, lw + 4 and Q of (group, row half)
    // The loaders run AHEAD of the item being computed: the K/V chunk c of the next (head, patch) is gathered as soon as
    // both groups have retired chunk c in their last query tile (kv_free), so the gather latency of an item hides
    // under the last tile pair of the item before it.
    const int lw = warp - 16, g = lw >> 1, hh = lw & 1;
    uint32_t ldmask = 0;   // bit c: parity of the number of loads of chunk c so far
    uint32_t ever = 0;     // bit c: chunk c has been loaded before
    int qn = 0;            // Q tiles stored so far for this group
    uint4 r[kItemsHalf];
    auto load_q = [&](int q_beg, int n_q, int h, int t) {
#pragma unroll
      for (int u = 0; u < kItemsHalf; ++u) {
        const int item = lane + 32 * u;
        const int rr = hh * 64 + item / kChunksPerRow, c = item % kChunksPerRow;
        const int qi = t * kQB + rr;
        r[u] = make_uint4(0u, 0u, 0u, 0u);
        if (qi < n_q)
          r[u] = __ldg(reinterpret_cast<const uint4*>(qkv + (size_t)order_row[q_beg + qi] * (3 * C) + h * D + c * 8));
      }
    };
    auto store_q = [&]() {
      if (qn > 0) tc::mbar_wait_sleep(&q_free[g], (qn - 1) & 1);
      uint8_t* sQ = smem + S::kOffQ + g * S::kQ;
#pragma unroll
      for (int u = 0; u < kItemsHalf; ++u) {
        const int item = lane + 32 * u;
        const int rr = hh * 64 + item / kChunksPerRow, c = item % kChunksPerRow;
        *reinterpret_cast<uint4*>(sQ + c * (kQB * 16) + (rr >> 3) * 128 + (rr & 7) * 16) = r[u];
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&q_full[g]);
      ++qn;
    };
    auto gather_kv = [&](int ch, int kv_beg, int kv_len, int h) {
      if ((ever >> ch) & 1) tc::mbar_wait_sleep(&kv_free[ch], ((ldmask >> ch) & 1) ^ 1);
      const uint32_t sK = tc::smem_u32(smem + S::kOffK), sV = tc::smem_u32(smem + S::kOffV);
#pragma unroll 2
      for (int i0 = 0; i0 < 2 * kItemsHalf; i0 += 4) {
        const __nv_bfloat16* src[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int item = lane + 32 * (i0 + u);
          const int rr = item / kChunksPerRow, c = item - rr * kChunksPerRow;
          const int j = ch * kKC + rr;
          ok[u] = j < kv_len;
          src[u] = qkv + (ok[u] ? (size_t)order_row[kv_beg + j] * (3 * C) : 0) + h * D + c * 8;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int item = lane + 32 * (i0 + u);
          const int rr = item / kChunksPerRow, c = item - rr * kChunksPerRow;
          const int j = ch * kKC + rr;
          const uint32_t off = (uint32_t)(c * (KMAX * 16) + (j >> 3) * 128 + (j & 7) * 16);
          tc::cp_async16(sK + off, src[u] + C, ok[u] ? 16u : 0u);
          tc::cp_async16(sV + off, src[u] + 2 * C, ok[u] ? 16u : 0u);
        }
      }
      tc::cp_async_mbar_arrive_noinc(&kv_full[ch]);
      ever |= 1u << ch;
      ldmask ^= 1u << ch;
    };
    // next non-empty item at or after w
    auto next_item = [&](int w, int4& e) {
      for (; w < n_items; w += gridDim.x) {
        e = table[w / H];
        if (e.y - e.x > 0) return w;
      }
      return n_items;
    };
    // One iteration per "slot" = (item, tile pair i): i = 0: the item's K/V chunks and the group's first Q tile; i > 0: the
    // group's next Q tile (prefetched into registers one slot earlier, stored when the Q buffer is released).
    int4 e;
    int w = next_item(blockIdx.x, e);
    int i = 0;
    bool have_q = false;
    if (w < n_items && (g == 0 || (e.y - e.x + kQB - 1) / kQB > 1)) {
      load_q(e.x, e.y - e.x, w % H, g);
      have_q = true;
    }
    while (w < n_items) {
      const int h = w % H;
      const int q_beg = e.x, n_q = e.y - e.x, kv_beg = e.z, kv_len = e.w - e.z;
      const int nch = (kv_len + kKC - 1) / kKC;
      const int nqb = (n_q + kQB - 1) / kQB;
      const int nt0 = (nqb + 1) / 2;
      if (i == 0) {
        if (lw < nch) gather_kv(lw, kv_beg, kv_len, h);
        if (have_q) store_q();
        if (lw + 4 < nch) gather_kv(lw + 4, kv_beg, kv_len, h);
      } else if (have_q) {
        store_q();
      }
      if (++i == nt0) {
        w = next_item(w + gridDim.x, e);
        i = 0;
      }
      have_q = false;
      if (w < n_items) {
        const int nq2 = e.y - e.x;
        const int nqb2 = (nq2 + kQB - 1) / kQB;
        if (i < (g == 0 ? (nqb2 + 1) / 2 : nqb2 / 2)) {
          load_q(e.x, nq2, w % H, 2 * i + g);
          have_q = true;
        }
      }
    }
  } else {
    // =========================================================== MMA issuers (warp 20: group 0, warp 21: group 1)
    const int g = warp - 20;
    constexpr uint32_t idesc_s = tc::umma_idesc_bf16(kQB, kKC, 0, 0);
    constexpr uint32_t idesc_o = tc::umma_idesc_bf16(kQB, D, 0, 1);
    constexpr uint32_t idesc_l = tc::umma_idesc_bf16(kQB, 16, 0, 1);
    const uint32_t sK = tc::smem_u32(smem + S::kOffK), sV = tc::smem_u32(smem + S::kOffV);
    const uint32_t q0 = tc::smem_u32(smem + S::kOffQ + g * S::kQ) >> 4;
    const uint32_t tSg = tmem_base + g * kKC;
    const uint32_t tPg = tmem_base + kPCol + g * kPStride;
    const uint32_t tOg = tmem_base + kOCol + g * kOStride;
    const uint64_t dq_base = tc::umma_desc_nosw(0, kQB * 16, 128);
    const uint64_t dk_base = tc::umma_desc_nosw(0, KMAX * 16, 128);
    const uint64_t dv_base = tc::umma_desc_nosw(0, 128, KMAX * 16);
    const uint64_t d_ones = tc::umma_desc_nosw(tc::smem_u32(smem + S::kOffOnes), 128, 256);
    uint32_t kvmask = 0;  // bit c: parity of the number of loads of chunk c consumed so far
    int sb = 0;           // steps of this group before the current item (barrier parities)
    int tn = 0;           // query tiles of this group so far
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int4 e = table[w / H];
      const int n_q = e.y - e.x, kv_len = e.w - e.z;
      if (n_q <= 0) continue;
      const int nch = (kv_len + kKC - 1) / kKC;
      const int nqb = (n_q + kQB - 1) / kQB;
      const int ntiles = g == 0 ? (nqb + 1) / 2 : nqb / 2;
      const int total = ntiles * nch;
      const bool both = nqb > 1;  // the other group has tiles in this item too
      int kv_ready = 0, jn = 0;
      auto issue_qk = [&]() {
        const int j = jn;
        if (j == 0) {
          tc::mbar_wait(&q_full[g], tn & 1);
          tc::tc_fence_after();
        }
        if (kv_ready <= j) {
          tc::mbar_wait(&kv_full[j], (kvmask >> j) & 1);
          tc::tc_fence_after();
          kv_ready = j + 1;
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
          tc::umma_commit_elect(&q_free[g]);
          jn = 0;
          ++tn;
        } else {
          jn = j + 1;
        }
      };
      if (total > 0) issue_qk();
      int pc = 0;
      for (int s = 0; s < total; ++s) {
        const bool qk_first = s + 1 < total && jn != 0;
        if (qk_first) {
          tc::mbar_wait(&s_free[g], (sb + s) & 1);
          tc::tc_fence_after();
          issue_qk();
        }
        tc::mbar_wait(&p_ready[g], (sb + s) & 1);
        tc::tc_fence_after();
        const uint32_t v0 = (sV + pc * (kKC / 8) * 128) >> 4;
#pragma unroll
        for (int t = 0; t < kKC / 16; ++t) {
          const uint64_t dv = dv_base | (uint64_t)((v0 + t * 16) & 0x3fff);
          tc::umma_bf16_ts_elect(tOg, tPg + 8 * t, dv, idesc_o, (pc | t) ? 1u : 0u);
        }
#pragma unroll
        for (int t = 0; t < kKC / 16; ++t)
          tc::umma_bf16_ts_elect(tOg + D, tPg + 8 * t, d_ones, idesc_l, (pc | t) ? 1u : 0u);
        tc::umma_commit_elect(&pv_done[g]);
        if (s >= total - nch) {  // last query tile of the group in this item: chunk pc is retired
          tc::umma_commit_elect(&kv_free[pc]);
          if (!both) tc::umma_commit_elect(&kv_free[pc]);  // on behalf of the group without tiles
        }
        pc = pc == nch - 1 ? 0 : pc + 1;
        if (s + 1 < total && !qk_first) issue_qk();
      }
      sb += total;
      kvmask ^= (1u << nch) - 1u;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 20) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

template <int D, int POLY>
static int launch_attention16(const void* qkv, const int64_t* order_row, const int32_t* table, int max_patches,
                              int heads, float scale, void* out, float* lse2, int64_t lse_stride, cudaStream_t stream) {
  constexpr int KMAX = 1024;
  using S = Att16Smem<D, KMAX>;
  auto kern = patch_attention_tc16_kernel<D, KMAX, POLY>;
  SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
  const int n_items = heads * max_patches;
  dim3 grid((unsigned)(n_items < kNumSMs ? n_items : kNumSMs));
  constexpr int pp = SS_ATT_PP;
  kern<<<grid, kAtt16Threads, S::kTotal, stream>>>((const __nv_bfloat16*)qkv, order_row, (const int4*)table, heads, n_items,
                                                   scale * 1.4426950408889634f, (__nv_bfloat16*)out, pp, lse2, lse_stride);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

// SS_ATT_POLY (compile time): exponentials per 8 evaluated on the FMA pipe (2 and 3 tie at d = 48, 2 is 3 % faster at
// d = 16 / 32); SS_ATT_PP: group ping-pong (1).  tools/micro/att_bench.cu builds this file with other values for A/B.
template <int D>
static int launch_attention(const void* qkv, const int64_t* order_row, const int32_t* table, int max_patches, int heads,
                            float scale, void* out, float* lse2, int64_t lse_stride, cudaStream_t stream) {
  return launch_attention16<D, SS_ATT_POLY>(qkv, order_row, table, max_patches, heads, scale, out, lse2, lse_stride, stream);
}

}  // namespace ss

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
    case 16: return ss::launch_attention<16>(qkv_bf16, order_row, table, max_patches, heads, scale, out_bf16, lse2, lse_stride, stream);
    case 32: return ss::launch_attention<32>(qkv_bf16, order_row, table, max_patches, heads, scale, out_bf16, lse2, lse_stride, stream);
    case 48: return ss::launch_attention<48>(qkv_bf16, order_row, table, max_patches, heads, scale, out_bf16, lse2, lse_stride, stream);
    default: return SS_BAD_ARGS;
  }
}

extern "C" int ss_patch_attention(const void* qkv_bf16, const int64_t* order_row, const int32_t* table, int max_patches,
                                  int patch_size, int heads, int head_dim, float scale, void* out_bf16, void* stream_) {
  return patch_attention_entry(qkv_bf16, order_row, table, max_patches, patch_size, heads, head_dim, scale, out_bf16,
                               nullptr, 0, stream_);
}

extern "C" int ss_patch_attention_lse(const void* qkv_bf16, const int64_t* order_row, const int32_t* table,
                                      int max_patches, int patch_size, int heads, int head_dim, float scale,
                                      void* out_bf16, float* lse2, int64_t n, void* stream_) {
  if (!lse2 || n < 0) return SS_BAD_ARGS;
  return patch_attention_entry(qkv_bf16, order_row, table, max_patches, patch_size, heads, head_dim, scale, out_bf16,
                               lse2, n, stream_);
}
