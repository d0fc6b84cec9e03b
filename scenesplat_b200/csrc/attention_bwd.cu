// Patch attention BACKWARD on the 5th-gen tensor cores (tcgen05 + TMEM), bf16 in / bf16 out.
//
// Replaces (reference): the autograd backward of flash_attn_varlen_qkvpacked_func together with the adjoints of the
// `qkv[order]` / `feat[inverse]` gathers around it (point_transformer_v3m1_base.py:181-216), on the same device patch
// table as the forward (window rule of :114-170).
//
// With P = softmax(scale * Q K^T) of one (head, patch), delta_i = sum_d dO_id O_id and dS = P o (dP - delta):
//     dP = dO V^T        dV = P^T dO        dQ = scale * dS K        dK = scale * dS^T Q
// P is recomputed from the forward's log2-domain log-sum-exp (ss_patch_attention_lse), so there is no running maximum
// and no rescaling here.  Every product has one of the two operand forms of the forward kernel (attention_tc.cu):
//     T   = X Y^T          X, Y K-major shared-memory tiles             (like S = Q K^T)
//     acc += A Y           A bf16 in TMEM, Y an MN-major shared-memory tile (like O += P V)
// and those forms need the OUTPUT rows on the TMEM lanes.  The kernel is therefore run twice with the roles swapped:
//   MODE 0 (dQ)      lanes = 128 queries (X1 = Q tile, X2 = dO tile), resident column side = K, V of the patch:
//                    T1 = Q K^T, T2 = dO V^T, U = P o (T2 - delta_row), dQ += U K
//   MODE 1 (dK, dV)  lanes = 128 keys (X1 = K tile, X2 = V tile), resident column side = Q, dO of the patch:
//                    T1 = K Q^T = S^T, T2 = V dO^T = dP^T, W = P^T, U = dS^T (lse / delta per COLUMN, from shared
//                    memory), dV += W dO, dK += U Q
// The byte layout of a shared-memory tile is the same for its K-major and MN-major use (8 x 16-byte core matrices),
// so K (mode 0) / Q (mode 1) are gathered once and serve both products.
//
// One CTA per (head, patch), 12 warps: 0-7 elementwise warps (row quarter, column half: 32 rows x 64 columns of every
// 128 x 128 tile), 8 / 11 loaders of X1 / X2 (next tile prefetched into registers), 9 MMA issuer, 10 vectors (mode 1).
// TMEM columns: T1 [0,128) T2 [128,256) | W [256,320) U [320,384) (bf16 pairs) | acc1 [384,448) acc2 [448,512).
// The next T1 / T2 are issued as soon as the current ones are in registers, i.e. they run under the exponentials.
// dQ rows are written straight to d(qkv) (every query belongs to one patch), and so are the dK / dV rows of keys only
// one patch attends to.  The window of an item's last patch shares K - r keys with the patch before it: those rows are
// added into a zeroed fp32 scratch with vector reductions (two addends per element, so the sum does not depend on
// the order) and converted by a last small kernel (attn_shared_rows_kernel).  Tables other than ss_patch_table's
// must keep that convention: kv ranges overlap only as [kv_begin, q_begin) of an entry with the entry before it.
#include "tc_common.cuh"
#include "attention_math.cuh"
#include "../../include/scenesplat_b200.h"

#ifndef SS_ATTB_POLY
#define SS_ATTB_POLY 0  // exponentials per 8 evaluated on the FMA pipe (compile-time; measured 0..3: 0 is fastest here)
#endif

namespace ss {

constexpr int kBwThreads = 384;
constexpr int kBwT = 128;  // tile edge (lane rows and column chunk)
constexpr int kBwColT1 = 0, kBwColT2 = 128, kBwColW = 256, kBwColU = 320, kBwColA1 = 384, kBwColA2 = 448;

template <int D, int KMAX>
struct AttBwdSmem {
  static constexpr int kY = KMAX * D * 2;  // one resident column-side matrix
  static constexpr int kX = kBwT * D * 2;  // one lane-side tile
  static constexpr int kOffY1 = 0;
  static constexpr int kOffY2 = kY;
  static constexpr int kOffX1 = 2 * kY;
  static constexpr int kOffX2 = 2 * kY + kX;
  static constexpr int kOffVec = 2 * kY + 2 * kX;  // mode 1: lse2[KMAX], delta[KMAX] of the patch's queries
  static constexpr int kOffBar = kOffVec + 2 * KMAX * 4;
  static constexpr int kTotal = kOffBar + 256 + 128;
};

// 16-byte shared-memory load by 32-bit shared address (the generic pointer would compile to LD.E, not LDS)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// delta[h, pos] = sum_d dO[order[pos], h, d] * O[order[pos], h, d]   (sorted positions, like lse2)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                  const int64_t* __restrict__ order_row, int64_t n, int H, int D, float* __restrict__ delta) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * H) return;
  const int64_t pos = idx / H;
  const int h = (int)(idx - pos * H);
  const size_t base = (size_t)order_row[pos] * ((size_t)H * D) + (size_t)h * D;
  float acc = 0.f;
  for (int c = 0; c < D; c += 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(out + base + c);
    const uint4 b = *reinterpret_cast<const uint4*>(dout + base + c);
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 fa = __bfloat1622float2(pa[u]), fb = __bfloat1622float2(pb[u]);
      acc = fmaf(fa.x, fb.x, acc);
      acc = fmaf(fa.y, fb.y, acc);
    }
  }
  delta[(size_t)h * n + pos] = acc;
}

// Keys that two patches attend to (the window of an item's last patch reaches K - r rows back into the patch before
// it: sorted positions [kv_begin, q_begin) of every table entry with kv_begin < q_begin) get their dK / dV from two
// CTAs.  Only those rows go through the fp32 scratch: STORE = 0 zeroes them before the backward kernels, STORE = 1
// converts the sums into d(qkv) afterwards.  grid = (max_patches, 8 row slices).
template <int STORE>
__global__ void __launch_bounds__(256)
attn_shared_rows_kernel(const int4* __restrict__ table, const int64_t* __restrict__ order_row, int C,
                        float* __restrict__ dkv32, __nv_bfloat16* __restrict__ dqkv) {
  const int4 e = table[blockIdx.x];
  if (e.y <= e.x || e.z >= e.x) return;  // unused entry / nothing borrowed
  const int per_row = 2 * C / 8;
  for (int pos = e.z + blockIdx.y; pos < e.x; pos += gridDim.y) {
    const size_t r = (size_t)order_row[pos];
    for (int v = threadIdx.x; v < per_row; v += 256) {
      float4* src = reinterpret_cast<float4*>(dkv32 + r * 2 * C + v * 8);
      if (STORE) {
        const float4 a = src[0], b = src[1];
        uint4 o;
        o.x = tc::pack_bf16(a.x, a.y);
        o.y = tc::pack_bf16(a.z, a.w);
        o.z = tc::pack_bf16(b.x, b.y);
        o.w = tc::pack_bf16(b.z, b.w);
        *reinterpret_cast<uint4*>(dqkv + r * 3 * C + C + v * 8) = o;
      } else {
        src[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        src[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

// POLY: of every 8 exponentials, POLY are evaluated on the FMA pipe (exp2_poly), the rest by MUFU.EX2
template <int D, int KMAX, int MODE, int POLY>
__global__ void __launch_bounds__(kBwThreads, 1)
patch_attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                           const float* __restrict__ lse2, const float* __restrict__ delta, int64_t vec_stride,
                           const int64_t* __restrict__ order_row, const int4* __restrict__ table, int n_patches, int H,
                           float scale, float scale_log2e, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dkv32) {
  using S = AttBwdSmem<D, KMAX>;
  const int4 e = table[blockIdx.x / H];
  const int q_beg = e.x, n_q = e.y - e.x, kv_beg = e.z, kv_len = e.w - e.z;
  if (n_q <= 0) return;  // block-uniform: unused table entry
  // first sorted position of this patch's keys that the NEXT table entry's window borrows (ss_patch_table puts an
  // item's last patch right after the patch it borrows from); INT_MAX if none
  int lend_beg = 0x7fffffff;
  if (MODE == 1 && (int)(blockIdx.x / H) + 1 < n_patches) {
    const int4 nx = table[blockIdx.x / H + 1];
    if (nx.y > nx.x && nx.z < nx.x && nx.z >= e.z && nx.z < e.w) lend_beg = nx.z;
  }
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  uint64_t* bars = (uint64_t*)(smem + S::kOffBar);
  uint64_t* y_full = bars;         // [8]  column-side chunk c landed (32 lane arrivals of the warp that gathered it)
  uint64_t* x_full = bars + 8;     //      lane tile landed (2 loader warps x 32 lanes)
  uint64_t* x_free = bars + 9;     //      last T1 / T2 of the tile done: X1 / X2 may be refilled
  uint64_t* t_full = bars + 10;    //      T1 and T2 ready
  uint64_t* t_free = bars + 11;    //      T1 and T2 are in registers (256 threads)
  uint64_t* wu_ready = bars + 12;  //      W and U written (256 threads)
  uint64_t* acc_done = bars + 13;  //      accumulator MMAs of the step done: W / U may be rewritten
  uint32_t* tmem_slot = (uint32_t*)(bars + 14);
  float* sL = (float*)(smem + S::kOffVec);
  float* sD = sL + KMAX;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int h = blockIdx.x % H;
  const int C = H * D;
  // lane side / column side in sorted positions
  const int l_beg = MODE == 0 ? q_beg : kv_beg, l_len = MODE == 0 ? n_q : kv_len;
  const int c_beg = MODE == 0 ? kv_beg : q_beg, c_len = MODE == 0 ? kv_len : n_q;
  const int nlt = (l_len + kBwT - 1) / kBwT;  // lane tiles
  const int nct = (c_len + kBwT - 1) / kBwT;  // column chunks
  // the four operand matrices: element (row r, channel c) at p + r * stride + c
  const __nv_bfloat16* pQ = qkv + h * D;
  const __nv_bfloat16* pK = qkv + C + h * D;
  const __nv_bfloat16* pV = qkv + 2 * C + h * D;
  const __nv_bfloat16* pG = dout + h * D;
  const __nv_bfloat16* pX1 = MODE == 0 ? pQ : pK;
  const __nv_bfloat16* pX2 = MODE == 0 ? pG : pV;
  const __nv_bfloat16* pY1 = MODE == 0 ? pK : pQ;
  const __nv_bfloat16* pY2 = MODE == 0 ? pV : pG;
  const size_t sX1 = 3 * (size_t)C, sX2 = MODE == 0 ? (size_t)C : 3 * (size_t)C;
  const size_t sY1 = 3 * (size_t)C, sY2 = MODE == 0 ? 3 * (size_t)C : (size_t)C;
  constexpr int kChunksPerRow = D / 8;       // 16-byte pieces per row
  constexpr int kItems = 4 * kChunksPerRow;  // pieces per lane per 128 rows

  if (threadIdx.x == 0) {
    for (int c = 0; c < 8; ++c) tc::mbar_init(&y_full[c], 32);
    tc::mbar_init(x_full, 64);
    tc::mbar_init(x_free, 1);
    tc::mbar_init(t_full, 1);
    tc::mbar_init(t_free, 256);
    tc::mbar_init(wu_ready, 256);
    tc::mbar_init(acc_done, 1);
    tc::mbar_fence_init();
  }
  if (warp == 9) tc::tmem_alloc<512>(tmem_slot);

  // Gathers: element (row j, piece c) -> c * (ROWS * 16) + (j / 8) * 128 + (j % 8) * 16 (UMMA no-swizzle core matrices)
  auto gather_y = [&](int ch) {
    const uint32_t a1 = tc::smem_u32(smem + S::kOffY1), a2 = tc::smem_u32(smem + S::kOffY2);
#pragma unroll
    for (int i0 = 0; i0 < kItems; i0 += 4) {
      int64_t src_row[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int item = lane + 32 * (i0 + u);
        const int j = ch * kBwT + item / kChunksPerRow;
        ok[u] = j < c_len;
        src_row[u] = ok[u] ? order_row[c_beg + j] : 0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int item = lane + 32 * (i0 + u);
        const int r = item / kChunksPerRow, c = item - r * kChunksPerRow;
        const int j = ch * kBwT + r;
        const uint32_t off = (uint32_t)(c * (KMAX * 16) + (j >> 3) * 128 + (j & 7) * 16);
        tc::cp_async16(a1 + off, pY1 + (size_t)src_row[u] * sY1 + c * 8, ok[u] ? 16u : 0u);
        tc::cp_async16(a2 + off, pY2 + (size_t)src_row[u] * sY2 + c * 8, ok[u] ? 16u : 0u);
      }
    }
  };
  auto gather_x0 = [&](uint32_t dst, const __nv_bfloat16* p, size_t stride) {  // lane tile 0, asynchronously
#pragma unroll
    for (int i0 = 0; i0 < kItems; i0 += 4) {
      int64_t src_row[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = (lane + 32 * (i0 + u)) / kChunksPerRow;
        ok[u] = r < l_len;
        src_row[u] = ok[u] ? order_row[l_beg + r] : 0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int item = lane + 32 * (i0 + u);
        const int r = item / kChunksPerRow, c = item - r * kChunksPerRow;
        tc::cp_async16(dst + (uint32_t)(c * (kBwT * 16) + (r >> 3) * 128 + (r & 7) * 16),
                       p + (size_t)src_row[u] * stride + c * 8, ok[u] ? 16u : 0u);
      }
    }
  };
  if (warp < 8 && warp < nct) gather_y(warp);
  if (warp == 8) gather_x0(tc::smem_u32(smem + S::kOffX1), pX1, sX1);
  if (warp == 11) gather_x0(tc::smem_u32(smem + S::kOffX2), pX2, sX2);
  if (MODE == 1 && warp == 10) {
    // per-COLUMN log-sum-exp and delta of the patch's queries; padding columns get W = U = 0 (lse = +inf)
    for (int i = lane; i < nct * kBwT; i += 32) {
      const bool ok = i < c_len;
      sL[i] = ok ? lse2[(size_t)h * vec_stride + c_beg + i] : INFINITY;
      sD[i] = ok ? delta[(size_t)h * vec_stride + c_beg + i] : 0.f;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // =========================================================== elementwise warps
    if (warp < nct) {  // publish the chunk this warp gathered
      tc::cp_async_wait_all();
      tc::fence_proxy_async();
      tc::mbar_arrive(&y_full[warp]);
    }
    const int quarter = warp & 3, hh = warp >> 2;
    const int row = quarter * 32 + lane;  // row inside the lane tile == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t tT1 = t_lane + kBwColT1 + 64 * hh, tT2 = t_lane + kBwColT2 + 64 * hh;
    const uint32_t tW = t_lane + kBwColW + 32 * hh, tU = t_lane + kBwColU + 32 * hh;
    const uint32_t sL_addr = tc::smem_u32(sL);
    int s = 0;
    for (int i = 0; i < nlt; ++i) {
      const int li = i * kBwT + row;
      const int64_t orow = li < l_len ? order_row[l_beg + li] : -1;
      float lrow = INFINITY, drow = 0.f;  // mode 0: per-row log-sum-exp / delta (rows past the end: W = U = 0)
      if (MODE == 0 && orow >= 0) {
        lrow = lse2[(size_t)h * vec_stride + l_beg + li];
        drow = delta[(size_t)h * vec_stride + l_beg + li];
      }
      for (int j = 0; j < nct; ++j, ++s) {
        uint32_t pkW[2][16], pkU[2][16];
        tc::mbar_wait(t_full, s & 1);
        tc::tc_fence_after();
        // column quarter 0 is loaded first; quarter 1's TMEM loads are in flight under quarter 0's arithmetic
        uint32_t a[2][32], b[2][32];
        tc::tmem_ld32(tT1, a[0]);
        tc::tmem_ld32(tT2, b[0]);
        tc::tmem_ld_wait();
        tc::tmem_ld32(tT1 + 32, a[1]);
        tc::tmem_ld32(tT2 + 32, b[1]);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (q == 1) {
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            tc::mbar_arrive(t_free);  // the next T1 / T2 run under this step's exponentials
          }
          const uint32_t aL = sL_addr + (uint32_t)(j * kBwT + 64 * hh + 32 * q) * 4u, aD = aL + KMAX * 4u;
#pragma unroll
          for (int u4 = 0; u4 < 8; ++u4) {
            float l[4] = {lrow, lrow, lrow, lrow}, dl[4] = {drow, drow, drow, drow};
            if (MODE == 1) {  // same address in every lane: shared-memory broadcast
              const float4 lv = lds128(aL + 16 * u4), dv = lds128(aD + 16 * u4);
              l[0] = lv.x; l[1] = lv.y; l[2] = lv.z; l[3] = lv.w;
              dl[0] = dv.x; dl[1] = dv.y; dl[2] = dv.z; dl[3] = dv.w;
            }
            float w[4], g[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const int u = 4 * u4 + v;
              const float x = fmaf(__uint_as_float(a[q][u]), scale_log2e, -l[v]);
              w[v] = (u & 7) < POLY ? exp2_poly(x) : ex2_approx(x);
              g[v] = w[v] * (__uint_as_float(b[q][u]) - dl[v]);
            }
            pkW[q][2 * u4] = tc::pack_bf16(w[0], w[1]);
            pkW[q][2 * u4 + 1] = tc::pack_bf16(w[2], w[3]);
            pkU[q][2 * u4] = tc::pack_bf16(g[0], g[1]);
            pkU[q][2 * u4 + 1] = tc::pack_bf16(g[2], g[3]);
          }
        }
        if (s > 0) {  // W / U are still being read by the previous step's accumulator MMAs
          tc::mbar_wait(acc_done, (s - 1) & 1);
          tc::tc_fence_after();
        }
        if (MODE == 1) {
          tc::tmem_st16(tW, pkW[0]);
          tc::tmem_st16(tW + 16, pkW[1]);
        }
        tc::tmem_st16(tU, pkU[0]);
        tc::tmem_st16(tU + 16, pkU[1]);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(wu_ready);
      }
      // ---- epilogue of the lane tile
      tc::mbar_wait(acc_done, (s - 1) & 1);
      tc::tc_fence_after();
      if (MODE == 0) {
        if (hh == 0) {  // dQ rows: scale * acc1 -> bf16 -> the query's own row of d(qkv)
          __nv_bfloat16* dst = orow >= 0 ? dqkv + (size_t)orow * (3 * C) + h * D : nullptr;
#pragma unroll
          for (int jo = 0; jo < D / 16; ++jo) {
            uint32_t o[16];
            tc::tmem_ld16(t_lane + kBwColA1 + jo * 16, o);
            tc::tmem_ld_wait();
            if (dst) {
              uint4 o0, o1;
              o0.x = tc::pack_bf16(__uint_as_float(o[0]) * scale, __uint_as_float(o[1]) * scale);
              o0.y = tc::pack_bf16(__uint_as_float(o[2]) * scale, __uint_as_float(o[3]) * scale);
              o0.z = tc::pack_bf16(__uint_as_float(o[4]) * scale, __uint_as_float(o[5]) * scale);
              o0.w = tc::pack_bf16(__uint_as_float(o[6]) * scale, __uint_as_float(o[7]) * scale);
              o1.x = tc::pack_bf16(__uint_as_float(o[8]) * scale, __uint_as_float(o[9]) * scale);
              o1.y = tc::pack_bf16(__uint_as_float(o[10]) * scale, __uint_as_float(o[11]) * scale);
              o1.z = tc::pack_bf16(__uint_as_float(o[12]) * scale, __uint_as_float(o[13]) * scale);
              o1.w = tc::pack_bf16(__uint_as_float(o[14]) * scale, __uint_as_float(o[15]) * scale);
              uint4* d4 = reinterpret_cast<uint4*>(dst + jo * 16);
              d4[0] = o0;
              d4[1] = o1;
            }
          }
        }
      } else {
        // column half 0: dK = scale * acc2, column half 1: dV = acc1.  A key that only this patch attends to is
        // written straight to d(qkv) as bf16; a key shared with the neighbouring patch (borrowed by this patch's
        // window, or lent to the next entry's window) is added into the zeroed fp32 scratch [n, 2C] (K | V)
        const float f = hh == 0 ? scale : 1.f;
        const uint32_t tA = t_lane + (hh == 0 ? kBwColA2 : kBwColA1);
        const int pos = l_beg + li;
        const bool shared_row = pos < q_beg || pos >= lend_beg;
        float* dst32 = orow >= 0 ? dkv32 + (size_t)orow * (2 * C) + (hh == 0 ? 0 : C) + h * D : nullptr;
        __nv_bfloat16* dst16 = orow >= 0 ? dqkv + (size_t)orow * (3 * C) + (hh == 0 ? C : 2 * C) + h * D : nullptr;
#pragma unroll
        for (int jo = 0; jo < D / 16; ++jo) {
          uint32_t o[16];
          tc::tmem_ld16(tA + jo * 16, o);
          tc::tmem_ld_wait();
          if (orow >= 0) {
            if (shared_row) {
#pragma unroll
              for (int v = 0; v < 4; ++v)
                red_add_v4(dst32 + jo * 16 + 4 * v, __uint_as_float(o[4 * v]) * f, __uint_as_float(o[4 * v + 1]) * f,
                           __uint_as_float(o[4 * v + 2]) * f, __uint_as_float(o[4 * v + 3]) * f);
            } else {
              uint4 o0, o1;
              o0.x = tc::pack_bf16(__uint_as_float(o[0]) * f, __uint_as_float(o[1]) * f);
              o0.y = tc::pack_bf16(__uint_as_float(o[2]) * f, __uint_as_float(o[3]) * f);
              o0.z = tc::pack_bf16(__uint_as_float(o[4]) * f, __uint_as_float(o[5]) * f);
              o0.w = tc::pack_bf16(__uint_as_float(o[6]) * f, __uint_as_float(o[7]) * f);
              o1.x = tc::pack_bf16(__uint_as_float(o[8]) * f, __uint_as_float(o[9]) * f);
              o1.y = tc::pack_bf16(__uint_as_float(o[10]) * f, __uint_as_float(o[11]) * f);
              o1.z = tc::pack_bf16(__uint_as_float(o[12]) * f, __uint_as_float(o[13]) * f);
              o1.w = tc::pack_bf16(__uint_as_float(o[14]) * f, __uint_as_float(o[15]) * f);
              uint4* d4 = reinterpret_cast<uint4*>(dst16 + jo * 16);
              d4[0] = o0;
              d4[1] = o1;
            }
          }
        }
      }
      tc::tc_fence_before();  // ordered before the next tile's first accumulator MMA by the next wu_ready arrival
    }
  } else if (warp == 8 || warp == 11) {
    // =========================================================== loaders of X1 (warp 8) / X2 (warp 11)
    const __nv_bfloat16* p = warp == 8 ? pX1 : pX2;
    const size_t stride = warp == 8 ? sX1 : sX2;
    uint8_t* sX = smem + (warp == 8 ? S::kOffX1 : S::kOffX2);
    tc::cp_async_wait_all();
    tc::fence_proxy_async();
    tc::mbar_arrive(x_full);
    // later tiles: rows prefetched into registers while the previous tile is processed, stored the moment its last
    // T1 / T2 have released the buffer
    for (int i = 1; i < nlt; ++i) {
      uint4 r[kItems];
#pragma unroll
      for (int u = 0; u < kItems; ++u) {
        const int item = lane + 32 * u;
        const int rr = item / kChunksPerRow, c = item - rr * kChunksPerRow;
        const int li = i * kBwT + rr;
        r[u] = make_uint4(0u, 0u, 0u, 0u);
        if (li < l_len)
          r[u] = __ldg(reinterpret_cast<const uint4*>(p + (size_t)order_row[l_beg + li] * stride + c * 8));
      }
      tc::mbar_wait_sleep(x_free, (i - 1) & 1);
#pragma unroll
      for (int u = 0; u < kItems; ++u) {
        const int item = lane + 32 * u;
        const int rr = item / kChunksPerRow, c = item - rr * kChunksPerRow;
        *reinterpret_cast<uint4*>(sX + c * (kBwT * 16) + (rr >> 3) * 128 + (rr & 7) * 16) = r[u];
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(x_full);
    }
  } else if (warp == 9) {
    // =========================================================== MMA issuer (whole warp, one elected lane per op)
    const int total = nlt * nct;
    constexpr uint32_t idesc_t = tc::umma_idesc_bf16(kBwT, kBwT, 0, 0);  // T = X Y^T : M=128, N=128
    constexpr uint32_t idesc_a = tc::umma_idesc_bf16(kBwT, D, 0, 1);     // acc += A Y : M=128, N=D, B MN-major
    const uint32_t aY1 = tc::smem_u32(smem + S::kOffY1), aY2 = tc::smem_u32(smem + S::kOffY2);
    const uint32_t x1 = tc::smem_u32(smem + S::kOffX1) >> 4, x2 = tc::smem_u32(smem + S::kOffX2) >> 4;
    const uint64_t dx_base = tc::umma_desc_nosw(0, kBwT * 16, 128);
    const uint64_t dy_base = tc::umma_desc_nosw(0, KMAX * 16, 128);
    const uint64_t dyt_base = tc::umma_desc_nosw(0, 128, KMAX * 16);
    int y_ready = 0;
    int jn = 0, tn = 0;  // chunk / tile index of the NEXT T1, T2

    auto issue_t = [&]() {
      const int j = jn;
      if (j == 0) {
        tc::mbar_wait(x_full, tn & 1);
        tc::tc_fence_after();
      }
      if (y_ready <= j) {
        tc::mbar_wait(&y_full[j], 0);
        tc::tc_fence_after();
        y_ready = j + 1;
      }
      const uint32_t y1 = (aY1 + j * (kBwT / 8) * 128) >> 4, y2 = (aY2 + j * (kBwT / 8) * 128) >> 4;
#pragma unroll
      for (int t = 0; t < D / 16; ++t)
        tc::umma_bf16_elect(tmem_base + kBwColT1, dx_base | (uint64_t)((x1 + 2 * t * kBwT) & 0x3fff),
                            dy_base | (uint64_t)((y1 + 2 * t * KMAX) & 0x3fff), idesc_t, t ? 1u : 0u);
#pragma unroll
      for (int t = 0; t < D / 16; ++t)
        tc::umma_bf16_elect(tmem_base + kBwColT2, dx_base | (uint64_t)((x2 + 2 * t * kBwT) & 0x3fff),
                            dy_base | (uint64_t)((y2 + 2 * t * KMAX) & 0x3fff), idesc_t, t ? 1u : 0u);
      tc::umma_commit_elect(t_full);
      if (j == nct - 1) {
        tc::umma_commit_elect(x_free);  // last read of this lane tile
        jn = 0;
        ++tn;
      } else {
        jn = j + 1;
      }
    };

    if (total > 0) issue_t();
    int pc = 0;  // column chunk of the next accumulator MMAs
    for (int s = 0; s < total; ++s) {
      const bool t_first = s + 1 < total && jn != 0;  // next T1 / T2 of the same lane tile: only needs t_free
      if (t_first) {
        tc::mbar_wait(t_free, s & 1);
        tc::tc_fence_after();
        issue_t();
      }
      tc::mbar_wait(wu_ready, s & 1);
      tc::tc_fence_after();
      const uint32_t y1 = (aY1 + pc * (kBwT / 8) * 128) >> 4, y2 = (aY2 + pc * (kBwT / 8) * 128) >> 4;
      if (MODE == 0) {
#pragma unroll
        for (int t = 0; t < kBwT / 16; ++t)  // dQ += U K
          tc::umma_bf16_ts_elect(tmem_base + kBwColA1, tmem_base + kBwColU + 8 * t,
                                 dyt_base | (uint64_t)((y1 + t * 16) & 0x3fff), idesc_a, (pc | t) ? 1u : 0u);
      } else {
#pragma unroll
        for (int t = 0; t < kBwT / 16; ++t)  // dV += W dO
          tc::umma_bf16_ts_elect(tmem_base + kBwColA1, tmem_base + kBwColW + 8 * t,
                                 dyt_base | (uint64_t)((y2 + t * 16) & 0x3fff), idesc_a, (pc | t) ? 1u : 0u);
#pragma unroll
        for (int t = 0; t < kBwT / 16; ++t)  // dK += U Q
          tc::umma_bf16_ts_elect(tmem_base + kBwColA2, tmem_base + kBwColU + 8 * t,
                                 dyt_base | (uint64_t)((y1 + t * 16) & 0x3fff), idesc_a, (pc | t) ? 1u : 0u);
      }
      tc::umma_commit_elect(acc_done);
      pc = pc == nct - 1 ? 0 : pc + 1;
      if (s + 1 < total && !t_first) issue_t();  // first T1 / T2 of the next lane tile (wu_ready: T was consumed)
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

template <int D, int MODE>
static int launch_attention_bwd(const void* qkv, const void* dout, const float* lse2, const float* delta, int64_t n,
                                const int64_t* order_row, const int32_t* table, int max_patches, int heads, float scale,
                                void* dqkv, float* dkv32, cudaStream_t stream) {
  constexpr int KMAX = 1024;
  using S = AttBwdSmem<D, KMAX>;
  auto kern = patch_attention_bwd_kernel<D, KMAX, MODE, SS_ATTB_POLY>;
  SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
  dim3 grid((unsigned)((size_t)heads * max_patches));  // heads fastest: the H CTAs of a patch share rows through L2
  kern<<<grid, kBwThreads, S::kTotal, stream>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dout, lse2, delta, n,
                                                order_row, (const int4*)table, max_patches, heads, scale,
                                                scale * 1.4426950408889634f, (__nv_bfloat16*)dqkv, dkv32);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

template <int D>
static int attention_bwd_all(const void* qkv, const void* out, const void* dout, const float* lse2, int64_t n,
                             const int64_t* order_row, const int32_t* table, int max_patches, int heads, float scale,
                             float* delta, float* dkv32, void* dqkv, cudaStream_t stream) {
  const int C = heads * D;
  dim3 sgrid((unsigned)max_patches, 8);
  attn_shared_rows_kernel<0><<<sgrid, 256, 0, stream>>>((const int4*)table, order_row, C, dkv32, (__nv_bfloat16*)dqkv);
  SS_CHECK_LAUNCH();
  const int64_t nd = n * heads;
  attn_delta_kernel<<<(unsigned)ceil_div64(nd, 256), 256, 0, stream>>>((const __nv_bfloat16*)out, (const __nv_bfloat16*)dout,
                                                                       order_row, n, heads, D, delta);
  SS_CHECK_LAUNCH();
  int rc = launch_attention_bwd<D, 0>(qkv, dout, lse2, delta, n, order_row, table, max_patches, heads, scale, dqkv, dkv32, stream);
  if (rc) return rc;
  rc = launch_attention_bwd<D, 1>(qkv, dout, lse2, delta, n, order_row, table, max_patches, heads, scale, dqkv, dkv32, stream);
  if (rc) return rc;
  attn_shared_rows_kernel<1><<<sgrid, 256, 0, stream>>>((const int4*)table, order_row, C, dkv32, (__nv_bfloat16*)dqkv);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // namespace ss

extern "C" size_t ss_patch_attention_backward_workspace_bytes(int64_t n, int heads, int head_dim) {
  if (n < 0 || heads < 1 || head_dim < 1) return 0;
  // delta [H, n] fp32 + dK | dV scratch [n, 2C] fp32
  return ((size_t)n * heads * sizeof(float) + 255) / 256 * 256 + (size_t)n * 2 * heads * head_dim * sizeof(float);
}

extern "C" int ss_patch_attention_backward(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16,
                                           const float* lse2, const int64_t* order_row, const int32_t* table,
                                           int max_patches, int patch_size, int heads, int head_dim, float scale,
                                           int64_t n, void* dqkv_bf16, void* workspace, size_t workspace_bytes,
                                           void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (max_patches < 0 || heads < 1 || patch_size < 1 || patch_size > 1024 || !(scale > 0.f) || n < 0) return SS_BAD_ARGS;
  if (head_dim != 16 && head_dim != 32 && head_dim != 48) return SS_BAD_ARGS;
  if (n == 0 || max_patches == 0) return SS_OK;
  if ((long long)max_patches * heads > 0x7fffffffLL) return SS_BAD_ARGS;
  if (!qkv_bf16 || !out_bf16 || !dout_bf16 || !lse2 || !order_row || !table || !dqkv_bf16 || !workspace) return SS_BAD_ARGS;
  if (((uintptr_t)qkv_bf16 | (uintptr_t)out_bf16 | (uintptr_t)dout_bf16 | (uintptr_t)dqkv_bf16 | (uintptr_t)workspace) % 16 != 0)
    return SS_BAD_ARGS;
  if (workspace_bytes < ss_patch_attention_backward_workspace_bytes(n, heads, head_dim)) return SS_BAD_ARGS;
  float* delta = (float*)workspace;
  float* dkv32 = (float*)((uint8_t*)workspace + ((size_t)n * heads * sizeof(float) + 255) / 256 * 256);
  switch (head_dim) {
    case 16: return ss::attention_bwd_all<16>(qkv_bf16, out_bf16, dout_bf16, lse2, n, order_row, table, max_patches, heads, scale, delta, dkv32, dqkv_bf16, stream);
    case 32: return ss::attention_bwd_all<32>(qkv_bf16, out_bf16, dout_bf16, lse2, n, order_row, table, max_patches, heads, scale, delta, dkv32, dqkv_bf16, stream);
    default: return ss::attention_bwd_all<48>(qkv_bf16, out_bf16, dout_bf16, lse2, n, order_row, table, max_patches, heads, scale, delta, dkv32, dqkv_bf16, stream);
  }
}
