// Shared device helpers for the scenesplat_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define SS_OK 0
#define SS_BAD_ARGS (-1)

// Every kernel launch site is followed by SS_CHECK_LAUNCH(): it also counts the launch (ss_launch_count()).
namespace ss { extern unsigned long long g_launch_count; }
#define SS_CHECK_LAUNCH()                         \
  do {                                            \
    ++ss::g_launch_count;                         \
    cudaError_t e__ = cudaGetLastError();         \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

#define SS_CUDA(x)                                \
  do {                                            \
    cudaError_t e__ = (x);                        \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

namespace ss {

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// GELU (exact / erf form, nn.GELU() default) without erff: 0.5 x erfc(|x|/sqrt2) from Abramowitz-Stegun 7.1.26
// (|error of erf| <= 1.5e-7), evaluated in the complementary form so the negative tail has no cancellation:
//   z = |x|/sqrt2, t = 1/(1 + p z), g = 0.5 |x| (a1 t + .. + a5 t^5) exp(-z^2),  GELU(x) = max(x, 0) - g.
// 15 instructions incl. MUFU.RCP + MUFU.EX2; max abs error 3.3e-7 over [-12, 12] (tools checked against float64;
// torch's own fp32 erf-GELU is at 1.2e-6), so the elementwise kernels stay HBM-bound instead of erff-bound.
__device__ __forceinline__ float gelu_fast(float x) {
  const float ax = fabsf(x);
  const float z = ax * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  p *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  return fmaxf(x, 0.f) - ax * (p * e);
}

// ---------------------------------------------------------------------------------------------
// Warp / block scans (sum), 32-bit.
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane_id() >= (unsigned)o) v += t;
  }
  return v;
}

// Exclusive block scan over blockDim.x (multiple of 32, <= 1024) threads. `smem` needs 33 words.
// Returns the exclusive prefix of `v`; `total` receives the block sum.  Contains __syncthreads.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem, uint32_t& total) {
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  uint32_t inc = warp_inclusive_scan(v);
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < nwarp ? smem[lane] : 0u;
    uint32_t winc = warp_inclusive_scan(w);
    smem[lane] = winc - w;
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  uint32_t res = smem[warp] + inc - v;
  total = smem[32];
  __syncthreads();
  return res;
}

// ---------------------------------------------------------------------------------------------
// Decoupled look-back over tiles for one 32-bit value per (tile, slot).
// status word: bit31 = inclusive prefix available, bit30 = tile aggregate available, low 30 bits value.
constexpr uint32_t kFlagInclusive = 0x80000000u;
constexpr uint32_t kFlagAggregate = 0x40000000u;
constexpr uint32_t kValueMask = 0x3fffffffu;

// Publishes this tile's aggregate, walks back over earlier tiles and returns the exclusive prefix
// of all earlier tiles; then publishes the inclusive prefix.  `status` points at slot 0 of tile 0,
// `stride` = words between consecutive tiles for the same slot.
__device__ __forceinline__ uint32_t lookback_exclusive(uint32_t* status, int stride, int tile, uint32_t aggregate) {
  uint32_t* mine = status + (size_t)tile * stride;
  if (tile == 0) {
    st_volatile_u32(mine, kFlagInclusive | aggregate);
    return 0u;
  }
  st_volatile_u32(mine, kFlagAggregate | aggregate);
  uint32_t excl = 0u;
  for (int t = tile - 1; t >= 0; --t) {
    const uint32_t* p = status + (size_t)t * stride;
    uint32_t v;
    do {
      v = ld_volatile_u32(p);
    } while ((v & (kFlagInclusive | kFlagAggregate)) == 0u);
    excl += v & kValueMask;
    if (v & kFlagInclusive) break;
  }
  st_volatile_u32(mine, kFlagInclusive | (excl + aggregate));
  return excl;
}

// ---------------------------------------------------------------------------------------------
// Space-filling-curve codes (bit-exact restatement targets: oracle/serialization.py).
__device__ __forceinline__ uint64_t spread3(uint64_t x) {  // 16 low bits -> every third bit
  x &= 0xffffull;
  x = (x | (x << 32)) & 0x001f00000000ffffull;
  x = (x | (x << 16)) & 0x001f0000ff0000ffull;
  x = (x | (x << 8)) & 0x100f00f00f00f00full;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

// bit i of x -> 3i+2, y -> 3i+1, z -> 3i  (reference z_order.py:42-49)
__device__ __forceinline__ uint64_t z_key(uint32_t x, uint32_t y, uint32_t z, int depth) {
  const uint32_t m = (1u << depth) - 1u;
  return (spread3(x & m) << 2) | (spread3(y & m) << 1) | spread3(z & m);
}

// Skilling transpose + interleave + Gray decode (reference hilbert.py:156-181)
__device__ __forceinline__ uint64_t hilbert_key(uint32_t x, uint32_t y, uint32_t z, int depth) {
  const uint32_t m = (1u << depth) - 1u;
  uint32_t X0 = x & m, X1 = y & m, X2 = z & m;
  for (uint32_t Q = 1u << (depth - 1); Q >= 1u; Q >>= 1) {
    const uint32_t P = Q - 1u;
    // dim 0
    if (X0 & Q) X0 ^= P;
    // dim 1
    if (X1 & Q) {
      X0 ^= P;
    } else {
      uint32_t t = (X0 ^ X1) & P;
      X0 ^= t;
      X1 ^= t;
    }
    // dim 2
    if (X2 & Q) {
      X0 ^= P;
    } else {
      uint32_t t = (X0 ^ X2) & P;
      X0 ^= t;
      X2 ^= t;
    }
  }
  uint64_t h = (spread3(X0) << 2) | (spread3(X1) << 1) | spread3(X2);
  h ^= h >> 1;
  h ^= h >> 2;
  h ^= h >> 4;
  h ^= h >> 8;
  h ^= h >> 16;
  h ^= h >> 32;
  return h;
}

// 32-bit forms for depth <= 10 (30-bit codes): the same bit patterns with a third of the instructions.
__device__ __forceinline__ uint32_t spread3_10(uint32_t x) {  // 10 low bits -> every third bit
  x &= 0x3ffu;
  x = (x | (x << 16)) & 0x030000ffu;
  x = (x | (x << 8)) & 0x0300f00fu;
  x = (x | (x << 4)) & 0x030c30c3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}

__device__ __forceinline__ uint32_t z_key32(uint32_t x, uint32_t y, uint32_t z, int depth) {
  const uint32_t m = (1u << depth) - 1u;
  return (spread3_10(x & m) << 2) | (spread3_10(y & m) << 1) | spread3_10(z & m);
}

__device__ __forceinline__ uint32_t hilbert_key32(uint32_t x, uint32_t y, uint32_t z, int depth) {
  const uint32_t m = (1u << depth) - 1u;
  uint32_t X0 = x & m, X1 = y & m, X2 = z & m;
  for (uint32_t Q = 1u << (depth - 1); Q >= 1u; Q >>= 1) {
    const uint32_t P = Q - 1u;
    if (X0 & Q) X0 ^= P;
    if (X1 & Q) {
      X0 ^= P;
    } else {
      const uint32_t t = (X0 ^ X1) & P;
      X0 ^= t;
      X1 ^= t;
    }
    if (X2 & Q) {
      X0 ^= P;
    } else {
      const uint32_t t = (X0 ^ X2) & P;
      X0 ^= t;
      X2 ^= t;
    }
  }
  uint32_t h = (spread3_10(X0) << 2) | (spread3_10(X1) << 1) | spread3_10(X2);
  h ^= h >> 1;
  h ^= h >> 2;
  h ^= h >> 4;
  h ^= h >> 8;
  h ^= h >> 16;
  return h;
}

// The same Hilbert encoder as an MSB-first finite-state machine over the octants of the z-order key (three key bits per
// level in, three out): state = the signed permutation the earlier levels of Skilling's transpose have accumulated on
// the lower bits + the running parity of the Gray->binary prefix XOR.  ~7 instructions per level instead of ~25 + the
// interleave + the XOR cascade.  hilbert(x, y, z) = fsm(z_key(x, y, z)); hilbert-trans = fsm(z_key(y, x, z)).
// Table derived from, and verified against, the pinned restatement of hilbert.py by tools/gen_hilbert_fsm.py.
// 48 states x 8 octants, entry = next_state << 3 | output bits; generated by tools/gen_hilbert_fsm.py
constexpr int kHilbertStates = 48;
static __device__ const uint16_t kHilbertFsm[kHilbertStates * 8] = {
    8, 17, 27, 2, 39, 46, 52, 61,
    64, 79, 81, 94, 99, 108, 10, 117,
    48, 57, 127, 134, 139, 106, 12, 21,
    150, 153, 29, 162, 87, 88, 172, 59,
    180, 187, 37, 194, 207, 208, 222, 225,
    236, 189, 35, 42, 31, 6, 240, 249,
    216, 231, 259, 4, 265, 278, 50, 285,
    8, 17, 27, 2, 39, 46, 52, 61,
    0, 291, 255, 220, 281, 66, 302, 269,
    308, 263, 227, 312, 77, 54, 274, 121,
    104, 203, 113, 218, 191, 68, 198, 85,
    212, 103, 229, 14, 75, 176, 90, 33,
    174, 287, 321, 296, 101, 332, 338, 115,
    48, 57, 127, 134, 139, 106, 12, 21,
    64, 79, 81, 94, 99, 108, 10, 117,
    316, 251, 295, 304, 125, 298, 70, 73,
    244, 253, 143, 110, 123, 130, 232, 185,
    350, 353, 271, 272, 141, 330, 340, 19,
    166, 205, 361, 146, 63, 348, 128, 83,
    210, 25, 157, 246, 355, 168, 92, 327,
    150, 153, 29, 162, 87, 88, 172, 59,
    342, 119, 261, 164, 369, 192, 170, 283,
    178, 379, 373, 196, 257, 160, 318, 367,
    236, 189, 35, 42, 31, 6, 240, 249,
    180, 187, 37, 194, 207, 208, 222, 225,
    166, 205, 361, 146, 63, 348, 128, 83,
    210, 25, 157, 246, 355, 168, 92, 327,
    104, 203, 113, 218, 191, 68, 198, 85,
    212, 103, 229, 14, 75, 176, 90, 33,
    234, 381, 371, 44, 289, 310, 144, 159,
    242, 365, 201, 214, 323, 132, 344, 359,
    244, 253, 143, 110, 123, 130, 232, 185,
    342, 119, 261, 164, 369, 192, 170, 283,
    0, 291, 255, 220, 281, 66, 302, 269,
    308, 263, 227, 312, 77, 54, 274, 121,
    216, 231, 259, 4, 265, 278, 50, 285,
    334, 293, 23, 148, 377, 346, 40, 267,
    316, 251, 295, 304, 125, 298, 70, 73,
    306, 137, 155, 336, 357, 238, 276, 375,
    314, 363, 97, 328, 325, 300, 182, 383,
    314, 363, 97, 328, 325, 300, 182, 383,
    350, 353, 271, 272, 141, 330, 340, 19,
    174, 287, 321, 296, 101, 332, 338, 115,
    334, 293, 23, 148, 377, 346, 40, 267,
    306, 137, 155, 336, 357, 238, 276, 375,
    242, 365, 201, 214, 323, 132, 344, 359,
    178, 379, 373, 196, 257, 160, 318, 367,
    234, 381, 371, 44, 289, 310, 144, 159,
};

// Copies the table to shared memory (caller synchronises).  `s_fsm` needs kHilbertStates * 8 entries.
__device__ __forceinline__ void hilbert_fsm_load(uint16_t* s_fsm) {
  for (int i = threadIdx.x; i < kHilbertStates * 8; i += blockDim.x) s_fsm[i] = kHilbertFsm[i];
}

template <typename KeyT>  // uint32_t for depth <= 10, uint64_t otherwise
__device__ __forceinline__ KeyT hilbert_from_zkey(KeyT zk, int depth, const uint16_t* __restrict__ s_fsm) {
  uint32_t e = 0;
  KeyT h = 0;
  for (int sh = 3 * (depth - 1); sh >= 0; sh -= 3) {
    e = s_fsm[(e & 0xfff8u) | ((uint32_t)(zk >> sh) & 7u)];
    h = (h << 3) | (KeyT)(e & 7u);
  }
  return h;
}

// order ids: 0 = z, 1 = z-trans, 2 = hilbert, 3 = hilbert-trans (reference serialization/default.py:8-24)
__device__ __forceinline__ uint64_t sfc_key(int order_id, uint32_t x, uint32_t y, uint32_t z, int depth) {
  switch (order_id) {
    case 0: return z_key(x, y, z, depth);
    case 1: return z_key(y, x, z, depth);
    case 2: return hilbert_key(x, y, z, depth);
    default: return hilbert_key(y, x, z, depth);
  }
}

__device__ __forceinline__ uint32_t sfc_key32(int order_id, uint32_t x, uint32_t y, uint32_t z, int depth) {  // depth <= 10
  switch (order_id) {
    case 0: return z_key32(x, y, z, depth);
    case 1: return z_key32(y, x, z, depth);
    case 2: return hilbert_key32(x, y, z, depth);
    default: return hilbert_key32(y, x, z, depth);
  }
}

}  // namespace ss
