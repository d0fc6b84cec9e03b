// Language head on the 5th-gen tensor cores + the gather-sum stage of the submanifold convolution.
//
// (1) head_gemm_kernel: logits = feat @ T^T (T = <= 256 text embeddings) with sigmoid + max / argmax (EPI = 1) or
//     probability accumulation (EPI = 2) fused into the TMEM epilogue, so the N x K logits never reach HBM.
//     Replaces (reference): pointcept/engines/hooks/evaluator.py:793-800, pointcept/engines/test.py:335-349.
//     CTA = 128 feature rows x 256 classes.  Warps 0-3 stage the A rows with 16-byte cp.async into a 128B-swizzled
//     K-major tile, warp 4 streams the T tile with TMA, warp 5 issues tcgen05.mma (M=128, N=256, K=16) into TMEM,
//     warps 0-3 then drain TMEM (tcgen05.ld).
// (2) conv_reduce_kernel: out[p, :] = bias + sum_t prod[ypos[t][p], :], the HBM-bound second stage of the xCPE conv
//     (first stage: the gather-GEMMs of conv_gemm2.cu / conv_gemm3.cu).
#include "tc_common.cuh"
#include "conv_reduce.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

constexpr int kGemmThreads = 192;
constexpr int kTileM = 128;
constexpr int kBK = 64;  // bf16 elements per K chunk = one 128-byte swizzle row

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int kABytes = kTileM * kBK * 2;  // 16 KB
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + 1024 /*barriers, indices*/ + 1024 /*alignment slack*/;
};

// Epilogue of the language head (EPI = 1: max/argmax of sigmoid(logits); EPI = 2: probs accumulate).
struct HeadEpi {
  int64_t n_rows;
  int n_classes;
  float threshold;
  float* max_prob;
  int64_t* label;
  float* probs_accum;
  const int64_t* idx;
};

template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(kGemmThreads)
head_gemm_kernel(const __nv_bfloat16* __restrict__ X, const __grid_constant__ CUtensorMap tmap_w, int cin,
                 HeadEpi head) {
  using S = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);
  int32_t* s_rows = (int32_t*)(smem + S::kBarOffset + 256);  // [128] gathered input rows

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably uniform
  const int tile = blockIdx.x;
  const int nk = (cin + kBK - 1) / kBK;

  if (threadIdx.x < kTileM) {  // identity rows, clamped (rows past n_rows are computed but never written)
    const int64_t r = (int64_t)tile * kTileM + threadIdx.x;
    s_rows[threadIdx.x] = (int32_t)(r < head.n_rows ? r : head.n_rows - 1);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&full_bar[s], 128 + 1);  // 128 cp.async producers + 1 TMA expect_tx arrive
      tc::mbar_init(&empty_bar[s], 1);       // one tcgen05.commit
    }
    tc::mbar_init(accum_bar, 1);
    tc::mbar_fence_init();
  }
  if (warp == 4 && lane == 0) tc::tma_prefetch_desc(&tmap_w);
  if (warp == 5) tc::tmem_alloc<BN>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp < 4) {
    // ------------------------------------------------------------------ A producers (gather)
    const int tid = threadIdx.x;  // 0..127
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % STAGES, it = kc / STAGES;
      tc::mbar_wait(&empty_bar[s], (it & 1) ^ 1);
      const uint32_t a_base = tc::smem_u32(smem + s * S::kStageBytes);
      const int k0 = kc * kBK;
      const int kchunks = min(kBK, cin - k0) >> 3;  // valid 16-byte chunks in this K chunk
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int q = i * 128 + tid;
        const int r = q >> 3, c = q & 7;
        if (c < kchunks) {
          const __nv_bfloat16* src = X + (size_t)s_rows[r] * cin + k0 + c * 8;
          tc::cp_async16(a_base + tc::sw128_offset(r, c), src);
        }
      }
      tc::cp_async_mbar_arrive_noinc(&full_bar[s]);
    }
    // ------------------------------------------------------------------ epilogue: TMEM -> bf16 -> global
    tc::mbar_wait(accum_bar, 0);
    tc::tc_fence_after();
    const int row = warp * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int64_t grow = (int64_t)tile * kTileM + row;
    float best = -INFINITY;
    int arg = 0;
    float* prow = nullptr;
    if (EPI == 2 && grow < head.n_rows)
      prow = head.probs_accum + (size_t)(head.idx ? head.idx[grow] : grow) * head.n_classes;
#pragma unroll 1
    for (int j = 0; j < BN / 32; ++j) {
      if (j * 32 >= head.n_classes) break;
      uint32_t v[32];
      tc::tmem_ld32(t_lane + j * 32, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        const int k = j * 32 + u;
        const float lg = __uint_as_float(v[u]);
        if (k < head.n_classes) {
          if (EPI == 1) {
            if (lg > best) { best = lg; arg = k; }
          } else if (prow) {
            prow[k] += 1.f / (1.f + __expf(-lg));
          }
        }
      }
    }
    if (EPI == 1 && grow < head.n_rows) {
      const float pr = 1.f / (1.f + __expf(-best));
      head.max_prob[grow] = pr;
      head.label[grow] = pr < head.threshold ? -1 : arg;
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ B producer (TMA, one lane)
    if (lane == 0) {
      for (int kc = 0; kc < nk; ++kc) {
        const int s = kc % STAGES, it = kc / STAGES;
        tc::mbar_wait(&empty_bar[s], (it & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&full_bar[s], S::kBBytes);
        tc::tma_load_2d(tc::smem_u32(smem + s * S::kStageBytes + S::kABytes), &tmap_w, kc * kBK, 0, &full_bar[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer (whole warp, elected lane per op)
    {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(kTileM, BN);
      const uint64_t d_base = tc::umma_desc_sw128(0);
      const uint32_t s0 = tc::smem_u32(smem);
      for (int kc = 0; kc < nk; ++kc) {
        const int s = kc % STAGES, it = kc / STAGES;
        tc::mbar_wait(&full_bar[s], it & 1);
        tc::tc_fence_after();
        const uint32_t a_addr = (s0 + s * S::kStageBytes) >> 4;
        const uint32_t b_addr = a_addr + (S::kABytes >> 4);
        const int ksteps = min(kBK, cin - kc * kBK) >> 4;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t da = d_base | (uint64_t)((a_addr + 2 * k) & 0x3fff);
          const uint64_t db = d_base | (uint64_t)((b_addr + 2 * k) & 0x3fff);
          tc::umma_bf16_elect(tmem_base, da, db, idesc, (kc | k) ? 1u : 0u);
        }
        tc::umma_commit_elect(&empty_bar[s]);  // frees the stage when these MMAs have read it
      }
      tc::umma_commit_elect(accum_bar);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc::tc_fence_after();
    tc::tmem_dealloc<BN>(tmem_base);
  }
}

// out[p, :] = bias + sum_t prod[ypos[t][p], :]; one warp per output voxel, 16-byte (8 x bf16) lanes, J x 256
// channels per pass.  The voxel's k3 product-row positions are fetched with ONE load (lane = tap), the active ones
// are walked in ascending tap order (fixed summation order) four at a time so four row loads are in flight per lane
// (conv_reduce.cuh).
template <typename TO, int J>
__global__ void __launch_bounds__(256)
conv_reduce_kernel(const __nv_bfloat16* __restrict__ prod, const int32_t* __restrict__ ypos,
                   const float* __restrict__ bias, int64_t n, int k3, int cout, TO* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp0; p < n; p += nwarp) {
    float acc[J][8];
    conv_acc_init<J>(bias, lane, cout, acc);
    for (int t0 = 0; t0 < k3; t0 += 32) {
      const int32_t mypos = (t0 + lane < k3) ? ypos[(size_t)(t0 + lane) * n + p] : -1;
      conv_gather_sum<J, false>(prod, mypos, lane, cout, acc);
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c0 = j * 256 + lane * 8;
      if (c0 >= cout) continue;
      if constexpr (sizeof(TO) == 2) {
        uint4 o;
        o.x = tc::pack_bf16(acc[j][0], acc[j][1]);
        o.y = tc::pack_bf16(acc[j][2], acc[j][3]);
        o.z = tc::pack_bf16(acc[j][4], acc[j][5]);
        o.w = tc::pack_bf16(acc[j][6], acc[j][7]);
        *reinterpret_cast<uint4*>(out + (size_t)p * cout + c0) = o;
      } else {
        float4* o = reinterpret_cast<float4*>(out + (size_t)p * cout + c0);
        o[0] = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
        o[1] = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
      }
    }
  }
}

// The same gather-sum for narrow outputs (cout = 8 * LPV, LPV = 4 / 8 / 16 lanes per voxel): with one warp per voxel only
// cout / 8 lanes carry data and every voxel pays 27 strided position loads of its own.  Here a warp takes 32 / LPV
// CONSECUTIVE voxels: a tap's positions are one coalesced load for the whole warp, and every lane group walks its voxel's
// active taps in ascending order (the summation order of conv_reduce_kernel: bit-identical output).
template <typename TO, int LPV>
__global__ void __launch_bounds__(256, 3)
conv_reduce_narrow_kernel(const __nv_bfloat16* __restrict__ prod, const int32_t* __restrict__ ypos,
                          const float* __restrict__ bias, int64_t n, int k3, TO* __restrict__ out) {
  constexpr int G = 32 / LPV, C = 8 * LPV, KMAX = 27, R = 9;
  const int lane = threadIdx.x & 31, sub = lane % LPV, c0 = sub * 8;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p0 = warp0 * G; p0 < n; p0 += nwarp * G) {
    const int64_t p = p0 + lane / LPV;
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = bias ? bias[c0 + u] : 0.f;
#pragma unroll 1
    for (int t0 = 0; t0 < KMAX; t0 += R) {  // R positions, then R row loads in flight per lane
      int32_t pos[R];
#pragma unroll
      for (int q = 0; q < R; ++q) pos[q] = (t0 + q < k3 && p < n) ? ypos[(size_t)(t0 + q) * n + p] : -1;
      uint4 v[R];
#pragma unroll
      for (int q = 0; q < R; ++q) {
        v[q] = make_uint4(0u, 0u, 0u, 0u);
        if (pos[q] >= 0) v[q] = *reinterpret_cast<const uint4*>(prod + (size_t)pos[q] * C + c0);
      }
#pragma unroll
      for (int q = 0; q < R; ++q) {
        if (pos[q] < 0) continue;  // (adding the zero vector would turn a -0 sum into +0)
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[q]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 f = __bfloat1622float2(h[u]);
          acc[2 * u] += f.x;
          acc[2 * u + 1] += f.y;
        }
      }
    }
    if (p < n) {
      if constexpr (sizeof(TO) == 2) {
        uint4 o;
        o.x = tc::pack_bf16(acc[0], acc[1]);
        o.y = tc::pack_bf16(acc[2], acc[3]);
        o.z = tc::pack_bf16(acc[4], acc[5]);
        o.w = tc::pack_bf16(acc[6], acc[7]);
        *reinterpret_cast<uint4*>(out + (size_t)p * C + c0) = o;
      } else {
        float4* o = reinterpret_cast<float4*>(out + (size_t)p * C + c0);
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
    }
  }
}

// The gather-sum fused with what follows the xCPE conv in a Block (point_transformer_v3m1_base.py:318-326 with the conv's
// Linear folded into the taps): z = bias + sum_t prod[ypos[t][p], :] stays in registers (fp32, never rounded to bf16),
// y = res + LN0(z) is written as the new fp32 residual stream and LN1(y) as the bf16 operand of the qkv GEMM.  One warp
// per voxel, as in conv_reduce_kernel; the two LayerNorm reductions are warp shuffles.  Saves the write + read of z and
// one launch per Block.  (For C >= 256 the same body runs inside the gather-GEMM itself: conv_gemm3.cu.)
template <int J>
__global__ void __launch_bounds__(256)
conv_reduce_add_ln_kernel(const __nv_bfloat16* __restrict__ prod, const int32_t* __restrict__ ypos,
                          const float* __restrict__ bias, const float* res, const float* __restrict__ g0,
                          const float* __restrict__ b0, const float* __restrict__ g1, const float* __restrict__ b1, float eps,
                          int64_t n, int k3, int C, float* res_out, __nv_bfloat16* __restrict__ norm_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float invC = 1.f / (float)C;
  for (int64_t p = warp0; p < n; p += nwarp) {
    float acc[J][8];
    conv_acc_init<J>(bias, lane, C, acc);
    for (int t0 = 0; t0 < k3; t0 += 32) {
      const int32_t mypos = (t0 + lane < k3) ? ypos[(size_t)(t0 + lane) * n + p] : -1;
      conv_gather_sum<J, false>(prod, mypos, lane, C, acc);
    }
    conv_ln_res_ln_store<J>(acc, p, lane, C, invC, eps, res, g0, b0, g1, b1, res_out, norm_out);
  }
}

template <int EPI>
static int launch_head(const void* feat, const CUtensorMap& tmap, int64_t n, int channels, const HeadEpi& head,
                       cudaStream_t stream) {
  using S = GemmSmem<256, 2>;
  auto kern = head_gemm_kernel<256, 2, EPI>;
  SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
  dim3 grid((unsigned)ceil_div64(n, kTileM), 1);
  kern<<<grid, kGemmThreads, S::kTotal, stream>>>((const __nv_bfloat16*)feat, tmap, channels, head);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // namespace ss

extern "C" {

int ss_lang_head_tc(const void* feat_bf16, const void* text_bf16, int64_t n, int channels, int n_classes,
                    float threshold, int mode, const int64_t* idx, float* max_prob, int64_t* label, float* probs_accum,
                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || channels < 16 || channels % 16 != 0 || n_classes < 1 || n_classes > 256 || (mode != 0 && mode != 1))
    return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!feat_bf16 || !text_bf16 || ((uintptr_t)feat_bf16 | (uintptr_t)text_bf16) % 16 != 0) return SS_BAD_ARGS;
  if (mode == 0 && (!max_prob || !label)) return SS_BAD_ARGS;
  if (mode == 1 && !probs_accum) return SS_BAD_ARGS;
  CUtensorMap tmap;  // text [n_classes, channels]: rows past n_classes read as zero (TMA OOB fill)
  int rc = ss::make_tmap_bf16_2d(&tmap, text_bf16, (uint64_t)n_classes, (uint64_t)channels, 256, ss::kBK);
  if (rc) return rc;
  ss::HeadEpi head{n, n_classes, threshold, max_prob, label, probs_accum, idx};
  return mode == 0 ? ss::launch_head<1>(feat_bf16, tmap, n, channels, head, stream)
                   : ss::launch_head<2>(feat_bf16, tmap, n, channels, head, stream);
}

int ss_subm_conv_reduce(const void* prod_bf16, const int32_t* ypos, const float* bias, int64_t n, int k3, int cout,
                        void* out, int out_is_bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || k3 < 1 || cout < 8 || cout % 8 != 0) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!prod_bf16 || !ypos || !out) return SS_BAD_ARGS;
  if (cout > 1024) return SS_BAD_ARGS;
  if (k3 <= 27 && (cout == 32 || cout == 64 || cout == 128)) {  // several voxels per warp
    const int lpv = cout / 8;
    const int nblocks = (int)ss::imin64(ss::ceil_div64(n, 8 * (32 / lpv)), 32 * ss::kNumSMs);
#define SS_REDN_(TO, LPV)                                                                                              \
  ss::conv_reduce_narrow_kernel<TO, LPV><<<nblocks, 256, 0, stream>>>((const __nv_bfloat16*)prod_bf16, ypos, bias, n, k3, \
                                                                      (TO*)out)
#define SS_REDN_L_(TO)                     \
  do {                                     \
    if (lpv == 4) SS_REDN_(TO, 4);         \
    else if (lpv == 8) SS_REDN_(TO, 8);    \
    else SS_REDN_(TO, 16);                 \
  } while (0)
    if (out_is_bf16) SS_REDN_L_(__nv_bfloat16);
    else SS_REDN_L_(float);
#undef SS_REDN_L_
#undef SS_REDN_
    SS_CHECK_LAUNCH();
    return SS_OK;
  }
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, 8), 32 * ss::kNumSMs);
  const int j = (cout + 255) / 256;
#define SS_RED_(TO, J) \
  ss::conv_reduce_kernel<TO, J><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)prod_bf16, ypos, bias, n, k3, cout, (TO*)out)
#define SS_RED_J_(TO)                 \
  do {                                \
    if (j == 1) SS_RED_(TO, 1);       \
    else if (j == 2) SS_RED_(TO, 2);  \
    else if (j == 3) SS_RED_(TO, 3);  \
    else SS_RED_(TO, 4);              \
  } while (0)
  if (out_is_bf16) SS_RED_J_(__nv_bfloat16);
  else SS_RED_J_(float);
#undef SS_RED_J_
#undef SS_RED_
  SS_CHECK_LAUNCH();
  return SS_OK;
}

int ss_subm_conv_reduce_add_ln(const void* prod_bf16, const int32_t* ypos, const float* bias, const float* res,
                               const float* g0, const float* b0, const float* g1, const float* b1, float eps, int64_t n, int k3,
                               int channels, float* res_out, void* norm_out_bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || k3 < 1 || channels < 8 || channels % 8 != 0 || channels > 1024) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!prod_bf16 || !ypos || !res || !g0 || !b0 || !g1 || !b1 || !res_out || !norm_out_bf16) return SS_BAD_ARGS;
  if (((uintptr_t)prod_bf16 | (uintptr_t)res | (uintptr_t)res_out | (uintptr_t)norm_out_bf16 | (uintptr_t)bias | (uintptr_t)g0 |
       (uintptr_t)b0 | (uintptr_t)g1 | (uintptr_t)b1) % 16 != 0)
    return SS_BAD_ARGS;
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, 8), 32 * ss::kNumSMs);
  const int j = (channels + 255) / 256;
#define SS_RLN_(J)                                                                                                        \
  ss::conv_reduce_add_ln_kernel<J><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)prod_bf16, ypos, bias, res, g0, b0, g1, b1, \
                                                               eps, n, k3, channels, res_out, (__nv_bfloat16*)norm_out_bf16)
  if (j == 1) SS_RLN_(1);
  else if (j == 2) SS_RLN_(2);
  else if (j == 3) SS_RLN_(3);
  else SS_RLN_(4);
#undef SS_RLN_
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // extern "C"
