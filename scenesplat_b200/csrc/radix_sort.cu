// The digit pass of the radix sort and its launch sequence (design notes: radix_sort.cuh).
#include "radix_sort.cuh"

namespace ss {

// ---------------------------------------------------------------------------------------------------------------
template <int BITS, typename KI>
constexpr size_t radix_pass_smem() {
  constexpr size_t tables = (size_t)kSortWarps * (1 << BITS) * (4 + 2);  // lane masks (u32) + counts (u16) per warp
  constexpr size_t stage = (size_t)kSortTile * (sizeof(KI) + 4);
  return (tables > stage ? tables : stage) + 3 * (size_t)(1 << BITS) * 4;
}

template <typename KI, typename KO, int BITS, int OUT>
__global__ void __launch_bounds__(kSortThreads)
radix_pass_kernel(const KI* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, KO* __restrict__ keys_out,
                  uint32_t* __restrict__ vals_out, int64_t* __restrict__ order_out, int64_t* __restrict__ inverse_out,
                  const uint32_t* __restrict__ ghist,  // this pass: [row * hist_row_stride + d]
                  uint32_t* __restrict__ ghist_next,   // this pass's slot too; the next pass's follows (nullptr: last)
                  uint32_t* __restrict__ counters,     // [rows] tickets of this pass
                  uint32_t* __restrict__ status,       // [rows][tiles][BINS] of this pass
                  int n, int tiles, int shift, size_t row_stride, size_t hist_row_stride) {
  constexpr int BINS = 1 << BITS;
  constexpr int PER = BINS > kSortThreads ? BINS / kSortThreads : 1;  // digits per thread in the per-digit phase
  extern __shared__ __align__(16) unsigned char s_raw[];
  // region A: per-warp lane masks [warps][BINS] u32, per-warp counts [warps][BINS] u16; later the staged tile
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_raw);
  uint16_t* s_whist = reinterpret_cast<uint16_t*>(s_raw + (size_t)kSortWarps * BINS * 4);
  KI* s_keys = reinterpret_cast<KI*>(s_raw);
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_raw + (size_t)kSortTile * sizeof(KI));
  constexpr size_t kRegionA = radix_pass_smem<BITS, KI>() - 3 * (size_t)BINS * 4;
  uint32_t* s_digit_start = reinterpret_cast<uint32_t*>(s_raw + kRegionA);  // start of the digit's run in the staged tile
  uint32_t* s_global_base = s_digit_start + BINS;                            // global destination of the run's first key
  uint32_t* s_next = s_global_base + BINS;                                   // histogram of the next pass's digit
  __shared__ uint32_t s_scan[66];
  __shared__ int s_tile;

  const int row = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  keys_in += (size_t)row * row_stride;
  if (vals_in) vals_in += (size_t)row * row_stride;
  if (OUT == kPairs) {
    keys_out += (size_t)row * row_stride;
    vals_out += (size_t)row * row_stride;
  } else {
    order_out += (size_t)row * row_stride;
    if (OUT == kFinal) inverse_out += (size_t)row * row_stride;
    if (keys_out != nullptr) keys_out += (size_t)row * row_stride;  // kFinalPairs: always; kFinal: optional
  }
  ghist += (size_t)row * hist_row_stride;
  const bool count_next = OUT == kPairs && ghist_next != nullptr;
  if (count_next) ghist_next += (size_t)row * hist_row_stride + BINS;

  if (tid == 0) s_tile = (int)atomicAdd(&counters[row], 1u);
  {
    uint32_t* z = reinterpret_cast<uint32_t*>(s_raw);
    for (int i = tid; i < kSortWarps * BINS * 6 / 4; i += kSortThreads) z[i] = 0u;
    for (int i = tid; i < BINS; i += kSortThreads) s_next[i] = 0u;
  }
  __syncthreads();
  const int tile = s_tile;
  const int base = tile * kSortTile;
  const int tile_n = min(kSortTile, n - base);

  // ---- load (warp-contiguous chunks keep the order stable), rank inside the warp, count the next digit
  KI key[kSortItems];
  uint32_t val[kSortItems];
  uint32_t rank[kSortItems];
  const int wbase = warp * (32 * kSortItems);
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    const int li = wbase + it * 32 + lane;
    const bool ok = li < tile_n;
    key[it] = ok ? keys_in[base + li] : (KI)0;
    val[it] = ok ? (vals_in ? vals_in[base + li] : (uint32_t)(base + li)) : 0u;
  }
  uint16_t* my_whist = s_whist + warp * BINS;
  uint32_t* my_mask = s_mask + warp * BINS;
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    const bool ok = wbase + it * 32 + lane < tile_n;
    rank[it] = warp_rank(my_whist, my_mask, (uint32_t)(key[it] >> shift) & (BINS - 1), ok);
    if (count_next && ok) atomicAdd(&s_next[(uint32_t)(key[it] >> (shift + BITS)) & (BINS - 1)], 1u);
  }
  __syncthreads();

  // ---- per digit: prefix over warps, tile aggregate, global bin start, look-back
  {
    uint32_t run[PER], gcount[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int d = tid * PER + j;
      run[j] = 0;
      gcount[j] = 0;
      if (d < BINS) {
        uint32_t r = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
          const uint32_t c = s_whist[w * BINS + d];
          s_whist[w * BINS + d] = (uint16_t)r;
          r += c;
        }
        run[j] = r;
        gcount[j] = ghist[d];
        if (count_next) {
          const uint32_t c = s_next[d];
          if (c) atomicAdd(&ghist_next[d], c);
        }
      }
    }
    uint32_t rsum = 0, gsum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      rsum += run[j];
      gsum += gcount[j];
    }
    uint32_t rex, gex, rtot, gtot;
    block_exclusive_scan2(rsum, gsum, s_scan, rex, gex, rtot, gtot);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int d = tid * PER + j;
      if (d < BINS) {
        s_digit_start[d] = rex;
        const uint32_t excl =
            lookback_exclusive_batched(status + ((size_t)row * tiles) * BINS + d, BINS, tile, run[j]);
        s_global_base[d] = gex + excl;
      }
      rex += run[j];
      gex += gcount[j];
    }
  }
  __syncthreads();

  // ---- position of every key in the digit-sorted tile (registers), then stage over the tables
  uint32_t pos[kSortItems];
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    const uint32_t d = (uint32_t)(key[it] >> shift) & (BINS - 1);
    pos[it] = s_digit_start[d] + my_whist[d] + rank[it];
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    if (wbase + it * 32 + lane < tile_n) {
      s_keys[pos[it]] = key[it];
      s_vals[pos[it]] = val[it];
    }
  }
  __syncthreads();

  // ---- scatter in runs
  for (int i = tid; i < tile_n; i += kSortThreads) {
    const KI k = s_keys[i];
    const uint32_t v = s_vals[i];
    const uint32_t d = (uint32_t)(k >> shift) & (BINS - 1);
    const uint32_t dst = s_global_base[d] + ((uint32_t)i - s_digit_start[d]);
    if (OUT == kPairs) {
      keys_out[dst] = (KO)k;
      vals_out[dst] = v;
    } else if (OUT == kFinal) {
      order_out[dst] = (int64_t)v;
      inverse_out[v] = (int64_t)dst;
      if (keys_out != nullptr) keys_out[dst] = (KO)k;
    } else {
      keys_out[dst] = (KO)k;
      order_out[dst] = (int64_t)v;
    }
  }
}

template <typename KI, typename KO, int BITS, int OUT>
inline cudaError_t radix_pass_launch(dim3 grid, cudaStream_t stream, const void* kin, const uint32_t* vin, void* kout,
                                     uint32_t* vout, int64_t* order_out, int64_t* inverse_out, const uint32_t* ghist,
                                     uint32_t* ghist_next, uint32_t* cnt, uint32_t* st, int n, int tiles, int shift,
                                     size_t row_stride, size_t hist_row_stride) {
  auto kern = radix_pass_kernel<KI, KO, BITS, OUT>;
  constexpr size_t smem = radix_pass_smem<BITS, KI>();
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  kern<<<grid, kSortThreads, smem, stream>>>((const KI*)kin, vin, (KO*)kout, vout, order_out, inverse_out, ghist,
                                             ghist_next, cnt, st, n, tiles, shift, row_stride, hist_row_stride);
  return cudaGetLastError();
}

template <int BITS>
inline cudaError_t radix_pass_dispatch(bool in64, bool out64, int out_mode, dim3 grid, cudaStream_t stream,
                                       const void* kin, const uint32_t* vin, void* kout, uint32_t* vout,
                                       int64_t* order_out, int64_t* inverse_out, const uint32_t* ghist,
                                       uint32_t* ghist_next, uint32_t* cnt, uint32_t* st, int n, int tiles, int shift,
                                       size_t row_stride, size_t hist_row_stride) {
#define SS_RP_(KI, KO, OUT)                                                                                           \
  return radix_pass_launch<KI, KO, BITS, OUT>(grid, stream, kin, vin, kout, vout, order_out, inverse_out, ghist,      \
                                              ghist_next, cnt, st, n, tiles, shift, row_stride, hist_row_stride)
  if (out_mode == kPairs) {
    if (in64 && out64) SS_RP_(uint64_t, uint64_t, kPairs);
    if (in64) SS_RP_(uint64_t, uint32_t, kPairs);
    SS_RP_(uint32_t, uint32_t, kPairs);
  }
  if (out_mode == kFinal) {
    if (in64) SS_RP_(uint64_t, uint64_t, kFinal);
    SS_RP_(uint32_t, uint64_t, kFinal);
  }
  if (in64) SS_RP_(uint64_t, uint64_t, kFinalPairs);
  SS_RP_(uint32_t, uint64_t, kFinalPairs);
#undef SS_RP_
}

int radix_sort_run(const RadixPlan& p, char* ws, const uint64_t* keys0, int final_mode, int64_t* order_out,
                          int64_t* inverse_out, uint64_t* sorted_keys, cudaStream_t stream) {
  if (p.n <= 0) return SS_OK;
  uint32_t* ghist = (uint32_t*)(ws + p.off_hist);
  uint32_t* counters = (uint32_t*)(ws + p.off_counter);
  uint32_t* status = (uint32_t*)(ws + p.off_status);
  void* kbuf[2] = {(void*)(ws + p.off_keys[0]), (void*)(ws + p.off_keys[1])};
  uint32_t* vbuf[2] = {(uint32_t*)(ws + p.off_vals[0]), (uint32_t*)(ws + p.off_vals[1])};
  dim3 grid(p.tiles, p.rows);
  const void* kin = keys0;
  const uint32_t* vin = nullptr;
  bool in64 = true;
  const int bins = p.bins();
  for (int pass = 0; pass < p.passes; ++pass) {
    uint32_t* cnt = counters + (size_t)pass * p.rows;
    uint32_t* st = status + (size_t)pass * p.rows * p.tiles * bins;
    const uint32_t* gh = ghist + (size_t)pass * bins;
    const int shift = pass * p.bits;
    const bool last = pass == p.passes - 1;
    const int mode = last ? final_mode : kPairs;
    const int o = pass & 1;
    void* kout = last ? (void*)sorted_keys : kbuf[o];
    const bool out64 = last ? true : !p.key32;
    uint32_t* gnext = last ? nullptr : ghist + (size_t)pass * bins;
    cudaError_t e;
#define SS_RD_(B)                                                                                                  \
  e = radix_pass_dispatch<B>(in64, out64, mode, grid, stream, kin, vin, kout, last ? nullptr : vbuf[o], order_out, \
                             inverse_out, gh, gnext, cnt, st, p.n, p.tiles, shift, (size_t)p.n, p.hist_row_stride())
    if (p.bits == 8) SS_RD_(8);
    else if (p.bits == 9) SS_RD_(9);
    else SS_RD_(10);
#undef SS_RD_
    ++g_launch_count;
    if (e != cudaSuccess) return (int)e;
    kin = kbuf[o];
    vin = vbuf[o];
    in64 = !p.key32;
  }
  return SS_OK;
}

}  // namespace ss
