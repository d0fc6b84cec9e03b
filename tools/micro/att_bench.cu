// Stand-alone A/B driver of the tcgen05 patch-attention kernels (developer tool): the shipped kernel
// (scenesplat_b200/csrc/attention_tc.cu) against the round-2 experimental kernel (tools/micro/attention_exp.cu) at a
// given shape, with the largest output difference between the two.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo \
//        "-DSS_ATT_ENTRY(name)=name##_exp" [-DSS_ATT_TRACE2] -o tools/micro/att_bench tools/micro/att_bench.cu \
//        scenesplat_b200/csrc/attention_tc.cu tools/micro/attention_exp.cu
//   att_bench n H d [reps] [order: 0 random | 1 identity]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <algorithm>
#include <numeric>
#include "../../include/scenesplat_b200.h"
#include <cuda_bf16.h>
#include <cmath>
extern "C" int ss_patch_attention_exp(const void*, const int64_t*, const int32_t*, int, int, int, int, float, void*, void*);
namespace ss { unsigned long long g_launch_count = 0; }
#ifdef SS_ATT_TRACE2
extern "C" void ss_att_trace_read(long long* host, int* slots, int* ctas);
#endif

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 299277, H = argc > 2 ? atoi(argv[2]) : 16, d = argc > 3 ? atoi(argv[3]) : 48;
  const int reps = argc > 4 ? atoi(argv[4]) : 5, ident = argc > 5 ? atoi(argv[5]) : 0;
  const int K = 1024, C = H * d;
  std::mt19937 rng(0);
  std::vector<__nv_bfloat16> hq((size_t)n * 3 * C);
  std::normal_distribution<float> nd(0.f, 1.f);
  for (size_t i = 0; i < hq.size(); ++i) hq[i] = __float2bfloat16(nd(rng));
  std::vector<int64_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  if (!ident) std::shuffle(order.begin(), order.end(), rng);
  // argv[7]: number of batch elements the n tokens are split into (each with its own patches and last-patch window rule)
  const int nbatch = argc > 7 ? atoi(argv[7]) : 1;
  std::vector<int> table;
  for (int b = 0; b < nbatch; ++b) {
    const int b0 = (int)((long long)n * b / nbatch), b1 = (int)((long long)n * (b + 1) / nbatch), nb = b1 - b0;
    const int npb = (nb + K - 1) / K;
    for (int p = 0; p < npb; ++p) {
      int qb = b0 + p * K, qe = std::min(b1, qb + K), kb = qb, ke = qe;
      if (p == npb - 1 && nb > K) kb = b1 - K;  // window rule of the last patch
      table.push_back(qb); table.push_back(qe); table.push_back(kb); table.push_back(ke);
    }
  }
  for (int z = 0; z < 2; ++z) for (int u = 0; u < 4; ++u) table.push_back(0);  // unused entries (n_q = 0), as the device table has
  const int np = (int)table.size() / 4;
  __nv_bfloat16 *dq, *dout; int64_t* dord; int* dtab;
  cudaMalloc(&dq, hq.size() * 2); cudaMalloc(&dout, (size_t)n * C * 2); cudaMalloc(&dord, n * 8); cudaMalloc(&dtab, table.size() * 4);
  cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dord, order.data(), n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dtab, table.data(), table.size() * 4, cudaMemcpyHostToDevice);
  char* flush; cudaMalloc(&flush, 256 << 20);
  float scale = 1.f / sqrtf((float)d);
  __nv_bfloat16* dout1; cudaMalloc(&dout1, (size_t)n * C * 2);
  typedef int (*fn_t)(const void*, const int64_t*, const int32_t*, int, int, int, int, float, void*, void*);
  struct { const char* name; fn_t fn; __nv_bfloat16* out; } impl[2] = {{"shipped (attention_tc.cu: 16 softmax warps, 128-key steps)", ss_patch_attention, dout1},
                                                                      {"experimental (attention_exp.cu: 64-key steps, 2 S/P buffers)", ss_patch_attention_exp, dout}};
  const int only = argc > 6 ? atoi(argv[6]) : -1;  // 0: shipped only, 1: experimental only (for ncu)
  for (int v = 0; v < 2; ++v) {
    if (only >= 0 && only != v) continue;
    for (int i = 0; i < 2; ++i) {
      int rc = impl[v].fn(dq, dord, dtab, np, K, H, d, scale, impl[v].out, 0);
      if (rc) { printf("launch rc=%d\n", rc); return 1; }
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float tot = 0;
    for (int i = 0; i < reps; ++i) {
      cudaMemsetAsync(flush, i, 256 << 20, 0);
      cudaEventRecord(e0);
      impl[v].fn(dq, dord, dtab, np, K, H, d, scale, impl[v].out, 0);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); tot += ms;
    }
    const double ms = tot / reps, exps = (double)n * K * H, flops = 4.0 * K * C * (double)n;
    printf("n=%d H=%d d=%d order=%s %-62s: %.3f ms  %.1f TFLOP/s  %.2f Texp/s (MUFU peak 4.65)  %.0f GB/s io\n", n, H, d,
           ident ? "identity" : "random", impl[v].name, ms, flops / ms / 1e9, exps / ms / 1e9, (double)n * C * 8 / ms / 1e6);
  }
  if (only < 0) {
    std::vector<__nv_bfloat16> a((size_t)n * C), b((size_t)n * C);
    cudaMemcpy(a.data(), dout1, a.size() * 2, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), dout, b.size() * 2, cudaMemcpyDeviceToHost);
    double mx = 0, sum = 0; size_t bad = 0;
    for (size_t i = 0; i < a.size(); ++i) {
      const float x = __bfloat162float(a[i]), y = __bfloat162float(b[i]);
      if (!(fabsf(x - y) <= 1e30f)) ++bad;
      mx = std::max(mx, (double)fabsf(x - y)); sum += fabs(x);
    }
    printf("  max |experimental - shipped| = %.5f (mean |out| %.4f, non-finite differences: %zu)\n", mx, sum / a.size(), bad);
  }
#ifdef SS_ATT_TRACE2
  {
    int kTraceSlots = 0, kTraceCtas = 0;
    ss_att_trace_read(nullptr, &kTraceSlots, &kTraceCtas);
    const int nb = np * H;
    const int ncta = std::min<int>(nb, kTraceCtas);
    std::vector<long long> tr((size_t)kTraceCtas * kTraceSlots);
    ss_att_trace_read(tr.data(), &kTraceSlots, &kTraceCtas);
    auto med = [&](int slot, int ref) {
      std::vector<long long> v;
      for (int b = 148; b < ncta; ++b) {  // skip the first wave (cold)
        long long t = tr[(size_t)b * kTraceSlots + slot], t0 = tr[(size_t)b * kTraceSlots + ref];
        if (t && t0) v.push_back(t - t0);
      }
      if (v.empty()) return -1LL;
      std::sort(v.begin(), v.end());
      return v[v.size() / 2];
    };
    const char* names[19] = {"entry", "after_first_sync", "kv0_landed", "kv1", "kv2", "kv3", "kv4", "kv5", "kv6", "kv7_landed",
                             "first_qk_issued", "first_S_seen", "q_norm_done", "tile0_steps_done", "tile0_epilogue_done",
                             "group0_done", "group1_done", "mma0_done", "cta_exit"};
    for (int sl = 1; sl < 19; ++sl) printf("  %-20s median %8lld clk after entry\n", names[sl], med(sl, 0));
    printf("  tile 1 of group 0, per step: S seen (rel. to step 0 of the tile) / p_ready arrived (rel. to own S seen)\n   ");
    for (int j = 0; j < 16; ++j) printf(" %6lld", med(20 + j, 20));
    printf("\n   ");
    for (int j = 0; j < 16; ++j) printf(" %6lld", med(36 + j, 20 + j));
    printf("\n");
  }
#endif
  return 0;
}
