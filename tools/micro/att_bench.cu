// Stand-alone A/B driver of the tcgen05 patch-attention kernels (developer tool): the shipped kernel
// (scenesplat_b200/csrc/attention_fwd.cu) against the round-1 kernel (tools/micro/attention_r1.cu) at a given shape,
// with the largest output difference between the two.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo \
//        -o tools/micro/att_bench tools/micro/att_bench.cu scenesplat_b200/csrc/attention_fwd.cu
//   att_bench n H d [reps] [order: 0 random | 1 identity]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <algorithm>
#include <numeric>
#include "attention_r1.cu"  // round-1 kernel (entry points renamed *_r1), kept for A/B; the shipped kernel is linked in
namespace ss { unsigned long long g_launch_count = 0; }

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
  const int np = (n + K - 1) / K;
  std::vector<int> table(4 * np);
  for (int p = 0; p < np; ++p) {
    int qb = p * K, qe = std::min(n, qb + K), kb = qb, ke = qe;
    if (p == np - 1 && n > K) kb = n - K;  // window rule of the last patch
    table[4 * p] = qb; table[4 * p + 1] = qe; table[4 * p + 2] = kb; table[4 * p + 3] = ke;
  }
  __nv_bfloat16 *dq, *dout; int64_t* dord; int* dtab;
  cudaMalloc(&dq, hq.size() * 2); cudaMalloc(&dout, (size_t)n * C * 2); cudaMalloc(&dord, n * 8); cudaMalloc(&dtab, table.size() * 4);
  cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dord, order.data(), n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dtab, table.data(), table.size() * 4, cudaMemcpyHostToDevice);
  char* flush; cudaMalloc(&flush, 256 << 20);
  float scale = 1.f / sqrtf((float)d);
  __nv_bfloat16* dout1; cudaMalloc(&dout1, (size_t)n * C * 2);
  typedef int (*fn_t)(const void*, const int64_t*, const int32_t*, int, int, int, int, float, void*, void*);
  struct { const char* name; fn_t fn; __nv_bfloat16* out; } impl[2] = {{"r1 (128-key steps, 1 S buffer/group)", ss_patch_attention_r1, dout1},
                                                                      {"shipped (64-key steps, 2 S/P buffers/group)", ss_patch_attention, dout}};
  const int only = argc > 6 ? atoi(argv[6]) : -1;  // 0: r1 only, 1: shipped only (for ncu)
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
    printf("n=%d H=%d d=%d order=%s %-44s: %.3f ms  %.1f TFLOP/s  %.2f Texp/s (MUFU peak 4.65)  %.0f GB/s io\n", n, H, d,
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
    printf("  max |shipped - r1| = %.5f (mean |out| %.4f, non-finite differences: %zu)\n", mx, sum / a.size(), bad);
  }
#ifdef SS_ATT_TRACE
  {
    const int nb = np * H;
    std::vector<long long> tr((size_t)nb * ss::kTraceSlots);
    cudaMemcpyFromSymbol(tr.data(), ss::g_att_trace, std::min(sizeof(long long) * tr.size(), sizeof(ss::g_att_trace)));
    const int ncta = std::min<int>(nb, ss::kTraceCtas);
    const char* names[16] = {"start", "loader_issued", "kv0", "kv1", "kv2", "kv3", "kv4", "kv5", "kv6", "kv7",
                             "sm_first_S", "sm_pair0_end", "sm_all_end", "cta_end", "mma_end", "ld_q_done"};
    auto med = [&](int slot, int ref) {
      std::vector<long long> v;
      for (int b = 0; b < ncta; ++b) {
        long long t = tr[(size_t)b * ss::kTraceSlots + slot], t0 = tr[(size_t)b * ss::kTraceSlots + ref];
        if (t && t0) v.push_back(t - t0);
      }
      if (v.empty()) return -1LL;
      std::sort(v.begin(), v.end());
      return v[v.size() / 2];
    };
    if (getenv("ATT_TRACE_WIDE")) {  // 16-softmax-warp kernel: warps 0 (group 0) and 8 (group 1), steps 8..15
      printf("  step |  g0: S_seen sweep1 exchanged exp_start exps_issued arrived |  g1: ...\n");
      for (int st = 0; st < 8; ++st) {
        printf("  %4d |", st + 8);
        for (int g = 0; g < 2; ++g) { for (int k = 0; k < 6; ++k) printf(" %7lld", med(16 + g * 48 + st * 6 + k, 16)); printf(" |"); }
        printf("\n");
      }
      return 0;
    }
    for (int s = 1; s < 16; ++s) printf("  %-14s median %8lld clk\n", names[s], med(s, 0));
    printf("  tile transition of group 0 (relative to its last p_ready arrival of tile 0): q_free seen by loader %lld, Q landed %lld, "
           "pv_done seen %lld, epilogue end %lld, q_full seen by MMA %lld, next S seen %lld\n",
           med(115, 112), med(116, 112), med(113, 112), med(11, 112), med(114, 112), med(16, 112));
    // second tile of each group (steady state): all relative to group 0's S-ready of step 8 (slot 16)
    printf("  step |  g0: S_seen ld_done max_done arrived |  g1: S_seen ld_done max_done arrived | mma: p0_seen p0_issued p1_seen p1_issued\n");
    for (int st = 0; st < 8; ++st) {
      printf("  %4d |", st + 8);
      for (int g = 0; g < 2; ++g) { for (int k = 0; k < 4; ++k) printf(" %7lld", med(16 + g * 32 + st * 4 + k, 16)); printf(" |"); }
      for (int k = 0; k < 4; ++k) printf(" %7lld", med(80 + st * 4 + k, 16));
      printf(" | sfree/qk g0 %7lld %7lld g1 %7lld %7lld", med(120 + st * 4, 16), med(121 + st * 4, 16), med(122 + st * 4, 16), med(123 + st * 4, 16));
      printf("\n");
    }
    printf("  warp 0 step 10 (rel. to exp start): exps done %lld, pv_done passed %lld, STTM done %lld, arrived %lld\n", med(160, 163), med(161, 163), med(162, 163), med(152, 163));
    printf("  group 0 step 10, warps 0..3: S_seen %lld %lld %lld %lld  arrived %lld %lld %lld %lld\n", med(156, 16), med(157, 16), med(158, 16),
           med(159, 16), med(152, 16), med(153, 16), med(154, 16), med(155, 16));
  }
#endif
  return 0;
}
