// Stand-alone timing driver of the tcgen05 patch-attention BACKWARD (developer tool; values are not checked here:
// tests/test_gpu_train.py does that).  lse2 / out are synthetic, which does not change the work done.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo [-DSS_ATTB_POLY=k] \
//        -o tools/micro/attbwd_bench tools/micro/attbwd_bench.cu
//   attbwd_bench n H d [reps]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <algorithm>
#include <numeric>
#include "../../scenesplat_b200/csrc/attention_bwd.cu"
namespace ss { unsigned long long g_launch_count = 0; }

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 163814, H = argc > 2 ? atoi(argv[2]) : 16, d = argc > 3 ? atoi(argv[3]) : 48;
  const int reps = argc > 4 ? atoi(argv[4]) : 5;
  const int K = 1024, C = H * d;
  std::mt19937 rng(0);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<__nv_bfloat16> hq((size_t)n * 3 * C), hg((size_t)n * C);
  for (auto& v : hq) v = __float2bfloat16(nd(rng));
  for (auto& v : hg) v = __float2bfloat16(nd(rng));
  std::vector<float> hl((size_t)n * H, 12.f);
  std::vector<int64_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::shuffle(order.begin(), order.end(), rng);
  const int np = (n + K - 1) / K;
  std::vector<int> table(4 * np);
  for (int p = 0; p < np; ++p) {
    int qb = p * K, qe = std::min(n, qb + K), kb = qb, ke = qe;
    if (p == np - 1 && n > K) kb = n - K;
    table[4 * p] = qb; table[4 * p + 1] = qe; table[4 * p + 2] = kb; table[4 * p + 3] = ke;
  }
  __nv_bfloat16 *dq, *dg, *dout, *dd; float* dl; int64_t* dord; int* dtab; void* ws;
  const size_t wsb = ss_patch_attention_backward_workspace_bytes(n, H, d);
  cudaMalloc(&dq, hq.size() * 2); cudaMalloc(&dg, hg.size() * 2); cudaMalloc(&dout, hg.size() * 2); cudaMalloc(&dd, hq.size() * 2);
  cudaMalloc(&dl, hl.size() * 4); cudaMalloc(&dord, n * 8); cudaMalloc(&dtab, table.size() * 4); cudaMalloc(&ws, wsb);
  cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dg, hg.data(), hg.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dout, hg.data(), hg.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dl, hl.data(), hl.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dord, order.data(), n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dtab, table.data(), table.size() * 4, cudaMemcpyHostToDevice);
  const float scale = 1.f / sqrtf((float)d);
  for (int i = 0; i < 2; ++i) {
    int rc = ss_patch_attention_backward(dq, dout, dg, dl, dord, dtab, np, K, H, d, scale, n, dd, ws, wsb, 0);
    if (rc) { printf("launch rc=%d\n", rc); return 1; }
  }
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float tot = 0;
  for (int i = 0; i < reps; ++i) {
    cudaEventRecord(e0);
    ss_patch_attention_backward(dq, dout, dg, dl, dord, dtab, np, K, H, d, scale, n, dd, ws, wsb, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); tot += ms;
  }
  const double ms = tot / reps;
  printf("poly=%d n=%d H=%d d=%d: backward %.3f ms  %.1f TFLOP/s (14 K C n)  %.2f Texp/s\n", SS_ATTB_POLY, n, H, d, ms,
         14.0 * K * C * (double)n / ms / 1e9, 2.0 * n * (double)K * H / ms / 1e9);
  return 0;
}
