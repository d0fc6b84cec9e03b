// Where does a pooling-level index launch spend its time?  Stand-alone copy of the launch shape of
// csrc/pool.cu::PoolRuns (same runs.cuh kernel, compiled with SS_RUNS_TRACE) on synthetic sorted codes; every CTA
// records %globaltimer at its phase boundaries, the host prints the distribution per row class.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DSS_RUNS_TRACE -I../../scenesplat_b200/csrc -o runs_trace runs_trace.cu
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>
#include "runs.cuh"
namespace ss { unsigned long long g_launch_count = 0; }

struct Pool {
  static constexpr bool kAfterRow0 = true;
  const int64_t *code, *order;
  int64_t n;
  int shift;
  int64_t *cluster, *seg_start, *ccode, *corder, *cinverse;
  __device__ bool head(int y, int64_t j) const {
    if (j == 0) return true;
    const int64_t* c = code + (size_t)y * n;
    const int64_t* o = order + (size_t)y * n;
    return (c[o[j]] >> shift) != (c[o[j - 1]] >> shift);
  }
  __device__ void emit(int y, int64_t j, uint32_t run, bool is_head) const {
    if (y == 0) {
      const int64_t p = order[j];
      cluster[p] = run;
      if (is_head) {
        seg_start[run] = j;
        for (int r = 0; r < 4; ++r) ccode[(size_t)r * n + run] = code[(size_t)r * n + p] >> shift;
        corder[run] = run;
        cinverse[run] = run;
      }
    } else if (is_head) {
      const int64_t c = __ldcg(cluster + order[(size_t)y * n + j]);
      corder[(size_t)y * n + run] = c;
      cinverse[(size_t)y * n + c] = run;
    }
  }
  __device__ void finish(int y, uint32_t total) const {
    if (y == 0) seg_start[total] = n;
  }
};

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 299277;
  const int rows = 4;
  std::mt19937_64 rng(1);
  std::vector<int64_t> code(rows * n), order(rows * n);
  for (int r = 0; r < rows; ++r) {
    std::vector<int64_t> keys(n);
    for (auto& k : keys) k = rng() & ((1ll << 27) - 1);
    std::sort(keys.begin(), keys.end());
    std::vector<int64_t> perm(n);
    std::iota(perm.begin(), perm.end(), 0);
    std::shuffle(perm.begin(), perm.end(), rng);
    for (int64_t j = 0; j < n; ++j) {
      order[r * n + j] = perm[j];
      code[r * n + perm[j]] = keys[j];
    }
  }
  int64_t *d_code, *d_order, *d_out;
  cudaMalloc(&d_code, rows * n * 8);
  cudaMalloc(&d_order, rows * n * 8);
  cudaMalloc(&d_out, (size_t)(2 + 3 * rows) * n * 8 + 64);
  cudaMemcpy(d_code, code.data(), rows * n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d_order, order.data(), rows * n * 8, cudaMemcpyHostToDevice);
  Pool f{d_code, d_order, n, 3, d_out, d_out + n, d_out + 2 * n + 8, d_out + (2 + rows) * n + 8, d_out + (2 + 2 * rows) * n + 8};
  void* ws;
  cudaMalloc(&ws, ss::runs_workspace_bytes(n, rows) + 256);
  const int tiles = (int)ss::ceil_div64(n, ss::kRunTile), ctas = tiles * rows;
  unsigned long long* d_tr;
  cudaMalloc(&d_tr, (size_t)ctas * 8 * 8);
  cudaMemcpyToSymbol(ss::g_runs_trace, &d_tr, sizeof(d_tr));
  int64_t* d_m;
  cudaMalloc(&d_m, 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(a);
    ss::runs_launch(f, n, ws, d_m, 0, rows);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("rep %d: memset + kernel %.1f us\n", rep, ms * 1e3);
  }
  std::vector<unsigned long long> tr((size_t)ctas * 8);
  cudaMemcpy(tr.data(), d_tr, tr.size() * 8, cudaMemcpyDeviceToHost);
  unsigned long long t0 = ~0ull;
  for (int c = 0; c < ctas; ++c) t0 = std::min(t0, tr[c * 8]);
  const char* names[7] = {"start", "ticket", "flags(loads)", "scan", "tile prefix", "row-0 wait", "emit+done"};
  for (int cls = 0; cls < 2; ++cls) {
    printf("%s (ns after the first CTA start; median / max over CTAs)\n", cls == 0 ? "row 0" : "rows 1..3");
    for (int ph = 0; ph < 7; ++ph) {
      std::vector<double> v;
      for (int c = 0; c < ctas; ++c) {
        const int row = (int)tr[c * 8 + 7];
        if ((row == 0) == (cls == 0)) v.push_back((double)(tr[c * 8 + ph] - t0));
      }
      std::sort(v.begin(), v.end());
      printf("  %-14s %8.0f %8.0f\n", names[ph], v[v.size() / 2], v.back());
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
