// Zero-shot evaluation tail on the GPU: k-nearest-neighbour label voting and the confusion-matrix update.
//
// Replaces (reference): pointcept/utils/misc.py:17-95 (neighbor_voting: scipy cKDTree.query(k) on the host +
// a numba majority vote) and the per-point Python loop of pointcept/engines/hooks/evaluator.py:830-834
// (`for gt, pred in zip(...)`: confusion[gt, pred] += 1 / fn_ignore[gt] += 1).
//
// kNN: the reference points are binned into a uniform grid (counting sort: count, host-side scan by the caller,
// fill), one thread per query walks cube shells of cells outwards and keeps the k best squared distances.
// Distances are evaluated in fp64 from the fp32 coordinates, exactly like cKDTree (which converts its input to
// double), so the neighbour SETS are the same; the search stops once the k-th best distance is <= the distance
// every unvisited cell is guaranteed to exceed.  Vote: most frequent valid label among the k neighbours, ties
// to the SMALLEST label (the reference scans classes upwards with a strict `>`), no valid label -> ignore_label.
#include "common.cuh"
#include "../../include/scenesplat_b200.h"

namespace ss {

constexpr int kVoteMaxK = 64;

struct VoteGrid {
  float ox, oy, oz;  // grid origin (min corner of the reference points)
  float inv_h, h;    // 1 / cell size, cell size
  int nx, ny, nz;
};

__device__ __forceinline__ int cell_coord(float x, float o, float inv_h) { return (int)floorf((x - o) * inv_h); }

__device__ __forceinline__ int64_t cell_of(const VoteGrid& g, float x, float y, float z) {
  const int cx = min(max(cell_coord(x, g.ox, g.inv_h), 0), g.nx - 1);
  const int cy = min(max(cell_coord(y, g.oy, g.inv_h), 0), g.ny - 1);
  const int cz = min(max(cell_coord(z, g.oz, g.inv_h), 0), g.nz - 1);
  return ((int64_t)cz * g.ny + cy) * g.nx + cx;
}

__global__ void __launch_bounds__(256)
vote_bin_count_kernel(const float* __restrict__ pts, int64_t m, VoteGrid g, int32_t* __restrict__ cell_count,
                      int64_t* __restrict__ cell_id) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = cell_of(g, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    cell_id[i] = c;
    atomicAdd(&cell_count[c], 1);
  }
}

// cell_start: exclusive prefix of cell_count; cursor: zero-initialised.  Writes the points of every cell
// contiguously (order inside a cell is arbitrary: only the SET of neighbours matters).
__global__ void __launch_bounds__(256)
vote_bin_fill_kernel(const float* __restrict__ pts, const int32_t* __restrict__ labels, const int64_t* __restrict__ cell_id,
                     int64_t m, const int64_t* __restrict__ cell_start, int32_t* __restrict__ cursor,
                     float4* __restrict__ binned) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = cell_id[i];
    const int64_t pos = cell_start[c] + atomicAdd(&cursor[c], 1);
    binned[pos] = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], __int_as_float(labels[i]));
  }
}

__global__ void __launch_bounds__(128)
knn_vote_kernel(const float4* __restrict__ binned, const int64_t* __restrict__ cell_start, VoteGrid g,
                const float* __restrict__ query, int64_t nq, int k, int ignore_label, int num_classes,
                int32_t* __restrict__ out) {
  const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  const float qx = query[3 * qi], qy = query[3 * qi + 1], qz = query[3 * qi + 2];
  const double dqx = qx, dqy = qy, dqz = qz;
  // unclamped cell of the query (it may lie outside the grid)
  const int cx = cell_coord(qx, g.ox, g.inv_h), cy = cell_coord(qy, g.oy, g.inv_h), cz = cell_coord(qz, g.oz, g.inv_h);
  double bd[kVoteMaxK];
  int bl[kVoteMaxK];
  int cnt = 0, worst = 0;
  double worst_d = -1.0;
  // shells beyond r_max contain no cell of the grid
  const int r_max = max(max(max(cx, g.nx - 1 - cx), max(cy, g.ny - 1 - cy)), max(cz, g.nz - 1 - cz));
  for (int r = 0; r <= r_max; ++r) {
    const int z0 = max(cz - r, 0), z1 = min(cz + r, g.nz - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, g.ny - 1);
    const int x0 = max(cx - r, 0), x1 = min(cx + r, g.nx - 1);
    for (int z = z0; z <= z1; ++z) {
      const bool zs = (z == cz - r) || (z == cz + r);
      for (int y = y0; y <= y1; ++y) {
        const bool ys = zs || (y == cy - r) || (y == cy + r);
        // on a z or y face the whole x range is shell; otherwise only the two x faces
        const int64_t row = ((int64_t)z * g.ny + y) * g.nx;
        for (int pass = 0; pass < 2; ++pass) {
          int xa, xb;
          if (ys) {
            if (pass || x0 > x1) break;
            xa = x0; xb = x1;
          } else {
            const int xf = pass ? cx + r : cx - r;
            if (xf < 0 || xf >= g.nx || (pass && r == 0)) continue;
            xa = xb = xf;
          }
          const int64_t s = cell_start[row + xa], e = cell_start[row + xb + 1];  // contiguous cells of an x run
          for (int64_t p = s; p < e; ++p) {
            const float4 v = binned[p];
            const double dx = (double)v.x - dqx, dy = (double)v.y - dqy, dz = (double)v.z - dqz;
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (cnt < k) {
              bd[cnt] = d2;
              bl[cnt] = __float_as_int(v.w);
              if (d2 > worst_d) { worst_d = d2; worst = cnt; }
              ++cnt;
            } else if (d2 < worst_d) {
              bd[worst] = d2;
              bl[worst] = __float_as_int(v.w);
              worst_d = bd[0];
              worst = 0;
              for (int u = 1; u < k; ++u)
                if (bd[u] > worst_d) { worst_d = bd[u]; worst = u; }
            }
          }
        }
      }
    }
    // every unvisited cell is farther than r * h from the query along at least one axis
    if (cnt == k) {
      const double bound = (double)r * (double)g.h;
      if (worst_d <= bound * bound) break;
    }
  }
  // majority vote over the valid labels; ties -> smallest label
  int best = ignore_label, best_count = 0;
  for (int a = 0; a < cnt; ++a) {
    const int la = bl[a];
    if (la == ignore_label || la < 0 || la >= num_classes) continue;
    int c = 0;
    for (int b = 0; b < cnt; ++b) c += (bl[b] == la);
    if (c > best_count || (c == best_count && la < best)) { best_count = c; best = la; }
  }
  out[qi] = best;
}

__global__ void __launch_bounds__(256)
confusion_update_kernel(const int64_t* __restrict__ gt, const int64_t* __restrict__ pred, int64_t n, int num_classes,
                        int64_t ignore_index, unsigned long long* __restrict__ confusion,
                        unsigned long long* __restrict__ fn_ignore) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = gt[i], p = pred[i];
    if (g < 0 || g >= num_classes) continue;  // (the caller passes gt already filtered by != ignore_index)
    if (p == ignore_index) atomicAdd(&fn_ignore[g], 1ull);
    else if (p >= 0 && p < num_classes) atomicAdd(&confusion[g * num_classes + p], 1ull);
  }
}

}  // namespace ss

extern "C" {

int ss_vote_bin_count(const float* pts, int64_t m, const float* origin3, float cell, int nx, int ny, int nz,
                      int32_t* cell_count, int64_t* cell_id, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (m < 0 || !(cell > 0.f) || nx < 1 || ny < 1 || nz < 1 || !origin3) return SS_BAD_ARGS;
  if (m == 0) return SS_OK;
  if (!pts || !cell_count || !cell_id) return SS_BAD_ARGS;
  ss::VoteGrid g{origin3[0], origin3[1], origin3[2], 1.f / cell, cell, nx, ny, nz};
  const int blocks = (int)ss::imin64(ss::ceil_div64(m, 256), 16 * ss::kNumSMs);
  ss::vote_bin_count_kernel<<<blocks, 256, 0, stream>>>(pts, m, g, cell_count, cell_id);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

int ss_vote_bin_fill(const float* pts, const int32_t* labels, const int64_t* cell_id, int64_t m,
                     const int64_t* cell_start, int32_t* cursor, void* binned_xyzl, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (m < 0) return SS_BAD_ARGS;
  if (m == 0) return SS_OK;
  if (!pts || !labels || !cell_id || !cell_start || !cursor || !binned_xyzl || (uintptr_t)binned_xyzl % 16 != 0)
    return SS_BAD_ARGS;
  const int blocks = (int)ss::imin64(ss::ceil_div64(m, 256), 16 * ss::kNumSMs);
  ss::vote_bin_fill_kernel<<<blocks, 256, 0, stream>>>(pts, labels, cell_id, m, cell_start, cursor, (float4*)binned_xyzl);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

int ss_knn_vote(const void* binned_xyzl, const int64_t* cell_start, const float* origin3, float cell, int nx, int ny,
                int nz, const float* query, int64_t nq, int k, int ignore_label, int num_classes, int32_t* out,
                void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (nq < 0 || k < 1 || k > ss::kVoteMaxK || !(cell > 0.f) || nx < 1 || ny < 1 || nz < 1 || num_classes < 1 || !origin3)
    return SS_BAD_ARGS;
  if (nq == 0) return SS_OK;
  if (!binned_xyzl || !cell_start || !query || !out) return SS_BAD_ARGS;
  ss::VoteGrid g{origin3[0], origin3[1], origin3[2], 1.f / cell, cell, nx, ny, nz};
  const int blocks = (int)ss::ceil_div64(nq, 128);
  ss::knn_vote_kernel<<<blocks, 128, 0, stream>>>((const float4*)binned_xyzl, cell_start, g, query, nq, k, ignore_label,
                                                  num_classes, out);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

int ss_confusion_update(const int64_t* gt, const int64_t* pred, int64_t n, int num_classes, int64_t ignore_index,
                        int64_t* confusion, int64_t* fn_ignore, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || num_classes < 1) return SS_BAD_ARGS;
  if (n == 0) return SS_OK;
  if (!gt || !pred || !confusion || !fn_ignore) return SS_BAD_ARGS;
  const int blocks = (int)ss::imin64(ss::ceil_div64(n, 256), 16 * ss::kNumSMs);
  ss::confusion_update_kernel<<<blocks, 256, 0, stream>>>(gt, pred, n, num_classes, ignore_index,
                                                          (unsigned long long*)confusion, (unsigned long long*)fn_ignore);
  SS_CHECK_LAUNCH();
  return SS_OK;
}

}  // extern "C"
