// nn_search.cu -- harness: ordered m-nearest-neighbour search on the GPU.
//
// NOT part of the reference hot path: neighbour search is an input producer that stays in the
// reference (GpGp::find_ordered_nn at R/vecchia_specify.R:159).  The measurement harness needs
// the same arrays at n = 1e6..1e7 on the GPU box, where they cannot be shipped, so this file
// generates them: row i gets (i, its min(m, i) nearest among rows < i, nearest first; ties by lower
// index), written as the column-reversed, 1-based, 0-padded revNNarray of R/U_sparsity.R:32.
//
// Method: rows < kBrute are brute force.  Later rows are handled in doubling levels: level L holds
// a uniform cell grid over points [0, n_L), n_L = kBrute * 2^L; a row i in [n_L/2, n_L) searches
// that grid ring by ring around its home cell, keeping the m best among ids < i, and stops when the
// m-th best distance is within the radius fully covered by the rings visited so far.
#include "../../include/gpvecchia_b200.h"

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cub/cub.cuh>
#include <cuda_runtime.h>

namespace {

constexpr int kBrute = 2048;
constexpr int kMaxM = 63;

struct GridSpec {
  double mn[3];
  double inv_h[3];
  double h_min;     // smallest cell edge (for the stopping rule)
  int g[3];         // cells per dimension
  int d;
};

__device__ __forceinline__ int cell_coord(double x, double mn, double inv_h, int g) {
  int c = (int)floor((x - mn) * inv_h);
  return c < 0 ? 0 : (c >= g ? g - 1 : c);
}
__device__ __forceinline__ int cell_linear(const int* c, const GridSpec& s) {
  int idx = c[0];
  if (s.d > 1) idx += s.g[0] * c[1];
  if (s.d > 2) idx += s.g[0] * s.g[1] * c[2];
  return idx;
}

__global__ void count_cells_kernel(const double* __restrict__ locs, int n, GridSpec s,
                                   int* __restrict__ cell_of, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c[3] = {0, 0, 0};
  for (int k = 0; k < s.d; ++k) c[k] = cell_coord(locs[(int64_t)i * s.d + k], s.mn[k], s.inv_h[k], s.g[k]);
  const int ci = cell_linear(c, s);
  cell_of[i] = ci;
  atomicAdd(&counts[ci], 1);
}
__global__ void scatter_cells_kernel(const int* __restrict__ cell_of, int n, int* __restrict__ cursor,
                                     int* __restrict__ sorted_ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pos = atomicAdd(&cursor[cell_of[i]], 1);
  sorted_ids[pos] = i;
}

// keeps the m best (dist2, id) pairs sorted ascending, ties by lower id
struct TopM {
  double bd[kMaxM];
  int bi[kMaxM];
  int cnt, m;
  __device__ void init(int m_) { cnt = 0; m = m_; }
  __device__ double worst() const { return cnt < m ? DBL_MAX : bd[m - 1]; }
  __device__ void offer(double d2, int id) {
    if (cnt == m) {
      if (d2 > bd[m - 1] || (d2 == bd[m - 1] && id > bi[m - 1])) return;
    }
    int pos = cnt < m ? cnt : m - 1;
    while (pos > 0 && (bd[pos - 1] > d2 || (bd[pos - 1] == d2 && bi[pos - 1] > id))) {
      bd[pos] = bd[pos - 1];
      bi[pos] = bi[pos - 1];
      --pos;
    }
    bd[pos] = d2;
    bi[pos] = id;
    if (cnt < m) ++cnt;
  }
};

__device__ __forceinline__ double dist2(const double* __restrict__ locs, int d, const double* x, int j) {
  // same summation order as src/dist.cpp:12-15
  double s = 0.0;
  for (int k = 0; k < d; ++k) {
    const double t = x[k] - locs[(int64_t)j * d + k];
    s += t * t;
  }
  return s;
}

__device__ void write_row(const TopM& t, int64_t row, int64_t r_local, int64_t nrows, int m,
                          int32_t* __restrict__ out) {
  // NNarray = (row, best[0..cnt-1], NA...) ; revNNarray column c holds NNarray column m - c
  out[r_local + nrows * (int64_t)m] = (int32_t)(row + 1);
  for (int j = 0; j < m; ++j)
    out[r_local + nrows * (int64_t)(m - 1 - j)] = (j < t.cnt) ? (t.bi[j] + 1) : 0;
}

__global__ void brute_rows_kernel(const double* __restrict__ locs, int d, int m, int64_t row_begin,
                                  int64_t row_end, int64_t lim, int32_t* __restrict__ out) {
  const int64_t row = row_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= row_end || row >= lim) return;
  double x[3];
  for (int k = 0; k < d; ++k) x[k] = locs[row * d + k];
  TopM t;
  t.init(m);
  for (int j = 0; j < (int)row; ++j) t.offer(dist2(locs, d, x, j), j);
  write_row(t, row, row - row_begin, row_end - row_begin, m, out);
}

__global__ void grid_rows_kernel(const double* __restrict__ locs, int m, GridSpec s,
                                 const int* __restrict__ cell_start, const int* __restrict__ sorted_ids,
                                 int64_t lo, int64_t hi, int64_t row_begin, int64_t row_end,
                                 int32_t* __restrict__ out) {
  const int64_t row = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= hi) return;
  const int d = s.d;
  double x[3] = {0, 0, 0};
  int c0[3] = {0, 0, 0};
  for (int k = 0; k < d; ++k) {
    x[k] = locs[row * d + k];
    c0[k] = cell_coord(x[k], s.mn[k], s.inv_h[k], s.g[k]);
  }
  TopM t;
  t.init(m);
  int gmax = s.g[0];
  if (d > 1 && s.g[1] > gmax) gmax = s.g[1];
  if (d > 2 && s.g[2] > gmax) gmax = s.g[2];
  for (int r = 0; r <= gmax; ++r) {
    const int z0 = d > 2 ? -r : 0, z1 = d > 2 ? r : 0;
    const int y0 = d > 1 ? -r : 0, y1 = d > 1 ? r : 0;
    for (int dz = z0; dz <= z1; ++dz) {
      const int cz = c0[2] + dz;
      if (d > 2 && (cz < 0 || cz >= s.g[2])) continue;
      for (int dy = y0; dy <= y1; ++dy) {
        const int cy = c0[1] + dy;
        if (d > 1 && (cy < 0 || cy >= s.g[1])) continue;
        const bool edge = (d > 2 && (dz == -r || dz == r)) || (d > 1 && (dy == -r || dy == r));
        const int step = (edge || r == 0) ? 1 : 2 * r;       // interior of the ring: only dx = -r, +r
        for (int dx = -r; dx <= r; dx += step) {
          const int cx = c0[0] + dx;
          if (cx < 0 || cx >= s.g[0]) continue;
          int c[3] = {cx, cy, cz};
          const int ci = cell_linear(c, s);
          const int e = cell_start[ci + 1];
          for (int q = cell_start[ci]; q < e; ++q) {
            const int j = sorted_ids[q];
            if (j < (int)row) t.offer(dist2(locs, d, x, j), j);
          }
        }
      }
    }
    // every unvisited point differs by more than r cells along some axis: distance > r * h_min
    const double cover = (double)r * s.h_min;
    if (t.cnt == m && t.worst() <= cover * cover) break;
  }
  write_row(t, row, row - row_begin, row_end - row_begin, m, out);
}

#define NN_TRY(expr)                                                                        \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      std::fprintf(stderr, "gpv_harness_ordered_nn: %s failed: %s\n", #expr, cudaGetErrorString(_e)); \
      st = GPV_ERR_CUDA;                                                                    \
      goto done;                                                                            \
    }                                                                                       \
  } while (0)

}  // namespace

extern "C" gpv_status gpv_harness_ordered_nn(int64_t Nlocs, int d, int m, const double* locs,
                                             int64_t row_begin, int64_t row_end,
                                             int32_t* revNNarray_out, int device) {
  if (!locs || !revNNarray_out || Nlocs <= 0 || d < 1 || d > 3 || m < 1 || m > kMaxM ||
      row_begin < 0 || row_end > Nlocs || row_begin > row_end || Nlocs > INT32_MAX)
    return GPV_ERR_ARG;
  gpv_status st = GPV_OK;
  const int64_t nrows = row_end - row_begin;
  if (nrows == 0) return GPV_OK;
  double *d_cm = nullptr, *d_locs = nullptr;
  int32_t* d_out = nullptr;
  int *d_cell_of = nullptr, *d_counts = nullptr, *d_start = nullptr, *d_cursor = nullptr, *d_sorted = nullptr;
  void* d_scan = nullptr;
  size_t scan_bytes = 0;
  double mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
  for (int k = 0; k < d; ++k) {
    double a = locs[(size_t)k * Nlocs], b = a;
    for (int64_t i = 1; i < Nlocs; ++i) {
      const double v = locs[(size_t)k * Nlocs + i];
      a = v < a ? v : a;
      b = v > b ? v : b;
    }
    mn[k] = a; mx[k] = b;
  }
  NN_TRY(cudaSetDevice(device));
  NN_TRY(cudaMalloc(&d_cm, sizeof(double) * (size_t)Nlocs * d));
  NN_TRY(cudaMalloc(&d_locs, sizeof(double) * (size_t)Nlocs * d));
  NN_TRY(cudaMalloc(&d_out, sizeof(int32_t) * (size_t)nrows * (m + 1)));
  NN_TRY(cudaMemcpy(d_cm, locs, sizeof(double) * (size_t)Nlocs * d, cudaMemcpyHostToDevice));
  {
    // column-major -> row-major with a strided 2-D copy per coordinate
    for (int k = 0; k < d; ++k)
      NN_TRY(cudaMemcpy2D(d_locs + k, sizeof(double) * d, d_cm + (size_t)k * Nlocs, sizeof(double),
                          sizeof(double), (size_t)Nlocs, cudaMemcpyDeviceToDevice));
  }
  if (row_begin < kBrute) {
    const int64_t lim = row_end < kBrute ? row_end : kBrute;
    const int64_t cnt = lim - row_begin;
    brute_rows_kernel<<<(unsigned)((cnt + 127) / 128), 128>>>(d_locs, d, m, row_begin, row_end, lim, d_out);
    NN_TRY(cudaGetLastError());
  }
  {
    const size_t max_cells = (size_t)Nlocs + 8;
    NN_TRY(cudaMalloc(&d_cell_of, sizeof(int) * (size_t)Nlocs));
    NN_TRY(cudaMalloc(&d_counts, sizeof(int) * (max_cells + 1)));
    NN_TRY(cudaMalloc(&d_start, sizeof(int) * (max_cells + 1)));
    NN_TRY(cudaMalloc(&d_cursor, sizeof(int) * (max_cells + 1)));
    NN_TRY(cudaMalloc(&d_sorted, sizeof(int) * (size_t)Nlocs));
    NN_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_counts, d_start, (int)(max_cells + 1)));
    NN_TRY(cudaMalloc(&d_scan, scan_bytes));
    for (int64_t lo = kBrute; lo < Nlocs; lo *= 2) {
      const int64_t hi = (2 * lo < Nlocs) ? 2 * lo : Nlocs;      // rows [lo, hi) search points [0, hi)
      const int64_t a = lo > row_begin ? lo : row_begin;
      const int64_t b = hi < row_end ? hi : row_end;
      if (a >= b) continue;
      GridSpec s;
      s.d = d;
      const double per_cell = 2.0;
      int g = (int)floor(pow((double)hi / per_cell, 1.0 / d));
      if (g < 1) g = 1;
      size_t ncell = 1;
      s.h_min = DBL_MAX;
      for (int k = 0; k < 3; ++k) { s.g[k] = 1; s.mn[k] = 0; s.inv_h[k] = 0; }
      for (int k = 0; k < d; ++k) {
        s.g[k] = g;
        const double ext = (mx[k] - mn[k]) > 0 ? (mx[k] - mn[k]) : 1.0;
        const double h = ext / g * (1.0 + 1e-12);
        s.mn[k] = mn[k];
        s.inv_h[k] = 1.0 / h;
        s.h_min = h < s.h_min ? h : s.h_min;
        ncell *= (size_t)g;
      }
      NN_TRY(cudaMemset(d_counts, 0, sizeof(int) * (ncell + 1)));
      count_cells_kernel<<<(unsigned)((hi + 255) / 256), 256>>>(d_locs, (int)hi, s, d_cell_of, d_counts);
      NN_TRY(cub::DeviceScan::ExclusiveSum(d_scan, scan_bytes, d_counts, d_start, (int)(ncell + 1)));
      NN_TRY(cudaMemcpy(d_cursor, d_start, sizeof(int) * (ncell + 1), cudaMemcpyDeviceToDevice));
      scatter_cells_kernel<<<(unsigned)((hi + 255) / 256), 256>>>(d_cell_of, (int)hi, d_cursor, d_sorted);
      grid_rows_kernel<<<(unsigned)((b - a + 127) / 128), 128>>>(d_locs, m, s, d_start, d_sorted, a, b,
                                                                 row_begin, row_end, d_out);
      NN_TRY(cudaGetLastError());
    }
  }
  NN_TRY(cudaMemcpy(revNNarray_out, d_out, sizeof(int32_t) * (size_t)nrows * (m + 1), cudaMemcpyDeviceToHost));
done:
  cudaFree(d_cm); cudaFree(d_locs); cudaFree(d_out); cudaFree(d_cell_of); cudaFree(d_counts);
  cudaFree(d_start); cudaFree(d_cursor); cudaFree(d_sorted); cudaFree(d_scan);
  return st;
}
