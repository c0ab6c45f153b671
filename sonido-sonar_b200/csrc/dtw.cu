// Dynamic time warping (float64, bit-exact with the reference's recurrence).
//
//   DTWAlignment.Align        algorithms/stats/dtw.go:55-103
//   fillCostMatrix            :106-135   C[i][j] = dist(q[i-1], r[j-1]) + step(C, i, j)
//   applyStepPattern          :138-162   symmetric2 = min(min(v, h), d); asymmetric; symmetric1
//   backtrack/findPreviousStep :165-217  strict-'<' scan in the order vertical, horizontal, diagonal
//   EuclideanDistanceFunc     algorithms/stats/distance.go:29-36
//
// The reference fills the matrix row by row; the value of a cell depends only on its three
// predecessors, so any schedule that respects those dependencies produces bit-identical
// values.  The fill kernel runs the anti-diagonal wavefront: one CTA per pair, one
// __syncthreads per diagonal d = i + j.  Cells are addressed by their offset o = i - j (+shift):
// on diagonal d the predecessors (i-1, j) and (i, j-1) are the cells of offsets o-1 and o+1
// written on diagonal d-1, and (i-1, j-1) is the cell of the same offset written on d-2, so a
// single shared-memory line of "latest value per offset" (initialised to +Inf, 0 at o = 0)
// carries the whole recurrence and also yields the +Inf borders and the Sakoe-Chiba band for
// free.  Cost cells are streamed to a banded store in HBM ((2*band+1) cells per row, or m cells
// per row when unconstrained) that the backtrack kernel and the optional CostMatrix export read.
//
// The backtrack is a pointer chase; it walks from (n, m) through shared-memory tiles of the
// store (64 x 64 cells loaded cooperatively, one thread walking) so that each step costs a
// shared-memory access instead of a DRAM round trip, and writes the path back to front.
#include <cmath>

#include "common.h"

namespace sonar {
namespace {

__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

__device__ __forceinline__ double local_dist(const double* __restrict__ q, const double* __restrict__ r, int dim) {
  double s = 0.0;
  for (int k = 0; k < dim; ++k) {
    const double d = __ldg(q + k) - __ldg(r + k);
    s += d * d;  // -fmad=false
  }
  return sqrt(s);
}

// implicit borders: C[0][0] = 0, first row/column +Inf, out-of-band +Inf
__device__ __forceinline__ double cell_get(const double* __restrict__ cells, const DtwGeom& g, int i, int j) {
  if (i == 0 && j == 0) return 0.0;
  if (i <= 0 || j <= 0) return d_inf();
  if (g.band > 0) {
    const int df = i - j;
    if (df > g.band || df < -g.band) return d_inf();
    return cells[(int64_t)(i - 1) * g.W + (j - i + g.band)];
  }
  return cells[(int64_t)(i - 1) * g.W + (j - 1)];
}

__global__ void __launch_bounds__(1024) dtw_fill_kernel(const double* __restrict__ qs, const double* __restrict__ rs,
                                                        DtwGeom g, int dim, int step, double* __restrict__ cells_all,
                                                        double* __restrict__ line_global, int line_in_smem) {
  extern __shared__ double s_line[];
  const int pair = blockIdx.x;
  const double* __restrict__ q = qs + (int64_t)pair * g.n * dim;
  const double* __restrict__ r = rs + (int64_t)pair * g.m * dim;
  double* __restrict__ cells = cells_all + (int64_t)pair * g.cells;
  const int shift = g.band > 0 ? g.band : g.m;  // o = i - j + shift in [0, n_off)
  const int n_off = g.n_off;
  double* line = line_in_smem ? s_line : line_global + (int64_t)pair * (n_off + 2);
  // line[o + 1]; line[0] and line[n_off + 1] are +Inf sentinels
  for (int k = threadIdx.x; k < n_off + 2; k += blockDim.x) line[k] = d_inf();
  __syncthreads();
  if (threadIdx.x == 0) line[shift + 1] = 0.0;  // C[0][0]
  __syncthreads();
  const int n = g.n, m = g.m, band = g.band;
  for (int d = 2; d <= n + m; ++d) {
    int ilo = d - m > 1 ? d - m : 1;
    int ihi = d - 1 < n ? d - 1 : n;
    if (band > 0) {
      const int lo2 = (d - band + 1) >> 1;  // ceil((d - band) / 2) for any sign (arithmetic shift)
      const int hi2 = (d + band) >> 1;      // floor((d + band) / 2)
      ilo = ilo > lo2 ? ilo : lo2;
      ihi = ihi < hi2 ? ihi : hi2;
    }
    for (int i = ilo + threadIdx.x; i <= ihi; i += blockDim.x) {
      const int j = d - i;
      const int o = i - j + shift + 1;
      const double ld = local_dist(q + (int64_t)(i - 1) * dim, r + (int64_t)(j - 1) * dim, dim);
      const double v = line[o - 1], h = line[o + 1], dg = line[o];
      double mc;
      if (step == SONAR_STEP_SYMMETRIC2)
        mc = fmin(fmin(v, h), dg);
      else if (step == SONAR_STEP_ASYMMETRIC)
        mc = fmin(v, h);
      else
        mc = fmin(v + 1.0, fmin(h + 1.0, dg));
      const double c = ld + mc;
      line[o] = c;
      const int64_t col = band > 0 ? (j - i + band) : (j - 1);
      cells[(int64_t)(i - 1) * g.W + col] = c;
    }
    __syncthreads();
  }
}

constexpr int kBtTile = 64;
constexpr int kBtThreads = 128;

__global__ void __launch_bounds__(kBtThreads) dtw_backtrack_kernel(const double* __restrict__ cells_all, DtwGeom g,
                                                                   int32_t* __restrict__ path_q,
                                                                   int32_t* __restrict__ path_r,
                                                                   double* __restrict__ path_c, int64_t path_cap,
                                                                   DtwPairOut* __restrict__ outs) {
  __shared__ double tile[kBtTile + 1][kBtTile + 1];  // tile[a][b] = C[i0 - a][j0 - b]
  __shared__ int s_i, s_j;
  __shared__ int64_t s_len;
  const int pair = blockIdx.x;
  const double* __restrict__ cells = cells_all + (int64_t)pair * g.cells;
  int32_t* pq = path_q + (int64_t)pair * path_cap;
  int32_t* pr = path_r + (int64_t)pair * path_cap;
  double* pc = path_c + (int64_t)pair * path_cap;
  if (threadIdx.x == 0) {
    s_i = g.n;
    s_j = g.m;
    s_len = 0;
  }
  __syncthreads();
  while (true) {
    const int i0 = s_i, j0 = s_j;
    if (i0 <= 0 && j0 <= 0) break;
    // load the tile anchored at (i0, j0): rows i0..i0-T, cols j0..j0-T
    for (int e = threadIdx.x; e < (kBtTile + 1) * (kBtTile + 1); e += kBtThreads) {
      const int a = e / (kBtTile + 1), b = e % (kBtTile + 1);
      tile[a][b] = cell_get(cells, g, i0 - a, j0 - b);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int i = i0, j = j0;
      int64_t len = s_len;
      // walk while the three predecessors are inside the tile
      while ((i > 0 || j > 0) && (i0 - i) < kBtTile && (j0 - j) < kBtTile) {
        const int a = i0 - i, b = j0 - j;
        double cost = 0.0;
        if (i > 0 && j > 0) cost = tile[a][b] - tile[a + 1][b + 1];
        const int64_t pos = path_cap - 1 - len;
        if (pos >= 0) {
          pq[pos] = i - 1;
          pr[pos] = j - 1;
          pc[pos] = cost;
        }
        ++len;
        if (i == 0) {
          j = j - 1;
        } else if (j == 0) {
          i = i - 1;
        } else {
          const double cv = tile[a + 1][b], ch = tile[a][b + 1], cd = tile[a + 1][b + 1];
          int mi = 0;
          double best = cv;
          if (ch < best) {
            mi = 1;
            best = ch;
          }
          if (cd < best) mi = 2;
          if (mi == 0)
            i = i - 1;
          else if (mi == 1)
            j = j - 1;
          else {
            i = i - 1;
            j = j - 1;
          }
        }
      }
      s_i = i;
      s_j = j;
      s_len = len;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    outs[pair].total_cost = cell_get(cells, g, g.n, g.m);
    outs[pair].path_len = s_len;
  }
}

// CostMatrix = costMatrix[1:] (dtw.go:96): full[n][m+1]
__global__ void dtw_expand_kernel(const double* __restrict__ cells, DtwGeom g, double* __restrict__ full) {
  const int64_t total = (int64_t)g.n * (g.m + 1);
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / (g.m + 1)) + 1, j = (int)(e % (g.m + 1));
    full[e] = cell_get(cells, g, i, j);
  }
}

}  // namespace

int dtw_geometry(int n, int m, int band, DtwGeom* g) {
  g->n = n;
  g->m = m;
  g->band = band > 0 ? band : 0;
  if (g->band > 0) {
    g->W = 2 * (int64_t)g->band + 1;
    g->n_off = (int)g->W;
  } else {
    g->W = m;
    g->n_off = n + m + 1;
  }
  g->cells = (int64_t)n * g->W;
  return SONAR_OK;
}

int launch_dtw(const double* q, const double* r, int n_pairs, const DtwGeom& g, int dim, int step, double* cells,
               double* line_scratch, int32_t* path_q, int32_t* path_r, double* path_c, int64_t path_cap,
               DtwPairOut* out, cudaStream_t st) {
  if (n_pairs <= 0) return SONAR_OK;
  const size_t line_bytes = sizeof(double) * (size_t)(g.n_off + 2);
  const int in_smem = line_bytes <= 200 * 1024;
  if (!in_smem && !line_scratch) return set_error(SONAR_ERR_INVALID, "DTW line scratch missing");
  int diag = g.n < g.m ? g.n : g.m;
  if (g.band > 0 && diag > g.band + 1) diag = g.band + 1;
  int threads = (diag + 31) & ~31;
  threads = threads < 32 ? 32 : (threads > 1024 ? 1024 : threads);
  const size_t smem = in_smem ? line_bytes : 0;
  SONAR_CUDA(cudaFuncSetAttribute(dtw_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
  prof_begin("dtw_fill_kernel", st);
  dtw_fill_kernel<<<n_pairs, threads, smem, st>>>(q, r, g, dim, step, cells, line_scratch, in_smem);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  prof_begin("dtw_backtrack_kernel", st);
  dtw_backtrack_kernel<<<n_pairs, kBtThreads, 0, st>>>(cells, g, path_q, path_r, path_c, path_cap, out);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

int launch_dtw_expand(const double* cells, const DtwGeom& g, double* full, cudaStream_t st) {
  const int64_t total = (int64_t)g.n * (g.m + 1);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  prof_begin("dtw_expand_kernel", st);
  dtw_expand_kernel<<<(unsigned)blocks, 256, 0, st>>>(cells, g, full);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

}  // namespace sonar
