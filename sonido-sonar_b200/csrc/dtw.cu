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

constexpr int kRing = 2048;       // staged q / r window (power of two), doubles each
constexpr int kStageBytes = 48 * 1024;  // cost cells staged in shared memory between flushes

// Two things keep the per-diagonal critical path short:
//  * bar.sync orders the CTA's global stores, i.e. a barrier that follows a global store waits for its
//    L2 round trip (~600 cycles).  When every thread owns at most one cell per diagonal (ONECELL) the
//    cost cells are therefore staged in shared memory and flushed to the HBM band store once every
//    `stage_diags` diagonals instead of once per diagonal.
//  * STAGED (dim == 1, 0 < band <= 1000): the q / r values a diagonal needs lie in a window of band+1
//    indices that advances by one every two diagonals, so they are staged in two shared-memory rings
//    (refilled every 256 diagonals) and the local distance of the NEXT diagonal is computed before the
//    barrier, off the dependent min/add chain.
template <bool STAGED, bool ONECELL>
__global__ void __launch_bounds__(1024) dtw_fill_kernel(const double* __restrict__ qs, const double* __restrict__ rs,
                                                        DtwGeom g, int dim, int step, double* __restrict__ cells_all,
                                                        double* __restrict__ line_global, int line_in_smem,
                                                        int stage_diags) {
  extern __shared__ double s_mem[];
  double* ring_q = s_mem;
  double* ring_r = s_mem + kRing;
  double* s_stage = s_mem + (STAGED ? 2 * kRing : 0);
  double* s_line = s_stage + (ONECELL ? (size_t)stage_diags * blockDim.x : 0);
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
  auto range = [&](int d, int& ilo, int& ihi) {
    ilo = d - m > 1 ? d - m : 1;
    ihi = d - 1 < n ? d - 1 : n;
    if (band > 0) {
      const int lo2 = (d - band + 1) >> 1;  // ceil((d - band) / 2) for any sign (arithmetic shift)
      const int hi2 = (d + band) >> 1;      // floor((d + band) / 2)
      ilo = ilo > lo2 ? ilo : lo2;
      ihi = ihi < hi2 ? ihi : hi2;
    }
  };
  auto staged_dist = [&](int i, int j) -> double {
    const double df = ring_q[(i - 1) & (kRing - 1)] - ring_r[(j - 1) & (kRing - 1)];
    const double ad = fabs(df);
    // sqrt(x*x) == |x| in binary floating point whenever x*x neither underflows nor overflows
    return (ad > 1e-150 && ad < 1e150) ? ad : sqrt(df * df);
  };
  auto flush = [&](int d_first, int count) {  // staged diagonals d_first .. d_first+count-1 -> HBM
    for (int e = threadIdx.x; e < count * (int)blockDim.x; e += blockDim.x) {
      const int sl = e / blockDim.x, t = e - sl * blockDim.x;
      const int d = d_first + sl;
      int ilo, ihi;
      range(d, ilo, ihi);
      const int i = ilo + t;
      if (i <= ihi) {
        const int j = d - i;
        const int64_t col = band > 0 ? (j - i + band) : (j - 1);
        cells[(int64_t)(i - 1) * g.W + col] = s_stage[e];
      }
    }
  };
  int loaded = 0, staged = 0, stage_first = 2;
  double ld_next = 0.0;
  for (int d = 2; d <= n + m; ++d) {
    if constexpr (STAGED) {
      if (((d - 2) & 255) == 0) {
        const int target = ((d + 256 + band) >> 1) + 2;
        for (int e = loaded + threadIdx.x; e < target; e += blockDim.x) {
          if (e < n) ring_q[e & (kRing - 1)] = q[e];
          if (e < m) ring_r[e & (kRing - 1)] = r[e];
        }
        loaded = target;
        __syncthreads();
      }
    }
    int ilo, ihi;
    range(d, ilo, ihi);
    for (int i = ilo + threadIdx.x; i <= ihi; i += blockDim.x) {
      const int j = d - i;
      const int o = i - j + shift + 1;
      double ld;
      if constexpr (STAGED)
        ld = d == 2 ? staged_dist(i, j) : ld_next;
      else
        ld = local_dist(q + (int64_t)(i - 1) * dim, r + (int64_t)(j - 1) * dim, dim);
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
      if constexpr (ONECELL) {
        s_stage[staged * blockDim.x + threadIdx.x] = c;
      } else {
        const int64_t col = band > 0 ? (j - i + band) : (j - 1);
        cells[(int64_t)(i - 1) * g.W + col] = c;
      }
    }
    if constexpr (STAGED) {  // one cell per thread (blockDim >= band + 1): prefetch its distance for d + 1
      int nlo, nhi;
      range(d + 1, nlo, nhi);
      const int i = nlo + threadIdx.x;
      if (i <= nhi) ld_next = staged_dist(i, d + 1 - i);
    }
    __syncthreads();
    if constexpr (ONECELL) {
      if (++staged == stage_diags || d == n + m) {
        flush(stage_first, staged);  // the next barrier (or kernel end) completes these stores
        stage_first = d + 1;
        staged = 0;
        __syncthreads();  // staging buffer free again
      }
    }
  }
}

// Fast path for the configuration the alignment pipeline uses (dim == 1, 0 < band <= 511): thread t owns
// one fixed offset i - j per diagonal parity, so row/column indices, the band-store address and the
// three shared-memory line slots only ever advance by constants; the per-diagonal work is ~25
// instructions and the dependent chain is LDS -> min -> min -> add -> STS -> barrier.
template <int STEP>
__global__ void __launch_bounds__(512) dtw_fill_band1_kernel(const double* __restrict__ qs,
                                                             const double* __restrict__ rs, DtwGeom g,
                                                             double* __restrict__ cells_all) {
  extern __shared__ double s_mem[];
  double* ring_q = s_mem;
  double* ring_r = s_mem + kRing;
  double* line = s_mem + 2 * kRing;  // line[delta + band + 1], +Inf sentinels at both ends
  const int pair = blockIdx.x, t = threadIdx.x;
  const int n = g.n, m = g.m, band = g.band, W = (int)g.W;
  const double* __restrict__ q = qs + (int64_t)pair * n;
  const double* __restrict__ r = rs + (int64_t)pair * m;
  double* __restrict__ cells = cells_all + (int64_t)pair * g.cells;
  for (int k = t; k < 2 * band + 3; k += blockDim.x) line[k] = d_inf();
  __syncthreads();
  if (t == 0) line[band + 1] = 0.0;  // C[0][0]
  // even diagonals carry offsets of even parity, odd diagonals of odd parity
  const int de = -band + (band & 1) + 2 * t, dod = -band + ((band + 1) & 1) + 2 * t;
  const bool act_e = de <= band, act_o = dod <= band;
  int ie = (2 + de) >> 1, je = 2 - ie;  // cell of this thread on d = 2
  int io = (3 + dod) >> 1, jo = 3 - io;  // ... and on d = 3
  int64_t offe = (int64_t)(ie - 1) * W + (band - de), offo = (int64_t)(io - 1) * W + (band - dod);
  const int oe = de + band + 1, oo = dod + band + 1;
  auto dist = [&](int i, int j) -> double {
    const double df = ring_q[(i - 1) & (kRing - 1)] - ring_r[(j - 1) & (kRing - 1)];
    const double ad = fabs(df);
    // sqrt(x*x) == |x| in binary floating point whenever x*x neither underflows nor overflows
    return (ad > 1e-150 && ad < 1e150) ? ad : sqrt(df * df);
  };
  auto relax = [&](int o, double ld) -> double {
    const double v = line[o - 1], h = line[o + 1], dg = line[o];
    double mc;
    if (STEP == SONAR_STEP_SYMMETRIC2)
      mc = fmin(fmin(v, h), dg);
    else if (STEP == SONAR_STEP_ASYMMETRIC)
      mc = fmin(v, h);
    else
      mc = fmin(v + 1.0, fmin(h + 1.0, dg));
    const double c = ld + mc;
    line[o] = c;
    return c;
  };
  int loaded = 0;
  double lde = 0.0, ldo = 0.0;
  const int last = n + m;
  // A barrier that follows a global store waits for the store's L2 round trip, so the cells of 16
  // diagonals (8 per parity and thread) are kept in registers and written out together.
  for (int d = 2; d <= last; d += 16) {
    if (((d - 2) & 255) == 0) {
      const int target = ((d + 258 + band) >> 1) + 2;
      for (int e = loaded + t; e < target; e += blockDim.x) {
        if (e < n) ring_q[e & (kRing - 1)] = q[e];
        if (e < m) ring_r[e & (kRing - 1)] = r[e];
      }
      loaded = target;
      __syncthreads();
      // distances of the two diagonals this refill makes reachable first
      if (act_e && (unsigned)(ie - 1) < (unsigned)n && (unsigned)(je - 1) < (unsigned)m) lde = dist(ie, je);
      if (act_o && (unsigned)(io - 1) < (unsigned)n && (unsigned)(jo - 1) < (unsigned)m) ldo = dist(io, jo);
    }
    double ce[8], co[8];
    unsigned me = 0, mo = 0;
    const int64_t be = offe, bo = offo;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      // ---- even diagonal d + 2u
      if (d + 2 * u <= last && act_e && (unsigned)(ie - 1) < (unsigned)n && (unsigned)(je - 1) < (unsigned)m) {
        ce[u] = relax(oe, lde);
        me |= 1u << u;
      }
      ++ie, ++je, offe += W;
      if (act_e && (unsigned)(ie - 1) < (unsigned)n && (unsigned)(je - 1) < (unsigned)m) lde = dist(ie, je);
      __syncthreads();
      // ---- odd diagonal d + 2u + 1
      if (d + 2 * u + 1 <= last && act_o && (unsigned)(io - 1) < (unsigned)n && (unsigned)(jo - 1) < (unsigned)m) {
        co[u] = relax(oo, ldo);
        mo |= 1u << u;
      }
      ++io, ++jo, offo += W;
      if (act_o && (unsigned)(io - 1) < (unsigned)n && (unsigned)(jo - 1) < (unsigned)m) ldo = dist(io, jo);
      __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (me & (1u << u)) cells[be + (int64_t)u * W] = ce[u];
      if (mo & (1u << u)) cells[bo + (int64_t)u * W] = co[u];
    }
  }
}

constexpr int kBtTile = 64;
constexpr int kBtThreads = 256;

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
#pragma unroll 6
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

// Banded store: the band is narrow enough to hold whole rows, so the walk proceeds through blocks of
// kBbRows consecutive rows (all 2*band+1 columns, borders normalised while loading).  Which block
// comes next does not depend on the path, so warps 1..7 prefetch block k+1 into the other buffer
// while lane 0 of warp 0 walks block k out of shared memory.
constexpr int kBbThreads = 256;

// Rows are dealt to warps two at a time; a lane loads up to 4 columns of each row with loads that do not
// depend on one another (address clamped to a valid location, value selected afterwards), so a warp has
// 8 DRAM requests in flight per lane.
__device__ __forceinline__ void bb_load_block(const double* __restrict__ cells, const DtwGeom& g, int ibase, int rows,
                                              double* __restrict__ buf, int t0, int nt) {
  const int W = (int)g.W, band = g.band, m = g.m;
  const int warp = t0 >> 5, lane = t0 & 31, nwarps = nt >> 5;
  for (int li0 = 2 * warp; li0 <= rows; li0 += 2 * nwarps) {
    for (int cb = 0; cb < W; cb += 128) {
      double v[2][4];
      bool ok[2][4];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int i = ibase - (li0 + rr);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = cb + lane + 32 * k;
          const int j = c + i - band;
          ok[rr][k] = (li0 + rr <= rows) && c < W && i > 0 && j >= 1 && j <= m;
          const int64_t off = ok[rr][k] ? (int64_t)(i - 1) * W + c : 0;
          v[rr][k] = __ldg(cells + off);
        }
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int li = li0 + rr, i = ibase - li;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = cb + lane + 32 * k;
          if (li <= rows && c < W) {
            double x = ok[rr][k] ? v[rr][k] : d_inf();
            if (i == 0 && c == band) x = 0.0;  // C[0][0]
            buf[li * W + c] = x;
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kBbThreads) dtw_backtrack_banded_kernel(const double* __restrict__ cells_all,
                                                                          DtwGeom g, int rows,
                                                                          int32_t* __restrict__ path_q,
                                                                          int32_t* __restrict__ path_r,
                                                                          double* __restrict__ path_c,
                                                                          int64_t path_cap,
                                                                          DtwPairOut* __restrict__ outs) {
  extern __shared__ double bb_smem[];
  __shared__ int s_i, s_j, s_done;
  __shared__ int64_t s_len;
  const int W = (int)g.W;
  const int blk = (rows + 1) * W;
  double* bufs[2] = {bb_smem, bb_smem + blk};
  const int pair = blockIdx.x;
  const double* __restrict__ cells = cells_all + (int64_t)pair * g.cells;
  int32_t* pq = path_q + (int64_t)pair * path_cap;
  int32_t* pr = path_r + (int64_t)pair * path_cap;
  double* pc = path_c + (int64_t)pair * path_cap;
  if (threadIdx.x == 0) {
    s_i = g.n;
    s_j = g.m;
    s_len = 0;
    s_done = 0;
  }
  int ibase = g.n;
  bb_load_block(cells, g, ibase, rows, bufs[0], threadIdx.x, kBbThreads);
  __syncthreads();
  for (int k = 0;; ++k) {
    const double* cur = bufs[k & 1];
    if (threadIdx.x >= 32) {
      if (ibase - rows > 0 || (ibase - rows == 0))  // a further block exists (it may only contain row 0)
        bb_load_block(cells, g, ibase - rows, rows, bufs[(k + 1) & 1], threadIdx.x - 32, kBbThreads - 32);
    } else if (threadIdx.x == 0) {
      int i = s_i, j = s_j;
      int64_t len = s_len;
      const int band = g.band, floor_i = ibase - rows;
      auto get = [&](int ii, int jj) -> double {
        const int c = jj - ii + band;
        if (c < 0 || c >= W) return d_inf();
        return cur[(ibase - ii) * W + c];
      };
      while ((i > 0 || j > 0) && (i == 0 || i - 1 >= floor_i)) {
        double cost = 0.0, cv = 0.0, ch = 0.0, cd = 0.0;
        if (i > 0 && j > 0) {
          cv = get(i - 1, j);
          ch = get(i, j - 1);
          cd = get(i - 1, j - 1);
          cost = get(i, j) - cd;
        }
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
          int mi = 0;
          double best = cv;
          if (ch < best) {
            mi = 1;
            best = ch;
          }
          if (cd < best) mi = 2;
          if (mi != 1) i = i - 1;
          if (mi != 0) j = j - 1;
        }
      }
      s_i = i;
      s_j = j;
      s_len = len;
      if (i <= 0 && j <= 0) s_done = 1;
    }
    __syncthreads();
    if (s_done) break;
    ibase -= rows;
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
  const int in_smem = line_bytes <= 140 * 1024;
  if (!in_smem && !line_scratch) return set_error(SONAR_ERR_INVALID, "DTW line scratch missing");
  int diag = g.n < g.m ? g.n : g.m;
  if (g.band > 0 && diag > g.band + 1) diag = g.band + 1;
  int threads = (diag + 31) & ~31;
  threads = threads < 32 ? 32 : (threads > 1024 ? 1024 : threads);
  if (dim == 1 && g.band > 0 && g.band <= 511) {
    const int thr = ((g.band + 1) + 31) & ~31;
    const size_t sm = sizeof(double) * (2 * kRing + 2 * (size_t)g.band + 3);
    prof_begin("dtw_fill_band1_kernel", st);
    if (step == SONAR_STEP_SYMMETRIC2)
      dtw_fill_band1_kernel<SONAR_STEP_SYMMETRIC2><<<n_pairs, thr, sm, st>>>(q, r, g, cells);
    else if (step == SONAR_STEP_ASYMMETRIC)
      dtw_fill_band1_kernel<SONAR_STEP_ASYMMETRIC><<<n_pairs, thr, sm, st>>>(q, r, g, cells);
    else
      dtw_fill_band1_kernel<SONAR_STEP_SYMMETRIC1><<<n_pairs, thr, sm, st>>>(q, r, g, cells);
    prof_end();
    SONAR_CUDA(cudaGetLastError());
  } else {
  const bool staged = dim == 1 && g.band > 0 && g.band <= 1000;
  const bool onecell = diag <= 1024;
  int stage_diags = 0;
  if (onecell) {
    stage_diags = kStageBytes / (int)(sizeof(double) * threads);
    if (stage_diags > 64) stage_diags = 64;
    if (stage_diags < 1) stage_diags = 1;
  }
  const size_t smem = (in_smem ? line_bytes : 0) + (staged ? sizeof(double) * 2 * kRing : 0) +
                      (onecell ? sizeof(double) * (size_t)stage_diags * threads : 0);
#define SONAR_DTW_FILL(S, O)                                                                                    \
  do {                                                                                                          \
    SONAR_CUDA(cudaFuncSetAttribute(dtw_fill_kernel<S, O>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                    (int)(224 * 1024)));                                                        \
    prof_begin("dtw_fill_kernel", st);                                                                          \
    dtw_fill_kernel<S, O><<<n_pairs, threads, smem, st>>>(q, r, g, dim, step, cells, line_scratch, in_smem,    \
                                                          stage_diags);                                         \
  } while (0)
  if (staged && onecell)
    SONAR_DTW_FILL(true, true);
  else if (onecell)
    SONAR_DTW_FILL(false, true);
  else
    SONAR_DTW_FILL(false, false);
#undef SONAR_DTW_FILL
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  }
  const int bb_rows = g.band > 0 ? (int)((100 * 1024 / sizeof(double)) / (size_t)g.W) - 1 : 0;
  if (bb_rows >= 16) {
    const int rows = bb_rows > 64 ? 64 : bb_rows;
    const size_t bsm = sizeof(double) * 2 * (size_t)(rows + 1) * (size_t)g.W;
    SONAR_CUDA(cudaFuncSetAttribute(dtw_backtrack_banded_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(216 * 1024)));
    prof_begin("dtw_backtrack_banded_kernel", st);
    dtw_backtrack_banded_kernel<<<n_pairs, kBbThreads, bsm, st>>>(cells, g, rows, path_q, path_r, path_c, path_cap,
                                                                  out);
    prof_end();
    SONAR_CUDA(cudaGetLastError());
    return SONAR_OK;
  }
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
