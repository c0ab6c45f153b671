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

#include <algorithm>
#include <type_traits>

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

// Banded cost store, DIAGONAL-major: cell (i, j) lives at (i + j) * Wd + ((i - j + band) >> 1), Wd = band + 1.
// The cells of one anti-diagonal are produced together, so this makes every store of the fill kernels a
// contiguous, fully covered run of sectors.  (A row-major band store turns each cell into an isolated
// 8-byte write: the L2 has to fetch the rest of the sector first, and with ~600 cycles per such write and a
// bounded number of writes in flight per SM the fill ran at one store per ~10 cycles.)
__device__ __forceinline__ int64_t band_index(const DtwGeom& g, int i, int j) {
  return (int64_t)(i + j) * g.W + ((i - j + g.band) >> 1);
}

// implicit borders: C[0][0] = 0, first row/column +Inf, out-of-band +Inf
__device__ __forceinline__ double cell_get(const double* __restrict__ cells, const DtwGeom& g, int i, int j) {
  if (i == 0 && j == 0) return 0.0;
  if (i <= 0 || j <= 0) return d_inf();
  if (g.band > 0) {
    const int df = i - j;
    if (df > g.band || df < -g.band) return d_inf();
    return cells[band_index(g, i, j)];
  }
  return cells[(int64_t)(i - 1) * g.W + (j - 1)];
}

constexpr int kRing = 2048;       // staged q / r window (power of two), doubles each
constexpr int kWRing = 512;       // register wavefront: staged q / r window per warp (needs kWRefill + band + 10 <= kWRing)
constexpr int kWRefill = 256;     // diagonals between two refills of the register wavefront's rings
constexpr int kDtwWarps = kDtwPairsPerCta;  // pairs (= warps) per CTA of the register wavefront (common.h)

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src));
}
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
// math.Min (Go): NaN when either argument is NaN; the costs are sums of non-negative distances, so -0 never occurs
__device__ __forceinline__ double go_min(double a, double b) {
  if (a != a || b != b) return a + b;  // NaN
  return a < b ? a : b;
}

template <bool STAGED, bool ONECELL>
__global__ void __launch_bounds__(1024) dtw_fill_kernel(const double* __restrict__ qs, const double* __restrict__ rs,
                                                        DtwGeom g, int dim, int step, double* __restrict__ cells_all,
                                                        double* __restrict__ line_global, int line_in_smem,
                                                        int stage_diags, const double* const* __restrict__ qptr,
                                                        const double* const* __restrict__ rptr) {
  extern __shared__ double s_mem[];
  double* ring_q = s_mem;
  double* ring_r = s_mem + kRing;
  double* s_stage = s_mem + (STAGED ? 2 * kRing : 0);
  double* s_line = s_stage + (ONECELL ? (size_t)stage_diags * blockDim.x : 0);
  const int pair = blockIdx.x;
  const double* __restrict__ q = qptr ? qptr[pair] : qs + (int64_t)pair * g.n * dim;
  const double* __restrict__ r = rptr ? rptr[pair] : rs + (int64_t)pair * g.m * dim;
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
        cells[band > 0 ? band_index(g, i, j) : (int64_t)(i - 1) * g.W + (j - 1)] = s_stage[e];
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
      double mc;  // Go's math.Min: NaN if either argument is NaN (fmin would drop it)
      if (step == SONAR_STEP_SYMMETRIC2)
        mc = go_min(go_min(v, h), dg);
      else if (step == SONAR_STEP_ASYMMETRIC)
        mc = go_min(v, h);
      else
        mc = go_min(v + 1.0, go_min(h + 1.0, dg));
      const double c = ld + mc;
      line[o] = c;
      if constexpr (ONECELL) {
        s_stage[staged * blockDim.x + threadIdx.x] = c;
      } else {
        cells[band > 0 ? band_index(g, i, j) : (int64_t)(i - 1) * g.W + (j - 1)] = c;
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

// |q - r| for dim == 1.  sqrt(x*x) == |x| in binary floating point whenever x*x neither underflows nor
// overflows; the rare remainder takes the out-of-line exact path (kept out of line so that the compiler
// cannot if-convert an FP64 square-root sequence into the wavefront's dependent chain).
__device__ __noinline__ double dist1_slow(double df) { return sqrt(df * df); }
__device__ __forceinline__ double dist1(double a, double b) {
  const double df = a - b;
  // biased exponent in [400, 1640]  <=>  2^-623 <= |df| < 2^618: df*df is a normal number
  const unsigned e = ((unsigned)__double2hiint(df) >> 20) & 0x7ffu;
  if (e - 400u > 1240u) return dist1_slow(df);
  return fabs(df);
}

// Register-resident wavefront for narrow bands (dim == 1, 2*band+3 <= 32*NPL): ONE warp per pair, lane l
// keeps the latest cost of the NPL offsets k = NPL*l - 1 .. NPL*l + NPL - 2 (k = i - j + band) in registers
// (slot x <-> k = NPL*l + x - 1, so slot 0 of lane 0 and the last slot of lane 31 are never inside the band
// and the values shuffled into them need no masking).  A diagonal only touches offsets of one parity and
// reads the two neighbouring offsets of the other parity, so a step is NPL/2 independent relaxations per
// lane plus ONE warp shuffle for the value that lives in the neighbouring lane — no shared-memory line, no
// block barrier.  The single warp issues in order, so everything that does not depend on the previous
// diagonal is kept off the dependent chain shuffle -> min -> add:
//   * the local distances of both diagonals of an iteration (and their validity, folded in as +Inf) are
//     computed from a register window of q / r that slides by one element per iteration;
//   * |q - r| needs no per-cell range test when the sequences were pre-screened (FAST, see dtw_safe_range);
// The backtrack's choice at every cell (findPreviousStep's strict-'<' scan: vertical, horizontal, diagonal)
// is a function of the finished cost store; dtw_dirs_kernel derives it for all cells in parallel afterwards.
// min of two NON-NEGATIVE doubles (finite, +0 or +Inf) through their bit patterns: IEEE-754 orders them like unsigned
// integers, and an integer compare + select has a third of the latency of DSETP + select on this chip (the FP64 pipe's
// dependent-issue latency is what the fill's chain of (min, min, add) per diagonal consists of).
__device__ __forceinline__ double umin_nonneg(double a, double b) {
  return (unsigned long long)__double_as_longlong(a) < (unsigned long long)__double_as_longlong(b) ? a : b;
}

template <int STEP, bool NANSAFE = false, bool INTMIN = false>
__device__ __forceinline__ double dtw_relax(double v, double hh, double dg, double ld) {
  double mc;
  if (INTMIN) {  // screened input: every cost is a sum of |q - r| >= +0 or the +Inf of a cell outside the band
    if (STEP == SONAR_STEP_SYMMETRIC2)
      mc = umin_nonneg(umin_nonneg(v, hh), dg);
    else if (STEP == SONAR_STEP_ASYMMETRIC)
      mc = umin_nonneg(v, hh);
    else
      mc = umin_nonneg(v + 1.0, umin_nonneg(hh + 1.0, dg));
    return ld + mc;
  }
  if (NANSAFE) {  // inputs that failed the screen may be NaN / Inf: math.Min propagates NaN
    if (STEP == SONAR_STEP_SYMMETRIC2)
      mc = go_min(go_min(v, hh), dg);
    else if (STEP == SONAR_STEP_ASYMMETRIC)
      mc = go_min(v, hh);
    else
      mc = go_min(v + 1.0, go_min(hh + 1.0, dg));
    return ld + mc;
  }
  if (STEP == SONAR_STEP_SYMMETRIC2) {  // min(a, b) as (a < b ? a : b) equals math.Min for non-NaN values
    const double t = v < hh ? v : hh;
    mc = t < dg ? t : dg;
  } else if (STEP == SONAR_STEP_ASYMMETRIC) {
    mc = v < hh ? v : hh;
  } else {
    const double a = v + 1.0, b = hh + 1.0;
    const double t = b < dg ? b : dg;
    mc = a < t ? a : t;
  }
  return ld + mc;
}
// findPreviousStep (dtw.go:191-217): 0 = vertical (i-1, j), 1 = horizontal (i, j-1), 2 = diagonal
__device__ __forceinline__ unsigned dtw_dir(double cv, double ch, double cd) {
  const bool p1 = ch < cv;
  const double best = p1 ? ch : cv;
  return cd < best ? 2u : (p1 ? 1u : 0u);
}

// |x| of every non-zero element in [2^-458, 2^500] (and finite): then every difference a - b is either 0 or
// has a square that is a normal number, so sqrt((a-b)^2) == |a - b| exactly and dist1's range test can go.
// One CTA per pair screens both sequences and leaves the verdict in the pair's flag slot.
__global__ void __launch_bounds__(1024) dtw_screen_kernel(const double* __restrict__ qs, const double* __restrict__ rs,
                                                          DtwGeom g, double* __restrict__ cells_all,
                                                          const double* const* __restrict__ qptr,
                                                          const double* const* __restrict__ rptr) {
  const int pair = blockIdx.x;
  const double* __restrict__ q = qptr ? qptr[pair] : qs + (int64_t)pair * g.n;
  const double* __restrict__ r = rptr ? rptr[pair] : rs + (int64_t)pair * g.m;
  bool ok = true;
  for (int e = threadIdx.x; e < g.n + g.m; e += blockDim.x) {
    const double v = e < g.n ? q[e] : r[e - g.n];
    const unsigned ex = (v == 0.0) ? 1023u : (((unsigned)__double2hiint(v) >> 20) & 0x7ffu);
    ok = ok && (ex - 565u <= 958u);  // biased exponent in [1023-458, 1023+500]
  }
  const int all = __syncthreads_and(ok ? 1 : 0);
  if (threadIdx.x == 0) cells_all[(int64_t)pair * g.cells + g.flag_off] = all ? 1.0 : 0.0;
}

template <int NPL, int STEP, bool FAST>
__device__ __forceinline__ void dtw_fill_warp_body(const double* __restrict__ q, const double* __restrict__ r,
                                                   const DtwGeom& g, double* __restrict__ cells, double* ring_q,
                                                   double* ring_r) {
  constexpr int H = NPL / 2;
  const int lane = threadIdx.x & 31;
  const int n = g.n, m = g.m, band = g.band, W = (int)g.W;
  const double inf = d_inf();
  const int kbase = NPL * lane - 1;  // offset of slot 0
  double L[NPL];
  int dlo[NPL];
  unsigned span[NPL];
  bool inband[NPL];
#pragma unroll
  for (int x = 0; x < NPL; ++x) {
    const int k = kbase + x, delta = k - band;
    L[x] = (k == band) ? 0.0 : inf;  // C[0][0] sits at offset i - j = 0
    const int lo = 2 + (delta < 0 ? -delta : delta);
    const int hi = (2 * n - delta) < (2 * m + delta) ? (2 * n - delta) : (2 * m + delta);
    inband[x] = k >= 0 && k <= 2 * band;
    const bool any = inband[x] && hi >= lo;
    dlo[x] = any ? lo : 0x3fffffff;  // cell of offset k on diagonal d is inside the matrix iff lo <= d <= hi
    span[x] = any ? (unsigned)(hi - lo) : 0u;
  }
  const int last = n + m;
  // q / r reach the rings through cp.async one refill period AHEAD of their use, so the warp never waits for
  // a global load: the call at diagonal d first waits for the copies issued at d - kWRefill (everything the
  // diagonals [d, d + kWRefill) read), then issues the elements of the following period.
  int loaded = 0;
  auto issue = [&](int target) {
    for (int e = loaded + lane; e < target; e += 32) {
      if (e < n) cp_async8(&ring_q[e & (kWRing - 1)], q + e);
      if (e < m) cp_async8(&ring_r[e & (kWRing - 1)], r + e);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    loaded = target > loaded ? target : loaded;
  };
  auto refill = [&](int d) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    issue(((d + 2 * kWRefill + band) >> 1) + 8);
  };
  // The single warp is issue-bound (one instruction every ~2 cycles), so the loop body is kept minimal: the
  // invalid slots get their +Inf through the local distance, stores are predicated (no branches), the store
  // pointers advance incrementally, and the steady-state loop below takes the validity flags as invariants.
  auto local = [&](double a, double b, bool ok) -> double {
    if (FAST) return fabs(a - b) + (ok ? 0.0 : inf);  // two DADDs; the penalty is loop-invariant in the steady state
    const double df = dist1(a, b);
    return ok ? df : inf;
  };
  // odd slots (even offsets k = NPL*l + 2h) of a diagonal whose row starts at `row`; store slot (k >> 1) = half_l + h
  auto half_b = [&](double* __restrict__ row, const double (&Q)[H], const double (&R)[H + 1], const bool (&ok)[H]) {
    const double edge = __shfl_down_sync(0xffffffffu, L[0], 1);  // offset k+1 of the last odd slot: lane+1, slot 0
    double c[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const int x = 2 * h + 1;
      const double hh = (h == H - 1) ? edge : L[x + 1 < NPL ? x + 1 : x];
      c[h] = dtw_relax<STEP, !FAST, FAST>(L[x - 1], hh, L[x], local(Q[h], R[h + 1], ok[h]));  // cell (ib+h, jb-h): q[ib+h-1], r[jb-h-1]
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
      L[2 * h + 1] = c[h];  // invalid cells carry +Inf, which is what the offset line must hold there
      if (ok[h]) row[h] = c[h];
    }
  };
  // even slots (odd offsets k = NPL*l + 2h - 1); store slot half_l + h - 1 (row is passed already shifted by -1)
  auto half_a = [&](double* __restrict__ row, const double (&Q)[H], const double (&R)[H + 1], const bool (&ok)[H]) {
    const double edge = __shfl_up_sync(0xffffffffu, L[NPL - 1], 1);  // offset k-1 of slot 0: lane-1, last slot
    double c[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const int x = 2 * h;
      const double v = (h == 0) ? edge : L[x > 0 ? x - 1 : 0];
      c[h] = dtw_relax<STEP, !FAST, FAST>(v, L[x + 1], L[x], local(Q[h], R[h], ok[h]));  // cell (ib+h, jb-h+1): q[ib+h-1], r[jb-h]
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
      L[2 * h] = c[h];
      if (ok[h]) row[h] = c[h];
    }
  };
  const int half_l = (NPL / 2) * lane;  // (NPL*lane) >> 1
  int d = 2;
  issue(((d + kWRefill + band) >> 1) + 8);
  refill(d);
  int next_refill = d + kWRefill;
  // Cells of an aligned iteration (d - band even): odd slots on d are (ib + h, jb - h), even slots on d + 1 are
  // (ib + h, jb - h + 1), with ib = (d - band + NPL*lane) / 2, jb = d - ib; Q[h] = q[ib + h - 1], R[u] = r[jb - u].
  double Q[H], R[H + 1];
  if ((d - band) & 1) {  // odd band: diagonal 2 carries the odd offsets (even slots); do it alone
    const int ib = (d - 1 - band + NPL * lane) >> 1, jb = d - 1 - ib;
    bool ok[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      ok[h] = (unsigned)(d - dlo[2 * h]) <= span[2 * h];
      Q[h] = ring_q[(ib + h - 1) & (kWRing - 1)];
    }
#pragma unroll
    for (int u = 0; u <= H; ++u) R[u] = ring_r[(jb - u) & (kWRing - 1)];
    half_a(cells + (int64_t)d * W + half_l - 1, Q, R, ok);
    ++d;
  }
  int ib = (d - band + NPL * lane) >> 1, jb = d - ib;
#pragma unroll
  for (int h = 0; h < H; ++h) Q[h] = ring_q[(ib + h - 1) & (kWRing - 1)];
#pragma unroll
  for (int u = 0; u <= H; ++u) R[u] = ring_r[(jb - u) & (kWRing - 1)];
  double* __restrict__ rowb = cells + (int64_t)d * W + half_l;  // diagonal d, odd slots; diagonal d + 1's even slots: + W - 1
  // steady state: every in-band offset has a cell on the diagonal; [sd0, sd1) in steps of two from d
  const int mn = n < m ? n : m;
  const int sd0 = band + 4 + ((band + 4 - d) & 1), sd1 = 2 * mn - band - 2;
  auto iterate = [&](auto steady_tag) {
    constexpr bool STEADY = decltype(steady_tag)::value;
    if (d >= next_refill) {
      refill(d);
      next_refill += kWRefill;
    }
    const double qn = ring_q[(ib + H - 1) & (kWRing - 1)], rn = ring_r[(jb + 1) & (kWRing - 1)];  // next iteration's
    bool okb[H], oka[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      okb[h] = STEADY ? inband[2 * h + 1] : ((unsigned)(d - dlo[2 * h + 1]) <= span[2 * h + 1]);
      oka[h] = STEADY ? inband[2 * h] : ((unsigned)(d + 1 - dlo[2 * h]) <= span[2 * h]);
    }
    half_b(rowb, Q, R, okb);
    half_a(rowb + W - 1, Q, R, oka);  // beyond the last diagonal every cell is invalid: nothing is stored
#pragma unroll
    for (int h = 0; h + 1 < H; ++h) Q[h] = Q[h + 1];
    Q[H - 1] = qn;
#pragma unroll
    for (int u = H; u > 0; --u) R[u] = R[u - 1];
    R[0] = rn;
    d += 2;
    ++ib;
    ++jb;
    rowb += 2 * W;
  };
  while (d <= last && d < sd0) iterate(std::false_type{});
  {
    const int n_steady = d + 1 < sd1 ? (sd1 - d) / 2 : 0;  // iterations with d + 1 < sd1
#pragma unroll 6  // the q / r register windows rotate with periods 2 and 3 (NPL = 4): no moves left after unrolling
    for (int it = 0; it < n_steady; ++it) iterate(std::true_type{});
  }
  while (d <= last) iterate(std::false_type{});
}

// kDtwWarps pairs per CTA, one warp each (a warp per scheduler: the chains do not compete for issue slots).  A CTA per
// pair put 32 single-warp CTAs with 32 KB of rings each on 32 SMs, and the persistent STFT kernels, whose CTAs need a
// whole SM's shared memory, ran on the remaining 116 SMs for as long as the fill lasted (+1 ms on the STFT pair in the
// step); eight CTAs with 8 KB of rings per warp take 8 SMs and fit beside a pitch-kernel CTA.
template <int NPL, int STEP>
__global__ void __launch_bounds__(32 * kDtwWarps) dtw_fill_warp_kernel(const double* __restrict__ qs, const double* __restrict__ rs,
                                                                       DtwGeom g, double* __restrict__ cells_all,
                                                                       const double* const* __restrict__ qptr,
                                                                       const double* const* __restrict__ rptr, int n_pairs) {
  __shared__ double rings[kDtwWarps][2][kWRing];
  const int w = threadIdx.x >> 5;
  const int pair = blockIdx.x * kDtwWarps + w;
  if (pair >= n_pairs) return;  // the body synchronises warps only
  const double* __restrict__ q = qptr ? qptr[pair] : qs + (int64_t)pair * g.n;
  const double* __restrict__ r = rptr ? rptr[pair] : rs + (int64_t)pair * g.m;
  double* cells = cells_all + (int64_t)pair * g.cells;
  if (cells[g.flag_off] != 0.0)  // warp-uniform verdict of dtw_screen_kernel
    dtw_fill_warp_body<NPL, STEP, true>(q, r, g, cells, rings[w][0], rings[w][1]);
  else
    dtw_fill_warp_body<NPL, STEP, false>(q, r, g, cells, rings[w][0], rings[w][1]);
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

// Banded backtrack.  The store is diagonal-major, a step moves to diagonal d-1 (vertical / horizontal) or
// d-2 (diagonal), and which diagonals come next does not depend on the path: the walk proceeds through
// blocks of `nd` consecutive diagonals held in shared memory (borders normalised to +Inf / C[0][0] = 0
// while loading, one +Inf sentinel slot on each side of every diagonal so the walker needs no bounds
// checks), and warps 1..7 prefetch block k+1 — one contiguous, coalesced region of HBM — into the other
// buffer while lane 0 of warp 0 walks block k.
constexpr int kBbThreads = 256;

__device__ __forceinline__ void bb_load_block(const double* __restrict__ cells, const DtwGeom& g, int dbase, int nd,
                                              double* __restrict__ buf, int t0, int nt) {
  const int Wd = (int)g.W, Wp = Wd + 2, band = g.band;
  const int total = (nd + 1) * Wp;
  for (int e0 = t0; e0 < total; e0 += 4 * nt) {
    double v[4];
    bool ok[4], zero[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * nt;
      const int dl = e / Wp, kk = e - dl * Wp - 1;  // kk == -1 / Wd are the sentinel slots
      const int d = dbase - dl;
      const int s2 = 2 * kk + ((d + band) & 1);     // i - j + band
      const int i = (d + s2 - band) >> 1, j = d - i;
      ok[u] = e < total && kk >= 0 && kk < Wd && s2 <= 2 * band && d >= 2 && i >= 1 && i <= g.n && j >= 1 && j <= g.m;
      zero[u] = e < total && d == 0 && s2 == band;  // C[0][0]
      v[u] = __ldg(cells + (ok[u] ? (int64_t)d * Wd + kk : 0));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * nt;
      if (e < total) buf[e] = zero[u] ? 0.0 : (ok[u] ? v[u] : d_inf());
    }
  }
}

__global__ void __launch_bounds__(kBbThreads) dtw_backtrack_banded_kernel(const double* __restrict__ cells_all,
                                                                          DtwGeom g, int nd,
                                                                          int32_t* __restrict__ path_q,
                                                                          int32_t* __restrict__ path_r,
                                                                          double* __restrict__ path_c,
                                                                          int64_t path_cap,
                                                                          DtwPairOut* __restrict__ outs,
                                                                          int only_flagged) {
  extern __shared__ double bb_smem[];
  if (only_flagged && outs[blockIdx.x].path_len >= 0) return;  // the parallel backtrack already produced this path
  __shared__ int s_i, s_j, s_done;
  __shared__ int64_t s_len;
  const int Wd = (int)g.W, Wp = Wd + 2;
  const int blk = (nd + 1) * Wp;
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
  int dbase = g.n + g.m;
  bb_load_block(cells, g, dbase, nd, bufs[0], threadIdx.x, kBbThreads);
  __syncthreads();
  for (int k = 0;; ++k) {
    const double* cur = bufs[k & 1];
    if (threadIdx.x >= 32) {
      if (dbase - nd + 1 >= 0)  // the next block starts one diagonal above this block's floor
        bb_load_block(cells, g, dbase - nd + 1, nd, bufs[(k + 1) & 1], threadIdx.x - 32, kBbThreads - 32);
    } else if (threadIdx.x == 0) {
      int i = s_i, j = s_j;
      int64_t pos = path_cap - 1 - s_len;
      const int band = g.band;
      const int floor_d = dbase - nd;  // lowest diagonal held by this block
      // a cell reads diagonals d-1 and d-2: stay while d - 2 >= floor_d (or no cell value is needed: i == 0 / j == 0)
      while ((i > 0 || j > 0) && (i == 0 || j == 0 || i + j - 2 >= floor_d)) {
        {  // fast path: interior cell inside the band, indices maintained incrementally
          int s2 = i - j + band;
          if (i > 1 && j > 1 && s2 >= 0 && s2 <= 2 * band) {
            int idx = (dbase - (i + j)) * Wp + (s2 >> 1) + 1;
            int dd = i + j;
            do {
              const int b0 = s2 & 1;
              const double cij = cur[idx], cv = cur[idx + Wp - 1 + b0], ch = cur[idx + Wp + b0], cd = cur[idx + 2 * Wp];
              if (pos >= 0) {
                pq[pos] = i - 1;
                pr[pos] = j - 1;
                pc[pos] = cij - cd;
              }
              --pos;
              double best = cv;
              int step_idx = Wp - 1 + b0, di = 1, dj = 0, ds = -1;
              if (ch < best) {
                best = ch;
                step_idx = Wp + b0, di = 0, dj = 1, ds = 1;
              }
              if (cd < best) step_idx = 2 * Wp, di = 1, dj = 1, ds = 0;
              idx += step_idx;
              i -= di;
              j -= dj;
              s2 += ds;
              dd -= di + dj;
            } while (i > 1 && j > 1 && (unsigned)s2 <= (unsigned)(2 * band) && dd - 2 >= floor_d);
            continue;
          }
        }
        double cost = 0.0;
        int mi;
        if (i > 0 && j > 0) {
          const int s2 = i - j + band;
          double cij = d_inf(), cv = d_inf(), ch = d_inf(), cd = d_inf();
          if (s2 >= -1 && s2 <= 2 * band + 1) {  // within one step of the band: slots (incl. sentinels) exist
            const int b0 = s2 & 1, kk = s2 >> 1;  // arithmetic shift: s2 == -1 -> kk == -1 (sentinel)
            const int idx = (dbase - (i + j)) * Wp + kk + 1;
            if (s2 >= 0 && s2 <= 2 * band) {
              cij = cur[idx];
              cd = cur[idx + 2 * Wp];
            }
            cv = cur[idx + Wp - 1 + b0];  // (i-1, j): offset s2 - 1 on diagonal d - 1
            ch = cur[idx + Wp + b0];      // (i, j-1): offset s2 + 1 on diagonal d - 1
          }
          cost = cij - cd;
          mi = 0;
          double best = cv;
          if (ch < best) {
            mi = 1;
            best = ch;
          }
          if (cd < best) mi = 2;
        } else {
          mi = (i == 0) ? 1 : 0;
        }
        if (pos >= 0) {
          pq[pos] = i - 1;
          pr[pos] = j - 1;
          pc[pos] = cost;
        }
        --pos;
        if (mi != 1) --i;
        if (mi != 0) --j;
      }
      s_i = i;
      s_j = j;
      s_len = path_cap - 1 - pos;
      if (i <= 0 && j <= 0) s_done = 1;
    }
    __syncthreads();
    if (s_done) break;
    dbase -= nd - 1;
  }
  if (threadIdx.x == 0) {
    outs[pair].total_cost = cell_get(cells, g, g.n, g.m);
    outs[pair].path_len = s_len;
  }
}

// ------------------------------------------------------------------------------------------------
// Parallel backtrack over the direction bytes the register wavefront recorded.
//
// The predecessor of a cell is a function of the cell alone, so the walk from (n, m) is a pointer chase
// through a forest whose edges are already known.  The diagonals n+m .. 0 are cut into blocks of
// kDtwBtDiags; a step moves down one or two diagonals, so a walk enters a block on its top diagonal or the
// one below it, in one of W slots: 2W possible entry states per block.
//   1. dtw_bt_exits_kernel  — one CTA per (block, pair): the block's bytes are staged in shared memory and
//      one thread per entry state walks to the block's floor, recording the state in which it enters the
//      next block and the number of steps taken (all blocks and states in parallel).
//   2. dtw_bt_chain_kernel  — one thread per pair hops from block to block through those tables starting at
//      (n, m): entry state and path position of every block, total path length.
//   3. dtw_bt_emit_kernel   — one CTA per (block, pair) re-walks its block from the now known entry state
//      and writes the path indices at their final positions.
//   4. dtw_bt_cost_kernel   — point.Cost = C[i][j] - C[i-1][j-1] for every path point, fully parallel.
// A walk that leaves the band or the matrix (only possible when |n - m| > band or the costs are not
// finite: the reference then follows +Inf/NaN comparisons) flags the pair (path_len = -1) and the
// cell-comparing walker below handles it.
// One direction byte per stored cell (same diagonal-major index as the cost store).
__global__ void __launch_bounds__(256) dtw_dirs_kernel(double* __restrict__ cells_all, DtwGeom g) {
  const int pair = blockIdx.y, W = (int)g.W, band = g.band;
  double* cells = cells_all + (int64_t)pair * g.cells;
  unsigned char* dirs = reinterpret_cast<unsigned char*>(cells + g.dirs_off);
  const int last = g.n + g.m;
  for (int d = 2 + blockIdx.x * 4 + (threadIdx.x >> 6); d <= last; d += gridDim.x * 4) {
    const int par = (d + band) & 1;
    for (int kk = threadIdx.x & 63; kk < W; kk += 64) {
      const int s2 = 2 * kk + par;
      const int i = (d + s2 - band) >> 1, j = d - i;
      if (s2 > 2 * band || i < 1 || i > g.n || j < 1 || j > g.m) continue;
      dirs[(int64_t)d * W + kk] =
          (unsigned char)dtw_dir(cell_get(cells, g, i - 1, j), cell_get(cells, g, i, j - 1), cell_get(cells, g, i - 1, j - 1));
    }
  }
}

constexpr int kBtDone = 0xfffe, kBtBail = 0xffff;

struct BtState {
  int i, j;
};
// entry state c of the block whose top diagonal is dtop: c < W -> (dtop, slot c), else (dtop - 1, slot c - W)
__device__ __forceinline__ bool bt_state_cell(const DtwGeom& g, int dtop, int c, BtState* st) {
  const int W = (int)g.W;
  const int d = c < W ? dtop : dtop - 1, slot = c < W ? c : c - W;
  const int s2 = 2 * slot + ((d + g.band) & 1);
  st->i = (d + s2 - g.band) >> 1;
  st->j = d - st->i;
  if (d == 0 && s2 == g.band) return true;  // (0, 0)
  return d >= 2 && s2 <= 2 * g.band && st->i >= 1 && st->i <= g.n && st->j >= 1 && st->j <= g.m;
}

// Walks from (i, j) down to below diagonal dbot.  sm holds the bytes of diagonals dbot .. (row (d - dbot) * W).
// Returns the exit code; EMIT writes the path points at pos, pos-1, ...
template <bool EMIT>
__device__ __forceinline__ int bt_walk(const DtwGeom& g, const unsigned char* __restrict__ sm, int dbot, int i, int j,
                                       int* steps_out, int32_t* __restrict__ pq, int32_t* __restrict__ pr,
                                       int64_t pos) {
  const int W = (int)g.W, band = g.band;
  int steps = 0, code;
  for (;;) {
    const int d = i + j, s2 = i - j + band;
    if (i == 0 && j == 0) {
      code = kBtDone;
      break;
    }
    if (i < 1 || j < 1 || (unsigned)s2 > (unsigned)(2 * band)) {
      code = kBtBail;
      break;
    }
    if (d < dbot) {
      code = (d == dbot - 1) ? (s2 >> 1) : W + (s2 >> 1);
      break;
    }
    const unsigned dir = sm[(d - dbot) * W + (s2 >> 1)];
    if (EMIT) {
      if (pos >= 0) {
        pq[pos] = i - 1;
        pr[pos] = j - 1;
      }
      --pos;
    }
    ++steps;
    i -= (dir != 1u);
    j -= (dir != 0u);
  }
  *steps_out = steps;
  return code;
}

__device__ __forceinline__ void bt_stage_block(const unsigned char* __restrict__ dirs, const DtwGeom& g, int dbot,
                                               int dtop, unsigned char* __restrict__ sm) {
  // bytes of diagonals dbot .. dtop are contiguous in the store; copy as 4-byte words from an aligned base
  const int64_t b0 = (int64_t)dbot * g.W, b1 = (int64_t)(dtop + 1) * g.W;
  const int64_t a0 = b0 & ~(int64_t)3;
  const int nw = (int)((b1 - a0 + 3) >> 2);
  const unsigned* __restrict__ src = reinterpret_cast<const unsigned*>(dirs + a0);
  unsigned* dst = reinterpret_cast<unsigned*>(sm);
  for (int e = threadIdx.x; e < nw; e += blockDim.x) dst[e] = src[e];
}

constexpr int kBtxThreads = 256;

__global__ void __launch_bounds__(kBtxThreads) dtw_bt_exits_kernel(double* __restrict__ cells_all, DtwGeom g) {
  extern __shared__ __align__(16) unsigned char bt_sm[];
  const int blk = blockIdx.x, pair = blockIdx.y, W = (int)g.W;
  double* cells = cells_all + (int64_t)pair * g.cells;
  const unsigned char* dirs = reinterpret_cast<const unsigned char*>(cells + g.dirs_off);
  int* tbl = reinterpret_cast<int*>(cells + g.tbl_off) + (int64_t)blk * 2 * W;
  const int dtop = g.n + g.m - blk * kDtwBtDiags;
  const int dbot = dtop - kDtwBtDiags + 1 > 0 ? dtop - kDtwBtDiags + 1 : 0;
  bt_stage_block(dirs, g, dbot, dtop, bt_sm);
  __syncthreads();
  const unsigned char* sm = bt_sm + (((int64_t)dbot * W) & 3);
  for (int c = threadIdx.x; c < 2 * W; c += blockDim.x) {
    BtState st;
    int steps = 0, code = kBtBail;
    if (bt_state_cell(g, dtop, c, &st)) code = bt_walk<false>(g, sm, dbot, st.i, st.j, &steps, nullptr, nullptr, 0);
    tbl[c] = (steps << 16) | code;
  }
}

__global__ void dtw_bt_chain_kernel(double* __restrict__ cells_all, DtwGeom g, int n_pairs, DtwPairOut* __restrict__ outs) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= n_pairs) return;
  double* cells = cells_all + (int64_t)pair * g.cells;
  const int* tbl = reinterpret_cast<const int*>(cells + g.tbl_off);
  int* chain = reinterpret_cast<int*>(cells + g.chain_off);
  const int W = (int)g.W;
  int64_t len = -1;
  const int s2 = g.n - g.m + g.band;
  if (s2 >= 0 && s2 <= 2 * g.band) {
    int c = s2 >> 1;
    int64_t total = 0;
    for (int b = 0; b < g.bt_nb; ++b) {
      chain[2 * b] = c;
      chain[2 * b + 1] = (int)total;
      const int e = tbl[(int64_t)b * 2 * W + c];
      total += e >> 16;
      const int code = e & 0xffff;
      if (code == kBtDone) {
        len = total;
        for (int bb = b + 1; bb < g.bt_nb; ++bb) chain[2 * bb] = -1;
        break;
      }
      if (code == kBtBail) break;
      c = code;
    }
  }
  outs[pair].path_len = len;
  outs[pair].total_cost = cell_get(cells, g, g.n, g.m);
}

__global__ void __launch_bounds__(kBtxThreads) dtw_bt_emit_kernel(const double* __restrict__ cells_all, DtwGeom g,
                                                                  int32_t* __restrict__ path_q,
                                                                  int32_t* __restrict__ path_r, int64_t path_cap,
                                                                  const DtwPairOut* __restrict__ outs) {
  extern __shared__ __align__(16) unsigned char bt_sm[];
  const int blk = blockIdx.x, pair = blockIdx.y, W = (int)g.W;
  if (outs[pair].path_len < 0) return;
  const double* cells = cells_all + (int64_t)pair * g.cells;
  const int* chain = reinterpret_cast<const int*>(cells + g.chain_off);
  const int c = chain[2 * blk];
  if (c < 0) return;  // the walk ended in an earlier block
  const unsigned char* dirs = reinterpret_cast<const unsigned char*>(cells + g.dirs_off);
  const int dtop = g.n + g.m - blk * kDtwBtDiags;
  const int dbot = dtop - kDtwBtDiags + 1 > 0 ? dtop - kDtwBtDiags + 1 : 0;
  bt_stage_block(dirs, g, dbot, dtop, bt_sm);
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned char* sm = bt_sm + (((int64_t)dbot * W) & 3);
    BtState st;
    bt_state_cell(g, dtop, c, &st);
    int steps;
    bt_walk<true>(g, sm, dbot, st.i, st.j, &steps, path_q + (int64_t)pair * path_cap, path_r + (int64_t)pair * path_cap,
                  path_cap - 1 - (int64_t)chain[2 * blk + 1]);
  }
}

__global__ void dtw_bt_cost_kernel(const double* __restrict__ cells_all, DtwGeom g, const int32_t* __restrict__ path_q,
                                   const int32_t* __restrict__ path_r, double* __restrict__ path_c, int64_t path_cap,
                                   const DtwPairOut* __restrict__ outs) {
  const int pair = blockIdx.y;
  const int64_t len = outs[pair].path_len;
  if (len < 0) return;
  const double* cells = cells_all + (int64_t)pair * g.cells;
  const int64_t take = len < path_cap ? len : path_cap;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < take; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pos = (int64_t)pair * path_cap + path_cap - 1 - e;
    const int i = path_q[pos] + 1, j = path_r[pos] + 1;
    path_c[pos] = cell_get(cells, g, i, j) - cell_get(cells, g, i - 1, j - 1);  // dtw.go:173-175
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
  g->dirs_off = g->tbl_off = g->chain_off = g->flag_off = 0;
  g->bt_nb = 0;
  g->n = n;
  g->m = m;
  g->band = band > 0 ? band : 0;
  if (g->band > 0) {
    g->W = (int64_t)g->band + 1;  // cells per stored anti-diagonal
    g->n_off = 2 * g->band + 1;
    g->cells = ((int64_t)n + m + 1) * g->W;
    if (2 * g->band + 3 <= 32 * 8) {  // the register wavefront applies (dim == 1): room for its side arrays
      g->bt_nb = (n + m) / kDtwBtDiags + 1;
      g->dirs_off = g->cells;
      g->tbl_off = g->dirs_off + (g->cells + 7) / 8 + 2;                           // bytes -> doubles (+ slack for slot -1)
      g->chain_off = g->tbl_off + ((int64_t)g->bt_nb * 2 * g->W + 1) / 2;          // int32 [bt_nb][2W]
      g->flag_off = g->chain_off + (int64_t)g->bt_nb;                              // int32 [bt_nb][2]
      g->cells = g->flag_off + 1;                                                  // screening verdict
    }
  } else {
    g->W = m;  // cells per stored row
    g->n_off = n + m + 1;
    g->cells = (int64_t)n * g->W;
  }
  return SONAR_OK;
}

int launch_dtw(const double* q, const double* r, int n_pairs, const DtwGeom& g, int dim, int step, double* cells,
               double* line_scratch, int32_t* path_q, int32_t* path_r, double* path_c, int64_t path_cap,
               DtwPairOut* out, cudaStream_t st, const double* const* qptr, const double* const* rptr) {
  if (n_pairs <= 0) return SONAR_OK;
  const size_t line_bytes = sizeof(double) * (size_t)(g.n_off + 2);
  const int in_smem = line_bytes <= 140 * 1024;
  if (!in_smem && !line_scratch) return set_error(SONAR_ERR_INVALID, "DTW line scratch missing");
  int diag = g.n < g.m ? g.n : g.m;
  if (g.band > 0 && diag > g.band + 1) diag = g.band + 1;
  int threads = (diag + 31) & ~31;
  threads = threads < 32 ? 32 : (threads > 1024 ? 1024 : threads);
  const bool warp_path = dim == 1 && g.band > 0 && g.dirs_off > 0;
  if (warp_path) {
    // min(a, b) as (a < b ? a : b) equals fmin for the non-NaN values this recurrence produces
#define SONAR_DTW_WARP(NPL)                                                                       \
  do {                                                                                            \
    if (step == SONAR_STEP_SYMMETRIC2)                                                            \
      dtw_fill_warp_kernel<NPL, SONAR_STEP_SYMMETRIC2><<<wgrid, 32 * kDtwWarps, 0, st>>>(q, r, g, cells, qptr, rptr, n_pairs);   \
    else if (step == SONAR_STEP_ASYMMETRIC)                                                       \
      dtw_fill_warp_kernel<NPL, SONAR_STEP_ASYMMETRIC><<<wgrid, 32 * kDtwWarps, 0, st>>>(q, r, g, cells, qptr, rptr, n_pairs);   \
    else                                                                                          \
      dtw_fill_warp_kernel<NPL, SONAR_STEP_SYMMETRIC1><<<wgrid, 32 * kDtwWarps, 0, st>>>(q, r, g, cells, qptr, rptr, n_pairs);   \
  } while (0)
    const int offs = 2 * g.band + 3;  // one never-valid slot on each side (see dtw_fill_warp_body)
    const unsigned wgrid = (unsigned)((n_pairs + kDtwWarps - 1) / kDtwWarps);
    prof_begin("dtw_screen_kernel", st);
    dtw_screen_kernel<<<n_pairs, 1024, 0, st>>>(q, r, g, cells, qptr, rptr);
    prof_end();
    prof_begin("dtw_fill_warp_kernel", st);
    if (offs <= 64)
      SONAR_DTW_WARP(2);
    else if (offs <= 128)
      SONAR_DTW_WARP(4);
    else
      SONAR_DTW_WARP(8);
#undef SONAR_DTW_WARP
    prof_end();
    SONAR_CUDA(cudaGetLastError());
    // parallel backtrack: direction bytes from the finished cost store, then block exits / chain / emit / costs
    prof_begin("dtw_dirs_kernel", st);
    dtw_dirs_kernel<<<dim3(148 * 4, (unsigned)n_pairs), 256, 0, st>>>(cells, g);
    prof_end();
    const size_t bsm = (size_t)kDtwBtDiags * (size_t)g.W + 16;
    const dim3 bgrid((unsigned)g.bt_nb, (unsigned)n_pairs);
    prof_begin("dtw_bt_exits_kernel", st);
    dtw_bt_exits_kernel<<<bgrid, kBtxThreads, bsm, st>>>(cells, g);
    prof_end();
    prof_begin("dtw_bt_chain_kernel", st);
    dtw_bt_chain_kernel<<<(n_pairs + 31) / 32, 32, 0, st>>>(cells, g, n_pairs, out);
    prof_end();
    prof_begin("dtw_bt_emit_kernel", st);
    dtw_bt_emit_kernel<<<bgrid, kBtxThreads, bsm, st>>>(cells, g, path_q, path_r, path_cap, out);
    prof_end();
    const int64_t max_len = (int64_t)g.n + g.m;
    prof_begin("dtw_bt_cost_kernel", st);
    dtw_bt_cost_kernel<<<dim3((unsigned)std::min<int64_t>((max_len + 255) / 256, 64), (unsigned)n_pairs), 256, 0, st>>>(
        cells, g, path_q, path_r, path_c, path_cap, out);
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
                                                          stage_diags, qptr, rptr);                             \
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
  // diagonals per shared-memory block of the banded backtrack (two buffers of (nd + 1) * (band + 3) doubles)
  const int bb_nd = g.band > 0 ? (int)((100 * 1024 / sizeof(double)) / (size_t)(g.W + 2)) - 1 : 0;
  if (bb_nd >= 16) {
    const int nd = bb_nd > 192 ? 192 : bb_nd;
    const size_t bsm = sizeof(double) * 2 * (size_t)(nd + 1) * (size_t)(g.W + 2);
    SONAR_CUDA(cudaFuncSetAttribute(dtw_backtrack_banded_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(216 * 1024)));
    prof_begin("dtw_backtrack_banded_kernel", st);
    dtw_backtrack_banded_kernel<<<n_pairs, kBbThreads, bsm, st>>>(cells, g, nd, path_q, path_r, path_c, path_cap,
                                                                  out, warp_path ? 1 : 0);
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
