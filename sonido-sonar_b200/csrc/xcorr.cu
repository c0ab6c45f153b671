// Normalised cross-correlation over +-maxLag (float64, reference summation order).
//
//   CrossCorrelation.normalize            algorithms/stats/correlation.go:464-501
//   computeTimeDomain                     :203-228
//   normalizedCrossCorrelation            :373-409
//   calculateOverlapRegion                :421-449
//   findPeak / SNR / sharpness / second peak / peak-to-sidelobe   :526-661
//
// The detected lag index must be bit-identical to the reference (north star), and the
// reference's arg-max runs over values that are each a *sequential* float64 sum.  The
// kernels here therefore reproduce that order exactly (compiled with -fmad=false):
//   * znorm_kernel: one warp per sequence; the warp stages a tile in shared memory with
//     coalesced loads, lane 0 replays the reference's left-to-right sums (mean, then
//     population variance), all lanes write the z-scores;
//   * ncc_exact_kernel: one thread per lag, i ascending; consecutive lags sit in
//     consecutive lanes so a[i] is a broadcast and b[i+lag] a coalesced load (mirrored
//     for negative lags).  Every correlation value is bit-exact, so the arg-max needs no
//     guard band;
//   * xcorr_finalize_kernel: one CTA per pair (or lag shard): first-index arg-max of |c|,
//     then the peak-relative reductions (noise power outside +-5, side lobe outside +-10,
//     second peak, neighbours of the peak).
#include <cmath>

#include "common.h"

namespace sonar {
namespace {

constexpr int kZTile = 2048;

__global__ void __launch_bounds__(32) znorm_kernel(const XcorrSeq* __restrict__ seqs) {
  __shared__ double tile[kZTile];
  const XcorrSeq q = seqs[blockIdx.x];
  const double* __restrict__ x = q.in;
  double* __restrict__ z = q.out;
  const int64_t n = q.n;
  const int lane = threadIdx.x;
  double acc = 0.0;
  for (int64_t base = 0; base < n; base += kZTile) {
    const int cnt = (int)((n - base < kZTile) ? (n - base) : kZTile);
    for (int i = lane; i < cnt; i += 32) tile[i] = x[base + i];
    __syncwarp();
    if (lane == 0) {
#pragma unroll 8
      for (int i = 0; i < cnt; ++i) acc += tile[i];
    }
    __syncwarp();
  }
  const double mean = __shfl_sync(0xffffffffu, acc, 0) / (double)n;
  acc = 0.0;
  for (int64_t base = 0; base < n; base += kZTile) {
    const int cnt = (int)((n - base < kZTile) ? (n - base) : kZTile);
    for (int i = lane; i < cnt; i += 32) {
      const double d = x[base + i] - mean;
      tile[i] = d * d;
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll 8
      for (int i = 0; i < cnt; ++i) acc += tile[i];
    }
    __syncwarp();
  }
  const double var = __shfl_sync(0xffffffffu, acc, 0) / (double)n;
  const double sd = sqrt(var);
  if (sd < 1e-10) {
    for (int64_t i = lane; i < n; i += 32) z[i] = x[i] - mean;
  } else {
    for (int64_t i = lane; i < n; i += 32) z[i] = (x[i] - mean) / sd;
  }
}

constexpr int kNccThreads = 128;

__global__ void __launch_bounds__(kNccThreads) ncc_exact_kernel(const XcorrPair* __restrict__ pairs) {
  const XcorrPair p = pairs[blockIdx.y];
  const int64_t j = p.idx_lo + (int64_t)blockIdx.x * kNccThreads + threadIdx.x;
  if (j >= p.idx_hi) return;
  const int64_t lag = j - p.aml;
  const int64_t na = p.na, nb = p.nb;
  // calculateOverlapRegion (correlation.go:421-449)
  int64_t s1, s2, len;
  if (lag >= 0) {
    s1 = 0;
    s2 = lag;
    len = na < nb - lag ? na : nb - lag;
  } else {
    s1 = -lag;
    s2 = 0;
    const int64_t e2 = nb < na + lag ? nb : na + lag;
    len = (na + lag) < e2 ? (na + lag) : e2;
  }
  double c = 0.0;
  if (len > 0) {
    const double* __restrict__ a = p.za + s1;
    const double* __restrict__ b = p.zb + s2;
    double sum = 0.0, q1 = 0.0, q2 = 0.0;
#pragma unroll 4
    for (int64_t i = 0; i < len; ++i) {
      const double v1 = a[i], v2 = b[i];
      sum += v1 * v2;
      q1 += v1 * v1;
      q2 += v2 * v2;
    }
    const double den = sqrt(q1 * q2);
    c = den < 1e-10 ? 0.0 : sum / den;
  }
  p.corr[j - p.idx_lo] = c;
}

struct PeakKey {
  double a;
  int64_t i;
};
__device__ __forceinline__ bool better(double a, int64_t i, double b, int64_t k) {
  // findPeak scans ascending with strict '>' on |c|: larger |c| wins, ties -> smaller index
  if (k < 0) return i >= 0;
  if (i < 0) return false;
  return a > b || (a == b && i < k);
}

constexpr int kFinThreads = 256;

__device__ void block_best(double& a, int64_t& i, double* sa, int64_t* si) {
  const int t = threadIdx.x;
  sa[t] = a;
  si[t] = i;
  __syncthreads();
  for (int o = kFinThreads / 2; o > 0; o >>= 1) {
    if (t < o && better(sa[t + o], si[t + o], sa[t], si[t])) {
      sa[t] = sa[t + o];
      si[t] = si[t + o];
    }
    __syncthreads();
  }
  a = sa[0];
  i = si[0];
  __syncthreads();
}

__device__ double block_sum(double v, double* sa) {
  const int t = threadIdx.x;
  sa[t] = v;
  __syncthreads();
  for (int o = kFinThreads / 2; o > 0; o >>= 1) {
    if (t < o) sa[t] += sa[t + o];
    __syncthreads();
  }
  const double r = sa[0];
  __syncthreads();
  return r;
}

__device__ double block_max(double v, double* sa) {
  const int t = threadIdx.x;
  sa[t] = v;
  __syncthreads();
  for (int o = kFinThreads / 2; o > 0; o >>= 1) {
    if (t < o) sa[t] = fmax(sa[t], sa[t + o]);
    __syncthreads();
  }
  const double r = sa[0];
  __syncthreads();
  return r;
}

// peak_override < 0: use the shard's own arg-max; otherwise the given global index.
__global__ void __launch_bounds__(kFinThreads) xcorr_finalize_kernel(const XcorrPair* __restrict__ pairs,
                                                                     int64_t peak_override,
                                                                     XcorrPairOut* __restrict__ outs) {
  __shared__ double sa[kFinThreads];
  __shared__ int64_t si[kFinThreads];
  const XcorrPair p = pairs[blockIdx.x];
  const double* __restrict__ c = p.corr;
  const int64_t lo = p.idx_lo, hi = p.idx_hi;
  const int t = threadIdx.x;
  // ---- local arg-max of |c| ----
  double ba = 0.0;
  int64_t bi = -1;
  for (int64_t i = lo + t; i < hi; i += kFinThreads) {
    const double v = fabs(c[i - lo]);
    if (better(v, i, ba, bi)) {
      ba = v;
      bi = i;
    }
  }
  block_best(ba, bi, sa, si);
  const int64_t local_peak = bi;
  const int64_t pk = peak_override >= 0 ? peak_override : local_peak;
  // ---- peak-relative reductions ----
  double ns = 0.0, nc = 0.0, ms = 0.0, s2a = 0.0;
  int64_t s2i = -1;
  for (int64_t i = lo + t; i < hi; i += kFinThreads) {
    const double v = c[i - lo], av = fabs(v);
    const int64_t d = i > pk ? i - pk : pk - i;
    if (d > 5) {
      ns += v * v;
      nc += 1.0;
    }
    if (d > 10) ms = fmax(ms, av);
    // findSecondPeak starts from 0.0 with strict '>', so zeros never qualify
    if (i != pk && av > 0.0 && better(av, i, s2a, s2i)) {
      s2a = av;
      s2i = i;
    }
  }
  ns = block_sum(ns, sa);
  nc = block_sum(nc, sa);
  ms = block_max(ms, sa);
  block_best(s2a, s2i, sa, si);
  if (t == 0) {
    XcorrPairOut o;
    o.peak_index = local_peak;
    o.peak = (local_peak >= 0) ? c[local_peak - lo] : 0.0;
    o.noise_sum = ns;
    o.noise_cnt = nc;
    o.max_sidelobe = ms;
    o.second_abs = s2a;
    o.second_index = s2i;
    o.second_val = s2i >= 0 ? c[s2i - lo] : 0.0;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    o.c_peak = (pk >= lo && pk < hi) ? c[pk - lo] : nan;
    o.c_prev = (pk - 1 >= lo && pk - 1 < hi) ? c[pk - 1 - lo] : nan;
    o.c_next = (pk + 1 >= lo && pk + 1 < hi) ? c[pk + 1 - lo] : nan;
    o.n_candidates = (int32_t)((hi - lo) > 0x7fffffff ? 0x7fffffff : (hi - lo));
    o.pad = 0;
    outs[blockIdx.x] = o;
  }
}

__global__ void xcorr_trim_kernel(const XcorrSeq* __restrict__ seqs, const XcorrPair* __restrict__ pairs,
                                  const XcorrPairOut* __restrict__ outs, int n_pairs, const double** __restrict__ qptr,
                                  const double** __restrict__ rptr) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const int64_t lag = outs[p].peak_index - pairs[p].aml;
  qptr[p] = seqs[2 * p].in + (lag < 0 ? -lag : 0);
  rptr[p] = seqs[2 * p + 1].in + (lag >= 0 ? lag : 0);
}

}  // namespace

int launch_xcorr_trim(const XcorrSeq* seqs_dev, const XcorrPair* pairs_dev, const XcorrPairOut* outs_dev, int n_pairs,
                      const double** qptr_dev, const double** rptr_dev, cudaStream_t st) {
  if (n_pairs <= 0) return SONAR_OK;
  prof_begin("xcorr_trim_kernel", st);
  xcorr_trim_kernel<<<(n_pairs + 127) / 128, 128, 0, st>>>(seqs_dev, pairs_dev, outs_dev, n_pairs, qptr_dev, rptr_dev);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

int launch_znorm(const XcorrSeq* seqs_dev, int count, cudaStream_t st) {
  if (count <= 0) return SONAR_OK;
  prof_begin("znorm_kernel", st);
  znorm_kernel<<<count, 32, 0, st>>>(seqs_dev);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

int launch_xcorr(const XcorrPair* pairs_dev, int n_pairs, int64_t max_shard_lags, cudaStream_t st) {
  if (n_pairs <= 0 || max_shard_lags <= 0) return SONAR_OK;
  dim3 grid((unsigned)((max_shard_lags + kNccThreads - 1) / kNccThreads), (unsigned)n_pairs);
  prof_begin("ncc_exact_kernel", st);
  ncc_exact_kernel<<<grid, kNccThreads, 0, st>>>(pairs_dev);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

int launch_xcorr_finalize(const XcorrPair* pairs_dev, int n_pairs, int64_t peak_override, XcorrPairOut* outs_dev,
                          cudaStream_t st) {
  if (n_pairs <= 0) return SONAR_OK;
  prof_begin("xcorr_finalize_kernel", st);
  xcorr_finalize_kernel<<<n_pairs, kFinThreads, 0, st>>>(pairs_dev, peak_override, outs_dev);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

}  // namespace sonar
