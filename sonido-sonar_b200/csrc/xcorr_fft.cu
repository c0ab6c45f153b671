// Screened normalised cross-correlation: an FP64 FFT finds WHERE the peaks are, the reference-order
// kernel (xcorr.cu) then computes those lags exactly.
//
//   CrossCorrelation.computeTimeDomain / normalizedCrossCorrelation   algorithms/stats/correlation.go:203-228, 373-409
//   findPeak / findSecondPeak                                          :526-548, :617-638
//
// The reference evaluates sum_i a[i] b[i+lag] for every lag as a sequential float64 sum; the detected lag index
// must be bit-identical, which ncc_tiled_kernel guarantees by replaying that order -- at O(lags * n) FP64
// operations per pair.  The arg-max only needs the exact value of the few lags that can win:
//   1. c~(lag) for all lags from one forward and one inverse complex FFT of z = a + i b (numerators) and
//      prefix sums of squares (denominators; znorm_kernel leaves them behind, see XcorrSeq::prefix).  |c~ - c| is ~1e-14, far below kDelta;
//   2. every lag with |c~| >= (second largest |c~|) - kDelta is a candidate for the peak or the second peak;
//      lags whose denominator is too small for the error bound to hold are candidates too;
//   3. ncc_tiled_kernel runs only for the 64-lag blocks that hold a candidate and overwrites c~ there with the
//      exact values, so the arg-max, its value, the second peak and every tie between candidates are resolved
//      on bit-exact numbers (so are the peak's two neighbours, which enter the sharpness); the remaining
//      (approximate) entries only enter the noise / side-lobe reductions at relative error ~1e-13.
// A flat curve (everything within kDelta of the second peak, e.g. silence) degenerates to the full exact kernel.
// This file is not in the -fmad=false set: nothing here feeds a bit-exact output.
#include "common.h"
#include "xcorr_fft.cuh"

namespace sonar {
namespace {

constexpr double kDelta = 1e-8;     // candidate band below the second largest |c~|
constexpr double kSmallQ = 1e-3;    // overlap energy below this fraction of the total: lag is verified exactly
constexpr int kXsThreads = 256;

struct XsBuf {  // scratch of one chunk, carved from one allocation
  double2* x;
  double2* y;
  double* pre;          // per pair: PA[0..pre_stride), PB[0..pre_stride)
  unsigned char* need;  // per pair: 2 * bps flags
  int64_t N;
  int64_t pre_stride;
  int need_stride;
  int bps;
};

__global__ void __launch_bounds__(kXsThreads) xs_pack_kernel(const XcorrPair* __restrict__ pairs, XsBuf w) {
  const XcorrPair p = pairs[blockIdx.y];
  const int64_t n = (int64_t)blockIdx.x * kXsThreads + threadIdx.x;
  if (n >= w.N) return;
  w.x[(int64_t)blockIdx.y * w.N + n] = make_double2(n < p.na ? p.za[n] : 0.0, n < p.nb ? p.zb[n] : 0.0);
}

template <int RADIX>
__global__ void __launch_bounds__(kXsThreads) xs_pass_kernel(const double2* __restrict__ src, double2* __restrict__ dst,
                                                             int64_t N, int64_t n, int64_t s, int dir) {
  const int64_t t = (int64_t)blockIdx.x * kXsThreads + threadIdx.x;
  if (t >= N / RADIX) return;
  const double2* x = src + (int64_t)blockIdx.y * N;
  double2* y = dst + (int64_t)blockIdx.y * N;
  if (RADIX == 4)
    xs_radix4(x, y, t, n, s, dir);
  else
    xs_radix2(x, y, t, n, s, dir);
}

__global__ void __launch_bounds__(kXsThreads) xs_spectrum_kernel(const double2* __restrict__ src, double2* __restrict__ dst,
                                                                 int64_t N) {
  const int64_t k = (int64_t)blockIdx.x * kXsThreads + threadIdx.x;
  if (k >= N) return;
  const double2* Z = src + (int64_t)blockIdx.y * N;
  dst[(int64_t)blockIdx.y * N + k] = xs_cross_spectrum(Z[k], Z[(N - k) & (N - 1)]);
}

// which CTA of ncc_tiled_kernel owns global lag index j of this pair (same arithmetic as the kernel)
__device__ __forceinline__ int xs_block_of(const XcorrPair& p, int64_t j, int bps) {
  const int64_t lag = j - p.aml;
  if (lag >= 0) {
    const int64_t l_lo = p.idx_lo - p.aml > 0 ? p.idx_lo - p.aml : 0;
    return (int)((lag - l_lo) / kNccFlagLags);
  }
  const int64_t l_lo = p.aml - p.idx_hi + 1 > 1 ? p.aml - p.idx_hi + 1 : 1;
  return bps + (int)((-lag - l_lo) / kNccFlagLags);
}

// c~[j] for j in [idx_lo, idx_hi) from the inverse transform (unnormalised) and the prefix sums
__global__ void __launch_bounds__(kXsThreads) xs_curve_kernel(const XcorrPair* __restrict__ pairs, const double2* __restrict__ inv,
                                                              XsBuf w) {
  const XcorrPair p = pairs[blockIdx.y];
  const int64_t j = p.idx_lo + (int64_t)blockIdx.x * kXsThreads + threadIdx.x;
  if (j >= p.idx_hi) return;
  const double* __restrict__ PA = w.pre + (int64_t)(2 * blockIdx.y) * w.pre_stride;
  const double* __restrict__ PB = PA + w.pre_stride;
  const int64_t lag = j - p.aml;
  int64_t s1, s2, len;
  xs_overlap(lag, p.na, p.nb, &s1, &s2, &len);
  double c = 0.0;
  if (len > 0) {
    // znorm_kernel left prefix sums of the squared deviations and, behind them, the factor to z^2
    const double r1 = PA[s1 + len] - PA[s1], r2 = PB[s2 + len] - PB[s2];
    const double q1 = r1 * PA[p.na + 1], q2 = r2 * PB[p.nb + 1];
    const double den = sqrt(q1 * q2);
    const double num = inv[(int64_t)blockIdx.y * w.N + (lag & (w.N - 1))].x / (double)w.N;
    c = den < 1e-10 ? 0.0 : num / den;
    // exact evaluation wanted: overlap energy too small for the error bound, or a denominator near the reference's
    // `den < 1e-10 -> 0` rule (correlation.go:401-405), which the prefix-sum denominator could take differently from the
    // per-lag sequential sums (sequences that took the non-normalised branch of z-scoring: ADVICE r1)
    if (!(r1 > kSmallQ * PA[p.na]) || !(r2 > kSmallQ * PB[p.nb]) || !(den > 1e-9))
      w.need[(int64_t)blockIdx.y * w.need_stride + xs_block_of(p, j, w.bps)] = 1;
  }
  p.corr[j - p.idx_lo] = c;
}

__global__ void __launch_bounds__(kXsThreads) xs_select_kernel(const XcorrPair* __restrict__ pairs, XsBuf w) {
  __shared__ double s1[kXsThreads], s2[kXsThreads];
  const XcorrPair p = pairs[blockIdx.x];
  const double* __restrict__ c = p.corr;
  const int64_t cnt = p.idx_hi - p.idx_lo;
  const int t = threadIdx.x;
  double m1 = 0.0, m2 = 0.0;  // two largest |c~| (with multiplicity)
  for (int64_t i = t; i < cnt; i += kXsThreads) {
    const double v = fabs(c[i]);
    if (v > m1) {
      m2 = m1;
      m1 = v;
    } else if (v > m2) {
      m2 = v;
    }
  }
  s1[t] = m1;
  s2[t] = m2;
  __syncthreads();
  for (int o = kXsThreads / 2; o > 0; o >>= 1) {
    if (t < o) {
      const double a1 = s1[t], a2 = s2[t], b1 = s1[t + o], b2 = s2[t + o];
      s1[t] = fmax(a1, b1);
      s2[t] = fmax(fmin(a1, b1), fmax(a2, b2));
    }
    __syncthreads();
  }
  // candidate band: kDelta, widened for long sequences to the worst-case cancellation error of a prefix-sum difference
  // relative to the smallest overlap energy the screen accepts (n eps / kSmallQ; 9e-8 at 10 minutes)
  const double nmax = (double)(p.na > p.nb ? p.na : p.nb);
  const double delta = fmax(kDelta, 4.0 * nmax * 2.220446049250313e-16 / kSmallQ);
  const double thr = s2[0] - delta, thr_peak = s1[0] - delta;
  unsigned char* need = w.need + (int64_t)blockIdx.x * w.need_stride;
  for (int64_t i = t; i < cnt; i += kXsThreads) {
    const double v = fabs(c[i]);
    if (!(v < thr)) need[xs_block_of(p, p.idx_lo + i, w.bps)] = 1;  // NaN counts as a candidate
    if (!(v < thr_peak)) {  // a possible peak: its neighbours enter the sharpness (second difference) -> exact too
      if (i > 0) need[xs_block_of(p, p.idx_lo + i - 1, w.bps)] = 1;
      if (i + 1 < cnt) need[xs_block_of(p, p.idx_lo + i + 1, w.bps)] = 1;
    }
  }
}

inline size_t up256(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace

XcorrScreen xcorr_screen_geom(int64_t max_n, int aml_max, int64_t max_shard_lags, int n_pairs) {
  XcorrScreen g{};
  int64_t need = max_n + aml_max;
  if (need < 1024) need = 1024;
  g.log2n = 0;
  while (((int64_t)1 << g.log2n) < need) ++g.log2n;
  g.N = (int64_t)1 << g.log2n;
  g.bps = (int)((max_shard_lags + kNccFlagLags - 1) / kNccFlagLags);
  g.pre_stride = (max_n + 3) & ~(int64_t)1;  // P[0..n] and the scale
  g.need_stride = (2 * g.bps + 15) & ~15;
  g.n_pairs = n_pairs;
  size_t o = 0;
  g.o_x = o;
  o += up256(sizeof(double2) * (size_t)g.N * n_pairs);
  g.o_y = o;
  o += up256(sizeof(double2) * (size_t)g.N * n_pairs);
  g.o_pre = o;
  o += up256(sizeof(double) * (size_t)g.pre_stride * 2 * n_pairs);
  g.o_need = o;
  o += up256((size_t)g.need_stride * n_pairs);
  g.bytes = o;
  return g;
}

double* xcorr_screen_prefix(const XcorrScreen& g, void* scratch, int seq) {
  return reinterpret_cast<double*>(static_cast<unsigned char*>(scratch) + g.o_pre) + (int64_t)seq * g.pre_stride;
}

int launch_xcorr_screened(const XcorrPair* pairs_dev, int n_pairs, int64_t max_shard_lags, const XcorrScreen& g,
                          void* scratch, cudaStream_t st) {
  if (n_pairs <= 0 || max_shard_lags <= 0) return SONAR_OK;
  if (n_pairs > g.n_pairs) return set_error(SONAR_ERR_INVALID, "xcorr screen scratch too small");
  unsigned char* base = static_cast<unsigned char*>(scratch);
  XsBuf w;
  w.x = reinterpret_cast<double2*>(base + g.o_x);
  w.y = reinterpret_cast<double2*>(base + g.o_y);
  w.pre = reinterpret_cast<double*>(base + g.o_pre);
  w.need = base + g.o_need;
  w.N = g.N;
  w.pre_stride = g.pre_stride;
  w.need_stride = g.need_stride;
  w.bps = g.bps;
  const unsigned np = (unsigned)n_pairs;
  SONAR_CUDA(cudaMemsetAsync(w.need, 0, (size_t)g.need_stride * n_pairs, st));
  prof_begin("xs_pack_kernel", st);
  xs_pack_kernel<<<dim3((unsigned)(g.N / kXsThreads), np), kXsThreads, 0, st>>>(pairs_dev, w);
  prof_end();
  double2* src = w.x;
  double2* dst = w.y;
  auto transform = [&](int dir) {
    int64_t n = g.N, s = 1;
    while (n >= 4) {
      prof_begin("xs_pass_kernel", st);
      xs_pass_kernel<4><<<dim3((unsigned)(g.N / 4 / kXsThreads), np), kXsThreads, 0, st>>>(src, dst, g.N, n, s, dir);
      prof_end();
      std::swap(src, dst);
      n /= 4;
      s *= 4;
    }
    if (n == 2) {
      prof_begin("xs_pass_kernel", st);
      xs_pass_kernel<2><<<dim3((unsigned)(g.N / 2 / kXsThreads), np), kXsThreads, 0, st>>>(src, dst, g.N, n, s, dir);
      prof_end();
      std::swap(src, dst);
    }
  };
  transform(-1);
  prof_begin("xs_spectrum_kernel", st);
  xs_spectrum_kernel<<<dim3((unsigned)(g.N / kXsThreads), np), kXsThreads, 0, st>>>(src, dst, g.N);
  prof_end();
  std::swap(src, dst);
  transform(+1);
  prof_begin("xs_curve_kernel", st);
  xs_curve_kernel<<<dim3((unsigned)((max_shard_lags + kXsThreads - 1) / kXsThreads), np), kXsThreads, 0, st>>>(pairs_dev, src, w);
  prof_end();
  prof_begin("xs_select_kernel", st);
  xs_select_kernel<<<np, kXsThreads, 0, st>>>(pairs_dev, w);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return launch_xcorr_flagged(pairs_dev, n_pairs, max_shard_lags, w.need, g.need_stride, st);
}

}  // namespace sonar
