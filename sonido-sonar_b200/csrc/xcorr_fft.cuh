// Index arithmetic and butterflies of the FP64 FFT that screens the cross-correlation (xcorr_fft.cu).
// __host__ __device__ so tests/cpp/xcorr_fft_selftest.cu can run the very same passes on the CPU.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#define XS_HD __host__ __device__ __forceinline__

namespace sonar {

XS_HD double2 xs_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
XS_HD double2 xs_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
XS_HD double2 xs_mul(double2 a, double2 w) { return make_double2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }

// exp(dir * 2*pi*i * num / den), dir = -1 forward, +1 inverse
XS_HD double2 xs_twiddle(int64_t num, int64_t den, int dir) {
  double s, c;
#ifdef __CUDA_ARCH__
  sincospi(2.0 * (double)num / (double)den, &s, &c);
#else
  const double a = 6.283185307179586476925286766559 * (double)num / (double)den;
  s = std::sin(a);
  c = std::cos(a);
#endif
  return make_double2(c, dir < 0 ? -s : s);
}

// One butterfly of a Stockham autosort pass.  The transform of length n_total proceeds through passes
// (n, s): n = remaining sub-transform length, s = stride (n * s == n_total); a radix-4 pass maps (n, s) to
// (n / 4, 4 s), a radix-2 pass to (n / 2, 2 s).  Butterfly t in [0, n_total / radix): q = t % s, p = t / s.
XS_HD void xs_radix4(const double2* __restrict__ x, double2* __restrict__ y, int64_t t, int64_t n, int64_t s, int dir) {
  const int64_t q = t % s, p = t / s;
  const int64_t n1 = n / 4;
  const double2 a = x[q + s * p];
  const double2 b = x[q + s * (p + n1)];
  const double2 c = x[q + s * (p + 2 * n1)];
  const double2 d = x[q + s * (p + 3 * n1)];
  const double2 apc = xs_add(a, c), amc = xs_sub(a, c), bpd = xs_add(b, d), bmd = xs_sub(b, d);
  // forward: W_4 = -i, so the odd outputs use amc -/+ i*bmd; the inverse swaps the sign
  const double2 jb = dir < 0 ? make_double2(bmd.y, -bmd.x) : make_double2(-bmd.y, bmd.x);
  const double2 w1 = xs_twiddle(p, n, dir), w2 = xs_twiddle(2 * p, n, dir), w3 = xs_twiddle(3 * p, n, dir);
  y[q + s * (4 * p + 0)] = xs_add(apc, bpd);
  y[q + s * (4 * p + 1)] = xs_mul(xs_add(amc, jb), w1);
  y[q + s * (4 * p + 2)] = xs_mul(xs_sub(apc, bpd), w2);
  y[q + s * (4 * p + 3)] = xs_mul(xs_sub(amc, jb), w3);
}

XS_HD void xs_radix2(const double2* __restrict__ x, double2* __restrict__ y, int64_t t, int64_t n, int64_t s, int dir) {
  const int64_t q = t % s, p = t / s;
  const int64_t m = n / 2;
  const double2 a = x[q + s * p];
  const double2 b = x[q + s * (p + m)];
  y[q + s * (2 * p + 0)] = xs_add(a, b);
  y[q + s * (2 * p + 1)] = xs_mul(xs_sub(a, b), xs_twiddle(p, n, dir));
}

// Spectrum of the correlation from the transform Z of z = a + i b (both real):
//   A[k] = (Z[k] + conj Z[N-k]) / 2,  B[k] = (Z[k] - conj Z[N-k]) / (2 i),  C[k] = conj(A[k]) B[k]
// whose inverse transform is N * sum_i a[i] b[i + lag] at index lag mod N.
XS_HD double2 xs_cross_spectrum(double2 zk, double2 zm /* Z[(N - k) % N] */) {
  const double2 A = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
  const double2 D = make_double2(0.5 * (zk.x - zm.x), 0.5 * (zk.y + zm.y));  // (Z[k] - conj Z[N-k]) / 2
  const double2 B = make_double2(D.y, -D.x);                                 // D / i
  return make_double2(A.x * B.x + A.y * B.y, A.x * B.y - A.y * B.x);          // conj(A) * B
}

// calculateOverlapRegion (algorithms/stats/correlation.go:421-449) for lag = j - aml
XS_HD void xs_overlap(int64_t lag, int64_t na, int64_t nb, int64_t* s1, int64_t* s2, int64_t* len) {
  if (lag >= 0) {
    *s1 = 0;
    *s2 = lag;
    *len = na < nb - lag ? na : nb - lag;
  } else {
    *s1 = -lag;
    *s2 = 0;
    const int64_t e2 = nb < na + lag ? nb : na + lag;
    *len = (na + lag) < e2 ? (na + lag) : e2;
  }
}

}  // namespace sonar
