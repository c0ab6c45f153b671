// In-register radix-R FFT building blocks in float64 (R <= 16), used by the YIN autocorrelation.
// Same structure as fft_regs.cuh (compile-time recursion, natural order in and out), __host__ __device__ so
// the butterflies can be checked on the CPU (tests/cpp/fft_selftest.cu).
#pragma once
#include <cuda_runtime.h>

#define SONAR_HD64 __host__ __device__ __forceinline__

namespace sonar {

SONAR_HD64 double2 dadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
SONAR_HD64 double2 dsub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// explicit fma(): the including translation units are compiled with -fmad=false
SONAR_HD64 double2 dmul(double2 a, double2 w) {
  return make_double2(fma(a.x, w.x, -(a.y * w.y)), fma(a.x, w.y, a.y * w.x));
}
SONAR_HD64 double2 dmul_conj(double2 a, double2 w) {  // a * conj(w)
  return make_double2(fma(a.x, w.x, a.y * w.y), fma(a.y, w.x, -(a.x * w.y)));
}
SONAR_HD64 double2 dsqr(double2 w) { return make_double2(fma(w.x, w.x, -(w.y * w.y)), 2.0 * (w.x * w.y)); }
SONAR_HD64 double2 dswap(double2 a) { return make_double2(a.y, a.x); }

// a * W_R^K = a * exp(-2 pi i K / R), K in [0, R/2), R in {2, 4, 8, 16}
template <int R, int K>
SONAR_HD64 double2 dmul_tw(double2 a) {
  constexpr int K16 = K * (16 / R);  // index into the 16th roots of unity
  if constexpr (K16 == 0) {
    return a;
  } else if constexpr (K16 == 4) {  // -i
    return make_double2(a.y, -a.x);
  } else if constexpr (K16 == 2) {  // (1 - i)/sqrt2
    const double c = 0.70710678118654752440;
    return make_double2(c * (a.x + a.y), c * (a.y - a.x));
  } else if constexpr (K16 == 6) {  // (-1 - i)/sqrt2
    const double c = 0.70710678118654752440;
    return make_double2(c * (a.y - a.x), -c * (a.x + a.y));
  } else {
    const double c8 = 0.92387953251128675613, s8 = 0.38268343236508977173;  // cos, sin of pi/8
    if constexpr (K16 == 1) return dmul(a, make_double2(c8, -s8));
    if constexpr (K16 == 3) return dmul(a, make_double2(s8, -c8));
    if constexpr (K16 == 5) return dmul(a, make_double2(-s8, -c8));
    if constexpr (K16 == 7) return dmul(a, make_double2(-c8, -s8));
    return a;
  }
}

template <int R, int K>
SONAR_HD64 void ddit_combine(double2 (&v)[R], const double2 (&e)[R / 2], const double2 (&o)[R / 2]) {
  if constexpr (K < R / 2) {
    const double2 t = dmul_tw<R, K>(o[K]);
    v[K] = dadd(e[K], t);
    v[K + R / 2] = dsub(e[K], t);
    ddit_combine<R, K + 1>(v, e, o);
  }
}
template <int R, int K>
SONAR_HD64 void ddit_split(const double2 (&v)[R], double2 (&e)[R / 2], double2 (&o)[R / 2]) {
  if constexpr (K < R / 2) {
    e[K] = v[2 * K];
    o[K] = v[2 * K + 1];
    ddit_split<R, K + 1>(v, e, o);
  }
}

// Forward DFT of R points held in registers, natural order in and out.
template <int R>
struct FftReg64 {
  SONAR_HD64 static void run(double2 (&v)[R]) {
    double2 e[R / 2], o[R / 2];
    ddit_split<R, 0>(v, e, o);
    FftReg64<R / 2>::run(e);
    FftReg64<R / 2>::run(o);
    ddit_combine<R, 0>(v, e, o);
  }
};
template <>
struct FftReg64<2> {
  SONAR_HD64 static void run(double2 (&v)[2]) {
    const double2 a = v[0], b = v[1];
    v[0] = dadd(a, b);
    v[1] = dsub(a, b);
  }
};
template <>
struct FftReg64<4> {
  SONAR_HD64 static void run(double2 (&v)[4]) {
    const double2 a = dadd(v[0], v[2]), b = dsub(v[0], v[2]);
    const double2 c = dadd(v[1], v[3]), d = dsub(v[1], v[3]);
    const double2 dj = make_double2(d.y, -d.x);  // d * (-i)
    v[0] = dadd(a, c);
    v[2] = dsub(a, c);
    v[1] = dadd(b, dj);
    v[3] = dsub(b, dj);
  }
};

}  // namespace sonar
