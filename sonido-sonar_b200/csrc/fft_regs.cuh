// In-register radix-R FFT building blocks (FP32) for the framed STFT kernel.
//
// Everything here is __host__ __device__ so the index arithmetic and the
// butterflies can be unit-tested on the CPU (tests/cpp/fft_selftest.cu) before
// any GPU time is spent.  All loops are compile-time recursions so that the
// float2 arrays live in registers (no dynamic indexing, no local memory).
#pragma once
#include <cuda_runtime.h>

#define SONAR_HD __host__ __device__ __forceinline__

namespace sonar {

// W_64^k = exp(-2*pi*i*k/64), k = 0..63.  After inlining the switch folds to a
// literal because every call site passes a compile-time constant.
SONAR_HD float2 w64_half(int k) {  // k in [0, 32]
  switch (k) {
    case 0: return make_float2(1.000000000e+00f, 0.000000000e+00f);
    case 1: return make_float2(9.951847267e-01f, -9.801714033e-02f);
    case 2: return make_float2(9.807852804e-01f, -1.950903220e-01f);
    case 3: return make_float2(9.569403357e-01f, -2.902846773e-01f);
    case 4: return make_float2(9.238795325e-01f, -3.826834324e-01f);
    case 5: return make_float2(8.819212643e-01f, -4.713967368e-01f);
    case 6: return make_float2(8.314696123e-01f, -5.555702330e-01f);
    case 7: return make_float2(7.730104534e-01f, -6.343932842e-01f);
    case 8: return make_float2(7.071067812e-01f, -7.071067812e-01f);
    case 9: return make_float2(6.343932842e-01f, -7.730104534e-01f);
    case 10: return make_float2(5.555702330e-01f, -8.314696123e-01f);
    case 11: return make_float2(4.713967368e-01f, -8.819212643e-01f);
    case 12: return make_float2(3.826834324e-01f, -9.238795325e-01f);
    case 13: return make_float2(2.902846773e-01f, -9.569403357e-01f);
    case 14: return make_float2(1.950903220e-01f, -9.807852804e-01f);
    case 15: return make_float2(9.801714033e-02f, -9.951847267e-01f);
    case 16: return make_float2(0.000000000e+00f, -1.000000000e+00f);
    case 17: return make_float2(-9.801714033e-02f, -9.951847267e-01f);
    case 18: return make_float2(-1.950903220e-01f, -9.807852804e-01f);
    case 19: return make_float2(-2.902846773e-01f, -9.569403357e-01f);
    case 20: return make_float2(-3.826834324e-01f, -9.238795325e-01f);
    case 21: return make_float2(-4.713967368e-01f, -8.819212643e-01f);
    case 22: return make_float2(-5.555702330e-01f, -8.314696123e-01f);
    case 23: return make_float2(-6.343932842e-01f, -7.730104534e-01f);
    case 24: return make_float2(-7.071067812e-01f, -7.071067812e-01f);
    case 25: return make_float2(-7.730104534e-01f, -6.343932842e-01f);
    case 26: return make_float2(-8.314696123e-01f, -5.555702330e-01f);
    case 27: return make_float2(-8.819212643e-01f, -4.713967368e-01f);
    case 28: return make_float2(-9.238795325e-01f, -3.826834324e-01f);
    case 29: return make_float2(-9.569403357e-01f, -2.902846773e-01f);
    case 30: return make_float2(-9.807852804e-01f, -1.950903220e-01f);
    case 31: return make_float2(-9.951847267e-01f, -9.801714033e-02f);
    case 32: return make_float2(-1.000000000e+00f, 0.000000000e+00f);
    default: return make_float2(0.f, 0.f);
  }
}
SONAR_HD float2 w64(int k) {
  k &= 63;
  if (k <= 32) return w64_half(k);
  float2 w = w64_half(k - 32);  // W^k = -W^(k-32)
  return make_float2(-w.x, -w.y);
}

SONAR_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SONAR_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
SONAR_HD float2 cmul(float2 a, float2 w) {
  return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}

// a * W_R^K with K in [0, R/2); trivial twiddles cost no multiplies.
template <int R, int K>
SONAR_HD float2 mul_tw(float2 a) {
  if constexpr (K == 0) {
    return a;
  } else if constexpr (4 * K == R) {  // -i
    return make_float2(a.y, -a.x);
  } else if constexpr (8 * K == R) {  // (1 - i)/sqrt2
    const float c = 0.70710678118654752f;
    return make_float2(c * (a.x + a.y), c * (a.y - a.x));
  } else if constexpr (8 * K == 3 * R) {  // (-1 - i)/sqrt2
    const float c = 0.70710678118654752f;
    return make_float2(c * (a.y - a.x), -c * (a.x + a.y));
  } else {
    static_assert(64 % R == 0, "radix must divide 64");
    return cmul(a, w64(K * (64 / R)));
  }
}

template <int R, int K>
SONAR_HD void dit_combine(float2 (&v)[R], const float2 (&e)[R / 2], const float2 (&o)[R / 2]) {
  if constexpr (K < R / 2) {
    float2 t = mul_tw<R, K>(o[K]);
    v[K] = cadd(e[K], t);
    v[K + R / 2] = csub(e[K], t);
    dit_combine<R, K + 1>(v, e, o);
  }
}

template <int R, int K>
SONAR_HD void dit_split(const float2 (&v)[R], float2 (&e)[R / 2], float2 (&o)[R / 2]) {
  if constexpr (K < R / 2) {
    e[K] = v[2 * K];
    o[K] = v[2 * K + 1];
    dit_split<R, K + 1>(v, e, o);
  }
}

// Forward DFT of R points held in registers, natural order in and out.
template <int R>
struct FftReg {
  SONAR_HD static void run(float2 (&v)[R]) {
    float2 e[R / 2], o[R / 2];
    dit_split<R, 0>(v, e, o);
    FftReg<R / 2>::run(e);
    FftReg<R / 2>::run(o);
    dit_combine<R, 0>(v, e, o);
  }
};
template <>
struct FftReg<1> {
  SONAR_HD static void run(float2 (&)[1]) {}
};
template <>
struct FftReg<2> {
  SONAR_HD static void run(float2 (&v)[2]) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};
template <>
struct FftReg<4> {
  SONAR_HD static void run(float2 (&v)[4]) {
    float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
    float2 c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
    float2 dj = make_float2(d.y, -d.x);  // d * (-i)
    v[0] = cadd(a, c);
    v[2] = csub(a, c);
    v[1] = cadd(b, dj);
    v[3] = csub(b, dj);
  }
};

// Geometry of the two-pass decomposition M = R1*R2 of the N = 2M real FFT
// (one M-point complex FFT of z[n] = x[2n] + i x[2n+1], then a split pass).
//   pass 1: R2 columns (n2), radix-R1 over n1, element n = R2*n1 + n2, twiddle W_M^(n2*k1)
//   pass 2: R1 columns (k1), radix-R2 over n2, output bin k = k1 + R1*k2
template <int R1_, int R2_>
struct FftGeom {
  static constexpr int R1 = R1_, R2 = R2_;
  static constexpr int M = R1 * R2, N = 2 * M, B = M + 1;
  static constexpr int F = 32 / R1;                 // frames per warp iteration (pass-2 slots)
  static constexpr int P1_FPR = 32 / R2;            // frames handled per pass-1 round
  static constexpr int P1_ROUNDS = F / P1_FPR;
  static constexpr int XROW = R2 + 1;               // exchange row stride, float2 units
  static constexpr int XSLOT = R1 * XROW;
  static constexpr int MAGROW = (B + (B >> 5) + 2 + 31) & ~31;  // padded magnitude row, floats (multiple of 32 banks)
  static_assert(R1 <= 32 && R2 <= 32 && F >= 1 && P1_ROUNDS >= 1, "unsupported geometry");
  SONAR_HD static int mag_index(int k) { return k + (k >> 5); }
};

}  // namespace sonar
