// YIN pitch track of the speech extractor's harmonic block (float64).
//
//   extractHarmonicFeatures      fingerprint/extractors/speech.go:464-509
//   PitchDetector.DetectPitch    algorithms/tonal/pitch_detection.go:225-279
//   preprocessFrame              :282-314 (second pre-emphasis 0.97 + un-normalised Hann)
//   detectPitchYin               :349-420 (difference function, CMNDF, first dip < 0.15)
//   parabolicInterpolation       :743-764
//   postProcessResult / updateTemporalTracking   :767-921,978-1007 (sequential over frames)
//
// Kernel 1 (four 1024/512 frames per CTA, 64 threads each) produces the raw (frequency, confidence)
// of every frame; kernel 2 (one thread per stream) replays the reference's
// sequential 20-deep history logic (octave correction vs the median of the last
// five, confidence gate, median-of-three smoothing) and writes the six
// HarmonicFeatures arrays.
#include <cmath>

#include <cstdlib>

#include "common.h"
#include "fft_regs_f64.cuh"

namespace sonar {
namespace {

constexpr int kYinFrame = 1024, kYinHop = 512, kYinHalf = 512;
constexpr int kYinFpb = 4;                     // frames per CTA
constexpr int kYinTpf = 64;                    // threads per frame: 8 lags each
constexpr int kYinThreads = kYinFpb * kYinTpf;
constexpr int kYinPad = kYinFrame + kYinFrame / 8 + 8;  // p[e] stored at e + (e >> 3): stride-8 reads hit distinct banks

__device__ __forceinline__ int pidx(int e) { return e + (e >> 3); }

// CMNDF + first dip below 0.15 + parabolic refinement (pitch_detection.go:363-420,743-764) of one frame by one
// warp: d = the 512 difference-function values in shared memory, 16 lags per lane.
template <int LS>  // LS = distance between the 16-lag runs of consecutive lanes (16, or 17 for padded rows)
__device__ __forceinline__ void yin_pick(const double* __restrict__ d, int64_t f, int64_t Tp, int sr, int lane,
                                         double* __restrict__ r) {
  double dv[16], loc[16], run = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int tau = lane * 16 + k;
    dv[k] = d[lane * LS + k];
    run += tau >= 1 ? dv[k] : 0.0;
    loc[k] = run;
  }
  double incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  const double base = incl - run;
  // Screen before dividing: cmndf[tau] = d[tau] / (cum[tau] / tau) can only be below the 0.15 threshold when
  // d[tau] * tau < 0.16 * cum[tau] (a 6 % margin against the few-ulp rounding of the two divisions).  A frame
  // with no such lag has no pitch (pitch_detection.go:377-383 finds no dip) and skips the 1024 FP64 divisions.
  {
    bool cand = false;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int tau = lane * 16 + k;
      cand = cand || (tau >= 1 && dv[k] * (double)tau < 0.16 * (base + loc[k]));
    }
    if (!__any_sync(0xffffffffu, cand)) {
      if (lane == 0 && f < Tp) {
        r[f] = 0.0;
        r[Tp + f] = 0.0;
      }
      return;
    }
  }
  double cm[17];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int tau = lane * 16 + k;
    cm[k] = tau == 0 ? 1.0 : dv[k] / ((base + loc[k]) / (double)tau);
  }
  cm[16] = __shfl_down_sync(0xffffffffu, cm[0], 1);
  int first = kYinHalf;
#pragma unroll
  for (int k = 15; k >= 0; --k) {
    const int tau = lane * 16 + k;
    if (tau >= 1 && tau + 1 < kYinHalf && cm[k] < 0.15 && cm[k] < cm[k + 1]) first = tau;
  }
  const int mt = __reduce_min_sync(0xffffffffu, first);
  if (f >= Tp) return;
  // the owner lane of mt has cm[mt-1..mt+1] at hand except across lane borders: fetch from shared d-derived values
  const int owner = mt < kYinHalf ? mt / 16 : 0;
  double y1 = 0.0, y2 = 0.0, y3 = 0.0;
  if (mt < kYinHalf) {
    const int k = mt % 16;
    // cm[k-1] may live in the previous lane
    double prev_last = __shfl_up_sync(0xffffffffu, cm[15], 1);
    double c_m1 = 0.0, c_0 = 0.0, c_p1 = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q)
      if (q == k) {
        c_0 = cm[q];
        c_p1 = cm[q + 1];
        c_m1 = q > 0 ? cm[q - 1] : prev_last;
      }
    y1 = __shfl_sync(0xffffffffu, c_m1, owner);
    y2 = __shfl_sync(0xffffffffu, c_0, owner);
    y3 = __shfl_sync(0xffffffffu, c_p1, owner);
  }
  if (lane == 0) {
    double pitch = 0.0, conf = 0.0;
    if (mt < kYinHalf && mt > 0) {
      double period = (double)mt;
      if (!(mt <= 0 || mt >= kYinHalf - 1)) {  // parabolicInterpolation :743-764
        const double a = (y1 - 2 * y2 + y3) / 2, b = (y3 - y1) / 2;
        if (a != 0) period = (double)mt + (-b / (2 * a));
      }
      const double freq = (double)sr / period;
      if (freq >= 80.0 && freq <= 1000.0) {
        pitch = freq;
        conf = 1.0 - y2;
      }
    }
    r[f] = pitch;
    r[Tp + f] = conf;
  }
}

// Difference function through the autocorrelation identity
//   d[tau] = sum_j (p[j] - p[j+tau])^2 = E(0) + E(tau) - 2 r[tau],
//   r[tau] = sum_{j<512} p[j] p[j+tau],  E(tau) = sum_{j<512} p[j+tau]^2 = S[tau+512] - S[tau]
// (one DFMA per (j, tau) instead of subtract/multiply/add; values agree with the reference's direct
// form to ~1e-13 relative, far inside the feature tolerance).  Each thread owns 8 consecutive lags
// and slides an 8-value register window over p, so the inner loop is 2 shared loads per 8 DFMAs.
__global__ void __launch_bounds__(kYinThreads) yin_frame_kernel(const double* __restrict__ pcm, int64_t stride,
                                                                double alpha, int sr, int64_t Tp,
                                                                const double* __restrict__ hann,
                                                                double* __restrict__ raw, int64_t raw_stride) {
  extern __shared__ double yin_smem[];
  double (*sp)[kYinPad] = reinterpret_cast<double (*)[kYinPad]>(yin_smem);
  double (*sS)[kYinFrame + 2] = reinterpret_cast<double (*)[kYinFrame + 2]>(yin_smem + kYinFpb * kYinPad);  // prefix sums of p^2
  double (*sd)[kYinHalf] = reinterpret_cast<double (*)[kYinHalf]>(yin_smem + kYinFpb * (kYinPad + kYinFrame + 2));
  const int s = blockIdx.y;
  const int fl = threadIdx.x / kYinTpf, tf = threadIdx.x % kYinTpf;
  const int64_t f0 = (int64_t)blockIdx.x * kYinFpb;
  const double* __restrict__ x = pcm + (int64_t)s * stride;
  // ---- stage: stream-level pre-emphasis (speech.go:161), the detector's own (pitch_detection.go:299-314), Hann
  for (int e = threadIdx.x; e < kYinFpb * kYinFrame; e += kYinThreads) {
    const int ff = e / kYinFrame, i = e % kYinFrame;
    const int64_t f = f0 + ff;
    double v = 0.0;
    if (f < Tp) {
      const int64_t g = f * kYinHop + i;
      const double xm1 = g > 0 ? x[g - 1] : 0.0, xm2 = g > 1 ? x[g - 2] : 0.0;
      const double y = x[g] - alpha * xm1;
      v = y;
      if (i > 0) {
        const double ym1 = xm1 - alpha * xm2;
        v = y - 0.97 * ym1;
      }
      v *= hann[i];
    }
    sp[ff][pidx(i)] = v;
  }
  __syncthreads();
  // ---- prefix sums of squares: warp w of the CTA scans frame w (8 warps: two passes of 4 frames x ... )
  {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w < kYinFpb) {
      double* S = sS[w];
      const double* P = sp[w];
      double run = 0.0;  // lane handles 32 consecutive samples
      double loc[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const double v = P[pidx(lane * 32 + k)];
        run = fma(v, v, run);
        loc[k] = run;
      }
      double incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      const double base = incl - run;
      if (lane == 0) S[0] = 0.0;
#pragma unroll
      for (int k = 0; k < 32; ++k) S[lane * 32 + k + 1] = base + loc[k];
    }
  }
  // ---- autocorrelation r[tau0 .. tau0+8) with a sliding register window
  const double* __restrict__ P = sp[fl];
  const int tau0 = tf * 8;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double w[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) w[k] = P[pidx(tau0 + k)];
  for (int j = 0; j < kYinHalf; j += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const double a = P[pidx(j + u)];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fma(a, w[(u + k) & 7], acc[k]);
      w[u & 7] = P[pidx(j + u + tau0 + 8)];  // j+u+tau0+8 <= 511+504+8 = 1023
    }
  }
  __syncthreads();  // sS complete
  {
    const double* S = sS[fl];
    const double e0 = S[kYinHalf];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int tau = tau0 + k;
      const double et = S[tau + kYinHalf] - S[tau];
      double dv = (e0 + et) - 2.0 * acc[k];
      sd[fl][tau] = dv > 0.0 ? dv : 0.0;
    }
  }
  __syncthreads();
  // ---- CMNDF + first dip below 0.15: one warp per frame
  const int wv = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (wv >= kYinFpb) return;
  yin_pick<16>(sd[wv], f0 + wv, Tp, sr, lane, raw + (int64_t)s * raw_stride);
}

// ------------------------------------------------------------------------------------------------
// FFT form of the same difference function (default path).
//
// r[tau] = sum_{j<512} p[j] p[j+tau] is the linear cross-correlation of a = p[0..512) (zero padded) with
// b = p[0..1024), so r = IFFT(conj(A) B) with 1024-point transforms and no wrap-around (j + tau <= 1022).
// Per frame: ONE complex forward FFT of z = u + i v, the two 512-sample halves of the frame (their zero-padded
// spectra U, V fall out of the Hermitian split, A = U, B = U + (-1)^k V; half of z is zero, which prunes the first
// pass), and HALF an
// inverse FFT — the spectra conj(A)B of two frames are packed as Q = P0 + i P1, whose inverse carries r of
// frame 0 in its real part and r of frame 1 in its imaginary part.  ~66 k FP64 operations per frame
// instead of 262 k DFMA.
//
// 1024 = 16 x 4 x 16, 64 threads per transform, data in shared memory as a 16 x 64 tile of double2
// (row stride 68, one pad every 16 columns: every pass below is bank-conflict free):
//   pass 1  column n2 (stride-64 elements): 16-point DFT over n1, twiddle W_1024^(n2 k1)
//   pass 2a row k1, n2 = 16 m1 + m2:        4-point DFT over m1, twiddle W_64^(m2 j1)
//   pass 2b row k1, fixed j1:               16-point DFT over m2
// which leaves X[k1 + 16 j1 + 64 j2] at (row k1, column 16 j1 + j2).  The spectra are only multiplied
// point-wise, so they stay in that digit-reversed order and the inverse runs the three passes backwards
// (conjugate twiddle first, then the inverse butterflies), ending in natural order in registers.
constexpr int kFRow = 68;
constexpr int kFBuf = 16 * kFRow;  // double2 per transform
constexpr int kFftFpb = 4;         // frames per CTA iteration (two packed pairs)
constexpr int kFftThreads = 64 * kFftFpb;

__device__ __forceinline__ int fpos(int row, int c) { return row * kFRow + c + (c >> 4); }
__device__ __forceinline__ int kpos(int k) { return (k & 15) * kFRow + 17 * ((k >> 4) & 3) + (k >> 6); }
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int K>
__device__ __forceinline__ void ld16(double2 (&v)[16], const double2* __restrict__ p, int stride) {
  if constexpr (K < 16) {
    v[K] = p[K * stride];
    ld16<K + 1>(v, p, stride);
  }
}
template <int K>
__device__ __forceinline__ void st16(const double2 (&v)[16], double2* __restrict__ p, int stride) {
  if constexpr (K < 16) {
    p[K * stride] = v[K];
    st16<K + 1>(v, p, stride);
  }
}
template <int K>
__device__ __forceinline__ void swap16(double2 (&v)[16]) {
  if constexpr (K < 16) {
    v[K] = dswap(v[K]);
    swap16<K + 1>(v);
  }
}

// v[k] *= W_1024^(n2 k) (CONJ: the conjugate), k = 1..15.  Odd powers come from the table (conflict-free:
// odd strides), even powers by squaring.
template <bool CONJ>
__device__ __forceinline__ void twiddle16(double2 (&v)[16], const double2* __restrict__ W, int n2) {
  const double2 w1 = W[n2], w3 = W[3 * n2], w5 = W[5 * n2], w7 = W[7 * n2];
  const double2 w2 = dsqr(w1), w6 = dsqr(w3), w10 = dsqr(w5), w14 = dsqr(w7);
  const double2 w4 = dsqr(w2), w12 = dsqr(w6);
  const double2 w8 = dsqr(w4);
  const double2 w9 = W[9 * n2], w11 = W[11 * n2], w13 = W[13 * n2], w15 = W[15 * n2];
#define SONAR_TW(K, WK) v[K] = CONJ ? dmul_conj(v[K], WK) : dmul(v[K], WK)
  SONAR_TW(1, w1); SONAR_TW(2, w2); SONAR_TW(3, w3); SONAR_TW(4, w4); SONAR_TW(5, w5);
  SONAR_TW(6, w6); SONAR_TW(7, w7); SONAR_TW(8, w8); SONAR_TW(9, w9); SONAR_TW(10, w10);
  SONAR_TW(11, w11); SONAR_TW(12, w12); SONAR_TW(13, w13); SONAR_TW(14, w14); SONAR_TW(15, w15);
#undef SONAR_TW
}

template <int K>
__device__ __forceinline__ void pre_tw16(double2 (&v)[8]) {  // v[n] *= W_16^n
  if constexpr (K < 8) {
    v[K] = dmul_tw<16, K>(v[K]);
    pre_tw16<K + 1>(v);
  }
}

// forward transform of the tile Z by the 64 threads tf of barrier `bar`; result in digit-reversed order.
// HALF: only the first 512 points (rows 0..7 of the tile) are non-zero, so the 16-point transforms of pass 1
// reduce to two 8-point transforms, X[2j] = DFT8(x)[j] and X[2j+1] = DFT8(x W_16^n)[j], and rows 8..15 are not read.
template <bool HALF>
__device__ __forceinline__ void fft1024_fwd(double2* __restrict__ Z, const double2* __restrict__ W,
                                            const double2* __restrict__ W64, int tf, int bar) {
  {
    double2 v[16];
    double2* col = Z + tf + (tf >> 4);
    if (HALF) {
      double2 e[8], o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        e[k] = col[k * kFRow];
        o[k] = e[k];
      }
      pre_tw16<0>(o);
      FftReg64<8>::run(e);
      FftReg64<8>::run(o);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[2 * k] = e[k];
        v[2 * k + 1] = o[k];
      }
    } else {
      ld16<0>(v, col, kFRow);
      FftReg64<16>::run(v);
    }
    twiddle16<false>(v, W, tf);
    st16<0>(v, col, kFRow);
  }
  bar_sync(bar, 64);
  {
    const int m2 = tf & 15;
    const double2 w1 = W64[m2], w3 = W64[3 * m2];
    const double2 w2 = dsqr(w1);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double2* p = Z + ((tf >> 4) + 4 * q) * kFRow + m2;
      double2 u[4] = {p[0], p[17], p[34], p[51]};
      FftReg64<4>::run(u);
      p[0] = u[0];
      p[17] = dmul(u[1], w1);
      p[34] = dmul(u[2], w2);
      p[51] = dmul(u[3], w3);
    }
  }
  bar_sync(bar, 64);
  {
    double2 v[16];
    double2* p = Z + (tf >> 2) * kFRow + 17 * (tf & 3);
    ld16<0>(v, p, 1);
    FftReg64<16>::run(v);
    st16<0>(v, p, 1);
  }
}

// unscaled inverse of fft1024_fwd: digit-reversed spectrum in Q -> x[64 n1 + tf] in v[n1] (natural order)
__device__ __forceinline__ void fft1024_inv(double2* __restrict__ Q, const double2* __restrict__ W,
                                            const double2* __restrict__ W64, int tf, int bar, double2 (&v)[16]) {
  {
    double2* p = Q + (tf >> 2) * kFRow + 17 * (tf & 3);
    ld16<0>(v, p, 1);
    swap16<0>(v);  // IDFT(x) = swap(DFT(swap(x)))
    FftReg64<16>::run(v);
    swap16<0>(v);
    st16<0>(v, p, 1);
  }
  bar_sync(bar, 64);
  {
    const int m2 = tf & 15;
    const double2 w1 = W64[m2], w3 = W64[3 * m2];
    const double2 w2 = dsqr(w1);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double2* p = Q + ((tf >> 4) + 4 * q) * kFRow + m2;
      double2 u[4] = {dswap(p[0]), dswap(dmul_conj(p[17], w1)), dswap(dmul_conj(p[34], w2)),
                      dswap(dmul_conj(p[51], w3))};
      FftReg64<4>::run(u);
      p[0] = dswap(u[0]);
      p[17] = dswap(u[1]);
      p[34] = dswap(u[2]);
      p[51] = dswap(u[3]);
    }
  }
  bar_sync(bar, 64);
  ld16<0>(v, Q + tf + (tf >> 4), kFRow);
  twiddle16<true>(v, W, tf);
  swap16<0>(v);
  FftReg64<16>::run(v);
  swap16<0>(v);
}

constexpr int kERow = kYinHalf + kYinHalf / 16;  // E / d rows: one pad double per 16 (16-lag runs per lane, no conflicts)
__device__ __forceinline__ int epos(int tau) { return tau + (tau >> 4); }
constexpr int kRawPair = 2 * kYinHop + kYinHop + 2;  // raw samples a pair of frames needs: 2 history + 1536

__device__ __forceinline__ void cp_async8_zfill(double* smem_dst, const double* gmem_src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gmem_src), "r"(n));
}

// Raw PCM of the pair of frames starting at frame fp of stream x -> `dst` (the pair's idle second tile), padded
// by two doubles per 16 so that the 18-sample windows of the staging threads are conflict-free LDS.128 runs.
// Issued by the 64 threads of the pair's second frame while the first frame's threads run the inverse FFT.
__device__ __forceinline__ void prefetch_pair(double* __restrict__ dst, const double* __restrict__ x, int64_t fp,
                                              int64_t limit, int tf) {
  const int64_t g0 = fp * kYinHop - 2;
  for (int j = tf; j < kRawPair; j += 64) {
    const int64_t gi = g0 + j;
    const bool ok = gi >= 0 && gi < limit;
    cp_async8_zfill(dst + j + 2 * (j >> 4), x + (ok ? gi : 0), ok);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(kFftThreads, 2)
    yin_frame_fft_kernel(const double* __restrict__ pcm, int64_t stride, double alpha, int sr, int64_t Tp,
                         int groups_per_stream, int64_t total_groups, const double* __restrict__ hann,
                         double* __restrict__ raw, int64_t raw_stride) {
  extern __shared__ __align__(16) double yin_smem[];
  double2* sW = reinterpret_cast<double2*>(yin_smem);                         // W_1024^k
  double2* sW64 = sW + 1024;                                                  // W_64^k
  double2* sZ = sW64 + 64;                                                    // kFftFpb tiles
  double (*sE)[kERow] = reinterpret_cast<double (*)[kERow]>(sZ + kFftFpb * kFBuf);  // E(tau), then d[tau]
  const int tid = threadIdx.x;
  const int fl = tid >> 6, tf = tid & 63;
  const int pr = fl >> 1, hf = fl & 1, tp = tid & 127;
  const int64_t limit = (Tp + 1) * kYinHop;  // samples [0, limit) belong to the Tp frames
  for (int k = tid; k < 1024; k += kFftThreads) {
    double sn, cs;
    sincospi(-(double)k / 512.0, &sn, &cs);
    sW[k] = make_double2(cs, sn);
    if (k < 64) {
      sincospi(-(double)k / 32.0, &sn, &cs);
      sW64[k] = make_double2(cs, sn);
    }
  }
  double2* Z = sZ + fl * kFBuf;
  double* rawbuf = reinterpret_cast<double*>(sZ + (2 * pr + 1) * kFBuf);  // the pair's second tile
  if (hf && blockIdx.x < total_groups) {
    const int64_t grp = blockIdx.x;
    prefetch_pair(rawbuf, pcm + (grp / groups_per_stream) * stride, (grp % groups_per_stream) * kFftFpb + 2 * pr, limit, tf);
  }

  __syncthreads();  // twiddle tables complete
  for (int64_t grp = blockIdx.x; grp < total_groups; grp += gridDim.x) {
    const int s = (int)(grp / groups_per_stream);
    const int64_t f0 = (grp % groups_per_stream) * kFftFpb;
    // From here on the two pairs of a CTA never wait for each other: everything a pair touches (its two tiles, its
    // raw-sample buffer, its E / d rows) is private to its 128 threads, so every barrier below is pair-wide at most.
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    bar_sync(5 + pr, 128);  // raw samples landed; previous iteration's CMNDF is done with sE
    // ---- stage: thread tf of frame fl owns samples [16 tf, 16 tf + 16): stream-level pre-emphasis (speech.go:161),
    //      the detector's own (pitch_detection.go:299-314), un-normalised Hann; exclusive prefix sums of p^2
    double v[16], ex[16], run = 0.0;
    {
      double w[18];
      const double2* rp = reinterpret_cast<const double2*>(rawbuf + 18 * (hf * 32 + tf));
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const double2 t = rp[k < 8 ? k : 9];  // samples 16, 17 of the window sit behind the pad pair
        w[2 * k] = t.x;
        w[2 * k + 1] = t.y;
      }
      const bool live = f0 + fl < Tp;
      const double2* hp = reinterpret_cast<const double2*>(hann + 16 * tf);
#pragma unroll
      for (int k = 0; k < 16; k += 2) {
        const double2 h2 = __ldg(hp + (k >> 1));
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const double y = w[k + u + 2] - alpha * w[k + u + 1];
          double t = y;
          if (tf > 0 || k + u > 0) t = y - 0.97 * (w[k + u + 1] - alpha * w[k + u]);
          t *= u ? h2.y : h2.x;
          t = live ? t : 0.0;
          v[k + u] = t;
          ex[k + u] = run;
          run = fma(t, t, run);
        }
      }
    }
    bar_sync(5 + pr, 128);  // every window is in registers before the pair's tiles are overwritten
    {
      // z[n] = p[n] + i p[n + 512], n < 512 (rows 0..7 of the tile; rows 8..15 are zeros the pruned first pass
      // never reads): thread tf < 32 owns the real parts of n = 16 tf + k, thread tf + 32 their imaginary parts
      const int th = tf & 31;
      double* zp = reinterpret_cast<double*>(Z + (th >> 2) * kFRow + 17 * (th & 3)) + (tf >> 5);
#pragma unroll
      for (int k = 0; k < 16; ++k) zp[2 * k] = v[k];
    }
    // ---- E(tau) = S[tau + 512] - S[tau]
    {
      double incl = run;
      const int lane = tid & 31;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      double base = incl - run;  // sum over the lower lanes of this warp
      double* E = sE[fl];
      if (tf == 31) E[0] = incl;  // total of the lower warp (samples 0..511)
      bar_sync(1 + fl, 64);
      if (tf >= 32) base += E[0];
      bar_sync(1 + fl, 64);
      if (tf >= 32) {
#pragma unroll
        for (int k = 0; k < 16; ++k) E[17 * (tf - 32) + k] = base + ex[k];  // S[tau + 512]
      }
      bar_sync(1 + fl, 64);  // also orders the z stores before pass 1
      if (tf < 32) {
#pragma unroll
        for (int k = 0; k < 16; ++k) E[17 * tf + k] -= base + ex[k];  // - S[tau]
      }
    }
    fft1024_fwd<true>(Z, sW, sW64, tf, 1 + fl);
    // ---- P = conj(A) B per frame, Q = P0 + i P1 per pair, written over the pair's first tile
    bar_sync(5 + pr, 128);
    {
      double2* Z0 = sZ + (2 * pr) * kFBuf;
      double2* Z1 = Z0 + kFBuf;
      for (int it = 0; it < 5; ++it) {
        int k;
        if (it < 4) {
          k = 64 * (tp & 7) + (tp >> 3) + 16 * it;
        } else {
          if (tp != 0) break;
          k = 512;
        }
        const int pk = kpos(k), pn = kpos((1024 - k) & 1023);
        double2 P[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double2* Zh = h ? Z1 : Z0;
          const double2 zk = Zh[pk], zn = Zh[pn];
          // U = (zk + conj(zn))/2 and V = -i (zk - conj(zn))/2 are the spectra of the two zero-padded halves;
          // A = U, B = U + (-1)^k V, P = conj(A) B = |U|^2 + (-1)^k conj(U) V (the 1/4 is folded into the final scale)
          const double sr_ = zk.x + zn.x, si = zk.y - zn.y, dr = zk.x - zn.x, di = zk.y + zn.y;
          const double cr = fma(sr_, di, -(si * dr)), ci = -fma(sr_, dr, si * di);
          const double uu = fma(sr_, sr_, si * si);
          P[h] = (k & 1) ? make_double2(uu - cr, -ci) : make_double2(uu + cr, ci);
        }
        Z0[pk] = make_double2(P[0].x - P[1].y, P[0].y + P[1].x);
        if (pn != pk) Z0[pn] = make_double2(P[0].x + P[1].y, P[1].x - P[0].y);
      }
    }
    bar_sync(5 + pr, 128);
    if (hf) {
      // the second tile is idle from here on: fetch the next group's samples into it
      const int64_t nxt = grp + gridDim.x;
      if (nxt < total_groups)
        prefetch_pair(rawbuf, pcm + (nxt / groups_per_stream) * stride, (nxt % groups_per_stream) * kFftFpb + 2 * pr, limit,
                      tf);
    } else {
      double2 q[16];
      fft1024_inv(Z, sW, sW64, tf, 1 + fl, q);
      const double scale = 1.0 / 4096.0;  // 1/4 (split) * 1/1024 (inverse)
      double* E0 = sE[fl];
      double* E1 = sE[fl + 1];
      const double e00 = E0[0], e10 = E1[0];
      bar_sync(1 + fl, 64);  // everyone has read E[0] before tau = 0 is overwritten
#pragma unroll
      for (int n1 = 0; n1 < 8; ++n1) {
        const int ep = epos(64 * n1 + tf);
        const double d0 = (e00 + E0[ep]) - 2.0 * (q[n1].x * scale);
        const double d1 = (e10 + E1[ep]) - 2.0 * (q[n1].y * scale);
        E0[ep] = d0 > 0.0 ? d0 : 0.0;
        E1[ep] = d1 > 0.0 ? d1 : 0.0;
      }
    }
    bar_sync(5 + pr, 128);
    // ---- CMNDF + first dip below 0.15 (pitch_detection.go:363-383): the first warp of each frame, 16 lags per lane
    if (tf < 32) yin_pick<17>(sE[fl], f0 + fl, Tp, sr, tf, raw + (int64_t)s * raw_stride);
  }
}

// medianFilter over the non-zero entries (pitch_detection.go:978-1007), register-only.
// Zero entries are "absent"; every value is >= 0.
__device__ __forceinline__ double median_nonzero3(double a, double b, double c) {
  const int m = (a > 0) + (b > 0) + (c > 0);
  if (m == 0) return 0.0;
  const double lo = fmin(a, b), hi = fmax(a, b);
  if (m == 3) return fmax(lo, fmin(hi, c));  // middle of three
  const double mx = fmax(hi, c);
  if (m == 1) return mx;
  // two present: (smaller + larger) / 2; the absent one is 0 and adding it is exact
  const double mid = fmax(lo, fmin(hi, c));  // with one zero, the middle of {0, x, y} is min(x, y)
  return (mid + mx) / 2.0;
}

__device__ __forceinline__ void cswap(double& x, double& y) {
  const double lo = fmin(x, y), hi = fmax(x, y);
  x = lo;
  y = hi;
}

// last `n` (3..5) history entries h[5-n..4]; zeros are absent
__device__ __forceinline__ double median_nonzero5(const double (&h)[5], int n) {
  double v0 = n >= 5 ? h[0] : 0.0, v1 = n >= 4 ? h[1] : 0.0, v2 = h[2], v3 = h[3], v4 = h[4];
  const int m = (v0 > 0) + (v1 > 0) + (v2 > 0) + (v3 > 0) + (v4 > 0);
  if (m == 0) return 0.0;
  // 9-comparator sorting network, ascending: the m present values end up in the top m slots
  cswap(v0, v1); cswap(v3, v4); cswap(v2, v4); cswap(v2, v3); cswap(v0, v3);
  cswap(v0, v2); cswap(v1, v4); cswap(v1, v3); cswap(v1, v2);
  const double s[5] = {v0, v1, v2, v3, v4};
  const int first = 5 - m;
  double lo = 0.0, hi = 0.0;  // sorted[m/2 - 1] and sorted[m/2] of the present values
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (k == first + m / 2) hi = s[k];
    if (k == first + m / 2 - 1) lo = s[k];
  }
  return (m & 1) ? hi : (lo + hi) / 2.0;
}

// One warp per stream, 32 frames per round.  Only the octave correction is a recurrence (it looks at the
// last five corrected pitches), and it only fires for frames that carry a pitch with confidence >= 0.5;
// everything else — the confidence gate, the median-of-three smoothing, the derived arrays — is a pure
// function of a three-frame window and runs one frame per lane.  Lane 0 therefore walks just the gated
// frames of a round (ballot + find-first-set), in order, with the reference's arithmetic.
constexpr int kTrackBatch = 4;
__global__ void __launch_bounds__(32) yin_track_kernel(const double* __restrict__ raw, int64_t raw_stride,
                                                       int n_streams, int64_t Tp, double* __restrict__ feat,
                                                       int64_t feat_stride, int64_t o_pitch, int64_t o_conf,
                                                       int64_t o_voicing, int64_t o_hratio, int64_t o_inharm,
                                                       int64_t o_tonal, const double* __restrict__ speech_gate,
                                                       int64_t gate_stride) {
  constexpr int kB = kTrackBatch, kRound = 32 * kTrackBatch;  // frames per round: kB per lane
  __shared__ double c_sm[5 + kRound];  // corrected pitches: [0..4] = the five frames before this round
  __shared__ double raw_sm[kRound];
  __shared__ unsigned todo_sm[kB];
  const int s = blockIdx.x;
  const int lane = threadIdx.x;
  const double* r = raw + (int64_t)s * raw_stride;
  double* fo = feat + (int64_t)s * feat_stride;
  if (lane < 5) c_sm[lane] = 0.0;
  __syncwarp();
  // The speech-specific group (extractors/speech.go:194-205,529-549) sweeps the SAME detector over the same frames before
  // the harmonic block when the signal is judged to be speech: its history (20 entries, pitch_detection.go:876-902) is
  // what the harmonic block starts from.  Pass 0 of two replays that sweep without writing anything.
  const int passes = (speech_gate && speech_gate[(int64_t)s * gate_stride] != 0.0) ? 2 : 1;
  for (int pass = 0; pass < passes; ++pass) {
    const bool write = pass == passes - 1;
    const int64_t hist0 = pass == 0 ? 0 : (Tp < 20 ? Tp : 20);  // history entries before the pass's first frame
    // A round covers kRound frames (kB per lane) and its raw values are fetched one round AHEAD of the walk: the fixed
    // cost of a round (shared-memory hand-overs around the sequential octave correction, ~900 cycles) is paid once per
    // 128 frames instead of once per 32, and the two global loads no longer add their latency to it (0.61 -> 0.37 ms per
    // 25,838 frames with the prefetch alone; this kernel is the tail of the fingerprint step and was 2.7 of the 8.7 ms of
    // the one-hour streams of C3).
    double nraw[kB], nconf[kB];
    auto fetch = [&](int64_t b0) {
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const int64_t i = b0 + 32 * u + lane;
        nraw[u] = i < Tp ? r[i] : 0.0;
        nconf[u] = i < Tp ? r[Tp + i] : 0.0;
      }
    };
    fetch(0);
    for (int64_t base = 0; base < Tp; base += kRound) {
      const int cnt = (int)((Tp - base < kRound) ? (Tp - base) : kRound);
      double conf[kB];
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const int e = 32 * u + lane;
        const double rawp = nraw[u];  // zero beyond the last frame
        conf[u] = nconf[u];
        const bool gate = e < cnt && rawp != 0.0 && conf[u] >= 0.5;  // :782-786: below 0.5 everything is zeroed
        raw_sm[e] = rawp;
        c_sm[5 + e] = gate ? rawp : 0.0;
        const unsigned td = __ballot_sync(0xffffffffu, gate);
        if (lane == 0) todo_sm[u] = td;
      }
      fetch(base + kRound);
      __syncwarp();
      if (lane == 0) {
#pragma unroll 1  // ONE instance of the correction code: unrolled, the round was 33 KB of instructions fetched by a lone warp
        for (int u = 0; u < kB; ++u) {
          unsigned td = todo_sm[u];
          while (td) {
            const int k = 32 * u + __ffs(td) - 1;
            td &= td - 1;
            int64_t hlen = hist0 + base + k;  // history entries before this frame
            if (hlen > 20) hlen = 20;
            double pitch = raw_sm[k];
            if (hlen >= 3) {  // applyOctaveCorrection :792-829 (needs >= 3 of the last five)
              const double h[5] = {c_sm[k], c_sm[k + 1], c_sm[k + 2], c_sm[k + 3], c_sm[k + 4]};
              const double med = median_nonzero5(h, hlen < 5 ? (int)hlen : 5);
              const double ratios[4] = {0.5, 2.0, 1.0 / 3.0, 3.0};
#pragma unroll
              for (int q = 0; q < 4; q++) {
                const double expect = med * ratios[q];
                if (fabs(pitch - expect) / expect < 0.1) {
                  if (fabs(pitch - med) > fabs(expect - med)) pitch = expect;
                  break;
                }
              }
            }
            c_sm[5 + k] = pitch;
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const int e = 32 * u + lane;
        if (e < cnt) {
          const int64_t i = base + e;
          const double c0 = c_sm[5 + e], c1 = c_sm[4 + e], c2 = c_sm[3 + e];
          double pitch = c0;  // applyTemporalSmoothing :905-921 on the history that already includes this frame
          const int64_t hsize = hist0 + i + 1;  // history length including this frame (capped at 20: >= 3 either way)
          if (hsize >= 3)
            pitch = median_nonzero3(c2, c1, c0);
          else if (hsize == 2)
            pitch = 0.3 * c0 + (1 - 0.3) * c1;  // history of two: blend with the previous (unsmoothed) output
          const double cf = conf[u] < 0.5 ? 0.0 : conf[u];
          if (write) {
            fo[o_pitch + i] = pitch;
            fo[o_conf + i] = cf;
            fo[o_voicing + i] = cf;
            fo[o_hratio + i] = cf * 10.0;               // speech.go:499
            fo[o_inharm + i] = 1.0 - cf;                // speech.go:500
            fo[o_tonal + i] = pitch > 0 ? pitch : 0.0;  // speech.go:503-505
          }
        }
      }
      __syncwarp();
      const double carry = (lane < 5) ? c_sm[cnt + lane] : 0.0;  // the five entries that end with this round's last frame
      __syncwarp();
      if (lane < 5) c_sm[lane] = carry;
      __syncwarp();
    }
  }
}

__global__ void yin_zero_kernel(int n_streams, int64_t Tp, double* __restrict__ feat, int64_t feat_stride,
                                int64_t o_pitch, int64_t o_conf, int64_t o_voicing, int64_t o_hratio,
                                int64_t o_inharm, int64_t o_tonal) {
  const int s = blockIdx.y;
  double* fo = feat + (int64_t)s * feat_stride;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Tp; i += (int64_t)gridDim.x * blockDim.x) {
    fo[o_pitch + i] = 0.0;
    fo[o_conf + i] = 0.0;
    fo[o_voicing + i] = 0.0;
    fo[o_hratio + i] = 0.0;
    fo[o_inharm + i] = 1.0;
    fo[o_tonal + i] = 0.0;
  }
}

}  // namespace

// scratch: per stream 2*Tp doubles (raw pitch, raw confidence), scratch_stride apart; hann: 1024 doubles (device)
int launch_yin(const double* pcm, int64_t stride, int n_streams, double alpha, int sr, int64_t Tp,
               const double* hann_dev, double* feat, int64_t feat_stride, int64_t o_pitch, int64_t o_conf,
               int64_t o_voicing, int64_t o_hratio, int64_t o_inharm, int64_t o_tonal, double* scratch,
               int64_t scratch_stride, cudaStream_t st, cudaStream_t track_st, cudaEvent_t fork, cudaEvent_t join,
               bool* forked, int* lists, int64_t list_stride, const double* speech_gate, int64_t gate_stride) {
  if (forked) *forked = false;
  if (Tp <= 0 || n_streams <= 0) return SONAR_OK;
  if (sr <= 0) {
    // frequency = sampleRate/period = 0 fails the [80,1000] Hz gate for every frame
    // (pitch_detection.go:394-400): the outputs are constants, no kernel work needed.
    dim3 grid((unsigned)std::min<int64_t>((Tp + 255) / 256, 64), (unsigned)n_streams);
    prof_begin("yin_zero_kernel", st);
    yin_zero_kernel<<<grid, 256, 0, st>>>(n_streams, Tp, feat, feat_stride, o_pitch, o_conf, o_voicing,
                                          o_hratio, o_inharm, o_tonal);
    prof_end();
    SONAR_CUDA(cudaGetLastError());
    return SONAR_OK;
  }
  if (Tp > 0x7fffffffLL) return set_error(SONAR_ERR_UNSUPPORTED, "too many pitch frames");
  static const bool direct = std::getenv("SONAR_YIN_DIRECT") != nullptr;  // diagnostic: the O(W^2) form
  static const bool fp64 = std::getenv("SONAR_YIN_FP64") != nullptr;      // diagnostic: the float64 FFT kernel below
  if (!direct && !fp64 && lists) {  // default: packed-FP32 transforms + exact re-evaluation of borderline frames (yin32.cu)
    SONAR_CUDA(cudaMemset2DAsync(lists, sizeof(int) * (size_t)list_stride, 0, sizeof(int), (size_t)n_streams, st));
    int rc = launch_yin32(pcm, stride, n_streams, alpha, sr, Tp, hann_dev, scratch, scratch_stride, lists, list_stride, st);
    if (rc) return rc;
  } else if (direct) {
    dim3 grid((unsigned)((Tp + kYinFpb - 1) / kYinFpb), (unsigned)n_streams);
    const size_t smem = sizeof(double) * kYinFpb * (kYinPad + kYinFrame + 2 + kYinHalf);
    SONAR_CUDA(cudaFuncSetAttribute(yin_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_begin("yin_frame_kernel", st);
    yin_frame_kernel<<<grid, kYinThreads, smem, st>>>(pcm, stride, alpha, sr, Tp, hann_dev, scratch, scratch_stride);
    prof_end();
  } else {
    const int gps = (int)((Tp + kFftFpb - 1) / kFftFpb);
    const int64_t total = (int64_t)gps * n_streams;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = sizeof(double2) * (1024 + 64 + kFftFpb * kFBuf) + sizeof(double) * kFftFpb * kERow;
    SONAR_CUDA(cudaFuncSetAttribute(yin_frame_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // Not persistent: a CTA takes a bounded share of the groups (its twiddle tables cost ~1 % of that) and retires,
    // so the latency-bound kernels of the alignment branch, which run beside this kernel on a higher-priority
    // stream, get shared memory on an SM within ~100 us instead of after the whole launch.
    const int64_t per_cta = 12;
    const unsigned ctas = (unsigned)std::max<int64_t>(std::min<int64_t>(total, (int64_t)2 * sms), (total + per_cta - 1) / per_cta);
    prof_begin("yin_frame_kernel", st);
    yin_frame_fft_kernel<<<ctas, kFftThreads, smem, st>>>(pcm, stride, alpha, sr, Tp, gps, total, hann_dev, scratch,
                                                         scratch_stride);
    prof_end();
  }
  SONAR_CUDA(cudaGetLastError());
  // The tracker is a sequential walk, one warp per stream: with a side stream it runs beside whatever the caller
  // enqueues next on `st` (the caller makes `st` wait for `join` before it is done with the features).
  cudaStream_t ts = st;
  if (track_st && fork && join) {
    SONAR_CUDA(cudaEventRecord(fork, st));
    SONAR_CUDA(cudaStreamWaitEvent(track_st, fork, 0));
    ts = track_st;
  }
  prof_begin("yin_track_kernel", ts);
  yin_track_kernel<<<n_streams, 32, 0, ts>>>(scratch, scratch_stride, n_streams, Tp, feat, feat_stride, o_pitch, o_conf,
                                             o_voicing, o_hratio, o_inharm, o_tonal, speech_gate, gate_stride);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  if (ts != st) {
    SONAR_CUDA(cudaEventRecord(join, ts));
    if (forked) *forked = true;
  }
  return SONAR_OK;
}

}  // namespace sonar
