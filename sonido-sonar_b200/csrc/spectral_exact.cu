// Exact float64 re-evaluation of single STFT frames in the reference's own order of operations.
//
// The fused kernels (stft_v3.cu) compute in FP32.  That is within the stated 1e-4 for every continuous feature of a
// well-conditioned frame, but three kinds of frame need more:
//   * the rolloff is a DISCRETE choice (first bin whose cumulative energy reaches 85 %, spectral_rolloff.go:19-55): when
//     the cumulative sum passes the threshold within the FP32 error, the bin can flip;
//   * flatness, slope (means of ln|X_k| over ALL bins, spectral_flatness.go:31-70, spectral_slope.go:24-64) and the log
//     mel energies (mfcc.go:136-145) of a band are dominated by weak bins; bins at the FP32 transform's noise floor
//     (~1e-7 of the frame's RMS level) carry percent-level errors;
//   * bins <= 1e-10 (digital silence) leave the log sums by a threshold test on the float64 magnitude.
// The fused kernel lists such frames (per stream: [count, t0, t1, ...]); this kernel redoes them from the float64 PCM:
//   window -> the radix-2 decimation-in-time FFT of go-dsp (github.com/mjibson/go-dsp/fft, radix2.go: bit reversal, then
//   log2 N stages of t = r[i2] * factor[blocks * j]; r[i] +- t, the same butterfly graph with the same factor table, so
//   every bin is bit-identical to the sequential evaluation) -> math.Hypot -> each feature's sums left to right.
// Bit-identical inputs to the reference-order sums make rolloff / centroid / bandwidth / crest / the band ratios
// bit-identical to the oracle; flatness, slope and the MFCCs agree to the last bits of log().
// This file is compiled with -fmad=false (Go on amd64 never contracts a * b + c).
#include <math_constants.h>

#include <cmath>

#include "common.h"

namespace sonar {
namespace {

// One CTA works on batches of up to 32 listed frames.  Phase A: each of the four warps runs whole 1024-point float64
// transforms (one frame at a time, its own 16 KB buffer) and leaves the frame's magnitudes in a [32][B] table.
// Phase B: the reference's sums are sequential in the bin index but independent between frames, so a lane takes a
// FRAME and walks its bins in order with every chain of its group in flight (warp 0: centroid, bandwidth, rolloff,
// crest, band ratios; warp 1: flatness and slope, the group that needs ln|X|; warps 2 and 3: the mel filters and the
// DCT).  (The first version gave each chain one thread of one frame: 1 / 32 of the FP64 pipe, 56 us per frame.)
constexpr int kXW = 4;    // warps per CTA
constexpr int kXB = 32;   // frames per batch
constexpr int kXT = kXW * 32;
constexpr unsigned kFullX = 0xffffffffu;

struct XSmem {
  size_t mags, fft, tw, fb, x10, lmel, total;
};
__host__ __device__ inline XSmem x_layout(int N, int n_mel) {
  XSmem L;
  const size_t B = (size_t)N / 2 + 1;  // odd: a lane-per-frame walk of the rows is bank-conflict free
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += (bytes + 15) & ~(size_t)15;
    return r;
  };
  L.mags = take(sizeof(double) * kXB * B);
  L.fft = take(sizeof(double2) * (size_t)kXW * N);
  L.tw = take(sizeof(double2) * (size_t)(N / 2));  // the factors every stage reads (global loads stalled the butterflies)
  L.fb = take(sizeof(double) * B);
  L.x10 = take(sizeof(double) * B);
  L.lmel = take(sizeof(double) * kXB * (size_t)((n_mel > 0 ? n_mel : 1) | 1));
  L.total = o;
  return L;
}

// Go math.Hypot (cmplx.Abs of the bin, analyzers/spectral.go:490-494)
__device__ __forceinline__ double go_hypot_dev(double p, double q) {
  p = fabs(p);
  q = fabs(q);
  if (isinf(p) || isinf(q)) return CUDART_INF;
  if (isnan(p) || isnan(q)) return CUDART_NAN;
  if (p < q) {
    const double t = p;
    p = q;
    q = t;
  }
  if (p == 0) return 0;
  q = q / p;
  return p * sqrt(1 + q * q);
}

template <int LOGN>
__global__ void __launch_bounds__(kXT, 1) spectral_exact_kernel(const StftArgs a) {
  extern __shared__ __align__(16) unsigned char xsm[];
  constexpr int N = 1 << LOGN, B = N / 2 + 1, logn = LOGN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const XSmem L = x_layout(N, a.n_mel);
  double* mags = reinterpret_cast<double*>(xsm + L.mags);                 // [kXB][B]
  double2* buf = reinterpret_cast<double2*>(xsm + L.fft) + (size_t)warp * N;  // this warp's transform
  double2* tw = reinterpret_cast<double2*>(xsm + L.tw);                   // go-dsp factors [0, N / 2)
  double* fb = reinterpret_cast<double*>(xsm + L.fb);                     // bin frequencies
  double* x10 = reinterpret_cast<double*>(xsm + L.x10);                   // log10 of them
  double* lmel = reinterpret_cast<double*>(xsm + L.lmel);                 // [kXB][n_mel | 1] log mel energies
  const int lms = a.n_mel | 1;

  for (int k = tid; k < N / 2; k += kXT) tw[k] = a.fac64[k];
  for (int k = tid; k < B; k += kXT) {  // spectral_centroid.go:59-65
    const double f = (double)k * (double)a.algo_sr / (double)((B - 1) * 2);
    fb[k] = f;
    x10[k] = f > 0 ? log10(f) : 0.0;
  }

  for (int s = blockIdx.y; s < a.n_streams; s += gridDim.y) {
    const int* lst = a.xlist + (int64_t)s * a.xlist_stride;
    const int count = lst[0];
    const double* __restrict__ x = a.pcm + (int64_t)s * a.stride;
    double* __restrict__ fo = a.feat + (int64_t)s * a.feat_stride;
    for (int b0 = blockIdx.x * kXB; b0 < count; b0 += gridDim.x * kXB) {
      const int nb = count - b0 < kXB ? count - b0 : kXB;
      __syncthreads();  // the previous batch's readers are done (and the tables are written)
      // ---- phase A: transforms, one frame per warp at a time ----------------------------------------------------
      for (int f = warp; f < nb; f += kXW) {
        const int64_t t = lst[1 + b0 + f];
        const double* __restrict__ fr = x + t * a.hop;
        // analyzers/spectral.go:477-480, stored at the bit-reversed position; lane l takes the samples 32 l + u so that
        // the stores of one instruction fall on consecutive slots
#pragma unroll 16
        for (int u = 0; u < N / 32; ++u) {
          const int i = (N / 32) * lane + u;
          buf[__brev((unsigned)i) >> (32 - logn)] = make_double2(fr[i] * a.win64[i], 0.0);
        }
        __syncwarp();
#pragma unroll 1
        for (int sh = 0; sh < logn; ++sh) {  // stage = 2 << sh
          const int s2 = 1 << sh, tshift = logn - 1 - sh;  // factor index = (N / stage) * j
#pragma unroll 8
          for (int u = 0; u < N / 64; ++u) {
            const int b = lane + 32 * u;
            const int j = b & (s2 - 1), i1 = ((b >> sh) << (sh + 1)) | j, i2 = i1 + s2;
            const double2 r1 = buf[i1], r2 = buf[i2];
            double2 w = r2;
            if (sh != 0) {
              const double2 fc = tw[j << tshift];
              w = make_double2(r2.x * fc.x - r2.y * fc.y, r2.x * fc.y + r2.y * fc.x);
            }
            buf[i1] = make_double2(r1.x + w.x, r1.y + w.y);
            buf[i2] = make_double2(r1.x - w.x, r1.y - w.y);
          }
          __syncwarp();
        }
        double* mrow = mags + (size_t)f * B;
#pragma unroll 4
        for (int k = lane; k < B; k += 32) mrow[k] = go_hypot_dev(buf[k].x, buf[k].y);
        __syncwarp();
      }
      __syncthreads();
      // ---- phase B: lane = frame, every sum left to right -------------------------------------------------------
      const bool live = lane < nb;
      const int64_t t = live ? lst[1 + b0 + lane] : 0;
      const double* __restrict__ m = mags + (size_t)(live ? lane : 0) * B;
      if (warp == 0) {
        // centroid (spectral_centroid.go:18-40), total energy and maximum (spectral_rolloff.go:19-31,
        // spectral_crest.go:18-39), band ratios (extractors/speech.go:436-456)
        double num = 0.0, den = 0.0, total = 0.0, mx = 0.0, le = 0.0, he = 0.0;
        const int split = B / 4;
#pragma unroll 4
        for (int i = 0; i < B; i++) {
          const double v = m[i], en = v * v;
          num += fb[i] * v;
          den += v;
          total += en;
          mx = v > mx ? v : mx;
          if (i < split)
            le += en;
          else
            he += en;
        }
        const double c = den == 0 ? 0.0 : num / den;
        // bandwidth (spectral_bandwidth.go:22-46; its denominator is the centroid's sum), rolloff crossing
        const double target = 0.85 * total;
        double bn = 0.0, cum = 0.0;
        int first = B - 1;
        bool found = false;
#pragma unroll 4
        for (int i = 0; i < B; i++) {
          const double v = m[i], d = fb[i] - c;
          bn += d * d * v;
          cum += v * v;
          const bool hit = !found && cum >= target;
          first = hit ? i : first;
          found = found || hit;
        }
        if (live) {
          fo[a.o_centroid + t] = c;
          fo[a.o_bandwidth + t] = den == 0 ? 0.0 : sqrt(bn / den);
          fo[a.o_rolloff + t] = total == 0 ? 0.0 : fb[first];
          const double rms = sqrt(total / (double)B);
          fo[a.o_crest + t] = rms == 0 ? 0.0 : mx / rms;
          if (t < a.Te) {
            fo[a.o_low + t] = total > 0 ? le / total : 0.0;
            fo[a.o_high + t] = total > 0 ? he / total : 0.0;
          }
        }
      } else if (warp == 1) {
        // flatness (spectral_flatness.go:31-70) and slope (spectral_slope.go:24-64): one ln|X| serves both
        // (log10 y = ln y / ln 10 to the last bits)
        double log_sum = 0.0, am = 0.0, sx = 0, sy = 0, sxy = 0, sxx = 0;
        int valid = 0, n = 0;
        const double inv_ln10 = 0.43429448190325182765;
#pragma unroll 4  // four logarithms in flight; the additions stay in order
        for (int i = 0; i < B; i++) {
          const double v = m[i];
          const bool ok = v > 1e-10;
          const double lv = ok ? log(v) : 0.0;
          log_sum += lv;  // + 0.0 where the bin is skipped: the sum never holds -0.0, so this is the skipped sum
          valid += ok ? 1 : 0;
          am += v;
          const bool oks = ok && fb[i] > 0;
          const double xx = oks ? x10[i] : 0.0, yy = oks ? lv * inv_ln10 : 0.0;
          sx += xx;
          sy += yy;
          sxy += xx * yy;
          sxx += xx * xx;
          n += oks ? 1 : 0;
        }
        double fl = 0.0;
        if (valid > 0) {
          const double gm = exp(log_sum / (double)valid);
          am /= (double)B;
          if (am > 1e-10) {
            fl = gm / am;
            if (fl > 1.0) fl = 1.0;
          }
        }
        double sl = 0.0;
        if (B >= 2 && n >= 2) {
          const double den = (double)n * sxx - sx * sx;
          if (den != 0) sl = ((double)n * sxy - sx * sy) / den;
        }
        if (live) {
          fo[a.o_flatness + t] = fl;
          fo[a.o_slope + t] = sl;
        }
      } else if (a.mfcc_on) {  // mel energies (mel_scale.go:58-105), zero weights skipped: warps 2 and 3 share the filters
        const int half = (a.n_mel + 1) / 2, f0 = warp == 2 ? 0 : half, f1 = warp == 2 ? half : a.n_mel;
        for (int f = f0; f < f1; ++f) {
          const int64_t l = a.melbins[f], c = a.melbins[f + 1], r = a.melbins[f + 2];
          double sum = 0.0;
          for (int64_t k = l; k < c && k < B; k++)
            if (c != l && k >= 0) {
              const double v = m[k];
              sum += (v * v) * ((double)(k - l) / (double)(c - l));
            }
          for (int64_t k = c; k < r && k < B; k++)
            if (r != c && k >= 0) {
              const double v = m[k];
              sum += (v * v) * ((double)(r - k) / (double)(r - c));
            }
          lmel[lane * lms + f] = sum > 0 ? log(sum) : log(1e-10);
        }
      }
      __syncthreads();
      if (warp >= 2 && a.mfcc_on) {  // DCT-II + lifter (mfcc.go:215-245): the coefficients split between the two warps
        const int half = (a.n_mfcc + 1) / 2, c0 = warp == 2 ? 0 : half, c1 = warp == 2 ? half : a.n_mfcc;
        for (int c = c0; c < c1; ++c) {
          double sum = 0.0;
          for (int n = 0; n < a.n_mel; n++) sum += lmel[lane * lms + n] * a.dct64[(size_t)c * a.n_mel + n];
          if (c >= 1) sum = sum * a.lift64[c];
          if (live) fo[a.o_mfcc + t * a.n_mfcc + c] = sum;
        }
      }
    }
  }
}

}  // namespace

int launch_spectral_exact(const StftArgs& a, cudaStream_t st) {
  if (!a.xlist || a.n_streams <= 0) return SONAR_OK;
  if (a.n_mel > 64 || (a.N != 1024 && a.N != 512))
    return set_error(SONAR_ERR_UNSUPPORTED, "exact frame re-evaluation: N = 512 or 1024, <= 64 mel filters");
  const XSmem L = x_layout(a.N, a.n_mel);
  if (L.total > 227 * 1024) return set_error(SONAR_ERR_UNSUPPORTED, "exact frame re-evaluation: mel bank too large for one CTA");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // one resident wave (one CTA per SM); a stream's list is walked by up to T / 32 CTAs
  int gx = sms / a.n_streams;
  if (gx < 1) gx = 1;
  const int64_t maxb = (a.T + kXB - 1) / kXB;
  if ((int64_t)gx > maxb) gx = (int)maxb;
  int gy = a.n_streams < sms ? a.n_streams : sms;
  if (a.N == 1024) {
    SONAR_CUDA(cudaFuncSetAttribute(spectral_exact_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    prof_begin("spectral_exact_kernel", st);
    spectral_exact_kernel<10><<<dim3((unsigned)gx, (unsigned)gy), kXT, L.total, st>>>(a);
  } else {
    SONAR_CUDA(cudaFuncSetAttribute(spectral_exact_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    prof_begin("spectral_exact_kernel", st);
    spectral_exact_kernel<9><<<dim3((unsigned)gx, (unsigned)gy), kXT, L.total, st>>>(a);
  }
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

}  // namespace sonar
