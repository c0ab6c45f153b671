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

// One CTA works on batches of up to 32 listed frames.  Phase A: four pairs of warps run whole 1024-point float64
// transforms (one frame per pair at a time, 16 KB buffer each) and leave the frame's magnitudes in a [32][B] table.
// Phase B: the reference's sums are sequential in the bin index but independent between frames, so a lane takes a
// FRAME and walks its bins in order with every chain of its group in flight (warp 0: centroid, bandwidth, rolloff,
// crest, band ratios; warp 1: flatness; warp 2: slope; warps 4..7: the mel filters and the DCT).
// (The first version gave each chain one thread of one frame: 1 / 32 of the FP64 pipe, 56 us per frame.)
constexpr int kXF = 4;    // transforms in flight per CTA (16 KB each)
constexpr int kXW = 8;    // warps per CTA: a PAIR of warps per transform (half the dependent latency per frame; with
                          // one warp per scheduler nothing hid the shared-memory and FP64 latencies: 0.13 issue / cycle)
constexpr int kXB = 32;   // frames per batch
constexpr int kXT = kXW * 32;
constexpr unsigned kFullX = 0xffffffffu;

struct XSmem {
  size_t mags, fft, tw, fb, x10, lmel, total;
  int batch, rows, nfft;  // frames per batch (<= 32: one per lane), magnitude rows (+ one halo row for the flux),
                          // transforms in flight
};
// N: window length, MF: radix-2 transform length (N, or NextPowerOf2(2 N - 1) on the Bluestein route)
__host__ __device__ inline XSmem x_layout(int N, int MF, int n_mel, bool all_frames) {
  XSmem L;
  const size_t B = (size_t)N / 2 + 1, Bs = B | 1;  // odd row stride: a lane-per-frame walk of the rows is conflict free
  L.nfft = MF > 2048 ? 1 : (MF > 1024 ? 2 : kXF);  // 16 bytes per point
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += (bytes + 15) & ~(size_t)15;
    return r;
  };
  L.fft = take(sizeof(double2) * (size_t)L.nfft * MF);
  L.tw = take(MF <= 1024 ? sizeof(double2) * (size_t)(MF / 2) : 16);  // the factors every stage reads
  L.fb = take(sizeof(double) * B);
  L.x10 = take(sizeof(double) * B);
  L.lmel = take(sizeof(double) * kXB * (size_t)((n_mel > 0 ? n_mel : 1) | 1));
  // as many magnitude rows as fit beside them (227 KB per CTA)
  const size_t budget = (size_t)227 * 1024, row = sizeof(double) * Bs;
  int rows = o + row <= budget ? (int)((budget - o) / row) : 0;
  const int halo = all_frames ? 1 : 0;
  L.batch = rows - halo > kXB ? kXB : rows - halo;
  if (L.batch < 0) L.batch = 0;
  L.rows = L.batch + halo;
  L.mags = take(row * (size_t)(L.rows > 0 ? L.rows : 1));
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

// One radix-2 decimation-in-time transform of length 2^logm in place (input already at the bit-reversed positions):
// go-dsp's butterfly graph, t = r[i2] * factor[(M / stage) j]; r[i1] +- t, by the 64 threads of a warp pair.
template <class Sync>
__device__ __forceinline__ void radix2_stages(double2* __restrict__ buf, const double2* __restrict__ tw, bool tw_brev,
                                              int logm, int pl, Sync&& pair_sync) {
  const int MF = 1 << logm;
#pragma unroll 1
  for (int sh = 0; sh < logm; ++sh) {  // stage = 2 << sh
    const int s2 = 1 << sh, tshift = logm - 1 - sh;  // factor index = (M / stage) * j
#pragma unroll 8
    for (int b = pl; b < MF / 2; b += 64) {
      const int j = b & (s2 - 1), i1 = ((b >> sh) << (sh + 1)) | j, i2 = i1 + s2;
      const double2 r1 = buf[i1], r2 = buf[i2];
      double2 w = r2;
      if (sh != 0) {
        // factor (M / stage) j.  The shared-memory copy of the table is stored bit-reversed: entry brev(j 2^tshift) =
        // the sh-bit reversal of j, so a stage reads one contiguous block of s2 entries (the natural order put the
        // lanes of the middle stages 128 .. 256 bytes apart: 8 .. 32-way bank conflicts, half of all wavefronts)
        const double2 fc = tw_brev ? tw[__brev((unsigned)j) >> (32 - sh)] : tw[j << tshift];
        w = make_double2(r2.x * fc.x - r2.y * fc.y, r2.x * fc.y + r2.y * fc.x);
      }
      buf[i1] = make_double2(r1.x + w.x, r1.y + w.y);
      buf[i2] = make_double2(r1.x - w.x, r1.y - w.y);
    }
    pair_sync();
  }
}

// LOGM: log2 of the transform length.  BLUE: the window length a.N is not a power of two (go-dsp's Bluestein route,
// transform length 2^LOGM = NextPowerOf2(2 N - 1)); otherwise N = 2^LOGM.  ALL: every frame 0 .. T-1 of every stream
// (lengths without a fused FP32 kernel; also writes the flux and, on request, the spectrum itself) instead of the lists.
template <int LOGM, bool BLUE, bool ALL>
__global__ void __launch_bounds__(kXT, 1) spectral_exact_kernel(const StftArgs a) {
  extern __shared__ __align__(16) unsigned char xsm[];
  int lg = LOGM;  // LOGM == 0: taken from the arguments (the all-frames route is not the fast path)
  if (LOGM == 0)
    while ((1 << lg) < a.fft_len) ++lg;
  const int logm = lg, MF = 1 << logm;
  const int N = BLUE ? a.N : MF, B = N / 2 + 1, Bs = B | 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const XSmem L = x_layout(N, MF, a.n_mel, ALL);
  double* mags = reinterpret_cast<double*>(xsm + L.mags);                 // [rows][Bs]
  const int pair = warp >> 1, pl = tid & 63;  // transform slot, lane of the pair
  const int nfft = L.nfft, BT = L.batch;  // BT <= 32 frames per batch; the halo row (flux) is row BT
  double2* buf = reinterpret_cast<double2*>(xsm + L.fft) + (size_t)(pair < nfft ? pair : 0) * MF;  // this pair's transform
  auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory"); };
  // go-dsp factors [0, MF / 2): in shared memory up to 1024 points, from the plan's table (L1) beyond
  const double2* tw = MF <= 1024 ? reinterpret_cast<const double2*>(xsm + L.tw) : a.fac64;
  double* fb = reinterpret_cast<double*>(xsm + L.fb);                     // bin frequencies
  double* x10 = reinterpret_cast<double*>(xsm + L.x10);                   // log10 of them
  double* lmel = reinterpret_cast<double*>(xsm + L.lmel);                 // [kXB][n_mel | 1] log mel energies
  const int lms = a.n_mel | 1;

  if (MF <= 1024)  // bit-reversed over logm - 1 bits, see radix2_stages
    for (int k = tid; k < MF / 2; k += kXT)
      reinterpret_cast<double2*>(xsm + L.tw)[logm > 1 ? (__brev((unsigned)k) >> (33 - logm)) : 0] = a.fac64[k];
  for (int k = tid; k < B; k += kXT) {  // spectral_centroid.go:59-65
    const double f = (double)k * (double)a.algo_sr / (double)((B - 1) * 2);
    fb[k] = f;
    x10[k] = f > 0 ? log10(f) : 0.0;
  }

  // one frame: window -> transform -> |X_k| into `mrow` (and the spectrum itself to the materialising outputs)
  auto transform = [&](const double* __restrict__ fr, double* __restrict__ mrow, int64_t t, int s) {
    if (!BLUE) {
      // analyzers/spectral.go:477-480, stored at the bit-reversed position; thread p of the pair takes the samples
      // (N / 64) p + u so that the stores of one instruction fall on consecutive slots
#pragma unroll 16
      for (int u = 0; u < MF / 64; ++u) {
        const int i = (MF / 64) * pl + u;
        buf[__brev((unsigned)i) >> (32 - logm)] = make_double2(fr[i] * a.win64[i], 0.0);
      }
      pair_sync();
      radix2_stages(buf, tw, MF <= 1024, logm, pl, pair_sync);
    } else {
      // go-dsp fft/bluestein.go: a[i] = x[i] conj(w_i) zero padded; r = IFFT(FFT(a) .* FFT(b)); X[k] = r[k] conj(w_k)
      for (int i = pl; i < MF; i += 64) buf[i] = make_double2(0.0, 0.0);
      pair_sync();
      for (int i = pl; i < N; i += 64) {
        const double xr = fr[i] * a.win64[i], xi = 0.0;
        const double2 c = a.chirp_inv[i];
        buf[__brev((unsigned)i) >> (32 - logm)] = make_double2(xr * c.x - xi * c.y, xr * c.y + xi * c.x);
      }
      pair_sync();
      radix2_stages(buf, tw, MF <= 1024, logm, pl, pair_sync);
      for (int i = pl; i < MF; i += 64) {  // Convolve: pointwise product with the chirp spectrum
        const double2 v = buf[i], g = a.blue_fb[i];
        buf[i] = make_double2(v.x * g.x - v.y * g.y, v.x * g.y + v.y * g.x);
      }
      pair_sync();
      for (int i = 1 + pl; i < MF / 2; i += 64) {  // IFFT: the input reversed modulo MF (r[i] <-> r[MF - i]) ...
        const double2 u = buf[i], v = buf[MF - i];
        buf[i] = v;
        buf[MF - i] = u;
      }
      pair_sync();
      for (int i = pl; i < MF; i += 64) {  // ... and moved to the bit-reversed positions (an involution: swaps)
        const int r = (int)(__brev((unsigned)i) >> (32 - logm));
        if (i < r) {
          const double2 u = buf[i], v = buf[r];
          buf[i] = v;
          buf[r] = u;
        }
      }
      pair_sync();
      radix2_stages(buf, tw, MF <= 1024, logm, pl, pair_sync);
      for (int k = pl; k < B; k += 64) {  // r / MF, times conj(w_k)
        const double2 v = make_double2(buf[k].x / (double)MF, buf[k].y / (double)MF), c = a.chirp_inv[k];
        buf[k] = make_double2(v.x * c.x - v.y * c.y, v.x * c.y + v.y * c.x);
      }
      pair_sync();
    }
#pragma unroll 4
    for (int k = pl; k < B; k += 64) {
      const double2 v = buf[k];
      const double mg = go_hypot_dev(v.x, v.y);
      mrow[k] = mg;
      if (ALL && a.mag && t >= 0) {  // sonar_stft_f64: |X|, atan2(im, re), (re, im)  (analyzers/spectral.go:490-494)
        const int64_t o = ((int64_t)s * a.T + t) * B + k;
        a.mag[o] = mg;
        if (a.phase) a.phase[o] = atan2(v.y, v.x);
        if (a.cplx) {
          a.cplx[2 * o] = v.x;
          a.cplx[2 * o + 1] = v.y;
        }
      }
    }
    pair_sync();
  };

  for (int s = blockIdx.y; s < a.n_streams; s += gridDim.y) {
    const int* lst = ALL ? nullptr : a.xlist + (int64_t)s * a.xlist_stride;
    const int64_t count = ALL ? a.T : (int64_t)lst[0];
    const double* __restrict__ x = a.pcm + (int64_t)s * a.stride;
    double* __restrict__ fo = a.feat ? a.feat + (int64_t)s * a.feat_stride : nullptr;
    for (int64_t b0 = (int64_t)blockIdx.x * BT; b0 < count; b0 += (int64_t)gridDim.x * BT) {
      const int nb = count - b0 < BT ? (int)(count - b0) : BT;
      __syncthreads();  // the previous batch's readers are done (and the tables are written)
      // ---- phase A: transforms, one frame per pair of warps at a time (ALL: + the frame before the batch, whose
      //      magnitudes the first flux needs, in the halo row) ----------------------------------------------------
      if (pair < nfft) {
        const int nfr = nb + ((ALL && b0 > 0) ? 1 : 0);
        for (int f = pair; f < nfr; f += nfft) {
          const bool halo = f == nb;
          const int64_t t = ALL ? (halo ? b0 - 1 : b0 + f) : (int64_t)lst[1 + b0 + f];
          transform(x + t * a.hop, mags + (size_t)(halo ? BT : f) * Bs, halo ? -1 : t, s);
        }
      }
      __syncthreads();
      if (fo == nullptr) continue;  // materialising call: the spectrum is all that was asked for
      // ---- phase B: lane = frame, every sum left to right -------------------------------------------------------
      const bool live = lane < nb && fo != nullptr;
      const int64_t t = ALL ? b0 + lane : (lane < nb ? (int64_t)lst[1 + b0 + lane] : 0);
      const double* __restrict__ m = mags + (size_t)(lane < nb ? lane : 0) * Bs;
      // ---- B1: flatness (spectral_flatness.go:31-70) and slope (spectral_slope.go:24-64) need ln|X_k| of every bin.
      //      513 logarithms one after the other on a single warp were the whole kernel's critical path (~0.4 ms per
      //      batch), so ALL warps evaluate them, a chunk of bins at a time, into the idle transform buffers
      //      ([bin][frame]); warps 1 and 2 then add the chunk up in order.  A skipped bin adds + 0.0: the sums never hold
      //      -0.0, so that is the skipped sum.  (log10 y = ln y / ln 10 to the last bits.)
      {
        double* lnbuf = reinterpret_cast<double*>(xsm + L.fft);
        int CH = nfft * MF / 16;  // bins per chunk: kXB * CH doubles fit the transform buffers
        CH = CH > 128 ? 128 : CH;
        double log_sum = 0.0, am = 0.0, sx = 0.0, sy = 0.0, sxy = 0.0, sxx = 0.0;
        int valid = 0, nreg = 0;
        const double inv_ln10 = 0.43429448190325182765;
        for (int c0 = 0; c0 < B; c0 += CH) {
          const int cw = B - c0 < CH ? B - c0 : CH;
#pragma unroll 4
          for (int e = tid; e < kXB * cw; e += kXT) {
            const int f = e & (kXB - 1), j = e >> 5;
            const double v = mags[(size_t)(f < nb ? f : 0) * Bs + c0 + j];
            lnbuf[j * kXB + f] = v > 1e-10 ? log(v) : 0.0;
          }
          __syncthreads();
          if (warp == 1) {
#pragma unroll 4
            for (int j = 0; j < cw; ++j) {
              const double v = m[c0 + j];
              log_sum += lnbuf[j * kXB + lane];
              valid += v > 1e-10 ? 1 : 0;
              am += v;
            }
          } else if (warp == 2) {
#pragma unroll 4
            for (int j = 0; j < cw; ++j) {
              const bool oks = m[c0 + j] > 1e-10 && fb[c0 + j] > 0;
              const double xx = oks ? x10[c0 + j] : 0.0, yy = oks ? lnbuf[j * kXB + lane] * inv_ln10 : 0.0;
              sx += xx;
              sy += yy;
              sxy += xx * yy;
              sxx += xx * xx;
              nreg += oks ? 1 : 0;
            }
          }
          __syncthreads();
        }
        if (warp == 1) {
          double fl = 0.0;
          if (valid > 0) {
            const double gm = exp(log_sum / (double)valid);
            am /= (double)B;
            if (am > 1e-10) {
              fl = gm / am;
              if (fl > 1.0) fl = 1.0;
            }
          }
          if (live) fo[a.o_flatness + t] = fl;
        } else if (warp == 2) {
          double sl = 0.0;
          if (B >= 2 && nreg >= 2) {
            const double den = (double)nreg * sxx - sx * sx;
            if (den != 0) sl = ((double)nreg * sxy - sx * sy) / den;
          }
          if (live) fo[a.o_slope + t] = sl;
        }
      }
      // ---- B2: the other chains, lane = frame ----
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
      } else if (warp == 3 && ALL) {
        // flux against the previous frame (spectral_flux.go:17-36): flux[t - 1] for t >= 1
        const double* __restrict__ pm = mags + (size_t)(lane > 0 ? (lane < nb ? lane - 1 : 0) : BT) * Bs;
        double sum = 0.0;
#pragma unroll 4
        for (int i = 0; i < B; i++) {
          const double d = m[i] - pm[i];
          sum += d > 0 ? d * d : 0.0;
        }
        if (live && t >= 1) fo[a.o_flux + t - 1] = sqrt(sum);
      } else if (warp >= 4 && a.mfcc_on) {  // mel energies (mel_scale.go:58-105), zero weights skipped: warps 4..7 share the filters
        const int q = (a.n_mel + 3) / 4, f0 = (warp - 4) * q, f1 = (f0 + q < a.n_mel) ? f0 + q : a.n_mel;
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
      if (warp >= 4 && a.mfcc_on) {  // DCT-II + lifter (mfcc.go:215-245): the coefficients split between four warps
        const int q = (a.n_mfcc + 3) / 4, c0 = (warp - 4) * q, c1 = (c0 + q < a.n_mfcc) ? c0 + q : a.n_mfcc;
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

namespace {
template <int LOGM, bool BLUE, bool ALL>
int launch_x(const StftArgs& a, cudaStream_t st) {
  const XSmem L = x_layout(a.N, a.fft_len, a.n_mel, ALL);
  if (L.batch < 1 || L.total > 227 * 1024)
    return set_error(SONAR_ERR_UNSUPPORTED, "float64 frame evaluation: window / mel bank too large for one CTA");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // one resident wave (one CTA per SM); a stream's frames are walked by up to T / 32 CTAs
  int gx = sms / a.n_streams;
  if (gx < 1) gx = 1;
  const int64_t maxb = (a.T + L.batch - 1) / L.batch;
  if ((int64_t)gx > maxb) gx = (int)maxb;
  const int gy = a.n_streams < sms ? a.n_streams : sms;
  SONAR_CUDA(cudaFuncSetAttribute(spectral_exact_kernel<LOGM, BLUE, ALL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)L.total));
  prof_begin("spectral_exact_kernel", st);
  spectral_exact_kernel<LOGM, BLUE, ALL><<<dim3((unsigned)gx, (unsigned)gy), kXT, L.total, st>>>(a);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}
}  // namespace

// Listed frames of the fused kernels (a.xlist, N = 512 / 1024), or -- a.exact_all -- every frame of a length that has no
// fused kernel (any N in [8, 1024]: radix-2 for powers of two, go-dsp's Bluestein otherwise).
int launch_spectral_exact(const StftArgs& a, cudaStream_t st) {
  if (a.n_streams <= 0 || a.T <= 0) return SONAR_OK;
  if (a.n_mel > 64) return set_error(SONAR_ERR_UNSUPPORTED, "at most 64 mel filters");
  if (a.exact_all) return a.fft_len != a.N ? launch_x<0, true, true>(a, st) : launch_x<0, false, true>(a, st);
  if (!a.xlist) return SONAR_OK;
  if (a.N == 1024 && a.fft_len == 1024) return launch_x<10, false, false>(a, st);
  if (a.N == 512 && a.fft_len == 512) return launch_x<9, false, false>(a, st);
  return set_error(SONAR_ERR_UNSUPPORTED, "listed-frame re-evaluation: N = 512 or 1024");
}

}  // namespace sonar
