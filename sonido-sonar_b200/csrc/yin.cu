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

#include "common.h"

namespace sonar {
namespace {

constexpr int kYinFrame = 1024, kYinHop = 512, kYinHalf = 512;
constexpr int kYinFpb = 4;                     // frames per CTA
constexpr int kYinTpf = 64;                    // threads per frame: 8 lags each
constexpr int kYinThreads = kYinFpb * kYinTpf;
constexpr int kYinPad = kYinFrame + kYinFrame / 8 + 8;  // p[e] stored at e + (e >> 3): stride-8 reads hit distinct banks

__device__ __forceinline__ int pidx(int e) { return e + (e >> 3); }

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
  // ---- CMNDF + first dip below 0.15 (pitch_detection.go:363-383): one warp per frame, 16 lags per lane
  const int wv = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (wv >= kYinFpb) return;
  const int64_t f = f0 + wv;
  const double* d = sd[wv];
  double loc[16], run = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int tau = lane * 16 + k;
    run += tau >= 1 ? d[tau] : 0.0;
    loc[k] = run;
  }
  double incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  const double base = incl - run;
  double cm[17];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int tau = lane * 16 + k;
    cm[k] = tau == 0 ? 1.0 : d[tau] / ((base + loc[k]) / (double)tau);
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
    double* r = raw + (int64_t)s * raw_stride;
    r[f] = pitch;
    r[Tp + f] = conf;
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
__global__ void __launch_bounds__(32) yin_track_kernel(const double* __restrict__ raw, int64_t raw_stride,
                                                       int n_streams, int64_t Tp, double* __restrict__ feat,
                                                       int64_t feat_stride, int64_t o_pitch, int64_t o_conf,
                                                       int64_t o_voicing, int64_t o_hratio, int64_t o_inharm,
                                                       int64_t o_tonal) {
  __shared__ double c_sm[5 + 32];  // corrected pitches: [0..4] = the five frames before this round
  __shared__ double raw_sm[32];
  const int s = blockIdx.x;
  const int lane = threadIdx.x;
  const double* r = raw + (int64_t)s * raw_stride;
  double* fo = feat + (int64_t)s * feat_stride;
  if (lane < 5) c_sm[lane] = 0.0;
  __syncwarp();
  for (int64_t base = 0; base < Tp; base += 32) {
    const int cnt = (int)((Tp - base < 32) ? (Tp - base) : 32);
    const int64_t i = base + lane;
    double rawp = 0.0, conf = 0.0;
    if (lane < cnt) {
      rawp = r[i];
      conf = r[Tp + i];
    }
    const bool gate = lane < cnt && rawp != 0.0 && conf >= 0.5;  // :782-786: below 0.5 everything is zeroed
    raw_sm[lane] = rawp;
    c_sm[5 + lane] = gate ? rawp : 0.0;
    unsigned todo = __ballot_sync(0xffffffffu, gate);
    __syncwarp();
    if (lane == 0) {
      while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        const int64_t hlen = base + k;  // history entries before this frame
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
    __syncwarp();
    if (lane < cnt) {
      const double c0 = c_sm[5 + lane], c1 = c_sm[4 + lane], c2 = c_sm[3 + lane];
      double pitch = c0;  // applyTemporalSmoothing :905-921 on the history that already includes this frame
      if (i >= 2)
        pitch = median_nonzero3(c2, c1, c0);
      else if (i == 1)
        pitch = 0.3 * c0 + (1 - 0.3) * c1;  // history of two: blend with the previous (unsmoothed) output
      const double cf = conf < 0.5 ? 0.0 : conf;
      fo[o_pitch + i] = pitch;
      fo[o_conf + i] = cf;
      fo[o_voicing + i] = cf;
      fo[o_hratio + i] = cf * 10.0;               // speech.go:499
      fo[o_inharm + i] = 1.0 - cf;                // speech.go:500
      fo[o_tonal + i] = pitch > 0 ? pitch : 0.0;  // speech.go:503-505
    }
    __syncwarp();
    const double carry = (lane < 5) ? c_sm[32 + lane] : 0.0;
    __syncwarp();
    if (lane < 5) c_sm[lane] = carry;
    __syncwarp();
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
               int64_t scratch_stride, cudaStream_t st) {
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
  dim3 grid((unsigned)((Tp + kYinFpb - 1) / kYinFpb), (unsigned)n_streams);
  const size_t smem = sizeof(double) * kYinFpb * (kYinPad + kYinFrame + 2 + kYinHalf);
  SONAR_CUDA(cudaFuncSetAttribute(yin_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_begin("yin_frame_kernel", st);
  yin_frame_kernel<<<grid, kYinThreads, smem, st>>>(pcm, stride, alpha, sr, Tp, hann_dev, scratch, scratch_stride);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  prof_begin("yin_track_kernel", st);
  yin_track_kernel<<<n_streams, 32, 0, st>>>(scratch, scratch_stride, n_streams, Tp, feat, feat_stride,
                                                         o_pitch, o_conf, o_voicing, o_hratio, o_inharm,
                                                         o_tonal);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

}  // namespace sonar
