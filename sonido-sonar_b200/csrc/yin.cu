// YIN pitch track of the speech extractor's harmonic block (float64).
//
//   extractHarmonicFeatures      fingerprint/extractors/speech.go:464-509
//   PitchDetector.DetectPitch    algorithms/tonal/pitch_detection.go:225-279
//   preprocessFrame              :282-314 (second pre-emphasis 0.97 + un-normalised Hann)
//   detectPitchYin               :349-420 (difference function, CMNDF, first dip < 0.15)
//   parabolicInterpolation       :743-764
//   postProcessResult / updateTemporalTracking   :767-921,978-1007 (sequential over frames)
//
// Kernel 1 (one CTA per 1024/512 frame) produces the raw (frequency, confidence)
// of every frame; kernel 2 (one thread per stream) replays the reference's
// sequential 20-deep history logic (octave correction vs the median of the last
// five, confidence gate, median-of-three smoothing) and writes the six
// HarmonicFeatures arrays.
#include <cmath>

#include "common.h"

namespace sonar {
namespace {

constexpr int kYinFrame = 1024, kYinHop = 512, kYinHalf = 512, kYinThreads = 256;

__global__ void __launch_bounds__(kYinThreads) yin_frame_kernel(const double* __restrict__ pcm, int64_t stride,
                                                                double alpha, int sr, int64_t Tp,
                                                                const double* __restrict__ hann,
                                                                double* __restrict__ raw, int64_t raw_stride) {
  __shared__ double p[kYinFrame];
  __shared__ double d[kYinHalf];
  __shared__ double cm[kYinHalf];
  __shared__ double wsum[kYinThreads / 32];
  __shared__ int s_min;
  const int64_t f = blockIdx.x;
  const int s = blockIdx.y;
  const double* __restrict__ x = pcm + (int64_t)s * stride;
  const int64_t s0 = f * kYinHop;
  const int tid = threadIdx.x;
  // stream-level pre-emphasis (speech.go:161) then the detector's own (pitch_detection.go:299-314)
  for (int i = tid; i < kYinFrame; i += kYinThreads) {
    const int64_t g = s0 + i;
    const double xm1 = g > 0 ? x[g - 1] : 0.0, xm2 = g > 1 ? x[g - 2] : 0.0;
    const double y = x[g] - alpha * xm1;
    double v = y;
    if (i > 0) {
      const double ym1 = xm1 - alpha * xm2;
      v = y - 0.97 * ym1;
    }
    p[i] = v * hann[i];
  }
  if (tid == 0) s_min = kYinHalf;
  __syncthreads();
  // difference function d[tau] = sum_j (p[j] - p[j+tau])^2
  for (int tau = tid; tau < kYinHalf; tau += kYinThreads) {
    double acc = 0.0;
#pragma unroll 8
    for (int j = 0; j < kYinHalf; ++j) {
      const double dl = p[j] - p[j + tau];
      acc += dl * dl;
    }
    d[tau] = acc;
  }
  __syncthreads();
  // inclusive prefix sum of d[1..] (two values per thread, block scan)
  {
    const int i0 = 2 * tid, i1 = 2 * tid + 1;
    const double a = i0 >= 1 ? d[i0] : 0.0, b = d[i1];
    double v = a + b;
    const int lane = tid & 31, w = tid >> 5;
    double incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    double base = 0.0;
    for (int k = 0; k < w; ++k) base += wsum[k];
    const double run1 = base + incl;      // running sum through i1
    const double run0 = run1 - b;          // running sum through i0
    cm[i0] = i0 == 0 ? 1.0 : d[i0] / (run0 / (double)i0);
    cm[i1] = d[i1] / (run1 / (double)i1);
  }
  __syncthreads();
  for (int tau = tid; tau < kYinHalf; tau += kYinThreads)
    if (tau >= 1 && tau + 1 < kYinHalf && cm[tau] < 0.15 && cm[tau] < cm[tau + 1]) atomicMin(&s_min, tau);
  __syncthreads();
  if (tid == 0) {
    double pitch = 0.0, conf = 0.0;
    const int mt = s_min;
    if (mt < kYinHalf && mt > 0) {
      double period = (double)mt;
      if (!(mt <= 0 || mt >= kYinHalf - 1)) {
        const double y1 = cm[mt - 1], y2 = cm[mt], y3 = cm[mt + 1];
        const double a = (y1 - 2 * y2 + y3) / 2, b = (y3 - y1) / 2;
        if (a != 0) period = (double)mt + (-b / (2 * a));
      }
      const double freq = (double)sr / period;
      if (freq >= 80.0 && freq <= 1000.0) {
        pitch = freq;
        conf = 1.0 - cm[mt];
      }
    }
    double* r = raw + (int64_t)s * raw_stride;
    r[f] = pitch;
    r[Tp + f] = conf;
  }
}

__device__ double median_nonzero(const double* v, int n) {  // pitch_detection.go:978-1007
  double f[5];
  int m = 0;
  for (int i = 0; i < n; i++)
    if (v[i] > 0) f[m++] = v[i];
  if (m == 0) return 0.0;
  for (int i = 1; i < m; i++) {  // insertion sort, m <= 5
    const double key = f[i];
    int j = i - 1;
    while (j >= 0 && f[j] > key) {
      f[j + 1] = f[j];
      j--;
    }
    f[j + 1] = key;
  }
  return (m % 2 == 0) ? (f[m / 2 - 1] + f[m / 2]) / 2.0 : f[m / 2];
}

__global__ void yin_track_kernel(const double* __restrict__ raw, int64_t raw_stride, int n_streams, int64_t Tp,
                                 double* __restrict__ feat, int64_t feat_stride, int64_t o_pitch, int64_t o_conf,
                                 int64_t o_voicing, int64_t o_hratio, int64_t o_inharm, int64_t o_tonal) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_streams) return;
  const double* r = raw + (int64_t)s * raw_stride;
  double* fo = feat + (int64_t)s * feat_stride;
  double hist[5] = {0, 0, 0, 0, 0};  // last five history entries, hist[4] newest
  int64_t hlen = 0;
  double previous = 0.0;
  for (int64_t i = 0; i < Tp; i++) {
    double pitch = r[i], conf = r[Tp + i], voicing = conf;
    // applyOctaveCorrection :792-829
    if (!(pitch == 0.0 || hlen == 0)) {
      const int cnt = hlen < 5 ? (int)hlen : 5;
      if (cnt >= 3) {
        const double med = median_nonzero(hist + 5 - cnt, cnt);
        const double ratios[4] = {0.5, 2.0, 1.0 / 3.0, 3.0};
        for (int k = 0; k < 4; k++) {
          const double expect = med * ratios[k];
          if (fabs(pitch - expect) / expect < 0.1) {
            if (fabs(pitch - med) > fabs(expect - med)) pitch = expect;
            break;
          }
        }
      }
    }
    if (conf < 0.5) {  // :782-786
      pitch = 0.0;
      conf = 0.0;
      voicing = 0.0;
    }
    // updateTemporalTracking :876-902 (only the last five entries are ever read)
    hist[0] = hist[1];
    hist[1] = hist[2];
    hist[2] = hist[3];
    hist[3] = hist[4];
    hist[4] = pitch;
    hlen++;
    if (hlen > 1) {  // applyTemporalSmoothing :905-921
      if (hlen >= 3)
        pitch = median_nonzero(hist + 2, 3);
      else
        pitch = 0.3 * pitch + (1 - 0.3) * previous;
    }
    previous = pitch;
    fo[o_pitch + i] = pitch;
    fo[o_conf + i] = conf;
    fo[o_voicing + i] = voicing;
    fo[o_hratio + i] = voicing * 10.0;          // speech.go:499
    fo[o_inharm + i] = 1.0 - voicing;           // speech.go:500
    fo[o_tonal + i] = pitch > 0 ? pitch : 0.0;  // speech.go:503-505
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
  dim3 grid((unsigned)Tp, (unsigned)n_streams);
  prof_begin("yin_frame_kernel", st);
  yin_frame_kernel<<<grid, kYinThreads, 0, st>>>(pcm, stride, alpha, sr, Tp, hann_dev, scratch, scratch_stride);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  prof_begin("yin_track_kernel", st);
  yin_track_kernel<<<(n_streams + 63) / 64, 64, 0, st>>>(scratch, scratch_stride, n_streams, Tp, feat, feat_stride,
                                                         o_pitch, o_conf, o_voicing, o_hratio, o_inharm,
                                                         o_tonal);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

}  // namespace sonar
