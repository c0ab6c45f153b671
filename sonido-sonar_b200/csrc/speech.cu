// The data-parallel part of the speech-specific feature group:
//   SpeechAnalyzer.detectSpeech            algorithms/speech/speech_analysis.go:113-207   (the IsSpeech gate)
//   extractSpectralTilt                    fingerprint/extractors/speech.go:551-584
// (extractVoicingProbability, speech.go:529-549, is the pitch detector's voicing on the same pre-emphasised frames as the
// harmonic block: yin32.cu + the tracker's first pass in yin.cu; pause durations and the speech rate are O(T) scans of the
// short-time energies on the host side, fingerprint_api.cu.)
// float64 in the reference's order (compiled with -fmad=false).  One exception, stated: the RMS level of the gate is a
// sum over the WHOLE signal; it is taken hierarchically (fixed order: 256 partial sums per stream, combined in order),
// so the comparison `rms < 0.001` can differ from the sequential sum only when the level is within ~1e-12 of 0.001.
#include <cmath>

#include "common.h"

namespace sonar {
namespace {

constexpr int kGateParts = 256;  // partial sums per stream
constexpr int kGateThreads = 256;

// partial zero-crossing counts and sums of squares of the pre-emphasised signal y[i] = x[i] - alpha x[i-1]
__global__ void __launch_bounds__(kGateThreads) speech_gate_partials_kernel(const double* __restrict__ pcm, int64_t n,
                                                                           int64_t stride, double alpha,
                                                                           double* __restrict__ parts, int64_t parts_stride) {
  __shared__ double s_sum[kGateThreads];
  __shared__ unsigned long long s_cnt[kGateThreads];
  const int s = blockIdx.y, part = blockIdx.x;
  const double* __restrict__ x = pcm + (int64_t)s * stride;
  const int64_t per = (n + kGateParts - 1) / kGateParts;
  const int64_t lo = (int64_t)part * per, hi = (lo + per < n) ? lo + per : n;
  double sum = 0.0;
  unsigned long long cnt = 0;
  // thread t takes a contiguous slice of the part: its own left-to-right order
  const int64_t len = hi > lo ? hi - lo : 0, slice = (len + kGateThreads - 1) / kGateThreads;
  const int64_t a = lo + (int64_t)threadIdx.x * slice, b = (a + slice < hi) ? a + slice : hi;
  if (a < b) {
    double xm1 = a > 0 ? x[a - 1] : 0.0, xm2 = a > 1 ? x[a - 2] : 0.0;
    double yprev = xm1 - alpha * xm2;  // y[a-1] (unused when a == 0)
    for (int64_t i = a; i < b; ++i) {
      const double xi = x[i], y = xi - alpha * xm1;
      sum += y * y;
      if (i >= 1 && ((yprev >= 0 && y < 0) || (yprev < 0 && y >= 0))) ++cnt;
      xm1 = xi;
      yprev = y;
    }
  }
  s_sum[threadIdx.x] = sum;
  s_cnt[threadIdx.x] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    unsigned long long c = 0;
    for (int k = 0; k < kGateThreads; ++k) {
      t += s_sum[k];
      c += s_cnt[k];
    }
    double* p = parts + (int64_t)s * parts_stride + part * 2;
    p[0] = t;
    p[1] = (double)c;  // < 2^53
  }
}

// combines the partials, runs checkPeriodicity on the first 1024 samples, writes gate[0] = 1.0 / 0.0 (+ diagnostics)
__global__ void __launch_bounds__(512) speech_gate_kernel(const double* __restrict__ pcm, int64_t n, int64_t stride,
                                                          double alpha, int sr, const double* __restrict__ parts,
                                                          int64_t parts_stride, double* __restrict__ feat, int64_t feat_stride, int64_t o_gate) {
  __shared__ double f[1024];
  __shared__ double corr[512];
  const int s = blockIdx.x, t = threadIdx.x;
  const double* __restrict__ x = pcm + (int64_t)s * stride;
  double* g = feat + (int64_t)s * feat_stride + o_gate;
  const bool has = n >= 1024;
  if (has)
    for (int i = t; i < 1024; i += 512) f[i] = x[i] - alpha * (i > 0 ? x[i - 1] : 0.0);
  __syncthreads();
  double c = 0.0;
  const int lag = t;
  const bool mine = has && lag >= 20 && lag < 400;  // lag < maxLag && lag < len / 2 (speech_analysis.go:186)
  if (mine) {
    double acc = 0.0;
    for (int i = 0; i < 1024 - lag; ++i) acc += f[i] * f[i + lag];
    c = acc / (double)(1024 - lag);
  }
  corr[t] = (mine && c > 0.0) ? c : 0.0;  // maxCorr starts at 0 and only grows
  __syncthreads();
  if (t == 0) {
    double tot = 0.0, cnt = 0.0;
    const double* p = parts + (int64_t)s * parts_stride;
    for (int k = 0; k < kGateParts; ++k) {
      tot += p[2 * k];
      cnt += p[2 * k + 1];
    }
    const double zcr = n <= 1 ? 0.0 : cnt / (double)(n - 1);
    const double rms = sqrt(tot / (double)n);
    double mx = 0.0;
    for (int k = 20; k < 400; ++k) mx = corr[k] > mx ? corr[k] : mx;
    double energy = 0.0;
    if (has)
      for (int i = 0; i < 1024; ++i) energy += f[i] * f[i];
    energy /= 1024.0;
    if (energy > 0) mx /= energy;
    bool ok = n >= (int64_t)(sr / 4);
    ok = ok && !(zcr < 0.01 || zcr > 0.3);
    ok = ok && !(rms < 0.001);
    ok = ok && has && mx > 0.1;
    g[0] = ok ? 1.0 : 0.0;
    g[1] = zcr;
    g[2] = rms;
    g[3] = mx;
  }
}

// extractSpectralTilt: one thread per frame, both sums left to right over the frame's pre-emphasised samples
__global__ void __launch_bounds__(128) speech_tilt_kernel(const double* __restrict__ pcm, int64_t n, int64_t stride,
                                                          double alpha, int64_t nf, double* __restrict__ feat,
                                                          int64_t feat_stride, int64_t o_tilt) {
  const int s = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nf) return;
  const double* __restrict__ x = pcm + (int64_t)s * stride;
  const int64_t s0 = i * 512, e0 = (s0 + 1024 < n) ? s0 + 1024 : n;
  double hi = 0.0, lo = 0.0;
  if (s0 < e0) {
    double xm1 = x[s0], yprev = xm1 - alpha * (s0 > 0 ? x[s0 - 1] : 0.0);
    for (int64_t j = s0 + 1; j < e0; ++j) {
      const double xj = x[j], y = xj - alpha * xm1, d = y - yprev;
      hi += d * d;
      lo += y * y;
      xm1 = xj;
      yprev = y;
    }
  }
  feat[(int64_t)s * feat_stride + o_tilt + i] = lo > 0 ? -10 * log10(hi / lo) : 0.0;
}

}  // namespace

size_t speech_gate_scratch_doubles() { return (size_t)kGateParts * 2; }

// gate -> feat[o_gate .. o_gate+3] per stream; tilt -> feat[o_tilt .. o_tilt+nf); scratch: per stream
// speech_gate_scratch_doubles() doubles, scratch_stride apart
int launch_speech(const double* pcm, int64_t n, int64_t stride, int n_streams, double alpha, int sr, int64_t nf, double* feat,
                  int64_t feat_stride, int64_t o_gate, int64_t o_tilt, double* scratch, int64_t scratch_stride,
                  cudaStream_t st) {
  if (n_streams <= 0) return SONAR_OK;
  prof_begin("speech_gate_partials_kernel", st);
  speech_gate_partials_kernel<<<dim3(kGateParts, (unsigned)n_streams), kGateThreads, 0, st>>>(pcm, n, stride, alpha, scratch,
                                                                                                scratch_stride);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  prof_begin("speech_gate_kernel", st);
  speech_gate_kernel<<<(unsigned)n_streams, 512, 0, st>>>(pcm, n, stride, alpha, sr, scratch, scratch_stride, feat, feat_stride,
                                                         o_gate);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  if (nf > 0) {
    prof_begin("speech_tilt_kernel", st);
    speech_tilt_kernel<<<dim3((unsigned)((nf + 127) / 128), (unsigned)n_streams), 128, 0, st>>>(pcm, n, stride, alpha, nf, feat,
                                                                                              feat_stride, o_tilt);
    prof_end();
    SONAR_CUDA(cudaGetLastError());
  }
  return SONAR_OK;
}

}  // namespace sonar
