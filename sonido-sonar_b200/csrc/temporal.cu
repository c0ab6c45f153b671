// Temporal feature group of the speech extractor (news / talk configurations), on the device.
//
//   extractTemporalFeatures    fingerprint/extractors/speech.go:370-408
//   calculateSilenceRatio      :625-653   (bubble sort -> sorted[len/10] threshold -> count)
//   detectOnsets               :657-680,  calculateAdaptiveThreshold :682-704 (mean + 2 sigma of the derivative)
//   calculateAttackTimes       :706-737
//   extractSimpleEnvelope      :739-767   (512 / 256 RMS)
//
// The reference spends O(T^2) compare-swaps on the bubble sort (6.5e10 for a one-hour stream); the
// value it extracts is just an order statistic, found here exactly with a byte-wise radix select.  All inputs are
// already on the device: the pre-emphasised PCM is recomputed from the raw samples, the short-time energies are
// the ones frame_walk_kernel wrote.
#include <cmath>

#include "common.h"

namespace sonar {
namespace {

constexpr int kAmpBlocks = 64;
constexpr int kTfThreads = 256;

// per (stream, block): sum |y| and max |y| of the pre-emphasised samples it strides over
__global__ void __launch_bounds__(256) amp_partial_kernel(const double* __restrict__ pcm, int64_t n, int64_t stride,
                                                          double alpha, double* __restrict__ tmp, int64_t tmp_stride,
                                                          int64_t o_part) {
  __shared__ double ssum[256], smax[256];
  const double* __restrict__ x = pcm + (int64_t)blockIdx.y * stride;
  double s = 0.0, m = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)kAmpBlocks * 256) {
    const double prev = i > 0 ? x[i - 1] : 0.0;
    const double a = fabs(x[i] - alpha * prev);
    s += a;
    m = fmax(m, a);
  }
  ssum[threadIdx.x] = s;
  smax[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      ssum[threadIdx.x] += ssum[threadIdx.x + o];
      smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double* part = tmp + (int64_t)blockIdx.y * tmp_stride + o_part;
    part[2 * blockIdx.x] = ssum[0];
    part[2 * blockIdx.x + 1] = smax[0];
  }
}

__device__ double tf_block_sum(double v, double* red) {
  const int t = threadIdx.x;
  red[t] = v;
  __syncthreads();
  for (int o = kTfThreads / 2; o > 0; o >>= 1) {
    if (t < o) red[t] += red[t + o];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

struct TemporalArgs {
  double* feat;
  int64_t feat_stride, o_energy, o_scalars, o_att;
  const double* tmp;
  int64_t tmp_stride, o_part;
  int64_t Te, n;
  int call_sr, algo_sr, energy_hop;
};

// one CTA per stream
__global__ void __launch_bounds__(kTfThreads) temporal_finalize_kernel(const TemporalArgs a) {
  __shared__ double red[kTfThreads];
  __shared__ unsigned hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ long long s_k;
  __shared__ int s_wsum[kTfThreads / 32];
  __shared__ int s_base;
  double* fo = a.feat + (int64_t)blockIdx.x * a.feat_stride;
  const double* __restrict__ ste = fo + a.o_energy;
  const int64_t Te = a.Te;
  const int t = threadIdx.x;

  // ---- silence ratio: threshold = sorted[Te / 10] (energies are >= 0, so the IEEE bit pattern orders them)
  double silence = 0.0;
  if (Te > 0) {
    if (t == 0) {
      s_prefix = 0ull;
      s_k = Te / 10;
    }
    __syncthreads();
    unsigned long long mask = 0ull;
    for (int pass = 7; pass >= 0; --pass) {
      hist[t] = 0u;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      for (int64_t i = t; i < Te; i += kTfThreads) {
        const unsigned long long k = (unsigned long long)__double_as_longlong(ste[i]);
        if ((k & mask) == prefix) atomicAdd(&hist[(unsigned)(k >> (8 * pass)) & 0xffu], 1u);
      }
      __syncthreads();
      if (t == 0) {
        long long k = s_k;
        int b = 0;
        for (; b < 256; ++b) {
          if (k < (long long)hist[b]) break;
          k -= hist[b];
        }
        s_k = k;
        s_prefix = prefix | ((unsigned long long)b << (8 * pass));
      }
      mask |= 0xffull << (8 * pass);
      __syncthreads();
    }
    const double thr = __longlong_as_double((long long)s_prefix);
    double cnt = 0.0;
    for (int64_t i = t; i < Te; i += kTfThreads) cnt += ste[i] <= thr ? 1.0 : 0.0;
    silence = tf_block_sum(cnt, red) / (double)Te;
  }

  // ---- onsets: local maxima of the energy derivative above mean + 2 sigma
  int n_onsets = 0;
  if (Te >= 3) {
    const int64_t nd = Te - 1;
    double s = 0.0;
    for (int64_t i = t; i < nd; i += kTfThreads) s += ste[i + 1] - ste[i];
    const double mean = tf_block_sum(s, red) / (double)nd;
    double v = 0.0;
    for (int64_t i = t; i < nd; i += kTfThreads) {
      const double d = (ste[i + 1] - ste[i]) - mean;
      v += d * d;
    }
    const double thr = mean + 2 * sqrt(tf_block_sum(v, red) / (double)nd);
    const double frame_time = (double)a.energy_hop / (double)a.algo_sr;  // +Inf when the extractor's rate is 0 (F2)
    double* att = fo + a.o_att;
    if (t == 0) s_base = 0;
    __syncthreads();
    for (int64_t base = 1; base + 1 < nd; base += kTfThreads) {
      const int64_t i = base + t;
      bool on = false;
      if (i + 1 < nd) {
        const double d0 = ste[i] - ste[i - 1], d1 = ste[i + 1] - ste[i], d2 = ste[i + 2] - ste[i + 1];
        on = d1 > d0 && d1 > d2 && d1 > thr;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, on);
      const int lane = t & 31, w = t >> 5;
      if (lane == 0) s_wsum[w] = __popc(bal);
      __syncthreads();
      int before = s_base;
      for (int q = 0; q < w; ++q) before += s_wsum[q];
      if (on) {
        const int pos = before + __popc(bal & ((1u << lane) - 1u));
        const int64_t onset = i;  // index into the derivative == energy frame index (speech.go:706-737)
        int64_t start = onset;
        const double pk = ste[onset];
        for (int64_t j = onset - 1; j >= 0 && j > onset - 10; --j)
          if (ste[j] < 0.1 * pk) {
            start = j;
            break;
          }
        double at = (double)(onset - start) * frame_time;
        if (at > 0.1) at = 0.1;
        att[pos] = at;
      }
      __syncthreads();
      if (t == 0) {
        int tot = 0;
        for (int q = 0; q < kTfThreads / 32; ++q) tot += s_wsum[q];
        s_base += tot;
      }
      __syncthreads();
    }
    n_onsets = s_base;
  }

  if (t == 0) {
    const double* part = a.tmp + (int64_t)blockIdx.x * a.tmp_stride + a.o_part;
    double sum = 0.0, peak = 0.0;
    for (int b = 0; b < kAmpBlocks; ++b) {
      sum += part[2 * b];
      peak = fmax(peak, part[2 * b + 1]);
    }
    double* sc = fo + a.o_scalars;
    sc[2] = silence;
    sc[3] = peak;
    sc[4] = a.n > 0 ? sum / (double)a.n : 0.0;
    sc[5] = (double)n_onsets / ((double)a.n / (double)a.call_sr);
    sc[6] = (double)n_onsets;
  }
}

}  // namespace

int launch_temporal(const double* pcm, int64_t n, int64_t stride, int n_streams, double alpha, double* feat,
                    int64_t feat_stride, int64_t o_energy, int64_t o_scalars, int64_t o_att, int64_t Te, double* tmp,
                    int64_t tmp_stride, int64_t o_part, int call_sr, int algo_sr, int energy_hop, cudaStream_t st) {
  if (n_streams <= 0) return SONAR_OK;
  prof_begin("amp_partial_kernel", st);
  amp_partial_kernel<<<dim3(kAmpBlocks, (unsigned)n_streams), 256, 0, st>>>(pcm, n, stride, alpha, tmp, tmp_stride, o_part);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  TemporalArgs a{feat, feat_stride, o_energy, o_scalars, o_att, tmp, tmp_stride, o_part, Te, n, call_sr, algo_sr, energy_hop};
  prof_begin("temporal_finalize_kernel", st);
  temporal_finalize_kernel<<<n_streams, kTfThreads, 0, st>>>(a);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

int temporal_partials_doubles() { return 2 * kAmpBlocks; }

}  // namespace sonar
