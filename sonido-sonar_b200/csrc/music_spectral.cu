// MusicFeatureExtractor's additions to the per-frame spectral block, on a materialised magnitude
// spectrogram (the fused kernels never materialise it; the music extractor is reached by direct construction only,
// so this is a separate, optional pass and not part of the fingerprint hot loop):
//   SpectralContrast.Compute             algorithms/spectral/spectral_contrast.go:26-187
//   ChromaSTFT.convertSTFTToChroma       algorithms/chroma/chroma_stft.go:63-138
//   BarkScale.ComputeBarkSpectrum        algorithms/spectral/bark_scale.go:36-128
// One CTA per frame.  The band edges, the bin -> pitch-class map and the Bark bank are built on the host in float64
// exactly as the reference's constructors do.  The contrast needs order statistics of every band (mean of the
// bottom / top 20 % of the sorted power): each value's rank is counted against the band in shared memory (ties by
// index, i.e. a stable sort), which is exact and embarrassingly parallel — sum of n^2 over the six bands is ~10^5
// comparisons per frame against the reference's insertion sort of the same order.
#include <cmath>

#include "common.h"

namespace sonar {
namespace {

constexpr int kMsThreads = 128;

__device__ __forceinline__ double block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < kMsThreads / 32; ++w) s += red[w];
  return s;
}

__global__ void __launch_bounds__(kMsThreads) music_spectral_kernel(const double* __restrict__ mag, int64_t T, int B,
                                                                    int n_bands, const int* __restrict__ edges,
                                                                    double* __restrict__ contrast,
                                                                    const signed char* __restrict__ cmap,
                                                                    double* __restrict__ chroma, int n_bark,
                                                                    const double* __restrict__ bank,
                                                                    const int2* __restrict__ bank_range,
                                                                    double* __restrict__ bark) {
  extern __shared__ double pw[];  // power spectrum of the frame
  __shared__ double red[kMsThreads / 32];
  __shared__ double s_chroma[12];
  const int64_t t = blockIdx.x;
  if (t >= T) return;
  const double* m = mag + t * B;
  for (int k = threadIdx.x; k < B; k += kMsThreads) {
    const double v = m[k];
    pw[k] = v * v;
  }
  __syncthreads();
  if (contrast) {
    for (int b = 0; b < n_bands; ++b) {
      const int s = edges[b];
      int e = edges[b + 1];
      e = e < B ? e : B;
      const int len = e - s;
      double res = 0.0;
      if (len > 0) {  // uniform across the CTA
        int vc = (int)(0.2 * (double)len), pc = vc;
        if (vc == 0) vc = pc = 1;
        double valley = 0.0, peak = 0.0;
        for (int i = threadIdx.x; i < len; i += kMsThreads) {
          const double x = pw[s + i];
          int rank = 0;
          for (int j = 0; j < len; ++j) {
            const double y = pw[s + j];
            rank += (y < x || (y == x && j < i)) ? 1 : 0;
          }
          if (rank < vc) valley += x;
          if (rank >= len - pc) peak += x;
        }
        valley = block_sum(valley, red) / (double)vc;
        peak = block_sum(peak, red) / (double)pc;
        if (valley <= 0) valley = 1e-10;
        res = peak <= 0 ? 0.0 : 10.0 * log10(peak / valley);
      }
      if (threadIdx.x == 0) contrast[t * n_bands + b] = res;
    }
  }
  if (chroma) {
    for (int c = 0; c < 12; ++c) {
      double acc = 0.0;
      for (int k = threadIdx.x; k < B; k += kMsThreads)
        if (cmap[k] == c) acc += pw[k];
      acc = block_sum(acc, red);
      if (threadIdx.x == 0) s_chroma[c] = acc;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
      double total = 0.0;
      for (int c = 0; c < 12; ++c) total += s_chroma[c];
      const double v = s_chroma[threadIdx.x];
      chroma[t * 12 + threadIdx.x] = total > 1e-10 ? v / total : v;
    }
  }
  if (bark) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < n_bark; i += kMsThreads / 32) {
      const int2 r = bank_range[i];  // non-zero weights live in [r.x, r.y)
      const double* row = bank + (size_t)i * B;
      double acc = 0.0;
      for (int k = r.x + lane; k < r.y; k += 32) acc += pw[k] * row[k];
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) bark[t * n_bark + i] = acc;
    }
  }
}

}  // namespace

int launch_music_spectral(const double* mag, int64_t T, int B, int n_bands, const int* edges, double* contrast,
                          const signed char* cmap, double* chroma, int n_bark, const double* bank,
                          const int2* bank_range, double* bark, cudaStream_t st) {
  if (T <= 0) return SONAR_OK;
  prof_begin("music_spectral_kernel", st);
  music_spectral_kernel<<<(unsigned)T, kMsThreads, sizeof(double) * (size_t)B, st>>>(
      mag, T, B, n_bands, edges, contrast, cmap, chroma, n_bark, bank, bank_range, bark);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

}  // namespace sonar
