// YIN difference function on the packed-FP32 pipe, with an exact float64 re-evaluation of every borderline frame.
//
//   PitchDetector.preprocessFrame / detectPitchYin / parabolicInterpolation
//                                    algorithms/tonal/pitch_detection.go:282-314, 349-420, 743-764
//   (called per 1024 / 512 frame of the pre-emphasised stream by extractHarmonicFeatures, extractors/speech.go:464-509)
//
// The reference evaluates d[tau] = sum_{j<512} (p[j] - p[j+tau])^2 directly: 262,144 multiply-adds per frame, the
// largest flop count of the fingerprint.  yin.cu replaced that by float64 FFTs in shared memory (17.5 ms of the
// 29.6 ms bench step, bound by shared-memory wavefronts and barriers).  Here:
//
//  1. yin32_kernel -- ONE WARP owns a pair of frames.  The frame is staged in float64 exactly as the reference does
//     (stream pre-emphasis, the detector's own pre-emphasis, Hann), scaled by a power of two and rounded to FP32 once.
//     d[tau] = E(0) + E(tau) - 2 r[tau];  r = IFFT(conj(U) P) comes from 1024-point transforms built like the STFT's
//     (stft_v3.cu): radix-32 in registers on FADD2 / FMUL2 / FFMA2, one transposition through a padded tile, radix-32
//     again; z = u + i v (the two halves of the frame) gives U and P from one forward transform per frame, and the two
//     frames' products share one inverse transform (real part / imaginary part).  CMNDF, first dip below 0.15 and the
//     parabolic refinement follow in FP32, 16 lags per lane.
//  2. Every decision the reference takes on the CMNDF (cm < 0.15, cm[tau] < cm[tau+1], the 80..1000 Hz gate) is
//     checked against an a-priori bound on the FP32 error of cm (kKappa * frame energy, propagated through the
//     division).  A frame with any comparison inside the bound -- or whose refinement is ill conditioned, or whose
//     samples are not finite -- is appended to its stream's list.
//  3. yin_exact_kernel re-evaluates the listed frames in float64 IN THE REFERENCE'S ORDER (sequential sums over j and
//     over tau, no FMA contraction: this file is compiled with -fmad=false), so their raw pitch / confidence are bit
//     identical to the reference's.  For all other frames the discrete choices (tau, voiced or not) provably agree and
//     the values agree to ~1e-6 relative (asserted at 1e-4 in the tests, the north star's feature tolerance).
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.h"
#include "fft_packed.cuh"

namespace sonar {
namespace {

constexpr int kYW = 8;                  // warps per CTA, one CTA per SM: TWO per scheduler.  Measured per 64 x 300 s (ms): 4 warps
                                        // 8.46, 6 warps 10.56, 8 warps 8.01, 10 warps 8.99 -- the kernel is bound by instruction
                                        // delivery (its loop body exceeds the 32 KB instruction cache level), a second warp per
                                        // scheduler adds 5 %, and a scheduler left with fewer warps than its neighbours (6, 10)
                                        // makes the whole CTA wait for the crowded ones.  21 KB of shared memory per warp.
constexpr unsigned kFullY = 0xffffffffu;
constexpr int kN = 1024, kHalf = 512, kHop = 512;
constexpr int kTileRowY = 34;           // exchange tile row stride (float2)
constexpr float kKappa = 2e-6f;         // |d32[tau] - d[tau]| <= kKappa * (frame energy): several times the observed FP32
                                        // error of the transform route (|cm32 - cm| ~ 3e-7 measured, scripts/dev_check_yin.py)
constexpr float kThresh = 0.15f;

__device__ __forceinline__ int pidx(int i) { return i + (i >> 5); }  // one pad per 32: stride-32 and stride-1 both conflict free

struct Y32Smem {
  size_t tw, hann, warp0, per_warp, total;
  size_t w_tile, w_p, w_e;  // w_p: two buffers (frame a, frame b)
};
__host__ __device__ inline Y32Smem y32_layout() {
  Y32Smem L;
  L.tw = 0;
  L.hann = sizeof(float2) * 32 * 32;
  size_t o = L.hann + sizeof(double) * (kN + 32);
  o = (o + 127) & ~(size_t)127;
  L.warp0 = o;
  size_t w = 0;
  auto wtake = [&](size_t bytes) {
    size_t r = w;
    w += (bytes + 127) & ~(size_t)127;
    return r;
  };
  L.w_tile = wtake(sizeof(float2) * 32 * kTileRowY);  // also the float64 staging buffer (1026 samples + pads) and, after
                                                       // the inverse transform, r of both frames (float2 x 512)
  L.w_p = wtake(sizeof(float) * 2 * (kN + 32));        // p of both frames in FP32 (kept for the direct refinement)
  L.w_e = wtake(sizeof(float) * 2 * (kHalf + 4));      // E(tau) of the two frames
  L.per_warp = w;
  L.total = o + w * kYW;
  return L;
}

template <int K>
__device__ __forceinline__ void tw_apply_y(float2 (&v)[32], const float2* __restrict__ tw) {
  if constexpr (K < 32) {
    v[K] = pk::mul(v[K], tw[K * 32]);
    tw_apply_y<K + 1>(v, tw);
  }
}

// 1024-point forward transform of the warp: lane l holds x[l + 32 j] in v[j]; returns X[l + 32 k2] in v[k2].
__device__ __forceinline__ void warp_fft1024(float2 (&v)[32], float2* __restrict__ tile, const float2* __restrict__ s_tw,
                                             int lane) {
  // ONE instance of the radix-32 butterflies serves both passes (not unrolled): the kernel's straight-line code was
  // 230 KB, far beyond the 32 KB instruction cache level, and instruction fetch was its largest stall
  // (profiles/r02_kernels_ncu_full.md)
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    pk::Fft<32>::run(v);
    if (h == 0) {
      tw_apply_y<1>(v, s_tw + lane);
      float2* tp = tile + lane;
#pragma unroll
      for (int q = 0; q < 32; ++q) tp[q * kTileRowY] = v[q];
      __syncwarp();
      const float4* rp = reinterpret_cast<const float4*>(tile + lane * kTileRowY);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 f = rp[i];
        v[2 * i] = make_float2(f.x, f.y);
        v[2 * i + 1] = make_float2(f.z, f.w);
      }
      __syncwarp();
    }
  }
}

// Stages one frame: float64 pre-processing exactly as the reference, power-of-two scaling, FP32 copy in pbuf (padded
// natural order), E(tau) = sum_{j<512} p[j+tau]^2 for tau < 512 in ebuf (swizzled), returns E(0) and the frame's energy.
// `live` = the frame exists; `finite` comes back false if a sample is NaN / Inf.
__device__ __forceinline__ void stage_frame(const double* __restrict__ x, int64_t g0, int64_t limit, bool live, double alpha,
                                            const double* __restrict__ s_hann, double* __restrict__ stage,
                                            float* __restrict__ pbuf, float* __restrict__ ebuf, int lane, float* e0,
                                            float* etot, bool* finite) {
  // raw samples g0 - 2 .. g0 + 1023, coalesced, into the staging buffer (x[-1] = x[-2] = 0: pre_emphasis.go:135-155)
  if (live && g0 >= 2 && g0 + kN <= limit) {  // the whole window lies inside the stream (all but the first and last frames)
    const double* __restrict__ xs = x + (g0 - 2 + lane);
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) stage[pidx(lane + 32 * jj)] = __ldg(xs + 32 * jj);
    if (lane < 2) stage[pidx(lane + kN)] = __ldg(xs + kN);
  } else {
#pragma unroll 1
    for (int jj = 0; jj < 33; ++jj) {
      const int e = lane + 32 * jj;
      if (e < kN + 2) {
        const int64_t g = g0 - 2 + e;
        stage[pidx(e)] = (live && g >= 0 && g < limit) ? __ldg(x + g) : 0.0;
      }
    }
  }
  __syncwarp();
  // lane l owns samples i = 32 l .. 32 l + 31 (window of 34 raw values)
  double t[32];
  unsigned hi_max = 0u;
  {
    double xm2 = stage[pidx(32 * lane)], xm1 = stage[pidx(32 * lane + 1)];
    double yprev = xm1 - alpha * xm2;  // y[i-1] of the lane's first sample
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int i = 32 * lane + k;
      const double xi = stage[pidx(i + 2)];
      const double y = xi - alpha * xm1;  // stream-level pre-emphasis (speech.go:161, pre_emphasis.go:184-190)
      double v = y;
      if (i > 0) v = y - 0.97 * yprev;     // the detector's own (pitch_detection.go:299-314): result[0] = signal[0]
      v *= s_hann[pidx(i)];
      t[k] = v;
      hi_max = max(hi_max, (unsigned)__double2hiint(v) & 0x7fffffffu);
      xm1 = xi;
      yprev = y;
    }
  }
  __syncwarp();
  hi_max = __reduce_max_sync(kFullY, hi_max);
  *finite = hi_max < 0x7ff00000u;
  // scale by 2^-e, e = exponent of the largest |p|: the CMNDF is scale invariant, FP32 keeps its full range
  const int ex = (int)(hi_max >> 20) - 1023;
  const double sc = (hi_max >= 0x00100000u && hi_max < 0x7fe00000u) ? __hiloint2double((1023 - ex) << 20, 0) : 1.0;
  float loc[32], run = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float pf = (float)(t[k] * sc);
    pbuf[pidx(32 * lane + k)] = pf;
    loc[k] = run;  // exclusive prefix of p^2 inside the lane
    run = fmaf(pf, pf, run);
  }
  float incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float up = __shfl_up_sync(kFullY, incl, o);
    if (lane >= o) incl += up;
  }
  const float base = incl - run;  // S[32 l]
  *etot = __shfl_sync(kFullY, incl, 31);
  *e0 = __shfl_sync(kFullY, base, 16);  // S[512]
  // E(tau) = S[tau + 512] - S[tau].  The prefix sums S[32 l + k] = base + loc[k] cross the warp through the staging buffer
  // (its float64 samples are in registers by now; rows of 36 floats: the 128-bit stores and loads below are conflict free),
  // then lane l forms the sixteen lags 16 l .. 16 l + 15 it will also read in pick32: 8 + 8 + 4 vector accesses per frame
  // where one shuffle and one predicated store per lag (32 + 32, half the lanes idle) were needed before.
  float* sbuf = reinterpret_cast<float*>(stage);
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4)
    *reinterpret_cast<float4*>(sbuf + lane * 36 + 4 * c4) =
        make_float4(base + loc[4 * c4], base + loc[4 * c4 + 1], base + loc[4 * c4 + 2], base + loc[4 * c4 + 3]);
  __syncwarp();
  const float* lo = sbuf + (lane >> 1) * 36 + (lane & 1) * 16;  // S[16 l ..]: row (16 l) / 32, column (16 l) % 32
  const float* hi = lo + 16 * 36;                               // S[16 l + 512 ..]
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    const float4 a = *reinterpret_cast<const float4*>(lo + 4 * c4), b = *reinterpret_cast<const float4*>(hi + 4 * c4);
    *reinterpret_cast<float4*>(ebuf + spos(16 * lane + 4 * c4)) = make_float4(b.x - a.x, b.y - a.y, b.z - a.z, b.w - a.w);
  }
  __syncwarp();
}

// conj(U) P of one frame from Z = FFT(u + i v): returned for the lane's bins k = lane + 32 k2, k2 < 16 (q[k2]) and,
// lane 0 only, for k = 512 (*nyq).  Unscaled (x 4).
__device__ __forceinline__ void spectrum_product(const float2 (&z)[32], int lane, float2 (&q)[16], float2* nyq) {
  const int src = (32 - lane) & 31;
  const bool lane0 = lane == 0;
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    const float2 mine = lane0 ? z[(32 - k2) & 31] : z[31 - k2];
    const float2 zn = make_float2(__shfl_sync(kFullY, mine.x, src), __shfl_sync(kFullY, mine.y, src));
    const float2 zk = z[k2];
    // 2U = zk + conj(zn) = (sr, si);  2V = -i (zk - conj(zn)) = (di, -dr);  A = U, B = U + (-1)^k V
    const float sr = zk.x + zn.x, si = zk.y - zn.y, dr = zk.x - zn.x, di = zk.y + zn.y;
    const float cr = fmaf(sr, di, -(si * dr)), ci = -fmaf(sr, dr, si * di);  // conj(2U) 2V
    const float uu = fmaf(sr, sr, si * si);
    // k = lane + 32 k2 has the parity of the lane
    q[k2] = (lane & 1) ? make_float2(uu - cr, -ci) : make_float2(uu + cr, ci);
  }
  {  // k = 512 (even): zn = zk = Z[512]
    const float2 zk = z[16];
    const float sr = 2.f * zk.x, di = 2.f * zk.y;
    *nyq = make_float2(fmaf(sr, sr, 0.f) + sr * di, 0.f);  // si = dr = 0: uu + cr, ci = 0
  }
}

struct PickOut {
  float pitch, conf;
  int flag;  // 0 = every decision is outside the FP32 error bound; else reason bits: 1 unsure comparison, 2 ill-conditioned
             // refinement, 4 frequency gate within rounding
};

// CMNDF + first dip + parabolic refinement of one frame in FP32 with the decision margins (lane l: lags 16 l .. 16 l + 15).
__device__ __forceinline__ PickOut pick32(const float* __restrict__ rb, int comp, const float* __restrict__ ebuf,
                                          const float* __restrict__ pb, float e0, float etot, float rscale, int sr,
                                          int lane) {
  float d[16], loc[16], run = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 ev = *reinterpret_cast<const float4*>(ebuf + spos(16 * lane + 4 * c));
    const float4 r01 = *reinterpret_cast<const float4*>(rb + 2 * ppos(16 * lane + 4 * c));
    const float4 r23 = *reinterpret_cast<const float4*>(rb + 2 * ppos(16 * lane + 4 * c + 2));
    const float rr[4] = {comp ? r01.y : r01.x, comp ? r01.w : r01.z, comp ? r23.y : r23.x, comp ? r23.w : r23.z};
    const float ee[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = 4 * c + u;
      const float dv = (e0 + ee[u]) - 2.f * (rr[u] * rscale);
      d[k] = dv > 0.f ? dv : 0.f;
      run += (lane | k) ? d[k] : 0.f;  // tau >= 1
      loc[k] = run;
    }
  }
  float incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float up = __shfl_up_sync(kFullY, incl, o);
    if (lane >= o) incl += up;
  }
  const float base = incl - run;
  const float bd = kKappa * (etot + e0);  // bound on |d32 - d|
  // Screen without divisions: cm[tau] < 0.15 + eb[tau] (the only way a lag can be chosen OR make the choice unsure) needs
  // d tau < 0.15 cum + bd tau (1 + cm) with cm < 0.15 + eb; 1.25 bd tau covers it for eb < 0.1, and eb >= 0.1 means
  // bd tau >= 0.087 cum, which the second test catches.  Frames without any such lag (noise, silence with energy,
  // most unvoiced audio) have no pitch for sure: the CMNDF itself is never formed for them.  Negated comparisons: NaN passes.
  {
    bool cand = false;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int tau = 16 * lane + k;
      const float cum = base + loc[k], ft = (float)tau;
      const bool in = tau >= 1 && tau + 1 < kHalf;
      cand = cand || (in && (!(d[k] * ft >= fmaf(kThresh, cum, 1.25f * bd * ft)) || !(bd * ft < 0.08f * cum)));
    }
    if (!__any_sync(kFullY, cand)) {
      PickOut none;
      none.pitch = 0.f;
      none.conf = 0.f;
      none.flag = 0;
      return none;
    }
  }
  float cm[17], eb[17];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int tau = 16 * lane + k;
    const float cum = base + loc[k];
    const float inv = __fdividef((float)tau, cum);     // tau / cum (cum = 0: inf / nan, handled by the flag below)
    cm[k] = tau == 0 ? 1.f : d[k] * inv;               // cmndf[tau] = d / (cum / tau)  (pitch_detection.go:372-376)
    eb[k] = tau == 0 ? 0.f : bd * inv * (1.f + cm[k]);  // |cm32 - cm| <= (|dd| + cm |dcum| / tau) tau / cum
  }
  cm[16] = __shfl_down_sync(kFullY, cm[0], 1);
  eb[16] = __shfl_down_sync(kFullY, eb[0], 1);
  // first tau >= 1 with cm < 0.15, tau + 1 < 512, cm[tau] < cm[tau+1]  (pitch_detection.go:378-388)
  int first = kHalf, unsure = kHalf;
#pragma unroll
  for (int k = 15; k >= 0; --k) {
    const int tau = 16 * lane + k;
    const bool in = tau >= 1 && tau + 1 < kHalf;
    if (in && cm[k] < kThresh && cm[k] < cm[k + 1]) first = tau;
    // a comparison inside its error bound (or not a number while the frame has energy): the reference may decide otherwise
    const float m1 = fabsf(cm[k] - kThresh), m2 = fabsf(cm[k] - cm[k + 1]);
    const bool close1 = !(m1 > eb[k]);
    const bool close2 = cm[k] < kThresh + eb[k] && !(m2 > eb[k] + eb[k + 1]);
    if (in && (close1 || close2)) unsure = tau;
  }
  const int mt = __reduce_min_sync(kFullY, first);
  const int mu = __reduce_min_sync(kFullY, unsure);
  PickOut out;
  out.pitch = 0.f;
  out.conf = 0.f;
  out.flag = (mu < kHalf && mu <= mt && etot > 0.f) ? 1 : 0;  // an unsure comparison at or before the chosen dip
  if (mt < kHalf) {
    // The three CMNDF values of the refinement are re-taken from DIRECT differences d = sum (p[j] - p[j+tau])^2 in FP32:
    // their error is relative to d itself (~1e-6), not to the frame energy as on the transform route, which is what the
    // parabola through a deep dip needs (its curvature falls with the square of the period).  The running sum cum[mt] is
    // the transform route's (a factor common to the three values; it cancels in -b / 2a).
    const int owner = mt >> 4, k = mt & 15;
    float cum_sel = 0.f;
#pragma unroll
    for (int qd = 0; qd < 16; ++qd)
      if (qd == k) cum_sel = base + loc[qd];
    const float cum0 = __shfl_sync(kFullY, cum_sel, owner);
    float dd[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int j = lane + 32 * i;
      const float pj = pb[pidx(j)];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float df = pj - pb[pidx(j + mt - 1 + c)];  // mt >= 1
        dd[c] = fmaf(df, df, dd[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) dd[c] += __shfl_xor_sync(kFullY, dd[c], o);
    const float cm1 = cum0 - dd[1], cp1 = cum0 + dd[2];
    const float y1 = mt == 1 ? 1.f : dd[0] * (float)(mt - 1) / cm1;  // cmndf[0] = 1
    const float y2 = dd[1] * (float)mt / cum0, y3 = dd[2] * (float)(mt + 1) / cp1;
    float period = (float)mt;
    bool ill = false;
    if (mt > 0 && mt < kHalf - 1) {  // parabolicInterpolation (pitch_detection.go:743-764)
      const float a = (y1 - 2.f * y2 + y3) * 0.5f, b = (y3 - y1) * 0.5f;
      if (a != 0.f) period += -b / (2.f * a);
      ill = !(fabsf(a) > 1e-3f * (fabsf(y1) + 2.f * fabsf(y2) + fabsf(y3)));  // the curvature itself cancels: leave it to float64
    }
    const float freq = (float)sr / period;
    if (freq >= 80.f && freq <= 1000.f) {
      out.pitch = freq;
      out.conf = 1.f - y2;
    }
    const bool gate = fabsf(freq - 80.f) < 0.08f || fabsf(freq - 1000.f) < 1.f;  // the 80..1000 Hz gate within rounding
    out.flag |= (ill ? 2 : 0) | (gate ? 4 : 0);
  }
  return out;
}

__global__ void __launch_bounds__(kYW * 32, 1)
    yin32_kernel(const double* __restrict__ pcm, int64_t stride, double alpha, int sr, int64_t Tp, int pairs_per_stream,
                 int64_t total_pairs, const double* __restrict__ hann, double* __restrict__ raw, int64_t raw_stride,
                 int* __restrict__ lists, int64_t list_stride) {
  extern __shared__ __align__(128) unsigned char smem[];
  const Y32Smem L = y32_layout();
  float2* s_tw = reinterpret_cast<float2*>(smem + L.tw);
  double* s_hann = reinterpret_cast<double*>(smem + L.hann);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wb = smem + L.warp0 + (size_t)warp * L.per_warp;
  float2* tile = reinterpret_cast<float2*>(wb + L.w_tile);
  double* stage = reinterpret_cast<double*>(wb + L.w_tile);
  float* pbuf = reinterpret_cast<float*>(wb + L.w_p);  // [2][kN + 32]
  float* rbuf = reinterpret_cast<float*>(wb + L.w_tile);  // r of both frames, float2 per lag, swizzled by ppos
  float* ebuf = reinterpret_cast<float*>(wb + L.w_e);

  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int k1 = i / 32, l = i % 32;
    double dsn, dcs;
    sincospi(-2.0 * (double)((k1 * l) % kN) / (double)kN, &dsn, &dcs);
    s_tw[i] = make_float2((float)dcs, (float)dsn);
  }
  for (int i = threadIdx.x; i < kN; i += blockDim.x) s_hann[pidx(i)] = hann[i];
  __syncthreads();

  const int64_t limit = (Tp + 1) * kHop;  // samples [0, limit) belong to the Tp frames
  const float rscale = 1.0f / 4096.0f;    // 1/4 (Hermitian split of both factors) * 1/1024 (inverse transform)
  for (int64_t base = (int64_t)blockIdx.x * kYW; base < total_pairs; base += (int64_t)gridDim.x * kYW) {
    const int64_t pr = base + warp;
    if (pr >= total_pairs) continue;
    const int s = (int)(pr / pairs_per_stream);
    const int64_t fa = (pr % pairs_per_stream) * 2, fb = fa + 1;
    const double* __restrict__ x = pcm + (int64_t)s * stride;
    float2 qa[16], qb[16], nyqa, nyqb;
    float e0a = 0.f, e0b = 0.f, etota = 0.f, etotb = 0.f;
    bool fina = true, finb = true;
    // ---- three passes through ONE instance of the transform (the loop body must stay near the 32 KB instruction cache
    //      level, profiles/r02_stft_v5_ncu.md): frame a, frame b (stage, forward transform, conj(U) P), then the inverse
    //      of Q = Q_a + i Q_b.  Frame b's pass leaves Q, conjugated for the inverse-by-forward transform, in z itself, so
    //      nothing but qa is carried from one pass to the next.
    float2 z[32];
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      if (c < 2) {
        float* pb = pbuf + c * (kN + 32);
        float e0c, etc;
        bool fc;
        stage_frame(x, (fa + c) * kHop, limit, c == 0 || fb < Tp, alpha, s_hann, stage, pb, ebuf + c * (kHalf + 4), lane, &e0c,
                    &etc, &fc);
        if (c == 0)
          e0a = e0c, etota = etc, fina = fc;
        else
          e0b = e0c, etotb = etc, finb = fc;
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = make_float2(pb[pidx(lane + 32 * j)], pb[pidx(lane + 32 * j + kHalf)]);
#pragma unroll
        for (int j = 16; j < 32; ++j) z[j] = make_float2(0.f, 0.f);
        __syncwarp();
      }
      warp_fft1024(z, tile, s_tw, lane);
      if (c < 2) {
        spectrum_product(z, lane, qb, &nyqb);
        if (c == 0) {
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2) qa[k2] = qb[k2];
          nyqa = nyqb;
        } else {
          const int src = (32 - lane) & 31;
          const bool lane0 = lane == 0;
          float2 g[16];  // Q[1024 - k] = conj(Q_a[k]) + i conj(Q_b[k]), destined for the partner lane's upper registers
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2) {
            const float2 a = qa[k2], b = qb[k2];
            z[k2] = make_float2(a.x - b.y, -(a.y + b.x));      // conj(Q_a + i Q_b)
            g[k2] = make_float2(a.x + b.y, -(b.x - a.y));      // conj(conj Q_a + i conj Q_b)
          }
          const float2 wn = make_float2(nyqa.x - nyqb.y, -(nyqa.y + nyqb.x));
#pragma unroll
          for (int r = 16; r < 32; ++r) {
            // lanes != 0: register r <- partner's g[31 - r]; lane 0: r = 16 <- the Nyquist bin, r > 16 <- own g[32 - r]
            const float2 give = lane0 ? (r == 16 ? wn : g[32 - r]) : g[31 - r];
            z[r] = make_float2(__shfl_sync(kFullY, give.x, src), __shfl_sync(kFullY, give.y, src));
          }
        }
      } else {
        __syncwarp();  // every lane has its row of the tile in registers: the tile becomes r
        // IFFT(Q) = conj(FFT(conj Q)) / N: r_a = Re, r_b = -Im; lags tau = lane + 32 k2 < 512
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2)
          *reinterpret_cast<float2*>(rbuf + 2 * ppos(lane + 32 * k2)) = make_float2(z[k2].x, -z[k2].y);
      }
    }
    __syncwarp();
    // ---- CMNDF, first dip, refinement; borderline frames go to the exact list ---------------------------------
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int64_t f = fa + c;
      const bool fin_c = c ? finb : fina;
      const PickOut o = pick32(rbuf, c, ebuf + c * (kHalf + 4), pbuf + c * (kN + 32), c ? e0b : e0a, c ? etotb : etota, rscale,
                               sr, lane);
      if (lane == 0 && f < Tp) {
        double* r = raw + (int64_t)s * raw_stride;
        r[f] = (double)o.pitch;
        r[Tp + f] = (double)o.conf;
        if (o.flag || !fin_c) {
          int* lst = lists + (int64_t)s * list_stride;
          const int pos = atomicAdd(lst, 1);
          lst[1 + pos] = (int)f | ((o.flag | (fin_c ? 0 : 8)) << 27);  // frame (< 2^27) + reason bits (diagnostic)
        }
      }
    }
    __syncwarp();
  }
}

// Exact float64 re-evaluation of the listed frames, in the reference's summation order (bit identical results).
constexpr int kExThreads = 512;
__global__ void __launch_bounds__(kExThreads)
    yin_exact_kernel(const double* __restrict__ pcm, int64_t stride, double alpha, int sr, int64_t Tp,
                     const double* __restrict__ hann, double* __restrict__ raw, int64_t raw_stride,
                     const int* __restrict__ lists, int64_t list_stride) {
  __shared__ double p[kN];
  __shared__ double dsh[kHalf];
  __shared__ double cmsh[kHalf];
  __shared__ int s_first;
  const int s = blockIdx.y, t = threadIdx.x;
  const int* lst = lists + (int64_t)s * list_stride;
  const int count = lst[0];
  const double* __restrict__ x = pcm + (int64_t)s * stride;
  for (int idx = blockIdx.x; idx < count; idx += gridDim.x) {
    const int64_t f = lst[1 + idx] & 0x7ffffff;
    const int64_t g0 = f * kHop;
    for (int i = t; i < kN; i += kExThreads) {
      const int64_t g = g0 + i;
      const double xm1 = g > 0 ? x[g - 1] : 0.0, xm2 = g > 1 ? x[g - 2] : 0.0;
      const double y = x[g] - alpha * xm1;
      double v = y;
      if (i > 0) v = y - 0.97 * (xm1 - alpha * xm2);
      p[i] = v * hann[i];
    }
    if (t == 0) s_first = kHalf;
    __syncthreads();
    {  // d[tau], sequential in j (pitch_detection.go:355-363)
      double sum = 0.0;
      for (int j = 0; j < kHalf; ++j) {
        const double delta = p[j] - p[j + t];
        sum += delta * delta;
      }
      dsh[t] = sum;
    }
    __syncthreads();
    if (t == 0) {  // running sum, sequential in tau (:369-376)
      double rs = 0.0;
      cmsh[0] = 1.0;
      for (int tau = 1; tau < kHalf; ++tau) {
        rs += dsh[tau];
        cmsh[tau] = dsh[tau] / (rs / (double)tau);
      }
    }
    __syncthreads();
    if (t >= 1 && t + 1 < kHalf && cmsh[t] < 0.15 && cmsh[t] < cmsh[t + 1]) atomicMin(&s_first, t);
    __syncthreads();
    if (t == 0) {
      const int mt = s_first;
      double pitch = 0.0, conf = 0.0;
      if (mt < kHalf && mt > 0) {
        double period = (double)mt;
        if (!(mt <= 0 || mt >= kHalf - 1)) {  // parabolicInterpolation :743-764
          const double y1 = cmsh[mt - 1], y2 = cmsh[mt], y3 = cmsh[mt + 1];
          const double a = (y1 - 2 * y2 + y3) / 2, b = (y3 - y1) / 2;
          if (a != 0) period = (double)mt + (-b / (2 * a));
        }
        const double freq = (double)sr / period;
        if (freq >= 80.0 && freq <= 1000.0) {
          pitch = freq;
          conf = 1.0 - cmsh[mt];
        }
      }
      double* r = raw + (int64_t)s * raw_stride;
      r[f] = pitch;
      r[Tp + f] = conf;
    }
    __syncthreads();
  }
}

}  // namespace

// raw pitch / confidence of every frame into scratch (per stream 2 Tp doubles); lists: per stream 1 + Tp ints (count,
// then the frames re-evaluated exactly), list_stride ints apart, counts zeroed by the caller.
int launch_yin32(const double* pcm, int64_t stride, int n_streams, double alpha, int sr, int64_t Tp, const double* hann_dev,
                 double* scratch, int64_t scratch_stride, int* lists, int64_t list_stride, cudaStream_t st) {
  const int pps = (int)((Tp + 1) / 2);
  const int64_t total = (int64_t)pps * n_streams;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const Y32Smem L = y32_layout();
  SONAR_CUDA(cudaFuncSetAttribute(yin32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  // not persistent for the whole launch: a CTA takes a bounded share and retires, so the latency-bound kernels of the
  // alignment branch (other stream, higher priority) get an SM's shared memory within ~100 us
  const int64_t per_cta = (int64_t)kYW * 32;
  const unsigned ctas = (unsigned)std::max<int64_t>(std::min<int64_t>((total + kYW - 1) / kYW, (int64_t)sms),
                                                    (total + per_cta - 1) / per_cta);
  prof_begin("yin_frame_kernel", st);
  yin32_kernel<<<ctas, kYW * 32, L.total, st>>>(pcm, stride, alpha, sr, Tp, pps, total, hann_dev, scratch, scratch_stride,
                                                lists, list_stride);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  static const bool debug = std::getenv("SONAR_YIN_DEBUG") != nullptr;  // diagnostic: how many frames went to the exact list, why
  if (debug) {
    cudaStreamSynchronize(st);
    std::vector<int> h((size_t)(1 + Tp));
    long tot = 0, why[16] = {0};
    for (int s = 0; s < n_streams; ++s) {
      cudaMemcpy(h.data(), lists + (int64_t)s * list_stride, sizeof(int) * (size_t)(1 + Tp), cudaMemcpyDeviceToHost);
      tot += h[0];
      for (int i = 0; i < h[0]; ++i) why[(h[1 + i] >> 27) & 15]++;
    }
    fprintf(stderr, "[yin32] %ld of %lld frames re-evaluated exactly; by reason bits (1 unsure, 2 ill, 4 gate, 8 non-finite):",
            tot, (long long)Tp * n_streams);
    for (int i = 1; i < 16; ++i)
      if (why[i]) fprintf(stderr, " %d:%ld", i, why[i]);
    fprintf(stderr, "\n");
  }
  static const bool noexact = std::getenv("SONAR_YIN_NOEXACT") != nullptr;  // diagnostic: keep the FP32 values of listed frames
  if (noexact) return SONAR_OK;
  prof_begin("yin_exact_kernel", st);
  yin_exact_kernel<<<dim3(8, (unsigned)n_streams), kExThreads, 0, st>>>(pcm, stride, alpha, sr, Tp, hann_dev, scratch,
                                                                        scratch_stride, lists, list_stride);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

}  // namespace sonar
