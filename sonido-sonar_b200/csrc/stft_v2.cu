// Fused framed-STFT + MFCC + spectral-descriptor kernel, second generation (N = 1024 only; FP32 arithmetic,
// f64 I/O).  Same outputs as stft_features.cu (which stays the general path for the other window sizes,
// unaligned PCM and the materialising spectrum mode), replacing
//   analyzers.ComputeSTFTWithWindow        fingerprint/analyzers/spectral.go:385-545
//   spectral.MFCC.Compute                  algorithms/spectral/mfcc.go:113-164
//   MelScale.ApplyFilterBank               algorithms/spectral/mel_scale.go:89-105
//   centroid/rolloff/bandwidth/flatness/crest/slope/flux   algorithms/spectral/spectral_*.go
//   low/high band energy ratios            fingerprint/extractors/speech.go:436-456
//
// The first-generation kernel keeps a 32-point FFT per lane in registers (255 registers, 8 warps per SM)
// and is bound by instruction latency.  Here ONE WARP OWNS ONE FRAME and every lane holds 16 of the 512
// complex points (z[n] = x[2n] + i x[2n+1]), so the kernel fits 12 warps per SM:
//   * the warp walks consecutive frames of one stream; the 256 new samples of a frame are read from HBM
//     once (coalesced 16-byte loads), converted to FP32 once and kept in a 1024-sample ring in shared memory
//     (every sample is used by four frames);
//   * 512 = 8 x 8 x 8: three radix-8 passes in registers, two butterflies per lane per pass, exchanged
//     through a padded per-warp tile (every access below is bank-conflict free, 16-byte where possible):
//       pass 1   n = 64 a + n',  lane -> n' = lane, lane + 32          twiddle W_512^(n' ka)
//       pass 2a  n' = 8 b + c,   lane -> (ka, c = 2 (lane&3) + {0,1})  twiddle W_64^(c kb)
//       pass 2b                  lane -> (ka, kb = (lane&3) + {0,4})   -> Z[ka + 8 kb + 64 kc]
//   * split pass: lane -> bins k = lane + 32 i and their mirrors 512 - k (both come out of the same pair
//     of Z values and share the twiddle product); |X| goes to a padded row in shared memory, the centroid
//     sums are taken on the way;
//   * phase B: every lane scans 16 CONTIGUOUS bins with 16-byte loads (magnitudes, previous frame's
//     magnitudes, per-bin mel weights and slope abscissae from tables built once per CTA); mel partial sums
//     go to lane-private slots and are combined per filter afterwards; ln + DCT-II + lifter per frame;
//   * the FP64 finishing arithmetic of the nine scalar descriptors is deferred: the raw FP32 sums of the
//     frames of a run are parked in shared memory and finished 32 frames at a time, one frame per lane.
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "common.h"
#include "fft_regs.cuh"

namespace sonar {
namespace {

constexpr int kW = 12;                 // warps per CTA (one CTA per SM)
constexpr int kRunFrames = 32;         // frames per run, the first is the flux warm-up
constexpr int kRunOut = kRunFrames - 1;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kN = 1024, kM = 512, kB = 513;
constexpr int kT1Row = 72;             // pass-1 tile: row ka, 64 columns + 8 pad (float2)
constexpr int kT2Ka = 88, kT2Kb = 10;  // pass-2 tile: [ka][kb][c] with padded strides (float2)
constexpr int kTile = 8 * kT2Ka;       // float2 per warp (also holds Z in split order: 640)
constexpr int kRow = 644;              // padded bin row, floats: bin k at k + 4 (k >> 4)
constexpr int kSlots = 24;             // lane-private mel slots (alias the tile during phase B)
constexpr int kRaw = 16;               // raw sums parked per frame
constexpr int kMaxContrib = 12;        // lanes that may hold a part of one mel filter

__device__ __forceinline__ int bpos(int k) { return k + 4 * (k >> 4); }

struct V2Smem {
  size_t tw1, tw2, xtab, wlo, whi, fmask, moff, dct, lift, r0, warp0, per_warp, total;
  size_t w_ring, w_tile, w_mag, w_raw, w_macc;
};

__host__ __device__ inline V2Smem v2_layout(int n_mel, int n_mfcc) {
  V2Smem L;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += (bytes + 15) & ~(size_t)15;
    return r;
  };
  L.tw1 = take(sizeof(float2) * 7 * 64);
  L.tw2 = take(sizeof(float2) * 7 * 8);
  L.xtab = take(sizeof(float) * kRow);
  L.wlo = take(sizeof(float) * kRow);
  L.whi = take(sizeof(float) * kRow);
  L.fmask = take(sizeof(unsigned) * 32);
  L.moff = take(sizeof(unsigned short) * kMaxContrib * kMaxMel);
  L.dct = take(sizeof(float) * (size_t)n_mfcc * (n_mel | 1));
  L.lift = take(sizeof(float) * n_mfcc);
  L.r0 = take(sizeof(int) * 33);
  o = (o + 127) & ~(size_t)127;
  L.warp0 = o;
  size_t w = 0;
  auto wtake = [&](size_t bytes) {
    size_t r = w;
    w += (bytes + 127) & ~(size_t)127;
    return r;
  };
  L.w_ring = wtake(sizeof(float) * kN);
  L.w_tile = wtake(sizeof(float2) * kTile);
  L.w_mag = wtake(sizeof(float) * 2 * kRow);
  L.w_raw = wtake(sizeof(float) * kRaw * kRunFrames);
  L.w_macc = wtake(sizeof(float) * (kMaxMel + 4));
  L.per_warp = w;
  L.total = o + w * kW;
  return L;
}

template <int K>
__device__ __forceinline__ void tw_apply8(float2 (&v)[8], const float2* __restrict__ tw, int stride) {
  if constexpr (K < 8) {
    v[K] = cmul(v[K], tw[(K - 1) * stride]);
    tw_apply8<K + 1>(v, tw, stride);
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

struct BinAcc {
  float seg, mx, sl, sxy, fl, bw, mlo, mhi, pend;
  int ninv;
  float* pp;  // next lane-private mel slot
};

__device__ __forceinline__ float sqrt_fast(float x) {  // MUFU; sqrt(0) = 0
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_fast(float x) {  // MUFU.LG2; the callers only pass normal numbers
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One bin of phase B.  `flush`: the bin opens a new mel region, i.e. the falling part of the filter being left
// joins its pending rising part in the next private slot (regions are at least one bin wide: eligibility).
__device__ __forceinline__ void bin_step(BinAcc& s, bool flush, float dk, float m, float mp, float xv, float wl,
                                         float wh, bool count_inv) {
  if (flush) {
    *s.pp = s.pend + s.mlo;
    s.pend = s.mhi;
    s.mlo = 0.f;
    s.mhi = 0.f;
    s.pp += 32;
  }
  const float p = m * m;
  s.mlo = fmaf(p, wl, s.mlo);
  s.mhi = fmaf(p, wh, s.mhi);
  s.seg += p;
  s.mx = fmaxf(s.mx, m);
  s.bw = fmaf(dk * dk, m, s.bw);
  const bool valid = m > 1e-10f;
  const float l2 = valid ? lg2_fast(m) : 0.f;
  s.sl += l2;
  s.sxy = fmaf(xv, l2, s.sxy);  // xtab[0] == 0: bin 0 never enters the regression
  if (!valid && count_inv) ++s.ninv;
  const float d = fmaxf(m - mp, 0.f);
  s.fl = fmaf(d, d, s.fl);
}

template <int NP>  // NP = hop / 64: new sample pairs per lane per frame; NP = 0: any even hop <= 1024 (no register prefetch)
__global__ void __launch_bounds__(kW * 32, 1) stft_v2_kernel(const StftArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const V2Smem L = v2_layout(a.n_mel, a.n_mfcc);
  float2* s_tw1 = reinterpret_cast<float2*>(smem + L.tw1);
  float2* s_tw2 = reinterpret_cast<float2*>(smem + L.tw2);
  float* s_xtab = reinterpret_cast<float*>(smem + L.xtab);
  float* s_wlo = reinterpret_cast<float*>(smem + L.wlo);
  float* s_whi = reinterpret_cast<float*>(smem + L.whi);
  unsigned* s_fmask = reinterpret_cast<unsigned*>(smem + L.fmask);
  unsigned short* s_moff = reinterpret_cast<unsigned short*>(smem + L.moff);
  float* s_dct = reinterpret_cast<float*>(smem + L.dct);
  float* s_lift = reinterpret_cast<float*>(smem + L.lift);
  int* s_r0 = reinterpret_cast<int*>(smem + L.r0);
  __shared__ int s_ncontrib;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wb = smem + L.warp0 + (size_t)warp * L.per_warp;
  float* wbf = reinterpret_cast<float*>(wb);
  float2* ring2 = reinterpret_cast<float2*>(wb + L.w_ring);
  float2* tile = reinterpret_cast<float2*>(wb + L.w_tile);
  float* priv = reinterpret_cast<float*>(tile);  // lane-private mel slots [slot][lane], phase B only
  float* mag = reinterpret_cast<float*>(wb + L.w_mag);
  float* rawsum = reinterpret_cast<float*>(wb + L.w_raw);
  float* macc = reinterpret_cast<float*>(wb + L.w_macc);
  constexpr int kZeroSlot = kMaxMel + 2;  // macc[kZeroSlot] stays 0: padding target of the combine table

  // ---- tables, once per CTA ------------------------------------------------------------------------
  for (int i = threadIdx.x; i < 7 * 64; i += blockDim.x) {
    const int ka = i / 64 + 1, np = i % 64;
    double dsn, dcs;
    sincospi(-(double)((ka * np) % 512) / 256.0, &dsn, &dcs);
    s_tw1[i] = make_float2((float)dcs, (float)dsn);
  }
  for (int i = threadIdx.x; i < 7 * 8; i += blockDim.x) {
    const int kb = i / 8 + 1, c = i % 8;
    double dsn, dcs;
    sincospi(-(double)(kb * c) / 32.0, &dsn, &dcs);
    s_tw2[i] = make_float2((float)dcs, (float)dsn);
  }
  for (int k = threadIdx.x; k < kRow; k += blockDim.x) {
    s_xtab[k] = 0.f;
    s_wlo[k] = 0.f;
    s_whi[k] = 0.f;
  }
  if (threadIdx.x < 32) s_fmask[threadIdx.x] = 0u;
  if (threadIdx.x == 0) s_ncontrib = 0;
  if (lane == 0) macc[kZeroSlot] = 0.f;
  __syncthreads();
  for (int k = threadIdx.x; k < kB; k += blockDim.x) {
    s_xtab[bpos(k)] = a.xtab[k];
    int r = 0;
    while (k >= a.regions[r].next_b) ++r;
    const MelRegion reg = a.regions[r];
    const float kf = (float)k;
    s_wlo[bpos(k)] = (reg.bhi - kf) * reg.inv_f;
    s_whi[bpos(k)] = (kf - reg.blo) * reg.inv_r;
    int rp = 0;
    if (k > 0)
      while (k - 1 >= a.regions[rp].next_b) ++rp;
    // bit j of a lane's mask: bin 16 lane + j opens a new region (a lane's first bin starts inside its region,
    // nothing to close; bin 512 is lane 31's seventeenth)
    if (r != rp && ((k & 15) || k == kM)) atomicOr(&s_fmask[k == kM ? 31 : (k >> 4)], 1u << (k == kM ? 16 : (k & 15)));
    if ((k & 15) == 0 && k < kM) s_r0[k >> 4] = r;
  }
  {
    const int nmp = a.n_mel | 1;
    for (int i = threadIdx.x; i < a.n_mfcc * a.n_mel; i += blockDim.x)
      s_dct[(i / a.n_mel) * nmp + (i % a.n_mel)] = a.dct[i];
    for (int i = threadIdx.x; i < a.n_mfcc; i += blockDim.x) s_lift[i] = a.lift[i];
  }
  __syncthreads();
  // combine table: the private slots (as float offsets from the warp's base) that hold a part of filter f
  // (absolute slot f + 1; lane j's slot q is r0[j] - 1 + q), padded with the zero slot to a uniform count
  const unsigned short zero_off = (unsigned short)((macc + kZeroSlot) - wbf);
  const unsigned short priv_off = (unsigned short)(priv - wbf);
  for (int f = threadIdx.x; f < kMaxMel; f += blockDim.x) {
    int cnt = 0;
    if (f < a.n_mel) {
      for (int j = 0; j < 32; ++j) {
        int rl = 0;
        const int kl = (j == 31) ? kB - 1 : 16 * j + 15;
        while (kl >= a.regions[rl].next_b) ++rl;
        const int first = s_r0[j] - 1, last = rl;  // slots first .. last are written by lane j
        if (f + 1 >= first && f + 1 <= last && cnt < kMaxContrib)
          s_moff[(cnt++) * kMaxMel + f] = (unsigned short)(priv_off + (f + 1 - first) * 32 + j);
      }
      atomicMax(&s_ncontrib, cnt);
    }
    for (int i = cnt; i < kMaxContrib; ++i) s_moff[i * kMaxMel + f] = zero_off;
  }
  __syncthreads();
  const int ncontrib = s_ncontrib;

  // ---- per-lane constants ----------------------------------------------------------------------------
  float2 wv[16];  // window (with the split pass's 1/2) for z[64 a + lane + 32 h] at wv[8 h + a]
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int x = 0; x < 8; ++x) wv[8 * h + x] = __ldg(a.win2 + 64 * x + lane + 32 * h);
  float2 wN[8];  // W_1024^(lane + 32 i)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    double dsn, dcs;
    sincospi(-(double)(lane + 32 * i) / 512.0, &dsn, &dcs);
    wN[i] = make_float2((float)dcs, (float)dsn);
  }
  const int ka2 = lane >> 2, q4 = lane & 3;
  const int H = NP ? 64 * NP : a.hop;
  constexpr int NPR = NP ? NP : 1;
  const int64_t T = a.T;
  const unsigned fmask = s_fmask[lane];
  const int lo_k = lane + 4 * (lane >> 4);           // bpos(lane + 32 i) = lo_k + 40 i
  const int hi_k = lane + 4 * ((lane + 15) >> 4);    // bpos(512 - lane - 32 i) = 640 - 40 i - hi_k
  const int pb = 20 * lane;                          // bpos(16 lane)

  for (int64_t run = (int64_t)blockIdx.x * kW + warp; run < a.total_runs; run += (int64_t)gridDim.x * kW) {
    const int s = (int)(run / a.runs_per_stream);
    const int64_t t0 = (run % a.runs_per_stream) * (int64_t)kRunOut;
    const int64_t tend = (t0 + kRunOut < T) ? t0 + kRunOut : T;
    const double* __restrict__ x = a.pcm + (int64_t)s * a.stride;
    double* __restrict__ fo = a.feat + (int64_t)s * a.feat_stride;
    const int nfr = (int)(tend - t0) + 1;      // frames t0 - 1 .. tend - 1; the first only warms the flux up
    const int it0 = t0 == 0 ? 1 : 0;           // the stream's first frame has no predecessor (flux[t-1] starts at t = 1)

    // ---- ring: all 1024 samples of the run's first frame -------------------------------------------------
    {
      const int64_t tf = t0 - 1 + it0;
      const double2* __restrict__ src = reinterpret_cast<const double2*>(x + tf * H);
      const int r2 = (int)(((tf * H) >> 1) & (kM - 1));
#pragma unroll 4
      for (int j = 0; j < 16; ++j) {
        const double2 d = __ldg(src + lane + 32 * j);
        ring2[(r2 + lane + 32 * j) & (kM - 1)] = make_float2((float)d.x, (float)d.y);
      }
      __syncwarp();
    }

    for (int it = it0; it < nfr; ++it) {
      const int64_t t = t0 - 1 + it;
      const bool out_ok = it > 0;
      const int base2 = (int)(((t * H) >> 1) & (kM - 1));  // float2 index of the frame's first sample pair
      // ================= pass 1 =================
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float2 v[8];
#pragma unroll
        for (int x8 = 0; x8 < 8; ++x8) {
          const float2 sm2 = ring2[(base2 + 64 * x8 + lane + 32 * h) & (kM - 1)];
          v[x8] = make_float2(sm2.x * wv[8 * h + x8].x, sm2.y * wv[8 * h + x8].y);
        }
        FftReg<8>::run(v);
        tw_apply8<1>(v, s_tw1 + lane + 32 * h, 64);
#pragma unroll
        for (int k8 = 0; k8 < 8; ++k8) tile[k8 * kT1Row + lane + 32 * h] = v[k8];
      }
      // the next frame's H new samples start their trip from HBM now and are parked in the ring at the end
      // of this iteration (they replace the oldest H samples, which pass 1 above was the last to read)
      double2 nx[NPR];
      const bool more = it + 1 < nfr;
      if (NP && more) {
        const double2* __restrict__ src = reinterpret_cast<const double2*>(x + (t + 1) * H + (kN - H));
#pragma unroll
        for (int j = 0; j < NPR; ++j) nx[j] = __ldg(src + lane + 32 * j);
      }
      __syncwarp();
      // ================= pass 2a =================
      {
        float2 v0[8], v1[8];
        const float4* tp = reinterpret_cast<const float4*>(tile + ka2 * kT1Row + 2 * q4);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const float4 f = tp[4 * b];  // T[ka][8 b + c0], T[ka][8 b + c0 + 1]
          v0[b] = make_float2(f.x, f.y);
          v1[b] = make_float2(f.z, f.w);
        }
        FftReg<8>::run(v0);
        FftReg<8>::run(v1);
        tw_apply8<1>(v0, s_tw2 + 2 * q4, 8);
        tw_apply8<1>(v1, s_tw2 + 2 * q4 + 1, 8);
        __syncwarp();  // every lane has read the pass-1 tile
        float4* up = reinterpret_cast<float4*>(tile + ka2 * kT2Ka + 2 * q4);
#pragma unroll
        for (int kb = 0; kb < 8; ++kb) up[(kT2Kb / 2) * kb] = make_float4(v0[kb].x, v0[kb].y, v1[kb].x, v1[kb].y);
      }
      __syncwarp();
      // ================= pass 2b =================
      {
        float2 z0[8], z1[8];  // Z[ka + 8 kb + 64 kc], kb = q4 (z0) and q4 + 4 (z1)
        const float4* u0 = reinterpret_cast<const float4*>(tile + ka2 * kT2Ka + kT2Kb * q4);
        const float4* u1 = reinterpret_cast<const float4*>(tile + ka2 * kT2Ka + kT2Kb * (q4 + 4));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 f = u0[j], g4 = u1[j];
          z0[2 * j] = make_float2(f.x, f.y);
          z0[2 * j + 1] = make_float2(f.z, f.w);
          z1[2 * j] = make_float2(g4.x, g4.y);
          z1[2 * j + 1] = make_float2(g4.z, g4.w);
        }
        FftReg<8>::run(z0);
        FftReg<8>::run(z1);
        __syncwarp();
        // bpos(ka + 8 kb + 64 kc): (ka + 8 q4) >> 4 == q4 >> 1
        float2* zp = tile + ka2 + 8 * q4 + 4 * (q4 >> 1);
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) {
          zp[80 * kc] = z0[kc];       // k + 64 kc
          zp[80 * kc + 40] = z1[kc];  // k + 32 + 64 kc
        }
      }
      __syncwarp();
      // ================= split pass + magnitudes =================
      float* mrow = mag + (it & 1) * kRow;
      const float* prow = mag + ((it + 1) & 1) * kRow;
      float sm = 0.f, skm = 0.f;
      {
        const float kfl = (float)lane;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 za = tile[lo_k + 40 * i];
          const float2 zp = tile[(i == 0 && lane == 0) ? 0 : 640 - 40 * i - hi_k];  // Z[512] is Z[0]
          const float er = za.x + zp.x, ei = za.y - zp.y, dr = za.x - zp.x, di = za.y + zp.y;
          const float t1 = wN[i].x * di + wN[i].y * dr, t2 = wN[i].y * di - wN[i].x * dr;
          const float xr = er + t1, xi = ei + t2;  // X[k], k = lane + 32 i
          const float yr = er - t1, yi = t2 - ei;  // X[512 - k]
          const float m0 = sqrt_fast(xr * xr + xi * xi), m1 = sqrt_fast(yr * yr + yi * yi);  // |X|; exact 0 stays 0
          mrow[lo_k + 40 * i] = m0;
          mrow[640 - 40 * i - hi_k] = m1;
          sm += m0 + m1;
          skm = fmaf(m0, kfl + (float)(32 * i), skm);
          skm = fmaf(m1, (float)(kM - 32 * i) - kfl, skm);
        }
      }
      if (lane == 0) {  // bin 256 pairs with itself: W_1024^256 = -i, X = (2 Re Z, -2 Im Z)
        const float2 za = tile[bpos(256)];
        const float xr = 2.f * za.x, xi = -2.f * za.y;
        const float m0 = sqrt_fast(xr * xr + xi * xi);
        mrow[bpos(256)] = m0;
        sm += m0;
        skm = fmaf(m0, 256.f, skm);
      }
      sm = warp_sum(sm);
      skm = warp_sum(skm);
      const float kc = sm > 0.f ? skm / sm : 0.f;  // centroid in bin units
      __syncwarp();  // magnitude row complete; Z no longer needed: the tile becomes the private mel slots

      // ================= phase B: 16 contiguous bins per lane =================
      BinAcc ac;
      ac.seg = ac.mx = ac.sl = ac.sxy = ac.fl = ac.bw = ac.mlo = ac.mhi = ac.pend = 0.f;
      ac.ninv = 0;
      ac.pp = priv + lane;
      {
        const float4* m4 = reinterpret_cast<const float4*>(mrow + pb);
        const float4* p4 = reinterpret_cast<const float4*>(prow + pb);
        const float4* x4 = reinterpret_cast<const float4*>(s_xtab + pb);
        const float4* l4 = reinterpret_cast<const float4*>(s_wlo + pb);
        const float4* h4 = reinterpret_cast<const float4*>(s_whi + pb);
        const float dk0 = (float)(16 * lane) - kc;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 mv = m4[g], pv = p4[g], xv = x4[g], lv = l4[g], hv = h4[g];
          const float mm[4] = {mv.x, mv.y, mv.z, mv.w}, pp[4] = {pv.x, pv.y, pv.z, pv.w};
          const float xx[4] = {xv.x, xv.y, xv.z, xv.w}, ll[4] = {lv.x, lv.y, lv.z, lv.w}, hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
          for (int u = 0; u < 4; ++u)
            bin_step(ac, (fmask >> (4 * g + u)) & 1u, dk0 + (float)(4 * g + u), mm[u], pp[u], xx[u], ll[u], hh[u],
                     (lane | g | u) != 0);
        }
      }
      const float m00 = mrow[0];
      const bool v00 = m00 > 1e-10f;
      const float l2k0 = v00 ? lg2_fast(m00) : 0.f;  // flatness counts bin 0, the slope regression does not
      if (lane == 31) {  // Nyquist bin
        const int bp = bpos(kM);
        bin_step(ac, (fmask >> 16) & 1u, (float)kM - kc, mrow[bp], prow[bp], s_xtab[bp], s_wlo[bp], s_whi[bp], true);
      }
      ac.pp[0] = ac.pend + ac.mlo;
      ac.pp[32] = ac.mhi;

      // ---- reductions ----
      float pre = ac.seg;  // inclusive prefix of the lanes' energies (bins ascend with the lane)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(kFull, pre, o);
        if (lane >= o) pre += up;
      }
      const float etot = __shfl_sync(kFull, pre, 31);
      const float plow = __shfl_sync(kFull, pre, 7);  // bins 0 .. 127 = B / 4 (speech.go:442)
      const float mx = warp_max(ac.mx);
      const float sl = warp_sum(ac.sl);
      const float sxy = warp_sum(ac.sxy);
      const float fl = warp_sum(ac.fl);
      const float bw = warp_sum(ac.bw);
      const int ninv = __reduce_add_sync(kFull, ac.ninv);
      float sxinv = 0.f, sxxinv = 0.f;
      if (ninv > 0) {  // rare: some bin (k >= 1) is below 1e-10 and leaves the slope regression
        for (int j = 0; j < 16; ++j) {
          if ((lane | j) == 0) continue;
          const float m = mrow[pb + j], xv = s_xtab[pb + j];
          if (!(m > 1e-10f)) {
            sxinv += xv;
            sxxinv = fmaf(xv, xv, sxxinv);
          }
        }
        if (lane == 31 && !(mrow[bpos(kM)] > 1e-10f)) {
          const float xv = s_xtab[bpos(kM)];
          sxinv += xv;
          sxxinv = fmaf(xv, xv, sxxinv);
        }
        sxinv = warp_sum(sxinv);
        sxxinv = warp_sum(sxxinv);
      }
      // ---- rolloff: first bin whose cumulative energy reaches 85 % (spectral_rolloff.go:19-55).  The lane whose
      //      range contains the crossing is found from the prefix; its 16 (17) bins are then scanned by the warp.
      int rk = kB - 1;
      {
        const float target = 0.85f * etot;
        const float excl = pre - ac.seg;
        const unsigned cb = __ballot_sync(kFull, (pre >= target) && (excl < target || lane == 0));
        if (cb) {
          const int cl = __ffs(cb) - 1;
          const float ex0 = __shfl_sync(kFull, excl, cl);
          const int nbn = cl == 31 ? 17 : 16;
          const float mj = lane < nbn ? mrow[bpos(16 * cl + lane)] : 0.f;
          float cum = mj * mj;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const float up = __shfl_up_sync(kFull, cum, o);
            if (lane >= o) cum += up;
          }
          const unsigned hit = __ballot_sync(kFull, lane < nbn && ex0 + cum >= target);
          rk = 16 * cl + (hit ? __ffs(hit) - 1 : nbn - 1);
        }
      }

      __syncwarp();  // private mel slots visible
      // ---- ln + DCT-II + lifter (mfcc.go:136-157) ----
      if (a.mfcc_on) {
        for (int f = lane; f < a.n_mel; f += 32) {
          float v = 0.f;
          for (int i = 0; i < ncontrib; ++i) v += wbf[s_moff[i * kMaxMel + f]];
          macc[f] = v > 0.f ? __logf(v) : -23.025850929940457f;  // ln(1e-10)
        }
        __syncwarp();
        // coefficient c by the lane pair (2c, 2c+1): each half sums every other filter
        const int nmp = a.n_mel | 1;
        for (int c0 = 0; c0 < a.n_mfcc; c0 += 16) {
          const int c = c0 + (lane >> 1);
          float acc = 0.f;
          if (c < a.n_mfcc)
            for (int f = lane & 1; f < a.n_mel; f += 2) acc = fmaf(macc[f], s_dct[c * nmp + f], acc);
          acc += __shfl_xor_sync(kFull, acc, 1);
          if (c < a.n_mfcc && !(lane & 1) && out_ok) fo[a.o_mfcc + t * a.n_mfcc + c] = (double)(acc * s_lift[c]);
        }
      }
      // ---- park the raw sums of this frame; finished in FP64 one frame per lane at the end of the run ----
      if (lane == 0) {
        float4* rs = reinterpret_cast<float4*>(rawsum + it * kRaw);
        rs[0] = make_float4(sm, kc, etot, __int_as_float(rk));
        rs[1] = make_float4(bw, sl, __int_as_float(ninv), mx);
        rs[2] = make_float4(sxy, l2k0, sxinv, sxxinv);
        rs[3] = make_float4(fl, plow, v00 ? 1.f : 0.f, 0.f);
      }
      // ---- ring: park the next frame's new samples over the oldest ones -----------------------------------
      if (more) {
        const int r2 = (int)((((t + 1) * H + (kN - H)) >> 1) & (kM - 1));
        if (NP) {
#pragma unroll
          for (int j = 0; j < NPR; ++j) ring2[(r2 + lane + 32 * j) & (kM - 1)] = make_float2((float)nx[j].x, (float)nx[j].y);
        } else {  // generic hop: straight from global memory (the latency is exposed once per frame)
          const double2* __restrict__ src = reinterpret_cast<const double2*>(x + (t + 1) * H + (kN - H));
          for (int j = lane; j < (H >> 1); j += 32) {
            const double2 d = __ldg(src + j);
            ring2[(r2 + j) & (kM - 1)] = make_float2((float)d.x, (float)d.y);
          }
        }
      }
      __syncwarp();  // tile / rows / macc / ring reused by the next frame
    }

    // ---- finish the run: lane i <-> frame t0 - 1 + i (i >= 1) -------------------------------------------
    if (lane >= 1 && lane < nfr) {
      const float4* rs4 = reinterpret_cast<const float4*>(rawsum + lane * kRaw);
      const float4 r0 = rs4[0], r1 = rs4[1], r2 = rs4[2], r3 = rs4[3];
      const int64_t t = t0 - 1 + lane;
      const float sm = r0.x, kc = r0.y, etot = r0.z, bw = r1.x, sl = r1.y, mx = r1.w, sxy = r2.x, l2k0 = r2.y,
                  sxinv = r2.z, sxxinv = r2.w, fl = r3.x, plow = r3.y, val0 = r3.z;
      const int rk = __float_as_int(r0.w), ninv = __float_as_int(r1.z);
      const double fs = a.freq_scale;
      const double dsm = (double)sm;
      fo[a.o_centroid + t] = (double)kc * fs;
      fo[a.o_rolloff + t] = etot > 0.f ? (double)rk * fs : 0.0;
      fo[a.o_bandwidth + t] = sm > 0.f ? sqrt((double)bw / dsm) * fs : 0.0;
      const float cnt = (float)(kB - 1 - ninv) + val0;  // bins with m > 1e-10 (spectral_flatness.go:31-70)
      double flat = 0.0;
      if (cnt > 0.f) {
        const double gm = exp2((double)sl / (double)cnt);
        const double am = dsm / (double)kB;
        if (am > 1e-10) {
          flat = gm / am;
          if (flat > 1.0) flat = 1.0;
        }
      }
      fo[a.o_flatness + t] = flat;
      const double rms = sqrt((double)etot / (double)kB);
      fo[a.o_crest + t] = rms > 0.0 ? (double)mx / rms : 0.0;
      double slope = 0.0;
      if (a.slope_on) {
        const double LG = 0.30102999566398120;  // log10(2)
        const double n = a.slope_ntot - (double)ninv;
        if (n >= 2.0) {
          const double sx = -(double)sxinv, sxx = a.slope_xxtot - (double)sxxinv;
          const double sy = LG * ((double)sl - (double)l2k0), sxyd = LG * (double)sxy;
          const double den = n * sxx - sx * sx;
          if (den != 0.0) slope = (n * sxyd - sx * sy) / den;
        }
      }
      fo[a.o_slope + t] = slope;
      if (t >= 1) fo[a.o_flux + t - 1] = sqrt((double)fl);
      if (t < a.Te) {
        fo[a.o_low + t] = etot > 0.f ? (double)plow / (double)etot : 0.0;
        fo[a.o_high + t] = etot > 0.f ? ((double)etot - (double)plow) / (double)etot : 0.0;
      }
    }
    __syncwarp();
  }
}

}  // namespace

// Eligibility of the second-generation kernel: N = 1024, 16-byte aligned PCM rows, and a mel bank whose
// filters are wide enough for kSlots lane-private slots (host copy of the region table).
bool stft_v2_eligible(const FpPlan& plan, const StftArgs& a) {
  static const bool off = std::getenv("SONAR_STFT_V1") != nullptr;  // diagnostic: force the first-generation kernel
  if (off || plan.N != kN) return false;
  if ((a.stride & 1) || (reinterpret_cast<uintptr_t>(a.pcm) & 15)) return false;
  if (a.hop <= 0 || a.hop > kN || (a.hop & 1)) return false;
  if (plan.h_regions.empty() || a.n_mel > kMaxMel || a.n_mfcc > kMaxMfcc || plan.split != 128) return false;
  auto region_of = [&](int k) {
    int r = 0;
    while (k >= plan.h_regions[r].next_b) ++r;
    return r;
  };
  int first[32], last[32];
  for (int j = 0; j < 32; ++j) {
    first[j] = region_of(16 * j);
    last[j] = region_of(j == 31 ? kB - 1 : 16 * j + 15);
    if (last[j] - first[j] + 2 > kSlots) return false;
  }
  for (int k = 1; k < kB; ++k)  // every mel region at least one bin wide (one private slot per boundary)
    if (region_of(k) - region_of(k - 1) > 1) return false;
  for (int f = 0; f < a.n_mel; ++f) {
    int cnt = 0;
    for (int j = 0; j < 32; ++j) cnt += (f + 1 >= first[j] - 1 && f + 1 <= last[j]);
    if (cnt > kMaxContrib) return false;
  }
  return true;
}

int launch_stft_v2(const FpPlan& plan, StftArgs& a, cudaStream_t st) {
  a.runs_per_stream = (int)((a.T + kRunOut - 1) / kRunOut);
  a.total_runs = (int64_t)a.runs_per_stream * a.n_streams;
  const V2Smem L = v2_layout(a.n_mel, a.n_mfcc);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t ctas = (a.total_runs + kW - 1) / kW;
  if (ctas > sms) ctas = sms;  // persistent: one CTA per SM, warps stride over the runs
  if (ctas < 1) ctas = 1;
  prof_begin("stft_features_kernel", st);
#define SONAR_V2_LAUNCH(NP)                                                                                      \
  do {                                                                                                           \
    SONAR_CUDA(cudaFuncSetAttribute(stft_v2_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total)); \
    stft_v2_kernel<NP><<<(unsigned)ctas, kW * 32, L.total, st>>>(a);                                               \
  } while (0)
  switch (a.hop) {
    case 64: SONAR_V2_LAUNCH(1); break;
    case 128: SONAR_V2_LAUNCH(2); break;
    case 256: SONAR_V2_LAUNCH(4); break;
    case 512: SONAR_V2_LAUNCH(8); break;
    default: SONAR_V2_LAUNCH(0); break;  // any other even hop
  }
#undef SONAR_V2_LAUNCH
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

}  // namespace sonar
