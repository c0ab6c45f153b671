// Fused framed-STFT + MFCC + spectral-descriptor kernel (FP32 arithmetic, f64 I/O).
//
// Replaces, for every frame t of every stream, the reference chain
//   analyzers.ComputeSTFTWithWindow        fingerprint/analyzers/spectral.go:385-545
//   spectral.MFCC.Compute                  algorithms/spectral/mfcc.go:113-164
//   MelScale.ApplyFilterBank               algorithms/spectral/mel_scale.go:89-105
//   centroid/rolloff/bandwidth/flatness/   algorithms/spectral/spectral_*.go
//     crest/slope/flux
//   low/high band energy ratios            fingerprint/extractors/speech.go:436-456
// without ever materialising the spectrogram in HBM.
//
// Mapping (geometry FftGeom<R1,R2>, N = 2*R1*R2 real samples per frame):
//   * one warp owns a run of consecutive frames of one stream and walks it F = 32/R1
//     frames at a time;
//   * pass 1: lane = column n2, radix-R1 DFT in registers on z[n] = x[2n] + i x[2n+1]
//     (f64 samples loaded as coalesced 16-byte pairs, converted, window applied from
//     registers), twiddle W_M^(n2*k1), transposed through a padded per-warp smem tile;
//   * pass 2: lane = (frame slot, k1), radix-R2 DFT in registers -> Z[k1 + R1*k2];
//   * split pass: partner Z[M-k] fetched with warp shuffles, X[k] formed in registers,
//     |X| written once to a padded smem row (bank-conflict free both for the strided
//     writes and for the contiguous reads of the next phase);
//   * phase B: each lane scans R2 contiguous bins and accumulates every per-frame sum
//     (mel regions with on-the-fly triangular weights, rolloff segment energy, flatness
//     log-sum, slope regression sums, positive flux against the previous frame's row,
//     bandwidth around the already-reduced centroid), followed by warp-shuffle
//     reductions, ln + DCT-II + lifter, and f64 stores of the 13 + 9 outputs.
#include <cfloat>
#include <cmath>

#include "common.h"
#include "fft_regs.cuh"

namespace sonar {

namespace {

#ifndef SONAR_STFT_WARPS
#define SONAR_STFT_WARPS 8
#endif
#ifndef SONAR_STFT_MINBLOCKS
#define SONAR_STFT_MINBLOCKS 1
#endif
constexpr int kWarps = SONAR_STFT_WARPS;  // warps per CTA
constexpr int kRunIters = 16;     // warp iterations per run (first frame of a run is flux warm-up)
constexpr unsigned kFull = 0xffffffffu;
// Per-lane row of mel partial sums (slot r = falling part of filter r-1 + rising part of filter r).  68 floats,
// 16-byte aligned; the row stride of 68 words puts lane j's column c in bank (4j + c) mod 32, so the strided
// column reads of the final reduction are at worst 2-way conflicted.
constexpr int kMelRow = kMaxMel + 4;

enum { MODE_FEATURES = 0, MODE_SPECTRUM = 1 };

__device__ __forceinline__ float2 ld_pair_f32(const double* p, bool aligned16) {
  if (aligned16) {
    double2 d = __ldg(reinterpret_cast<const double2*>(p));
    return make_float2((float)d.x, (float)d.y);
  }
  return make_float2((float)__ldg(p), (float)__ldg(p + 1));
}

template <int R, int K>
__device__ __forceinline__ void p1_store(const float2 (&v)[R], const float2* __restrict__ tw, int tw_stride,
                                         float2* __restrict__ xb, int xrow) {
  if constexpr (K < R) {
    float2 y = v[K];
    if constexpr (K > 0) y = cmul(y, tw[K * tw_stride]);
    xb[K * xrow] = y;
    p1_store<R, K + 1>(v, tw, tw_stride, xb, xrow);
  }
}

template <int R, int K>
__device__ __forceinline__ void p1_load(float2 (&v)[R], const float2 (&wv)[R], const double* __restrict__ px,
                                        int step, bool aligned16) {
  if constexpr (K < R) {
    float2 s = ld_pair_f32(px + (int64_t)K * step, aligned16);
    v[K] = make_float2(s.x * wv[K].x, s.y * wv[K].y);
    p1_load<R, K + 1>(v, wv, px, step, aligned16);
  }
}

template <int R, int K>
__device__ __forceinline__ void p2_load(float2 (&z)[R], const float2* __restrict__ xb) {
  if constexpr (K < R) {
    z[K] = xb[K];
    p2_load<R, K + 1>(z, xb);
  }
}

// Split pass for one output bin k = k1 + R1*K2 (compile-time K2).  Returns X[k].
template <int R1, int R2, int K2>
__device__ __forceinline__ float2 split_bin(const float2 (&z)[R2], int k1, int src_lane, float2 wn) {
  constexpr int PK = R2 - 1 - K2;          // partner register for k1 >= 1
  constexpr int PK0 = (R2 - K2) % R2;      // partner register for k1 == 0 (same lane)
  float pr = __shfl_sync(kFull, z[PK].x, src_lane);
  float pi = __shfl_sync(kFull, z[PK].y, src_lane);
  if (k1 == 0) {
    pr = z[PK0].x;
    pi = z[PK0].y;
  }
  const float2 a = z[K2];
  const float er = a.x + pr, ei = a.y - pi, dr = a.x - pr, di = a.y + pi;
  const float2 w = cmul(wn, w64(K2 * (32 / R2)));  // W_N^(k1 + R1*K2)
  return make_float2(er + (w.x * di + w.y * dr), ei + (w.y * di - w.x * dr));
}

struct FrameSums {  // per-lane partial sums of phase A
  float sm, skm;
};

template <int R1, int R2, int K2>
__device__ __forceinline__ void split_all_features(const float2 (&z)[R2], int k1, int src_lane, float2 wn,
                                                   float* __restrict__ mrow, FrameSums& s) {
  if constexpr (K2 < R2) {
    const float2 x = split_bin<R1, R2, K2>(z, k1, src_lane, wn);
    const float p = x.x * x.x + x.y * x.y;
    const float m = p > 0.f ? p * rsqrtf(p) : 0.f;
    constexpr int kb = R1 * K2;
    mrow[k1 + kb + (kb >> 5)] = m;
    s.sm += m;
    s.skm += m * (float)(k1 + kb);
    split_all_features<R1, R2, K2 + 1>(z, k1, src_lane, wn, mrow, s);
  }
}

template <int R1, int R2, int K2>
__device__ __forceinline__ void split_all_spectrum(const float2 (&z)[R2], int k1, int src_lane, float2 wn,
                                                   bool ok, double* __restrict__ mag, double* __restrict__ ph,
                                                   double* __restrict__ cx) {
  if constexpr (K2 < R2) {
    const float2 x = split_bin<R1, R2, K2>(z, k1, src_lane, wn);
    const int k = k1 + R1 * K2;
    if (ok) {
      // cmplx.Abs is a scaled hypot (spectral.go:492); sqrt(re^2+im^2) in f64 from the f32 bins
      mag[k] = sqrt((double)x.x * (double)x.x + (double)x.y * (double)x.y);
      if (ph) ph[k] = (double)atan2f(x.y, x.x);
      if (cx) {
        cx[2 * k] = (double)x.x;
        cx[2 * k + 1] = (double)x.y;
      }
    }
    split_all_spectrum<R1, R2, K2 + 1>(z, k1, src_lane, wn, ok, mag, ph, cx);
  }
}

template <int R1>
__device__ __forceinline__ float slot_sum(float v) {
#pragma unroll
  for (int off = R1 / 2; off >= 1; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  return v;
}
template <int R1>
__device__ __forceinline__ float slot_max(float v) {
#pragma unroll
  for (int off = R1 / 2; off >= 1; off >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, off));
  return v;
}

struct SmemLayout {
  size_t off_tw1, off_xtab, off_dct, off_lift, off_regions, off_chunk, off_warp, per_warp, total;
  size_t w_xbuf, w_carry, w_mel, w_melpriv;
};

template <int R1, int R2>
__host__ __device__ inline SmemLayout smem_layout(int n_mel, int n_mfcc, int n_regions) {
  using G = FftGeom<R1, R2>;
  SmemLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += (bytes + 15) & ~(size_t)15;
    return r;
  };
  L.off_tw1 = take(sizeof(float2) * G::M);
  L.off_xtab = take(sizeof(float) * (G::B + (G::B >> 5) + 2));  // padded like a magnitude row (index k + (k >> 5))
  L.off_dct = take(sizeof(float) * (size_t)n_mfcc * (n_mel | 1));
  L.off_lift = take(sizeof(float) * n_mfcc);
  L.off_regions = take(sizeof(MelRegion) * n_regions);
  L.off_chunk = take(sizeof(int) * R1);
  o = (o + 127) & ~(size_t)127;  // per-warp regions start on a bank-0 boundary
  L.off_warp = o;
  size_t w = 0;
  auto wtake = [&](size_t bytes) {
    size_t r = w;
    w += (bytes + 15) & ~(size_t)15;
    return r;
  };
  // Phase B reads one magnitude row per slot with the same instruction; with two slots per warp (F == 2) slot 0
  // reads a row of the exchange tile and slot 1 a carry row (or vice versa for the previous-frame rows).  The
  // exchange tile starts on bank 0, every row is a multiple of 32 words long and the carry rows start on bank
  // 16, so the two half-warps always hit disjoint halves of the banks.
  L.w_xbuf = wtake(sizeof(float2) * G::F * G::XSLOT);
  w = ((w + 127) & ~(size_t)127) + 64;
  L.w_carry = wtake(sizeof(float) * 2 * G::MAGROW);
  w = (w + 127) & ~(size_t)127;
  L.w_mel = wtake(sizeof(float) * G::F * (kMaxMel + 4));
  L.w_melpriv = wtake(sizeof(float) * 32 * kMelRow);  // one private row of mel partials per lane
  L.per_warp = (w + 127) & ~(size_t)127;
  L.total = o + L.per_warp * kWarps;
  return L;
}

template <int R1, int R2, int MODE>
__global__ void __launch_bounds__(kWarps * 32, SONAR_STFT_MINBLOCKS) stft_kernel(const StftArgs a) {
  using G = FftGeom<R1, R2>;
  constexpr int F = G::F;
  constexpr int WARM = (MODE == MODE_FEATURES) ? 1 : 0;
  constexpr int RUN_OUT = kRunIters * F - WARM;
  // current-iteration magnitude rows of slots 0..F-2 alias the exchange tile
  static_assert(sizeof(float) * (F > 1 ? (F - 1) : 1) * G::MAGROW <= sizeof(float2) * F * G::XSLOT,
                "magnitude rows must fit in the exchange tile");

  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = smem_layout<R1, R2>(a.n_mel, a.n_mfcc, a.n_regions);
  float2* s_tw1 = reinterpret_cast<float2*>(smem + L.off_tw1);
  float* s_xtab = reinterpret_cast<float*>(smem + L.off_xtab);
  float* s_dct = reinterpret_cast<float*>(smem + L.off_dct);
  float* s_lift = reinterpret_cast<float*>(smem + L.off_lift);
  MelRegion* s_reg = reinterpret_cast<MelRegion*>(smem + L.off_regions);
  int* s_chunk = reinterpret_cast<int*>(smem + L.off_chunk);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wbase = smem + L.off_warp + (size_t)warp * L.per_warp;
  float2* xbuf = reinterpret_cast<float2*>(wbase + L.w_xbuf);
  float* carry = reinterpret_cast<float*>(wbase + L.w_carry);
  float* melacc = reinterpret_cast<float*>(wbase + L.w_mel);
  float* melpriv = reinterpret_cast<float*>(wbase + L.w_melpriv) + lane * kMelRow;

  // ---- stage the tables -------------------------------------------------------
  for (int i = threadIdx.x; i < G::M; i += blockDim.x) s_tw1[i] = a.tw1[i];
  if (MODE == MODE_FEATURES) {
    for (int i = threadIdx.x; i < G::B; i += blockDim.x) s_xtab[G::mag_index(i)] = a.xtab[i];
    const int nmp = a.n_mel | 1;
    for (int i = threadIdx.x; i < a.n_mfcc * a.n_mel; i += blockDim.x)
      s_dct[(i / a.n_mel) * nmp + (i % a.n_mel)] = a.dct[i];
    for (int i = threadIdx.x; i < a.n_mfcc; i += blockDim.x) s_lift[i] = a.lift[i];
    for (int i = threadIdx.x; i < a.n_regions; i += blockDim.x) s_reg[i] = a.regions[i];
    for (int i = threadIdx.x; i < R1; i += blockDim.x) s_chunk[i] = a.chunk_region[i];
  }
  __syncthreads();

  // ---- lane identities ----------------------------------------------------------
  const int n2 = lane % R2, p1slot = lane / R2;  // pass 1
  const int k1 = lane % R1, slot = lane / R1;    // pass 2 and later
  const int src_lane = slot * R1 + ((R1 - k1) % R1);
  float2 wv[R1];
#pragma unroll
  for (int i = 0; i < R1; i++) wv[i] = __ldg(a.win2 + R2 * i + n2);
  const float2 wn = __ldg(a.wn + k1);
  const bool aligned16 = ((a.hop & 1) == 0) && ((a.stride & 1) == 0) &&
                         ((reinterpret_cast<uintptr_t>(a.pcm) & 15) == 0);
  const int H = a.hop;
  const int64_t T = a.T;

  for (int64_t run = (int64_t)blockIdx.x * kWarps + warp; run < a.total_runs;
       run += (int64_t)gridDim.x * kWarps) {
    const int s = (int)(run / a.runs_per_stream);
    const int64_t t0 = (run % a.runs_per_stream) * (int64_t)RUN_OUT;
    const int64_t tend = (t0 + RUN_OUT < T) ? t0 + RUN_OUT : T;
    const double* __restrict__ x = a.pcm + (int64_t)s * a.stride;

    for (int it = 0; it < kRunIters; ++it) {
      const int64_t tbase = t0 - WARM + (int64_t)it * F;
      if (tbase >= tend) break;  // warp-uniform

      // ================= pass 1 =================
#pragma unroll
      for (int rd = 0; rd < G::P1_ROUNDS; ++rd) {
        const int fs = rd * G::P1_FPR + p1slot;
        int64_t t = tbase + fs;
        t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
        float2 v[R1];
        p1_load<R1, 0>(v, wv, x + t * H + 2 * n2, 2 * R2, aligned16);
        FftReg<R1>::run(v);
        p1_store<R1, 0>(v, s_tw1 + n2, R2, xbuf + fs * G::XSLOT + n2, G::XROW);
      }
      __syncwarp();

      // ================= pass 2 =================
      float2 z[R2];
      p2_load<R2, 0>(z, xbuf + slot * G::XSLOT + k1 * G::XROW);
      FftReg<R2>::run(z);
      __syncwarp();  // exchange tile is free again

      const int64_t t = tbase + slot;
      const bool out_ok = (t >= t0) && (t < tend);

      if constexpr (MODE == MODE_SPECTRUM) {
        double* mg = a.mag + t * G::B;
        double* ph = a.phase ? a.phase + t * G::B : nullptr;
        double* cx = a.cplx ? a.cplx + t * 2 * G::B : nullptr;
        split_all_spectrum<R1, R2, 0>(z, k1, src_lane, wn, out_ok, mg, ph, cx);
        if (k1 == 0 && out_ok) {
          const float xm = (z[0].x + z[0].x) - (z[0].y + z[0].y);
          mg[G::M] = fabs((double)xm);
          if (ph) ph[G::M] = (double)atan2f(0.f, xm);
          if (cx) {
            cx[2 * G::M] = (double)xm;
            cx[2 * G::M + 1] = 0.0;
          }
        }
      } else {
        // ================= split pass + magnitudes =================
        // rows: slots 0..F-2 -> alias of the exchange tile; slot F-1 -> carry[it&1]
        float* cur_rows = reinterpret_cast<float*>(xbuf);
        float* mrow = (slot == F - 1) ? carry + (it & 1) * G::MAGROW : cur_rows + slot * G::MAGROW;
        const float* prow = (slot == 0) ? carry + ((it + 1) & 1) * G::MAGROW
                                        : ((slot - 1 == F - 1) ? carry + (it & 1) * G::MAGROW
                                                               : cur_rows + (slot - 1) * G::MAGROW);
        FrameSums ps{0.f, 0.f};
        split_all_features<R1, R2, 0>(z, k1, src_lane, wn, mrow, ps);
        if (k1 == 0) {  // Nyquist bin M
          const float xm = fabsf((z[0].x + z[0].x) - (z[0].y + z[0].y));
          mrow[G::mag_index(G::M)] = xm;
          ps.sm += xm;
          ps.skm += xm * (float)G::M;
        }
        const float sm = slot_sum<R1>(ps.sm);
        const float skm = slot_sum<R1>(ps.skm);
        const float kc = sm > 0.f ? skm / sm : 0.f;  // centroid in bin units
        float* macc = melacc + slot * (kMaxMel + 4);
        // Mel partials: every lane scans a contiguous run of bins, i.e. a contiguous run of filter regions.
        // It writes its partial sums to a PRIVATE row with plain stores (each slot once: the falling part of
        // the region being left is merged with the pending rising part of the previous one), and the rows
        // of the slot's R1 lanes are summed afterwards — no shared-memory atomics (float atomicAdd on
        // shared memory is a compare-and-swap loop).
        {
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          float4* zr = reinterpret_cast<float4*>(melpriv);
          for (int i = 0; i < (a.n_mel + 3 + 3) / 4; ++i) zr[i] = z;
        }
        float pend = 0.f;
        __syncwarp();

        // ================= phase B: contiguous scan =================
        const int kb0 = k1 * R2;
        const int nb = R2 + ((k1 == R1 - 1) ? 1 : 0);
        float seg = 0.f, plow = 0.f, mx = 0.f, sl = 0.f, cnt = 0.f, sxy = 0.f, l2k0 = 0.f;
        float ninv = 0.f, sxinv = 0.f, sxxinv = 0.f, fl = 0.f, bw = 0.f;
        int r = s_chunk[k1];
        MelRegion reg = s_reg[r];
        float mlo = 0.f, mhi = 0.f;
#pragma unroll 4
        for (int i = 0; i < nb; ++i) {
          const int k = kb0 + i;
          const int idx = G::mag_index(k);
          const float m = mrow[idx];
          const float mp = prow[idx];
          const float xv = s_xtab[idx];  // padded like the magnitude rows: the 16 lanes of a slot hit 16 banks
          const float p = m * m;
          const float kf = (float)k;
          while (k >= reg.next_b) {
            melpriv[r > 0 ? r - 1 : 0] = pend + mlo;
            pend = mhi;
            mlo = 0.f;
            mhi = 0.f;
            ++r;
            reg = s_reg[r];
          }
          mlo = fmaf(p, (reg.bhi - kf) * reg.inv_f, mlo);
          mhi = fmaf(p, (kf - reg.blo) * reg.inv_r, mhi);
          seg += p;
          if (k < a.split) plow += p;
          mx = fmaxf(mx, m);
          const float dk = kf - kc;
          bw = fmaf(dk * dk, m, bw);
          const bool valid = m > 1e-10f;
          const float l2 = __log2f(m);
          if (valid) {
            sl += l2;
            cnt += 1.f;
          }
          if (k >= 1) {
            if (valid) {
              sxy = fmaf(xv, l2, sxy);
            } else {
              ninv += 1.f;
              sxinv += xv;
              sxxinv = fmaf(xv, xv, sxxinv);
            }
          } else if (valid) {
            l2k0 = l2;
          }
          const float d = m - mp;
          if (d > 0.f) fl = fmaf(d, d, fl);
        }
        melpriv[r > 0 ? r - 1 : 0] = pend + mlo;
        melpriv[r] = mhi;

        // ---- reductions over the R1 lanes of this slot ----
        float pre = seg;  // inclusive prefix of segment energies (bins ascending with k1)
#pragma unroll
        for (int off = 1; off < R1; off <<= 1) {
          const float up = __shfl_up_sync(kFull, pre, off, R1);
          if (k1 >= off) pre += up;
        }
        const float etot = __shfl_sync(kFull, pre, slot * R1 + R1 - 1);
        plow = slot_sum<R1>(plow);
        mx = slot_max<R1>(mx);
        sl = slot_sum<R1>(sl);
        cnt = slot_sum<R1>(cnt);
        sxy = slot_sum<R1>(sxy);
        fl = slot_sum<R1>(fl);
        bw = slot_sum<R1>(bw);
        l2k0 = __shfl_sync(kFull, l2k0, slot * R1);
        ninv = slot_sum<R1>(ninv);
        if (ninv > 0.f) {  // uniform across the slot after the reduction
          sxinv = slot_sum<R1>(sxinv);
          sxxinv = slot_sum<R1>(sxxinv);
        }

        // ---- rolloff: first bin whose cumulative energy reaches 85 % (spectral_rolloff.go:19-55)
        const float target = 0.85f * etot;
        const float excl = pre - seg;
        const bool crossing = (pre >= target) && (excl < target || k1 == 0);
        int rk = G::B - 1;
        if (crossing) {
          float cum = excl;
          rk = kb0 + nb - 1;
          for (int i = 0; i < nb; ++i) {
            const float m = mrow[G::mag_index(kb0 + i)];
            cum = fmaf(m, m, cum);
            if (cum >= target) {
              rk = kb0 + i;
              break;
            }
          }
        }
        // lowest crossing lane wins (cumulative sums are monotone, so there is exactly one
        // unless rounding makes two adjacent lanes both qualify)
        const unsigned slot_mask = (R1 == 32) ? kFull : (((1u << R1) - 1u) << (slot * R1));
        const unsigned cb = __ballot_sync(kFull, crossing) & slot_mask;
        if (cb) rk = __shfl_sync(kFull, rk, __ffs(cb) - 1);

        __syncwarp();  // private mel rows visible

        // ---- ln + DCT-II + lifter (mfcc.go:136-157) ----
        if (a.mfcc_on) {
          const float* rows = melpriv - lane * kMelRow + (slot * R1) * kMelRow;  // the R1 rows of this slot
          for (int f = k1; f < a.n_mel; f += R1) {
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < R1; ++j) v += rows[j * kMelRow + f + 1];
            macc[f + 1] = v > 0.f ? __logf(v) : -23.025850929940457f;  // ln(1e-10)
          }
          __syncwarp();
          const int nmp = a.n_mel | 1;
          for (int c = k1; c < a.n_mfcc; c += R1) {
            float acc = 0.f;
            for (int f = 0; f < a.n_mel; ++f) acc = fmaf(macc[f + 1], s_dct[c * nmp + f], acc);
            acc *= s_lift[c];
            if (out_ok)
              a.feat[(int64_t)s * a.feat_stride + a.o_mfcc + t * a.n_mfcc + c] = (double)acc;
          }
        }

        // ---- scalar outputs ----
        if (k1 == 0 && out_ok) {
          double* fo = a.feat + (int64_t)s * a.feat_stride;
          const double fs = a.freq_scale;
          const double dsm = (double)sm;
          fo[a.o_centroid + t] = (double)kc * fs;
          fo[a.o_rolloff + t] = etot > 0.f ? (double)rk * fs : 0.0;
          fo[a.o_bandwidth + t] = sm > 0.f ? sqrt((double)bw / dsm) * fs : 0.0;
          double flat = 0.0;
          if (cnt > 0.f) {
            const double gm = exp2((double)sl / (double)cnt);
            const double am = dsm / (double)G::B;
            if (am > 1e-10) {
              flat = gm / am;
              if (flat > 1.0) flat = 1.0;
            }
          }
          fo[a.o_flatness + t] = flat;
          const double rms = sqrt((double)etot / (double)G::B);
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
        __syncwarp();  // rows/melacc reused next iteration
      }
    }
  }
}

template <int R1, int R2>
int launch_geom(const FpPlan& plan, StftArgs& a, bool spectrum, cudaStream_t st) {
  using G = FftGeom<R1, R2>;
  const int warm = spectrum ? 0 : 1;
  const int run_out = kRunIters * G::F - warm;
  a.runs_per_stream = (int)((a.T + run_out - 1) / run_out);
  a.total_runs = (int64_t)a.runs_per_stream * a.n_streams;
  const SmemLayout L = smem_layout<R1, R2>(a.n_mel, a.n_mfcc, a.n_regions);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t ctas = (a.total_runs + kWarps - 1) / kWarps;
  if (ctas > (int64_t)sms * SONAR_STFT_MINBLOCKS) ctas = (int64_t)sms * SONAR_STFT_MINBLOCKS;  // persistent: resident CTAs only, warps stride over the runs
  if (ctas < 1) ctas = 1;
  prof_begin(spectrum ? "stft_spectrum_kernel" : "stft_features_kernel", st);
  if (spectrum) {
    auto k = stft_kernel<R1, R2, MODE_SPECTRUM>;
    SONAR_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    k<<<(unsigned)ctas, kWarps * 32, L.total, st>>>(a);
  } else {
    auto k = stft_kernel<R1, R2, MODE_FEATURES>;
    SONAR_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    k<<<(unsigned)ctas, kWarps * 32, L.total, st>>>(a);
  }
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

}  // namespace

bool stft_supported(int w) { return w == 256 || w == 512 || w == 1024 || w == 2048; }
// lengths served by the float64 route only (spectral_exact.cu; go-dsp's Bluestein for lengths that are not powers of two)
bool stft_exact_only(int w) { return w >= 8 && w <= 2048 && !stft_supported(w); }

int launch_stft_features(const FpPlan& plan, StftArgs& a, bool spectrum, cudaStream_t st) {
  if (!spectrum && stft_v3_eligible(plan, a)) return launch_stft_v3(plan, a, st);
  if (!spectrum && stft_v2_eligible(plan, a)) return launch_stft_v2(plan, a, st);
  switch (plan.N) {
    case 256: return launch_geom<8, 16>(plan, a, spectrum, st);
    case 512: return launch_geom<16, 16>(plan, a, spectrum, st);
    case 1024: return launch_geom<16, 32>(plan, a, spectrum, st);
    case 2048: return launch_geom<32, 32>(plan, a, spectrum, st);
    default:
      return set_error(SONAR_ERR_UNSUPPORTED,
                       "window size must be 256, 512, 1024 or 2048 on the fused GPU path");
  }
}

}  // namespace sonar
