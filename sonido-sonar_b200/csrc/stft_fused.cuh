// Shared pieces of the fused STFT + MFCC + spectral-descriptor kernels (stft_v3.cu: single-role and warp-specialised
// forms; stft_v5.cu: the transform / scan kernel pair): geometry, per-bin scan step, rolloff search with its error
// margins, mbarrier / TMA bulk-copy wrappers.  See stft_v3.cu for the description of the data flow.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>

#include "common.h"
#include "fft_packed.cuh"

namespace sonar {
namespace {


#ifndef V3_WARPS
#define V3_WARPS 12
#endif
constexpr int kW3 = V3_WARPS;         // warps per CTA (one CTA per SM)
constexpr int kRun3 = 32;             // frames per segment: finished in FP64 one frame per lane
constexpr int kSeg3 = 4;              // segments per run: a warp walks kSeg3 * 32 consecutive frames of one stream, the
constexpr int kRunOut3 = kSeg3 * kRun3 - 1;  // first of which only warms the flux up (it has no predecessor at hand)
constexpr unsigned kFull3 = 0xffffffffu;
constexpr int kTileRow = 34;          // exchange tile row stride (float2): 16-byte rows, conflict-free LDS.128
constexpr int kSlots3 = 24;           // lane-private mel slots (float2: both frames of a pack; alias the tile in the scan)
#ifdef V3_RAW16
constexpr int kRaw3 = 16;
#else
constexpr int kRaw3 = 12;             // raw sums parked per frame
#endif
constexpr int kMaxContrib3 = 12;      // lanes that may hold a part of one mel filter
// ---- frames handed to the float64 re-evaluation (spectral_exact.cu) ---------------------------------------------
// The FP32 transform leaves an absolute error of a few 1e-7 of the frame's RMS spectral level on every bin (random, white;
// kEta bounds it generously: 2^-21 of the RMS level ~ 5 sigma, of the strongest bin where a bound must hold for sure).
constexpr float kEta = 4.76837158e-7f;       // 2^-21
constexpr int kExactBit = 0x40000000;        // set in the parked rolloff bin of a listed frame
constexpr float kRollSum = 4e-6f;            // relative error of an FP32 cumulative sum of <= 1025 squares, with margin
constexpr float kLogTau = 2e-5f;             // largest tolerated error bound of the mean of ln|X_k| (flatness; slope x 4)
constexpr float kMelRatio = 9.2e-5f;         // (2 kEta / 1e-4)^2: a mel band this far below the mean level errs by > 1e-4 in ln E
constexpr float kTinyMag = 1e-9f;            // magnitudes near the reference's 1e-10 validity threshold

template <int LOGN, int HR_>
struct V3G {
  static constexpr int N = 1 << LOGN, M = N / 2, B = M + 1;
  static constexpr int J = N / 32;            // samples per lane per frame = pass-1 radix
  static constexpr int FR = 64 / J;           // frames per warp iteration (2 or 4)
  static constexpr int PK = FR / 2;           // complex packs per iteration
  static constexpr int HR = HR_, H = 32 * HR_;
  static constexpr int RR = J + (FR - 1) * HR;  // ring rows per lane
  static constexpr int NEW = FR * HR;           // new rows per iteration
  static constexpr int BPL = M / 32;            // contiguous bins per lane in the scan (16 / 8)
  static constexpr int KSTR = 32 / PK;          // bin stride of pass 2's outputs: k = k1 + KSTR k2
  static constexpr int ROW = M + 4;             // floats per table row (xtab / wlo / whi), swizzled by spos
  static constexpr int PROW = M + 4;            // float2 per magnitude pair row (|X_a|, |X_b|), swizzled by ppos
  static_assert(LOGN == 10 || LOGN == 9, "N = 1024 or 512");
  static_assert(NEW <= RR, "hop must not exceed the window");
};

struct V3Smem {
  size_t tw, win, xtab, wlo, whi, fmask, moff, dct, lift, r0, warp0, per_warp, total;
  size_t w_tile, w_mag, w_raw, w_macc;
};

template <class G>
__host__ __device__ inline V3Smem v3_layout(int n_mel, int n_mfcc) {
  V3Smem L;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += (bytes + 15) & ~(size_t)15;
    return r;
  };
  L.tw = take(sizeof(float2) * G::J * 32);
  L.win = take(sizeof(float) * G::N);
  L.xtab = take(sizeof(float) * G::ROW);
  L.wlo = take(sizeof(float) * G::ROW);
  L.whi = take(sizeof(float) * G::ROW);
  L.fmask = take(sizeof(unsigned) * 32);
  L.moff = take(sizeof(unsigned short) * kMaxContrib3 * kMaxMel);
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
  L.w_tile = wtake(sizeof(float2) * 32 * kTileRow);
  L.w_mag = wtake(sizeof(float2) * G::PK * G::PROW);
  L.w_raw = wtake(sizeof(float) * kRaw3 * kRun3);
  L.w_macc = wtake(sizeof(float2) * (kMaxMel + 4));
  L.per_warp = w;
  L.total = o + w * kW3;
  return L;
}

__device__ __forceinline__ float warp_sum3(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull3, v, o);
  return v;
}
__device__ __forceinline__ float2 warp_sum3(float2 v) {  // both frames of a pack at once
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1)
    v = pk::add(v, make_float2(__shfl_xor_sync(kFull3, v.x, o), __shfl_xor_sync(kFull3, v.y, o)));
  return v;
}
__device__ __forceinline__ float warp_max3(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull3, v, o));
  return v;
}
__device__ __forceinline__ float sqrt_fast3(float x) {  // MUFU; sqrt(0) = 0
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rsqrt_fast3(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_fast3(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Per-lane accumulators of the scan; every float2 is (frame a, frame b) of a pack.
struct BinAcc3 {
  float2 seg, sl, sxy, fl, s0, s1, s2, mlo, mhi, pend;
  float mxa, mxb;
  float2* pp;  // next lane-private mel slot
};
__device__ __forceinline__ void acc_init(BinAcc3& s, float2* pp) {
  const float2 z = make_float2(0.f, 0.f);
  s.seg = s.sl = s.sxy = s.fl = s.s0 = s.s1 = s.s2 = s.mlo = s.mhi = s.pend = z;
  s.mxa = s.mxb = 0.f;
  s.pp = pp;
}
// One bin of the scan for both frames of a pack: m = (|X_a[k]|, |X_b[k]|), pv = |X[k]| of the frame before a.
// `flush`: the bin opens a new mel region, i.e. the falling part of the filter being left joins its pending rising
// part in the next private slot.  The centroid / bandwidth sums are taken relative to the lane's first bin (J = k -
// k0 is a compile-time constant: sum m, sum J m, sum J^2 m), which also removes them from the split pass; bins below
// 1e-10 (silence) are not tested here: the frame's smallest magnitude is tracked and the rare frame is redone exactly.
__device__ __forceinline__ void bin_step3(BinAcc3& s, bool flux_a, int jj, bool flush, float2 m, float pv, float xv,
                                          float wl, float wh) {
  if (flush) {
    *s.pp = pk::add(s.pend, s.mlo);
    s.pend = s.mhi;
    s.mlo = make_float2(0.f, 0.f);
    s.mhi = make_float2(0.f, 0.f);
    s.pp += 32;
  }
  const float2 p = __fmul2_rn(m, m);
  s.mlo = pk::fma(p, wl, s.mlo);
  s.mhi = pk::fma(p, wh, s.mhi);
  s.seg = pk::add(s.seg, p);
  s.s0 = pk::add(s.s0, m);
  if (jj > 0) {  // jj is a compile-time constant after unrolling
    s.s1 = pk::fma(m, (float)jj, s.s1);
    s.s2 = pk::fma(m, (float)(jj * jj), s.s2);
  }
  s.mxa = fmaxf(s.mxa, m.x);
  s.mxb = fmaxf(s.mxb, m.y);
  const float2 l2 = make_float2(lg2_fast3(m.x), lg2_fast3(m.y));
  s.sl = pk::add(s.sl, l2);
  s.sxy = pk::fma(l2, xv, s.sxy);  // xtab[0] == 0: bin 0 never enters the regression
  // flux: frame b against frame a; frame a against its predecessor only when that one is at hand (flux_a, a compile-time constant after unrolling: packs after
  // the first read it from the previous pack's row) -- the first pack's frame a is done in pass 2, see there
  const float2 d = make_float2(flux_a ? fmaxf(m.x - pv, 0.f) : 0.f, fmaxf(m.y - m.x, 0.f));
  s.fl = __ffma2_rn(d, d, s.fl);
}

template <int R, int K>
__device__ __forceinline__ void tw_apply(float2 (&v)[R], const float2* __restrict__ tw) {
  if constexpr (K < R) {
    v[K] = pk::mul(v[K], tw[K * 32]);
    tw_apply<R, K + 1>(v, tw);
  }
}

// Second look at a frame whose rolloff threshold sits within the FP32 SUMMATION error of a cumulative sum: the same
// magnitudes summed in float64 (their squares and partial sums are exact to ~1e-16), which leaves only the transform's
// own error `dfft` (relative to the total energy) as the margin.  Rare (~1 % of broadband frames), hence not inlined.
template <class G>
__device__ __noinline__ int rolloff_refine(const float* __restrict__ mrow2, float dfft, int lane) {
  constexpr int BPL = G::BPL;
  double e[BPL + 1], seg = 0.0;
#pragma unroll
  for (int j = 0; j < BPL; ++j) {
    const double m = (double)mrow2[2 * ppos(BPL * lane + j)];
    e[j] = m * m;
    seg += e[j];
  }
  e[BPL] = 0.0;
  if (lane == 31) {
    const double m = (double)mrow2[2 * ppos(G::M)];
    e[BPL] = m * m;
    seg += e[BPL];
  }
  double pre = seg;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double up = __shfl_up_sync(kFull3, pre, o);
    if (lane >= o) pre += up;
  }
  const double tot = __shfl_sync(kFull3, pre, 31), target = 0.85 * tot, excl = pre - seg;
  const unsigned cb = __ballot_sync(kFull3, (pre >= target) && (excl < target || lane == 0));
  if (!cb) return (G::B - 1) | kExactBit;  // not a number
  const int cl = __ffs(cb) - 1;
  int k = 0;
  double gap = 0.0;
  if (lane == cl) {
    double cum = excl, below = excl;
    bool found = false;
#pragma unroll
    for (int j = 0; j <= BPL; ++j) {
      if (j == BPL && lane != 31) break;
      const double nxt = cum + e[j];
      if (!found && nxt >= target) {
        found = true;
        k = j;
        below = cum;
        gap = fmin(nxt - target, target - cum);
      }
      cum = nxt;
    }
    (void)below;
  }
  k = __shfl_sync(kFull3, k, cl);
  gap = __shfl_sync(kFull3, gap, cl);
  return (BPL * cl + k) | ((gap > (double)dfft * tot) ? 0 : kExactBit);
}

// rolloff: first bin whose cumulative energy reaches 85 % (spectral_rolloff.go:19-55).  The lane whose range holds the
// crossing is found from the prefix of the lanes' energies; its BPL (+1) bins are then scanned by the warp.  The choice is
// safe when the threshold is further than `delta` (transform error + FP32 summation error, both relative to the total)
// from the cumulative sums on both sides of the chosen bin; otherwise rolloff_refine removes the summation error, and
// only a threshold within the transform's error `dfft` of a cumulative sum sets kExactBit (float64 from the PCM decides).
template <class G>
__device__ __forceinline__ int rolloff_bin(const float* __restrict__ mrow2, float pre, float seg, float etot, float dfft,
                                           int lane) {
  constexpr int BPL = G::BPL;
  int rk = (G::B - 1) | kExactBit;
  const float target = 0.85f * etot, delta = (dfft + kRollSum) * etot;
  const float excl = pre - seg;
  const unsigned cb = __ballot_sync(kFull3, (pre >= target) && (excl < target || lane == 0));
  if (cb) {
    const int cl = __ffs(cb) - 1;
    const float ex0 = __shfl_sync(kFull3, excl, cl);
    const int nbn = cl == 31 ? BPL + 1 : BPL;
    const float mj = lane < nbn ? mrow2[2 * ppos(BPL * cl + lane)] : 0.f;
    float cum = mj * mj;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(kFull3, cum, o);
      if (lane >= o) cum += up;
    }
    cum += ex0;
    const unsigned hit = __ballot_sync(kFull3, lane < nbn && cum >= target);
    const int h = hit ? __ffs(hit) - 1 : nbn - 1;
    float below = __shfl_up_sync(kFull3, cum, 1);
    if (lane == 0) below = ex0;
    const float gap = fminf(cum - target, target - below);  // both positive at lane h when the search was clean
    const float gh = __shfl_sync(kFull3, gap, h);
    if (hit && gh > delta) return BPL * cl + h;
#ifdef V3_NO_REFINE
    return (BPL * cl + h) | kExactBit;
#endif
  }
#ifdef V3_NO_REFINE
  return rk;
#else
  return rolloff_refine<G>(mrow2, dfft, lane);
#endif
}


// ---- mbarrier / TMA bulk copy ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit (cp.async.bulk), completion counted in bytes on an mbarrier:
// no registers are tied up while the rows travel from HBM
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("fence.proxy.async.shared::cta;\n"
               "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}

// Named barrier of two warps (the two warps of one role that share a scheduler): keeps them within a few instructions
// of each other, so the second one finds the code the first one fetched in the instruction caches -- the two loops
// together are 50 KB of straight-line code, more than the 32 KB L1.5 holds, and four independent streams per scheduler
// were fetch-bound (ncu: no_instruction 1.2 cycles per issue).
__device__ __forceinline__ void mate_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }


}  // namespace
}  // namespace sonar
