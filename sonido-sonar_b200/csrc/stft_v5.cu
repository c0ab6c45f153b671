// Fused framed-STFT + MFCC + spectral descriptors as a KERNEL PAIR (fifth generation; same arithmetic and results as
// stft_v3.cu, replacing the same reference functions: analyzers/spectral.go:385-545, algorithms/spectral/mfcc.go:113-164,
// mel_scale.go:89-105, spectral_*.go, extractors/speech.go:436-456).
//
// Why two kernels: an SM delivers instructions at full rate only while a loop body fits its 32 KB L1.5 instruction cache;
// beyond ~40 KB the issue rate of ANY kernel drops to 0.5-0.6 slots per scheduler-cycle (12-16 warps in step) and to
// 0.25-0.4 with the warps out of step (scripts/microbench/icache_bw.cu, measured on this B200).  The single-role kernel
// (stft_v3.cu) has a 43.5 KB loop body and issues 0.56; its warp-specialised form (stft_v4_kernel: 50 KB for both roles)
// gained nothing from four more resident warps.  Here each half gets its own kernel, ~20 KB (transform) and ~22 KB
// (scan) of loop body, and the occupancy that suits it:
//   * stft_v5_transform_kernel: 12 warps x 168 registers.  Sample ring in registers, the NEW rows of the next iteration
//     staged by ONE TMA bulk copy (cp.async.bulk + mbarrier) into the idle exchange tile, window, radix-32 x radix-32 on
//     the packed FP32 pipe, Hermitian split, |X| of both frames of a pack into a swizzled shared-memory row, and the row
//     leaves for a global workspace by ONE TMA bulk store per iteration;
//   * stft_v5_scan_kernel: up to 16 warps.  A warp pulls the rows of its run with TMA bulk loads, double buffered (the
//     next row travels while the current one is scanned), keeps the previous frame's magnitudes of its own bins in
//     registers for the flux, and does everything from the per-bin scan to the float64 finishing.
// The price is the workspace traffic: 2 x 4.1 KB per frame pair written and read once (L2 / HBM; 13.6 GB per 64 x 300 s
// against 7.1 GB of algorithmic bytes), which both kernels hide behind their arithmetic.
#include <map>
#include <mutex>

#include "stft_fused.cuh"

namespace sonar {
namespace {

#ifndef V5_TW
#define V5_TW 12
#endif
// (The register file is partitioned per scheduler, 16,384 registers each: 12 warps = 3 per scheduler x 168 registers and
// 16 warps = 4 x 128 are the two shapes that fill it; 13-15 warps put four warps on one scheduler and are capped at 128
// registers as well -- `too many resources requested` at 14 x 144 -- and at 128 the transform spills ~60 words per
// thread, which 217 KB of configured shared memory leave no L1 for.)
#ifdef V5_TREGS  // variant builds: an explicit register cap
#define V5_TATTR __maxnreg__(V5_TREGS)
#else
#define V5_TATTR __launch_bounds__(V5_TW * 32, 1)
#endif
constexpr int kTW5 = V5_TW;        // transform warps per CTA
constexpr int kSW5 = 16;           // scan warps per CTA (fewer when the mel bank's private slots do not fit)
constexpr int kRunFrames5 = 128;   // frames per run: a multiple of the 32-frame finishing segment and of FR

template <class G>
__host__ __device__ constexpr size_t row_bytes5() {
  return (sizeof(float2) * G::PK * G::PROW + 127) & ~(size_t)127;
}

__device__ __forceinline__ void tma_store_1d(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("fence.proxy.async.shared::cta;\n"
               "cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
               "cp.async.bulk.commit_group;" ::"l"(gdst),
               "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct V5Args {
  unsigned char* ws;        // workspace: row (s - s0) * iters_per_stream + i = iteration i (FR frames) of stream s
  int s0, ns;               // streams of this group
  int64_t iters_per_stream;
  int runs_per_stream;
  int64_t total_runs;
  int n_slots;              // lane-private mel slots per lane (scan kernel)
};

// ================================================ transform ==========================================================
struct V5TSmem {
  size_t tw, win, warp0, per_warp, total, w_tile, w_row;
};
template <class G>
__host__ __device__ inline V5TSmem v5t_layout() {
  V5TSmem L;
  size_t o = 0;
  L.tw = o;
  o += sizeof(float2) * G::J * 32;
  L.win = o;
  o += sizeof(float) * G::N;
  o = (o + 127) & ~(size_t)127;
  L.warp0 = o;
  L.w_tile = 0;
  L.w_row = (sizeof(float2) * 32 * kTileRow + 127) & ~(size_t)127;
  L.per_warp = L.w_row + row_bytes5<G>();
  L.total = o + L.per_warp * kTW5;
  return L;
}

template <int LOGN, int HR>
__global__ void V5_TATTR stft_v5_transform_kernel(const StftArgs a, const V5Args v) {
  using G = V3G<LOGN, HR>;
  constexpr int N = G::N, M = G::M, J = G::J, FR = G::FR, PK = G::PK, RR = G::RR, NEW = G::NEW, KSTR = G::KSTR,
                PROW = G::PROW, H = G::H;
  extern __shared__ __align__(128) unsigned char smem[];
  const V5TSmem L = v5t_layout<G>();
  float2* s_tw = reinterpret_cast<float2*>(smem + L.tw);
  float* s_win = reinterpret_cast<float*>(smem + L.win);
  __shared__ __align__(8) unsigned long long s_tbar[kTW5];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wb = smem + L.warp0 + (size_t)warp * L.per_warp;
  float2* tile = reinterpret_cast<float2*>(wb + L.w_tile);
  float2* rowbuf = reinterpret_cast<float2*>(wb + L.w_row);
  unsigned long long* tbar = &s_tbar[warp];
  constexpr unsigned kRowBytes = (unsigned)row_bytes5<G>();

  for (int i = threadIdx.x; i < J * 32; i += blockDim.x) {
    const int k1 = i / 32, l = i % 32;
    double dsn, dcs;
    sincospi(-2.0 * (double)((k1 * l) % N) / (double)N, &dsn, &dcs);
    s_tw[i] = make_float2((float)dcs, (float)dsn);
  }
  {
    const float* wsrc = reinterpret_cast<const float*>(a.win2);  // 0.5 w[n]: the 1/2 of the Hermitian split
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_win[i] = __ldg(wsrc + i);
  }
  for (int i = lane; i < (int)(kRowBytes / sizeof(float2)); i += 32) rowbuf[i] = make_float2(0.f, 0.f);  // pads: defined bytes
  if (threadIdx.x < kTW5) mbar_init(&s_tbar[threadIdx.x], 1u);
  __syncthreads();

  const int k1 = PK == 1 ? lane : (lane & 15);
  const int src = PK == 1 ? ((32 - lane) & 31) : ((lane & 16) | ((16 - k1) & 15));
  const bool k1zero = k1 == 0;
  const int64_t T = a.T;
  const int nyq = ppos(M);
  constexpr int NV = PK == 1 ? 4 : 8;
  int woff[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) woff[q] = ppos(k1 + KSTR * q) - KSTR * q;
  unsigned tphase = 0;
  const double* stage = reinterpret_cast<const double*>(tile);

  for (int64_t run = (int64_t)blockIdx.x * kTW5 + warp; run < v.total_runs; run += (int64_t)gridDim.x * kTW5) {
    const int sl = (int)(run / v.runs_per_stream);
    const int s = v.s0 + sl;
    const int64_t t0 = (run % v.runs_per_stream) * (int64_t)kRunFrames5;
    const int64_t tend = (t0 + kRunFrames5 < T) ? t0 + kRunFrames5 : T;
    const double* __restrict__ x = a.pcm + (int64_t)s * a.stride;
    const int nit = (int)((tend - t0 + FR - 1) / FR);
    unsigned char* wrow = v.ws + ((int64_t)sl * v.iters_per_stream + t0 / FR) * (int64_t)kRowBytes;

    float ring[RR];
    const double* __restrict__ xl = x + (t0 * H + lane);
    int64_t rows_left;
    {
      const int64_t g0 = t0 * H + lane;
      rows_left = (a.n - g0 + 31) >> 5;
      const int jhi = rows_left < RR ? (int)(rows_left < 0 ? 0 : rows_left) : RR;
#pragma unroll
      for (int j = 0; j < RR; ++j) ring[j] = (j < jhi) ? (float)__ldg(xl + 32 * j) : 0.f;
    }

    for (int it = 0; it < nit; ++it) {
      // ---- pass 1: radix-J over the lane's own samples, both frames of a pack at once ----
#pragma unroll
      for (int p = 0; p < PK; ++p) {
        float2 c[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const float w = s_win[lane + 32 * j];
          c[j] = make_float2(ring[j + 2 * p * G::HR] * w, ring[j + (2 * p + 1) * G::HR] * w);
        }
        pk::Fft<J>::run(c);
        tw_apply<J, 1>(c, s_tw + lane);
        float2* tp = tile + (p * J) * kTileRow + lane;
#pragma unroll
        for (int q = 0; q < J; ++q) tp[q * kTileRow] = c[q];
      }
      __syncwarp();
      // ---- pass 2 ----
      const bool more = it + 1 < nit;
      bool staged = false;
      {
        float2 z[32];
        const float4* rp = reinterpret_cast<const float4*>(tile + lane * kTileRow);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 f = rp[i];
          z[2 * i] = make_float2(f.x, f.y);
          z[2 * i + 1] = make_float2(f.z, f.w);
        }
        // the tile is idle until the next pass 1: the next iteration's NEW sample rows (one contiguous block of the
        // stream) land in it by one bulk copy, in flight behind pass 2 and the split
        __syncwarp();
        if (more) {
          const int r0 = FR * G::HR * (it + 1) + (RR - NEW);
          const int64_t gs = t0 * H + 32 * (int64_t)r0;
          staged = gs + 32 * NEW <= a.n && ((reinterpret_cast<uintptr_t>(x + gs) & 15) == 0);
          if (staged && lane == 0) {
            mbar_expect_tx(tbar, 32 * NEW * sizeof(double));
            tma_load_1d(tile, x + gs, 32 * NEW * sizeof(double), tbar);
          }
        }
        pk::Fft<32>::run(z);
        // the previous iteration's row must have left the buffer (its bulk store has READ it) before it is rewritten
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        float2* row = rowbuf + (PK == 1 ? 0 : (lane >> 4)) * PROW;
        float2 rinv = make_float2(0.f, 0.f);
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
          const float2 mine = k1zero ? z[(32 - k2) & 31] : z[31 - k2];
          const float2 pz = make_float2(__shfl_sync(kFull3, mine.x, src), __shfl_sync(kFull3, mine.y, src));
          const float2 zz = z[k2];
          const float2 xa = __fadd2_rn(zz, make_float2(pz.x, -pz.y));
          const float2 xb = __fadd2_rn(make_float2(zz.y, -zz.x), make_float2(pz.y, pz.x));
          const float2 qa = __fmul2_rn(xa, xa), qb = __fmul2_rn(xb, xb);
          const int e = woff[k2 % NV] + KSTR * k2;
          const float2 q = make_float2(qa.x + qa.y, qb.x + qb.y);
          const float2 qt = __fadd2_rn(q, make_float2(1e-36f, 1e-36f));
          const float2 ri = make_float2(rsqrt_fast3(qt.x), rsqrt_fast3(qt.y));
          const float2 m = __fmul2_rn(q, ri);  // |X| = q rsqrt(q); 1 / |X| feeds the weak-bin test
          rinv = __fadd2_rn(rinv, ri);
          row[e] = m;
        }
        if (k1zero) {  // Z[M] pairs with itself
          const float2 m = make_float2(fabsf(2.f * z[16].x), fabsf(2.f * z[16].y));
          rinv = __fadd2_rn(rinv, make_float2(__fdividef(1.f, fmaxf(m.x, 1e-18f)), __fdividef(1.f, fmaxf(m.y, 1e-18f))));
          row[nyq] = m;
        }
#pragma unroll
        for (int o = (PK == 1 ? 16 : 8); o >= 1; o >>= 1)
          rinv = __fadd2_rn(rinv, make_float2(__shfl_xor_sync(kFull3, rinv.x, o), __shfl_xor_sync(kFull3, rinv.y, o)));
        if (k1zero) row[M + 2] = rinv;  // sum 1 / |X_k| of the pack's two frames, in a free slot behind the Nyquist bin
      }
      __syncwarp();
      if (lane == 0) tma_store_1d(wrow + (int64_t)it * kRowBytes, rowbuf, kRowBytes);
      // ---- the next iteration's new rows join the ring ----
      if (staged) {
        mbar_wait(tbar, tphase);
        tphase ^= 1u;
#pragma unroll
        for (int j = 0; j < RR - NEW; ++j) ring[j] = ring[j + NEW];
#pragma unroll
        for (int j = 0; j < NEW; ++j) ring[RR - NEW + j] = (float)stage[lane + 32 * j];
        __syncwarp();  // every lane has its samples: pass 1 may overwrite the tile
      } else if (more) {
        double nx[NEW];
        const int r0 = FR * G::HR * (it + 1) + (RR - NEW);
        const double* __restrict__ src_p = xl + 32 * (int64_t)r0;
        const int64_t left = rows_left - r0;
        const int jhi = left < NEW ? (int)(left < 0 ? 0 : left) : NEW;
#pragma unroll
        for (int j = 0; j < NEW; ++j) nx[j] = (j < jhi) ? __ldg(src_p + 32 * j) : 0.0;
#pragma unroll
        for (int j = 0; j < RR - NEW; ++j) ring[j] = ring[j + NEW];
#pragma unroll
        for (int j = 0; j < NEW; ++j) ring[RR - NEW + j] = (float)nx[j];
      }
    }
  }
  if (lane == 0) tma_store_wait_all();  // the last rows have reached global memory before the CTA's shared memory goes
}

// ================================================== scan =============================================================
struct V5SSmem {
  size_t xtab, wlo, whi, fmask, moff, dct, lift, r0, warp0, per_warp, total, w_rows, w_priv, w_raw, w_macc;
};
template <class G>
__host__ __device__ inline V5SSmem v5s_layout(int n_mel, int n_mfcc, int n_slots, int warps) {
  V5SSmem L;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += (bytes + 15) & ~(size_t)15;
    return r;
  };
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
  L.w_rows = wtake(2 * row_bytes5<G>());
  L.w_priv = wtake(sizeof(float2) * (size_t)n_slots * 32);
  L.w_raw = wtake(sizeof(float) * kRaw3 * kRun3);
  L.w_macc = wtake(sizeof(float2) * (kMaxMel + 4));
  L.per_warp = w;
  L.total = o + w * (size_t)warps;
  return L;
}

template <int LOGN, int HR>
__global__ void __launch_bounds__(kSW5 * 32, 1) stft_v5_scan_kernel(const StftArgs a, const V5Args v) {
  using G = V3G<LOGN, HR>;
  constexpr int M = G::M, B = G::B, FR = G::FR, PK = G::PK, BPL = G::BPL, ROW = G::ROW, PROW = G::PROW;
  extern __shared__ __align__(128) unsigned char smem[];
  const int warps = blockDim.x >> 5;
  const V5SSmem L = v5s_layout<G>(a.n_mel, a.n_mfcc, v.n_slots, warps);
  float* s_xtab = reinterpret_cast<float*>(smem + L.xtab);
  float* s_wlo = reinterpret_cast<float*>(smem + L.wlo);
  float* s_whi = reinterpret_cast<float*>(smem + L.whi);
  unsigned* s_fmask = reinterpret_cast<unsigned*>(smem + L.fmask);
  unsigned short* s_moff = reinterpret_cast<unsigned short*>(smem + L.moff);
  float* s_dct = reinterpret_cast<float*>(smem + L.dct);
  float* s_lift = reinterpret_cast<float*>(smem + L.lift);
  int* s_r0 = reinterpret_cast<int*>(smem + L.r0);
  __shared__ int s_ncontrib;
  __shared__ float s_invw[kMaxMel];
  __shared__ __align__(8) unsigned long long s_bar[kSW5][2];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wb = smem + L.warp0 + (size_t)warp * L.per_warp;
  float2* wbf2 = reinterpret_cast<float2*>(wb);
  unsigned char* rows = wb + L.w_rows;
  float2* priv = reinterpret_cast<float2*>(wb + L.w_priv);
  float* rawsum = reinterpret_cast<float*>(wb + L.w_raw);
  float2* macc = reinterpret_cast<float2*>(wb + L.w_macc);
  constexpr int kZeroSlot = kMaxMel + 2;
  constexpr unsigned kRowBytes = (unsigned)row_bytes5<G>();

  // ---- tables, once per CTA (as in stft_v3_kernel) ----
  for (int k = threadIdx.x; k < ROW; k += blockDim.x) {
    s_xtab[k] = 0.f;
    s_wlo[k] = 0.f;
    s_whi[k] = 0.f;
  }
  if (threadIdx.x < 32) s_fmask[threadIdx.x] = 0u;
  if (threadIdx.x < kMaxMel) s_invw[threadIdx.x] = a.mel_invw[threadIdx.x];
  if (threadIdx.x == 0) s_ncontrib = 0;
  if (lane == 0) macc[kZeroSlot] = make_float2(0.f, 0.f);
  if (threadIdx.x < 2 * warps) mbar_init(&s_bar[0][0] + threadIdx.x, 1u);
  __syncthreads();
  for (int k = threadIdx.x; k < B; k += blockDim.x) {
    s_xtab[spos(k)] = a.xtab[k];
    int r = 0;
    while (k >= a.regions[r].next_b) ++r;
    const MelRegion reg = a.regions[r];
    const float kf = (float)k;
    s_wlo[spos(k)] = (reg.bhi - kf) * reg.inv_f;
    s_whi[spos(k)] = (kf - reg.blo) * reg.inv_r;
    int rp = 0;
    if (k > 0)
      while (k - 1 >= a.regions[rp].next_b) ++rp;
    if (r != rp && ((k % BPL) || k == M)) atomicOr(&s_fmask[k == M ? 31 : (k / BPL)], 1u << (k == M ? BPL : (k % BPL)));
    if ((k % BPL) == 0 && k < M) s_r0[k / BPL] = r;
  }
  {
    const int nmp = a.n_mel | 1;
    for (int i = threadIdx.x; i < a.n_mfcc * a.n_mel; i += blockDim.x)
      s_dct[(i / a.n_mel) * nmp + (i % a.n_mel)] = a.dct[i];
    for (int i = threadIdx.x; i < a.n_mfcc; i += blockDim.x) s_lift[i] = a.lift[i];
  }
  __syncthreads();
  const unsigned short zero_off = (unsigned short)((macc + kZeroSlot) - wbf2);  // the same for every warp
  const unsigned short priv_off = (unsigned short)(priv - wbf2);
  for (int f = threadIdx.x; f < kMaxMel; f += blockDim.x) {
    int cnt = 0;
    if (f < a.n_mel) {
      for (int j = 0; j < 32; ++j) {
        int rl = 0;
        const int kl = (j == 31) ? B - 1 : BPL * j + BPL - 1;
        while (kl >= a.regions[rl].next_b) ++rl;
        const int first = s_r0[j] - 1, last = rl;
        if (f + 1 >= first && f + 1 <= last && cnt < kMaxContrib3)
          s_moff[(cnt++) * kMaxMel + f] = (unsigned short)(priv_off + (f + 1 - first) * 32 + j);
      }
      atomicMax(&s_ncontrib, cnt);
    }
    for (int i = cnt; i < kMaxContrib3; ++i) s_moff[i * kMaxMel + f] = zero_off;
  }
  __syncthreads();
  const int ncontrib = s_ncontrib;
  const unsigned fmask = s_fmask[lane];
  const int64_t T = a.T;
  const int nyq = ppos(M);
  const float k0f = (float)(BPL * lane);
  unsigned long long* bar = &s_bar[warp][0];
  unsigned ph[2] = {0u, 0u};  // parity of each buffer's next completion

  for (int64_t run = (int64_t)blockIdx.x * warps + warp; run < v.total_runs; run += (int64_t)gridDim.x * warps) {
    const int sl = (int)(run / v.runs_per_stream);
    const int s = v.s0 + sl;
    const int64_t t0 = (run % v.runs_per_stream) * (int64_t)kRunFrames5;
    const int64_t tend = (t0 + kRunFrames5 < T) ? t0 + kRunFrames5 : T;
    double* __restrict__ fo = a.feat + (int64_t)s * a.feat_stride;
    const int nit = (int)((tend - t0 + FR - 1) / FR);
    const unsigned char* grow = v.ws + ((int64_t)sl * v.iters_per_stream + t0 / FR) * (int64_t)kRowBytes;

    // Two row buffers: the current iteration's and, for the flux of its first frame, the previous iteration's (the
    // .y components of its last pack), which is only replaced by the NEXT row once the bin loop has read it.  The row
    // before the run's first one comes from the workspace as well (none before the stream's first frame, whose flux is
    // not an output: the buffer then holds zeros).
    __syncwarp();  // both buffers are behind every lane
    if (t0 == 0) {
      float2* z = reinterpret_cast<float2*>(rows + kRowBytes);
      for (int i = lane; i < (int)(kRowBytes / sizeof(float2)); i += 32) z[i] = make_float2(0.f, 0.f);
      __syncwarp();
    }
    if (lane == 0) {
      if (t0 > 0) {
        mbar_expect_tx(bar + 1, kRowBytes);
        tma_load_1d(rows + kRowBytes, grow - kRowBytes, kRowBytes, bar + 1);
      }
      mbar_expect_tx(bar, kRowBytes);
      tma_load_1d(rows, grow, kRowBytes, bar);
    }
    if (t0 > 0) {
      mbar_wait(bar + 1, ph[1]);
      ph[1] ^= 1u;
    }

    for (int it = 0; it < nit; ++it) {
      const int64_t tf = t0 + (int64_t)FR * it;
      const bool more = it + 1 < nit;
      const unsigned b = (unsigned)it & 1u;
      if (b == 0) {
        mbar_wait(bar, ph[0]);
        ph[0] ^= 1u;
      } else if (it > 0) {
        mbar_wait(bar + 1, ph[1]);
        ph[1] ^= 1u;
      }
      const float2* prevrow = reinterpret_cast<const float2*>(rows + (b ^ 1u) * kRowBytes) + (PK - 1) * PROW;
      const float2* buf = reinterpret_cast<const float2*>(rows + b * kRowBytes);
#pragma unroll
      for (int p = 0; p < PK; ++p) {
        const float2* row = buf + p * PROW;
        const float* rowf = reinterpret_cast<const float*>(row);
        BinAcc3 ac;
        acc_init(ac, priv + lane);
#pragma unroll
        for (int q = 0; q < BPL / 4; ++q) {
          const int tc = spos(BPL * lane + 4 * q);
          const float4 xv = *reinterpret_cast<const float4*>(s_xtab + tc);
          const float4 lv = *reinterpret_cast<const float4*>(s_wlo + tc);
          const float4 hv = *reinterpret_cast<const float4*>(s_whi + tc);
          const float4 m01 = *reinterpret_cast<const float4*>(row + ppos(BPL * lane + 4 * q));
          const float4 m23 = *reinterpret_cast<const float4*>(row + ppos(BPL * lane + 4 * q + 2));
          float pv[4];
          if (p > 0) {  // the frame before this pack's first one is the previous pack's second one
            const float4 r01 = *reinterpret_cast<const float4*>(row - PROW + ppos(BPL * lane + 4 * q));
            const float4 r23 = *reinterpret_cast<const float4*>(row - PROW + ppos(BPL * lane + 4 * q + 2));
            pv[0] = r01.y, pv[1] = r01.w, pv[2] = r23.y, pv[3] = r23.w;
          } else {
            const float4 r01 = *reinterpret_cast<const float4*>(prevrow + ppos(BPL * lane + 4 * q));
            const float4 r23 = *reinterpret_cast<const float4*>(prevrow + ppos(BPL * lane + 4 * q + 2));
            pv[0] = r01.y, pv[1] = r01.w, pv[2] = r23.y, pv[3] = r23.w;
          }
          const float2 mm[4] = {make_float2(m01.x, m01.y), make_float2(m01.z, m01.w), make_float2(m23.x, m23.y),
                                make_float2(m23.z, m23.w)};
          const float xx[4] = {xv.x, xv.y, xv.z, xv.w}, ll[4] = {lv.x, lv.y, lv.z, lv.w}, hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
          for (int u = 0; u < 4; ++u)
            bin_step3(ac, true, 4 * q + u, (fmask >> (4 * q + u)) & 1u, mm[u], pv[u], xx[u], ll[u], hh[u]);
        }
        if (lane == 31) {  // Nyquist bin
          const float2 mq = row[nyq];
          const float pv = p == 0 ? prevrow[nyq].y : row[nyq - PROW].y;
          bin_step3(ac, true, BPL, (fmask >> BPL) & 1u, mq, pv, s_xtab[spos(M)], s_wlo[spos(M)], s_whi[spos(M)]);
        }
        ac.pp[0] = pk::add(ac.pend, ac.mlo);
        ac.pp[32] = ac.mhi;
        if (p == 0) {  // the previous row has been read: the next one may take its place (it has the rest of the
                       // iteration -- reductions, rolloff, mel, DCT -- to arrive)
          __syncwarp();
          if (more && lane == 0) {
            mbar_expect_tx(bar + (b ^ 1u), kRowBytes);
            tma_load_1d(rows + (b ^ 1u) * kRowBytes, grow + (int64_t)(it + 1) * kRowBytes, kRowBytes, bar + (b ^ 1u));
          }
        }

        // ---- the frames' sums: reductions over the warp, both frames at once ----
        const int64_t ta = tf + 2 * p, tb = ta + 1;
        const bool oka = ta < tend, okb = tb < tend;
        float2 pre = ac.seg;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float2 up = make_float2(__shfl_up_sync(kFull3, pre.x, o), __shfl_up_sync(kFull3, pre.y, o));
          if (lane >= o) pre = pk::add(pre, up);
        }
        const float2 etot = make_float2(__shfl_sync(kFull3, pre.x, 31), __shfl_sync(kFull3, pre.y, 31));
        const float2 plow = make_float2(__shfl_sync(kFull3, pre.x, 7), __shfl_sync(kFull3, pre.y, 7));  // bins < B / 4
        const float2 sm = warp_sum3(ac.s0);
        const float2 skm = warp_sum3(pk::fma(ac.s0, k0f, ac.s1));
        const float2 kc = make_float2(sm.x > 0.f ? __fdividef(skm.x, sm.x) : 0.f, sm.y > 0.f ? __fdividef(skm.y, sm.y) : 0.f);
        const float2 dk = make_float2(k0f - kc.x, k0f - kc.y);
        const float2 bw = warp_sum3(__ffma2_rn(__fmul2_rn(dk, dk), ac.s0, __ffma2_rn(pk::scale(dk, 2.f), ac.s1, ac.s2)));
        float2 slg = warp_sum3(ac.sl), sxy = warp_sum3(ac.sxy);
        const float2 fl = warp_sum3(ac.fl);
        const float mxa = warp_max3(ac.mxa), mxb = warp_max3(ac.mxb);
        const float2 ri = row[M + 2];
        // frames for the float64 re-evaluation (spectral_exact.cu), as in stft_v3_kernel
        const float clog = kLogTau * (float)B * sqrt_fast3((float)B) / kEta;
        bool xa = !(ri.x * sqrt_fast3(etot.x) <= clog) || !(ri.x < 1.f / kTinyMag) || !(etot.x > 0.f);
        bool xb = !(ri.y * sqrt_fast3(etot.y) <= clog) || !(ri.y < 1.f / kTinyMag) || !(etot.y > 0.f);
        const float irb = 2.f * kEta * rsqrt_fast3((float)B);
        int rka = rolloff_bin<G>(rowf, pre.x, ac.seg.x, etot.x, irb * sm.x * rsqrt_fast3(fmaxf(etot.x, 1e-36f)) + 1e-7f, lane);
        int rkb = rolloff_bin<G>(rowf + 1, pre.y, ac.seg.y, etot.y, irb * sm.y * rsqrt_fast3(fmaxf(etot.y, 1e-36f)) + 1e-7f, lane);

        __syncwarp();  // private mel slots visible
        if (a.mfcc_on) {
          float2 dens = make_float2(FLT_MAX, FLT_MAX);
          for (int f = lane; f < a.n_mel; f += 32) {
            float2 acc = make_float2(0.f, 0.f);
            for (int i = 0; i < ncontrib; ++i) acc = pk::add(acc, wbf2[s_moff[i * kMaxMel + f]]);
            macc[f] = make_float2(acc.x > 0.f ? __logf(acc.x) : -23.025850929940457f,
                                  acc.y > 0.f ? __logf(acc.y) : -23.025850929940457f);  // ln(1e-10)
            const float iw = s_invw[f];
            if (iw > 0.f) dens = make_float2(fminf(dens.x, acc.x * iw), fminf(dens.y, acc.y * iw));
          }
          const float lim = kMelRatio / (float)B;
          xa = xa || __any_sync(kFull3, !(dens.x >= lim * etot.x));
          xb = xb || __any_sync(kFull3, !(dens.y >= lim * etot.y));
          __syncwarp();
          const int nmp = a.n_mel | 1;
          for (int c0 = 0; c0 < a.n_mfcc; c0 += 16) {
            const int c = c0 + (lane >> 1);
            float2 acc = make_float2(0.f, 0.f);
            if (c < a.n_mfcc)
              for (int f = lane & 1; f < a.n_mel; f += 2) acc = pk::fma(macc[f], s_dct[c * nmp + f], acc);
            acc = pk::add(acc, make_float2(__shfl_xor_sync(kFull3, acc.x, 1), __shfl_xor_sync(kFull3, acc.y, 1)));
            if (c < a.n_mfcc && !(lane & 1)) {
              const float lf = s_lift[c];
              if (oka) fo[a.o_mfcc + ta * a.n_mfcc + c] = (double)(acc.x * lf);
              if (okb) fo[a.o_mfcc + tb * a.n_mfcc + c] = (double)(acc.y * lf);
            }
          }
        }
        if (lane == 0) {  // park the raw sums; finished in FP64 one frame per lane at the end of the 32-frame segment
          const int slot = (int)(ta - t0) & (kRun3 - 1);
          float4* rs = reinterpret_cast<float4*>(rawsum + slot * kRaw3);
          if (xa) rka |= kExactBit;
          if (xb) rkb |= kExactBit;
          rs[0] = make_float4(sm.x, kc.x, etot.x, __int_as_float(rka));
          rs[1] = make_float4(bw.x, slg.x, sxy.x, mxa);
          rs[2] = make_float4(fl.x, plow.x, 0.f, 0.f);
          rs[3] = make_float4(sm.y, kc.y, etot.y, __int_as_float(rkb));
          rs[4] = make_float4(bw.y, slg.y, sxy.y, mxb);
          rs[5] = make_float4(fl.y, plow.y, 0.f, 0.f);
        }
        __syncwarp();  // private slots / macc / parked sums: reused by the next pack, read by the finishing below
      }

      if (((FR * (it + 1)) & (kRun3 - 1)) == 0 || !more) {
        const int seg = (FR * it) / kRun3;
        const int64_t t = t0 + (int64_t)kRun3 * seg + lane;
        if (t < tend) {
          const float4* rs4 = reinterpret_cast<const float4*>(rawsum + lane * kRaw3);
          const float4 r0 = rs4[0], r1 = rs4[1], r2 = rs4[2];
          const float sm = r0.x, kc = r0.y, etot = r0.z, bw = r1.x, slg = r1.y, sxy = r1.z, mx = r1.w, fl = r2.x, plow = r2.y;
          const int rkx = __float_as_int(r0.w), rk = rkx & ~kExactBit;
          if ((rkx & kExactBit) && a.xlist) {  // listed for spectral_exact.cu, which overwrites what is stored below
            int* lst = a.xlist + (int64_t)s * a.xlist_stride;
            lst[1 + atomicAdd(lst, 1)] = (int)t;
          }
          const double fs = a.freq_scale;
          const double dsm = (double)sm;
          fo[a.o_centroid + t] = (double)kc * fs;
          fo[a.o_rolloff + t] = etot > 0.f ? (double)rk * fs : 0.0;
          fo[a.o_bandwidth + t] = sm > 0.f ? sqrt((double)bw / dsm) * fs : 0.0;
          double flat = 0.0;
          {
            const double gm = exp2((double)slg / (double)B);
            const double am = dsm / (double)B;
            if (am > 1e-10) {
              flat = gm / am;
              if (flat > 1.0) flat = 1.0;
            }
          }
          fo[a.o_flatness + t] = flat;
          const double rms = sqrt((double)etot / (double)B);
          fo[a.o_crest + t] = rms > 0.0 ? (double)mx / rms : 0.0;
          double slope = 0.0;
          if (a.slope_on) {  // sum x = 0 for the centred abscissae (spectral_slope.go:42-64)
            const double LG = 0.30102999566398120;  // log10(2)
            const double n = a.slope_ntot;
            if (n >= 2.0 && a.slope_xxtot != 0.0) slope = LG * (double)sxy / a.slope_xxtot;
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
  }
}

// ---- workspace: one per (device, stream), grown on demand, kept for the life of the process ---------------------------
struct Ws5 {
  void* p = nullptr;
  size_t bytes = 0;
};
std::mutex g_ws_mu;
std::map<std::pair<int, cudaStream_t>, Ws5> g_ws;

int ws_get(int dev, cudaStream_t st, size_t bytes, unsigned char** out) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  Ws5& w = g_ws[std::make_pair(dev, st)];
  if (w.bytes < bytes) {
    if (w.p) {
      SONAR_CUDA(cudaStreamSynchronize(st));  // the old block may still be in use by this stream's last launch
      SONAR_CUDA(cudaFree(w.p));
      w.p = nullptr;
      w.bytes = 0;
    }
    if (cudaMalloc(&w.p, bytes) != cudaSuccess) {  // no room for the workspace: the caller falls back
      cudaGetLastError();
      w.p = nullptr;
      return SONAR_ERR_UNSUPPORTED;
    }
    w.bytes = bytes;
  }
  *out = static_cast<unsigned char*>(w.p);
  return SONAR_OK;
}

template <class G>
int mel_slots(const FpPlan& plan) {  // lane-private slots a lane needs: the regions its bins touch + 2
  int mx = 2;
  auto region_of = [&](int k) {
    int r = 0;
    while (k >= plan.h_regions[r].next_b) ++r;
    return r;
  };
  for (int j = 0; j < 32; ++j) {
    const int first = region_of(G::BPL * j), last = region_of(j == 31 ? G::B - 1 : G::BPL * j + G::BPL - 1);
    mx = std::max(mx, last - first + 2);
  }
  return mx;
}

template <int LOGN, int HR>
int v5_launch(const FpPlan& plan, StftArgs& a, cudaStream_t st) {
  using G = V3G<LOGN, HR>;
  constexpr size_t kSmemMax = 227 * 1024;
  V5Args v;
  v.n_slots = mel_slots<G>(plan);
  int warps = kSW5;
  while (warps > 4 && v5s_layout<G>(a.n_mel, a.n_mfcc, v.n_slots, warps).total + 1024 > kSmemMax) --warps;
  const V5SSmem LS = v5s_layout<G>(a.n_mel, a.n_mfcc, v.n_slots, warps);
  const V5TSmem LT = v5t_layout<G>();
  if (LS.total + 1024 > kSmemMax) return SONAR_ERR_UNSUPPORTED;
  v.iters_per_stream = (a.T + G::FR - 1) / G::FR;
  v.runs_per_stream = (int)((a.T + kRunFrames5 - 1) / kRunFrames5);
  const size_t per_stream = (size_t)v.iters_per_stream * row_bytes5<G>();
  static const size_t cap = std::getenv("SONAR_STFT_WS_MB") ? (size_t)std::atoll(std::getenv("SONAR_STFT_WS_MB")) << 20
                                                           : (size_t)8 << 30;
  int group = (int)std::max<size_t>(1, cap / std::max<size_t>(per_stream, 1));
  if (group > a.n_streams) group = a.n_streams;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // Both kernels are persistent, one CTA per SM with a static share of the runs: a CTA that finds its SM taken by a kernel
  // of another stream (the alignment branch of the pair pipeline) starts late and the whole launch waits for it.  Leaving
  // a few SMs unclaimed costs their share of the throughput and keeps the launch off that wait (tl_stft_sm_reserve, common.h).
  {
    const char* e = std::getenv("SONAR_STFT_SM_RESERVE");  // diagnostic override (read per launch: scripts/sm_reserve_ab.py)
    const int reserve = e ? std::atoi(e) : tl_stft_sm_reserve;
    if (reserve > 0) sms = std::max(sms / 2, sms - reserve);
  }
  unsigned char* ws = nullptr;
  int rc = ws_get(dev, st, per_stream * (size_t)group, &ws);
  while (rc == SONAR_ERR_UNSUPPORTED && group > 1) {  // smaller groups of streams, more launches
    group = (group + 1) / 2;
    rc = ws_get(dev, st, per_stream * (size_t)group, &ws);
  }
  if (rc) return rc;
  v.ws = ws;
  SONAR_CUDA(cudaFuncSetAttribute(stft_v5_transform_kernel<LOGN, HR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LT.total));
  SONAR_CUDA(cudaFuncSetAttribute(stft_v5_scan_kernel<LOGN, HR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LS.total));
  for (int s0 = 0; s0 < a.n_streams; s0 += group) {
    v.s0 = s0;
    v.ns = std::min(group, a.n_streams - s0);
    v.total_runs = (int64_t)v.runs_per_stream * v.ns;
    int64_t ct = (v.total_runs + kTW5 - 1) / kTW5, cs = (v.total_runs + warps - 1) / warps;
    if (ct > sms) ct = sms;
    if (cs > sms) cs = sms;
    prof_begin("stft_features_kernel", st);  // the pair is timed as one unit: the roofline's bytes belong to both
    stft_v5_transform_kernel<LOGN, HR><<<(unsigned)ct, kTW5 * 32, LT.total, st>>>(a, v);
    stft_v5_scan_kernel<LOGN, HR><<<(unsigned)cs, warps * 32, LS.total, st>>>(a, v);
    prof_end();
    prof_count_launch();  // two kernels behind one timing pair
    SONAR_CUDA(cudaGetLastError());
  }
  return SONAR_OK;
}

}  // namespace

thread_local int tl_stft_sm_reserve = 0;

void stft_workspace_release(int device, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  auto it = g_ws.find(std::make_pair(device, st));
  if (it == g_ws.end()) return;
  if (it->second.p) cudaFree(it->second.p);  // the caller has synchronised the device
  g_ws.erase(it);
}

// the kernel pair serves the geometries of stft_v3.cu (its eligibility test applies)
int launch_stft_v5(const FpPlan& plan, StftArgs& a, cudaStream_t st) {
  if (plan.N == 1024) return v5_launch<10, 8>(plan, a, st);
  return v5_launch<9, 5>(plan, a, st);
}

}  // namespace sonar
