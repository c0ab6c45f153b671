// Fused framed-STFT + MFCC + spectral-descriptor kernel, third generation (N = 1024 or 512, hop a multiple of 32;
// FP32 arithmetic, f64 I/O).  Same outputs as stft_v2.cu / stft_features.cu, replacing
//   analyzers.ComputeSTFTWithWindow        fingerprint/analyzers/spectral.go:385-545
//   spectral.MFCC.Compute                  algorithms/spectral/mfcc.go:113-164
//   MelScale.ApplyFilterBank               algorithms/spectral/mel_scale.go:89-105
//   centroid/rolloff/bandwidth/flatness/crest/slope/flux   algorithms/spectral/spectral_*.go
//   low/high band energy ratios            fingerprint/extractors/speech.go:436-456
//
// The second generation was bound by shared-memory wavefronts (575 per frame at 74 % of the pipe: three tile
// exchanges of the 8 x 8 x 8 FFT, the ring, five table / row streams of the per-bin scan), not by FP32 issue or HBM
// (profiles/r01_stft_v2_kernel_ncu_full.md; one SM moves 1 wavefront per clock but issues 4 FP32 warp instructions).
// This generation is built around the wavefront count (scripts/proto/stft_v3_dataflow.py is the numpy model of it):
//   * a warp transforms TWO consecutive frames at once as ONE complex FFT of c = a_w + i b_w (N = 512: two such
//     packs, four frames).  Lane l owns the samples n = l + 32 j, so consecutive frames share their samples IN
//     REGISTERS (the hop is a whole number of rows j): the sample ring costs no shared memory at all, and every new
//     sample is loaded from HBM exactly once, converted to FP32 once;
//   * 1024 = 32 x 32 (512 = 16 x 32): pass 1 is a radix-32 (radix-16) transform in registers, ONE transposition through
//     a padded tile, pass 2 a radix-32 transform in registers that leaves Z[k1 + 32 k2] in lane k1;
//   * the Hermitian split X_a = Z[k] + conj Z[N-k], X_b = (Z[k] - conj Z[N-k]) / i needs the partner bin from lane
//     (32 - k1) & 31: 32 warp shuffles instead of a second trip through shared memory;
//   * all complex arithmetic runs on the Blackwell packed-FP32 pipe instructions (FADD2 / FMUL2 / FFMA2, fft_packed.cuh):
//     a radix-32 transform is 228 issue slots instead of ~510, and the scan processes (frame a, frame b) pairs;
//   * magnitudes go to XOR-swizzled rows (conflict free for the stride-1 writes by bin AND for the 16-byte reads of
//     the scan, no padding), the per-bin scan walks BOTH frames of a pack in one pass: the mel weights / slope
//     abscissae are read once per two frames and frame b's predecessor is frame a (registers); the predecessor of the
//     first frame of an iteration is kept in registers from the previous iteration.
#include "stft_fused.cuh"

namespace sonar {
namespace {

template <int LOGN, int HR>
__global__ void __launch_bounds__(kW3 * 32, 1) stft_v3_kernel(const StftArgs a) {
  using G = V3G<LOGN, HR>;
  constexpr int N = G::N, M = G::M, B = G::B, J = G::J, FR = G::FR, PK = G::PK, RR = G::RR, NEW = G::NEW, BPL = G::BPL,
                KSTR = G::KSTR, ROW = G::ROW, PROW = G::PROW, H = G::H;
  extern __shared__ __align__(128) unsigned char smem[];
  const V3Smem L = v3_layout<G>(a.n_mel, a.n_mfcc);
  float2* s_tw = reinterpret_cast<float2*>(smem + L.tw);
  float* s_win = reinterpret_cast<float*>(smem + L.win);
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

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wb = smem + L.warp0 + (size_t)warp * L.per_warp;
  float2* wbf2 = reinterpret_cast<float2*>(wb);
  float2* tile = reinterpret_cast<float2*>(wb + L.w_tile);
  float2* priv = tile;  // lane-private mel slots [slot][lane], (frame a, frame b)
  float2* mag = reinterpret_cast<float2*>(wb + L.w_mag);
  float* rawsum = reinterpret_cast<float*>(wb + L.w_raw);
  float2* macc = reinterpret_cast<float2*>(wb + L.w_macc);
  constexpr int kZeroSlot = kMaxMel + 2;  // macc[kZeroSlot] stays 0: padding target of the combine table

  // ---- tables, once per CTA ------------------------------------------------------------------------
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
  for (int k = threadIdx.x; k < ROW; k += blockDim.x) {
    s_xtab[k] = 0.f;
    s_wlo[k] = 0.f;
    s_whi[k] = 0.f;
  }
  if (threadIdx.x < 32) s_fmask[threadIdx.x] = 0u;
  if (threadIdx.x < kMaxMel) s_invw[threadIdx.x] = a.mel_invw[threadIdx.x];
  if (threadIdx.x == 0) s_ncontrib = 0;
  if (lane == 0) macc[kZeroSlot] = make_float2(0.f, 0.f);
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
    // bit j of a lane's mask: bin BPL lane + j opens a new region (a lane's first bin starts inside its region,
    // nothing to close; the Nyquist bin is lane 31's extra one, bit BPL)
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
  // combine table: the private slots (float2 offsets from the warp's base) that hold a part of filter f (absolute slot
  // f + 1; lane j's slot q is r0[j] - 1 + q), padded with the zero slot to a uniform count
  const unsigned short zero_off = (unsigned short)((macc + kZeroSlot) - wbf2);
  const unsigned short priv_off = (unsigned short)(priv - wbf2);
  for (int f = threadIdx.x; f < kMaxMel; f += blockDim.x) {
    int cnt = 0;
    if (f < a.n_mel) {
      for (int j = 0; j < 32; ++j) {
        int rl = 0;
        const int kl = (j == 31) ? B - 1 : BPL * j + BPL - 1;
        while (kl >= a.regions[rl].next_b) ++rl;
        const int first = s_r0[j] - 1, last = rl;  // slots first .. last are written by lane j
        if (f + 1 >= first && f + 1 <= last && cnt < kMaxContrib3)
          s_moff[(cnt++) * kMaxMel + f] = (unsigned short)(priv_off + (f + 1 - first) * 32 + j);
      }
      atomicMax(&s_ncontrib, cnt);
    }
    for (int i = cnt; i < kMaxContrib3; ++i) s_moff[i * kMaxMel + f] = zero_off;
  }
  __syncthreads();
  const int ncontrib = s_ncontrib;

  // ---- per-lane constants ----------------------------------------------------------------------------
  const int k1 = PK == 1 ? lane : (lane & 15);                               // pass-2 row of this lane
  const int src = PK == 1 ? ((32 - lane) & 31) : ((lane & 16) | ((16 - k1) & 15));  // lane of the partner bins
  const bool k1zero = k1 == 0;
  const int64_t T = a.T;
  const unsigned fmask = s_fmask[lane];
  const int nyq = ppos(M);
  const float k0f = (float)(BPL * lane);
  // pass 2 stores bin k1 + KSTR k2 of the lane's pair row: ppos() of it is woff[k2 % NV] + KSTR k2 (the XOR of the swizzle
  // only depends on k2 mod NV), so the 16 stores need no address arithmetic
  constexpr int NV = PK == 1 ? 4 : 8;
  int woff[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) woff[v] = ppos(k1 + KSTR * v) - KSTR * v;

  // (V3_DYNAMIC_RUNS: runs handed out through a counter, so that a CTA that starts late -- its SM busy with another
  // stream's kernel -- simply takes fewer of them)
#ifndef V3_DYNAMIC_RUNS  // static striding measured 2.5 % faster at 64 streams (7.36 vs 7.57 ms); the counter only pays
                         // off when CTAs start late, which the walk -> STFT order avoids
  for (int64_t run = (int64_t)blockIdx.x * kW3 + warp; run < a.total_runs; run += (int64_t)gridDim.x * kW3) {
#else
  for (;;) {
    int64_t run = 0;
    if (lane == 0) run = (int64_t)atomicAdd(a.work_counter, 1u);
    run = __shfl_sync(kFull3, run, 0);
    if (run >= a.total_runs) break;
#endif
    const int s = (int)(run / a.runs_per_stream);
    const int64_t t0 = (run % a.runs_per_stream) * (int64_t)kRunOut3;
    const int64_t tend = (t0 + kRunOut3 < T) ? t0 + kRunOut3 : T;
    const double* __restrict__ x = a.pcm + (int64_t)s * a.stride;
    double* __restrict__ fo = a.feat + (int64_t)s * a.feat_stride;
    const int64_t first = t0 - 1;                       // the run's first frame only warms the flux up (-1: none)
    const int nfr = (int)(tend - first);                // frames first .. tend - 1
    const int nit = (nfr + FR - 1) / FR;

    // ---- ring: RR rows of the first iteration; samples outside [0, n) read as silence --------------------
    // rows [jlo, jhi) of the lane's column lie inside the stream: two 32-bit compares per load instead of 64-bit ones
    float ring[RR];
    const double* __restrict__ xl = x + (first * H + lane);  // row j of the first iteration is xl[32 j]
    int64_t rows_left;                                        // rows from xl's row 0 to the end of the stream
    {
      const int64_t g0 = first * H + lane;
      rows_left = (a.n - g0 + 31) >> 5;
      const int jlo = g0 < 0 ? (int)((-g0 + 31) >> 5) : 0;
      const int jhi = rows_left < RR ? (int)(rows_left < 0 ? 0 : rows_left) : RR;
#pragma unroll
      for (int j = 0; j < RR; ++j) ring[j] = (j >= jlo && j < jhi) ? (float)__ldg(xl + 32 * j) : 0.f;
    }

    for (int it = 0; it < nit; ++it) {
      const int64_t tf = first + (int64_t)FR * it;  // first frame of the iteration
      // ================= pass 1: radix-J over the lane's own samples, both frames of a pack at once ==========
#pragma unroll
      for (int p = 0; p < PK; ++p) {
        float2 c[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const float w = s_win[lane + 32 * j];
          c[j] = make_float2(ring[j + 2 * p * HR] * w, ring[j + (2 * p + 1) * HR] * w);
        }
        pk::Fft<J>::run(c);
        tw_apply<J, 1>(c, s_tw + lane);
        float2* tp = tile + (p * J) * kTileRow + lane;
#pragma unroll
        for (int q = 0; q < J; ++q) tp[q * kTileRow] = c[q];
      }
      __syncwarp();
      // ================= pass 2: radix-32 over the lanes of pass 1 ==========================================
      float fla = 0.f;  // this lane's part of the first frame's flux
      float2 rinv = make_float2(0.f, 0.f);  // this lane's part of sum 1 / |X_k| of its pack's two frames
      {
        float2 z[32];
        const float4* rp = reinterpret_cast<const float4*>(tile + lane * kTileRow);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 f = rp[i];
          z[2 * i] = make_float2(f.x, f.y);
          z[2 * i + 1] = make_float2(f.z, f.w);
        }
        pk::Fft<32>::run(z);
        // Hermitian split + magnitudes: bins k = k1 + KSTR k2, k2 < 16; partner Z[N - k] = register 31 - k2 of lane
        // `src` (k1 == 0: register (32 - k2) & 31 of this lane):  X_a = Z + conj P,  X_b = (Z - conj P) / i
        // Flux of the iteration's FIRST frame: its predecessor is the previous iteration's last frame, whose magnitudes
        // are still in the last pack's row (component .y) -- read bin k's old value right before the row is overwritten
        // (the lanes of the last pack store to that row later in program order; shared memory is in order per warp).
        float2* row = mag + (PK == 1 ? 0 : (lane >> 4)) * PROW;
        const float* oldb = reinterpret_cast<const float*>(mag + (PK - 1) * PROW) + 1;
        const bool first_pack = PK == 1 || lane < 16;
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
          const float2 mine = k1zero ? z[(32 - k2) & 31] : z[31 - k2];
          const float2 pz = make_float2(__shfl_sync(kFull3, mine.x, src), __shfl_sync(kFull3, mine.y, src));
          const float2 zz = z[k2];
          const float2 xa = __fadd2_rn(zz, make_float2(pz.x, -pz.y));
          const float2 xb = __fadd2_rn(make_float2(zz.y, -zz.x), make_float2(pz.y, pz.x));
          const float2 qa = __fmul2_rn(xa, xa), qb = __fmul2_rn(xb, xb);
          const int e = woff[k2 % NV] + KSTR * k2;
          // |X| = q rsqrt(q): one MUFU gives the magnitude AND its reciprocal (the weak-bin test)
          const float2 q = make_float2(qa.x + qa.y, qb.x + qb.y);
          // (the 1e-36 keeps q = 0 finite: |X| = 0, 1 / |X| = 1e18; one packed add for both frames)
          const float2 qt = __fadd2_rn(q, make_float2(1e-36f, 1e-36f));
          const float2 ri = make_float2(rsqrt_fast3(qt.x), rsqrt_fast3(qt.y));
          const float2 m = __fmul2_rn(q, ri);
#ifndef V3_NO_WEAK
          rinv = __fadd2_rn(rinv, ri);
#endif
          const float d = first_pack ? fmaxf(m.x - oldb[2 * e], 0.f) : 0.f;
          fla = fmaf(d, d, fla);
          row[e] = m;
        }
        if (k1zero) {  // Z[M] pairs with itself
          const float2 m = make_float2(fabsf(2.f * z[16].x), fabsf(2.f * z[16].y));
          rinv = __fadd2_rn(rinv, make_float2(__fdividef(1.f, fmaxf(m.x, 1e-18f)), __fdividef(1.f, fmaxf(m.y, 1e-18f))));
          const float d = first_pack ? fmaxf(m.x - oldb[2 * nyq], 0.f) : 0.f;
          fla = fmaf(d, d, fla);
          row[nyq] = m;
        }
#pragma unroll
        for (int o = (PK == 1 ? 16 : 8); o >= 1; o >>= 1)  // over the lanes of the pack
          rinv = __fadd2_rn(rinv, make_float2(__shfl_xor_sync(kFull3, rinv.x, o), __shfl_xor_sync(kFull3, rinv.y, o)));
      }
      // the next iteration's NEW rows start their trip from HBM now and join the ring at the end of the iteration
      double nx[NEW];
      const bool more = it + 1 < nit;
      if (more) {
        const int r0 = FR * HR * (it + 1) + (RR - NEW);  // first new row, counted from xl's row 0
        const double* __restrict__ src_p = xl + 32 * (int64_t)r0;
        const int64_t left = rows_left - r0;
        const int jhi = left < NEW ? (int)(left < 0 ? 0 : left) : NEW;
        const int jlo = (t0 == 0 && it == 0) ? ((H - lane + 31) >> 5) - r0 : 0;  // rows before the stream's first sample
#pragma unroll
        for (int j = 0; j < NEW; ++j) nx[j] = (j >= jlo && j < jhi) ? __ldg(src_p + 32 * j) : 0.0;
      }
      __syncwarp();  // magnitude rows complete; the tile becomes the private mel slots

      // ================= scan: BPL contiguous bins per lane, both frames of a pack in one pass ===============
#pragma unroll
      for (int p = 0; p < PK; ++p) {
        const float2* row = mag + p * PROW;
        const float* rowf = reinterpret_cast<const float*>(row);
        BinAcc3 ac;
        acc_init(ac, priv + lane);
#pragma unroll
        for (int q = 0; q < BPL / 4; ++q) {  // four bins per step: two 16-byte pair chunks, one chunk of each table
          const int tc = spos(BPL * lane + 4 * q);
          const float4 xv = *reinterpret_cast<const float4*>(s_xtab + tc);
          const float4 lv = *reinterpret_cast<const float4*>(s_wlo + tc);
          const float4 hv = *reinterpret_cast<const float4*>(s_whi + tc);
          const float4 m01 = *reinterpret_cast<const float4*>(row + ppos(BPL * lane + 4 * q));
          const float4 m23 = *reinterpret_cast<const float4*>(row + ppos(BPL * lane + 4 * q + 2));
          float pv[4] = {0.f, 0.f, 0.f, 0.f};
          if (p > 0) {  // the frame before this pack's first one is the previous pack's second one
            const float4 r01 = *reinterpret_cast<const float4*>(row - PROW + ppos(BPL * lane + 4 * q));
            const float4 r23 = *reinterpret_cast<const float4*>(row - PROW + ppos(BPL * lane + 4 * q + 2));
            pv[0] = r01.y, pv[1] = r01.w, pv[2] = r23.y, pv[3] = r23.w;
          }
          const float2 mm[4] = {make_float2(m01.x, m01.y), make_float2(m01.z, m01.w), make_float2(m23.x, m23.y),
                                make_float2(m23.z, m23.w)};
          const float xx[4] = {xv.x, xv.y, xv.z, xv.w}, ll[4] = {lv.x, lv.y, lv.z, lv.w}, hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
          for (int u = 0; u < 4; ++u)
            bin_step3(ac, p > 0, 4 * q + u, (fmask >> (4 * q + u)) & 1u, mm[u], pv[u], xx[u], ll[u], hh[u]);
        }
        if (lane == 31) {  // Nyquist bin
          const float2 mq = row[nyq];
          const float pv = p == 0 ? 0.f : row[nyq - PROW].y;
          bin_step3(ac, p > 0, BPL, (fmask >> BPL) & 1u, mq, pv, s_xtab[spos(M)], s_wlo[spos(M)], s_whi[spos(M)]);
        }
        if (p == 0) ac.fl.x += fla;  // the first frame's flux was taken in pass 2
        ac.pp[0] = pk::add(ac.pend, ac.mlo);
        ac.pp[32] = ac.mhi;

        // ---- the frames' sums: reductions over the warp, both frames at once --------------------------------
        const int64_t ta = tf + 2 * p, tb = ta + 1;
        const bool oka = ta >= t0 && ta < tend, okb = tb >= t0 && tb < tend;
        float2 pre = ac.seg;  // inclusive prefix of the lanes' energies (bins ascend with the lane)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float2 up = make_float2(__shfl_up_sync(kFull3, pre.x, o), __shfl_up_sync(kFull3, pre.y, o));
          if (lane >= o) pre = pk::add(pre, up);
        }
        const float2 etot = make_float2(__shfl_sync(kFull3, pre.x, 31), __shfl_sync(kFull3, pre.y, 31));
        const float2 plow = make_float2(__shfl_sync(kFull3, pre.x, 7), __shfl_sync(kFull3, pre.y, 7));  // bins < B / 4
        const float2 sm = warp_sum3(ac.s0);
        const float2 skm = warp_sum3(pk::fma(ac.s0, k0f, ac.s1));
        const float2 kc = make_float2(sm.x > 0.f ? __fdividef(skm.x, sm.x) : 0.f,
                                      sm.y > 0.f ? __fdividef(skm.y, sm.y) : 0.f);  // centroids in bin units
        // sum (k - kc)^2 m over the lane's bins = dk^2 S0 + 2 dk S1 + S2, dk = k0 - kc
        const float2 dk = make_float2(k0f - kc.x, k0f - kc.y);
        const float2 bw = warp_sum3(__ffma2_rn(__fmul2_rn(dk, dk), ac.s0, __ffma2_rn(pk::scale(dk, 2.f), ac.s1, ac.s2)));
        float2 sl = warp_sum3(ac.sl), sxy = warp_sum3(ac.sxy);
        const float2 fl = warp_sum3(ac.fl);
        const float mxa = warp_max3(ac.mxa), mxb = warp_max3(ac.mxb);
        const float2 ri = PK == 1 ? rinv : make_float2(__shfl_sync(kFull3, rinv.x, 16 * p), __shfl_sync(kFull3, rinv.y, 16 * p));
        // float64 re-evaluation wanted (spectral_exact.cu): the mean of ln|X_k| cannot be trusted when
        // kEta * RMS level * mean(1 / |X_k|) exceeds kLogTau; a magnitude near the reference's 1e-10 validity threshold
        // shows as sum 1 / |X_k| >= 1 / kTinyMag; silence; anything not finite (the negated comparisons catch NaN)
        const float clog = kLogTau * (float)B * sqrt_fast3((float)B) / kEta;
#ifdef V3_NO_XFLAGS
        bool xa = false, xb = false;
        (void)clog, (void)ri;
#else
        bool xa = !(ri.x * sqrt_fast3(etot.x) <= clog) || !(ri.x < 1.f / kTinyMag) || !(etot.x > 0.f);
        bool xb = !(ri.y * sqrt_fast3(etot.y) <= clog) || !(ri.y < 1.f / kTinyMag) || !(etot.y > 0.f);
#endif
        // transform error of a cumulative sum of squares: sum 2 m_k e_k with |e_k| <= kEta * RMS level, all aligned
        // (~50 standard deviations of the actual, random, sum) + 1e-7 for the rounding of the window / twiddle tables
        const float irb = 2.f * kEta * rsqrt_fast3((float)B);
        int rka = rolloff_bin<G>(rowf, pre.x, ac.seg.x, etot.x, irb * sm.x * rsqrt_fast3(fmaxf(etot.x, 1e-36f)) + 1e-7f, lane);
        int rkb = rolloff_bin<G>(rowf + 1, pre.y, ac.seg.y, etot.y, irb * sm.y * rsqrt_fast3(fmaxf(etot.y, 1e-36f)) + 1e-7f, lane);

        __syncwarp();  // private mel slots visible
        // ---- ln + DCT-II + lifter (mfcc.go:136-157), both frames ----
        if (a.mfcc_on) {
          float2 dens = make_float2(FLT_MAX, FLT_MAX);  // smallest mean energy density of a mel band
          for (int f = lane; f < a.n_mel; f += 32) {
            float2 v = make_float2(0.f, 0.f);
            for (int i = 0; i < ncontrib; ++i) v = pk::add(v, wbf2[s_moff[i * kMaxMel + f]]);
            macc[f] = make_float2(v.x > 0.f ? __logf(v.x) : -23.025850929940457f,
                                  v.y > 0.f ? __logf(v.y) : -23.025850929940457f);  // ln(1e-10)
#ifndef V3_NO_WEAK
            const float iw = s_invw[f];
            if (iw > 0.f) dens = make_float2(fminf(dens.x, v.x * iw), fminf(dens.y, v.y * iw));
#endif
          }
          // a band whose level is below kMelRatio of the frame's mean level sits in the transform's noise: ln E errs by > 1e-4
          const float lim = kMelRatio / (float)B;
          xa = xa || __any_sync(kFull3, !(dens.x >= lim * etot.x));
          xb = xb || __any_sync(kFull3, !(dens.y >= lim * etot.y));
          __syncwarp();
          // coefficient c by the lane pair (2c, 2c+1): each half sums every other filter
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
        // ---- park the raw sums of the two frames; finished in FP64 one frame per lane at the end of the run ----
        if (lane == 0) {
          const int slot = (int)(ta - first) & (kRun3 - 1);  // position inside the 32-frame segment
          float4* rs = reinterpret_cast<float4*>(rawsum + slot * kRaw3);
          if (xa) rka |= kExactBit;
          if (xb) rkb |= kExactBit;
          rs[0] = make_float4(sm.x, kc.x, etot.x, __int_as_float(rka));
          rs[1] = make_float4(bw.x, sl.x, sxy.x, mxa);
          rs[2] = make_float4(fl.x, plow.x, 0.f, 0.f);
          if (slot + 1 < kRun3) {
            rs[3] = make_float4(sm.y, kc.y, etot.y, __int_as_float(rkb));
            rs[4] = make_float4(bw.y, sl.y, sxy.y, mxb);
            rs[5] = make_float4(fl.y, plow.y, 0.f, 0.f);
          }
        }
        __syncwarp();  // private slots / macc reused by the next pack
      }
      // ---- ring: drop the oldest NEW rows, append the prefetched ones ----------------------------------------
      if (more) {
#pragma unroll
        for (int j = 0; j < RR - NEW; ++j) ring[j] = ring[j + NEW];
#pragma unroll
        for (int j = 0; j < NEW; ++j) ring[RR - NEW + j] = (float)nx[j];
      }
      __syncwarp();  // tile / rows / parked sums visible, reused by the next iteration

      // ---- a 32-frame segment is complete: finish it in FP64, lane i <-> frame first + 32 seg + i ----------
      if (((FR * (it + 1)) & (kRun3 - 1)) == 0 || !more) {
        const int seg = (FR * it) / kRun3;
        const int64_t t = first + (int64_t)kRun3 * seg + lane;
        if (t >= t0 && t < tend) {
          const float4* rs4 = reinterpret_cast<const float4*>(rawsum + lane * kRaw3);
          const float4 r0 = rs4[0], r1 = rs4[1], r2 = rs4[2];
          const float sm = r0.x, kc = r0.y, etot = r0.z, bw = r1.x, sl = r1.y, sxy = r1.z, mx = r1.w, fl = r2.x, plow = r2.y;
          const int rkx = __float_as_int(r0.w), rk = rkx & ~kExactBit;
#ifndef V3_NO_LIST
          if ((rkx & kExactBit) && a.xlist) {  // listed for spectral_exact.cu, which overwrites what is stored below
            int* lst = a.xlist + (int64_t)s * a.xlist_stride;
            lst[1 + atomicAdd(lst, 1)] = (int)t;
          }
#endif
          const double fs = a.freq_scale;
          const double dsm = (double)sm;
          fo[a.o_centroid + t] = (double)kc * fs;
          fo[a.o_rolloff + t] = etot > 0.f ? (double)rk * fs : 0.0;
          fo[a.o_bandwidth + t] = sm > 0.f ? sqrt((double)bw / dsm) * fs : 0.0;
          double flat = 0.0;
          {  // every bin counts: frames with a magnitude near 1e-10 are listed (spectral_flatness.go:31-70)
            const double gm = exp2((double)sl / (double)B);
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
        __syncwarp();  // the parked sums are overwritten by the next segment
      }
    }
  }
}


// =====================================================================================================================
// Fourth generation: the same arithmetic, WARP SPECIALISED.  The third-generation kernel is bound by the dependent-issue
// latency of one warp's instruction stream: 12 warps x 168 registers fill the register file, its time is inversely
// proportional to the resident warps (profiles/r02_stft_v3_ncu.md), and the 168 registers are needed by the transform
// (64 for the radix-32 points + 40 for the sample ring), not by the per-bin scan.  Here a frame pair's work is split
// between TWO warps that run concurrently:
//   * a TRANSFORM warp (sample ring, window, pass 1, transposition, pass 2, Hermitian split, magnitudes) and
//   * a SCAN warp (per-bin scan, reductions, rolloff, mel combine, ln + DCT, float64 finishing),
// handing the magnitude rows over through a double-buffered shared-memory row and two mbarriers per buffer (full /
// empty; one arrival each, by lane 0 after a __syncwarp).  `setmaxnreg` moves registers from the scan warpgroups (96) to
// the transform warpgroups (160): 8 x 32 x (160 + 96) = 65,536, i.e. SIXTEEN resident warps per SM instead of twelve,
// each with half the instruction stream per frame.  The double buffer also removes the flux work-around of the third
// generation (the first frame's flux taken in pass 2 from the row about to be overwritten): the scan warp keeps the
// previous frame's magnitudes of its own bins in registers.
constexpr int kPairs4 = 8;                  // transform / scan warp pairs per CTA (one CTA per SM)
constexpr int kThreads4 = kPairs4 * 64;     // warps 0..7 transform, 8..15 scan (two warpgroups each)
constexpr int kRegsFft4 = 160, kRegsScan4 = 96;

struct V4Smem {
  size_t tw, win, xtab, wlo, whi, fmask, moff, dct, lift, r0, pair0, per_pair, total;
  size_t p_tile, p_mag, p_priv, p_raw, p_macc, mag_bytes;
};
template <class G>
__host__ __device__ inline V4Smem v4_layout(int n_mel, int n_mfcc) {
  V4Smem L;
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
  L.pair0 = o;
  size_t w = 0;
  auto wtake = [&](size_t bytes) {
    size_t r = w;
    w += (bytes + 127) & ~(size_t)127;
    return r;
  };
  L.mag_bytes = (sizeof(float2) * G::PK * G::PROW + 127) & ~(size_t)127;
  L.p_tile = wtake(sizeof(float2) * 32 * kTileRow);
  L.p_mag = wtake(2 * L.mag_bytes);
  L.p_priv = wtake(sizeof(float2) * kSlots3 * 32);
  L.p_raw = wtake(sizeof(float) * kRaw3 * kRun3);
  L.p_macc = wtake(sizeof(float2) * (kMaxMel + 4));
  L.per_pair = w;
  L.total = o + w * kPairs4;
  return L;
}

// ---- transform warp ---------------------------------------------------------------------------------------------------
template <class G>
__device__ __forceinline__ void v4_transform(const StftArgs& a, int pr, int lane, const float2* s_tw,
                                             const float* s_win, float2* tile,
                                             unsigned char* magbuf, size_t mag_bytes,
                                             unsigned long long* bar, unsigned long long* tbar) {
  constexpr int N = G::N, M = G::M, J = G::J, FR = G::FR, PK = G::PK, RR = G::RR, NEW = G::NEW, KSTR = G::KSTR,
                PROW = G::PROW, H = G::H;
  (void)N;
  const int k1 = PK == 1 ? lane : (lane & 15);
  const int src = PK == 1 ? ((32 - lane) & 31) : ((lane & 16) | ((16 - k1) & 15));
  const bool k1zero = k1 == 0;
  const int64_t T = a.T;
  const int nyq = ppos(M);
  constexpr int NV = PK == 1 ? 4 : 8;
  int woff[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) woff[v] = ppos(k1 + KSTR * v) - KSTR * v;
  unsigned gi = 0;  // iterations handed over so far (the scan warp counts the same way)
  unsigned tphase = 0;  // parity of the staging barrier's next completion
  const double* stage = reinterpret_cast<const double*>(tile);  // the tile is idle between pass 2's loads and pass 1

  // lockstep: every warp runs the same number of rounds and of iterations per round (a run past the end of the work, or
  // the frames past the end of a stream's last run, only keep the barrier counts equal: loads read silence, stores are
  // masked by the scan warp)
  const int lock = a.v4_lockstep;
  const int bid = 1 + (pr & 3);
  const int64_t stride_runs = (int64_t)gridDim.x * kPairs4;
  const int64_t run_end = lock ? ((a.total_runs + stride_runs - 1) / stride_runs) * stride_runs : a.total_runs;
  constexpr int kNitFull = (kRunOut3 + 1 + FR - 1) / FR;
  for (int64_t run = (int64_t)blockIdx.x * kPairs4 + pr; run < run_end; run += stride_runs) {
    if (run >= a.total_runs) {  // barrier counts only
      for (int it = 0; it < kNitFull * (lock > 1 ? 2 : 1); ++it) mate_sync(bid);
      continue;
    }
    const int s = (int)(run / a.runs_per_stream);
    const int64_t t0 = (run % a.runs_per_stream) * (int64_t)kRunOut3;
    const int64_t tend = (t0 + kRunOut3 < T) ? t0 + kRunOut3 : T;
    const double* __restrict__ x = a.pcm + (int64_t)s * a.stride;
    const int64_t first = t0 - 1;
    const int nfr = (int)(tend - first);
    const int nit = lock ? kNitFull : (nfr + FR - 1) / FR;

    float ring[RR];
    const double* __restrict__ xl = x + (first * H + lane);
    int64_t rows_left;
    {
      const int64_t g0 = first * H + lane;
      rows_left = (a.n - g0 + 31) >> 5;
      const int jlo = g0 < 0 ? (int)((-g0 + 31) >> 5) : 0;
      const int jhi = rows_left < RR ? (int)(rows_left < 0 ? 0 : rows_left) : RR;
#pragma unroll
      for (int j = 0; j < RR; ++j) ring[j] = (j >= jlo && j < jhi) ? (float)__ldg(xl + 32 * j) : 0.f;
    }

    for (int it = 0; it < nit; ++it, ++gi) {
      if (lock) mate_sync(bid);
      // ---- pass 1 ----
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
      if (lock > 1) mate_sync(bid);
      // ---- pass 2 + Hermitian split + magnitudes into the buffer the scan warp has released ----
      const unsigned b = gi & 1u, use = gi >> 1;
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
        // The tile is idle from here to the next pass 1: the NEW sample rows of the next iteration (one contiguous block
        // of the stream) are staged in it by ONE bulk copy, in flight behind pass 2 and the split.  Rows that leave the
        // stream (its first and last iterations) take the register path below.
        __syncwarp();
        if (more) {
          const int r0 = FR * G::HR * (it + 1) + (RR - NEW);
          const int64_t gs = first * H + 32 * (int64_t)r0;  // first sample of the block
          staged = gs >= 0 && gs + 32 * NEW <= a.n && ((reinterpret_cast<uintptr_t>(x + gs) & 15) == 0);
          if (staged && lane == 0) {
            mbar_expect_tx(tbar, 32 * NEW * sizeof(double));
            tma_load_1d(tile, x + gs, 32 * NEW * sizeof(double), tbar);
          }
        }
        pk::Fft<32>::run(z);
        if (use > 0) mbar_wait(bar + 2 + b, (use - 1) & 1u);
        float2* row = reinterpret_cast<float2*>(magbuf + b * mag_bytes) + (PK == 1 ? 0 : (lane >> 4)) * PROW;
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
          const float2 m = __fmul2_rn(q, ri);
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
        if (k1zero) row[M + 2] = rinv;  // sum 1 / |X_k| of the pack's two frames (a free slot behind the Nyquist bin)
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar + b);
      // ---- the next iteration's new rows ----
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
        const int jlo = (t0 == 0 && it == 0) ? ((H - lane + 31) >> 5) - r0 : 0;
#pragma unroll
        for (int j = 0; j < NEW; ++j) nx[j] = (j >= jlo && j < jhi) ? __ldg(src_p + 32 * j) : 0.0;
#pragma unroll
        for (int j = 0; j < RR - NEW; ++j) ring[j] = ring[j + NEW];
#pragma unroll
        for (int j = 0; j < NEW; ++j) ring[RR - NEW + j] = (float)nx[j];
      }
    }
  }
}

// ---- scan warp --------------------------------------------------------------------------------------------------------
template <class G>
__device__ __forceinline__ void v4_scan(const StftArgs& a, int pr, int lane, const float* s_xtab,
                                        const float* s_wlo, const float* s_whi,
                                        unsigned fmask, const unsigned short* s_moff,
                                        const float* s_dct, const float* s_lift,
                                        const float* s_invw, int ncontrib, float2* wbf2,
                                        const unsigned char* magbuf, size_t mag_bytes,
                                        float2* priv, float* rawsum,
                                        float2* macc, unsigned long long* bar) {
  constexpr int M = G::M, B = G::B, FR = G::FR, PK = G::PK, BPL = G::BPL, PROW = G::PROW;
  const int64_t T = a.T;
  const int nyq = ppos(M);
  const float k0f = (float)(BPL * lane);
  unsigned gi = 0;

  const int lock = a.v4_lockstep;
  const int bid = 5 + (pr & 3);
  const int64_t stride_runs = (int64_t)gridDim.x * kPairs4;
  const int64_t run_end = lock ? ((a.total_runs + stride_runs - 1) / stride_runs) * stride_runs : a.total_runs;
  constexpr int kNitFull = (kRunOut3 + 1 + FR - 1) / FR;
  for (int64_t run = (int64_t)blockIdx.x * kPairs4 + pr; run < run_end; run += stride_runs) {
    if (run >= a.total_runs) {  // barrier counts only
      for (int it = 0; it < kNitFull * (lock > 1 ? 1 + PK : 1); ++it) mate_sync(bid);
      continue;
    }
    const int s = (int)(run / a.runs_per_stream);
    const int64_t t0 = (run % a.runs_per_stream) * (int64_t)kRunOut3;
    const int64_t tend = (t0 + kRunOut3 < T) ? t0 + kRunOut3 : T;
    double* __restrict__ fo = a.feat + (int64_t)s * a.feat_stride;
    const int64_t first = t0 - 1;
    const int nfr = (int)(tend - first);
    const int nit = lock ? kNitFull : (nfr + FR - 1) / FR;
    // |X| of the frame before the iteration's first one, this lane's own bins (the run's first frame only warms the
    // flux up: its predecessor does not matter)
    float prevb[BPL], prevny = 0.f;
#pragma unroll
    for (int j = 0; j < BPL; ++j) prevb[j] = 0.f;

    for (int it = 0; it < nit; ++it, ++gi) {
      const int64_t tf = first + (int64_t)FR * it;
      const bool more = it + 1 < nit;
      const unsigned b = gi & 1u, use = gi >> 1;
      if (lock) mate_sync(bid);
      mbar_wait(bar + b, use & 1u);
      const float2* buf = reinterpret_cast<const float2*>(magbuf + b * mag_bytes);
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
          if (p > 0) {
            const float4 r01 = *reinterpret_cast<const float4*>(row - PROW + ppos(BPL * lane + 4 * q));
            const float4 r23 = *reinterpret_cast<const float4*>(row - PROW + ppos(BPL * lane + 4 * q + 2));
            pv[0] = r01.y, pv[1] = r01.w, pv[2] = r23.y, pv[3] = r23.w;
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) pv[u] = prevb[4 * q + u];
          }
          const float2 mm[4] = {make_float2(m01.x, m01.y), make_float2(m01.z, m01.w), make_float2(m23.x, m23.y),
                                make_float2(m23.z, m23.w)};
          const float xx[4] = {xv.x, xv.y, xv.z, xv.w}, ll[4] = {lv.x, lv.y, lv.z, lv.w}, hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            bin_step3(ac, true, 4 * q + u, (fmask >> (4 * q + u)) & 1u, mm[u], pv[u], xx[u], ll[u], hh[u]);
            if (p == PK - 1) prevb[4 * q + u] = mm[u].y;
          }
        }
        if (lane == 31) {  // Nyquist bin
          const float2 mq = row[nyq];
          const float pv = p == 0 ? prevny : row[nyq - PROW].y;
          bin_step3(ac, true, BPL, (fmask >> BPL) & 1u, mq, pv, s_xtab[spos(M)], s_wlo[spos(M)], s_whi[spos(M)]);
          if (p == PK - 1) prevny = mq.y;
        }
        ac.pp[0] = pk::add(ac.pend, ac.mlo);
        ac.pp[32] = ac.mhi;
        if (lock > 1) mate_sync(bid);

        const int64_t ta = tf + 2 * p, tb = ta + 1;
        const bool oka = ta >= t0 && ta < tend, okb = tb >= t0 && tb < tend;
        float2 pre = ac.seg;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float2 up = make_float2(__shfl_up_sync(kFull3, pre.x, o), __shfl_up_sync(kFull3, pre.y, o));
          if (lane >= o) pre = pk::add(pre, up);
        }
        const float2 etot = make_float2(__shfl_sync(kFull3, pre.x, 31), __shfl_sync(kFull3, pre.y, 31));
        const float2 plow = make_float2(__shfl_sync(kFull3, pre.x, 7), __shfl_sync(kFull3, pre.y, 7));
        const float2 sm = warp_sum3(ac.s0);
        const float2 skm = warp_sum3(pk::fma(ac.s0, k0f, ac.s1));
        const float2 kc = make_float2(sm.x > 0.f ? __fdividef(skm.x, sm.x) : 0.f, sm.y > 0.f ? __fdividef(skm.y, sm.y) : 0.f);
        const float2 dk = make_float2(k0f - kc.x, k0f - kc.y);
        const float2 bw = warp_sum3(__ffma2_rn(__fmul2_rn(dk, dk), ac.s0, __ffma2_rn(pk::scale(dk, 2.f), ac.s1, ac.s2)));
        float2 sl = warp_sum3(ac.sl), sxy = warp_sum3(ac.sxy);
        const float2 fl = warp_sum3(ac.fl);
        const float mxa = warp_max3(ac.mxa), mxb = warp_max3(ac.mxb);
        const float2 ri = row[M + 2];
        const float clog = kLogTau * (float)B * sqrt_fast3((float)B) / kEta;
        bool xa = !(ri.x * sqrt_fast3(etot.x) <= clog) || !(ri.x < 1.f / kTinyMag) || !(etot.x > 0.f);
        bool xb = !(ri.y * sqrt_fast3(etot.y) <= clog) || !(ri.y < 1.f / kTinyMag) || !(etot.y > 0.f);
        const float irb = 2.f * kEta * rsqrt_fast3((float)B);
        int rka = rolloff_bin<G>(rowf, pre.x, ac.seg.x, etot.x, irb * sm.x * rsqrt_fast3(fmaxf(etot.x, 1e-36f)) + 1e-7f, lane);
        int rkb = rolloff_bin<G>(rowf + 1, pre.y, ac.seg.y, etot.y, irb * sm.y * rsqrt_fast3(fmaxf(etot.y, 1e-36f)) + 1e-7f, lane);

        __syncwarp();  // private mel slots visible; the last read of this pack's row is behind every lane
        if (p == PK - 1 && lane == 0) mbar_arrive(bar + 2 + b);  // the transform warp may overwrite the buffer
        if (a.mfcc_on) {
          float2 dens = make_float2(FLT_MAX, FLT_MAX);
          for (int f = lane; f < a.n_mel; f += 32) {
            float2 v = make_float2(0.f, 0.f);
            for (int i = 0; i < ncontrib; ++i) v = pk::add(v, wbf2[s_moff[i * kMaxMel + f]]);
            macc[f] = make_float2(v.x > 0.f ? __logf(v.x) : -23.025850929940457f,
                                  v.y > 0.f ? __logf(v.y) : -23.025850929940457f);  // ln(1e-10)
            const float iw = s_invw[f];
            if (iw > 0.f) dens = make_float2(fminf(dens.x, v.x * iw), fminf(dens.y, v.y * iw));
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
        if (lane == 0) {
          const int slot = (int)(ta - first) & (kRun3 - 1);
          float4* rs = reinterpret_cast<float4*>(rawsum + slot * kRaw3);
          if (xa) rka |= kExactBit;
          if (xb) rkb |= kExactBit;
          rs[0] = make_float4(sm.x, kc.x, etot.x, __int_as_float(rka));
          rs[1] = make_float4(bw.x, sl.x, sxy.x, mxa);
          rs[2] = make_float4(fl.x, plow.x, 0.f, 0.f);
          if (slot + 1 < kRun3) {
            rs[3] = make_float4(sm.y, kc.y, etot.y, __int_as_float(rkb));
            rs[4] = make_float4(bw.y, sl.y, sxy.y, mxb);
            rs[5] = make_float4(fl.y, plow.y, 0.f, 0.f);
          }
        }
        __syncwarp();  // private slots / macc / parked sums: reused by the next pack, read by the finishing below
      }

      if (((FR * (it + 1)) & (kRun3 - 1)) == 0 || !more) {
        const int seg = (FR * it) / kRun3;
        const int64_t t = first + (int64_t)kRun3 * seg + lane;
        if (t >= t0 && t < tend) {
          const float4* rs4 = reinterpret_cast<const float4*>(rawsum + lane * kRaw3);
          const float4 r0 = rs4[0], r1 = rs4[1], r2 = rs4[2];
          const float sm = r0.x, kc = r0.y, etot = r0.z, bw = r1.x, sl = r1.y, sxy = r1.z, mx = r1.w, fl = r2.x, plow = r2.y;
          const int rkx = __float_as_int(r0.w), rk = rkx & ~kExactBit;
          if ((rkx & kExactBit) && a.xlist) {
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
            const double gm = exp2((double)sl / (double)B);
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
          if (a.slope_on) {
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

template <int LOGN, int HR>
__global__ void __launch_bounds__(kThreads4, 1) stft_v4_kernel(const StftArgs a) {
  using G = V3G<LOGN, HR>;
  constexpr int N = G::N, M = G::M, B = G::B, J = G::J, BPL = G::BPL, ROW = G::ROW;
  extern __shared__ __align__(128) unsigned char smem[];
  const V4Smem L = v4_layout<G>(a.n_mel, a.n_mfcc);
  float2* s_tw = reinterpret_cast<float2*>(smem + L.tw);
  float* s_win = reinterpret_cast<float*>(smem + L.win);
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
  __shared__ __align__(8) unsigned long long s_bar[kPairs4][4];  // full[0], full[1], empty[0], empty[1]
  __shared__ __align__(8) unsigned long long s_tbar[kPairs4];    // staging copies of the transform warps

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pr = warp & (kPairs4 - 1);
  unsigned char* pb = smem + L.pair0 + (size_t)pr * L.per_pair;
  float2* wbf2 = reinterpret_cast<float2*>(pb);
  float2* tile = reinterpret_cast<float2*>(pb + L.p_tile);
  unsigned char* magbuf = pb + L.p_mag;
  float2* priv = reinterpret_cast<float2*>(pb + L.p_priv);
  float* rawsum = reinterpret_cast<float*>(pb + L.p_raw);
  float2* macc = reinterpret_cast<float2*>(pb + L.p_macc);
  constexpr int kZeroSlot = kMaxMel + 2;

  // ---- tables, once per CTA (as in stft_v3_kernel) ----
  for (int i = threadIdx.x; i < J * 32; i += blockDim.x) {
    const int k1 = i / 32, l = i % 32;
    double dsn, dcs;
    sincospi(-2.0 * (double)((k1 * l) % N) / (double)N, &dsn, &dcs);
    s_tw[i] = make_float2((float)dcs, (float)dsn);
  }
  {
    const float* wsrc = reinterpret_cast<const float*>(a.win2);
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_win[i] = __ldg(wsrc + i);
  }
  for (int k = threadIdx.x; k < ROW; k += blockDim.x) {
    s_xtab[k] = 0.f;
    s_wlo[k] = 0.f;
    s_whi[k] = 0.f;
  }
  if (threadIdx.x < 32) s_fmask[threadIdx.x] = 0u;
  if (threadIdx.x < kMaxMel) s_invw[threadIdx.x] = a.mel_invw[threadIdx.x];
  if (threadIdx.x == 0) s_ncontrib = 0;
  if (warp >= kPairs4 && lane == 0) macc[kZeroSlot] = make_float2(0.f, 0.f);
  if (threadIdx.x < kPairs4 * 4) mbar_init(&s_bar[0][0] + threadIdx.x, 1u);
  if (threadIdx.x < kPairs4) mbar_init(&s_tbar[threadIdx.x], 1u);
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
  const unsigned short zero_off = (unsigned short)((macc + kZeroSlot) - wbf2);  // the same for every pair
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

  if (warp < kPairs4) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(kRegsFft4));
    v4_transform<G>(a, pr, lane, s_tw, s_win, tile, magbuf, L.mag_bytes, &s_bar[pr][0], &s_tbar[pr]);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(kRegsScan4));
    v4_scan<G>(a, pr, lane, s_xtab, s_wlo, s_whi, s_fmask[lane], s_moff, s_dct, s_lift, s_invw, s_ncontrib, wbf2, magbuf,
               L.mag_bytes, priv, rawsum, macc, &s_bar[pr][0]);
  }
}

template <int LOGN, int HR>
int v4_launch(StftArgs& a, cudaStream_t st) {
  using G = V3G<LOGN, HR>;
  a.runs_per_stream = (int)((a.T + kRunOut3 - 1) / kRunOut3);
  a.total_runs = (int64_t)a.runs_per_stream * a.n_streams;
  const V4Smem L = v4_layout<G>(a.n_mel, a.n_mfcc);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t ctas = (a.total_runs + kPairs4 - 1) / kPairs4;
  if (ctas > sms) ctas = sms;  // persistent: one CTA per SM, pairs stride over the runs
  if (ctas < 1) ctas = 1;
  static const int lockstep = std::getenv("SONAR_V4_LOCKSTEP") ? std::atoi(std::getenv("SONAR_V4_LOCKSTEP")) : 0;
  a.v4_lockstep = lockstep;
  prof_begin("stft_features_kernel", st);
  SONAR_CUDA(cudaFuncSetAttribute(stft_v4_kernel<LOGN, HR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  stft_v4_kernel<LOGN, HR><<<(unsigned)ctas, kThreads4, L.total, st>>>(a);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

template <class G>
bool v3_mel_eligible(const FpPlan& plan, const StftArgs& a) {
  constexpr int BPL = G::BPL, B = G::B;
  if (plan.h_regions.empty() || a.n_mel > kMaxMel || a.n_mfcc > kMaxMfcc || plan.split != G::M / 4) return false;
  auto region_of = [&](int k) {
    int r = 0;
    while (k >= plan.h_regions[r].next_b) ++r;
    return r;
  };
  int first[32], last[32];
  for (int j = 0; j < 32; ++j) {
    first[j] = region_of(BPL * j);
    last[j] = region_of(j == 31 ? B - 1 : BPL * j + BPL - 1);
    if (last[j] - first[j] + 2 > kSlots3) return false;
  }
  for (int k = 1; k < B; ++k)  // every mel region at least one bin wide (one private slot per boundary)
    if (region_of(k) - region_of(k - 1) > 1) return false;
  for (int f = 0; f < a.n_mel; ++f) {
    int cnt = 0;
    for (int j = 0; j < 32; ++j) cnt += (f + 1 >= first[j] - 1 && f + 1 <= last[j]);
    if (cnt > kMaxContrib3) return false;
  }
  return true;
}

template <int LOGN, int HR>
int v3_launch(StftArgs& a, cudaStream_t st) {
  using G = V3G<LOGN, HR>;
  a.runs_per_stream = (int)((a.T + kRunOut3 - 1) / kRunOut3);
  a.total_runs = (int64_t)a.runs_per_stream * a.n_streams;
  const V3Smem L = v3_layout<G>(a.n_mel, a.n_mfcc);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t ctas = (a.total_runs + kW3 - 1) / kW3;
  if (ctas > sms) ctas = sms;  // persistent: one CTA per SM, warps stride over the runs
  if (ctas < 1) ctas = 1;
  prof_begin("stft_features_kernel", st);
  SONAR_CUDA(cudaFuncSetAttribute(stft_v3_kernel<LOGN, HR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  stft_v3_kernel<LOGN, HR><<<(unsigned)ctas, kW3 * 32, L.total, st>>>(a);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

}  // namespace

// The third-generation kernel covers the two geometries the BASELINE configurations use (1024 / 256 and 512 / 160);
// everything else stays with stft_v2.cu / stft_features.cu.
bool stft_v3_eligible(const FpPlan& plan, const StftArgs& a) {
  static const bool off = std::getenv("SONAR_STFT_V2") != nullptr || std::getenv("SONAR_STFT_V1") != nullptr;  // diagnostic
  if (off) return false;
  if (plan.N == 1024 && a.hop == 256) return v3_mel_eligible<V3G<10, 8>>(plan, a);
  if (plan.N == 512 && a.hop == 160) return v3_mel_eligible<V3G<9, 5>>(plan, a);
  return false;
}

int launch_stft_v3(const FpPlan& plan, StftArgs& a, cudaStream_t st) {
  // The warp-specialised form (stft_v4_kernel: transform / scan warp pairs, TMA-staged sample rows, mbarrier hand-over,
  // setmaxnreg) is bit-identical in its results but measures 7.85 ms against 7.16 ms per 64 x 300 s: sixteen warps issue
  // no more than twelve (0.52 against 0.56 slots per scheduler-cycle) because the two roles' loops together are 50 KB of
  // straight-line code and the SM delivers instructions at full rate only out of its 32 KB L1.5 instruction cache
  // (scripts/microbench/icache_bw.cu).  It stays selectable for measurements: SONAR_STFT_V4=1.
  static const bool v4 = std::getenv("SONAR_STFT_V4") != nullptr;
  static const bool v3 = std::getenv("SONAR_STFT_V3") != nullptr;
  if (!v3 && !v4) {
    const int rc5 = launch_stft_v5(plan, a, st);
    if (rc5 != SONAR_ERR_UNSUPPORTED) return rc5;  // a mel bank whose private slots do not fit: the single-role kernel
  }
  constexpr size_t kSmemMax = 227 * 1024;
  const bool fits = plan.N == 1024 ? v4_layout<V3G<10, 8>>(a.n_mel, a.n_mfcc).total <= kSmemMax
                                   : v4_layout<V3G<9, 5>>(a.n_mel, a.n_mfcc).total <= kSmemMax;
  if (v4 && fits) {
    if (plan.N == 1024) return v4_launch<10, 8>(a, st);
    return v4_launch<9, 5>(a, st);
  }
  if (plan.N == 1024) return v3_launch<10, 8>(a, st);
  return v3_launch<9, 5>(a, st);
}

}  // namespace sonar
