// Exact float64 time-domain kernels.  Compiled with -fmad=false: Go on amd64 never
// contracts a*b+c, and the short-time energies computed here feed the
// cross-correlation arg-max and the DTW path, which must be bit-identical to
// the reference (SURVEY.md §7 "hard parts" #1).
//
//   pre-emphasis          algorithms/filters/pre_emphasis.go:135-155,184-190
//   short-time RMS energy algorithms/temporal/energy.go:25-50   (sequential sum, ascending j)
//   energy entropy        fingerprint/extractors/speech.go:429-434
//   zero-crossing rate    algorithms/spectral/zero_crossing_rate.go:37-53 on pre[tH : tH+W]
//   energy variance       algorithms/temporal/energy.go:97-118
//   loudness range        algorithms/temporal/energy.go:157-225
//   RMS envelope 512/256  fingerprint/extractors/speech.go:739-767
//
// Layout: a CTA stages the raw samples of 32 consecutive frames of one stream in shared
// memory (cp.async, one pad word per hop so that the per-lane sequential walks are
// bank-conflict free) and one warp accumulates ONE frame per lane in the reference's
// order.  The chain of dependent DADDs is what bounds this kernel, not HBM.
#include <cmath>
#include <cstdlib>

#include "common.h"

namespace sonar {
namespace {

constexpr int kTdThreads = 64;     // two warps stage a tile, warp 0 walks it
constexpr int kTdFrames = 32;      // frames (sequential float64 chains) per CTA: one per lane of warp 0

__device__ __forceinline__ double preemph(const double* __restrict__ x, int64_t i, double alpha) {
  const double prev = i > 0 ? x[i - 1] : 0.0;
  return x[i] - alpha * prev;  // -fmad=false: separate multiply and subtract
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src));
}

// One CTA = 32 consecutive frames of one stream.  The RAW samples they cover (plus the one sample before
// the first frame, needed by the pre-emphasis) are copied to shared memory with cp.async — no registers,
// every request in flight at once — one pad word per hop so that the 32 sequential walks are bank-conflict
// free.  Lane t of warp 0 then walks frame t in the reference's order: y = x[i] - alpha*x[i-1] (x[-1] = 0),
// sum += y*y, sign changes of y counted on the way.  The dependent DADD chain of each frame is the only
// serial part; loads, the pre-emphasis and the squares of later samples are independent of it.  Tiles are
// ~70 KB so three CTAs share an SM and one CTA's staging overlaps the others' walks.
template <bool ENERGY, bool ZCR>
__global__ void __launch_bounds__(kTdThreads) frame_walk_kernel(
    const double* __restrict__ pcm, int64_t n, int64_t stride, double alpha, int frame, int hop, int64_t Tn,
    int sr, double* __restrict__ out, int64_t out_stride, int64_t o_energy, int64_t o_entropy,
    int64_t o_zcr, int fpc) {  // fpc <= kTdFrames frames per CTA (fewer when a long hop would overflow the tile)
  extern __shared__ double tile[];
  const int s = blockIdx.y;
  const int64_t f0 = (int64_t)blockIdx.x * fpc;
  if (f0 >= Tn) return;
  const int nf = (int)((Tn - f0 < fpc) ? (Tn - f0) : fpc);
  const double* __restrict__ x = pcm + (int64_t)s * stride;
  const int64_t g0 = f0 * hop - 1;                // global index of staged element 0
  const int count = (nf - 1) * hop + frame + 1;  // staged elements
  // element e lives at e + e / hop: one pad word per hop-sized block (no per-element division: blocks outside)
  for (int blk = 0, e0 = 0; e0 < count; ++blk, e0 += hop) {
    const int cnt_b = count - e0 < hop ? count - e0 : hop;
    const int64_t gb = g0 + e0;
    if (gb >= 0) {  // every block but the stream's very first: plain pointer walk, 16 bytes per request when aligned
      const double* __restrict__ src = x + gb;
      unsigned dst = (unsigned)__cvta_generic_to_shared(tile + e0 + blk);
      if (((reinterpret_cast<uintptr_t>(src) | dst) & 15) == 0 && (cnt_b & 1) == 0) {
        const double* sp = src + 2 * threadIdx.x;
        unsigned dp = dst + 16 * threadIdx.x;
#pragma unroll 4
        for (int o = 2 * threadIdx.x; o < cnt_b; o += 2 * kTdThreads, sp += 2 * kTdThreads, dp += 16 * kTdThreads)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dp), "l"(sp));
      } else {
        const double* sp = src + threadIdx.x;
        unsigned dp = dst + 8 * threadIdx.x;
#pragma unroll 4
        for (int o = threadIdx.x; o < cnt_b; o += kTdThreads, sp += kTdThreads, dp += 8 * kTdThreads)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dp), "l"(sp));
      }
    } else {
      double* dst = tile + e0 + blk;
      for (int o = threadIdx.x; o < cnt_b; o += kTdThreads) {
        if (gb + o >= 0)
          cp_async8(dst + o, x + gb + o);
        else
          dst[o] = 0.0;  // x[-1] = 0 (pre_emphasis.go:135-155, lastSample starts at 0)
      }
    }
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int t = threadIdx.x;
  if (t >= nf) return;
  // thread t: elements t*hop + u, u = 0 (previous sample), 1 .. frame; physical = t*(hop+1) + u + u/hop
  const double* __restrict__ base = tile + t * (hop + 1);
  double xprev = base[0];
  double sum = 0.0;
  int crossings = 0;
  bool prev_neg = false;
  // one sample of the walk; y < 0 is tested on the bit pattern (sign set and not -0.0) to keep the two
  // comparisons per sample off the FP64 pipe, which the dependent sum chain needs
  auto step = [&](double xv) {
    const double y = xv - alpha * xprev;
    xprev = xv;
    if (ENERGY) sum += y * y;
    if (ZCR) {
      const bool neg = (unsigned long long)__double_as_longlong(y) > 0x8000000000000000ull;
      crossings += (neg != prev_neg) ? 1 : 0;
      prev_neg = neg;
    }
  };
  int u = 1, j = 0, b = 0;  // b = u / hop
  while (j < frame) {
    int cnt = (b + 1) * hop - u;
    if (cnt > frame - j) cnt = frame - j;
    const double* __restrict__ p = base + u + b;
    // eight samples are in registers one block ahead of their use: the shared-memory latency stays off the chain
    int c = 0;
    double cur[8];
    if (cnt >= 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) cur[k] = p[k];
    }
    for (; c + 8 <= cnt; c += 8) {
      double nxt[8];
      const bool more = c + 16 <= cnt;
      if (more) {
#pragma unroll
        for (int k = 0; k < 8; ++k) nxt[k] = p[c + 8 + k];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) step(cur[k]);
      if (more) {
#pragma unroll
        for (int k = 0; k < 8; ++k) cur[k] = nxt[k];
      }
    }
    for (; c < cnt; ++c) step(p[c]);
    u += cnt;
    j += cnt;
    if (u == (b + 1) * hop) ++b;
  }
  if (ZCR) {  // the first sample has no predecessor inside the frame (zero_crossing_rate.go:43: i from 1)
    const double y0 = base[1] - alpha * base[0];
    if (y0 < 0.0) crossings -= 1;  // undo the spurious change against the initial prev_neg = false
  }
  double* __restrict__ o = out + (int64_t)s * out_stride;
  const int64_t f = f0 + t;
  if (ENERGY) {
    const double e = sqrt(sum / (double)frame);
    o[o_energy + f] = e;
    if (o_entropy >= 0) o[o_entropy + f] = e > 0.0 ? -e * log(e + 1e-10) : 0.0;
  }
  if (ZCR) {
    const double dur = (double)frame / (double)sr;  // len(frame)/sampleRate; sr==0 -> +Inf -> zcr 0
    o[o_zcr + f] = frame < 2 ? 0.0 : (double)crossings / dur;
  }
}

// The same walk without the shared-memory tile, for frames that are a whole number R of hops long: a thread owns
// G CONSECUTIVE frames and walks their union once (hop-sized segments), so every pre-emphasised sample and its
// square are computed once and added to the (up to R) running sums of the frames that contain it — each frame's
// sum still receives its terms in ascending sample order, i.e. bit-identical.  G independent chains per thread hide
// the FP64 add latency, and with no tile the SM holds as many warps as registers allow instead of ~100 chains.
// Every lane streams through its own part of the signal (16-byte loads, each 128-byte line serves eight of them
// from L1).
// With BS the walk also leaves partial sums of y^2 over the hop-sized blocks of the loudness windows (energy.go:157-179:
// 400 ms windows every 100 ms): every thread owns the samples of its G hops (the last thread of a stream everything
// up to wb.limit), a block boundary falls inside that range at most once (the launcher checks the sizes), so a thread
// writes two sums: the part of its range before the boundary and the part after it.  rms_from_parts_kernel adds the
// 3-4 parts of each block in thread order.  This replaces a second pass over the PCM (rms_blocks_kernel).
struct WalkBlocks {
  double* part;         // per stream 2 doubles per walk thread, part_stride apart
  int64_t part_stride;
  int64_t hopL;         // block length
  int64_t limit;        // samples [0, limit) belong to the blocks in use
};

// STAGED (hop a multiple of 16, aligned rows): the lanes of a warp walk ranges G hops apart, so a direct 16-byte load
// touches 32 different 128-byte lines per instruction and the kernel is bound by the L1 tag stage (ncu: l1tex 85 %,
// 0.15 issue slots per scheduler-cycle; 2.0 of its 2.9 ms per 64 x 300 s are tag cycles).  Here the warp fetches the next
// 16 samples of all its 32 lanes TOGETHER -- eight cp.async instructions of four whole lines each -- into a padded
// shared-memory tile (row stride 144 B: the lanes' 16-byte reads of their own rows are conflict free), one chunk ahead
// of the arithmetic.  Every lane still sees its samples in the same order: results are bit-identical.
// RT > 0 (staged form, frame = RT hops): the sums of the RT frames that contain the current hop live in a SLIDING window
// of RT registers that every sample is added to unconditionally (slot j = the frame that started j hops ago; at the end
// of a hop the oldest frame is finished and written, the others move up): the generic form below tests eight
// frame-active flags per sample, 51 instructions per sample with its ALU pipe 71 % busy (ncu), this one ~20.  Each
// frame's sum still receives its squares one by one in ascending order from 0.0 (bit-identical); the crossing counts are
// integers, so they are taken per hop and added to the frames at the end of the hop (the frame that STARTS in a hop
// does not count the hop's first sample: zero_crossing_rate.go:43).
template <int G, bool ALIGNED, bool BS, bool STAGED = false, int RT = 0>
__global__ void __launch_bounds__(128) frame_walk_multi_kernel(
    const double* __restrict__ pcm, int64_t stride, double alpha, int frame, int hop, int R, int64_t Tn, int sr,
    double* __restrict__ out, int64_t out_stride, int64_t o_energy, int64_t o_entropy, int64_t o_zcr, WalkBlocks wb) {
  const int s = blockIdx.y;
  const int64_t tw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t f0 = tw * G;
  if (!STAGED && f0 >= Tn) return;  // (staged: lanes without frames still help to load)
  const int nf = f0 >= Tn ? 0 : (int)((Tn - f0 < G) ? (Tn - f0) : G);
  const double* __restrict__ x = pcm + (int64_t)s * stride;
  const int64_t s0 = f0 * hop;
  double sum[G];
  int cnt[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    sum[g] = 0.0;
    cnt[g] = 0;
  }
  double xprev = (s0 > 0 && nf > 0) ? x[s0 - 1] : 0.0;  // x[-1] = 0 (pre_emphasis.go:135-155)
  bool prev_neg = false;
  const int nseg = nf > 0 ? nf - 1 + R : 0;
  // block sums (BS): bcur collects the owned samples, p_first keeps what was collected before the boundary
  const bool last_thread = nf > 0 && f0 + G >= Tn;
  const int64_t next_b = BS ? (s0 / wb.hopL + 1) * wb.hopL : 0;  // first sample of the next block
  double bcur = 0.0, p_first = 0.0;
  bool flushed = false;
  if constexpr (STAGED) {
    constexpr int kRowD = 18;  // doubles per tile row: 16 samples + 16 bytes of padding
    __shared__ __align__(16) double s_tile[4][2][32][kRowD];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int64_t f0w = f0 - (int64_t)lane * G;  // first frame of lane 0
    const int cps = hop / 16;                    // chunks per segment
    const int nseg_w = __shfl_sync(0xffffffffu, nseg, 0);  // lane 0 has the most segments
    const int total = nseg_w * cps;
    auto issue = [&](int q, int buf) {
      const int sg = q / cps, c = q - sg * cps;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = 4 * i + (lane >> 3), col = lane & 7;
        const int64_t f0r = f0w + (int64_t)row * G;
        const int nfr = f0r >= Tn ? 0 : (int)((Tn - f0r < G) ? (Tn - f0r) : G);
        if (nfr > 0 && sg < nfr - 1 + R) {
          const double* src = x + f0r * hop + (int64_t)sg * hop + 16 * c + 2 * col;
          const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_tile[wrp][buf][row][2 * col]);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (total > 0) issue(0, 0);
    bool act[G];
    int rel_b = -1;
    bool own = false;
    double wsum[RT > 0 ? RT : 1];
    int wcnt[RT > 0 ? RT : 1], segc = 0, firstc = 0;
#pragma unroll
    for (int j = 0; j < (RT > 0 ? RT : 1); ++j) {
      wsum[j] = 0.0;
      wcnt[j] = 0;
    }
    double* __restrict__ o = out + (int64_t)s * out_stride;
    const double dur = (double)frame / (double)sr;
    for (int q = 0; q < total; ++q) {
      const int sg = q / cps, c = q - sg * cps, buf = q & 1;
      if (q + 1 < total) {
        issue(q + 1, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncwarp();
      if (c == 0) {
        if (RT == 0) {
#pragma unroll
          for (int g = 0; g < G; ++g) act[g] = g < nf && g <= sg && sg < g + R;
        }
        own = BS && (last_thread || sg < G);
        const int64_t seg0 = s0 + (int64_t)sg * hop;
        rel_b = (own && next_b >= seg0 && next_b < seg0 + hop) ? (int)(next_b - seg0) : -1;
        segc = 0;
      }
      auto step = [&](double xv, bool first, int idx) {
        const double y = xv - alpha * xprev;
        xprev = xv;
        const double sq = y * y;
        const bool neg = (unsigned long long)__double_as_longlong(y) > 0x8000000000000000ull;
        const int cross = (neg != prev_neg) ? 1 : 0;
        prev_neg = neg;
        if (BS) {
          const bool hit = idx == rel_b;
          p_first = hit ? bcur : p_first;
          flushed = flushed || hit;
          const double add = own ? sq : 0.0;
          bcur = hit ? add : bcur + add;
        }
        if (RT > 0) {
#pragma unroll
          for (int j = 0; j < (RT > 0 ? RT : 1); ++j) wsum[j] += sq;
          segc += cross;
          if (first) firstc = cross;
        } else {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (act[g]) {
              sum[g] += sq;
              if (!(first && sg == g)) cnt[g] += cross;
            }
          }
        }
      };
      if (sg < nseg) {
        const double2* __restrict__ row = reinterpret_cast<const double2*>(&s_tile[wrp][buf][lane][0]);
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = row[u];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          step(v[u].x, c == 0 && u == 0, 16 * c + 2 * u);
          step(v[u].y, false, 16 * c + 2 * u + 1);
        }
        if (RT > 0 && c == cps - 1) {  // end of the hop: the oldest frame of the window is complete
#pragma unroll
          for (int j = 0; j < (RT > 0 ? RT : 1); ++j) wcnt[j] += segc;
          wcnt[0] -= firstc;
          const int g = sg - (RT - 1);
          if (g >= 0 && g < nf) {
            const int64_t f = f0 + g;
            const double e = sqrt(wsum[RT > 0 ? RT - 1 : 0] / (double)frame);
            o[o_energy + f] = e;
            if (o_entropy >= 0) o[o_entropy + f] = e > 0.0 ? -e * log(e + 1e-10) : 0.0;
            o[o_zcr + f] = frame < 2 ? 0.0 : (double)wcnt[RT > 0 ? RT - 1 : 0] / dur;
          }
#pragma unroll
          for (int j = (RT > 0 ? RT : 1) - 1; j > 0; --j) {
            wsum[j] = wsum[j - 1];
            wcnt[j] = wcnt[j - 1];
          }
          wsum[0] = 0.0;
          wcnt[0] = 0;
        }
      }
      __syncwarp();  // the buffer is rewritten two chunks later
    }
  } else
  for (int sg = 0; sg < nseg; ++sg) {
    bool act[G];
#pragma unroll
    for (int g = 0; g < G; ++g) act[g] = g < nf && g <= sg && sg < g + R;
    const double* __restrict__ p = x + s0 + (int64_t)sg * hop;
    const bool own = BS && (last_thread || sg < G);
    // offset of the block boundary inside this segment, -1 = none (no divergent path: lanes meet their boundary in
    // different segments, a branch would make nearly every warp run both paths)
    const int64_t seg0 = s0 + (int64_t)sg * hop;
    const int rel_b = (own && next_b >= seg0 && next_b < seg0 + hop) ? (int)(next_b - seg0) : -1;
    auto step = [&](double xv, bool first, int idx) {
      const double y = xv - alpha * xprev;
      xprev = xv;
      const double sq = y * y;
      const bool neg = (unsigned long long)__double_as_longlong(y) > 0x8000000000000000ull;
      const int cross = (neg != prev_neg) ? 1 : 0;
      prev_neg = neg;
      if (BS) {
        const bool hit = idx == rel_b;  // first sample of the next block: park what was collected so far
        p_first = hit ? bcur : p_first;
        flushed = flushed || hit;
        const double add = own ? sq : 0.0;
        bcur = hit ? add : bcur + add;
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (act[g]) {
          sum[g] += sq;
          // the first sample of a frame has no predecessor inside the frame (zero_crossing_rate.go:43)
          if (!(first && sg == g)) cnt[g] += cross;
        }
      }
    };
    if (ALIGNED) {  // hop even, 16-byte aligned rows
      const double2* __restrict__ p2 = reinterpret_cast<const double2*>(p);
      {
        const double2 v = p2[0];
        step(v.x, true, 0);
        step(v.y, false, 1);
      }
      int e = 1;
      for (; e + 4 <= hop / 2; e += 4) {
        const double2 a = p2[e], b = p2[e + 1], c = p2[e + 2], d = p2[e + 3];
        step(a.x, false, 2 * e); step(a.y, false, 2 * e + 1); step(b.x, false, 2 * e + 2); step(b.y, false, 2 * e + 3);
        step(c.x, false, 2 * e + 4); step(c.y, false, 2 * e + 5); step(d.x, false, 2 * e + 6); step(d.y, false, 2 * e + 7);
      }
      for (; e < hop / 2; ++e) {
        const double2 a = p2[e];
        step(a.x, false, 2 * e);
        step(a.y, false, 2 * e + 1);
      }
    } else {
      step(p[0], true, 0);
      for (int e = 1; e < hop; ++e) step(p[e], false, e);
    }
  }
  if (BS) {
    if (last_thread) {  // the samples behind the last frame still belong to loudness blocks
      for (int64_t i = s0 + (int64_t)nseg * hop; i < wb.limit; ++i) {
        if (i == next_b) {
          p_first = bcur;
          bcur = 0.0;
          flushed = true;
        }
        const double xv = x[i];
        const double y = xv - alpha * xprev;
        xprev = xv;
        bcur += y * y;
      }
    }
    if (nf > 0) {
      double* part = wb.part + (int64_t)s * wb.part_stride + 2 * tw;
      part[0] = flushed ? p_first : bcur;  // block s0 / hopL
      part[1] = flushed ? bcur : 0.0;      // the block after it
    }
  }
  if (STAGED && RT > 0) return;  // the sliding form wrote every frame as it finished
  double* __restrict__ o = out + (int64_t)s * out_stride;
  const double dur = (double)frame / (double)sr;  // len(frame)/sampleRate; sr==0 -> +Inf -> zcr 0
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (g >= nf) break;
    const int64_t f = f0 + g;
    const double e = sqrt(sum[g] / (double)frame);
    o[o_energy + f] = e;
    if (o_entropy >= 0) o[o_entropy + f] = e > 0.0 ? -e * log(e + 1e-10) : 0.0;
    o[o_zcr + f] = frame < 2 ? 0.0 : (double)cnt[g] / dur;
  }
}

// Unbiased variance (two passes, block tree reduction; tolerance 1e-4, not bit-exact).
__global__ void __launch_bounds__(256) variance_kernel(const double* __restrict__ xs, int64_t n, int64_t stride,
                                                       double* __restrict__ out, int64_t out_stride) {
  __shared__ double red[256];
  __shared__ double s_mean;
  const double* x = xs + (int64_t)blockIdx.x * stride;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += x[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) s_mean = n > 0 ? red[0] / (double)n : 0.0;
  __syncthreads();
  const double mean = s_mean;
  acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = x[i] - mean;
    acc += d * d;
  }
  __syncthreads();
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[(int64_t)blockIdx.x * out_stride] = n < 2 ? 0.0 : red[0] / (double)(n - 1);
}

// RMS of pre-emphasised windows (one warp per window; tree reduction).
__global__ void __launch_bounds__(256) rms_windows_kernel(const double* __restrict__ pcm, int64_t n,
                                                          int64_t stride, double alpha, int win, int hop,
                                                          int64_t nw, double* __restrict__ out,
                                                          int64_t out_stride) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= nw) return;
  const int s = blockIdx.y;
  const double* x = pcm + (int64_t)s * stride;
  const int64_t s0 = w * hop;
  double acc = 0.0;
  for (int j = lane; j < win; j += 32) {
    const double y = preemph(x, s0 + j, alpha);
    acc += y * y;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[(int64_t)s * out_stride + w] = sqrt(acc / (double)win);
}

// Same, for windows that are a whole number R of hops long (the 400 ms / 100 ms loudness windows and the 512 / 256
// envelope): a CTA produces 16 consecutive windows from the 16 + R - 1 hop-sized block sums they share, so every
// sample is squared ~1.2 times instead of R times.
constexpr int kRbWin = 16;
constexpr int kRbWarps = 16;
// short hops (the 512 / 256 envelope): one warp per hop block
__global__ void __launch_bounds__(kRbWin * 32) rms_blocks_warp_kernel(const double* __restrict__ pcm, int64_t stride,
                                                                 double alpha, int win, int hop, int R, int64_t nw,
                                                                 double* __restrict__ out, int64_t out_stride) {
  __shared__ double part[kRbWin + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.y;
  const double* __restrict__ x = pcm + (int64_t)s * stride;
  const int64_t w0 = (int64_t)blockIdx.x * kRbWin;
  const int64_t last_block = nw - 1 + R - 1;
  for (int b = warp; b < kRbWin + R - 1; b += kRbWin) {
    const int64_t blk = w0 + b;
    double acc = 0.0;
    if (blk <= last_block) {
      const int64_t s0 = blk * hop;
      // four independent strips per lane: eight loads in flight (this kernel waits on DRAM, not on arithmetic)
      double a4[4] = {0.0, 0.0, 0.0, 0.0};
      int j = lane;
      for (; j + 96 < hop; j += 128) {
        double cur[4], prv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t i = s0 + j + 32 * u;
          cur[u] = x[i];
          prv[u] = i > 0 ? x[i - 1] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double y = cur[u] - alpha * prv[u];
          a4[u] += y * y;
        }
      }
      for (; j < hop; j += 32) {
        const int64_t i = s0 + j;
        const double y = x[i] - alpha * (i > 0 ? x[i - 1] : 0.0);
        a4[0] += y * y;
      }
      acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (lane == 0) part[b] = acc;
  }
  __syncthreads();
  if (threadIdx.x < kRbWin && w0 + threadIdx.x < nw) {
    double acc = 0.0;
    for (int k = 0; k < R; ++k) acc += part[threadIdx.x + k];
    out[(int64_t)s * out_stride + w0 + threadIdx.x] = sqrt(acc / (double)win);
  }
}

// long hops (the 100 ms loudness hop): every warp takes a slice of every block
__global__ void __launch_bounds__(kRbWarps * 32) rms_blocks_kernel(const double* __restrict__ pcm, int64_t stride,
                                                                   double alpha, int win, int hop, int R, int64_t nw,
                                                                   double* __restrict__ out, int64_t out_stride) {
  __shared__ double part[kRbWin + 8][kRbWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.y;
  const double* __restrict__ x = pcm + (int64_t)s * stride;
  const int64_t w0 = (int64_t)blockIdx.x * kRbWin;
  const int64_t last_block = nw - 1 + R - 1;
  // every warp takes the same slice of every hop block (16 + R - 1 blocks do not divide over 16 warps); the slice
  // sums meet in shared memory and are added in warp order, so the result does not depend on scheduling
  const int per = (hop + kRbWarps - 1) / kRbWarps;
  const int j0 = warp * per, j1 = (j0 + per < hop) ? j0 + per : hop;
#pragma unroll 2
  for (int b = 0; b < kRbWin + R - 1; ++b) {
    const int64_t blk = w0 + b;
    double acc = 0.0;
    if (blk <= last_block) {
      const int64_t s0 = blk * hop;
      // independent strips per lane: this kernel waits on DRAM, not on arithmetic
      double a4[4] = {0.0, 0.0, 0.0, 0.0};
      int j = j0 + lane;
      for (; j + 96 < j1; j += 128) {
        double cur[4], prv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t i = s0 + j + 32 * u;
          cur[u] = x[i];
          prv[u] = i > 0 ? x[i - 1] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double y = cur[u] - alpha * prv[u];
          a4[u] += y * y;
        }
      }
      for (int u = 0; j < j1; j += 32, ++u) {
        const int64_t i = s0 + j;
        const double y = x[i] - alpha * (i > 0 ? x[i - 1] : 0.0);
        a4[u & 3] += y * y;
      }
      acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (lane == 0) part[b][warp] = acc;
  }
  __syncthreads();
  if (threadIdx.x < kRbWin + R - 1) {  // block sums, slices in warp order
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kRbWarps; ++k) t += part[threadIdx.x][k];
    part[threadIdx.x][0] = t;
  }
  __syncthreads();
  if (threadIdx.x < kRbWin && w0 + threadIdx.x < nw) {
    double acc = 0.0;
    for (int k = 0; k < R; ++k) acc += part[threadIdx.x + k][0];
    out[(int64_t)s * out_stride + w0 + threadIdx.x] = sqrt(acc / (double)win);
  }
}

// RMS of the loudness windows from the block parts frame_walk_multi_kernel<.., BS = true> left: block b is covered by the
// walk threads whose sample range [t U, (t + 1) U) (the last thread: to the end) meets [b hopL, (b + 1) hopL); their
// parts are added in thread order, the R blocks of a window in block order.
__global__ void __launch_bounds__(128) rms_from_parts_kernel(const double* __restrict__ part, int64_t part_stride,
                                                            int64_t n_threads, int64_t U, int64_t hopL, int R, int win,
                                                            int64_t nw, double* __restrict__ out, int64_t out_stride) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  const double* __restrict__ P = part + (int64_t)blockIdx.y * part_stride;
  const int64_t last = n_threads - 1;
  double acc = 0.0;
  for (int k = 0; k < R; ++k) {
    const int64_t b = w + k;
    int64_t t_lo = (b * hopL) / U, t_hi = ((b + 1) * hopL - 1) / U;
    if (t_lo > last) t_lo = last;
    if (t_hi > last) t_hi = last;
    double bs = 0.0;
    for (int64_t t = t_lo; t <= t_hi; ++t) {
      const int64_t b0 = (t * U) / hopL;
      if (b0 == b)
        bs += P[2 * t];
      else if (b0 + 1 == b)
        bs += P[2 * t + 1];
    }
    acc += bs;
  }
  out[(int64_t)blockIdx.y * out_stride + w] = sqrt(acc / (double)win);
}

// loudness units + 10th/95th percentile range on <= 4096 values per stream (bitonic sort in smem)
__global__ void __launch_bounds__(256) loudness_range_kernel(const double* __restrict__ rms, int64_t nw,
                                                             int64_t in_stride, double* __restrict__ out,
                                                             int64_t out_stride) {
  extern __shared__ double v[];
  const int s = blockIdx.x;
  int np2 = 1;
  while (np2 < nw) np2 <<= 1;
  for (int i = threadIdx.x; i < np2; i += blockDim.x) {
    double lv = INFINITY;
    if (i < nw) {
      const double e = rms[(int64_t)s * in_stride + i];
      lv = e > 0.0 ? -0.691 + 10.0 * log10(e * e) : -70.0;
    }
    v[i] = lv;
  }
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool up = (i & k) == 0;
          const double a = v[i], b = v[ixj];
          if ((a > b) == up) {
            v[i] = b;
            v[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  if (threadIdx.x == 0) {
    double res = 0.0;
    if (nw > 0) {
      const int lo = (int)(0.10 * (double)(nw - 1)), hi = (int)(0.95 * (double)(nw - 1));
      double lov = v[lo];
      const double hiv = v[hi];
      if (lov <= 0.0) lov = 1e-10;
      res = hiv <= 0.0 ? 0.0 : 20.0 * log10(hiv / lov);
    }
    out[(int64_t)s * out_stride] = res;
  }
}

// Same result for any number of windows (1 h at 16 kHz has 36,000): the two order statistics are found
// exactly with an 8-pass byte-wise radix select over the order-preserving 64-bit keys of the loudness
// values (converted in place in the scratch array), one CTA per stream.
__device__ __forceinline__ unsigned long long f64_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(256) loudness_select_kernel(double* __restrict__ rms, int64_t nw, int64_t in_stride,
                                                              double* __restrict__ out, int64_t out_stride) {
  __shared__ unsigned hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ long long s_k;
  double* v = rms + (int64_t)blockIdx.x * in_stride;
  for (int64_t i = threadIdx.x; i < nw; i += blockDim.x) {
    const double e = v[i];
    v[i] = e > 0.0 ? -0.691 + 10.0 * log10(e * e) : -70.0;
  }
  __syncthreads();
  double picked[2] = {0.0, 0.0};
  for (int which = 0; which < 2; ++which) {
    if (threadIdx.x == 0) {
      s_prefix = 0ull;
      s_k = (long long)((which == 0 ? 0.10 : 0.95) * (double)(nw - 1));  // int() truncation as in energy.go:203-213
    }
    __syncthreads();
    unsigned long long mask = 0ull;
    for (int pass = 7; pass >= 0; --pass) {
      hist[threadIdx.x] = 0u;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      for (int64_t i = threadIdx.x; i < nw; i += blockDim.x) {
        const unsigned long long k = f64_key(v[i]);
        if ((k & mask) == prefix) atomicAdd(&hist[(unsigned)(k >> (8 * pass)) & 0xffu], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
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
    picked[which] = key_f64(s_prefix);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double res = 0.0;
    if (nw > 0) {
      double lov = picked[0];
      const double hiv = picked[1];
      if (lov <= 0.0) lov = 1e-10;
      res = hiv <= 0.0 ? 0.0 : 20.0 * log10(hiv / lov);
    }
    out[(int64_t)blockIdx.x * out_stride] = res;
  }
}

__global__ void fill_kernel(double* p, int64_t n, double v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

__global__ void fill_strided_kernel(double* p, int64_t count, int64_t stride, double v) {
  double* q = p + (int64_t)blockIdx.y * stride;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    q[i] = v;
}

}  // namespace

int launch_frame_walk(const double* pcm, int64_t n, int64_t stride, int n_streams, double alpha, int frame,
                      int hop, int64_t Tn, int sr, double* out, int64_t out_stride, int64_t o_energy,
                      int64_t o_entropy, int64_t o_zcr, cudaStream_t st, const WalkLoudness* wl, bool* wl_done) {
  if (wl_done) *wl_done = false;
  if (Tn <= 0 || n_streams <= 0) return SONAR_OK;
  int fpc = kTdFrames;  // frames per CTA: as many as keep the staged tile under 200 KB
  while (fpc > 1 && sizeof(double) * (size_t)(((int64_t)(fpc - 1) * hop + frame + 1) * (hop + 1) / hop + 2) > 200 * 1024) --fpc;
  const int64_t count = (int64_t)(fpc - 1) * hop + frame + 1;
  const size_t smem = sizeof(double) * (size_t)(count + count / hop + 2);
  dim3 grid((unsigned)((Tn + fpc - 1) / fpc), (unsigned)n_streams);
  const bool en = o_energy >= 0, zc = o_zcr >= 0;
  if (!en && !zc) return SONAR_OK;
  static const bool tiled_only = std::getenv("SONAR_FRAME_WALK_TILED") != nullptr;  // diagnostic
  if (en && zc && !tiled_only && hop > 0 && frame % hop == 0 && frame / hop <= 8) {
    const bool aligned = (hop % 2 == 0) && (stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(pcm) & 15) == 0);
    static const bool direct = std::getenv("SONAR_FRAME_WALK_DIRECT") != nullptr;  // diagnostic: per-lane 16-byte loads
    static const bool flags = std::getenv("SONAR_FRAME_WALK_FLAGS") != nullptr;    // diagnostic: per-sample active flags
    const bool staged = aligned && !direct && hop % 16 == 0;
    const bool slide = staged && !flags && frame / hop == 4;
    // Frames per thread.  A thread reads G + R - 1 hops to finish G frames (R = frame / hop), so the sliding-window form,
    // whose registers do not grow with G, takes 12: 15 hops per 12 frames instead of 11 per 8 (-9 % samples walked;
    // measured 1.90 -> 1.85 ms per 64 x 300 s: the kernel waits on its staged loads, long_scoreboard 1.5 warps per issue,
    // more than on its arithmetic).  12 is also the most the fused loudness block sums allow at 44.1 kHz (a thread's range
    // may hold one 4,410-sample block boundary).
    auto run = [&](auto g_tag) -> int {
    constexpr int G = decltype(g_tag)::value;
    const int64_t groups = (Tn + G - 1) / G;
    const dim3 mg((unsigned)((groups + 127) / 128), (unsigned)n_streams);
    // loudness by-product: a thread's range (G hops, the last thread up to a frame and a hop more) must hold at most
    // one block boundary, and the blocks in use must end inside the stream
    WalkBlocks wb{nullptr, 0, 0, 0};
    const bool fuse = wl && wl->nw > 0 && wl->hop > 0 && wl->win % wl->hop == 0 && wl->win / wl->hop <= 8 &&
                      wl->hop > (int64_t)(G + 1) * hop + frame && (wl->nw - 1) * wl->hop + wl->win <= n &&
                      wl->part_stride >= 2 * groups;
    if (fuse) wb = WalkBlocks{wl->part, wl->part_stride, wl->hop, (wl->nw - 1) * wl->hop + wl->win};
    prof_begin("frame_walk_kernel", st);
    if (slide && fuse)
      frame_walk_multi_kernel<G, true, true, true, 4><<<mg, 128, 0, st>>>(pcm, stride, alpha, frame, hop, frame / hop, Tn,
                                                                         sr, out, out_stride, o_energy, o_entropy, o_zcr, wb);
    else if (slide)
      frame_walk_multi_kernel<G, true, false, true, 4><<<mg, 128, 0, st>>>(pcm, stride, alpha, frame, hop, frame / hop, Tn,
                                                                          sr, out, out_stride, o_energy, o_entropy, o_zcr, wb);
    else if (staged && fuse)
      frame_walk_multi_kernel<G, true, true, true><<<mg, 128, 0, st>>>(pcm, stride, alpha, frame, hop, frame / hop, Tn, sr,
                                                                      out, out_stride, o_energy, o_entropy, o_zcr, wb);
    else if (staged)
      frame_walk_multi_kernel<G, true, false, true><<<mg, 128, 0, st>>>(pcm, stride, alpha, frame, hop, frame / hop, Tn, sr,
                                                                       out, out_stride, o_energy, o_entropy, o_zcr, wb);
    else if (aligned && fuse)
      frame_walk_multi_kernel<G, true, true><<<mg, 128, 0, st>>>(pcm, stride, alpha, frame, hop, frame / hop, Tn, sr, out,
                                                                out_stride, o_energy, o_entropy, o_zcr, wb);
    else if (aligned)
      frame_walk_multi_kernel<G, true, false><<<mg, 128, 0, st>>>(pcm, stride, alpha, frame, hop, frame / hop, Tn, sr, out,
                                                                 out_stride, o_energy, o_entropy, o_zcr, wb);
    else if (fuse)
      frame_walk_multi_kernel<G, false, true><<<mg, 128, 0, st>>>(pcm, stride, alpha, frame, hop, frame / hop, Tn, sr, out,
                                                                 out_stride, o_energy, o_entropy, o_zcr, wb);
    else
      frame_walk_multi_kernel<G, false, false><<<mg, 128, 0, st>>>(pcm, stride, alpha, frame, hop, frame / hop, Tn, sr,
                                                                  out, out_stride, o_energy, o_entropy, o_zcr, wb);
    prof_end();
    SONAR_CUDA(cudaGetLastError());
    if (fuse) {
      prof_begin("rms_windows_kernel", st);
      rms_from_parts_kernel<<<dim3((unsigned)((wl->nw + 127) / 128), (unsigned)n_streams), 128, 0, st>>>(
          wl->part, wl->part_stride, groups, (int64_t)G * hop, wl->hop, (int)(wl->win / wl->hop), (int)wl->win, wl->nw,
          wl->rms, wl->rms_stride);
      prof_end();
      SONAR_CUDA(cudaGetLastError());
      if (wl_done) *wl_done = true;
    }
    return SONAR_OK;
    };
    static const bool g8 = std::getenv("SONAR_FRAME_WALK_G8") != nullptr;  // diagnostic: eight frames per thread everywhere
    if (slide && !g8) return run(std::integral_constant<int, 12>{});
    return run(std::integral_constant<int, 8>{});
  }
  if (smem > 200 * 1024)
    return set_error(SONAR_ERR_UNSUPPORTED, "energy frame / hop too long for the shared-memory tile");
#define LAUNCH_FW(E, Z)                                                                                  \
  do {                                                                                                   \
    auto k = frame_walk_kernel<E, Z>;                                                                    \
    SONAR_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    k<<<grid, kTdThreads, smem, st>>>(pcm, n, stride, alpha, frame, hop, Tn, sr, out, out_stride,       \
                                      o_energy, o_entropy, o_zcr, fpc);                                 \
  } while (0)
  prof_begin("frame_walk_kernel", st);
  if (en && zc)
    LAUNCH_FW(true, true);
  else if (en)
    LAUNCH_FW(true, false);
  else
    LAUNCH_FW(false, true);
#undef LAUNCH_FW
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

namespace {
// Zero-crossing rate of the lone incomplete frame of a stream shorter than the window (speech.go:351-357 hands
// pre[0 : min(W, N)] to zero_crossing_rate.go:37-53): one thread per stream, n < W samples, rate = crossings / (n / sr).
__global__ void short_zcr_kernel(const double* __restrict__ pcm, int64_t n, int64_t stride, int n_streams, double alpha,
                                 int sr, double* __restrict__ out, int64_t out_stride, int64_t o_zcr) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_streams) return;
  const double* __restrict__ x = pcm + (int64_t)s * stride;
  double xprev = 0.0;
  int crossings = 0;
  bool prev_neg = false;
  for (int64_t i = 0; i < n; ++i) {
    const double y = x[i] - alpha * xprev;
    xprev = x[i];
    const bool neg = y < 0.0;
    if (i > 0 && neg != prev_neg) ++crossings;
    prev_neg = neg;
  }
  double z = 0.0;
  if (n >= 2) z = (double)crossings / ((double)n / (double)sr);  // sr == 0: x / +Inf = 0, as in Go
  out[(int64_t)s * out_stride + o_zcr] = z;
}
}  // namespace

int launch_short_zcr(const double* pcm, int64_t n, int64_t stride, int n_streams, double alpha, int sr, double* out,
                     int64_t out_stride, int64_t o_zcr, cudaStream_t st) {
  if (n_streams <= 0) return SONAR_OK;
  prof_begin("frame_walk_kernel", st);
  short_zcr_kernel<<<(n_streams + 63) / 64, 64, 0, st>>>(pcm, n, stride, n_streams, alpha, sr, out, out_stride, o_zcr);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

int launch_variance(const double* x, int64_t n, int64_t stride, int n_streams, double* out, int64_t out_stride,
                    cudaStream_t st) {
  if (n_streams <= 0) return SONAR_OK;
  prof_begin("variance_kernel", st);
  variance_kernel<<<n_streams, 256, 0, st>>>(x, n, stride, out, out_stride);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

int launch_rms_windows(const double* pcm, int64_t n, int64_t stride, int n_streams, double alpha, int win,
                       int hop, int64_t nw, double* out, int64_t out_stride, cudaStream_t st) {
  if (nw <= 0 || n_streams <= 0) return SONAR_OK;
  if (hop > 0 && win % hop == 0 && win / hop <= 8) {
    dim3 bgrid((unsigned)((nw + kRbWin - 1) / kRbWin), (unsigned)n_streams);
    prof_begin("rms_windows_kernel", st);
    if (hop >= 32 * kRbWarps * 4)
      rms_blocks_kernel<<<bgrid, kRbWarps * 32, 0, st>>>(pcm, stride, alpha, win, hop, win / hop, nw, out, out_stride);
    else
      rms_blocks_warp_kernel<<<bgrid, kRbWin * 32, 0, st>>>(pcm, stride, alpha, win, hop, win / hop, nw, out, out_stride);
    prof_end();
    SONAR_CUDA(cudaGetLastError());
    return SONAR_OK;
  }
  dim3 grid((unsigned)((nw + 7) / 8), (unsigned)n_streams);
  prof_begin("rms_windows_kernel", st);
  rms_windows_kernel<<<grid, 256, 0, st>>>(pcm, n, stride, alpha, win, hop, nw, out, out_stride);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

int launch_loudness_range(double* rms, int64_t nw, int64_t in_stride, int n_streams, double* out,
                          int64_t out_stride, cudaStream_t st) {
  if (n_streams <= 0) return SONAR_OK;
  if (nw > 4096) {
    prof_begin("loudness_select_kernel", st);
    loudness_select_kernel<<<n_streams, 256, 0, st>>>(rms, nw, in_stride, out, out_stride);
    prof_end();
    SONAR_CUDA(cudaGetLastError());
    return SONAR_OK;
  }
  int np2 = 1;
  while (np2 < nw) np2 <<= 1;
  prof_begin("loudness_range_kernel", st);
  loudness_range_kernel<<<n_streams, 256, sizeof(double) * np2, st>>>(rms, nw, in_stride, out, out_stride);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

int launch_fill(double* p, int64_t n, double v, cudaStream_t st) {
  if (n <= 0) return SONAR_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  prof_begin("fill_kernel", st);
  fill_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, n, v);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

int launch_fill_strided(double* p, int64_t count, int64_t stride, int n_streams, double v, cudaStream_t st) {
  if (count <= 0 || n_streams <= 0) return SONAR_OK;
  int64_t blocks = (count + 255) / 256;
  if (blocks > 64) blocks = 64;
  prof_begin("fill_strided_kernel", st);
  fill_strided_kernel<<<dim3((unsigned)blocks, (unsigned)n_streams), 256, 0, st>>>(p, count, stride, v);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

}  // namespace sonar
