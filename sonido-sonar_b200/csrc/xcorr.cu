// Normalised cross-correlation over +-maxLag (float64, reference summation order).
//
//   CrossCorrelation.normalize            algorithms/stats/correlation.go:464-501
//   computeTimeDomain                     :203-228
//   normalizedCrossCorrelation            :373-409
//   calculateOverlapRegion                :421-449
//   findPeak / SNR / sharpness / second peak / peak-to-sidelobe   :526-661
//
// The detected lag index must be bit-identical to the reference (north star), and the
// reference's arg-max runs over values that are each a *sequential* float64 sum.  The
// kernels here therefore reproduce that order exactly (compiled with -fmad=false):
//   * znorm_kernel: one warp per sequence; the warp stages a tile in shared memory with
//     coalesced loads, lane 0 replays the reference's left-to-right sums (mean, then
//     population variance), all lanes write the z-scores;
//   * ncc_exact_kernel: one thread per lag, i ascending; consecutive lags sit in
//     consecutive lanes so a[i] is a broadcast and b[i+lag] a coalesced load (mirrored
//     for negative lags).  Every correlation value is bit-exact, so the arg-max needs no
//     guard band;
//   * xcorr_finalize_kernel: one CTA per pair (or lag shard): first-index arg-max of |c|,
//     then the peak-relative reductions (noise power outside +-5, side lobe outside +-10,
//     second peak, neighbours of the peak).
#include <cmath>
#include <cstdlib>

#include "common.h"

namespace sonar {
namespace {

constexpr int kZTile = 2048;
constexpr int kZThreads = 128;

__device__ __forceinline__ void z_cp8(double* smem_dst, const double* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src));
}

// One CTA per sequence.  Thread 0 replays the reference's left-to-right sums; the other threads only move data:
// tiles arrive by cp.async one tile ahead of the summation, so the kernel's length is the dependent-add chain.
// With q.prefix set, the running sum of squared deviations (which the second pass forms anyway) is kept as
// P[i + 1] = sum_{k <= i} (x[k] - mean)^2 with P[n + 1] = the factor that turns it into a prefix sum of z^2
// (xcorr_fft.cu reads the denominators of the screened curve from it).
__global__ void __launch_bounds__(kZThreads) znorm_kernel(const XcorrSeq* __restrict__ seqs) {
  __shared__ double tile[2][kZTile];
  __shared__ double s_val;
  const XcorrSeq q = seqs[blockIdx.x];
  const double* __restrict__ x = q.in;
  double* __restrict__ z = q.out;
  double* __restrict__ P = q.prefix;
  const int64_t n = q.n;
  const int t = threadIdx.x;
  const int n_tiles = (int)((n + kZTile - 1) / kZTile);
  auto count = [&](int k) { return (int)((n - (int64_t)k * kZTile < kZTile) ? (n - (int64_t)k * kZTile) : kZTile); };
  auto stage = [&](int k) {
    const int cnt = count(k);
    double* dst = tile[k & 1];
    const double* src = x + (int64_t)k * kZTile;
    for (int e = t; e < cnt; e += kZThreads) z_cp8(dst + e, src + e);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto arrive = [&](int k) {  // tile k is in shared memory for everybody; tile k + 1 is on its way
    if (k + 1 < n_tiles) {
      stage(k + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
  };
  double acc = 0.0;
  if (n_tiles > 0) stage(0);
  for (int k = 0; k < n_tiles; ++k) {
    arrive(k);
    if (t == 0) {
      const double* c = tile[k & 1];
      const int cnt = count(k);
#pragma unroll 8
      for (int i = 0; i < cnt; ++i) acc += c[i];
    }
    __syncthreads();  // the buffer is restaged in the next iteration
  }
  if (t == 0) s_val = acc / (double)n;
  __syncthreads();
  const double mean = s_val;
  acc = 0.0;
  if (n_tiles > 0) stage(0);
  for (int k = 0; k < n_tiles; ++k) {
    arrive(k);
    double* c = tile[k & 1];
    const int cnt = count(k);
    for (int i = t; i < cnt; i += kZThreads) {
      const double d = c[i] - mean;
      c[i] = d * d;
    }
    __syncthreads();
    if (t == 0) {
      if (P) {
#pragma unroll 8
        for (int i = 0; i < cnt; ++i) {
          acc += c[i];
          c[i] = acc;  // the tile turns into its own running sums
        }
      } else {
#pragma unroll 8
        for (int i = 0; i < cnt; ++i) acc += c[i];
      }
    }
    __syncthreads();
    if (P) {
      for (int i = t; i < cnt; i += kZThreads) P[(int64_t)k * kZTile + i + 1] = c[i];
      __syncthreads();  // before the buffer is restaged
    }
  }
  __syncthreads();
  if (t == 0) s_val = acc / (double)n;
  __syncthreads();
  const double var = s_val;
  const double sd = sqrt(var);
  if (sd < 1e-10) {
#pragma unroll 4
    for (int64_t i = t; i < n; i += kZThreads) z[i] = x[i] - mean;
  } else {
#pragma unroll 4
    for (int64_t i = t; i < n; i += kZThreads) z[i] = (x[i] - mean) / sd;
  }
  if (P && t == 0) {
    P[0] = 0.0;
    P[n + 1] = sd < 1e-10 ? 1.0 : 1.0 / var;
  }
}

constexpr int kNccThreads = 128;

__global__ void __launch_bounds__(kNccThreads) ncc_exact_kernel(const XcorrPair* __restrict__ pairs) {
  const XcorrPair p = pairs[blockIdx.y];
  const int64_t j = p.idx_lo + (int64_t)blockIdx.x * kNccThreads + threadIdx.x;
  if (j >= p.idx_hi) return;
  const int64_t lag = j - p.aml;
  const int64_t na = p.na, nb = p.nb;
  // calculateOverlapRegion (correlation.go:421-449)
  int64_t s1, s2, len;
  if (lag >= 0) {
    s1 = 0;
    s2 = lag;
    len = na < nb - lag ? na : nb - lag;
  } else {
    s1 = -lag;
    s2 = 0;
    const int64_t e2 = nb < na + lag ? nb : na + lag;
    len = (na + lag) < e2 ? (na + lag) : e2;
  }
  double c = 0.0;
  if (len > 0) {
    const double* __restrict__ a = p.za + s1;
    const double* __restrict__ b = p.zb + s2;
    double sum = 0.0, q1 = 0.0, q2 = 0.0;
#pragma unroll 4
    for (int64_t i = 0; i < len; ++i) {
      const double v1 = a[i], v2 = b[i];
      sum += v1 * v2;
      q1 += v1 * v1;
      q2 += v2 * v2;
    }
    const double den = sqrt(q1 * q2);
    c = den < 1e-10 ? 0.0 : sum / den;
  }
  p.corr[j - p.idx_lo] = c;
}

// ------------------------------------------------------------------------------------------------
// Tiled form of the same sums (default).  ncc_exact_kernel issues two global loads per (lag, i) and is bound
// by the L1 data path (ncu: l1tex 88 %, FP64 62 %).  Here a CTA owns kNtT * kNtR consecutive lags of one sign
// and walks i in chunks staged in shared memory (cp.async, double buffered); a thread owns kNtR CONSECUTIVE
// lags, so the shifted sequence slides through a register window: one conflict-free shared load (odd stride)
// and one broadcast load per i feed kNtR products.  Sums stay sequential in i per lag, i.e. bit-identical:
//   lag >= 0:  u = a, v = b;   lag < 0:  u = b, v = a, lag' = -lag   (a*b and q1*q2 commute exactly)
//   c = sum_i u[i] v[lag'+i] / sqrt(sum_i u[i]^2 * sum_i v[lag'+i]^2),  i < len = min(nu, nv - lag')
// sum_i u[i]^2 is one running sum per thread, sampled when each lag's overlap ends; squares of v are taken
// once when a value enters the window.  The last (< 2 kNtR) terms of each lag are added from global memory.
// Two shapes: <64, 5> (320 lags per CTA) evaluates whole curves at FP64-pipe throughput; <64, 1> (64 lags per CTA,
// kNccFlagLags) evaluates the few blocks the screen (xcorr_fft.cu) flags, where only the length of the dependent
// chain matters and a finer block wastes less work.
constexpr int kNtChunk = 640;            // i per staged chunk (multiple of every R in use)

__device__ __forceinline__ void ncc_cp8(double* smem_dst, const double* gmem_src, bool ok) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int nbytes = ok ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gmem_src), "r"(nbytes));
}

template <int kNtT /* threads per CTA */, int kNtR /* lags per thread; odd or 1: conflict-free strided loads */>
__global__ void __launch_bounds__(kNtT) ncc_tiled_kernel(const XcorrPair* __restrict__ pairs, int blocks_per_sign,
                                                         const unsigned char* __restrict__ need, int need_stride) {
  constexpr int kNtLags = kNtT * kNtR;      // lags per CTA
  constexpr int kNtV = kNtChunk + kNtLags;  // staged values of the shifted sequence
  static_assert(kNtChunk % kNtR == 0, "chunk must hold whole window rotations");
  __shared__ double su[2][kNtChunk];
  __shared__ double sv[2][kNtV];
  if (need && !need[(size_t)blockIdx.y * need_stride + blockIdx.x]) return;  // screened out (xcorr_fft.cu)
  const XcorrPair p = pairs[blockIdx.y];
  const bool neg = (int)blockIdx.x >= blocks_per_sign;
  const int blk = neg ? (int)blockIdx.x - blocks_per_sign : (int)blockIdx.x;
  // lags (as l = |lag|) of this sign inside the shard [idx_lo, idx_hi): j = aml + lag
  int64_t l_lo, l_hi;  // [l_lo, l_hi)
  if (!neg) {
    l_lo = p.idx_lo - p.aml > 0 ? p.idx_lo - p.aml : 0;
    l_hi = p.idx_hi - p.aml;
  } else {
    l_lo = p.aml - p.idx_hi + 1 > 1 ? p.aml - p.idx_hi + 1 : 1;
    l_hi = p.aml - p.idx_lo + 1;
  }
  const int64_t L0 = l_lo + (int64_t)blk * kNtLags;
  if (L0 >= l_hi) return;
  const double* __restrict__ u = neg ? p.zb : p.za;
  const double* __restrict__ v = neg ? p.za : p.zb;
  const int64_t nu = neg ? p.nb : p.na, nv = neg ? p.na : p.nb;
  const int t = threadIdx.x;
  int64_t len[kNtR];
#pragma unroll
  for (int r = 0; r < kNtR; ++r) {
    const int64_t l = L0 + (int64_t)kNtR * t + r;
    // lags past the end of the shard are computed like the others (the staged tiles are zero filled) and simply
    // not stored: a thread that straddles the end must not fall back to the slow tail for its valid lags
    const int64_t n = nu < nv - l ? nu : nv - l;
    len[r] = n > 0 ? n : 0;
  }
  const int64_t i_main = (len[kNtR - 1] / kNtR) * kNtR;  // all kNtR lags of the thread overlap on [0, i_main)
  int64_t cta_len = nu < nv - L0 ? nu : nv - L0;          // longest overlap in the CTA
  if (cta_len < 0) cta_len = 0;
  const int n_chunks = (int)((cta_len + kNtChunk - 1) / kNtChunk);
  auto stage = [&](int c) {
    const int64_t i0 = (int64_t)c * kNtChunk;
    double* du = su[c & 1];
    double* dv = sv[c & 1];
    for (int e = t; e < kNtChunk; e += kNtT) ncc_cp8(du + e, u + (i0 + e < nu ? i0 + e : 0), i0 + e < nu);
    for (int e = t; e < kNtV; e += kNtT) {
      const int64_t g = L0 + i0 + e;
      ncc_cp8(dv + e, v + (g < nv ? g : 0), g < nv);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  double sum[kNtR], qv[kNtR], qu[kNtR], run = 0.0;
#pragma unroll
  for (int r = 0; r < kNtR; ++r) sum[r] = qv[r] = qu[r] = 0.0;
  if (n_chunks > 0) stage(0);
  for (int c = 0; c < n_chunks; ++c) {
    if (c + 1 < n_chunks) {
      stage(c + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int64_t i0 = (int64_t)c * kNtChunk;
    const double* __restrict__ cu = su[c & 1];
    const double* __restrict__ cv = sv[c & 1] + kNtR * t;
    int64_t iend = i_main - i0;
    if (iend > kNtChunk) iend = kNtChunk;
    if (iend > 0) {
      double w[kNtR], w2[kNtR];
#pragma unroll
      for (int r = 0; r < kNtR; ++r) {
        w[r] = cv[r];
        w2[r] = w[r] * w[r];
      }
      for (int ii = 0; ii < (int)iend; ii += kNtR) {
#pragma unroll
        for (int s5 = 0; s5 < kNtR; ++s5) {  // window slot (s5 + r) % kNtR holds v[l_r + i]
          const double uv = cu[ii + s5];
          const double nv1 = cv[ii + s5 + kNtR];
          run += uv * uv;
#pragma unroll
          for (int r = 0; r < kNtR; ++r) {
            sum[r] += uv * w[(s5 + r) % kNtR];
            qv[r] += w2[(s5 + r) % kNtR];
          }
          w[s5] = nv1;  // the slot lag 0 just used now holds lag kNtR-1's next value
          w2[s5] = nv1 * nv1;
        }
      }
    }
    __syncthreads();  // the buffer is restaged two chunks later
  }
  // ---- tails: i in [i_main, len_r), at most 2 kNtR - 2 terms per lag, straight from global memory ----
  if (len[0] > 0) {
#pragma unroll
    for (int r = 0; r < kNtR; ++r)
      if (len[r] == i_main) qu[r] = run;
    for (int64_t i = i_main; i < len[0]; ++i) {
      const double uv = u[i];
      run += uv * uv;
#pragma unroll
      for (int r = 0; r < kNtR; ++r) {
        if (i < len[r]) {
          const double vv = v[L0 + (int64_t)kNtR * t + r + i];
          sum[r] += uv * vv;
          qv[r] += vv * vv;
        }
        if (i + 1 == len[r]) qu[r] = run;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kNtR; ++r) {
    const int64_t l = L0 + (int64_t)kNtR * t + r;
    if (l >= l_hi) continue;
    double c = 0.0;
    if (len[r] > 0) {
      const double den = sqrt(qu[r] * qv[r]);
      c = den < 1e-10 ? 0.0 : sum[r] / den;
    }
    const int64_t j = neg ? p.aml - l : p.aml + l;
    p.corr[j - p.idx_lo] = c;
  }
}

struct PeakKey {
  double a;
  int64_t i;
};
__device__ __forceinline__ bool better(double a, int64_t i, double b, int64_t k) {
  // findPeak scans ascending with strict '>' on |c|: larger |c| wins, ties -> smaller index
  if (k < 0) return i >= 0;
  if (i < 0) return false;
  return a > b || (a == b && i < k);
}

constexpr int kFinThreads = 256;

__device__ void block_best(double& a, int64_t& i, double* sa, int64_t* si) {
  const int t = threadIdx.x;
  sa[t] = a;
  si[t] = i;
  __syncthreads();
  for (int o = kFinThreads / 2; o > 0; o >>= 1) {
    if (t < o && better(sa[t + o], si[t + o], sa[t], si[t])) {
      sa[t] = sa[t + o];
      si[t] = si[t + o];
    }
    __syncthreads();
  }
  a = sa[0];
  i = si[0];
  __syncthreads();
}

__device__ double block_sum(double v, double* sa) {
  const int t = threadIdx.x;
  sa[t] = v;
  __syncthreads();
  for (int o = kFinThreads / 2; o > 0; o >>= 1) {
    if (t < o) sa[t] += sa[t + o];
    __syncthreads();
  }
  const double r = sa[0];
  __syncthreads();
  return r;
}

__device__ double block_max(double v, double* sa) {
  const int t = threadIdx.x;
  sa[t] = v;
  __syncthreads();
  for (int o = kFinThreads / 2; o > 0; o >>= 1) {
    if (t < o) sa[t] = fmax(sa[t], sa[t + o]);
    __syncthreads();
  }
  const double r = sa[0];
  __syncthreads();
  return r;
}

// peak_override < 0: use the shard's own arg-max; otherwise the given global index.
__global__ void __launch_bounds__(kFinThreads) xcorr_finalize_kernel(const XcorrPair* __restrict__ pairs,
                                                                     int64_t peak_override,
                                                                     XcorrPairOut* __restrict__ outs) {
  __shared__ double sa[kFinThreads];
  __shared__ int64_t si[kFinThreads];
  const XcorrPair p = pairs[blockIdx.x];
  const double* __restrict__ c = p.corr;
  const int64_t lo = p.idx_lo, hi = p.idx_hi;
  const int t = threadIdx.x;
  // ---- local arg-max of |c| ----
  double ba = 0.0;
  int64_t bi = -1;
  for (int64_t i = lo + t; i < hi; i += kFinThreads) {
    const double v = fabs(c[i - lo]);
    if (better(v, i, ba, bi)) {
      ba = v;
      bi = i;
    }
  }
  block_best(ba, bi, sa, si);
  const int64_t local_peak = bi;
  const int64_t pk = peak_override >= 0 ? peak_override : local_peak;
  // ---- peak-relative reductions ----
  double ns = 0.0, nc = 0.0, ms = 0.0, s2a = 0.0;
  int64_t s2i = -1;
  for (int64_t i = lo + t; i < hi; i += kFinThreads) {
    const double v = c[i - lo], av = fabs(v);
    const int64_t d = i > pk ? i - pk : pk - i;
    if (d > 5) {
      ns += v * v;
      nc += 1.0;
    }
    if (d > 10) ms = fmax(ms, av);
    // findSecondPeak starts from 0.0 with strict '>', so zeros never qualify
    if (i != pk && av > 0.0 && better(av, i, s2a, s2i)) {
      s2a = av;
      s2i = i;
    }
  }
  ns = block_sum(ns, sa);
  nc = block_sum(nc, sa);
  ms = block_max(ms, sa);
  block_best(s2a, s2i, sa, si);
  if (t == 0) {
    XcorrPairOut o;
    o.peak_index = local_peak;
    o.peak = (local_peak >= 0) ? c[local_peak - lo] : 0.0;
    o.noise_sum = ns;
    o.noise_cnt = nc;
    o.max_sidelobe = ms;
    o.second_abs = s2a;
    o.second_index = s2i;
    o.second_val = s2i >= 0 ? c[s2i - lo] : 0.0;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    o.c_peak = (pk >= lo && pk < hi) ? c[pk - lo] : nan;
    o.c_prev = (pk - 1 >= lo && pk - 1 < hi) ? c[pk - 1 - lo] : nan;
    o.c_next = (pk + 1 >= lo && pk + 1 < hi) ? c[pk + 1 - lo] : nan;
    o.n_candidates = (int32_t)((hi - lo) > 0x7fffffff ? 0x7fffffff : (hi - lo));
    o.pad = 0;
    outs[blockIdx.x] = o;
  }
}

__global__ void xcorr_trim_kernel(const XcorrSeq* __restrict__ seqs, const XcorrPair* __restrict__ pairs,
                                  const XcorrPairOut* __restrict__ outs, int n_pairs, const double** __restrict__ qptr,
                                  const double** __restrict__ rptr) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const int64_t lag = outs[p].peak_index - pairs[p].aml;
  qptr[p] = seqs[2 * p].in + (lag < 0 ? -lag : 0);
  rptr[p] = seqs[2 * p + 1].in + (lag >= 0 ? lag : 0);
}

}  // namespace

int launch_xcorr_trim(const XcorrSeq* seqs_dev, const XcorrPair* pairs_dev, const XcorrPairOut* outs_dev, int n_pairs,
                      const double** qptr_dev, const double** rptr_dev, cudaStream_t st) {
  if (n_pairs <= 0) return SONAR_OK;
  prof_begin("xcorr_trim_kernel", st);
  xcorr_trim_kernel<<<(n_pairs + 127) / 128, 128, 0, st>>>(seqs_dev, pairs_dev, outs_dev, n_pairs, qptr_dev, rptr_dev);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

int launch_znorm(const XcorrSeq* seqs_dev, int count, cudaStream_t st) {
  if (count <= 0) return SONAR_OK;
  prof_begin("znorm_kernel", st);
  znorm_kernel<<<count, kZThreads, 0, st>>>(seqs_dev);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

int launch_xcorr(const XcorrPair* pairs_dev, int n_pairs, int64_t max_shard_lags, cudaStream_t st) {
  if (n_pairs <= 0 || max_shard_lags <= 0) return SONAR_OK;
  static const bool direct = std::getenv("SONAR_NCC_DIRECT") != nullptr;  // diagnostic: one thread per lag
  prof_begin("ncc_exact_kernel", st);
  if (direct) {
    dim3 grid((unsigned)((max_shard_lags + kNccThreads - 1) / kNccThreads), (unsigned)n_pairs);
    ncc_exact_kernel<<<grid, kNccThreads, 0, st>>>(pairs_dev);
  } else {
    // a shard holds at most max_shard_lags lags of either sign
    const int bps = (int)((max_shard_lags + 319) / 320);
    ncc_tiled_kernel<64, 5><<<dim3((unsigned)(2 * bps), (unsigned)n_pairs), 64, 0, st>>>(pairs_dev, bps, nullptr, 0);
  }
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

int launch_xcorr_flagged(const XcorrPair* pairs_dev, int n_pairs, int64_t max_shard_lags, const unsigned char* need_dev,
                         int need_stride, cudaStream_t st) {
  if (n_pairs <= 0 || max_shard_lags <= 0) return SONAR_OK;
  const int bps = (int)((max_shard_lags + kNccFlagLags - 1) / kNccFlagLags);
  prof_begin("ncc_exact_kernel", st);
  ncc_tiled_kernel<64, 1><<<dim3((unsigned)(2 * bps), (unsigned)n_pairs), 64, 0, st>>>(pairs_dev, bps, need_dev, need_stride);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

int launch_xcorr_finalize(const XcorrPair* pairs_dev, int n_pairs, int64_t peak_override, XcorrPairOut* outs_dev,
                          cudaStream_t st) {
  if (n_pairs <= 0) return SONAR_OK;
  prof_begin("xcorr_finalize_kernel", st);
  xcorr_finalize_kernel<<<n_pairs, kFinThreads, 0, st>>>(pairs_dev, peak_override, outs_dev);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

}  // namespace sonar
