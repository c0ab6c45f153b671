// Column mean / unbiased standard deviation over frames (float64).
//
//   extractMFCCStatistics      fingerprint/comparison.go:774-800  (per coefficient: stat.Mean, sqrt(stat.Variance))
//   compareSequenceStats       fingerprint/comparison.go:827-842  (dim == 1)
//   gonum stat.Variance        two-pass with compensation: (sum d^2 - (sum d)^2 / n) / (n - 1)
//
// One CTA per column, strided (coalesced across the CTA for dim == 1, row-strided otherwise —
// the arrays are at most frames x 13) two-pass tree reduction.  Not a bit-exact output:
// similarity carries the 1e-4 tolerance, the tree order differs from gonum's by ~1e-16.
#include <cmath>

#include "common.h"

namespace sonar {
namespace {

constexpr int kCsThreads = 256;

__device__ double cs_block_sum(double v, double* red) {
  const int t = threadIdx.x;
  red[t] = v;
  __syncthreads();
  for (int o = kCsThreads / 2; o > 0; o >>= 1) {
    if (t < o) red[t] += red[t + o];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kCsThreads) colstats_kernel(const double* __restrict__ x, int64_t t, int dim,
                                                              double* __restrict__ stats) {
  __shared__ double red[kCsThreads];
  const int c = blockIdx.x;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < t; i += kCsThreads) acc += x[i * dim + c];
  const double mean = cs_block_sum(acc, red) / (double)t;
  double ss = 0.0, comp = 0.0;
  for (int64_t i = threadIdx.x; i < t; i += kCsThreads) {
    const double d = x[i * dim + c] - mean;
    ss += d * d;
    comp += d;
  }
  ss = cs_block_sum(ss, red);
  comp = cs_block_sum(comp, red);
  if (threadIdx.x == 0) {
    stats[c] = mean;
    stats[dim + c] = sqrt((ss - comp * comp / (double)t) / (double)(t - 1));  // NaN for t == 1, like gonum
  }
}

// Batched form: one CTA per (array, column) of a job table -- every statistic a Compare / BatchCompare call needs in ONE
// launch (the query's arrays appear once however many candidates follow).  The same two-pass tree reduction as above,
// so a column's result does not depend on which form computed it.
__global__ void __launch_bounds__(kCsThreads) colstats_batch_kernel(const double* __restrict__ base,
                                                                    const ColJob* __restrict__ jobs,
                                                                    double* __restrict__ out) {
  __shared__ double red[kCsThreads];
  const ColJob j = jobs[blockIdx.x];
  const double* __restrict__ x = base + j.off;
  const int64_t t = j.t;
  const int dim = j.dim;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < t; i += kCsThreads) acc += x[i * dim];
  const double mean = cs_block_sum(acc, red) / (double)t;
  double ss = 0.0, comp = 0.0;
  for (int64_t i = threadIdx.x; i < t; i += kCsThreads) {
    const double d = x[i * dim] - mean;
    ss += d * d;
    comp += d;
  }
  ss = cs_block_sum(ss, red);
  comp = cs_block_sum(comp, red);
  if (threadIdx.x == 0) {
    out[j.out_mean] = mean;
    out[j.out_std] = sqrt((ss - comp * comp / (double)t) / (double)(t - 1));
  }
}

}  // namespace

int launch_colstats_batch(const double* base, const ColJob* jobs, int n_jobs, double* out, cudaStream_t st) {
  if (n_jobs <= 0) return SONAR_OK;
  prof_begin("colstats_batch_kernel", st);
  colstats_batch_kernel<<<n_jobs, kCsThreads, 0, st>>>(base, jobs, out);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

int launch_colstats(const double* x, int64_t t, int dim, double* stats, cudaStream_t st) {
  if (t <= 0 || dim <= 0) return SONAR_OK;
  prof_begin("colstats_kernel", st);
  colstats_kernel<<<dim, kCsThreads, 0, st>>>(x, t, dim, stats);
  prof_end();
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

}  // namespace sonar
