// C-ABI entry points of the alignment and comparison paths: staging, kernel sequencing,
// sharding of pair batches over the context's devices, and the O(1) scalar formulas that
// turn kernel results into the reference's result structs.  The signal-sized arithmetic
// lives in xcorr.cu / dtw.cu / colstats.cu.
//
//   CrossCorrelation.Compute          algorithms/stats/correlation.go:131-200
//   AlignmentAnalyzer.alignWithCrossCorrelation   algorithms/stats/alignment.go:151-181
//   alignWithFeatures (lag clamp)     fingerprint/extractors/alignment.go:357-409
//   DTWAlignment.Align                algorithms/stats/dtw.go:55-103
//   FingerprintComparator.Compare     fingerprint/comparison.go:133-194,266-341,646-882,1011-1037
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <cstdlib>
#include <thread>

#include "common.h"

namespace sonar {
namespace {

const double kInf = std::numeric_limits<double>::infinity();
const double kNaN = std::numeric_limits<double>::quiet_NaN();

inline int64_t even(int64_t n) { return (n + 1) & ~(int64_t)1; }

// Peak-relative partials -> CorrelationResult scalars (correlation.go:526-667).
void summarize(double c0, double cm, double cp, double ns, double nc, double ms, double second_val, int64_t gp,
               int aml, int64_t na, int64_t nb, int64_t n_eval, sonar_xcorr_summary* o) {
  const int64_t nl = 2 * (int64_t)aml + 1;
  o->peak_correlation = c0;
  o->peak_index = (int32_t)gp;
  o->peak_lag = (int32_t)(gp - aml);
  o->actual_max_lag = aml;
  if (nc == 0) {
    o->snr = 0.0;  // calculateSNR :572-601
  } else {
    const double nlv = std::sqrt(ns / nc);
    o->snr = nlv < 1e-10 ? kInf : 20.0 * std::log10(std::fabs(c0) / nlv);
  }
  o->sharpness = (nl < 3 || gp <= 0 || gp >= nl - 1) ? 0.0 : -(cp - 2 * c0 + cm);  // :611-619
  o->second_peak = second_val;                                                      // :622-636
  o->peak_to_sidelobe = ms < 1e-10 ? kInf : 20.0 * std::log10(std::fabs(c0) / ms);  // :639-661
  xcorr_derive(o, na, nb);  // p-value, significance, overlap length
  o->n_candidates = (int32_t)std::min<int64_t>(n_eval, 0x7fffffff);
}

}  // namespace

void summarize_xcorr(const XcorrPairOut& o, int aml, int64_t na, int64_t nb, int64_t n_eval, sonar_xcorr_summary* s) {
  summarize(o.c_peak, o.c_prev, o.c_next, o.noise_sum, o.noise_cnt, o.max_sidelobe, o.second_val, o.peak_index, aml, na,
            nb, n_eval, s);
}

// AlignmentAnalyzer.alignWithCrossCorrelation's scalars (algorithms/stats/alignment.go:151-181)
void fill_align_from_xcorr(const sonar_xcorr_summary* xc, int64_t nq, int64_t nr, int max_lag, int hop, int sr,
                           sonar_align_result* out) {
  std::memset(out, 0, sizeof(*out));
  out->method = 1;
  out->query_length = (int32_t)nq;
  out->reference_length = (int32_t)nr;
  out->sample_rate = sr;
  out->offset = xc->peak_lag * hop;                                                    // :164
  out->offset_seconds = (double)out->offset / (double)sr;                              // :165
  out->similarity = std::fmin(1.0, std::fmax(0.0, std::fabs(xc->peak_correlation)));  // :171-174
  out->confidence = corr_confidence(xc);
  out->alignment_quality = corr_quality(xc, max_lag);
  out->noise_level = 1.0 - xc->snr / 20.0;  // :178
}

namespace {

struct PairJob {
  const double* a;
  int64_t na;
  const double* b;
  int64_t nb;
  double* corr;  // host, nullable
  sonar_xcorr_summary* out;
};

// One chunk of pairs on one device slot.  host_inputs: a/b are host pointers (copied in);
// otherwise they are device pointers on this device.
int run_xcorr_chunk(DevCtx& dev, Slot& s, const PairJob* jobs, int np, int max_lag, bool host_inputs,
                    double* corr_dev_out /* nullable: device-resident destination, packed 2*aml+1 per pair */) {
  if (np <= 0) return SONAR_OK;
  size_t in_d = 0, z_d = 0, corr_d = 0;
  int64_t max_lags = 0;
  std::vector<int> aml(np);
  for (int p = 0; p < np; p++) {
    if (!jobs[p].a || !jobs[p].b || jobs[p].na <= 0 || jobs[p].nb <= 0)
      return set_error(SONAR_ERR_EMPTY, "empty signals provided");  // correlation.go:133
    aml[p] = actual_max_lag(max_lag, jobs[p].na, jobs[p].nb);
    const int64_t nl = 2 * (int64_t)aml[p] + 1;
    in_d += (size_t)(even(jobs[p].na) + even(jobs[p].nb));
    corr_d += (size_t)even(nl);
    max_lags = std::max(max_lags, nl);
  }
  z_d = in_d;
  const size_t desc_bytes = sizeof(XcorrSeq) * 2 * np + sizeof(XcorrPair) * np + sizeof(XcorrPairOut) * np;
  // Nobody takes the curve home: screen the lags with an FFT and evaluate only the candidate blocks in reference
  // order (xcorr_fft.cu) -- same peak, same value, same second peak.
  static const bool full_env = std::getenv("SONAR_NCC_FULL") != nullptr;
  bool screen = !corr_dev_out && !full_env;
  int64_t max_n = 0;
  int aml_max = 0;
  for (int p = 0; p < np && screen; p++) {
    screen = jobs[p].corr == nullptr;
    max_n = std::max<int64_t>(max_n, std::max(jobs[p].na, jobs[p].nb));
    aml_max = std::max(aml_max, aml[p]);
  }
  XcorrScreen xs{};
  if (screen) {
    xs = xcorr_screen_geom(max_n, aml_max, max_lags, np);
    if (xs.bytes > ((size_t)8 << 30)) screen = false;
  }
  const size_t screen_off = (sizeof(double) * z_d + desc_bytes + 64 + 255) & ~(size_t)255;
  int rc;
  if (host_inputs && (rc = dev.ensure_dev(s.d_in, sizeof(double) * in_d))) return rc;
  if ((rc = dev.ensure_dev(s.d_tmp, screen ? screen_off + xs.bytes : sizeof(double) * z_d + desc_bytes + 64))) return rc;
  unsigned char* d_screen = static_cast<unsigned char*>(s.d_tmp.p) + screen_off;
  if (!corr_dev_out && (rc = dev.ensure_dev(s.d_out, sizeof(double) * corr_d))) return rc;

  double* d_in = static_cast<double*>(s.d_in.p);
  double* d_z = static_cast<double*>(s.d_tmp.p);
  unsigned char* d_desc = reinterpret_cast<unsigned char*>(d_z + z_d);
  XcorrSeq* d_seqs = reinterpret_cast<XcorrSeq*>(d_desc);
  XcorrPair* d_pairs = reinterpret_cast<XcorrPair*>(d_desc + sizeof(XcorrSeq) * 2 * np);
  XcorrPairOut* d_outs = reinterpret_cast<XcorrPairOut*>(d_desc + sizeof(XcorrSeq) * 2 * np + sizeof(XcorrPair) * np);
  double* d_corr = corr_dev_out ? corr_dev_out : static_cast<double*>(s.d_out.p);

  std::vector<XcorrSeq> seqs(2 * (size_t)np);
  std::vector<XcorrPair> pairs(np);
  size_t io = 0, co = 0;
  for (int p = 0; p < np; p++) {
    const PairJob& j = jobs[p];
    const double *a_dev, *b_dev;
    double* za = d_z + io;
    double* zb = d_z + io + even(j.na);
    if (host_inputs) {
      double* ad = d_in + io;
      double* bd = d_in + io + even(j.na);
      SONAR_CUDA(cudaMemcpyAsync(ad, j.a, sizeof(double) * (size_t)j.na, cudaMemcpyHostToDevice, s.st));
      SONAR_CUDA(cudaMemcpyAsync(bd, j.b, sizeof(double) * (size_t)j.nb, cudaMemcpyHostToDevice, s.st));
      a_dev = ad;
      b_dev = bd;
    } else {
      a_dev = j.a;
      b_dev = j.b;
    }
    io += (size_t)(even(j.na) + even(j.nb));
    seqs[2 * p] = XcorrSeq{a_dev, za, j.na, screen ? xcorr_screen_prefix(xs, d_screen, 2 * p) : nullptr};
    seqs[2 * p + 1] = XcorrSeq{b_dev, zb, j.nb, screen ? xcorr_screen_prefix(xs, d_screen, 2 * p + 1) : nullptr};
    const int64_t nl = 2 * (int64_t)aml[p] + 1;
    XcorrPair& pr = pairs[p];
    pr.za = za;
    pr.zb = zb;
    pr.corr = d_corr + co;
    pr.na = j.na;
    pr.nb = j.nb;
    pr.idx_lo = 0;
    pr.idx_hi = nl;
    pr.aml = aml[p];
    pr.pad = 0;
    co += corr_dev_out ? (size_t)nl : (size_t)even(nl);
  }
  SONAR_CUDA(cudaMemcpyAsync(d_seqs, seqs.data(), sizeof(XcorrSeq) * seqs.size(), cudaMemcpyHostToDevice, s.st));
  SONAR_CUDA(cudaMemcpyAsync(d_pairs, pairs.data(), sizeof(XcorrPair) * pairs.size(), cudaMemcpyHostToDevice, s.st));
  if ((rc = launch_znorm(d_seqs, 2 * np, s.st))) return rc;
  rc = screen ? launch_xcorr_screened(d_pairs, np, max_lags, xs, d_screen, s.st) : launch_xcorr(d_pairs, np, max_lags, s.st);
  if (rc) return rc;
  if ((rc = launch_xcorr_finalize(d_pairs, np, -1, d_outs, s.st))) return rc;
  std::vector<XcorrPairOut> outs(np);
  SONAR_CUDA(cudaMemcpyAsync(outs.data(), d_outs, sizeof(XcorrPairOut) * np, cudaMemcpyDeviceToHost, s.st));
  for (int p = 0; p < np; p++)
    if (jobs[p].corr)
      SONAR_CUDA(cudaMemcpyAsync(jobs[p].corr, pairs[p].corr, sizeof(double) * (size_t)(pairs[p].idx_hi),
                                 cudaMemcpyDeviceToHost, s.st));
  SONAR_CUDA(cudaStreamSynchronize(s.st));
  for (int p = 0; p < np; p++) {
    const XcorrPairOut& o = outs[p];
    if (jobs[p].out)
      summarize(o.c_peak, o.c_prev, o.c_next, o.noise_sum, o.noise_cnt, o.max_sidelobe, o.second_val, o.peak_index,
                aml[p], jobs[p].na, jobs[p].nb, pairs[p].idx_hi, jobs[p].out);
  }
  return SONAR_OK;
}

constexpr size_t kXcorrChunkBytes = (size_t)1 << 30;

struct DevJob {
  int rc = SONAR_OK;
  std::string err;
};

void run_xcorr_device(sonar_ctx* ctx, DevCtx* dev, const std::vector<PairJob>* jobs, int max_lag, DevJob* res) {
  set_current_ctx(ctx);
  auto fail = [&](int rc) {
    res->rc = rc;
    res->err = sonar_last_error();
  };
  cudaError_t e = cudaSetDevice(dev->device);
  if (e != cudaSuccess) return fail(cuda_error(e, "cudaSetDevice"));
  size_t i = 0;
  while (i < jobs->size()) {
    size_t k = i, bytes = 0;
    while (k < jobs->size()) {
      const size_t b = sizeof(double) * (size_t)(2 * ((*jobs)[k].na + (*jobs)[k].nb) + 2 * (int64_t)max_lag + 8);
      if (k > i && (bytes + b > kXcorrChunkBytes || k - i >= 32768)) break;  // grid.y carries the pair index
      bytes += b;
      k++;
    }
    const int rc = run_xcorr_chunk(*dev, dev->slot[0], jobs->data() + i, (int)(k - i), max_lag, true, nullptr);
    if (rc) return fail(rc);
    i = k;
  }
}

int xcorr_pairs(sonar_ctx* ctx, const std::vector<PairJob>& all, int max_lag) {
  const int nd = (int)ctx->devs.size();
  std::vector<std::vector<PairJob>> per(nd);
  for (size_t p = 0; p < all.size(); p++) per[p % nd].push_back(all[p]);
  std::vector<DevJob> res(nd);
  if (nd == 1) {
    run_xcorr_device(ctx, &ctx->devs[0], &per[0], max_lag, &res[0]);
  } else {
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++) th.emplace_back(run_xcorr_device, ctx, &ctx->devs[d], &per[d], max_lag, &res[d]);
    for (auto& t : th) t.join();
    cudaSetDevice(ctx->devs[0].device);
  }
  for (auto& r : res)
    if (r.rc) return set_error(r.rc, r.err);
  return SONAR_OK;
}

// ---- DTW -------------------------------------------------------------------

int dtw_validate(const double* q, int n, const double* r, int m, int dim, int step, int metric) {
  if (!q || !r || n <= 0 || m <= 0) return set_error(SONAR_ERR_EMPTY, "empty sequences provided");  // dtw.go:57
  if (metric != SONAR_METRIC_EUCLIDEAN)
    return set_error(SONAR_ERR_UNSUPPORTED, "only the Euclidean metric is on this path");
  if (step < 0 || step > 2) return set_error(SONAR_ERR_INVALID, "unknown step pattern");  // dtw.go:160
  if (dim <= 0) return set_error(SONAR_ERR_INVALID, "dim must be positive");
  return SONAR_OK;
}

// pairs [p0, p0+np) of a uniform batch on one device slot
int run_dtw_chunk(DevCtx& dev, Slot& s, const double* const* q, const double* const* r, int np, const DtwGeom& g,
                  int dim, int step, sonar_dtw_out* outs) {
  const int n = g.n, m = g.m;
  const int64_t cap = (int64_t)n + m;
  const size_t q_d = (size_t)np * n * dim, r_d = (size_t)np * m * dim;
  const bool need_line = sizeof(double) * (size_t)(g.n_off + 2) > 140 * 1024;
  const size_t line_d = need_line ? (size_t)np * (g.n_off + 2) : 0;
  const size_t cells_d = (size_t)np * (size_t)g.cells;
  const size_t path_bytes = (size_t)np * cap * (sizeof(int32_t) * 2 + sizeof(double)) + sizeof(DtwPairOut) * np;
  int rc;
  if ((rc = dev.ensure_dev(s.d_in, sizeof(double) * (q_d + r_d))) ||
      (rc = dev.ensure_dev(s.d_tmp, sizeof(double) * (cells_d + line_d))) ||
      (rc = dev.ensure_dev(s.d_out, path_bytes + 64)))
    return rc;
  double* d_q = static_cast<double*>(s.d_in.p);
  double* d_r = d_q + q_d;
  double* d_cells = static_cast<double*>(s.d_tmp.p);
  double* d_line = need_line ? d_cells + cells_d : nullptr;
  double* d_pc = static_cast<double*>(s.d_out.p);
  int32_t* d_pq = reinterpret_cast<int32_t*>(d_pc + (size_t)np * cap);
  int32_t* d_pr = d_pq + (size_t)np * cap;
  DtwPairOut* d_po = reinterpret_cast<DtwPairOut*>(d_pr + (size_t)np * cap);  // 2*np*cap int32 = 8-byte multiple
  for (int p = 0; p < np; p++) {
    SONAR_CUDA(cudaMemcpyAsync(d_q + (size_t)p * n * dim, q[p], sizeof(double) * (size_t)n * dim,
                               cudaMemcpyHostToDevice, s.st));
    SONAR_CUDA(cudaMemcpyAsync(d_r + (size_t)p * m * dim, r[p], sizeof(double) * (size_t)m * dim,
                               cudaMemcpyHostToDevice, s.st));
  }
  if ((rc = launch_dtw(d_q, d_r, np, g, dim, step, d_cells, d_line, d_pq, d_pr, d_pc, cap, d_po, s.st))) return rc;
  std::vector<DtwPairOut> po(np);
  SONAR_CUDA(cudaMemcpyAsync(po.data(), d_po, sizeof(DtwPairOut) * np, cudaMemcpyDeviceToHost, s.st));
  SONAR_CUDA(cudaStreamSynchronize(s.st));
  bool cap_fail = false;
  for (int p = 0; p < np; p++) {
    sonar_dtw_out& o = outs[p];
    const int64_t len = po[p].path_len;
    o.path_len = len;
    o.total_cost = po[p].total_cost;
    o.distance = po[p].total_cost / (double)len;  // dtw.go:88-91
    const int64_t take = std::min<int64_t>(len, o.path_cap);
    const size_t off = (size_t)p * cap + (size_t)(cap - len);
    if (take > 0) {
      if (o.path_query)
        SONAR_CUDA(cudaMemcpyAsync(o.path_query, d_pq + off, sizeof(int32_t) * take, cudaMemcpyDeviceToHost, s.st));
      if (o.path_ref)
        SONAR_CUDA(cudaMemcpyAsync(o.path_ref, d_pr + off, sizeof(int32_t) * take, cudaMemcpyDeviceToHost, s.st));
      if (o.path_cost)
        SONAR_CUDA(cudaMemcpyAsync(o.path_cost, d_pc + off, sizeof(double) * take, cudaMemcpyDeviceToHost, s.st));
    }
    if (len > o.path_cap && (o.path_query || o.path_ref || o.path_cost)) cap_fail = true;
  }
  SONAR_CUDA(cudaStreamSynchronize(s.st));
  for (int p = 0; p < np; p++) {
    if (!outs[p].cost_matrix) continue;  // CostMatrix export (dtw.go:96): opt-in, O(n*m)
    const size_t full = (size_t)n * (m + 1);
    if ((rc = dev.ensure_dev(s.d_out, path_bytes + 64 + sizeof(double) * full))) return rc;
    // ensure_dev may have reallocated d_out: the paths were already copied out, only the matrix lives there now
    double* d_full = static_cast<double*>(s.d_out.p);
    DtwGeom g1 = g;
    if ((rc = launch_dtw_expand(d_cells + (size_t)p * g.cells, g1, d_full, s.st))) return rc;
    SONAR_CUDA(cudaMemcpyAsync(outs[p].cost_matrix, d_full, sizeof(double) * full, cudaMemcpyDeviceToHost, s.st));
    SONAR_CUDA(cudaStreamSynchronize(s.st));
  }
  if (cap_fail) return set_error(SONAR_ERR_INVALID, "path capacity too small");
  return SONAR_OK;
}

void run_dtw_device(sonar_ctx* ctx, DevCtx* dev, const double* const* q, const double* const* r,
                    const std::vector<int>* ids, DtwGeom g, int dim, int step, sonar_dtw_out* outs, DevJob* res) {
  set_current_ctx(ctx);
  auto fail = [&](int rc) {
    res->rc = rc;
    res->err = sonar_last_error();
  };
  cudaError_t e = cudaSetDevice(dev->device);
  if (e != cudaSuccess) return fail(cuda_error(e, "cudaSetDevice"));
  size_t free_b = 0, total_b = 0;
  if ((e = cudaMemGetInfo(&free_b, &total_b)) != cudaSuccess) return fail(cuda_error(e, "cudaMemGetInfo"));
  const size_t budget = (free_b + dev->slot[0].d_tmp.bytes) / 2;
  const size_t per_pair = sizeof(double) * ((size_t)g.cells + (size_t)g.n_off + 2) * 9 / 8 + 4096;
  if (per_pair > budget) {
    fail(set_error(SONAR_ERR_NOMEM, "DTW cost store does not fit device memory; use a Sakoe-Chiba band"));
    return;
  }
  const size_t chunk = std::max<size_t>(1, budget / per_pair);
  std::vector<const double*> qq, rr;
  std::vector<sonar_dtw_out> oo;
  for (size_t i = 0; i < ids->size(); i += chunk) {
    const size_t k = std::min(ids->size(), i + chunk);
    qq.clear(), rr.clear(), oo.clear();
    for (size_t x = i; x < k; x++) {
      qq.push_back(q[(*ids)[x]]);
      rr.push_back(r[(*ids)[x]]);
      oo.push_back(outs[(*ids)[x]]);
    }
    const int rc = run_dtw_chunk(*dev, dev->slot[0], qq.data(), rr.data(), (int)(k - i), g, dim, step, oo.data());
    for (size_t x = i; x < k; x++) outs[(*ids)[x]] = oo[x - i];
    if (rc) return fail(rc);
  }
}

// ---- comparison --------------------------------------------------------------

// Statistics of a Compare / BatchCompare call are gathered first and computed together: the comparison logic runs twice
// over the same code -- a COLLECT pass that only registers the (array, frames, dim) triples it would ask for, then ONE
// upload of the distinct arrays, ONE colstats_batch_kernel launch, ONE copy back -- and a LOOKUP pass that reads the
// table.  The query's arrays are uploaded and reduced once however many candidates follow (round 1 / 2: fourteen
// blocking upload-launch-download round trips per candidate, the query's statistics recomputed for each).
struct StatsBatch {
  enum Mode { COLLECT, LOOKUP } mode = COLLECT;
  struct Key {
    const double* x;
    int64_t t;
    int dim;
    bool operator<(const Key& o) const {
      return x != o.x ? x < o.x : (t != o.t ? t < o.t : dim < o.dim);
    }
  };
  std::map<Key, size_t> index;  // -> offset of the 2 * dim statistics in `stats`
  std::vector<Key> order;
  std::vector<double> stats;
  size_t n_stats = 0;
};
static thread_local StatsBatch* g_stats_batch = nullptr;

static int stats_batch_run(sonar_ctx* ctx, StatsBatch& b) {
  b.stats.assign(b.n_stats, 0.0);
  if (b.order.empty()) return SONAR_OK;
  DevCtx& dev = ctx->devs[0];
  Slot& s = dev.slot[0];
  size_t n_in = 0, n_cols = 0;
  for (const auto& k : b.order) {
    n_in += (size_t)k.t * k.dim;
    n_cols += (size_t)k.dim;
  }
  int rc;
  if ((rc = dev.ensure_dev(s.d_in, sizeof(double) * n_in)) || (rc = dev.ensure_dev(s.d_out, sizeof(double) * b.n_stats)) ||
      (rc = dev.ensure_dev(s.d_tmp, sizeof(ColJob) * n_cols)))
    return rc;
  std::vector<ColJob> jobs;
  jobs.reserve(n_cols);
  size_t off = 0;
  for (const auto& k : b.order) {
    const size_t so = b.index[k];
    SONAR_CUDA(cudaMemcpyAsync(static_cast<double*>(s.d_in.p) + off, k.x, sizeof(double) * (size_t)k.t * k.dim,
                               cudaMemcpyHostToDevice, s.st));
    for (int c = 0; c < k.dim; c++)
      jobs.push_back(ColJob{(int64_t)(off + c), k.t, k.dim, (int64_t)(so + c), (int64_t)(so + k.dim + c)});
    off += (size_t)k.t * k.dim;
  }
  SONAR_CUDA(cudaMemcpyAsync(s.d_tmp.p, jobs.data(), sizeof(ColJob) * jobs.size(), cudaMemcpyHostToDevice, s.st));
  if ((rc = launch_colstats_batch(static_cast<const double*>(s.d_in.p), static_cast<const ColJob*>(s.d_tmp.p),
                                  (int)jobs.size(), static_cast<double*>(s.d_out.p), s.st)))
    return rc;
  SONAR_CUDA(cudaMemcpyAsync(b.stats.data(), s.d_out.p, sizeof(double) * b.n_stats, cudaMemcpyDeviceToHost, s.st));
  SONAR_CUDA(cudaStreamSynchronize(s.st));  // also keeps `jobs` alive until its copy has been read
  return SONAR_OK;
}

int gpu_colstats(sonar_ctx* ctx, const double* x, int64_t t, int dim, double* st) {
  if (StatsBatch* b = g_stats_batch) {
    const StatsBatch::Key k{x, t, dim};
    if (b->mode == StatsBatch::COLLECT) {
      if (!b->index.count(k)) {
        b->index[k] = b->n_stats;
        b->order.push_back(k);
        b->n_stats += 2 * (size_t)dim;
      }
      for (int i = 0; i < 2 * dim; i++) st[i] = 0.0;
    } else {
      const size_t so = b->index.at(k);
      for (int i = 0; i < 2 * dim; i++) st[i] = b->stats[so + i];
    }
    return SONAR_OK;
  }
  DevCtx& dev = ctx->devs[0];
  Slot& s = dev.slot[0];
  const size_t nd = (size_t)t * dim;
  int rc;
  if ((rc = dev.ensure_dev(s.d_in, sizeof(double) * nd)) || (rc = dev.ensure_dev(s.d_out, sizeof(double) * 2 * dim)))
    return rc;
  SONAR_CUDA(cudaMemcpyAsync(s.d_in.p, x, sizeof(double) * nd, cudaMemcpyHostToDevice, s.st));
  if ((rc = launch_colstats(static_cast<const double*>(s.d_in.p), t, dim, static_cast<double*>(s.d_out.p), s.st)))
    return rc;
  SONAR_CUDA(cudaMemcpyAsync(st, s.d_out.p, sizeof(double) * 2 * dim, cudaMemcpyDeviceToHost, s.st));
  SONAR_CUDA(cudaStreamSynchronize(s.st));
  return SONAR_OK;
}

double cosine(const double* a, const double* b, int n) {  // comparison.go:858-873
  if (n == 0) return 0.0;
  double dot = 0, n1 = 0, n2 = 0;
  for (int i = 0; i < n; i++) {
    dot += a[i] * b[i];
    n1 += a[i] * a[i];
    n2 += b[i] * b[i];
  }
  n1 = std::sqrt(n1);
  n2 = std::sqrt(n2);
  if (n1 == 0 || n2 == 0) return 0.0;
  return dot / (n1 * n2);
}

int seq_stats_sim(sonar_ctx* ctx, const double* a, int64_t na, const double* b, int64_t nb, double* sim) {
  *sim = 0.0;  // compareSequenceStats :827-842
  if (na == 0 || nb == 0 || !a || !b) return SONAR_OK;
  double f1[2], f2[2];
  int rc;
  if ((rc = gpu_colstats(ctx, a, na, 1, f1)) || (rc = gpu_colstats(ctx, b, nb, 1, f2))) return rc;
  *sim = cosine(f1, f2, 2);
  return SONAR_OK;
}

double scalar_sim(double a, double b) {  // compareScalarFeatures :844-856
  if (a == 0 && b == 0) return 1.0;
  const double mx = std::fmax(std::fabs(a), std::fabs(b));
  if (mx == 0) return 1.0;
  return std::fmax(0.0, 1.0 - std::fabs(a - b) / mx);
}

double mean_of(const std::vector<double>& s) {
  double t = 0.0;
  for (double v : s) t += v;
  return t / (double)s.size();
}

}  // namespace
}  // namespace sonar

using namespace sonar;

struct sonar_xcorr_shard {
  sonar_ctx* ctx = nullptr;
  int device = 0;
  double* d_buf = nullptr;   // za | zb | corr | descriptors
  XcorrPair pair{};          // host copy (device pointers inside)
  XcorrPair* d_pair = nullptr;
  XcorrPairOut* d_out = nullptr;
  cudaStream_t st = nullptr;
};

extern "C" {

int sonar_truncate_to_alignment(int64_t n1, int64_t n2, int sample_rate, double offset_seconds, int64_t* start1,
                                int64_t* start2, int64_t* length) {
  if (!start1 || !start2 || !length) return set_error(SONAR_ERR_INVALID, "nil argument");
  const double srf = (double)sample_rate;
  /* int(math.Round(math.Abs(offsetSeconds) * sampleRateFloat)): math.Round rounds half away from zero = std::round */
  const int64_t off = (int64_t)std::round(std::fabs(offset_seconds) * srf);
  int64_t s1 = 0, s2 = 0, common = 0;
  if (offset_seconds > 0) { /* :241-253 */
    s2 = off;
    if (s2 >= n2)
      return set_error(SONAR_ERR_INVALID, "offset too large: need to skip " + std::to_string(s2) + " samples but pcm2 only has " +
                                         std::to_string(n2));
    common = std::min(n1 - s1, n2 - s2);
  } else if (offset_seconds < 0) { /* :255-267 */
    s1 = off;
    if (s1 >= n1)
      return set_error(SONAR_ERR_INVALID, "offset too large: need to skip " + std::to_string(s1) + " samples but pcm1 only has " +
                                         std::to_string(n1));
    common = std::min(n1 - s1, n2 - s2);
  } else {
    common = std::min(n1, n2); /* :269-272 */
  }
  if (common <= 0) return set_error(SONAR_ERR_INVALID, "no overlapping audio after alignment"); /* :275-277 */
  const int64_t pad = (int64_t)(0.5 * srf); /* :281-286 */
  if (common > 2 * pad) {
    s1 += pad;
    s2 += pad;
    common -= 2 * pad;
  }
  *start1 = s1;
  *start2 = s2;
  *length = common;
  return SONAR_OK;
}

int sonar_xcorr_ncc_f64(sonar_ctx* ctx, const double* a, int64_t na, const double* b, int64_t nb, int max_lag,
                        double* corr, sonar_xcorr_summary* out) {
  if (!ctx) return set_error(SONAR_ERR_INVALID, "nil argument");
  if (!a || !b || na <= 0 || nb <= 0) return set_error(SONAR_ERR_EMPTY, "empty signals provided");
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  DevCtx& dev = ctx->devs[0];
  SONAR_CUDA(cudaSetDevice(dev.device));
  PairJob j{a, na, b, nb, corr, out};
  return run_xcorr_chunk(dev, dev.slot[0], &j, 1, max_lag, true, nullptr);
}

int sonar_xcorr_batch_f64(sonar_ctx* ctx, const double* const* a, const int64_t* na, const double* const* b,
                          const int64_t* nb, int n_pairs, int max_lag, double* const* corr,
                          sonar_xcorr_summary* outs) {
  if (!ctx || (n_pairs > 0 && (!a || !na || !b || !nb || !outs))) return set_error(SONAR_ERR_INVALID, "nil argument");
  if (n_pairs <= 0) return SONAR_OK;
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  std::vector<PairJob> jobs(n_pairs);
  for (int p = 0; p < n_pairs; p++) {
    if (!a[p] || !b[p] || na[p] <= 0 || nb[p] <= 0) return set_error(SONAR_ERR_EMPTY, "empty signals provided");
    jobs[p] = PairJob{a[p], na[p], b[p], nb[p], corr ? corr[p] : nullptr, &outs[p]};
  }
  return xcorr_pairs(ctx, jobs, max_lag);
}

int sonar_xcorr_batch_dev(sonar_ctx* ctx, const double* a_dev, int64_t na, const double* b_dev, int64_t nb,
                          int n_pairs, int max_lag, double* corr_dev, sonar_xcorr_summary* summ_host) {
  if (!ctx || !a_dev || !b_dev) return set_error(SONAR_ERR_INVALID, "nil argument");
  if (na <= 0 || nb <= 0) return set_error(SONAR_ERR_EMPTY, "empty signals provided");
  if (n_pairs <= 0) return SONAR_OK;
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  DevCtx& dev = ctx->devs[0];
  SONAR_CUDA(cudaSetDevice(dev.device));
  std::vector<PairJob> jobs(n_pairs);
  for (int p = 0; p < n_pairs; p++)
    jobs[p] = PairJob{a_dev + (int64_t)p * na, na, b_dev + (int64_t)p * nb, nb, nullptr,
                      summ_host ? &summ_host[p] : nullptr};
  return run_xcorr_chunk(dev, dev.slot[0], jobs.data(), n_pairs, max_lag, false, corr_dev);
}

int sonar_xcorr_shard_open(sonar_ctx* ctx, const double* a, int64_t na, const double* b, int64_t nb, int max_lag,
                           int64_t idx_lo, int64_t idx_hi, sonar_xcorr_shard** out, sonar_xcorr_shard_peak* peak) {
  if (!ctx || !out || !peak) return set_error(SONAR_ERR_INVALID, "nil argument");
  if (!a || !b || na <= 0 || nb <= 0) return set_error(SONAR_ERR_EMPTY, "empty signals provided");
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  DevCtx& dev = ctx->devs[0];
  SONAR_CUDA(cudaSetDevice(dev.device));
  const int aml = actual_max_lag(max_lag, na, nb);
  const int64_t nl = 2 * (int64_t)aml + 1;
  int64_t lo = std::max<int64_t>(0, idx_lo), hi = std::min<int64_t>(nl, idx_hi);
  if (hi < lo) hi = lo;
  auto* sh = new sonar_xcorr_shard();
  sh->ctx = ctx;
  sh->device = dev.device;
  sh->st = dev.slot[0].st;
  const size_t doubles = (size_t)(4 * (even(na) + even(nb)) / 2 + even(hi - lo) + 2);
  const size_t desc = sizeof(XcorrSeq) * 2 + sizeof(XcorrPair) + sizeof(XcorrPairOut);
  cudaError_t e = cudaMalloc(&sh->d_buf, sizeof(double) * doubles + desc + 64);
  if (e != cudaSuccess) {
    delete sh;
    return cuda_error(e, "cudaMalloc(shard)");
  }
  double* d_a = sh->d_buf;
  double* d_b = d_a + even(na);
  double* d_za = d_b + even(nb);
  double* d_zb = d_za + even(na);
  double* d_c = d_zb + even(nb);
  unsigned char* dd = reinterpret_cast<unsigned char*>(d_c + even(hi - lo) + 2);
  XcorrSeq* d_seqs = reinterpret_cast<XcorrSeq*>(dd);
  sh->d_pair = reinterpret_cast<XcorrPair*>(dd + sizeof(XcorrSeq) * 2);
  sh->d_out = reinterpret_cast<XcorrPairOut*>(dd + sizeof(XcorrSeq) * 2 + sizeof(XcorrPair));
  XcorrSeq seqs[2] = {{d_a, d_za, na}, {d_b, d_zb, nb}};
  sh->pair = XcorrPair{d_za, d_zb, d_c, na, nb, lo, hi, aml, 0};
  XcorrPairOut o;
  int rc = SONAR_OK;
  auto run = [&]() -> int {
    SONAR_CUDA(cudaMemcpyAsync(d_a, a, sizeof(double) * (size_t)na, cudaMemcpyHostToDevice, sh->st));
    SONAR_CUDA(cudaMemcpyAsync(d_b, b, sizeof(double) * (size_t)nb, cudaMemcpyHostToDevice, sh->st));
    SONAR_CUDA(cudaMemcpyAsync(d_seqs, seqs, sizeof(seqs), cudaMemcpyHostToDevice, sh->st));
    SONAR_CUDA(cudaMemcpyAsync(sh->d_pair, &sh->pair, sizeof(XcorrPair), cudaMemcpyHostToDevice, sh->st));
    int r2;
    if ((r2 = launch_znorm(d_seqs, 2, sh->st))) return r2;
    if ((r2 = launch_xcorr(sh->d_pair, 1, hi - lo, sh->st))) return r2;
    if ((r2 = launch_xcorr_finalize(sh->d_pair, 1, -1, sh->d_out, sh->st))) return r2;
    SONAR_CUDA(cudaMemcpyAsync(&o, sh->d_out, sizeof(o), cudaMemcpyDeviceToHost, sh->st));
    SONAR_CUDA(cudaStreamSynchronize(sh->st));
    return SONAR_OK;
  };
  rc = run();
  if (rc) {
    cudaFree(sh->d_buf);
    delete sh;
    return rc;
  }
  peak->index = o.peak_index;
  peak->abs_peak = o.peak_index >= 0 ? std::fabs(o.peak) : 0.0;
  *out = sh;
  return SONAR_OK;
}

int sonar_xcorr_shard_metrics_f64(sonar_xcorr_shard* sh, int64_t gp, sonar_xcorr_shard_metrics* m) {
  if (!sh || !m) return set_error(SONAR_ERR_INVALID, "nil argument");
  std::lock_guard<std::mutex> call_lock(sh->ctx->call_mu);
  set_current_ctx(sh->ctx);
  SONAR_CUDA(cudaSetDevice(sh->device));
  int rc = launch_xcorr_finalize(sh->d_pair, 1, gp, sh->d_out, sh->st);
  if (rc) return rc;
  XcorrPairOut o;
  SONAR_CUDA(cudaMemcpyAsync(&o, sh->d_out, sizeof(o), cudaMemcpyDeviceToHost, sh->st));
  SONAR_CUDA(cudaStreamSynchronize(sh->st));
  m->noise_sum = o.noise_sum;
  m->noise_count = o.noise_cnt;
  m->max_sidelobe = o.max_sidelobe;
  m->second_abs = o.second_abs;
  m->second_val = o.second_val;
  m->second_index = (double)o.second_index;
  m->c_peak = o.c_peak;
  m->c_prev = o.c_prev;
  m->c_next = o.c_next;
  return SONAR_OK;
}

int sonar_xcorr_shard_corr(sonar_xcorr_shard* sh, double* corr) {
  if (!sh || !corr) return set_error(SONAR_ERR_INVALID, "nil argument");
  std::lock_guard<std::mutex> call_lock(sh->ctx->call_mu);
  SONAR_CUDA(cudaSetDevice(sh->device));
  const int64_t cnt = sh->pair.idx_hi - sh->pair.idx_lo;
  if (cnt > 0) {
    SONAR_CUDA(cudaMemcpyAsync(corr, sh->pair.corr, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost, sh->st));
    SONAR_CUDA(cudaStreamSynchronize(sh->st));
  }
  return SONAR_OK;
}

void sonar_xcorr_shard_close(sonar_xcorr_shard* sh) {
  if (!sh) return;
  {
    std::lock_guard<std::mutex> call_lock(sh->ctx->call_mu);
    cudaSetDevice(sh->device);
    cudaStreamSynchronize(sh->st);
    cudaFree(sh->d_buf);
  }
  delete sh;
}

int sonar_xcorr_merge_peaks(const sonar_xcorr_shard_peak* pk, int n, int64_t* gi) {
  if (!pk || !gi) return set_error(SONAR_ERR_INVALID, "nil argument");
  // findPeak scans ascending with strict '>' on |c| (correlation.go:535-541):
  // larger |c| wins, ties -> smaller global index.
  int64_t best = -1;
  double bv = 0.0;
  for (int i = 0; i < n; i++) {
    if (pk[i].index < 0) continue;
    if (best < 0 || pk[i].abs_peak > bv || (pk[i].abs_peak == bv && pk[i].index < best)) {
      best = pk[i].index;
      bv = pk[i].abs_peak;
    }
  }
  if (best < 0) return set_error(SONAR_ERR_EMPTY, "empty signals provided");
  *gi = best;
  return SONAR_OK;
}

int sonar_xcorr_merge_metrics(const sonar_xcorr_shard_metrics* parts, int n, int64_t na, int64_t nb, int max_lag,
                              int64_t gp, sonar_xcorr_summary* o) {
  if (!parts || !o) return set_error(SONAR_ERR_INVALID, "nil argument");
  const int aml = actual_max_lag(max_lag, na, nb);
  double c0 = kNaN, cm = kNaN, cp = kNaN, ns = 0, nc = 0, ms = 0, sa = 0, sv = 0, si = -1;
  for (int i = 0; i < n; i++) {
    const auto& m = parts[i];
    if (!std::isnan(m.c_peak)) c0 = m.c_peak;
    if (!std::isnan(m.c_prev)) cm = m.c_prev;
    if (!std::isnan(m.c_next)) cp = m.c_next;
    ns += m.noise_sum;
    nc += m.noise_count;
    if (m.max_sidelobe > ms) ms = m.max_sidelobe;
    if (m.second_index >= 0 && (m.second_abs > sa || (m.second_abs == sa && (si < 0 || m.second_index < si)))) {
      sa = m.second_abs;
      sv = m.second_val;
      si = m.second_index;
    }
  }
  summarize(c0, cm, cp, ns, nc, ms, sv, gp, aml, na, nb, 0, o);
  return SONAR_OK;
}

int sonar_align_xcorr_f64(sonar_ctx* ctx, const double* q, int64_t nq, const double* r, int64_t nr,
                          int max_lag_frames, int hop, int sr, double* corr, sonar_xcorr_summary* xc,
                          sonar_align_result* out) {
  if (!ctx || !out) return set_error(SONAR_ERR_INVALID, "nil argument");
  if (!q || !r || nq <= 0 || nr <= 0)
    return set_error(SONAR_ERR_EMPTY, "empty feature sequences provided");  // stats/alignment.go:86
  const int64_t min_frames = std::min(nq, nr);  // extractors/alignment.go:372-374
  const int ml = (int)std::min<int64_t>(max_lag_frames, min_frames - 1);
  sonar_xcorr_summary local;
  if (!xc) xc = &local;
  int rc = sonar_xcorr_ncc_f64(ctx, q, nq, r, nr, ml, corr, xc);
  if (rc) return rc;
  fill_align_from_xcorr(xc, nq, nr, ml, hop, sr, out);
  return SONAR_OK;
}

int sonar_dtw_batch_f64(sonar_ctx* ctx, const double* const* q, const double* const* r, int n_pairs, int n, int m,
                        int dim, int band, int step_pattern, int metric, sonar_dtw_out* outs) {
  if (!ctx || (n_pairs > 0 && (!q || !r || !outs))) return set_error(SONAR_ERR_INVALID, "nil argument");
  if (n_pairs <= 0) return SONAR_OK;
  for (int p = 0; p < n_pairs; p++) {
    int rc = dtw_validate(q[p], n, r[p], m, dim, step_pattern, metric);
    if (rc) return rc;
  }
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  DtwGeom g;
  dtw_geometry(n, m, band, &g);
  const int nd = (int)ctx->devs.size();
  std::vector<std::vector<int>> ids(nd);
  for (int p = 0; p < n_pairs; p++) ids[p % nd].push_back(p);
  std::vector<DevJob> res(nd);
  if (nd == 1) {
    run_dtw_device(ctx, &ctx->devs[0], q, r, &ids[0], g, dim, step_pattern, outs, &res[0]);
  } else {
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++)
      th.emplace_back(run_dtw_device, ctx, &ctx->devs[d], q, r, &ids[d], g, dim, step_pattern, outs, &res[d]);
    for (auto& t : th) t.join();
    cudaSetDevice(ctx->devs[0].device);
  }
  for (auto& x : res)
    if (x.rc) return set_error(x.rc, x.err);
  return SONAR_OK;
}

int sonar_dtw_f64(sonar_ctx* ctx, const double* q, int n, const double* r, int m, int dim, int band,
                  int step_pattern, int metric, sonar_dtw_out* out) {
  if (!ctx || !out) return set_error(SONAR_ERR_INVALID, "nil argument");
  const double* qq[1] = {q};
  const double* rr[1] = {r};
  return sonar_dtw_batch_f64(ctx, qq, rr, 1, n, m, dim, band, step_pattern, metric, out);
}

int sonar_colstats_f64(sonar_ctx* ctx, const double* x, int64_t t, int dim, double* st) {
  if (!ctx || !st) return set_error(SONAR_ERR_INVALID, "nil argument");
  if (!x || t <= 0 || dim <= 0) return set_error(SONAR_ERR_EMPTY, "empty feature sequences provided");
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  SONAR_CUDA(cudaSetDevice(ctx->devs[0].device));
  return gpu_colstats(ctx, x, t, dim, st);
}

int sonar_colstats_cosine_f64(sonar_ctx* ctx, const double* x, int64_t tx, const double* y, int64_t ty, int dim,
                              double* sim) {
  if (!ctx || !sim) return set_error(SONAR_ERR_INVALID, "nil argument");
  *sim = 0.0;
  if (!x || !y || tx <= 0 || ty <= 0 || dim <= 0) return SONAR_OK;
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  SONAR_CUDA(cudaSetDevice(ctx->devs[0].device));
  std::vector<double> s1(2 * dim), s2(2 * dim);
  int rc;
  if ((rc = gpu_colstats(ctx, x, tx, dim, s1.data())) || (rc = gpu_colstats(ctx, y, ty, dim, s2.data()))) return rc;
  *sim = cosine(s1.data(), s2.data(), 2 * dim);
  return SONAR_OK;
}

static int compare_locked(sonar_ctx* ctx, const sonar_cmp_features* f1, const sonar_cmp_features* f2,
                          const sonar_cmp_weights* w, int content_filter, sonar_cmp_result* o);
static int compare_many(sonar_ctx* ctx, const sonar_cmp_features* query, const sonar_cmp_features* const* cands, int n,
                        const sonar_cmp_weights* w, int content_filter, sonar_cmp_result* results);

int sonar_compare_f64(sonar_ctx* ctx, const sonar_cmp_features* f1, const sonar_cmp_features* f2,
                      const sonar_cmp_weights* w, int content_filter, sonar_cmp_result* o) {
  if (!ctx || !f1 || !f2 || !w || !o) return set_error(SONAR_ERR_INVALID, "fingerprints cannot be nil");  // :135
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  SONAR_CUDA(cudaSetDevice(ctx->devs[0].device));
  const sonar_cmp_features* cands[1] = {f2};
  return compare_many(ctx, f1, cands, 1, w, content_filter, o);
}

int sonar_compare_batch_f64(sonar_ctx* ctx, const sonar_cmp_features* query, const sonar_cmp_features* const* cands,
                            int n, const sonar_cmp_weights* w, int content_filter, sonar_cmp_result* results) {
  if (!ctx || !query || !w || (n > 0 && (!cands || !results)))
    return set_error(SONAR_ERR_INVALID, "query fingerprint cannot be nil");  // comparison.go:1108-1110
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  SONAR_CUDA(cudaSetDevice(ctx->devs[0].device));
  return compare_many(ctx, query, cands, n, w, content_filter, results);
}

// COLLECT pass -> one batched statistics launch -> LOOKUP pass (see StatsBatch)
static int compare_many(sonar_ctx* ctx, const sonar_cmp_features* query, const sonar_cmp_features* const* cands, int n,
                        const sonar_cmp_weights* w, int content_filter, sonar_cmp_result* results) {
  StatsBatch batch;
  struct Guard {
    ~Guard() { g_stats_batch = nullptr; }
  } guard;
  g_stats_batch = &batch;
  for (int pass = 0; pass < 2; pass++) {
    batch.mode = pass == 0 ? StatsBatch::COLLECT : StatsBatch::LOOKUP;
    for (int i = 0; i < n; i++) {
      if (!cands[i]) {  // comparison.go:1123-1125: nil candidates are skipped
        std::memset(&results[i], 0, sizeof(results[i]));
        results[i].n_features = -1;
        continue;
      }
      int rc = compare_locked(ctx, query, cands[i], w, content_filter, &results[i]);
      if (rc) return rc;
    }
    if (pass == 0) {
      g_stats_batch = nullptr;  // the real launch
      int rc = stats_batch_run(ctx, batch);
      if (rc) return rc;
      g_stats_batch = &batch;
    }
  }
  return SONAR_OK;
}

static int compare_locked(sonar_ctx* ctx, const sonar_cmp_features* f1, const sonar_cmp_features* f2,
                          const sonar_cmp_weights* w, int content_filter, sonar_cmp_result* o) {
  std::memset(o, 0, sizeof(*o));
  o->dist_mfcc = o->dist_spectral = o->dist_temporal = o->dist_harmonic = kNaN;
  o->content_type_match = f1->content_type == f2->content_type;
  if (content_filter && !o->content_type_match) {  // comparison.go:160-166
    o->overall_similarity = 0.0;
    o->confidence = 0.25;
    return SONAR_OK;
  }
  int rc;
  std::vector<double> sims, wts;
  // order of w->w: mfcc, spectral, chroma, temporal, speech, harmonic, energy
  if (f1->mfcc_frames > 0 && f2->mfcc_frames > 0 && f1->mfcc && f2->mfcc) {  // :286-291,344-404
    double sim = 0.0;
    if (f1->mfcc_dim > 0 && f2->mfcc_dim > 0) {
      std::vector<double> s1(2 * f1->mfcc_dim), s2(2 * f2->mfcc_dim);
      if ((rc = gpu_colstats(ctx, f1->mfcc, f1->mfcc_frames, f1->mfcc_dim, s1.data())) ||
          (rc = gpu_colstats(ctx, f2->mfcc, f2->mfcc_frames, f2->mfcc_dim, s2.data())))
        return rc;
      if (f1->mfcc_dim == f2->mfcc_dim) sim = cosine(s1.data(), s2.data(), 2 * f1->mfcc_dim);
    }
    sims.push_back(sim);
    wts.push_back(w->w[0]);
    o->dist_mfcc = 1.0 - sim;
  }
  auto seq = [&](const double* a, int64_t na, const double* b, int64_t nb, std::vector<double>& s) -> int {
    if (na > 0 && nb > 0) {
      double v;
      int r2 = seq_stats_sim(ctx, a, na, b, nb, &v);
      if (r2) return r2;
      s.push_back(v);
    }
    return SONAR_OK;
  };
  if (f1->has_spectral && f2->has_spectral) {  // :294-298,646-671
    std::vector<double> s;
    if ((rc = seq(f1->spectral_centroid, f1->n_centroid, f2->spectral_centroid, f2->n_centroid, s)) ||
        (rc = seq(f1->spectral_rolloff, f1->n_rolloff, f2->spectral_rolloff, f2->n_rolloff, s)) ||
        (rc = seq(f1->spectral_flux, f1->n_flux, f2->spectral_flux, f2->n_flux, s)))
      return rc;
    double sim = 0.0, dist = 1.0;
    if (!s.empty()) {
      sim = mean_of(s);
      dist = 1.0 - sim;
    }
    sims.push_back(sim);
    wts.push_back(w->w[1]);
    o->dist_spectral = dist;
  }
  if (f1->has_temporal && f2->has_temporal) {  // :310-315,688-718
    std::vector<double> s;
    if (f1->dynamic_range > 0 && f2->dynamic_range > 0) s.push_back(scalar_sim(f1->dynamic_range, f2->dynamic_range));
    s.push_back(scalar_sim(f1->silence_ratio, f2->silence_ratio));
    if (f1->onset_density > 0 && f2->onset_density > 0) s.push_back(scalar_sim(f1->onset_density, f2->onset_density));
    if ((rc = seq(f1->rms_energy, f1->n_rms, f2->rms_energy, f2->n_rms, s))) return rc;
    const double sim = mean_of(s);
    sims.push_back(sim);
    wts.push_back(w->w[3]);
    o->dist_temporal = 1.0 - sim;
  }
  if (f1->has_harmonic && f2->has_harmonic) {  // :326-331,746-771
    std::vector<double> s;
    if ((rc = seq(f1->harmonic_ratio, f1->n_harmonic_ratio, f2->harmonic_ratio, f2->n_harmonic_ratio, s)) ||
        (rc = seq(f1->pitch_estimate, f1->n_pitch, f2->pitch_estimate, f2->n_pitch, s)))
      return rc;
    double sim = 0.0, dist = 1.0;
    if (!s.empty()) {
      sim = mean_of(s);
      dist = 1.0 - sim;
    }
    sims.push_back(sim);
    wts.push_back(w->w[5]);
    o->dist_harmonic = dist;
  }
  o->n_features = (int32_t)sims.size();
  double fs = 0.0;
  if (!sims.empty()) {  // stat.Mean(values, weights) = sum(w*x)/sum(w)  (:875-882)
    double sw = 0.0, swx = 0.0;
    for (size_t i = 0; i < sims.size(); i++) {
      swx += wts[i] * sims[i];
      sw += wts[i];
    }
    fs = swx / sw;
  }
  o->feature_similarity = fs;
  o->overall_similarity = fs;  // :886-889
  double conf = 0.5;           // :1011-1037
  if (o->overall_similarity > 0.8)
    conf += 0.3;
  else if (o->overall_similarity > 0.6)
    conf += 0.2;
  if (o->content_type_match) conf += 0.1;
  conf += (double)o->n_features * 0.05;
  o->confidence = std::fmax(0.0, std::fmin(1.0, conf));
  return SONAR_OK;
}

}  // extern "C"
