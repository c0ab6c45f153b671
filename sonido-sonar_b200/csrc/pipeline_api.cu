// Chained entry point for the whole CDN-latency measurement loop of a batch of source/CDN pairs:
//   GenerateFingerprint x2 -> ExtractAlignmentFeatures ("corr_energy") -> banded DTW of the lag-trimmed
//   energy series, i.e. BASELINE config[1], and the flow of AlignmentExtractor.AlignAudioFiles
//   (fingerprint/extractors/alignment.go:489-560: short-time energies of both PCM buffers, cross-correlation,
//   then DTW) with the full fingerprint features produced on the way (device-resident
//   chaining of the callers either side of the hot path).
//
// Everything between the H2D copy of a pair's PCM and the D2H copy of its results is enqueued on ONE stream
// without a host round trip: the z-score kernels read the short-time energies straight out of the feature
// blocks, a tiny kernel turns the detected lag into the trimmed start pointers the DTW kernels consume
// (TruncateToAlignmentPCM's sign convention, alignment.go:239-243).  Eight pairs are in flight per device: three
// PCM staging buffers feed the GPU-filling front half (fingerprint, NCC) on their own streams, and each pair's
// latency-bound tail (one DTW warp, then the D2H copy) runs on a second stream, so the PCIe copy and the
// kernels of the following pairs overlap it and the host-side scatter of finished pairs.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

#include "common.h"

namespace sonar {
namespace {

struct PairGeom {
  FpShape sh;
  int64_t n = 0, stride = 0;  // samples per stream, device stride
  int64_t Te = 0;             // energy frames per stream
  int aml = 0;                // clamped max lag in frames (extractors/alignment.go:372-374, correlation.go:452-461)
  int64_t nl = 0;             // 2*aml+1
  int dtw_len = 0;            // Te - aml: the overlap common to every admissible lag
  int band = 0;
  DtwGeom g{};
  int64_t path_cap = 0;
  // per-pair scratch layout in d_tmp (doubles unless noted)
  size_t fp_tmp = 0, z = 0, corr = 0, cells = 0, desc_bytes = 0, total_bytes = 0;
  // per-pair result layout in d_out / h_out (bytes)
  size_t feat_bytes = 0, res_off = 0, res_bytes = 0;
};

struct PairDesc {  // device-side descriptors of one pair, packed behind its scratch
  XcorrSeq seqs[2];
  XcorrPair pair;
  XcorrPairOut xo;
  DtwPairOut dout;
  const double* qptr;
  const double* rptr;
};

inline size_t up(size_t b) { return (b + 255) & ~(size_t)255; }

int pair_geometry(const sonar_fp_params* p, int64_t n, double max_lag_seconds, int band, PairGeom* G) {
  int rc = fp_shape(p, n, &G->sh);
  if (rc) return rc;
  G->n = n;
  G->stride = (n + 1) & ~(int64_t)1;
  G->Te = G->sh.sz.n_energy_frames;
  if (G->Te <= 0) return set_error(SONAR_ERR_EMPTY, "empty feature sequences provided");  // stats/alignment.go:86
  if (p->hop_size <= 0 || p->energy_hop <= 0) return set_error(SONAR_ERR_INVALID, "hop size must be positive");
  // NewAlignmentExtractorWithMaxLag: maxLagSamples = int(maxLagSeconds * SampleRate); frames = samples / HopSize
  const int64_t max_lag_samples = (int64_t)(max_lag_seconds * (double)p->call_sample_rate);
  int64_t ml = max_lag_samples / p->energy_hop;
  ml = std::min<int64_t>(ml, G->Te - 1);  // alignWithFeatures clamp
  G->aml = actual_max_lag((int)std::max<int64_t>(ml, 0), G->Te, G->Te);
  G->nl = 2 * (int64_t)G->aml + 1;
  G->dtw_len = (int)(G->Te - G->aml);
  G->band = band;
  dtw_geometry(G->dtw_len, G->dtw_len, band, &G->g);
  G->path_cap = 2 * (int64_t)G->dtw_len;
  if (sizeof(double) * (size_t)(G->g.n_off + 2) > 140 * 1024)
    return set_error(SONAR_ERR_UNSUPPORTED, "sonar_align_pairs needs a Sakoe-Chiba band (dtw_band > 0) for long streams");
  G->fp_tmp = 2 * G->sh.tmp_doubles_per_stream;
  G->z = 2 * (size_t)((G->Te + 1) & ~(int64_t)1);
  G->corr = (size_t)((G->nl + 1) & ~(int64_t)1);
  G->cells = (size_t)G->g.cells;
  G->desc_bytes = up(sizeof(PairDesc));
  G->total_bytes = up(sizeof(double) * (G->fp_tmp + G->z + G->corr + G->cells)) + G->desc_bytes;
  G->feat_bytes = up(sizeof(double) * 2 * (size_t)G->sh.L.total);
  // results: corr | path_c (double) | path_q, path_r (int32) | PairDesc copy (xo + dout)
  G->res_off = G->feat_bytes;
  G->res_bytes = up(sizeof(double) * G->corr) + up(sizeof(double) * (size_t)G->path_cap) +
                 up(sizeof(int32_t) * 2 * (size_t)G->path_cap) + up(sizeof(PairDesc));
  return SONAR_OK;
}

struct PairDevPtrs {
  double* feat;     // 2 feature blocks
  double* fp_tmp;
  double* z;
  double* corr;
  double* cells;
  PairDesc* desc;
  double* path_c;
  int32_t* path_q;
  int32_t* path_r;
};

// Front half of ONE pair on `st`: fingerprint -> NCC -> trim (these kernels fill the GPU).  The PCM (query at
// pcm_dev, reference at pcm_dev + stride) is already on the device or queued on the same stream.
int enqueue_pair_front(sonar_ctx* ctx, int device, const sonar_fp_params* p, const PairGeom& G, const double* pcm_dev,
                       const PairDevPtrs& d, cudaStream_t st) {
  int rc = enqueue_fingerprint(ctx, device, p, G.sh, pcm_dev, G.n, G.stride, 2, d.feat, d.fp_tmp, st);
  if (rc) return rc;
  PairDesc h;
  std::memset(&h, 0, sizeof(h));
  const double* ea = d.feat + G.sh.L.short_time_energy;
  const double* eb = d.feat + G.sh.L.total + G.sh.L.short_time_energy;
  double* za = d.z;
  double* zb = d.z + G.z / 2;
  h.seqs[0] = XcorrSeq{ea, za, G.Te};
  h.seqs[1] = XcorrSeq{eb, zb, G.Te};
  h.pair = XcorrPair{za, zb, d.corr, G.Te, G.Te, 0, G.nl, G.aml, 0};
  SONAR_CUDA(cudaMemcpyAsync(d.desc, &h, sizeof(h), cudaMemcpyHostToDevice, st));  // pageable source: staged before return
  if ((rc = launch_znorm(d.desc->seqs, 2, st))) return rc;
  if ((rc = launch_xcorr(&d.desc->pair, 1, G.nl, st))) return rc;
  if ((rc = launch_xcorr_finalize(&d.desc->pair, 1, -1, &d.desc->xo, st))) return rc;
  return launch_xcorr_trim(d.desc->seqs, &d.desc->pair, &d.desc->xo, 1, &d.desc->qptr, &d.desc->rptr, st);
}

// Tail of the pair on `st`: the banded DTW is a single latency-bound warp, so it runs on the lane's second
// stream where it overlaps the front halves of the following pairs.
int enqueue_pair_tail(const PairGeom& G, const PairDevPtrs& d, cudaStream_t st) {
  return launch_dtw(nullptr, nullptr, 1, G.g, 1, SONAR_STEP_SYMMETRIC2, d.cells, nullptr, d.path_q, d.path_r, d.path_c,
                    G.path_cap, &d.desc->dout, st, &d.desc->qptr, &d.desc->rptr);
}

PairDevPtrs carve(const PairGeom& G, unsigned char* tmp, unsigned char* out) {
  PairDevPtrs d;
  double* t = reinterpret_cast<double*>(tmp);
  d.fp_tmp = t;
  d.z = t + G.fp_tmp;
  d.corr = nullptr;  // lives in the result block
  d.cells = t + G.fp_tmp + G.z;
  d.desc = reinterpret_cast<PairDesc*>(tmp + up(sizeof(double) * (G.fp_tmp + G.z + G.corr + G.cells)));
  d.feat = reinterpret_cast<double*>(out);
  unsigned char* r = out + G.res_off;
  d.corr = reinterpret_cast<double*>(r);
  r += up(sizeof(double) * G.corr);
  d.path_c = reinterpret_cast<double*>(r);
  r += up(sizeof(double) * (size_t)G.path_cap);
  d.path_q = reinterpret_cast<int32_t*>(r);
  d.path_r = d.path_q + G.path_cap;
  return d;
}

// host side of one finished pair: h = pinned copy of the pair's result block (features | results | descriptor)
int finish_pair(const sonar_fp_params* p, const PairGeom& G, const unsigned char* h, bool have_features,
                sonar_pair_out* o) {
  if (have_features) {
    const double* f = reinterpret_cast<const double*>(h);
    scatter_block(f, G.sh, &o->query);
    scatter_block(f + G.sh.L.total, G.sh, &o->reference);
  }
  const unsigned char* r = h + G.res_off;
  const double* corr = reinterpret_cast<const double*>(r);
  r += up(sizeof(double) * G.corr);
  const double* pc = reinterpret_cast<const double*>(r);
  r += up(sizeof(double) * (size_t)G.path_cap);
  const int32_t* pq = reinterpret_cast<const int32_t*>(r);
  const int32_t* pr = pq + G.path_cap;
  r += up(sizeof(int32_t) * 2 * (size_t)G.path_cap);
  const PairDesc* d = reinterpret_cast<const PairDesc*>(r);
  summarize_xcorr(d->xo, G.aml, G.Te, G.Te, G.nl, &o->xcorr);
  fill_align_from_xcorr(&o->xcorr, G.Te, G.Te, G.aml, p->energy_hop, p->call_sample_rate, &o->corr_alignment);
  if (o->corr) std::memcpy(o->corr, corr, sizeof(double) * (size_t)G.nl);
  o->dtw_length = G.dtw_len;
  sonar_dtw_out& w = o->dtw;
  const int64_t len = d->dout.path_len;
  w.path_len = len;
  w.total_cost = d->dout.total_cost;
  w.distance = d->dout.total_cost / (double)len;
  const int64_t take = std::min<int64_t>(len, w.path_cap);
  const int64_t off = G.path_cap - len;
  if (take > 0) {
    if (w.path_query) std::memcpy(w.path_query, pq + off, sizeof(int32_t) * (size_t)take);
    if (w.path_ref) std::memcpy(w.path_ref, pr + off, sizeof(int32_t) * (size_t)take);
    if (w.path_cost) std::memcpy(w.path_cost, pc + off, sizeof(double) * (size_t)take);
  }
  if (len > w.path_cap && (w.path_query || w.path_ref || w.path_cost))
    return set_error(SONAR_ERR_INVALID, "path capacity too small");
  return SONAR_OK;
}

struct DevJob {
  int rc = SONAR_OK;
  std::string err;
};

// pcm_q / pcm_r: host pointers (host_pcm) or device pointers on this device
void run_pairs_device(sonar_ctx* ctx, DevCtx* dev, const double* const* pcm_q, const double* const* pcm_r,
                      const std::vector<int>* ids, const sonar_fp_params* p, const PairGeom* Gp, bool host_pcm,
                      sonar_pair_out* outs, DevJob* job) {
  const PairGeom& G = *Gp;
  set_current_ctx(ctx);
  auto fail = [&](int rc) {
    job->rc = rc;
    job->err = sonar_last_error();
  };
  cudaError_t e = cudaSetDevice(dev->device);
  if (e != cudaSuccess) return fail(cuda_error(e, "cudaSetDevice"));
  constexpr int NL = DevCtx::kSlots;       // lanes: result + scratch buffers, second stream, one pair in flight each
  constexpr int NP = DevCtx::kStageSlots;  // PCM staging buffers + front-half streams
  int pending[NL];
  for (int& x : pending) x = -1;
  const size_t out_bytes = G.feat_bytes + G.res_bytes;
  auto finish = [&](int li) -> int {
    if (pending[li] < 0) return SONAR_OK;
    Slot& s = dev->slot[li];
    SONAR_CUDA(cudaEventSynchronize(s.done));
    const int id = pending[li];
    pending[li] = -1;
    return finish_pair(p, G, static_cast<const unsigned char*>(s.h_out.p), true, &outs[id]);
  };
  int k = 0;
  for (int id : *ids) {
    const int li = k % NL, pi = k % NP;
    ++k;
    Slot& lane = dev->slot[li];
    Slot& stage = dev->slot[pi];
    int rc = finish(li);
    if (rc) return fail(rc);
    if ((host_pcm && (rc = dev->ensure_dev(stage.d_in, sizeof(double) * 2 * (size_t)G.stride))) ||
        (rc = dev->ensure_dev(lane.d_tmp, G.total_bytes)) || (rc = dev->ensure_dev(lane.d_out, out_bytes)) ||
        (rc = dev->ensure_host(lane.h_out, out_bytes)))
      return fail(rc);
    const double* pcm_dev;
    if (host_pcm) {
      // the staging buffer's previous user (pair k - NP) ran its fingerprint kernels on this same stream: ordered
      double* d_in = static_cast<double*>(stage.d_in.p);
      if ((e = cudaMemcpyAsync(d_in, pcm_q[id], sizeof(double) * (size_t)G.n, cudaMemcpyHostToDevice, stage.st)) !=
              cudaSuccess ||
          (e = cudaMemcpyAsync(d_in + G.stride, pcm_r[id], sizeof(double) * (size_t)G.n, cudaMemcpyHostToDevice,
                               stage.st)) != cudaSuccess)
        return fail(cuda_error(e, "cudaMemcpyAsync(H2D pcm)"));
      pcm_dev = d_in;
    } else {
      pcm_dev = pcm_q[id];  // device-resident: query and reference adjacent, `stride` apart
    }
    const PairDevPtrs d = carve(G, static_cast<unsigned char*>(lane.d_tmp.p), static_cast<unsigned char*>(lane.d_out.p));
    rc = enqueue_pair_front(ctx, dev->device, p, G, pcm_dev, d, stage.st);
    if (rc) return fail(rc);
    if ((e = cudaEventRecord(lane.mid, stage.st)) != cudaSuccess ||
        (e = cudaStreamWaitEvent(lane.st2, lane.mid, 0)) != cudaSuccess)
      return fail(cuda_error(e, "cudaEventRecord/cudaStreamWaitEvent"));
    rc = enqueue_pair_tail(G, d, lane.st2);
    if (rc) return fail(rc);
    // results: features + corr + paths in one copy, then the descriptor block (xcorr partials, DTW totals)
    unsigned char* h = static_cast<unsigned char*>(lane.h_out.p);
    const size_t body = G.feat_bytes + G.res_bytes - up(sizeof(PairDesc));
    if ((e = cudaMemcpyAsync(h, lane.d_out.p, body, cudaMemcpyDeviceToHost, lane.st2)) != cudaSuccess ||
        (e = cudaMemcpyAsync(h + body, d.desc, sizeof(PairDesc), cudaMemcpyDeviceToHost, lane.st2)) != cudaSuccess)
      return fail(cuda_error(e, "cudaMemcpyAsync(D2H results)"));
    if ((e = cudaEventRecord(lane.done, lane.st2)) != cudaSuccess) return fail(cuda_error(e, "cudaEventRecord"));
    pending[li] = id;
  }
  for (int j = 0; j < NL; j++) {
    int rc = finish((k + j) % NL);
    if (rc) return fail(rc);
  }
}

int align_pairs(sonar_ctx* ctx, const double* const* q, const double* const* r, int64_t n, int n_pairs,
                const sonar_fp_params* p, double max_lag_seconds, int dtw_band, bool host_pcm, sonar_pair_out* outs) {
  if (!ctx || !p || (n_pairs > 0 && (!q || !outs))) return set_error(SONAR_ERR_INVALID, "audio data cannot be nil");
  if (n_pairs <= 0) return SONAR_OK;
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  int rc = fp_validate(p);
  if (rc) return rc;
  if (p->enable & SONAR_FP_ENABLE_TEMPORAL)
    return set_error(SONAR_ERR_UNSUPPORTED, "temporal features are not part of the pair pipeline");
  PairGeom G;
  rc = pair_geometry(p, n, max_lag_seconds, dtw_band, &G);
  if (rc) return rc;
  for (int i = 0; i < n_pairs; i++)
    if (!q[i] || (host_pcm && (!r || !r[i]))) return set_error(SONAR_ERR_INVALID, "audio data cannot be nil");
  const int nd = host_pcm ? (int)ctx->devs.size() : 1;
  std::vector<std::vector<int>> ids(nd);
  for (int i = 0; i < n_pairs; i++) ids[i % nd].push_back(i);
  std::vector<DevJob> res(nd);
  if (nd == 1) {
    run_pairs_device(ctx, &ctx->devs[0], q, r, &ids[0], p, &G, host_pcm, outs, &res[0]);
  } else {
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++)
      th.emplace_back(run_pairs_device, ctx, &ctx->devs[d], q, r, &ids[d], p, &G, host_pcm, outs, &res[d]);
    for (auto& t : th) t.join();
    cudaSetDevice(ctx->devs[0].device);
  }
  for (auto& x : res)
    if (x.rc) return set_error(x.rc, x.err);
  return SONAR_OK;
}

}  // namespace
}  // namespace sonar

using namespace sonar;

extern "C" {

int sonar_align_pairs_sizes(const sonar_fp_params* p, int64_t n, double max_lag_seconds, int32_t* n_lags,
                            int32_t* dtw_length) {
  if (!p) return set_error(SONAR_ERR_INVALID, "nil argument");
  PairGeom G;
  int rc = pair_geometry(p, n, max_lag_seconds, 1, &G);
  if (rc) return rc;
  if (n_lags) *n_lags = (int32_t)G.nl;
  if (dtw_length) *dtw_length = G.dtw_len;
  return SONAR_OK;
}

int sonar_align_pairs_f64(sonar_ctx* ctx, const double* const* query_pcm, const double* const* reference_pcm, int64_t n,
                          int n_pairs, const sonar_fp_params* p, double max_lag_seconds, int dtw_band,
                          sonar_pair_out* outs) {
  return align_pairs(ctx, query_pcm, reference_pcm, n, n_pairs, p, max_lag_seconds, dtw_band, true, outs);
}

int sonar_align_pairs_dev(sonar_ctx* ctx, const double* pcm_dev, int64_t n, int64_t stride, int n_pairs,
                          const sonar_fp_params* p, double max_lag_seconds, int dtw_band, sonar_pair_out* outs) {
  if (!pcm_dev) return set_error(SONAR_ERR_INVALID, "audio data cannot be nil");
  if (stride != ((n + 1) & ~(int64_t)1)) return set_error(SONAR_ERR_INVALID, "stride must be n rounded up to even");
  std::vector<const double*> q(n_pairs > 0 ? n_pairs : 0);
  for (int i = 0; i < n_pairs; i++) q[i] = pcm_dev + (int64_t)2 * i * stride;
  return align_pairs(ctx, q.data(), nullptr, n, n_pairs, p, max_lag_seconds, dtw_band, false, outs);
}

}  // extern "C"
