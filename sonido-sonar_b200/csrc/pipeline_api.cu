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
// (TruncateToAlignmentPCM's sign convention, alignment.go:239-243).  Pairs travel in chunks (as many as fit a
// 256 MB PCM staging buffer, at most four; 32 when the PCM is already resident), up to eight chunks in flight per
// device: three staging buffers feed the fingerprint kernels on their own streams, and each chunk's alignment
// branch (z-score, NCC, one DTW warp per pair, then the D2H copy) is forked onto a second stream as soon as the
// short-time energies exist, so its latency-bound kernels run beside the YIN kernels of the same chunk and the
// PCIe copies, kernels and host-side scatter of the neighbouring chunks.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
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
  int64_t path_tail = 0;  // entries of a path array that travel with the result copy: the path is emitted at the END of its
                          // capacity and is rarely longer than the sequences plus a few percent (rest fetched on demand)
  size_t z_pair = 0, corr_pair = 0;  // doubles per pair (even)
};

inline size_t up(size_t b) { return (b + 255) & ~(size_t)255; }

// f32 / s16 samples -> the float64 the reference's decoder would have produced (exact in both cases)
template <class T>
__global__ void widen_pcm_kernel(const T* __restrict__ src, double* __restrict__ dst, int64_t n, int64_t src_stride,
                                 int64_t dst_stride) {
  const T* s = src + (int64_t)blockIdx.y * src_stride;
  double* d = dst + (int64_t)blockIdx.y * dst_stride;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (sizeof(T) == 2)
      d[i] = (double)s[i] * (1.0 / 32768.0);  // swresample's s16 -> dbl
    else
      d[i] = (double)s[i];
  }
}

}  // namespace

int launch_widen_pcm(const void* src, int fmt, double* dst, int64_t n, int64_t src_stride, int64_t dst_stride, int rows,
                     cudaStream_t st) {
  if (rows <= 0 || n <= 0) return SONAR_OK;
  const dim3 wg(296, (unsigned)rows);
  if (fmt == SONAR_PCM_S16)
    widen_pcm_kernel<int16_t><<<wg, 256, 0, st>>>(static_cast<const int16_t*>(src), dst, n, src_stride, dst_stride);
  else if (fmt == SONAR_PCM_F32)
    widen_pcm_kernel<float><<<wg, 256, 0, st>>>(static_cast<const float*>(src), dst, n, src_stride, dst_stride);
  else
    return set_error(SONAR_ERR_INVALID, "unknown PCM sample format");
  SONAR_CUDA(cudaGetLastError());
  return SONAR_OK;
}

namespace {

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
  G->path_tail = std::min<int64_t>(G->path_cap, (int64_t)G->dtw_len + G->dtw_len / 8 + 64);
  if (sizeof(double) * (size_t)(G->g.n_off + 2) > 140 * 1024)
    return set_error(SONAR_ERR_UNSUPPORTED, "sonar_align_pairs needs a Sakoe-Chiba band (dtw_band > 0) for long streams");
  G->z_pair = 2 * (size_t)((G->Te + 1) & ~(int64_t)1);
  G->corr_pair = (size_t)((G->nl + 1) & ~(int64_t)1);
  return SONAR_OK;
}

// Byte layout of one chunk of C pairs: scratch (d_tmp) and results (d_out, mirrored in pinned h_out).
struct ChunkLayout {
  int C = 0;
  // d_tmp
  size_t t_fp = 0, t_z = 0, t_cells = 0, t_screen = 0, tmp_bytes = 0;
  XcorrScreen xs{};  // scratch geometry of the screened cross-correlation
  // d_out / h_out
  size_t o_feat = 0, o_corr = 0, o_pc = 0, o_pq = 0, o_pr = 0, o_seqs = 0, o_pairs = 0, o_xo = 0, o_dout = 0, o_qptr = 0,
         o_rptr = 0, out_bytes = 0;
};

ChunkLayout chunk_layout(const PairGeom& G, int C) {
  ChunkLayout L;
  L.C = C;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t r = o;
    o += up(bytes);
    return r;
  };
  L.t_fp = take(sizeof(double) * 2 * (size_t)C * G.sh.tmp_doubles_per_stream);
  L.t_z = take(sizeof(double) * (size_t)C * G.z_pair);
  L.t_cells = take(sizeof(double) * (size_t)C * (size_t)G.g.cells);
  L.xs = xcorr_screen_geom(G.Te, G.aml, G.nl, C);
  L.t_screen = take(L.xs.bytes);
  L.tmp_bytes = o;
  o = 0;
  L.o_feat = take(sizeof(double) * 2 * (size_t)C * (size_t)G.sh.L.total);
  L.o_corr = take(sizeof(double) * (size_t)C * G.corr_pair);
  L.o_pc = take(sizeof(double) * (size_t)C * (size_t)G.path_cap);
  L.o_pq = take(sizeof(int32_t) * (size_t)C * (size_t)G.path_cap);
  L.o_pr = take(sizeof(int32_t) * (size_t)C * (size_t)G.path_cap);
  L.o_seqs = take(sizeof(XcorrSeq) * 2 * (size_t)C);
  L.o_pairs = take(sizeof(XcorrPair) * (size_t)C);
  L.o_xo = take(sizeof(XcorrPairOut) * (size_t)C);
  L.o_dout = take(sizeof(DtwPairOut) * (size_t)C);
  L.o_qptr = take(sizeof(double*) * (size_t)C);
  L.o_rptr = take(sizeof(double*) * (size_t)C);
  L.out_bytes = o;
  return L;
}

inline size_t sample_bytes(int fmt) { return pcm_sample_bytes(fmt); }

template <class T>
T* at(void* base, size_t off) {
  return reinterpret_cast<T*>(static_cast<unsigned char*>(base) + off);
}
template <class T>
const T* at(const void* base, size_t off) {
  return reinterpret_cast<const T*>(static_cast<const unsigned char*>(base) + off);
}

// A chunk of c pairs: the fingerprint of the 2c streams on `st` (these kernels fill the GPU), and, forked off as
// soon as the short-time energies exist, the alignment branch on `st2`: z-score, NCC, peak metrics, trim, banded
// DTW.  Its z-score / DTW / backtrack-chain kernels are latency bound (one warp per sequence or pair), so they
// hide under the YIN and loudness kernels that follow the energies on `st` instead of stalling one stream.
// Stream 2i is pair i's query, 2i+1 its reference, `stride` apart starting at pcm_dev.
int enqueue_chunk(sonar_ctx* ctx, int device, const sonar_fp_params* p, const PairGeom& G, const ChunkLayout& L, int c,
                  const double* pcm_dev, void* d_tmp, void* d_out, cudaStream_t st, cudaStream_t st2, cudaEvent_t mid,
                  cudaEvent_t fpdone, bool exact_curve, bool feat_travel) {
  double* feat = at<double>(d_out, L.o_feat);
  // the DTW fill of this (or the previous) chunk sits on ceil(c / kDtwPairsPerCta) SMs for several milliseconds beside the
  // persistent STFT kernels: they leave that many SMs unclaimed (common.h)
  tl_stft_sm_reserve = std::min((c + kDtwPairsPerCta - 1) / kDtwPairsPerCta, 16);
  int rc = enqueue_fingerprint(ctx, device, p, G.sh, pcm_dev, G.n, G.stride, 2 * c, feat, at<double>(d_tmp, L.t_fp), st, mid);
  tl_stft_sm_reserve = 0;
  if (rc) return rc;
  SONAR_CUDA(cudaEventRecord(fpdone, st));
  SONAR_CUDA(cudaStreamWaitEvent(st2, mid, 0));
  // a caller that takes the correlation curve home gets every lag in reference order; otherwise only the lags that
  // can be the peak or the second peak are (xcorr_fft.cu)
  static const bool full = std::getenv("SONAR_NCC_FULL") != nullptr;
  const bool screen = !exact_curve && !full;
  std::vector<XcorrSeq> seqs(2 * (size_t)c);
  std::vector<XcorrPair> pairs(c);
  double* z = at<double>(d_tmp, L.t_z);
  double* corr = at<double>(d_out, L.o_corr);
  for (int i = 0; i < c; i++) {
    const double* ea = feat + (size_t)(2 * i) * G.sh.L.total + G.sh.L.short_time_energy;
    const double* eb = feat + (size_t)(2 * i + 1) * G.sh.L.total + G.sh.L.short_time_energy;
    double* za = z + (size_t)i * G.z_pair;
    double* zb = za + G.z_pair / 2;
    seqs[2 * i] = XcorrSeq{ea, za, G.Te, screen ? xcorr_screen_prefix(L.xs, at<unsigned char>(d_tmp, L.t_screen), 2 * i) : nullptr};
    seqs[2 * i + 1] =
        XcorrSeq{eb, zb, G.Te, screen ? xcorr_screen_prefix(L.xs, at<unsigned char>(d_tmp, L.t_screen), 2 * i + 1) : nullptr};
    pairs[i] = XcorrPair{za, zb, corr + (size_t)i * G.corr_pair, G.Te, G.Te, 0, G.nl, G.aml, 0};
  }
  XcorrSeq* d_seqs = at<XcorrSeq>(d_out, L.o_seqs);
  XcorrPair* d_pairs = at<XcorrPair>(d_out, L.o_pairs);
  XcorrPairOut* d_xo = at<XcorrPairOut>(d_out, L.o_xo);
  // pageable sources: staged by the runtime before these calls return
  SONAR_CUDA(cudaMemcpyAsync(d_seqs, seqs.data(), sizeof(XcorrSeq) * seqs.size(), cudaMemcpyHostToDevice, st2));
  SONAR_CUDA(cudaMemcpyAsync(d_pairs, pairs.data(), sizeof(XcorrPair) * pairs.size(), cudaMemcpyHostToDevice, st2));
  if ((rc = launch_znorm(d_seqs, 2 * c, st2))) return rc;
  if (!screen)
    rc = launch_xcorr(d_pairs, c, G.nl, st2);
  else
    rc = launch_xcorr_screened(d_pairs, c, G.nl, L.xs, at<unsigned char>(d_tmp, L.t_screen), st2);
  if (rc) return rc;
  if ((rc = launch_xcorr_finalize(d_pairs, c, -1, d_xo, st2))) return rc;
  if ((rc = launch_xcorr_trim(d_seqs, d_pairs, d_xo, c, at<const double*>(d_out, L.o_qptr), at<const double*>(d_out, L.o_rptr),
                              st2)))
    return rc;
  rc = launch_dtw(nullptr, nullptr, c, G.g, 1, SONAR_STEP_SYMMETRIC2, at<double>(d_tmp, L.t_cells), nullptr,
                  at<int32_t>(d_out, L.o_pq), at<int32_t>(d_out, L.o_pr), at<double>(d_out, L.o_pc), G.path_cap,
                  at<DtwPairOut>(d_out, L.o_dout), st2, at<const double*>(d_out, L.o_qptr),
                  at<const double*>(d_out, L.o_rptr));
  if (rc) return rc;
  // the result copy that follows on st2 carries the features only when somebody asked for a feature array; otherwise
  // it (and the host-side scatter behind it) need not wait for the rest of the fingerprint
  if (feat_travel) SONAR_CUDA(cudaStreamWaitEvent(st2, fpdone, 0));
  return SONAR_OK;
}

// host side of pair i of a finished chunk: h = pinned copy of the chunk's result block
int finish_pair(const sonar_fp_params* p, const PairGeom& G, const ChunkLayout& L, const void* h, int i, bool feat,
                sonar_pair_out* o, const void* d_out) {
  if (feat) {
    const double* f = at<double>(h, L.o_feat) + (size_t)(2 * i) * G.sh.L.total;
    std::thread ref_side([&] { scatter_block(f + G.sh.L.total, G.sh, &o->reference); });
    scatter_block(f, G.sh, &o->query);
    ref_side.join();
  }
  const double* corr = at<double>(h, L.o_corr) + (size_t)i * G.corr_pair;
  const double* pc = at<double>(h, L.o_pc) + (size_t)i * G.path_cap;
  const int32_t* pq = at<int32_t>(h, L.o_pq) + (size_t)i * G.path_cap;
  const int32_t* pr = at<int32_t>(h, L.o_pr) + (size_t)i * G.path_cap;
  const XcorrPairOut& xo = at<XcorrPairOut>(h, L.o_xo)[i];
  const DtwPairOut& dout = at<DtwPairOut>(h, L.o_dout)[i];
  summarize_xcorr(xo, G.aml, G.Te, G.Te, G.nl, &o->xcorr);
  fill_align_from_xcorr(&o->xcorr, G.Te, G.Te, G.aml, p->energy_hop, p->call_sample_rate, &o->corr_alignment);
  if (o->corr) std::memcpy(o->corr, corr, sizeof(double) * (size_t)G.nl);
  o->dtw_length = G.dtw_len;
  sonar_dtw_out& w = o->dtw;
  const int64_t len = dout.path_len;
  if (len > G.path_tail && len <= G.path_cap && (w.path_query || w.path_ref || w.path_cost)) {
    // a path longer than the tail that travelled (a walk that wanders inside the band): fetch its head now
    const size_t lo = (size_t)(G.path_cap - len), cnt = (size_t)(len - G.path_tail);
    void* hm = const_cast<void*>(h);
    SONAR_CUDA(cudaMemcpy(at<double>(hm, L.o_pc) + (size_t)i * G.path_cap + lo,
                          at<double>(d_out, L.o_pc) + (size_t)i * G.path_cap + lo, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
    SONAR_CUDA(cudaMemcpy(at<int32_t>(hm, L.o_pq) + (size_t)i * G.path_cap + lo,
                          at<int32_t>(d_out, L.o_pq) + (size_t)i * G.path_cap + lo, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost));
    SONAR_CUDA(cudaMemcpy(at<int32_t>(hm, L.o_pr) + (size_t)i * G.path_cap + lo,
                          at<int32_t>(d_out, L.o_pr) + (size_t)i * G.path_cap + lo, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost));
  }
  w.path_len = len;
  w.total_cost = dout.total_cost;
  w.distance = dout.total_cost / (double)len;
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

bool wants_features(const sonar_fp_out& o) {
  const double* const ptrs[] = {o.mfcc, o.spectral_centroid, o.spectral_rolloff, o.spectral_bandwidth,
                                o.spectral_flatness, o.spectral_crest, o.spectral_slope, o.spectral_flux,
                                o.zero_crossing_rate, o.short_time_energy, o.energy_entropy, o.low_energy_ratio,
                                o.high_energy_ratio, o.pitch_estimate, o.pitch_confidence, o.voicing_strength,
                                o.harmonic_ratio, o.inharmonicity_ratio, o.tonal_centroid};
  for (const double* q : ptrs)
    if (q) return true;
  return false;
}

constexpr size_t kPairChunkBytes = (size_t)256 << 20;  // host PCM bytes staged per chunk (measured: scripts/e2e_*_sweep.py)

// pcm_q / pcm_r: host pointers (host_pcm) or, device-resident, pcm_q[i] = device pointer of pair i's query with the
// reference `stride` behind it and consecutive pairs 2*stride apart.
void run_pairs_device(sonar_ctx* ctx, DevCtx* dev, const double* const* pcm_q, const double* const* pcm_r,
                      const std::vector<int>* ids, const sonar_fp_params* p, const PairGeom* Gp, bool host_pcm, int fmt,
                      sonar_pair_out* outs, DevJob* job) {
  const PairGeom& G = *Gp;
  set_current_ctx(ctx);
  auto fail = [&](int rc) {
    job->rc = rc;
    job->err = sonar_last_error();
  };
  cudaError_t e = cudaSetDevice(dev->device);
  if (e != cudaSuccess) return fail(cuda_error(e, "cudaSetDevice"));
  constexpr int NL = DevCtx::kSlots;       // lanes: result + scratch buffers, second stream, one chunk in flight each
  constexpr int NP = DevCtx::kStageSlots;  // PCM staging buffers + front-half streams
  const int total = (int)ids->size();
  // pairs per chunk: host path = what fits the staging budget (the PCIe copy of the next chunk overlaps this one);
  // device-resident = contiguous runs as large as possible (batched launches hide the latency-bound kernels)
  static const int dev_chunk = [] {
    const char* e = std::getenv("SONAR_PAIR_CHUNK");
    const int v = e ? std::atoi(e) : 0;
    return v > 0 ? v : 32;  // 16 overlaps the chunks' kernels (-3 % step) but they then time-share the SMs
  }();
  static const size_t host_chunk_bytes = [] {
    const char* e = std::getenv("SONAR_PAIR_CHUNK_MB");
    const long v = e ? std::atol(e) : 0;
    return v > 0 ? (size_t)v << 20 : kPairChunkBytes;
  }();
  // host path: pairs per chunk by the bytes that cross PCIe (narrow formats travel in proportionally larger chunks:
  // the kernels run better on bigger batches and the copy of a chunk costs the same)
  int C = host_pcm ? (int)std::max<size_t>(1, host_chunk_bytes / (sample_bytes(fmt) * 2 * (size_t)G.stride)) : dev_chunk;
  static const int host_cap = [] {
    const char* e = std::getenv("SONAR_PAIR_CHUNK_MAX");
    const int v = e ? std::atoi(e) : 0;
    return v > 0 ? v : 4;
  }();
  if (host_pcm) C = std::min(C, host_cap);  // deeper pipelines beat bigger batches: 32 pairs of int16 take 63 ms at 4, 72 ms at 9 per chunk
  C = std::min(C, total);
  const ChunkLayout L = chunk_layout(G, C);
  struct Pending {
    int first = -1, count = 0;
    bool feat = false;
  } pending[NL];
  auto finish = [&](int li) -> int {
    if (pending[li].first < 0) return SONAR_OK;
    Slot& s = dev->slot[li];
    SONAR_CUDA(cudaEventSynchronize(s.done));
    const Pending pd = pending[li];
    pending[li].first = -1;
    // the scatter into the caller's arrays is plain memcpy at one core's bandwidth; with the feature blocks on
    // board (~24 MB per pair) it would outlast the GPU work, so the pairs of a chunk go to worker threads
    std::vector<int> rcs(pd.count, SONAR_OK);
    std::vector<std::string> errs(pd.count);
    auto one = [&](int i) {
      rcs[i] = finish_pair(p, G, L, s.h_out.p, i, pd.feat, &outs[(*ids)[pd.first + i]], s.d_out.p);
      if (rcs[i]) errs[i] = sonar_last_error();
    };
    if (pd.feat && pd.count > 1) {
      std::vector<std::thread> th;
      for (int i = 1; i < pd.count; i++) th.emplace_back(one, i);
      one(0);
      for (auto& t : th) t.join();
    } else {
      for (int i = 0; i < pd.count; i++) one(i);
    }
    // without feature arrays the copy did not wait for the fingerprint kernels (the scatter above ran beside them);
    // they still belong to the call, and the lane's buffers are theirs until they end
    if (!pd.feat) SONAR_CUDA(cudaEventSynchronize(s.fpdone));
    for (int i = 0; i < pd.count; i++)
      if (rcs[i]) return set_error(rcs[i], errs[i]);
    return SONAR_OK;
  };
  int k = 0;
  for (int first = 0; first < total; ++k) {
    int c = std::min(C, total - first);
    if (!host_pcm)  // a device-resident chunk must be contiguous in the caller's layout
      for (int i = 1; i < c; i++)
        if (pcm_q[(*ids)[first + i]] != pcm_q[(*ids)[first]] + (int64_t)2 * i * G.stride) {
          c = i;
          break;
        }
    const int li = k % NL, pi = k % NP;
    Slot& lane = dev->slot[li];
    Slot& stage = dev->slot[pi];
    int rc = finish(li);
    if (rc) return fail(rc);
    if ((host_pcm && (rc = dev->ensure_dev(stage.d_in, sizeof(double) * 2 * (size_t)C * (size_t)G.stride))) ||
        (rc = dev->ensure_dev(lane.d_tmp, L.tmp_bytes)) || (rc = dev->ensure_dev(lane.d_out, L.out_bytes)) ||
        (rc = dev->ensure_host(lane.h_out, L.out_bytes)))
      return fail(rc);
    const double* pcm_dev;
    if (host_pcm) {
      // the staging buffer's previous user (chunk k - NP) ran its fingerprint kernels on this same stream: ordered
      double* d_in = static_cast<double*>(stage.d_in.p);
      const size_t sb = sample_bytes(fmt);
      unsigned char* d_raw = nullptr;
      const int64_t raw_stride = (G.n + 7) & ~(int64_t)7;  // samples; keeps every stream 16-byte aligned
      if (fmt != SONAR_PCM_F64) {
        if ((rc = dev->ensure_dev(stage.d_raw, sb * 2 * (size_t)C * (size_t)raw_stride))) return fail(rc);
        d_raw = static_cast<unsigned char*>(stage.d_raw.p);
      }
      for (int i = 0; i < c; i++) {
        const int id = (*ids)[first + i];
        const void* hq = pcm_q[id];
        const void* hr = pcm_r[id];
        void* dq = d_raw ? (void*)(d_raw + sb * (size_t)(2 * i) * raw_stride) : (void*)(d_in + (int64_t)(2 * i) * G.stride);
        void* dr = d_raw ? (void*)(d_raw + sb * (size_t)(2 * i + 1) * raw_stride)
                         : (void*)(d_in + (int64_t)(2 * i + 1) * G.stride);
        if ((e = cudaMemcpyAsync(dq, hq, sb * (size_t)G.n, cudaMemcpyHostToDevice, stage.st)) != cudaSuccess ||
            (e = cudaMemcpyAsync(dr, hr, sb * (size_t)G.n, cudaMemcpyHostToDevice, stage.st)) != cudaSuccess)
          return fail(cuda_error(e, "cudaMemcpyAsync(H2D pcm)"));
      }
      if (d_raw && (rc = launch_widen_pcm(d_raw, fmt, d_in, G.n, raw_stride, G.stride, 2 * c, stage.st))) return fail(rc);
      pcm_dev = d_in;
    } else {
      pcm_dev = pcm_q[(*ids)[first]];
    }
    bool curve = false;  // somebody in the chunk takes the correlation curve home
    for (int i = 0; i < c && !curve; i++) curve = outs[(*ids)[first + i]].corr != nullptr;
    bool feat = false;  // the feature blocks only travel when somebody asked for a feature array
    for (int i = 0; i < c && !feat; i++) {
      const sonar_pair_out& o = outs[(*ids)[first + i]];
      feat = wants_features(o.query) || wants_features(o.reference);
    }
    rc = enqueue_chunk(ctx, dev->device, p, G, L, c, pcm_dev, lane.d_tmp.p, lane.d_out.p, stage.st, lane.st2, lane.mid,
                       lane.fpdone, curve, feat);
    if (rc) return fail(rc);
    // Result copy: [features] [curve] | the TAILS of the three path arrays (42 MB of capacity per 32 pairs, half of it
    // unused: at 8 ranks the host's D2H ceiling of 92 GB/s made the full copy 3.6 ms of a 25 ms step) | descriptors
    const size_t from = feat ? 0 : (curve ? L.o_corr : L.o_pc);
    unsigned char* hb = static_cast<unsigned char*>(lane.h_out.p);
    unsigned char* db = static_cast<unsigned char*>(lane.d_out.p);
    if (from < L.o_pc && (e = cudaMemcpyAsync(hb + from, db + from, L.o_pc - from, cudaMemcpyDeviceToHost, lane.st2)) != cudaSuccess)
      return fail(cuda_error(e, "cudaMemcpyAsync(D2H results)"));
    {
      const size_t skip = (size_t)(G.path_cap - G.path_tail);
      const struct { size_t off, elem; } arr[3] = {{L.o_pc, sizeof(double)}, {L.o_pq, sizeof(int32_t)}, {L.o_pr, sizeof(int32_t)}};
      for (const auto& a : arr) {
        const size_t pitch = a.elem * (size_t)G.path_cap, o0 = a.off + a.elem * skip;
        if ((e = cudaMemcpy2DAsync(hb + o0, pitch, db + o0, pitch, a.elem * (size_t)G.path_tail, (size_t)c,
                                   cudaMemcpyDeviceToHost, lane.st2)) != cudaSuccess)
          return fail(cuda_error(e, "cudaMemcpy2DAsync(D2H paths)"));
      }
    }
    if ((e = cudaMemcpyAsync(hb + L.o_seqs, db + L.o_seqs, L.out_bytes - L.o_seqs, cudaMemcpyDeviceToHost, lane.st2)) != cudaSuccess)
      return fail(cuda_error(e, "cudaMemcpyAsync(D2H results)"));
    if ((e = cudaEventRecord(lane.done, lane.st2)) != cudaSuccess) return fail(cuda_error(e, "cudaEventRecord"));
    pending[li].first = first;
    pending[li].count = c;
    pending[li].feat = feat;
    first += c;
  }
  for (int j = 0; j < NL; j++) {
    int rc = finish((k + j) % NL);  // oldest first: their host-side scatter overlaps the GPU work still queued
    if (rc) return fail(rc);
  }
}

int align_pairs(sonar_ctx* ctx, const double* const* q, const double* const* r, int64_t n, int n_pairs,
                const sonar_fp_params* p, double max_lag_seconds, int dtw_band, bool host_pcm, int fmt,
                sonar_pair_out* outs) {
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
    run_pairs_device(ctx, &ctx->devs[0], q, r, &ids[0], p, &G, host_pcm, fmt, outs, &res[0]);
  } else {
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++)
      th.emplace_back(run_pairs_device, ctx, &ctx->devs[d], q, r, &ids[d], p, &G, host_pcm, fmt, outs, &res[d]);
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
  return align_pairs(ctx, query_pcm, reference_pcm, n, n_pairs, p, max_lag_seconds, dtw_band, true, SONAR_PCM_F64, outs);
}

int sonar_align_pairs_pcm(sonar_ctx* ctx, const void* const* query_pcm, const void* const* reference_pcm, int sample_format,
                          int64_t n, int n_pairs, const sonar_fp_params* p, double max_lag_seconds, int dtw_band,
                          sonar_pair_out* outs) {
  if (sample_format != SONAR_PCM_F64 && sample_format != SONAR_PCM_F32 && sample_format != SONAR_PCM_S16)
    return set_error(SONAR_ERR_INVALID, "unknown PCM sample format");
  return align_pairs(ctx, reinterpret_cast<const double* const*>(query_pcm),
                     reinterpret_cast<const double* const*>(reference_pcm), n, n_pairs, p, max_lag_seconds, dtw_band, true,
                     sample_format, outs);
}

int sonar_align_pairs_dev(sonar_ctx* ctx, const double* pcm_dev, int64_t n, int64_t stride, int n_pairs,
                          const sonar_fp_params* p, double max_lag_seconds, int dtw_band, sonar_pair_out* outs) {
  if (!pcm_dev) return set_error(SONAR_ERR_INVALID, "audio data cannot be nil");
  if (stride != ((n + 1) & ~(int64_t)1)) return set_error(SONAR_ERR_INVALID, "stride must be n rounded up to even");
  std::vector<const double*> q(n_pairs > 0 ? n_pairs : 0);
  for (int i = 0; i < n_pairs; i++) q[i] = pcm_dev + (int64_t)2 * i * stride;
  return align_pairs(ctx, q.data(), nullptr, n, n_pairs, p, max_lag_seconds, dtw_band, false, SONAR_PCM_F64, outs);
}

}  // extern "C"
