// Internal declarations shared by the CUDA library sources (libsonar.so).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sonar.h"

namespace sonar {

// ---- errors ----------------------------------------------------------------
int set_error(int code, const std::string& msg);
int cuda_error(cudaError_t e, const char* what);
const std::string& last_error_string();
#define SONAR_CUDA(call)                                   \
  do {                                                     \
    cudaError_t e__ = (call);                              \
    if (e__ != cudaSuccess) return ::sonar::cuda_error(e__, #call); \
  } while (0)

// ---- host-side table construction (host_tables.cpp) ------------------------
// All tables are built in float64 exactly as the reference's constructors do
// and rounded to float32 once.
int host_window(int type, int n, bool symmetric, bool normalize, double beta, double alpha,
                double* w);
// mel_scale.go:29-86 bin points; returns false when a non-empty segment would
// index a negative bin (the reference would panic).
bool host_mel_bins(int n_mel, int fft_size, int sample_rate, double low, double high,
                   std::vector<int64_t>& bins);
void resolve_mfcc(const sonar_fp_params* p, int* n_mfcc, int* n_mel, double* high, double* lifter);
int host_fp_sizes(const sonar_fp_params* p, int64_t n, sonar_fp_sizes_t* out);
int host_fp_layout(const sonar_fp_params* p, int64_t n, sonar_fp_dev_layout_t* out);
int actual_max_lag(int max_lag, int64_t l1, int64_t l2);
int64_t overlap_len(int64_t l1, int64_t l2, int64_t lag);
void xcorr_derive(sonar_xcorr_summary* s, int64_t na, int64_t nb);  // p-value, overlap, significance
double corr_confidence(const sonar_xcorr_summary* c);
double corr_quality(const sonar_xcorr_summary* c, int max_lag);
int host_align_dtw_scalars(const sonar_dtw_out* d, int n, int m, int sr, sonar_align_result* out);

// ---- fused STFT + feature kernel (stft_features.cu) ------------------------
struct MelRegion {  // bins [prev next_b, next_b): fall=(bhi-k)*inv_f -> acc[r-1], rise=(k-blo)*inv_r -> acc[r]
  int next_b;
  float blo, bhi, inv_f, inv_r;
};

constexpr int kMaxMel = 64;
constexpr int kMaxMfcc = 32;

struct FpPlan {  // immutable once built; cached per context keyed by the parameters
  int R1 = 0, R2 = 0, N = 0, hop = 0, B = 0;
  int n_mfcc = 0, n_mel = 0, n_regions = 0;
  int algo_sr = 0, split = 0, slope_on = 0;
  double freq_scale = 0;  // sr / N
  double slope_ntot = 0, slope_xxtot = 0;
  // one device blob; offsets in bytes
  void* d_blob = nullptr;
  size_t blob_bytes = 0;
  size_t off_win2 = 0, off_tw1 = 0, off_wn = 0, off_xtab = 0, off_dct = 0, off_lift = 0, off_regions = 0,
         off_chunk_region = 0, off_hann = 0, off_zero = 0;
  // float64 tables of the exact re-evaluation (spectral_exact.cu): window, go-dsp radix-2 factors, mel bin points,
  // DCT-II, lifter; float32 1 / (sum of a filter's weights) for the fused kernel's weak-band test
  size_t off_win64 = 0, off_fac64 = 0, off_melbins = 0, off_dct64 = 0, off_lift64 = 0, off_melinvw = 0, off_chirp = 0,
         off_bluefb = 0;
  bool exact_only = false;  // no fused FP32 kernel for this length: every frame takes the float64 route (spectral_exact.cu)
  int fft_len = 0;          // its radix-2 transform length: N, or NextPowerOf2(2 N - 1) on the Bluestein route
  std::vector<MelRegion> h_regions;  // host copy of the mel region table (kernel eligibility checks)
  ~FpPlan();
};

struct StftArgs {
  const double* pcm;
  int64_t n, stride;
  int n_streams;
  int64_t T, Te;
  int hop;
  int runs_per_stream;
  int64_t total_runs;
  // tables
  const float2* win2;
  const float2* tw1;
  const float2* wn;
  const float* xtab;
  const float* dct;
  const float* lift;
  const MelRegion* regions;
  const int* chunk_region;
  int n_regions, n_mel, n_mfcc, split, slope_on, mfcc_on;
  double freq_scale, slope_ntot, slope_xxtot;
  // features mode outputs (device layout, per-stream stride in doubles)
  double* feat;
  int64_t feat_stride;
  int64_t o_mfcc, o_centroid, o_rolloff, o_bandwidth, o_flatness, o_crest, o_slope, o_flux, o_low, o_high;
  // spectrum mode outputs (single stream)
  double* mag;
  double* phase;
  double* cplx;
  // frames whose discrete or ill-conditioned results need float64 (rolloff bin within the FP32 error of the 85 %
  // threshold, bins or mel bands at the FP32 transform's noise floor, silence): per stream a list [count, t...] of
  // ints, `xlist_stride` ints apart, re-evaluated in the reference's order by spectral_exact.cu.  nullptr: no lists.
  int* xlist;
  int64_t xlist_stride;
  unsigned* work_counter;  // zeroed before the launch: the persistent kernel's warps draw their runs from it (stft_v3.cu)
  const double* win64;
  const double2* fac64;
  const int* melbins;
  const double* dct64;
  const double* lift64;
  const float* mel_invw;
  int algo_sr, N;
  // float64 route for every frame (lengths without a fused kernel): no lists, frames 0 .. T-1, flux included;
  // fft_len > N: go-dsp's Bluestein (chirp_inv = conj w_i, blue_fb = FFT of the chirp sequence b)
  int exact_all, fft_len;
  const double2* chirp_inv;
  const double2* blue_fb;
  int v4_lockstep;  // stft_v4_kernel: warps of one role on one scheduler meet at a named barrier every iteration
};

// exact float64 re-evaluation of the frames the fused kernel listed (spectral_exact.cu)
int launch_spectral_exact(const StftArgs& a, cudaStream_t st);

int build_fp_plan(const sonar_fp_params* p, std::shared_ptr<FpPlan>* out);
int launch_stft_features(const FpPlan& plan, StftArgs& a, bool spectrum_mode, cudaStream_t st);
bool stft_supported(int window_size);
bool stft_exact_only(int window_size);
// second-generation kernel (stft_v2.cu): N = 1024, aligned PCM, features mode
bool stft_v2_eligible(const FpPlan& plan, const StftArgs& a);
int launch_stft_v2(const FpPlan& plan, StftArgs& a, cudaStream_t st);
// third generation (stft_v3.cu): 1024 / 256 and 512 / 160, two frames per complex FFT, register-resident sample ring
bool stft_v3_eligible(const FpPlan& plan, const StftArgs& a);
int launch_stft_v3(const FpPlan& plan, StftArgs& a, cudaStream_t st);
// fifth generation (stft_v5.cu): the same work as a transform kernel + a scan kernel, each within the 32 KB instruction cache
int launch_stft_v5(const FpPlan& plan, StftArgs& a, cudaStream_t st);
// SMs the persistent STFT kernels leave unclaimed for the launches that follow on the calling thread.  The pair pipeline
// sets it to the number of CTAs its DTW fill occupies (kDtwPairsPerCta pairs each) while it enqueues a chunk's fingerprint:
// both STFT kernels give every CTA a static share of the runs, so a CTA whose SM is held by the alignment branch starts late
// and the launch waits for it (scripts/sm_reserve_ab.py, 32 pairs: 19.8 ms per step with 0, 18.7 with 8, 19.2 with 16).
extern thread_local int tl_stft_sm_reserve;
constexpr int kDtwPairsPerCta = 4;
void stft_workspace_release(int device, cudaStream_t st);  // frees the pair's workspace of a stream about to be destroyed

// ---- fingerprint sequencing shared by the host-pointer, device-resident and pipeline entry points ----------
struct FpShape {
  sonar_fp_sizes_t sz;
  sonar_fp_dev_layout_t L;
  int64_t lr_win = 0, lr_hop = 0, lr_nw = 0;  // loudness-range windows (energy.go:157-179)
  size_t tmp_doubles_per_stream = 0;
  // temporal group (only with SONAR_FP_ENABLE_TEMPORAL): two more arrays behind the public layout (L.total is
  // bumped accordingly) and a partial-sum area in the scratch
  bool temporal = false;
  int64_t o_env = 0, o_att = 0, o_part = 0;
  int64_t o_wpart = 0, wpart_doubles = 0;  // scratch of the frame walk's loudness block parts (timedomain.cu)
  // speech-specific group (only with SONAR_FP_ENABLE_SPEECH): the gate (4 doubles) and the tilt array behind the public
  // layout, the gate's partial sums in the scratch
  bool speech = false;
  int64_t o_sgate = 0, o_tilt = 0, n_speech_frames = 0, o_spart = 0;
  int64_t o_work = 0;   // scratch of the fused STFT kernel: its work counter (stream 0's copy is used)
  int64_t o_xlist = 0;  // scratch of the fused STFT kernel: 1 + T ints (count, frames re-evaluated in float64; spectral_exact.cu)
  int64_t o_ylist = 0;  // scratch of the pitch detector: 1 + Tp ints (count, frames re-evaluated exactly; yin32.cu)
};

// ---- exact FP64 time-domain kernels (timedomain.cu) -------------------------
// o_* are offsets (doubles) into each stream's output block; < 0 = not wanted.
// Optional by-product of the frame walk: the RMS of the loudness windows (energy.go:157-179) from block sums of the
// squared pre-emphasised samples the walk forms anyway, instead of a second pass over the PCM (launch_rms_windows).
struct WalkLoudness {
  int64_t win, hop, nw;            // window / hop in samples, windows per stream
  double* part;                    // scratch: per stream 2 doubles per walk thread (ceil(Tn / 8) threads)
  int64_t part_stride;
  double* rms;                     // nw values per stream
  int64_t rms_stride;
};
// *wl_done tells whether the by-product was produced (sizes permitting); otherwise the caller runs launch_rms_windows
int launch_frame_walk(const double* pcm, int64_t n, int64_t stride, int n_streams, double alpha, int frame,
                      int hop, int64_t Tn, int sr, double* out, int64_t out_stride, int64_t o_energy,
                      int64_t o_entropy, int64_t o_zcr, cudaStream_t st, const WalkLoudness* wl = nullptr,
                      bool* wl_done = nullptr);
// ZCR of the lone incomplete frame of a stream shorter than the window (n < W, T == 1)
int launch_short_zcr(const double* pcm, int64_t n, int64_t stride, int n_streams, double alpha, int sr, double* out,
                     int64_t out_stride, int64_t o_zcr, cudaStream_t st);
int launch_variance(const double* x, int64_t n, int64_t stride, int n_streams, double* out, int64_t out_stride,
                    cudaStream_t st);
int launch_rms_windows(const double* pcm, int64_t n, int64_t stride, int n_streams, double alpha, int win,
                       int hop, int64_t nw, double* out, int64_t out_stride, cudaStream_t st);
int launch_loudness_range(double* rms, int64_t nw, int64_t in_stride, int n_streams, double* out,
                          int64_t out_stride, cudaStream_t st);
int launch_fill(double* p, int64_t n, double v, cudaStream_t st);
int launch_fill_strided(double* p, int64_t count, int64_t stride, int n_streams, double v, cudaStream_t st);

// ---- YIN (yin.cu) ------------------------------------------------------------
// scratch: per stream 2*Tp doubles, scratch_stride apart; hann_dev: un-normalised symmetric Hann(1024)
int launch_yin(const double* pcm, int64_t stride, int n_streams, double alpha, int sr, int64_t Tp,
               const double* hann_dev, double* feat, int64_t feat_stride, int64_t o_pitch, int64_t o_conf,
               int64_t o_voicing, int64_t o_hratio, int64_t o_inharm, int64_t o_tonal, double* scratch,
               int64_t scratch_stride, cudaStream_t st, cudaStream_t track_st = nullptr, cudaEvent_t fork = nullptr,
               cudaEvent_t join = nullptr, bool* forked = nullptr, int* lists = nullptr, int64_t list_stride = 0,
               const double* speech_gate = nullptr, int64_t gate_stride = 0);
// speech-specific group (speech.cu): IsSpeech gate (4 doubles at o_gate: flag, zcr, rms, periodicity) and spectral tilt
size_t speech_gate_scratch_doubles();
int launch_speech(const double* pcm, int64_t n, int64_t stride, int n_streams, double alpha, int sr, int64_t nf, double* feat,
                  int64_t feat_stride, int64_t o_gate, int64_t o_tilt, double* scratch, int64_t scratch_stride,
                  cudaStream_t st);
// FP32 difference function on the packed pipe + exact float64 re-evaluation of the borderline frames (yin32.cu);
// lists: per stream 1 + Tp ints, list_stride ints apart, counts zeroed by the caller
int launch_yin32(const double* pcm, int64_t stride, int n_streams, double alpha, int sr, int64_t Tp, const double* hann_dev,
                 double* scratch, int64_t scratch_stride, int* lists, int64_t list_stride, cudaStream_t st);

// ---- cross-correlation (xcorr.cu) ------------------------------------------
struct XcorrSeq {  // one sequence to z-score (population sigma, reference summation order)
  const double* in;
  double* out;
  int64_t n;
  double* prefix = nullptr;  // optional: n + 2 doubles, running sum of squared deviations + scale (xcorr.cu znorm_kernel)
};
struct XcorrPair {  // one pair, or one lag shard [idx_lo, idx_hi) of a pair
  const double* za;
  const double* zb;
  double* corr;  // idx_hi - idx_lo values
  int64_t na, nb, idx_lo, idx_hi;
  int32_t aml, pad;
};
struct XcorrPairOut {  // device result of one pair (or one lag shard of a pair)
  double peak;            // c[peak_index]
  double noise_sum, noise_cnt, max_sidelobe;
  double second_abs, second_val;
  double c_peak, c_prev, c_next;  // around the (possibly overridden) peak; NaN when not inside [idx_lo, idx_hi)
  int64_t peak_index;     // shard-local arg-max as a global lag index, -1 = empty shard
  int64_t second_index;   // -1 = none
  int32_t n_candidates, pad;
};
int launch_znorm(const XcorrSeq* seqs_dev, int count, cudaStream_t st);
int launch_xcorr(const XcorrPair* pairs_dev, int n_pairs, int64_t max_shard_lags, cudaStream_t st);
int launch_xcorr_finalize(const XcorrPair* pairs_dev, int n_pairs, int64_t peak_override, XcorrPairOut* outs_dev,
                          cudaStream_t st);
// lags per CTA of the flagged form of ncc_tiled_kernel (xcorr.cu); the screen (xcorr_fft.cu) flags CTAs in these units
constexpr int kNccFlagLags = 64;
// exact NCC only for the CTAs whose flag is set: need[pair * need_stride + blockIdx.x], grid.x = 2 * ceil(max_shard_lags / 64)
int launch_xcorr_flagged(const XcorrPair* pairs_dev, int n_pairs, int64_t max_shard_lags, const unsigned char* need_dev,
                         int need_stride, cudaStream_t st);
// Screened NCC (xcorr_fft.cu): FFT estimate of the whole curve, exact values wherever the peak or the second peak
// can be.  The curve left in XcorrPair::corr is exact only there: callers that return the curve use launch_xcorr.
struct XcorrScreen {
  int log2n, bps, need_stride, n_pairs;
  int64_t N, pre_stride;
  size_t o_x, o_y, o_pre, o_need, bytes;
};
XcorrScreen xcorr_screen_geom(int64_t max_n, int aml_max, int64_t max_shard_lags, int n_pairs);
// where znorm_kernel leaves sequence `seq`'s prefix sums (XcorrSeq::prefix) for launch_xcorr_screened
double* xcorr_screen_prefix(const XcorrScreen& g, void* scratch, int seq);
int launch_xcorr_screened(const XcorrPair* pairs_dev, int n_pairs, int64_t max_shard_lags, const XcorrScreen& g,
                          void* scratch, cudaStream_t st);
// TruncateToAlignmentPCM's convention on the feature series (extractors/alignment.go:239-243): a positive lag skips
// the start of the second sequence, a negative one the start of the first.  Writes the trimmed start pointers.
int launch_xcorr_trim(const XcorrSeq* seqs_dev, const XcorrPair* pairs_dev, const XcorrPairOut* outs_dev, int n_pairs,
                      const double** qptr_dev, const double** rptr_dev, cudaStream_t st);

// ---- DTW (dtw.cu) -------------------------------------------------------------
struct DtwGeom {
  int n, m, band;   // band == 0: unconstrained
  int n_off;        // number of distinct offsets i - j the wavefront tracks
  int64_t W;        // banded: cells per stored anti-diagonal (band+1); unconstrained: cells per stored row (m)
  int64_t cells;    // doubles per pair: the cost store plus, for narrow bands, the backtrack side arrays below
  // narrow bands only (dirs_off > 0): one direction byte per cost cell (same diagonal-major indexing) and the
  // block tables of the parallel backtrack, as offsets in doubles from the start of the pair's region
  int64_t dirs_off, tbl_off, chain_off, flag_off;
  int bt_nb;        // number of kDtwBtDiags-diagonal blocks covering diagonals n+m .. 0
};
constexpr int kDtwBtDiags = 256;
int dtw_geometry(int n, int m, int band, DtwGeom* g);
struct DtwPairOut {
  double total_cost;
  int64_t path_len;
};
// q: n_pairs*n*dim, r: n_pairs*m*dim (device); cells: n_pairs*g.cells; line_scratch: n_pairs*(g.n_off+2)
// doubles, only needed when the offset line does not fit shared memory; paths are written back to front:
// pair p's path occupies [p*path_cap + path_cap - len, p*path_cap + path_cap).
// qptr / rptr (nullable, device arrays of n_pairs device pointers) override the packed q / r layout.
int launch_dtw(const double* q, const double* r, int n_pairs, const DtwGeom& g, int dim, int step, double* cells,
               double* line_scratch, int32_t* path_q, int32_t* path_r, double* path_c, int64_t path_cap,
               DtwPairOut* out, cudaStream_t st, const double* const* qptr = nullptr,
               const double* const* rptr = nullptr);
int launch_dtw_expand(const double* cells, const DtwGeom& g, double* full, cudaStream_t st);

// ---- music-extractor spectral additions (music_spectral.cu) -----------------
int launch_music_spectral(const double* mag, int64_t T, int B, int n_bands, const int* edges, double* contrast,
                          const signed char* cmap, double* chroma, int n_bark, const double* bank,
                          const int2* bank_range, double* bark, cudaStream_t st);
// host tables in float64 as the reference's constructors build them (host_tables.cpp)
std::vector<int> host_contrast_edges(int n_bands, int num_bins, int sample_rate);
std::vector<signed char> host_chroma_map(int freq_bins, double freq_resolution);
std::vector<double> host_bark_bank(int n_filters, int fft_size, int sample_rate, double low, double high);

// ---- column statistics (colstats.cu) ----------------------------------------
int launch_colstats(const double* x, int64_t t, int dim, double* stats, cudaStream_t st);
struct ColJob {  // one column of one array: x[off + i * dim], i < t; results at out[out_mean], out[out_std]
  int64_t off, t;
  int dim;
  int64_t out_mean, out_std;
};
int launch_colstats_batch(const double* base, const ColJob* jobs, int n_jobs, double* out, cudaStream_t st);

// ---- context ------------------------------------------------------------------
struct Buf {  // growable allocation (device or pinned host)
  void* p = nullptr;
  size_t bytes = 0;
};
struct Slot {  // one in-flight unit of work on a device: its streams and buffers
  cudaStream_t st = nullptr;   // H2D + the kernels that saturate the GPU
  cudaStream_t st2 = nullptr;  // latency-bound tails (DTW) + D2H of the pair pipeline, so they overlap other slots
  cudaEvent_t done = nullptr;
  cudaEvent_t mid = nullptr;   // hand-off st -> st2 (short-time energies ready)
  cudaEvent_t fpdone = nullptr;  // the rest of the fingerprint (st) has finished: st2 may copy the features out
  cudaStream_t st3 = nullptr;    // the pitch tracker (one warp per stream) beside the DRAM-bound loudness kernels on st
  cudaEvent_t fork = nullptr, join = nullptr;  // st -> st3 -> st
  Buf d_in, d_out, d_tmp, h_in, h_out;
  Buf d_raw;  // narrow (f32 / s16) PCM as it crossed PCIe, widened into d_in on the device
};
struct DevCtx {
  int device = 0;
  static constexpr int kSlots = 8;
  static constexpr int kStageSlots = 3;  // slots whose d_in stages host PCM
  Slot slot[kSlots];
  int ensure_dev(Buf& b, size_t bytes);
  int ensure_host(Buf& b, size_t bytes);
};

}  // namespace sonar

struct sonar_ctx {
  std::vector<sonar::DevCtx> devs;
  std::mutex mu;       // guards the plan cache
  std::mutex call_mu;  // one in-flight call per context (slots and their buffers are per context)
  std::map<std::string, std::shared_ptr<sonar::FpPlan>> plans;
  std::atomic<uint64_t> launches{0};
  // per-kernel CUDA-event timing (off by default)
  struct ProfRec {
    const char* name;
    cudaEvent_t a, b;
  };
  std::atomic<bool> profiling{false};
  std::mutex prof_mu;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> prof_pool;
  // where the most recent fingerprint batch keeps its lists of exactly re-evaluated frames (sonar_fp_exact_counts)
  struct LastLists {
    int device = -1;
    cudaStream_t st = nullptr;
    const double* tmp = nullptr;
    int64_t tstride = 0, o_xlist = 0, o_ylist = 0;
    int ns = 0;
  } last_lists;
  // lag-sharded cross-correlation over the ranks of a job (nccl_shard.cu): an NCCL communicator created from a unique id
  // the host distributes (sonar_nccl_init), and the resident workspace of the sharded call
  void* nccl_comm = nullptr;
  int nccl_world = 1, nccl_rank = 0;
  void* shard_buf = nullptr;
  size_t shard_bytes = 0;
};

namespace sonar {
// Every kernel launch is bracketed by prof_begin(name, stream) / prof_end(): the pair counts the
// launch on the calling thread's current context and, when profiling is enabled on it
// (sonar_profile_enable), records a CUDA event before and after the kernel on its own stream.
void prof_begin(const char* kernel, cudaStream_t st);
void prof_end();
void prof_count_launch();  // a second kernel launched inside one prof_begin / prof_end pair
void set_current_ctx(sonar_ctx* c);
void nccl_release(sonar_ctx* ctx);  // nccl_shard.cu
// speech.go:370-408 temporal block (temporal.cu)
int fp_validate(const sonar_fp_params* p);
int fp_shape(const sonar_fp_params* p, int64_t n, FpShape* s);
// enqueues every kernel of one uniform batch of streams on `st` (fingerprint_api.cu)
int enqueue_fingerprint(sonar_ctx* ctx, int device, const sonar_fp_params* p, const FpShape& sh, const double* pcm_dev,
                        int64_t n, int64_t stride, int ns, double* feat_dev, double* tmp_dev, cudaStream_t st,
                        cudaEvent_t energy_ready = nullptr);  // recorded on st once the short-time energies exist
void scatter_block(const double* f, const FpShape& sh, sonar_fp_out* o);
// f32 / s16 PCM rows (src_stride samples apart) -> float64 rows (pipeline_api.cu); fmt = SONAR_PCM_*
int launch_widen_pcm(const void* src, int fmt, double* dst, int64_t n, int64_t src_stride, int64_t dst_stride, int rows,
                     cudaStream_t st);
inline size_t pcm_sample_bytes(int fmt) { return fmt == SONAR_PCM_S16 ? 2 : (fmt == SONAR_PCM_F32 ? 4 : 8); }
void summarize_xcorr(const XcorrPairOut& o, int aml, int64_t na, int64_t nb, int64_t n_eval, sonar_xcorr_summary* s);
void fill_align_from_xcorr(const sonar_xcorr_summary* xc, int64_t nq, int64_t nr, int max_lag, int hop, int sr,
                           sonar_align_result* out);
// speech.go:370-408 temporal block (temporal.cu): scalars [2..6] of the stream's scalar area receive silence ratio,
// peak / average amplitude, onset density and the onset count; attack times go to feat + o_att
int launch_temporal(const double* pcm, int64_t n, int64_t stride, int n_streams, double alpha, double* feat,
                    int64_t feat_stride, int64_t o_energy, int64_t o_scalars, int64_t o_att, int64_t Te, double* tmp,
                    int64_t tmp_stride, int64_t o_part, int call_sr, int algo_sr, int energy_hop, cudaStream_t st);
int temporal_partials_doubles();
}  // namespace sonar
