// C-ABI entry points of the fingerprint path: kernel sequencing, H2D/D2H staging,
// multi-device sharding of stream batches.  The arithmetic lives in
// stft_features.cu / timedomain.cu / yin.cu.
//
//   GenerateFingerprint            fingerprint/fingerprint.go:137-236
//   SpeechFeatureExtractor         fingerprint/extractors/speech.go:135-243
//   ComputeSTFTWithWindow          fingerprint/analyzers/spectral.go:385-545
//   ComputeSTFTBatch (batch form)  fingerprint/analyzers/spectral.go:234-285
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sstream>
#include <thread>

#include "common.h"

namespace sonar {

namespace {

int64_t go_int(double x) {
  if (!(x > -9.2e18 && x < 9.2e18)) return INT64_MIN;
  return (int64_t)x;
}

std::string plan_key(int device, const sonar_fp_params* p) {
  std::ostringstream os;
  os.precision(17);
  os << device << '/' << p->window_size << '/' << p->hop_size << '/' << p->window_type << '/'
     << p->algo_sample_rate << '/' << p->n_mfcc << '/' << p->n_mel << '/' << p->use_liftering << '/' << p->low_hz
     << '/' << p->high_hz << '/' << p->lifter;
  return os.str();
}

int get_plan(sonar_ctx* ctx, int device, const sonar_fp_params* p, std::shared_ptr<FpPlan>* out) {
  std::lock_guard<std::mutex> lk(ctx->mu);
  const std::string key = plan_key(device, p);
  auto it = ctx->plans.find(key);
  if (it != ctx->plans.end()) {
    *out = it->second;
    return SONAR_OK;
  }
  std::shared_ptr<FpPlan> plan;
  int rc = build_fp_plan(p, &plan);  // allocates on the current device
  if (rc) return rc;
  ctx->plans[key] = plan;
  *out = plan;
  return SONAR_OK;
}

}  // namespace

int fp_validate(const sonar_fp_params* p) {
  if (p->call_sample_rate <= 0) return set_error(SONAR_ERR_INVALID, "sample rate must be positive");  // speech.go:143
  if (!stft_supported(p->window_size) && !stft_exact_only(p->window_size))
    return set_error(SONAR_ERR_UNSUPPORTED,
                     "window size must be 256, 512, 1024, 2048 (fused kernels) or any length in [8, 2048] (float64 route)");
  return SONAR_OK;
}

int fp_shape(const sonar_fp_params* p, int64_t n, FpShape* s) {
  int rc = host_fp_sizes(p, n, &s->sz);
  if (rc) return rc;
  rc = host_fp_layout(p, n, &s->L);
  if (rc) return rc;
  s->lr_win = s->lr_hop = s->lr_nw = 0;
  if (p->algo_sample_rate > 0) {
    s->lr_win = go_int(0.4 * (double)p->algo_sample_rate);
    s->lr_hop = s->lr_win / 4;
    if (s->lr_hop <= 0) s->lr_hop = 1;
    s->lr_nw = (n < s->lr_win || s->lr_win <= 0) ? 0 : (n - s->lr_win) / s->lr_hop + 1;
  }
  s->tmp_doubles_per_stream = (size_t)(2 * s->sz.n_pitch_frames + s->lr_nw + 2);
  s->o_wpart = (int64_t)s->tmp_doubles_per_stream;
  s->wpart_doubles = 2 * ((s->sz.n_frames + 7) / 8) + 2;
  s->tmp_doubles_per_stream += (size_t)s->wpart_doubles;
  s->o_ylist = (int64_t)s->tmp_doubles_per_stream;
  s->tmp_doubles_per_stream += (size_t)((s->sz.n_pitch_frames + 2) / 2 + 1);
  s->o_work = (int64_t)s->tmp_doubles_per_stream;
  s->tmp_doubles_per_stream += 2;
  s->o_xlist = (int64_t)s->tmp_doubles_per_stream;  // 1 + T ints: frames re-evaluated in float64 (spectral_exact.cu)
  s->tmp_doubles_per_stream += (size_t)((s->sz.n_frames + 2) / 2 + 1);
  s->speech = (p->enable & SONAR_FP_ENABLE_SPEECH) != 0;
  if (s->speech) {  // extractSpeechFeatures (speech.go:271-317): gate + tilt behind the public layout
    s->n_speech_frames = s->sz.n_pitch_frames;  // (N-1024)/512+1, clamped at 0 (speech.go:530-532, 552-554)
    s->o_sgate = s->L.total;
    s->o_tilt = s->o_sgate + 4;
    s->L.total = (s->o_tilt + s->n_speech_frames + 1) & ~(int64_t)1;
    s->o_spart = (int64_t)s->tmp_doubles_per_stream;
    s->tmp_doubles_per_stream += speech_gate_scratch_doubles();
  }
  s->temporal = (p->enable & SONAR_FP_ENABLE_TEMPORAL) != 0;
  if (s->temporal) {
    s->o_env = s->L.total;
    s->o_att = s->o_env + s->sz.n_envelope;
    s->L.total = (s->o_att + s->sz.n_energy_frames + 1) & ~(int64_t)1;
    s->o_part = (int64_t)s->tmp_doubles_per_stream;
    s->tmp_doubles_per_stream += (size_t)temporal_partials_doubles();
  }
  return SONAR_OK;
}

// Enqueues every kernel of one uniform batch on `st`.  pcm_dev: ns streams of n samples,
// `stride` apart; feat_dev: ns feature blocks of sh.L.total doubles; tmp_dev: ns *
// sh.tmp_doubles_per_stream doubles.
int enqueue_fingerprint(sonar_ctx* ctx, int device, const sonar_fp_params* p, const FpShape& sh,
                        const double* pcm_dev, int64_t n, int64_t stride, int ns, double* feat_dev,
                        double* tmp_dev, cudaStream_t st, cudaEvent_t energy_ready) {
  std::shared_ptr<FpPlan> plan;
  int rc = get_plan(ctx, device, p, &plan);
  if (rc) return rc;
  const auto& L = sh.L;
  const int64_t T = sh.sz.n_frames, Te = sh.sz.n_energy_frames, Tp = sh.sz.n_pitch_frames;
  const unsigned char* blob = static_cast<const unsigned char*>(plan->d_blob);
  const int64_t tstride = (int64_t)sh.tmp_doubles_per_stream;

  // The slot that owns `st` lends its side stream to the pitch tracker (one warp per stream), which then runs beside the
  // kernels that follow the YIN kernel on `st`; `st` waits for the side stream at the end.  (Tried: the loudness kernels
  // on the side stream beside the frame walk -- the two do not co-reside, either one fills the register file of an SM.)
  Slot* side = nullptr;
  for (auto& d : ctx->devs)
    if (d.device == device)
      for (auto& sl : d.slot)
        if (sl.st == st) side = &sl;
  // scalars: [0] energy variance (energy.go:97-118), [1] loudness range (energy.go:157-225), [2..] temporal block
  rc = launch_fill_strided(feat_dev + L.scalars, L.total - L.scalars, L.total, ns, 0.0, st);
  if (rc) return rc;
  // n in (W - H, W): Go's truncating division still yields one frame, which the reference's worker then skips because
  // it would end behind the signal (analyzers/spectral.go:409,472-474): its row stays zero.  The STFT kernel reads a
  // frame of silence instead of running past the buffer; same for the lone pitch frame of n in (512, 1024).
  const bool short_win = n < (int64_t)p->window_size;
  StftArgs a;
  std::memset(&a, 0, sizeof(a));
  a.pcm = short_win ? reinterpret_cast<const double*>(blob + plan->off_zero) : pcm_dev;
  a.n = short_win ? (int64_t)p->window_size : n;
  a.stride = short_win ? 0 : stride;
  a.n_streams = ns;
  a.T = T;
  a.Te = Te;
  a.hop = p->hop_size;
  a.win2 = reinterpret_cast<const float2*>(blob + plan->off_win2);
  a.tw1 = reinterpret_cast<const float2*>(blob + plan->off_tw1);
  a.wn = reinterpret_cast<const float2*>(blob + plan->off_wn);
  a.xtab = reinterpret_cast<const float*>(blob + plan->off_xtab);
  a.dct = reinterpret_cast<const float*>(blob + plan->off_dct);
  a.lift = reinterpret_cast<const float*>(blob + plan->off_lift);
  a.regions = reinterpret_cast<const MelRegion*>(blob + plan->off_regions);
  a.chunk_region = reinterpret_cast<const int*>(blob + plan->off_chunk_region);
  a.n_regions = plan->n_regions;
  a.n_mel = plan->n_mel;
  a.n_mfcc = plan->n_mfcc;
  a.split = plan->split;
  a.slope_on = plan->slope_on;
  a.mfcc_on = (p->enable & SONAR_FP_ENABLE_MFCC) ? 1 : 0;
  a.freq_scale = plan->freq_scale;
  a.slope_ntot = plan->slope_ntot;
  a.slope_xxtot = plan->slope_xxtot;
  a.feat = feat_dev;
  a.feat_stride = L.total;
  a.o_mfcc = L.mfcc;
  a.o_centroid = L.spectral_centroid;
  a.o_rolloff = L.spectral_rolloff;
  a.o_bandwidth = L.spectral_bandwidth;
  a.o_flatness = L.spectral_flatness;
  a.o_crest = L.spectral_crest;
  a.o_slope = L.spectral_slope;
  a.o_flux = L.spectral_flux;
  a.o_low = L.low_energy_ratio;
  a.o_high = L.high_energy_ratio;
  // The fused STFT kernel and the float64 re-evaluation of the frames it lists.  Enqueued AFTER the frame walk: the
  // alignment branch of the pair pipeline only waits for the walk's energies.  The STFT kernel's CTAs fill the register
  // file of an SM, so the branch's kernels mostly queue behind it and then run beside the pitch kernel, whose CTAs leave
  // room.  (Measured, 32 pairs: STFT before the walk 30.2 ms, after the pitch kernel 27.6 ms -- but then the pitch
  // tracker and the branch share the SMs with the STFT kernel and stretch it from 6.8 to 7.8 .. 10 ms.)
  auto run_stft = [&]() -> int {
    // only the third-generation kernel lists frames (the other geometries keep their stated FP32 bounds)
    a.xlist = stft_v3_eligible(*plan, a) ? reinterpret_cast<int*>(tmp_dev + sh.o_xlist) : nullptr;
    a.xlist_stride = 2 * tstride;
    a.win64 = reinterpret_cast<const double*>(blob + plan->off_win64);
    a.fac64 = reinterpret_cast<const double2*>(blob + plan->off_fac64);
    a.melbins = reinterpret_cast<const int*>(blob + plan->off_melbins);
    a.dct64 = reinterpret_cast<const double*>(blob + plan->off_dct64);
    a.lift64 = reinterpret_cast<const double*>(blob + plan->off_lift64);
    a.mel_invw = reinterpret_cast<const float*>(blob + plan->off_melinvw);
    a.algo_sr = p->algo_sample_rate;
    a.N = p->window_size;
    a.fft_len = plan->fft_len;
    a.chirp_inv = reinterpret_cast<const double2*>(blob + plan->off_chirp);
    a.blue_fb = reinterpret_cast<const double2*>(blob + plan->off_bluefb);
    if (plan->exact_only) {  // no fused kernel for this window length: every frame in float64 (Bluestein if not 2^k)
      a.exact_all = 1;
      a.xlist = nullptr;
      rc = launch_spectral_exact(a, st);
      if (rc) return rc;
      if (!a.mfcc_on) {
        rc = launch_fill_strided(feat_dev + L.mfcc, T * sh.sz.n_mfcc, L.total, ns, 0.0, st);
        if (rc) return rc;
      }
      return SONAR_OK;
    }
    a.work_counter = reinterpret_cast<unsigned*>(tmp_dev + sh.o_work);
    rc = launch_fill_strided(tmp_dev + sh.o_work, 1, tstride, 1, 0.0, st);
    if (rc) return rc;
    rc = launch_fill_strided(tmp_dev + sh.o_xlist, 1, tstride, ns, 0.0, st);  // empty lists
    if (rc) return rc;
    rc = launch_stft_features(*plan, a, false, st);
    if (rc) return rc;
    rc = launch_spectral_exact(a, st);  // the frames the fused kernel listed, in float64 and the reference's order
    if (rc) return rc;
    if (!a.mfcc_on) {
      rc = launch_fill_strided(feat_dev + L.mfcc, T * sh.sz.n_mfcc, L.total, ns, 0.0, st);
      if (rc) return rc;
    }
    return SONAR_OK;
  };
  // order on `st`: frame walk -> [energies ready: the alignment branch may start] -> STFT (+ float64 re-evaluation) ->
  // pitch detector -> small kernels.  SONAR_FP_ORDER=stft_first / yin_first select the other two orders (diagnostic).
  static const char* order_env = std::getenv("SONAR_FP_ORDER");
  static const int order = !order_env ? 0 : (std::string(order_env) == "stft_first" ? 1 : (std::string(order_env) == "yin_first" ? 2 : 0));
  if (order == 1 && (rc = run_stft())) return rc;

  // exact FP64 walks over the pre-emphasised PCM: short-time energy (+entropy) and ZCR
  const bool same_grid = !short_win && (Te == T) && p->energy_frame == p->window_size && p->energy_hop == p->hop_size;
  bool loudness_done = false;  // the walk can leave the loudness windows' RMS behind (one pass over the PCM less)
  if (same_grid) {
    WalkLoudness wl{sh.lr_win, sh.lr_hop, sh.lr_nw, tmp_dev + sh.o_wpart, tstride, tmp_dev + 2 * Tp, tstride};
    rc = launch_frame_walk(pcm_dev, n, stride, ns, p->pre_emph_alpha, p->window_size, p->hop_size, T,
                           p->algo_sample_rate, feat_dev, L.total, L.short_time_energy, L.energy_entropy,
                           L.zero_crossing_rate, st, sh.lr_nw > 0 ? &wl : nullptr, &loudness_done);
    if (rc) return rc;
  } else {
    if (Te > 0) {
      rc = launch_frame_walk(pcm_dev, n, stride, ns, p->pre_emph_alpha, p->energy_frame, p->energy_hop, Te,
                             p->algo_sample_rate, feat_dev, L.total, L.short_time_energy, L.energy_entropy, -1,
                             st);
      if (rc) return rc;
    }
    if (short_win)  // the frame handed to the ZCR is pre[0 : min(W, N)] = the whole short stream (speech.go:351-357)
      rc = launch_short_zcr(pcm_dev, n, stride, ns, p->pre_emph_alpha, p->algo_sample_rate, feat_dev, L.total,
                            L.zero_crossing_rate, st);
    else
      rc = launch_frame_walk(pcm_dev, n, stride, ns, p->pre_emph_alpha, p->window_size, p->hop_size, T,
                             p->algo_sample_rate, feat_dev, L.total, -1, -1, L.zero_crossing_rate, st);
    if (rc) return rc;
  }
  // the alignment branch of the pair pipeline only needs the short-time energies: it may start on its own stream
  // here, next to the rest of the fingerprint (loudness range, YIN)
  // The alignment branch starts right behind the walk; its first kernels (z-score, FFT screen) then share the SMs with
  // the STFT kernel, which costs that kernel ~10 % (7.3 -> 8.2 ms) but the step 2.8 ms less (24.8 vs 27.6 ms, 32 pairs)
  // than starting the branch behind the STFT kernel (SONAR_ALIGN_LATE, diagnostic)
  static const bool align_early = std::getenv("SONAR_ALIGN_LATE") == nullptr;
  if (energy_ready && (align_early || order != 0)) SONAR_CUDA(cudaEventRecord(energy_ready, st));
  if (order == 0 && (rc = run_stft())) return rc;
  if (energy_ready && !align_early && order == 0) SONAR_CUDA(cudaEventRecord(energy_ready, st));
  if (Te > T) {  // band ratios only exist where a magnitude frame does (speech.go:436-456)
    rc = launch_fill_strided(feat_dev + L.low_energy_ratio + T, Te - T, L.total, ns, 0.0, st);
    if (rc) return rc;
    rc = launch_fill_strided(feat_dev + L.high_energy_ratio + T, Te - T, L.total, ns, 0.0, st);
    if (rc) return rc;
  }

  // harmonic block (speech.go:464-509).  Enqueued before the variance / temporal kernels so that its tracker, a
  // sequential walk with one warp per stream, runs on the side stream beside them.
  if (sh.speech) {  // before the pitch tracker: its first pass depends on the gate
    rc = launch_speech(pcm_dev, n, stride, ns, p->pre_emph_alpha, p->algo_sample_rate, sh.n_speech_frames, feat_dev, L.total,
                       sh.o_sgate, sh.o_tilt, tmp_dev + sh.o_spart, tstride, st);
    if (rc) return rc;
  }
  bool forked = false;
  {
    const double* hann = reinterpret_cast<const double*>(blob + plan->off_hann);
    // n in (512, 1024): one frame shorter than the detector's window -> DetectPitch errors, the zeros stay
    // (speech.go:480-487, pitch_detection.go:226-228) = what the sample-rate-0 path writes
    rc = launch_yin(pcm_dev, stride, ns, p->pre_emph_alpha, n < 1024 ? 0 : p->algo_sample_rate, Tp, hann, feat_dev, L.total,
                    L.pitch_estimate, L.pitch_confidence, L.voicing_strength, L.harmonic_ratio, L.inharmonicity_ratio,
                    L.tonal_centroid, tmp_dev, tstride, st, side ? side->st3 : nullptr, side ? side->fork : nullptr,
                    side ? side->join : nullptr, &forked, reinterpret_cast<int*>(tmp_dev + sh.o_ylist), 2 * tstride,
                    sh.speech ? feat_dev + sh.o_sgate : nullptr, L.total);
    if (rc) return rc;
  }
  if (order == 2 && (rc = run_stft())) return rc;

  if (Te >= 2) {
    rc = launch_variance(feat_dev + L.short_time_energy, Te, L.total, ns, feat_dev + L.scalars, L.total, st);
    if (rc) return rc;
  }
  if (sh.lr_nw > 0) {
    double* rms = tmp_dev + 2 * Tp;
    if (!loudness_done) {
      rc = launch_rms_windows(pcm_dev, n, stride, ns, p->pre_emph_alpha, (int)sh.lr_win, (int)sh.lr_hop, sh.lr_nw, rms,
                              tstride, st);
      if (rc) return rc;
    }
    rc = launch_loudness_range(rms, sh.lr_nw, tstride, ns, feat_dev + L.scalars + 1, L.total, st);
    if (rc) return rc;
  }

  // temporal block (speech.go:370-408): envelope = 512/256 RMS of the pre-emphasised PCM, the rest in temporal.cu
  if (sh.temporal) {
    if (sh.sz.n_envelope > 0) {
      rc = launch_rms_windows(pcm_dev, n, stride, ns, p->pre_emph_alpha, 512, 256, sh.sz.n_envelope, feat_dev + sh.o_env,
                              L.total, st);
      if (rc) return rc;
    }
    rc = launch_temporal(pcm_dev, n, stride, ns, p->pre_emph_alpha, feat_dev, L.total, L.short_time_energy, L.scalars,
                         sh.o_att, Te, tmp_dev, tstride, sh.o_part, p->call_sample_rate, p->algo_sample_rate,
                         p->energy_hop, st);
    if (rc) return rc;
  }

  if (side && forked) SONAR_CUDA(cudaStreamWaitEvent(st, side->join, 0));
  ctx->last_lists.device = device;
  ctx->last_lists.st = st;
  ctx->last_lists.tmp = tmp_dev;
  ctx->last_lists.tstride = tstride;
  ctx->last_lists.o_xlist = sh.o_xlist;
  ctx->last_lists.o_ylist = sh.o_ylist;
  ctx->last_lists.ns = ns;
  return SONAR_OK;
}

void scatter_block(const double* f, const FpShape& sh, sonar_fp_out* o) {
  const auto& L = sh.L;
  const int64_t T = sh.sz.n_frames, Te = sh.sz.n_energy_frames, Tp = sh.sz.n_pitch_frames;
  auto cp = [&](double* dst, int64_t off, int64_t cnt) {
    if (dst && cnt > 0) std::memcpy(dst, f + off, sizeof(double) * (size_t)cnt);
  };
  cp(o->mfcc, L.mfcc, T * sh.sz.n_mfcc);
  cp(o->spectral_centroid, L.spectral_centroid, T);
  cp(o->spectral_rolloff, L.spectral_rolloff, T);
  cp(o->spectral_bandwidth, L.spectral_bandwidth, T);
  cp(o->spectral_flatness, L.spectral_flatness, T);
  cp(o->spectral_crest, L.spectral_crest, T);
  cp(o->spectral_slope, L.spectral_slope, T);
  cp(o->spectral_flux, L.spectral_flux, sh.sz.n_flux);
  cp(o->zero_crossing_rate, L.zero_crossing_rate, T);
  cp(o->short_time_energy, L.short_time_energy, Te);
  cp(o->energy_entropy, L.energy_entropy, Te);
  cp(o->low_energy_ratio, L.low_energy_ratio, Te);
  cp(o->high_energy_ratio, L.high_energy_ratio, Te);
  cp(o->pitch_estimate, L.pitch_estimate, Tp);
  cp(o->pitch_confidence, L.pitch_confidence, Tp);
  cp(o->voicing_strength, L.voicing_strength, Tp);
  cp(o->harmonic_ratio, L.harmonic_ratio, Tp);
  cp(o->inharmonicity_ratio, L.inharmonicity_ratio, Tp);
  cp(o->tonal_centroid, L.tonal_centroid, Tp);
  o->energy_variance = f[L.scalars];
  o->loudness_range = f[L.scalars + 1];
  o->dynamic_range = 0;
  o->silence_ratio = 0;
  o->peak_amplitude = 0;
  o->average_amplitude = 0;
  o->onset_density = 0;
  o->n_attack_time = 0;
  if (sh.temporal) {  // extractTemporalFeatures (speech.go:370-408)
    cp(o->rms_energy, L.short_time_energy, Te);  // RMSEnergy is the same ComputeShortTimeEnergy call
    cp(o->envelope_shape, sh.o_env, sh.sz.n_envelope);
    o->dynamic_range = o->loudness_range;  // ComputeLoudnessRange again
    o->silence_ratio = f[L.scalars + 2];
    o->peak_amplitude = f[L.scalars + 3];
    o->average_amplitude = f[L.scalars + 4];
    o->onset_density = f[L.scalars + 5];
    o->n_attack_time = (int64_t)f[L.scalars + 6];
    cp(o->attack_time, sh.o_att, std::min<int64_t>(o->n_attack_time, o->attack_time_cap));
  }
}

// SpeechFeatures of one stream from its feature block (extractSpeechFeatures, speech.go:271-317): the voicing sweep is the
// detector's voicing on the same frames as the harmonic block, the pauses and the speech rate are O(Te) scans of the
// short-time energies (feature-sized, host side: speech.go:586-655,779-797).
void scatter_speech(const double* f, const FpShape& sh, const sonar_fp_params* p, int64_t n, sonar_speech_out* o) {
  const auto& L = sh.L;
  o->is_speech = f[sh.o_sgate] != 0.0 ? 1 : 0;
  o->reserved = 0;
  o->n_frames = 0;
  o->n_pause = 0;
  o->speech_rate = 0.0;
  if (!o->is_speech) return;  // speech.go:281-291: empty arrays, rate 0
  const int64_t nf = sh.n_speech_frames, Te = sh.sz.n_energy_frames;
  o->n_frames = nf;
  if (o->voicing_probability && nf > 0) std::memcpy(o->voicing_probability, f + L.voicing_strength, sizeof(double) * (size_t)nf);
  if (o->spectral_tilt && nf > 0) std::memcpy(o->spectral_tilt, f + sh.o_tilt, sizeof(double) * (size_t)nf);
  const double* e = f + L.short_time_energy;
  double silence = 0.0, thr = 0.0;
  if (Te > 0) {  // sortedEnergies[len / 10] (the reference bubble-sorts a copy; the order statistic is the same value)
    std::vector<double> srt(e, e + Te);
    std::nth_element(srt.begin(), srt.begin() + Te / 10, srt.end());
    thr = srt[(size_t)(Te / 10)];
    int64_t silent = 0;
    for (int64_t i = 0; i < Te; i++)
      if (e[i] <= thr) silent++;
    silence = (double)silent / (double)Te;
  }
  const double dur = (double)n / (double)p->algo_sample_rate;  // estimateSpeechRate :779-797
  const double speech_time = dur * (1.0 - silence);
  o->speech_rate = speech_time > 0 ? 4.0 * speech_time / dur : 3.0;
  if (Te > 0) {  // extractPauseDurations :586-641
    const double frame_time = (double)p->energy_hop / (double)p->algo_sample_rate;
    bool in_pause = false;
    int64_t start = 0, np = 0;
    auto emit = [&](int64_t end) {
      const double d = (double)(end - start) * frame_time;
      if (d > 0.1) {
        if (o->pause_duration && np < o->pause_cap) o->pause_duration[np] = d;
        np++;
      }
    };
    for (int64_t i = 0; i < Te; i++) {
      if (e[i] <= thr) {
        if (!in_pause) {
          in_pause = true;
          start = i;
        }
      } else if (in_pause) {
        emit(i);
        in_pause = false;
      }
    }
    if (in_pause) emit(Te);
    o->n_pause = np;
  }
}

namespace {

struct Chunk {  // consecutive streams of one device with identical length
  std::vector<int> ids;
  int64_t n = 0;
  FpShape sh;
};

struct DeviceJob {
  int rc = SONAR_OK;
  std::string err;
};

// One device's share of a host batch: chunks rotate through the device's slots (stream + buffers each), so
// the H2D copy of chunk k+1 and k+2 overlaps the kernels and the D2H of chunk k, and the host-side scatter of
// a finished chunk into the caller's arrays overlaps the GPU work queued on the other slots.
void run_device_batch(sonar_ctx* ctx, DevCtx* dev, const double* const* pcm, const std::vector<Chunk>* chunks,
                      const sonar_fp_params* p, int fmt, sonar_fp_out* outs, DeviceJob* job, sonar_speech_out* speech) {
  set_current_ctx(ctx);
  auto fail = [&](int rc) {
    job->rc = rc;
    job->err = sonar_last_error();
  };
  cudaError_t e = cudaSetDevice(dev->device);
  if (e != cudaSuccess) return fail(cuda_error(e, "cudaSetDevice"));
  constexpr int NS = DevCtx::kStageSlots;
  const Chunk* pending[NS] = {};
  auto finish = [&](int si) -> int {
    const Chunk* c = pending[si];
    if (!c) return SONAR_OK;
    Slot& s = dev->slot[si];
    SONAR_CUDA(cudaEventSynchronize(s.done));
    const double* h = static_cast<const double*>(s.h_out.p);
    for (size_t i = 0; i < c->ids.size(); i++) {
      scatter_block(h + (int64_t)i * c->sh.L.total, c->sh, &outs[c->ids[i]]);
      if (speech && c->sh.speech) scatter_speech(h + (int64_t)i * c->sh.L.total, c->sh, p, c->n, &speech[c->ids[i]]);
    }
    pending[si] = nullptr;
    return SONAR_OK;
  };
  int k = 0;
  for (const Chunk& c : *chunks) {
    const int si = k++ % NS;
    Slot& s = dev->slot[si];
    int rc = finish(si);
    if (rc) return fail(rc);
    const int ns = (int)c.ids.size();
    const int64_t stride = (c.n + 1) & ~(int64_t)1;
    const size_t out_bytes = sizeof(double) * (size_t)c.sh.L.total * ns;
    if ((rc = dev->ensure_dev(s.d_in, sizeof(double) * (size_t)stride * ns)) ||
        (rc = dev->ensure_dev(s.d_out, out_bytes)) ||
        (rc = dev->ensure_dev(s.d_tmp, sizeof(double) * c.sh.tmp_doubles_per_stream * ns)) ||
        (rc = dev->ensure_host(s.h_out, out_bytes)))
      return fail(rc);
    double* d_in = static_cast<double*>(s.d_in.p);
    if (fmt == SONAR_PCM_F64) {
      for (int i = 0; i < ns; i++) {
        e = cudaMemcpyAsync(d_in + (int64_t)i * stride, pcm[c.ids[i]], sizeof(double) * (size_t)c.n,
                            cudaMemcpyHostToDevice, s.st);
        if (e != cudaSuccess) return fail(cuda_error(e, "cudaMemcpyAsync(H2D pcm)"));
      }
    } else {  // narrow samples cross PCIe and are widened on the device (sonar_fingerprint_batch_pcm)
      const size_t sb = pcm_sample_bytes(fmt);
      const int64_t raw_stride = (c.n + 7) & ~(int64_t)7;
      if ((rc = dev->ensure_dev(s.d_raw, sb * (size_t)raw_stride * ns))) return fail(rc);
      unsigned char* d_raw = static_cast<unsigned char*>(s.d_raw.p);
      for (int i = 0; i < ns; i++) {
        e = cudaMemcpyAsync(d_raw + sb * (size_t)i * raw_stride, pcm[c.ids[i]], sb * (size_t)c.n, cudaMemcpyHostToDevice,
                            s.st);
        if (e != cudaSuccess) return fail(cuda_error(e, "cudaMemcpyAsync(H2D pcm)"));
      }
      if ((rc = launch_widen_pcm(d_raw, fmt, d_in, c.n, raw_stride, stride, ns, s.st))) return fail(rc);
    }
    rc = enqueue_fingerprint(ctx, dev->device, p, c.sh, d_in, c.n, stride, ns, static_cast<double*>(s.d_out.p),
                             static_cast<double*>(s.d_tmp.p), s.st);
    if (rc) return fail(rc);
    e = cudaMemcpyAsync(s.h_out.p, s.d_out.p, out_bytes, cudaMemcpyDeviceToHost, s.st);
    if (e != cudaSuccess) return fail(cuda_error(e, "cudaMemcpyAsync(D2H features)"));
    e = cudaEventRecord(s.done, s.st);
    if (e != cudaSuccess) return fail(cuda_error(e, "cudaEventRecord"));
    pending[si] = &c;
  }
  for (int j = 0; j < NS; j++) {
    int rc = finish((k + j) % NS);  // oldest first
    if (rc) return fail(rc);
  }
}

constexpr size_t kChunkBytes = (size_t)256 << 20;  // PCM bytes per in-flight chunk and slot

}  // namespace
}  // namespace sonar

using namespace sonar;

extern "C" {

static int fingerprint_batch(sonar_ctx* ctx, const double* const* pcm, int fmt, const int64_t* n, int n_streams,
                             const sonar_fp_params* p, sonar_fp_out* outs, sonar_speech_out* speech = nullptr);

int sonar_fingerprint_batch_f64(sonar_ctx* ctx, const double* const* pcm, const int64_t* n, int n_streams,
                                const sonar_fp_params* p, sonar_fp_out* outs) {
  return fingerprint_batch(ctx, pcm, SONAR_PCM_F64, n, n_streams, p, outs);
}

int sonar_fingerprint_batch_pcm(sonar_ctx* ctx, const void* const* pcm, int sample_format, const int64_t* n,
                                int n_streams, const sonar_fp_params* p, sonar_fp_out* outs) {
  if (sample_format != SONAR_PCM_F64 && sample_format != SONAR_PCM_F32 && sample_format != SONAR_PCM_S16)
    return set_error(SONAR_ERR_INVALID, "unknown PCM sample format");
  return fingerprint_batch(ctx, reinterpret_cast<const double* const*>(pcm), sample_format, n, n_streams, p, outs);
}

static int fingerprint_batch(sonar_ctx* ctx, const double* const* pcm, int fmt, const int64_t* n, int n_streams,
                             const sonar_fp_params* p, sonar_fp_out* outs, sonar_speech_out* speech) {
  if (!ctx || !p || (n_streams > 0 && (!pcm || !n || !outs)))
    return set_error(SONAR_ERR_INVALID, "audio data cannot be nil");  // fingerprint.go:139
  if (n_streams <= 0) return SONAR_OK;
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  int rc = fp_validate(p);
  if (rc) return rc;
  const int nd = (int)ctx->devs.size();
  std::vector<std::vector<Chunk>> per_dev(nd);
  for (int s = 0; s < n_streams; s++) {
    if (!pcm[s]) return set_error(SONAR_ERR_INVALID, "audio data cannot be nil");
    auto& chunks = per_dev[s % nd];
    const bool extend = !chunks.empty() && chunks.back().n == n[s] &&
                        (chunks.back().ids.size() + 1) * (size_t)n[s] * sizeof(double) <= kChunkBytes;
    if (!extend) {
      Chunk c;
      c.n = n[s];
      rc = fp_shape(p, n[s], &c.sh);
      if (rc) return rc;
      chunks.push_back(std::move(c));
    }
    chunks.back().ids.push_back(s);
  }
  std::vector<DeviceJob> jobs(nd);
  if (nd == 1) {
    run_device_batch(ctx, &ctx->devs[0], pcm, &per_dev[0], p, fmt, outs, &jobs[0], speech);
  } else {
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++)
      th.emplace_back(run_device_batch, ctx, &ctx->devs[d], pcm, &per_dev[d], p, fmt, outs, &jobs[d], speech);
    for (auto& t : th) t.join();
    cudaSetDevice(ctx->devs[0].device);
  }
  for (auto& j : jobs)
    if (j.rc) return set_error(j.rc, j.err);
  return SONAR_OK;
}

int sonar_fp_exact_counts(sonar_ctx* ctx, int64_t* spectral, int64_t* pitch) {
  if (!ctx) return set_error(SONAR_ERR_INVALID, "nil argument");
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  const auto& ll = ctx->last_lists;
  if (spectral) *spectral = 0;
  if (pitch) *pitch = 0;
  if (ll.device < 0 || !ll.tmp || ll.ns <= 0) return SONAR_OK;
  SONAR_CUDA(cudaSetDevice(ll.device));
  SONAR_CUDA(cudaStreamSynchronize(ll.st));
  std::vector<int> cx((size_t)ll.ns), cy((size_t)ll.ns);
  SONAR_CUDA(cudaMemcpy2D(cx.data(), sizeof(int), ll.tmp + ll.o_xlist, sizeof(double) * (size_t)ll.tstride, sizeof(int),
                          (size_t)ll.ns, cudaMemcpyDeviceToHost));
  SONAR_CUDA(cudaMemcpy2D(cy.data(), sizeof(int), ll.tmp + ll.o_ylist, sizeof(double) * (size_t)ll.tstride, sizeof(int),
                          (size_t)ll.ns, cudaMemcpyDeviceToHost));
  for (int i = 0; i < ll.ns; i++) {
    if (spectral) *spectral += cx[i];
    if (pitch) *pitch += cy[i];
  }
  return SONAR_OK;
}

int sonar_fingerprint_f64(sonar_ctx* ctx, const double* pcm, int64_t n, const sonar_fp_params* p,
                          sonar_fp_out* out) {
  if (!ctx || !pcm || !p || !out) return set_error(SONAR_ERR_INVALID, "audio data cannot be nil");
  const double* ptrs[1] = {pcm};
  return sonar_fingerprint_batch_f64(ctx, ptrs, &n, 1, p, out);
}

int sonar_fingerprint_speech_f64(sonar_ctx* ctx, const double* pcm, int64_t n, const sonar_fp_params* p, sonar_fp_out* out,
                                 sonar_speech_out* speech) {
  if (!ctx || !pcm || !p || !out || !speech) return set_error(SONAR_ERR_INVALID, "audio data cannot be nil");
  sonar_fp_params q = *p;
  q.enable |= SONAR_FP_ENABLE_SPEECH;
  const double* ptrs[1] = {pcm};
  return fingerprint_batch(ctx, ptrs, SONAR_PCM_F64, &n, 1, &q, out, speech);
}

int sonar_fingerprint_batch_dev(sonar_ctx* ctx, const double* pcm_dev, int64_t n, int64_t stride, int n_streams,
                                const sonar_fp_params* p, double* feat_dev) {
  if (!ctx || !pcm_dev || !p || !feat_dev) return set_error(SONAR_ERR_INVALID, "audio data cannot be nil");
  if (n_streams <= 0) return SONAR_OK;
  if (stride < n) return set_error(SONAR_ERR_INVALID, "stride must be >= n");
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  int rc = fp_validate(p);
  if (rc) return rc;
  if (p->enable & (SONAR_FP_ENABLE_TEMPORAL | SONAR_FP_ENABLE_SPEECH))
    return set_error(SONAR_ERR_UNSUPPORTED, "temporal / speech features are not part of the device layout");
  FpShape sh;
  rc = fp_shape(p, n, &sh);
  if (rc) return rc;
  DevCtx& dev = ctx->devs[0];
  SONAR_CUDA(cudaSetDevice(dev.device));
  Slot& s = dev.slot[0];
  rc = dev.ensure_dev(s.d_tmp, sizeof(double) * sh.tmp_doubles_per_stream * (size_t)n_streams);
  if (rc) return rc;
  return enqueue_fingerprint(ctx, dev.device, p, sh, pcm_dev, n, stride, n_streams, feat_dev,
                             static_cast<double*>(s.d_tmp.p), s.st);
}

// Materialising STFT of the signal [carry | pcm]: `carry_n` samples already resident on the device (the streamer's
// buffer, `d_carry`) followed by n host samples.  On success the last `keep` samples of the signal are left in `d_carry`
// (device to device): the streaming form never uploads a sample twice.
static int stft_core(sonar_ctx* ctx, const double* pcm, int64_t n_host, double* d_carry, int64_t carry_n, int64_t keep,
                     int win, int hop, int window_type, double* mag, double* phase, double* cplx) {
  const int64_t n = n_host + carry_n;
  if (!ctx) return set_error(SONAR_ERR_INVALID, "nil argument");
  if (!pcm || n <= 0) return set_error(SONAR_ERR_EMPTY, "empty signal");  // analyzers/spectral.go:387
  if (win <= 0) return set_error(SONAR_ERR_INVALID, "window size must be positive");
  if (hop <= 0) return set_error(SONAR_ERR_INVALID, "hop size must be positive");
  const int64_t T = (n - win) / hop + 1;
  if (T <= 0) return set_error(SONAR_ERR_TOO_SHORT, "signal too short for given window size and hop size");
  if (!mag) return set_error(SONAR_ERR_INVALID, "nil argument");
  if (!stft_supported(win) && !stft_exact_only(win))
    return set_error(SONAR_ERR_UNSUPPORTED,
                     "window size must be 256, 512, 1024, 2048 (fused kernels) or any length in [8, 2048] (float64 route)");
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  sonar_fp_params p;
  sonar_fp_params_default(&p);
  p.window_size = win;
  p.hop_size = hop;
  p.window_type = window_type;
  DevCtx& dev = ctx->devs[0];
  SONAR_CUDA(cudaSetDevice(dev.device));
  std::shared_ptr<FpPlan> plan;
  int rc = get_plan(ctx, dev.device, &p, &plan);
  if (rc) return rc;
  Slot& s = dev.slot[0];
  const int B = win / 2 + 1;
  const size_t per = sizeof(double) * (size_t)T * B;
  if (n < win) {  // T == 1, the frame is skipped (analyzers/spectral.go:472-474): the rows stay as allocated
    std::memset(mag, 0, per);
    if (phase) std::memset(phase, 0, per);
    if (cplx) std::memset(cplx, 0, 2 * per);
    return SONAR_OK;
  }
  const size_t out_bytes = per * (1 + (phase ? 1 : 0) + (cplx ? 2 : 0));
  const int64_t stride = (n + 1) & ~(int64_t)1;
  if ((rc = dev.ensure_dev(s.d_in, sizeof(double) * (size_t)stride)) || (rc = dev.ensure_dev(s.d_out, out_bytes)))
    return rc;
  if (carry_n > 0)
    SONAR_CUDA(cudaMemcpyAsync(s.d_in.p, d_carry, sizeof(double) * (size_t)carry_n, cudaMemcpyDeviceToDevice, s.st));
  SONAR_CUDA(cudaMemcpyAsync(static_cast<double*>(s.d_in.p) + carry_n, pcm, sizeof(double) * (size_t)n_host,
                             cudaMemcpyHostToDevice, s.st));
  const unsigned char* blob = static_cast<const unsigned char*>(plan->d_blob);
  StftArgs a;
  std::memset(&a, 0, sizeof(a));
  a.pcm = static_cast<const double*>(s.d_in.p);
  a.n = n;
  a.stride = stride;
  a.n_streams = 1;
  a.T = T;
  a.hop = hop;
  a.win2 = reinterpret_cast<const float2*>(blob + plan->off_win2);
  a.tw1 = reinterpret_cast<const float2*>(blob + plan->off_tw1);
  a.wn = reinterpret_cast<const float2*>(blob + plan->off_wn);
  double* d = static_cast<double*>(s.d_out.p);
  a.mag = d;
  d += (size_t)T * B;
  if (phase) {
    a.phase = d;
    d += (size_t)T * B;
  }
  if (cplx) a.cplx = d;
  if (plan->exact_only) {  // go-dsp's route for this length (Bluestein when it is not a power of two), float64
    a.exact_all = 1;
    a.N = win;
    a.fft_len = plan->fft_len;
    a.win64 = reinterpret_cast<const double*>(blob + plan->off_win64);
    a.fac64 = reinterpret_cast<const double2*>(blob + plan->off_fac64);
    a.chirp_inv = reinterpret_cast<const double2*>(blob + plan->off_chirp);
    a.blue_fb = reinterpret_cast<const double2*>(blob + plan->off_bluefb);
    rc = launch_spectral_exact(a, s.st);
  } else {
    rc = launch_stft_features(*plan, a, true, s.st);
  }
  if (rc) return rc;
  if (d_carry && keep > 0)  // the streamer's next buffer: the tail of this signal (stream order: after the kernel)
    SONAR_CUDA(cudaMemcpyAsync(d_carry, static_cast<const double*>(s.d_in.p) + (n - keep), sizeof(double) * (size_t)keep,
                               cudaMemcpyDeviceToDevice, s.st));
  SONAR_CUDA(cudaMemcpyAsync(mag, a.mag, per, cudaMemcpyDeviceToHost, s.st));
  if (phase) SONAR_CUDA(cudaMemcpyAsync(phase, a.phase, per, cudaMemcpyDeviceToHost, s.st));
  if (cplx) SONAR_CUDA(cudaMemcpyAsync(cplx, a.cplx, 2 * per, cudaMemcpyDeviceToHost, s.st));
  SONAR_CUDA(cudaStreamSynchronize(s.st));
  return SONAR_OK;
}

int sonar_stft_f64(sonar_ctx* ctx, const double* pcm, int64_t n, int win, int hop, int window_type, double* mag,
                   double* phase, double* cplx) {
  return stft_core(ctx, pcm, n, nullptr, 0, 0, win, hop, window_type, mag, phase, cplx);
}

int sonar_music_spectral_f64(sonar_ctx* ctx, const double* pcm, int64_t n, int win, int hop, int window_type,
                             int sample_rate, int n_bands, double* contrast, double* chroma, int n_bark,
                             double bark_low_hz, double bark_high_hz, double* bark) {
  if (!ctx) return set_error(SONAR_ERR_INVALID, "nil argument");
  if (!pcm || n <= 0) return set_error(SONAR_ERR_EMPTY, "empty signal");  // analyzers/spectral.go:387
  if (win <= 0) return set_error(SONAR_ERR_INVALID, "window size must be positive");
  if (hop <= 0) return set_error(SONAR_ERR_INVALID, "hop size must be positive");
  if (sample_rate <= 0) return set_error(SONAR_ERR_INVALID, "sample rate must be positive");
  if ((contrast && n_bands <= 0) || (bark && n_bark <= 0)) return set_error(SONAR_ERR_INVALID, "band count must be positive");
  const int64_t T = (n - win) / hop + 1;
  if (T <= 0) return set_error(SONAR_ERR_TOO_SHORT, "signal too short for given window size and hop size");
  if (!stft_supported(win))
    return set_error(SONAR_ERR_UNSUPPORTED, "window size must be 256, 512, 1024 or 2048 on the fused GPU path");
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  sonar_fp_params p;
  sonar_fp_params_default(&p);
  p.window_size = win;
  p.hop_size = hop;
  p.window_type = window_type;
  DevCtx& dev = ctx->devs[0];
  SONAR_CUDA(cudaSetDevice(dev.device));
  std::shared_ptr<FpPlan> plan;
  int rc = get_plan(ctx, dev.device, &p, &plan);
  if (rc) return rc;
  Slot& s = dev.slot[0];
  const int B = win / 2 + 1;
  // host tables, float64, as the reference's constructors build them
  std::vector<int> edges;
  std::vector<signed char> cmap;
  std::vector<double> bank;
  std::vector<int2> ranges;
  if (contrast) edges = host_contrast_edges(n_bands, B, sample_rate);
  if (chroma) cmap = host_chroma_map(B, (double)sample_rate / (double)win);
  if (bark) {
    bank = host_bark_bank(n_bark, (B - 1) * 2, sample_rate, bark_low_hz, bark_high_hz);
    ranges.resize((size_t)n_bark);
    for (int i = 0; i < n_bark; i++) {
      int lo = B, hi = 0;
      for (int k = 0; k < B; k++)
        if (bank[(size_t)i * B + k] != 0.0) {
          lo = std::min(lo, k);
          hi = k + 1;
        }
      ranges[(size_t)i] = make_int2(lo < hi ? lo : 0, lo < hi ? hi : 0);
    }
  }
  auto up256 = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const int64_t chunk = std::min<int64_t>(T, 4096);  // frames per materialised spectrogram chunk (~17 MB at B = 513)
  const size_t o_mag = 0, o_con = o_mag + up256(sizeof(double) * (size_t)chunk * B),
               o_chr = o_con + up256(sizeof(double) * (size_t)chunk * (size_t)std::max(n_bands, 1)),
               o_brk = o_chr + up256(sizeof(double) * (size_t)chunk * 12),
               o_edg = o_brk + up256(sizeof(double) * (size_t)chunk * (size_t)std::max(n_bark, 1)),
               o_map = o_edg + up256(sizeof(int) * (edges.size() + 1)), o_bank = o_map + up256(cmap.size() + 1),
               o_rng = o_bank + up256(sizeof(double) * (bank.size() + 1)),
               total = o_rng + up256(sizeof(int2) * (ranges.size() + 1));
  const int64_t stride = (n + 1) & ~(int64_t)1;
  if ((rc = dev.ensure_dev(s.d_in, sizeof(double) * (size_t)stride)) || (rc = dev.ensure_dev(s.d_out, total))) return rc;
  unsigned char* d = static_cast<unsigned char*>(s.d_out.p);
  if (n < win) {  // the lone frame is skipped by the reference's STFT worker: a spectrogram row of zeros
    if ((rc = dev.ensure_dev(s.d_in, sizeof(double) * (size_t)win))) return rc;
    SONAR_CUDA(cudaMemsetAsync(s.d_in.p, 0, sizeof(double) * (size_t)win, s.st));
    n = win;
  } else {
    SONAR_CUDA(cudaMemcpyAsync(s.d_in.p, pcm, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s.st));
  }
  if (contrast) SONAR_CUDA(cudaMemcpyAsync(d + o_edg, edges.data(), sizeof(int) * edges.size(), cudaMemcpyHostToDevice, s.st));
  if (chroma) SONAR_CUDA(cudaMemcpyAsync(d + o_map, cmap.data(), cmap.size(), cudaMemcpyHostToDevice, s.st));
  if (bark) {
    SONAR_CUDA(cudaMemcpyAsync(d + o_bank, bank.data(), sizeof(double) * bank.size(), cudaMemcpyHostToDevice, s.st));
    SONAR_CUDA(cudaMemcpyAsync(d + o_rng, ranges.data(), sizeof(int2) * ranges.size(), cudaMemcpyHostToDevice, s.st));
  }
  const unsigned char* blob = static_cast<const unsigned char*>(plan->d_blob);
  for (int64_t t0 = 0; t0 < T; t0 += chunk) {
    const int64_t tc = std::min<int64_t>(chunk, T - t0);
    StftArgs a;
    std::memset(&a, 0, sizeof(a));
    a.pcm = static_cast<const double*>(s.d_in.p) + t0 * hop;
    a.n = n - t0 * hop;
    a.stride = stride;
    a.n_streams = 1;
    a.T = tc;
    a.hop = hop;
    a.win2 = reinterpret_cast<const float2*>(blob + plan->off_win2);
    a.tw1 = reinterpret_cast<const float2*>(blob + plan->off_tw1);
    a.wn = reinterpret_cast<const float2*>(blob + plan->off_wn);
    a.mag = reinterpret_cast<double*>(d + o_mag);
    if ((rc = launch_stft_features(*plan, a, true, s.st))) return rc;
    double* dc = contrast ? reinterpret_cast<double*>(d + o_con) : nullptr;
    double* dh = chroma ? reinterpret_cast<double*>(d + o_chr) : nullptr;
    double* db = bark ? reinterpret_cast<double*>(d + o_brk) : nullptr;
    rc = launch_music_spectral(a.mag, tc, B, n_bands, reinterpret_cast<const int*>(d + o_edg), dc,
                               reinterpret_cast<const signed char*>(d + o_map), dh, n_bark,
                               reinterpret_cast<const double*>(d + o_bank), reinterpret_cast<const int2*>(d + o_rng), db,
                               s.st);
    if (rc) return rc;
    if (contrast)
      SONAR_CUDA(cudaMemcpyAsync(contrast + t0 * n_bands, dc, sizeof(double) * (size_t)tc * n_bands, cudaMemcpyDeviceToHost, s.st));
    if (chroma) SONAR_CUDA(cudaMemcpyAsync(chroma + t0 * 12, dh, sizeof(double) * (size_t)tc * 12, cudaMemcpyDeviceToHost, s.st));
    if (bark) SONAR_CUDA(cudaMemcpyAsync(bark + t0 * n_bark, db, sizeof(double) * (size_t)tc * n_bark, cudaMemcpyDeviceToHost, s.st));
    SONAR_CUDA(cudaStreamSynchronize(s.st));  // the chunk buffers are reused
  }
  return SONAR_OK;
}

}  // extern "C"

// ---- STFTStreamer (analyzers/spectral.go:289-374) -----------------------------------------------------------------
struct sonar_stft_stream {
  sonar_ctx* ctx;
  int win, hop, wtype;
  double* d_carry;   // device: the samples the reference's streamer keeps between chunks (always fewer than `win`)
  int64_t n_carry;
};

namespace {
int64_t stream_frames_for(int64_t len, int win, int hop) { return len >= win ? (len - win) / hop + 1 : 0; }
}  // namespace

int sonar_stft_stream_open(sonar_ctx* ctx, int win, int hop, int window_type, sonar_stft_stream** out) {
  if (!out) return set_error(SONAR_ERR_INVALID, "nil argument");
  *out = nullptr;
  if (win <= 0) return set_error(SONAR_ERR_INVALID, "window size must be positive");
  if (hop <= 0) return set_error(SONAR_ERR_INVALID, "hop size must be positive");
  if (!ctx) return set_error(SONAR_ERR_INVALID, "nil argument");
  if (!stft_supported(win))
    return set_error(SONAR_ERR_UNSUPPORTED, "window size must be 256, 512, 1024 or 2048 on the fused GPU path");
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  SONAR_CUDA(cudaSetDevice(ctx->devs[0].device));
  double* d = nullptr;
  SONAR_CUDA(cudaMalloc(&d, sizeof(double) * (size_t)win));
  *out = new sonar_stft_stream{ctx, win, hop, window_type, d, 0};
  return SONAR_OK;
}

int64_t sonar_stft_stream_frames(const sonar_stft_stream* s, int64_t chunk_len) {
  if (!s || chunk_len <= 0) return 0;
  return stream_frames_for(s->n_carry + chunk_len, s->win, s->hop);
}

int64_t sonar_stft_stream_buffered(const sonar_stft_stream* s) { return s ? s->n_carry : 0; }

int sonar_stft_stream_process(sonar_stft_stream* s, const double* chunk, int64_t n, double* mag, double* phase,
                              double* cplx, int64_t cap_frames, int64_t* n_frames) {
  if (!s || !n_frames) return set_error(SONAR_ERR_INVALID, "nil argument");
  *n_frames = 0;
  if (n <= 0) return SONAR_OK;  // spectral.go:324-326
  if (!chunk) return set_error(SONAR_ERR_INVALID, "nil argument");
  const int64_t len = s->n_carry + n;
  const int64_t T = stream_frames_for(len, s->win, s->hop);
  if (T > 0 && (!mag || cap_frames < T)) return set_error(SONAR_ERR_INVALID, "frame capacity too small");
  if (T == 0) {  // not a window yet (len < win): the chunk joins the device-resident buffer
    sonar_ctx* ctx = s->ctx;
    std::lock_guard<std::mutex> call_lock(ctx->call_mu);
    set_current_ctx(ctx);
    SONAR_CUDA(cudaSetDevice(ctx->devs[0].device));
    cudaStream_t st = ctx->devs[0].slot[0].st;
    SONAR_CUDA(cudaMemcpyAsync(s->d_carry + s->n_carry, chunk, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    SONAR_CUDA(cudaStreamSynchronize(st));
    s->n_carry = len;
    return SONAR_OK;
  }
  // the frames the reference's loop extracts are those of the batch transform over [buffer | chunk]: frame k starts k
  // hops in.  spectral.go:364-369: T - 1 plain advances, then the last one empties the buffer if a hop does not fit any more
  const int64_t rem = len - (T - 1) * (int64_t)s->hop;
  const int64_t keep = (int64_t)s->hop >= rem ? 0 : rem - s->hop;
  const int rc = stft_core(s->ctx, chunk, n, s->d_carry, s->n_carry, keep, s->win, s->hop, s->wtype, mag, phase, cplx);
  if (rc) return rc;  // nothing consumed: the buffer is only rewritten behind a successful transform
  s->n_carry = keep;
  *n_frames = T;
  return SONAR_OK;
}

void sonar_stft_stream_close(sonar_stft_stream* s) {
  if (!s) return;
  if (s->d_carry) cudaFree(s->d_carry);
  delete s;
}
