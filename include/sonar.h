/*
 * sonar.h — C ABI of the B200-native fingerprint + alignment hot path.
 *
 * This is the drop-in boundary for RyanBlaney/sonido-sonar's data-parallel hot
 * path.  The reference is pure Go and has no FFI today; its
 * only plugin point is the Go interface `extractors.FeatureExtractor`
 * (fingerprint/extractors/feature_extractor.go:10-15).  A cgo shim (see
 * INTEGRATION.md and go/) binds exactly the entry points declared here.  Every
 * entry point names the reference function(s) it replaces (file:line relative
 * to the reference repository root).
 *
 * Conventions
 *   - plain pointers + sizes, no torch / CUDA types in any signature;
 *   - every function returns a sonar_status; on error a human readable message
 *     (using the reference's own error text where it has one) is available from
 *     sonar_last_error() on the calling thread;
 *   - `_f64` entry points take HOST pointers (Go `[]float64` backing arrays,
 *     flattened row-major for `[][]float64`), copy to the device, run the CUDA
 *     kernels and copy the results back into caller-allocated buffers;
 *   - `_dev` entry points take DEVICE pointers (allocated with
 *     sonar_dev_alloc or by the caller, e.g. a torch tensor's data_ptr) and do
 *     not touch the host: they are what a device-resident chain (§8 f3) and the
 *     bench's `value` leg use;
 *   - the library is re-entrant: no global mutable state except a
 *     mutex-guarded plan cache inside the context;
 *   - there is NO CPU fallback: with no usable CUDA device sonar_init fails.
 *
 * The same header is implemented a second time by the CPU oracle
 * (oracle/sonar_oracle.cpp -> oracle/libsonar_oracle.so), which is test
 * infrastructure only.
 */
#ifndef SONAR_H_
#define SONAR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SONAR_ABI_VERSION 2   /* 2: sonar_speech_out, sonar_fingerprint_speech_f64, sonar_fp_exact_counts, sonar_nccl_*,
                                 sonar_xcorr_lag_sharded, sonar_truncate_to_alignment (additions only) */

typedef struct sonar_ctx sonar_ctx;

typedef enum sonar_status {
  SONAR_OK = 0,
  SONAR_ERR_INVALID = 1,     /* nil / non-positive argument                       */
  SONAR_ERR_EMPTY = 2,       /* "empty signal", "empty signals provided", ...     */
  SONAR_ERR_TOO_SHORT = 3,   /* "signal too short for given window size and hop size" */
  SONAR_ERR_CUDA = 4,        /* CUDA runtime failure (message carries cudaGetErrorString) */
  SONAR_ERR_NOMEM = 5,       /* device / pinned allocation failed or would exceed the limit */
  SONAR_ERR_UNSUPPORTED = 6  /* valid in the reference but outside this path's scope */
} sonar_status;

/* analyzers.WindowType (fingerprint/analyzers/windowing.go:12-24) */
typedef enum sonar_window {
  SONAR_WINDOW_HANN = 0,
  SONAR_WINDOW_HAMMING = 1,
  SONAR_WINDOW_BLACKMAN = 2,
  SONAR_WINDOW_BLACKMAN_HARRIS = 3,
  SONAR_WINDOW_KAISER = 4,
  SONAR_WINDOW_TUKEY = 5,
  SONAR_WINDOW_RECTANGULAR = 6,
  SONAR_WINDOW_BARTLETT = 7,
  SONAR_WINDOW_WELCH = 8
} sonar_window;

/* config.FeatureConfig.Enable* as the speech extractor actually reads them
 * (fingerprint/extractors/speech.go:168,179,201).  Spectral, energy and
 * harmonic groups are unconditional in the reference (speech.go:193,215,224). */
#define SONAR_FP_ENABLE_MFCC      0x1u
#define SONAR_FP_ENABLE_TEMPORAL  0x2u   /* speech.go:370-408            */
#define SONAR_FP_ENABLE_SPEECH    0x4u   /* speech.go:194-205,271-317: see sonar_speech_out */

/* ------------------------------------------------------------------------- */
/* context                                                                    */
/* ------------------------------------------------------------------------- */

/* Creates a context bound to `n_devices` CUDA devices (device_ids == NULL:
 * devices 0..n_devices-1; n_devices <= 0: the current device only).
 * Fails with SONAR_ERR_CUDA when no CUDA device is usable. */
int sonar_init(int n_devices, const int* device_ids, sonar_ctx** out);
void sonar_destroy(sonar_ctx* ctx);
/* Thread-local message of the last failing call on this thread ("" if none). */
const char* sonar_last_error(void);
int sonar_abi_version(void);
/* "cuda-sm100a" for the product library, "cpu-oracle" for the oracle. */
const char* sonar_backend(void);

/* Pinned host memory for callers that want full PCIe rate (optional). */
int sonar_host_alloc(sonar_ctx* ctx, uint64_t bytes, void** out);
int sonar_host_free(sonar_ctx* ctx, void* p);
/* Page-locks caller-owned memory in place (cudaHostRegister) so that the H2D copies of the host-pointer entry points
 * run at full PCIe rate from it: the route for a Go []float64 or a decoder's own buffer (transcode/decoder.go:850-870
 * bytesToFloat64 output), which sonar_host_alloc cannot replace.  Pageable memory is accepted everywhere too; it is
 * staged by the driver at a fraction of the rate.  `p` need not be page aligned. */
int sonar_host_register(sonar_ctx* ctx, void* p, uint64_t bytes);
int sonar_host_unregister(sonar_ctx* ctx, void* p);
/* Device memory on the context's first device (for `_dev` entry points). */
int sonar_dev_alloc(sonar_ctx* ctx, uint64_t bytes, void** out);
int sonar_dev_free(sonar_ctx* ctx, void* p);
int sonar_memcpy_h2d(sonar_ctx* ctx, void* dst_dev, const void* src_host, uint64_t bytes);
int sonar_memcpy_d2h(sonar_ctx* ctx, void* dst_host, const void* src_dev, uint64_t bytes);
int sonar_synchronize(sonar_ctx* ctx);
/* Number of kernels this context has launched since creation (bench claim). */
uint64_t sonar_kernel_launches(sonar_ctx* ctx);

/* The CUDA stream (a cudaStream_t) the `_dev` entry points enqueue on, for callers that chain
 * their own device work behind them or time them with CUDA events. NULL for the oracle. */
void* sonar_stream(sonar_ctx* ctx);

/* Per-kernel device timing (replaces the reference's stubbed getTimeMs()/ProcessingTime fields,
 * algorithms/stats/correlation.go:777-780). While enabled every kernel launch is bracketed by
 * two CUDA events on its stream; sonar_profile_read synchronises, sums them per kernel name
 * into out[0..*n_out) (at most cap entries) and clears the log. */
typedef struct sonar_kernel_time {
  char kernel[48];
  double total_ms;
  int64_t launches;
} sonar_kernel_time;
int sonar_profile_enable(sonar_ctx* ctx, int on);
int sonar_profile_read(sonar_ctx* ctx, sonar_kernel_time* out, int cap, int* n_out);

/* Diagnostic: how many frames of the most recently enqueued fingerprint batch took the float64 re-evaluation instead
 * of the FP32 kernels' results -- *spectral: STFT frames whose rolloff bin, log-magnitude means or mel bands are not safe
 * in FP32 (spectral_rolloff.go:19-55, spectral_flatness.go:31-70, mfcc.go:136-145); *pitch: YIN frames with a
 * borderline threshold / dip decision (pitch_detection.go:363-383).  Synchronises the library stream. */
int sonar_fp_exact_counts(sonar_ctx* ctx, int64_t* spectral, int64_t* pitch);

/* ------------------------------------------------------------------------- */
/* windows                                                                    */
/* ------------------------------------------------------------------------- */

/* analyzers.WindowGenerator.Generate (fingerprint/analyzers/windowing.go:77-136,
 * 246-371, 427-437) and the un-normalised algorithms/windowing structs
 * (algorithms/windowing/hann.go:16-37 ...).  Host-side table generator; the
 * device kernels consume these tables. `beta` = Kaiser, `alpha` = Tukey. */
int sonar_window_f64(int window_type, int size, int symmetric, int normalize,
                     double beta, double alpha, double* out);

/* ------------------------------------------------------------------------- */
/* fingerprint (GenerateFingerprint)                                          */
/* ------------------------------------------------------------------------- */

/* Exactly what the reference's algorithm objects were constructed with
 * (SURVEY §0 F2-F4). Zero-initialise, then sonar_fp_params_default(). */
typedef struct sonar_fp_params {
  int32_t window_size;       /* FingerprintConfig.WindowSize  (fingerprint.go:177)            */
                             /* 256 / 512 / 1024 / 2048: fused FP32 kernels (1024/256 and 512/160 with float64 re-evaluation of
                                unsafe frames); any other length in [8, 2048]: float64 route, go-dsp's transform (Bluestein) */
  int32_t hop_size;          /* FingerprintConfig.HopSize     (fingerprint.go:180)            */
  int32_t window_type;       /* FeatureConfig.WindowType      (fingerprint.go:182)            */
  int32_t algo_sample_rate;  /* extractor's config.SampleRate (speech.go:70-96). Stock
                                GenerateFingerprint passes 0 (content_config.go:87-103, F2).  */
  int32_t call_sample_rate;  /* audioData.SampleRate handed to ExtractFeatures (fingerprint.go:207) */
  int32_t energy_frame;      /* FeatureConfig.WindowSize seen by temporal.NewEnergy (speech.go:85, F4) */
  int32_t energy_hop;        /* FeatureConfig.HopSize    seen by temporal.NewEnergy               */
  int32_t n_mfcc;            /* FeatureConfig.MFCCCoefficients (<=0 -> 13, mfcc.go:59)           */
  int32_t n_mel;             /* MFCCParams.NumMelFilters       (<=0 -> 26, mfcc.go:62)           */
  int32_t use_liftering;     /* MFCCParams.UseLiftering (NewMFCC: true, mfcc.go:50)              */
  uint32_t enable;           /* SONAR_FP_ENABLE_*                                                */
  int32_t reserved0;
  double low_hz;             /* MFCCParams.LowFreq  (0)                                          */
  double high_hz;            /* MFCCParams.HighFreq (<=0 -> algo_sample_rate/2, mfcc.go:65)      */
  double lifter;             /* MFCCParams.LifterCoeff (<=0 -> 22, mfcc.go:68)                   */
  double pre_emph_alpha;     /* filters.NewPreEmphasisForContent("speech") = 0.97 (pre_emphasis.go:114) */
} sonar_fp_params;

/* Output sizes, computable up front so the Go shim can allocate flat slices. */
typedef struct sonar_fp_sizes_t {
  int64_t n_frames;          /* T  = (N-W)/H+1                     (analyzers/spectral.go:409)  */
  int64_t n_bins;            /* B  = W/2+1                         (analyzers/spectral.go:428)  */
  int64_t n_flux;            /* T-1 if T>1 else 0                  (speech.go:361)              */
  int64_t n_energy_frames;   /* Te = (N-Fe)/He+1 or 0              (temporal/energy.go:26-30)   */
  int64_t n_pitch_frames;    /* Tp = (N-1024)/512+1                (speech.go:468-470)          */
  int64_t n_mfcc;            /* effective coefficient count                                     */
  int64_t n_envelope;        /* (N-512)/256+1 or 0                 (speech.go:751-761)          */
} sonar_fp_sizes_t;

/* Caller-allocated flat outputs; any pointer may be NULL (skipped).  Layout
 * mirrors extractors.ExtractedFeatures (fingerprint/extractors/features.go:5-124). */
typedef struct sonar_fp_out {
  /* MFCC [T][n_mfcc] row-major                        (features.go:11, mfcc.go:167)       */
  double* mfcc;
  /* SpectralFeatures, each [T] (flux: [T-1])          (features.go:30-40, speech.go:320-367) */
  double* spectral_centroid;
  double* spectral_rolloff;
  double* spectral_bandwidth;
  double* spectral_flatness;
  double* spectral_crest;
  double* spectral_slope;
  double* spectral_flux;
  double* zero_crossing_rate;
  /* EnergyFeatures, each [Te]                         (features.go:97-110, speech.go:411-461) */
  double* short_time_energy;
  double* energy_entropy;
  double* low_energy_ratio;
  double* high_energy_ratio;
  /* HarmonicFeatures, each [Tp]                       (features.go:115-124, speech.go:464-509) */
  double* pitch_estimate;
  double* pitch_confidence;
  double* voicing_strength;
  double* harmonic_ratio;
  double* inharmonicity_ratio;
  double* tonal_centroid;
  /* TemporalFeatures (only with SONAR_FP_ENABLE_TEMPORAL; features.go:70-92, speech.go:370-408) */
  double* rms_energy;        /* [Te]                                                         */
  double* envelope_shape;    /* [n_envelope]                                                 */
  double* attack_time;       /* [attack_time_cap]; n_attack_time entries written              */
  int64_t attack_time_cap;
  /* scalars, written by the call */
  double energy_variance;    /* speech.go:418 */
  double loudness_range;     /* speech.go:421 */
  double dynamic_range;      /* speech.go:377 */
  double silence_ratio;      /* speech.go:380 */
  double peak_amplitude;     /* speech.go:384-392 */
  double average_amplitude;  /* speech.go:393-395 */
  double onset_density;      /* speech.go:398-399 */
  int64_t n_attack_time;     /* number of onsets (speech.go:402) */
} sonar_fp_out;

/* The frame-level part of SpeechFeatures (features.go:45-65; extractSpeechFeatures, speech.go:271-317), produced with
 * SONAR_FP_ENABLE_SPEECH by sonar_fingerprint_speech_f64.  The group runs on the pre-emphasised PCM BEFORE the harmonic
 * block and on the same pitch detector: when the signal is judged to be speech, the detector's 20-frame history carries
 * over, so the FIRST frames of pitch_estimate / tonal_centroid differ from a run without the group -- reproduced here.
 * Not computed on this path (host-side LPC / voice-quality analyzers of algorithms/speech, no data-parallel work):
 * FormantFrequencies, VocalTractLength, Jitter, Shimmer. */
typedef struct sonar_speech_out {
  double* voicing_probability;  /* [n_frames] extractVoicingProbability (speech.go:529-549)                     */
  double* spectral_tilt;        /* [n_frames] extractSpectralTilt (speech.go:551-584)                           */
  double* pause_duration;       /* [pause_cap] extractPauseDurations (speech.go:586-641); n_pause found in all  */
  int64_t pause_cap;
  int64_t n_pause;              /* written by the call                                                          */
  int64_t n_frames;             /* (N-1024)/512+1, or 0 when the signal is not speech (speech.go:281-291)       */
  int32_t is_speech;            /* SpeechAnalyzer.detectSpeech (algorithms/speech/speech_analysis.go:113-207)   */
  int32_t reserved;
  double speech_rate;           /* estimateSpeechRate (speech.go:779-797)                                       */
} sonar_speech_out;

void sonar_fp_params_default(sonar_fp_params* p);

/* Replaces the size arithmetic of ComputeSTFTWithWindow / ComputeShortTimeEnergy
 * / extractHarmonicFeatures (see sonar_fp_sizes_t). Returns the reference's
 * errors for n_samples == 0, W <= 0, H <= 0, T <= 0. */
int sonar_fp_sizes(const sonar_fp_params* p, int64_t n_samples, sonar_fp_sizes_t* out);

/* Replaces FingerprintGenerator.GenerateFingerprint's compute
 * (fingerprint/fingerprint.go:190-207): ComputeSTFTWithWindow
 * (analyzers/spectral.go:385-545) fused with SpeechFeatureExtractor.
 * ExtractFeatures (extractors/speech.go:135-243) — the spectrogram is never
 * materialised. `pcm` is a host pointer to n float64 samples. */
int sonar_fingerprint_f64(sonar_ctx* ctx, const double* pcm, int64_t n,
                          const sonar_fp_params* p, sonar_fp_out* out);

/* sonar_fingerprint_f64 with the speech-specific group (EnableSpeechFeatures: what the news / talk configurations
 * turn on, config.go): `speech` receives the frame-level SpeechFeatures, `out` the fingerprint whose harmonic block saw
 * the shared detector history (see sonar_speech_out).  The voicing / tilt arrays hold sonar_fp_sizes_t.n_pitch_frames
 * entries at most. */
int sonar_fingerprint_speech_f64(sonar_ctx* ctx, const double* pcm, int64_t n, const sonar_fp_params* p,
                                 sonar_fp_out* out, sonar_speech_out* speech);

/* Batch form (the analogue of ComputeSTFTBatch, analyzers/spectral.go:234-285,
 * applied to whole fingerprints): n_streams host buffers, one sonar_fp_out
 * each.  Streams are sharded round-robin over the context's devices and
 * pipelined (pinned double-buffered H2D) on each. */
/* PCM sample formats of the *_pcm entry points */
enum { SONAR_PCM_F64 = 0, SONAR_PCM_F32 = 1, SONAR_PCM_S16 = 2 };

int sonar_fingerprint_batch_f64(sonar_ctx* ctx, const double* const* pcm, const int64_t* n,
                                int n_streams, const sonar_fp_params* p, sonar_fp_out* outs);

/* sonar_fingerprint_batch_f64 for PCM that has not been widened to float64 yet (SONAR_PCM_F32 / SONAR_PCM_S16, see
 * sonar_align_pairs_pcm below): pcm[i] points at n[i] samples of `sample_format`; identical results to the float64
 * call on the widened samples, a half / a quarter of the PCIe bytes. */
int sonar_fingerprint_batch_pcm(sonar_ctx* ctx, const void* const* pcm, int sample_format, const int64_t* n,
                                int n_streams, const sonar_fp_params* p, sonar_fp_out* outs);

/* Device-resident batch: `pcm_dev` holds n_streams streams of `n` samples
 * each, stream s starting at pcm_dev + s*stride (stride >= n, in samples,
 * even).  `feat_dev` receives the features of stream s at feat_dev +
 * s*sonar_fp_dev_stride(p, n) doubles, in the layout described by
 * sonar_fp_dev_layout().  Asynchronous on the context's compute stream;
 * call sonar_synchronize() before reading. */
int sonar_fingerprint_batch_dev(sonar_ctx* ctx, const double* pcm_dev, int64_t n, int64_t stride,
                                int n_streams, const sonar_fp_params* p, double* feat_dev);

/* Offsets (in doubles, relative to the stream's feature block) of each array
 * in the device layout; total = block size. Arrays are in sonar_fp_out order. */
typedef struct sonar_fp_dev_layout_t {
  int64_t mfcc, spectral_centroid, spectral_rolloff, spectral_bandwidth, spectral_flatness,
      spectral_crest, spectral_slope, spectral_flux, zero_crossing_rate, short_time_energy,
      energy_entropy, low_energy_ratio, high_energy_ratio, pitch_estimate, pitch_confidence,
      voicing_strength, harmonic_ratio, inharmonicity_ratio, tonal_centroid, scalars, total;
} sonar_fp_dev_layout_t;
int sonar_fp_dev_layout(const sonar_fp_params* p, int64_t n_samples, sonar_fp_dev_layout_t* out);

/* SpectralAnalyzer.ComputeSTFTWithWindow as a materialising call
 * (analyzers/spectral.go:385-545): mag [T][B] required; phase [T][B] and cplx
 * [T][B][2] (re,im) optional (NULL = skipped). */
int sonar_stft_f64(sonar_ctx* ctx, const double* pcm, int64_t n, int win, int hop, int window_type,
                   double* mag, double* phase, double* cplx);

/* SpectralAnalyzer.ComputeSTFTStreaming + STFTStreamer.ProcessChunk (analyzers/spectral.go:289-374):
 * samples are appended to the streamer's buffer; every complete window at the head of the buffer yields one frame
 * (same window and transform as sonar_stft_f64) and the buffer advances by `hop` -- or is emptied when fewer than
 * `hop` samples remain (spectral.go:364-369: with hop > win the samples a full hop would still skip are NOT skipped).
 * sonar_stft_stream_frames tells how many frames a chunk of `chunk_len` samples would complete, so the caller can
 * size mag [T][B] (required when T > 0), phase [T][B] and cplx [T][B][2] (optional, NULL = skipped), B = win/2 + 1.
 * An empty chunk is not an error: 0 frames (spectral.go:324-326). */
typedef struct sonar_stft_stream sonar_stft_stream;
int sonar_stft_stream_open(sonar_ctx* ctx, int win, int hop, int window_type, sonar_stft_stream** out);
int64_t sonar_stft_stream_frames(const sonar_stft_stream* s, int64_t chunk_len);
int64_t sonar_stft_stream_buffered(const sonar_stft_stream* s);
int sonar_stft_stream_process(sonar_stft_stream* s, const double* chunk, int64_t n, double* mag, double* phase,
                              double* cplx, int64_t cap_frames, int64_t* n_frames);
void sonar_stft_stream_close(sonar_stft_stream* s);

/* MusicFeatureExtractor's additions to the per-frame spectral block, computed on the same magnitude
 * spectrogram ComputeSTFTWithWindow produces (fingerprint/extractors/music.go:261-302; the factory has the music
 * extractor commented out, so this is reached by direct construction only):
 *   contrast [T][n_bands]  SpectralContrast.Compute (algorithms/spectral/spectral_contrast.go:26-187): n_bands
 *                          log-spaced bands from 200 Hz to Nyquist, 10*log10(mean of the top 20 % / mean of the
 *                          bottom 20 % of the band's sorted power); music.go:111 uses 6 bands;
 *   chroma   [T][12]       ChromaSTFT.convertSTFTToChroma (algorithms/chroma/chroma_stft.go:63-138): power folded by
 *                          round(69 + 12 log2(f / 440)) mod 12 over 80..8000 Hz, then normalised to unit sum;
 *   bark     [T][n_bark]   BarkScale.ComputeBarkSpectrum (algorithms/spectral/bark_scale.go:36-128): triangular
 *                          Traunmueller filter bank between bark_low_hz and bark_high_hz applied to the power.
 * Every output is optional (NULL = skipped; n_bands / n_bark may then be 0). */
int sonar_music_spectral_f64(sonar_ctx* ctx, const double* pcm, int64_t n, int win, int hop, int window_type,
                             int sample_rate, int n_bands, double* contrast, double* chroma, int n_bark,
                             double bark_low_hz, double bark_high_hz, double* bark);

/* ------------------------------------------------------------------------- */
/* alignment: cross-correlation                                               */
/* ------------------------------------------------------------------------- */

/* stats.CorrelationResult without the arrays (algorithms/stats/correlation.go:44-71) */
typedef struct sonar_xcorr_summary {
  double peak_correlation;   /* correlation.go:526-544 (max |c|, first index wins) */
  double p_value;            /* correlation.go:547-569 */
  double snr;                /* correlation.go:572-601 (+Inf possible) */
  double sharpness;          /* correlation.go:611-619 */
  double second_peak;        /* correlation.go:622-636 */
  double peak_to_sidelobe;   /* correlation.go:639-661 (+Inf possible) */
  int32_t peak_lag;
  int32_t peak_index;
  int32_t actual_max_lag;    /* correlation.go:452-461 */
  int32_t overlap_length;    /* correlation.go:664-667 */
  int32_t is_significant;    /* p < 1-0.95, correlation.go:168 */
  int32_t n_candidates;      /* lags re-evaluated in reference summation order (diagnostic) */
} sonar_xcorr_summary;

/* CrossCorrelation.Compute as configured by stats.NewAlignmentAnalyzer
 * (algorithms/stats/alignment.go:60-81): TimeDomain, NormalizedCrossCorrelation,
 * normalizeInputs=true -> correlation.go:131-228,373-409,421-501,526-667.
 * corr receives 2*actual_max_lag+1 values (caller allocates 2*max_lag+1; NULL =
 * not wanted); lag of corr[i] is i-actual_max_lag.
 * With corr == NULL the lags are screened by an FFT and only the lag blocks that can
 * hold the peak or the second peak are evaluated in the reference's summation order:
 * peak index, peak value, second peak and sharpness are bit-identical either way, the
 * noise / side-lobe reductions (snr, peak_to_sidelobe) agree to ~1e-13. */
int sonar_xcorr_ncc_f64(sonar_ctx* ctx, const double* a, int64_t na, const double* b, int64_t nb,
                        int max_lag, double* corr, sonar_xcorr_summary* out);

/* Batch of independent pairs (BASELINE config 5): pair p uses a[p][0..na[p]),
 * b[p][0..nb[p]); corr[p] may be NULL. Sharded by pair over the devices. */
int sonar_xcorr_batch_f64(sonar_ctx* ctx, const double* const* a, const int64_t* na,
                          const double* const* b, const int64_t* nb, int n_pairs, int max_lag,
                          double* const* corr, sonar_xcorr_summary* outs);

/* Device-resident uniform batch (bench `value` leg): a_dev/b_dev hold n_pairs
 * sequences of na / nb doubles back to back; corr_dev (nullable) receives
 * n_pairs*(2*max_lag+1) values; summ_host receives the summaries after an
 * internal synchronise. */
int sonar_xcorr_batch_dev(sonar_ctx* ctx, const double* a_dev, int64_t na, const double* b_dev,
                          int64_t nb, int n_pairs, int max_lag, double* corr_dev,
                          sonar_xcorr_summary* summ_host);

/* Lag-range shard of ONE correlation: evaluates only the lag
 * indices [idx_lo, idx_hi) of the 2L+1 and returns this shard's partials.
 * The ranks exchange `sonar_xcorr_shard_peak` (16 B) with an all-gather, pick
 * the global peak with sonar_xcorr_merge_peaks, then call
 * sonar_xcorr_shard_metrics for the peak-relative partial sums, all-gather
 * those, and finish with sonar_xcorr_merge_metrics. */
typedef struct sonar_xcorr_shard_peak {
  double abs_peak;           /* max |c| in the shard, after the exact re-check     */
  int64_t index;             /* global lag index of its first occurrence; -1 = empty shard */
} sonar_xcorr_shard_peak;
typedef struct sonar_xcorr_shard_metrics {
  double noise_sum;          /* sum c^2 over |i-peak|>5 in the shard  */
  double noise_count;
  double max_sidelobe;       /* max |c| over |i-peak|>10 in the shard */
  double second_abs;         /* max |c| over i != peak in the shard   */
  double second_val;         /* signed value of its first occurrence  */
  double second_index;       /* its global index (-1 = none)          */
  double c_peak, c_prev, c_next;   /* c[peak], c[peak-1], c[peak+1] when owned, else NaN */
} sonar_xcorr_shard_metrics;
typedef struct sonar_xcorr_shard sonar_xcorr_shard;   /* opaque: device state of one shard */
int sonar_xcorr_shard_open(sonar_ctx* ctx, const double* a, int64_t na, const double* b,
                           int64_t nb, int max_lag, int64_t idx_lo, int64_t idx_hi,
                           sonar_xcorr_shard** out, sonar_xcorr_shard_peak* peak);
int sonar_xcorr_shard_metrics_f64(sonar_xcorr_shard* sh, int64_t global_peak_index,
                                  sonar_xcorr_shard_metrics* out);
/* copies this shard's correlations (idx_hi-idx_lo doubles) to the host */
int sonar_xcorr_shard_corr(sonar_xcorr_shard* sh, double* corr);
void sonar_xcorr_shard_close(sonar_xcorr_shard* sh);
/* The same sharding done inside the library (one call per rank, no host round trip between the phases): every rank
 * z-scores both sequences, evaluates its ceil((2L+1)/world) lags in reference order (correlation.go:373-449), ONE
 * ncclAllGather of the curve shards (device buffers, the library's stream) assembles the curve on every rank, and
 * findPeak / SNR / sharpness / second peak / side lobe (correlation.go:526-661) run over it: every rank returns the
 * identical summary, bit-identical to sonar_xcorr_ncc_f64.  The communicator is the library's own: rank 0 calls
 * sonar_nccl_unique_id, the host distributes the 128 bytes by any means (torch.distributed broadcast, MPI, a file),
 * every rank calls sonar_nccl_init.  NCCL is bound with dlopen("libnccl.so.2") at the first call; without a
 * communicator the call degenerates to one shard = the unsharded evaluation.  inputs_on_device != 0: a / b are device
 * pointers on the context's device 0.  corr_host (nullable) receives the 2L+1 values. */
#define SONAR_NCCL_ID_BYTES 128
int sonar_nccl_unique_id(unsigned char* id, int cap);
int sonar_nccl_init(sonar_ctx* ctx, int world, int rank, const unsigned char* id);
int sonar_nccl_shutdown(sonar_ctx* ctx);
int sonar_xcorr_lag_sharded(sonar_ctx* ctx, const double* a, int64_t na, const double* b, int64_t nb, int max_lag,
                            int inputs_on_device, double* corr_host, sonar_xcorr_summary* out);

/* pure host arithmetic over the gathered partials (same tie rules as findPeak) */
int sonar_xcorr_merge_peaks(const sonar_xcorr_shard_peak* peaks, int n, int64_t* global_index);
int sonar_xcorr_merge_metrics(const sonar_xcorr_shard_metrics* parts, int n, int64_t na,
                              int64_t nb, int max_lag, int64_t global_peak_index,
                              sonar_xcorr_summary* out);

/* stats.AlignmentResult scalars (algorithms/stats/alignment.go:34-58) */
typedef struct sonar_align_result {
  int32_t method;            /* stats.AlignmentMethod: 0 DTW, 1 cross-correlation, 3 hybrid */
  int32_t offset;            /* alignment.go:164 / :140 (samples)                  */
  double offset_seconds;     /* alignment.go:165 / :141                            */
  double confidence;         /* alignment.go:183-243 / :420-452                    */
  double similarity;         /* alignment.go:171-174 / :380-405                    */
  double alignment_quality;  /* alignment.go:245-305 / :543-566                    */
  double noise_level;        /* alignment.go:178 (can be -Inf)                     */
  double stability;          /* alignment.go:618-643 (DTW only)                    */
  int32_t query_length;
  int32_t reference_length;
  int32_t sample_rate;
  int32_t reserved0;
} sonar_align_result;

/* AlignmentAnalyzer.AlignFeatures with method = AlignmentCrossCorrelation on
 * the first feature component (alignment.go:84-106,151-181,363-378) — the call
 * extractors.alignWithFeatures makes for "corr_energy"
 * (fingerprint/extractors/alignment.go:323-332,357-409), including its
 * max-lag clamp min(maxLagFrames, minFrames-1) (:372-374). */
int sonar_align_xcorr_f64(sonar_ctx* ctx, const double* query, int64_t nq, const double* reference,
                          int64_t nr, int max_lag_frames, int hop_size, int sample_rate,
                          double* corr, sonar_xcorr_summary* xc, sonar_align_result* out);

/* ------------------------------------------------------------------------- */
/* alignment: DTW                                                             */
/* ------------------------------------------------------------------------- */

typedef enum sonar_step_pattern {
  SONAR_STEP_SYMMETRIC2 = 0, /* dtw.go:140-142 (unweighted 3-way min) */
  SONAR_STEP_ASYMMETRIC = 1, /* dtw.go:144-146 */
  SONAR_STEP_SYMMETRIC1 = 2  /* dtw.go:148-157 */
} sonar_step_pattern;

typedef enum sonar_metric {
  SONAR_METRIC_EUCLIDEAN = 0 /* distance.go:29-36; others UNSUPPORTED on this path */
} sonar_metric;

/* stats.DTWResult (algorithms/stats/dtw.go:18-34). Caller allocates path_*
 * with capacity n+m; cost_matrix (nullable) is the full [n][m+1] matrix the
 * reference returns as CostMatrix (dtw.go:96) — opt-in, SURVEY F7. */
typedef struct sonar_dtw_out {
  int32_t* path_query;       /* AlignPoint.QueryIndex (can be -1, dtw.go:169-197) */
  int32_t* path_ref;         /* AlignPoint.RefIndex                               */
  double* path_cost;         /* AlignPoint.Cost = C[i][j]-C[i-1][j-1] (dtw.go:171-174) */
  int64_t path_cap;
  int64_t path_len;          /* written                                           */
  double distance;           /* C[n][m]/len(path)      (dtw.go:88-91)             */
  double total_cost;         /* C[n][m]                                           */
  double* cost_matrix;       /* optional [n][m+1]                                 */
} sonar_dtw_out;

/* DTWAlignment.Align (algorithms/stats/dtw.go:55-217): q [n][dim], r [m][dim]
 * row-major; band <= 0 = unconstrained (dtw.go:115). */
int sonar_dtw_f64(sonar_ctx* ctx, const double* q, int n, const double* r, int m, int dim,
                  int band, int step_pattern, int metric, sonar_dtw_out* out);

/* Batch of independent pairs of identical shape (replicas). */
int sonar_dtw_batch_f64(sonar_ctx* ctx, const double* const* q, const double* const* r,
                        int n_pairs, int n, int m, int dim, int band, int step_pattern, int metric,
                        sonar_dtw_out* outs);

/* AlignmentAnalyzer.alignWithDTW's scalars from a finished path
 * (algorithms/stats/alignment.go:129-148,380-643). Host arithmetic. */
int sonar_align_dtw_scalars(const sonar_dtw_out* dtw, int n, int m, int sample_rate,
                            sonar_align_result* out);

/* ------------------------------------------------------------------------- */
/* the chained pair pipeline                                                  */
/* ------------------------------------------------------------------------- */

/* Result of one source/CDN pair.  Feature pointers inside `query` / `reference`, `corr` and the path arrays of
 * `dtw` are caller-allocated (NULL = not wanted); everything else is written by the call. */
typedef struct sonar_pair_out {
  sonar_fp_out query, reference;      /* GenerateFingerprint's features of the two streams            */
  sonar_xcorr_summary xcorr;          /* "corr_energy" cross-correlation (extractors/alignment.go:323-332) */
  sonar_align_result corr_alignment;  /* stats.AlignmentResult scalars of it (stats/alignment.go:151-181) */
  double* corr;                       /* optional [n_lags] correlation curve                            */
  sonar_dtw_out dtw;                  /* banded DTW of the lag-trimmed energy series (dtw_length frames each) */
  int32_t dtw_length, reserved0;
} sonar_pair_out;

/* n_lags = 2*clamped_max_lag_frames+1 and dtw_length = energy_frames - clamped_max_lag_frames for streams of n
 * samples (so the caller can size `corr` and the path arrays: path_cap = 2*dtw_length). */
int sonar_align_pairs_sizes(const sonar_fp_params* p, int64_t n, double max_lag_seconds, int32_t* n_lags,
                            int32_t* dtw_length);

/* The whole CDN-latency loop for n_pairs pairs of equally long streams, chained on the device
 * (SURVEY §8 f3): GenerateFingerprint of both streams (fingerprint/fingerprint.go:137-236), the "corr_energy"
 * alignment of ExtractAlignmentFeatures (fingerprint/extractors/alignment.go:139-219,357-409) with the lag clamp
 * of NewAlignmentExtractorWithMaxLag (:99-136), and DTWAlignment.Align (algorithms/stats/dtw.go:55-217,
 * band = dtw_band > 0, "symmetric2", Euclidean) of the two short-time-energy series after trimming by the
 * detected lag as TruncateToAlignmentPCM does (alignment.go:239-243) and truncating both to dtw_length.  This is
 * also the shape of AlignmentExtractor.AlignAudioFiles (alignment.go:489-560).  Results are identical to calling
 * sonar_fingerprint_f64 x2, sonar_align_xcorr_f64 and sonar_dtw_f64 one after the other; the point of the call
 * is that nothing returns to the host in between and the PCIe copy of the next pairs overlaps the kernels. */
int sonar_align_pairs_f64(sonar_ctx* ctx, const double* const* query_pcm, const double* const* reference_pcm,
                          int64_t n, int n_pairs, const sonar_fp_params* p, double max_lag_seconds, int dtw_band,
                          sonar_pair_out* outs);

/* The same with the PCM in the sample format the decoder holds BEFORE the reference widens it to float64
 * (transcode/decoder.go:707-712 asks ffmpeg for "-f f64le" and :850-870 bytesToFloat64 reinterprets the bytes;
 * ffmpeg's own s16 -> dbl conversion is x / 32768 and flt -> dbl is exact, so the float64 values the reference
 * would see are reproduced bit for bit by widening on the device).  A quarter (s16) or half (f32) of the bytes
 * cross PCIe, which is what bounds the host-pointer calls.  query_pcm[i] / reference_pcm[i] point
 * at n samples of `sample_format`; results are identical to sonar_align_pairs_f64 on the widened samples. */
int sonar_align_pairs_pcm(sonar_ctx* ctx, const void* const* query_pcm, const void* const* reference_pcm,
                          int sample_format, int64_t n, int n_pairs, const sonar_fp_params* p,
                          double max_lag_seconds, int dtw_band, sonar_pair_out* outs);

/* Device-resident form: pair i = streams 2i (query) and 2i+1 (reference) of pcm_dev, `stride` (= n rounded up to
 * even) samples apart. */
int sonar_align_pairs_dev(sonar_ctx* ctx, const double* pcm_dev, int64_t n, int64_t stride, int n_pairs,
                          const sonar_fp_params* p, double max_lag_seconds, int dtw_band, sonar_pair_out* outs);

/* ------------------------------------------------------------------------- */
/* comparison                                                                 */
/* ------------------------------------------------------------------------- */

/* extractMFCCStatistics x2 + cosineSimilarity (fingerprint/comparison.go:
 * 774-800,858-873): per column gonum mean and unbiased variance over the
 * frames, 2*dim statistics per side, cosine of the two vectors.
 * dim == 1 gives compareSequenceStats (comparison.go:827-842). */
int sonar_colstats_cosine_f64(sonar_ctx* ctx, const double* x, int64_t tx, const double* y,
                              int64_t ty, int dim, double* sim);

/* Column mean / unbiased std only (2*dim doubles: means then stds). */
int sonar_colstats_f64(sonar_ctx* ctx, const double* x, int64_t t, int dim, double* stats);

/* One side of FingerprintComparator.Compare: the feature arrays Compare reads
 * (comparison.go:266-341,646-771). NULL / zero-length = feature absent. */
typedef struct sonar_cmp_features {
  const double* mfcc; int64_t mfcc_frames; int32_t mfcc_dim; int32_t content_type;
  const double* spectral_centroid; int64_t n_centroid;
  const double* spectral_rolloff; int64_t n_rolloff;
  const double* spectral_flux; int64_t n_flux;
  int32_t has_spectral; int32_t has_harmonic;
  const double* harmonic_ratio; int64_t n_harmonic_ratio;
  const double* pitch_estimate; int64_t n_pitch;
  const double* rms_energy; int64_t n_rms;       /* temporal (comparison.go:688-718) */
  int32_t has_temporal; int32_t reserved0;
  double dynamic_range, silence_ratio, onset_density;
} sonar_cmp_features;

/* AlignmentExtractor.TruncateToAlignmentPCM (fingerprint/extractors/alignment.go:223-297): where the aligned,
 * 0.5 s-padded segments of the two streams start and how long they are, from AlignmentFeatures.TemporalOffset
 * (seconds; > 0: stream 2 is ahead).  Pure index arithmetic -- the segments are pcm1[start1 : start1+len] and
 * pcm2[start2 : start2+len]; the shim slices, re-fingerprints both (sonar_fingerprint_batch_*) and compares
 * (sonar_compare_f64): the rest of the CDN-latency loop.  Errors as the reference's: "offset too large: ...",
 * "no overlapping audio after alignment". */
int sonar_truncate_to_alignment(int64_t n1, int64_t n2, int sample_rate, double offset_seconds, int64_t* start1,
                                int64_t* start2, int64_t* length);

/* feature weights in the order mfcc, spectral, chroma, temporal, speech,
 * harmonic, energy (comparison.go:1055-1104; fp1.Metadata["feature_weights"]) */
typedef struct sonar_cmp_weights { double w[7]; } sonar_cmp_weights;

/* SimilarityResult scalars (comparison.go:28-39) */
typedef struct sonar_cmp_result {
  double overall_similarity, feature_similarity, confidence;
  double dist_mfcc, dist_spectral, dist_temporal, dist_harmonic;   /* NaN = feature not compared */
  int32_t content_type_match; int32_t n_features;
} sonar_cmp_result;

/* FingerprintComparator.Compare (comparison.go:133-194) on flattened features. */
int sonar_compare_f64(sonar_ctx* ctx, const sonar_cmp_features* f1, const sonar_cmp_features* f2,
                      const sonar_cmp_weights* weights, int enable_content_filter,
                      sonar_cmp_result* out);

/* FingerprintComparator.BatchCompare / the comparison loop of FindBestMatches (comparison.go:1107-1151, 197-263):
 * one query against n candidates, results[i] for candidates[i] (a NULL candidate is skipped: the reference drops
 * it; its result is zeroed with n_features = -1).  Thresholding, sorting and ranking of FindBestMatches stay in
 * the shim: they are O(n log n) on n scalars. */
int sonar_compare_batch_f64(sonar_ctx* ctx, const sonar_cmp_features* query,
                            const sonar_cmp_features* const* candidates, int n, const sonar_cmp_weights* weights,
                            int enable_content_filter, sonar_cmp_result* results);

#ifdef __cplusplus
}
#endif
#endif /* SONAR_H_ */
