// Host-side mirror of the reference's Go API for the hot path, header-only C++17 over the C ABI
// (include/sonar.h).  The reference is Go; no Go toolchain exists in this image, so the host layer
// a maintainer would write in Go (go/) is mirrored here in C++ with the same names,
// argument meaning, error strings and configuration plumbing — including the quirks a drop-in must
// reproduce:
//   F1  every content type gets the speech extractor          fingerprint/extractors/feature_extractor.go:38-62
//   F2  the extractor's config.SampleRate is 0                fingerprint/content_config.go:87-103
//   F4  the extractor sees FeatureConfig.WindowSize/HopSize of the BASE config, copied before the
//       top-level values are patched in                       fingerprint/fingerprint.go:171 vs :177-181
// All signal-sized arithmetic happens in libsonar.so (CUDA); this file is orchestration, result
// structs and O(1) formulas only.  Go `(T, error)` returns become Result<T>.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <complex>
#include <vector>

#include "../../include/sonar.h"

namespace sonido {

template <class T>
struct Result {
  std::shared_ptr<T> value;  // nil on error, like the reference
  std::string err;           // "" == nil error
  bool ok() const { return err.empty(); }
  T* operator->() const { return value.get(); }
};
template <class T>
Result<T> Err(const std::string& m) {
  return Result<T>{nullptr, m};
}

// One lazily created sonar_ctx per process (the reference's objects are created per call and own no
// device state; the C layer is re-entrant).
class Runtime {
 public:
  static sonar_ctx* Ctx(std::string* err = nullptr) {
    static Runtime rt;
    if (!rt.ctx_ && err) *err = rt.err_;
    return rt.ctx_;
  }

 private:
  Runtime() {
    if (sonar_init(0, nullptr, &ctx_) != SONAR_OK) {
      err_ = sonar_last_error();
      ctx_ = nullptr;
    }
  }
  ~Runtime() {
    if (ctx_) sonar_destroy(ctx_);
  }
  sonar_ctx* ctx_ = nullptr;
  std::string err_;
};

// ----------------------------------------------------------------------------------------------
namespace config {  // fingerprint/config/config.go

using ContentType = std::string;
const ContentType ContentMusic = "music", ContentNews = "news", ContentSports = "sports", ContentTalk = "talk",
                  ContentMixed = "mixed", ContentUnknown = "unknown";

inline ContentType ToContentType(const std::string& s) {  // config.go:50-65
  for (const auto& c : {ContentMusic, ContentNews, ContentSports, ContentTalk, ContentMixed})
    if (s == c) return c;
  return ContentUnknown;
}

struct FeatureConfig {  // config.go:13-37
  int SampleRate = 0, WindowSize = 0, HopSize = 0;
  double FreqRange[2] = {0, 0};
  std::string WindowType;  // analyzers.WindowType ("hann", ...)
  bool EnableChroma = false, EnableMFCC = false, EnableSpectralContrast = false, EnableTemporalFeatures = false,
       EnableSpeechFeatures = false, EnableHarmonicFeatures = false;
  int MFCCCoefficients = 0, ChromaBins = 0, ContrastBands = 0;
  std::map<std::string, double> SimilarityWeights;
  double MatchThreshold = 0;
};

struct ContentAwareConfig {  // config.go:5-11
  bool EnableContentDetection = false;
  ContentType DefaultContentType;
  double AutoDetectThreshold = 0;
  std::string FallbackStrategy;
};

struct ComparisonConfig {  // config.go:68-80
  double SimilarityThreshold = 0.75;
  std::string Method = "auto";
  bool EnableDetailedMetrics = false;
  int MaxCandidates = 50;
  bool EnableContentFilter = false;
};
inline ComparisonConfig DefaultComparisonConfig() { return ComparisonConfig{}; }  // config.go:120-128

struct AlignmentConfig {  // config.go:82-101
  double MaxLagSeconds = 30.0, MinConfidence = 0.6;
  int StepSize = 1;
  std::string PreferredMethod = "hybrid", FallbackMethod = "correlation";
  double MinSimilarity = 0.3, MinQuality = 0.4;
  int DTWBandRadius = 50;
  bool CorrNormalize = true;
  int ConsistencyTrials = 5;
  double NoiseThreshold = 0.1;
};
inline AlignmentConfig DefaultAlignmentConfig() { return AlignmentConfig{}; }  // config.go:103-117

}  // namespace config

// ----------------------------------------------------------------------------------------------
namespace transcode {  // transcode/decoder.go:21-46 (types only; decoding is out of scope)
struct StreamMetadata {
  std::string URL, ContentType;
};
struct AudioData {
  std::vector<double> PCM;
  int SampleRate = 0, Channels = 1;
  std::shared_ptr<StreamMetadata> Metadata;
};
}  // namespace transcode

namespace analyzers {
// fingerprint/analyzers/spectral.go:22-33. On the fused GPU path the spectrogram never leaves the
// device: Magnitude stays empty and only the header fields (the ones speech.go reads) are filled.
struct SpectrogramResult {
  int TimeFrames = 0, FreqBins = 0, SampleRate = 0, WindowSize = 0, HopSize = 0;
  std::string WindowType = "hann";
  std::vector<std::vector<double>> Magnitude;
};
inline int WindowTypeId(const std::string& w) {  // analyzers/windowing.go:12-24
  static const std::map<std::string, int> ids = {
      {"hann", SONAR_WINDOW_HANN},         {"hamming", SONAR_WINDOW_HAMMING},
      {"blackman", SONAR_WINDOW_BLACKMAN}, {"blackman_harris", SONAR_WINDOW_BLACKMAN_HARRIS},
      {"kaiser", SONAR_WINDOW_KAISER},     {"tukey", SONAR_WINDOW_TUKEY},
      {"rectangular", SONAR_WINDOW_RECTANGULAR}, {"bartlett", SONAR_WINDOW_BARTLETT},
      {"welch", SONAR_WINDOW_WELCH}};
  auto it = ids.find(w);
  return it == ids.end() ? SONAR_WINDOW_HANN : it->second;
}

// analyzers/spectral.go:376-381
struct SpectrogramFrame {
  std::vector<double> Magnitude, Phase;
  std::vector<std::complex<double>> Complex;
};

// STFTStreamer (analyzers/spectral.go:312-374) over sonar_stft_stream_*: same method, same results, same quirks
// (an empty chunk yields no frames and no error; with hop > window an emptied buffer does not skip ahead).
class STFTStreamer {
 public:
  ~STFTStreamer() {
    if (h_) sonar_stft_stream_close(h_);
  }
  STFTStreamer(const STFTStreamer&) = delete;
  STFTStreamer& operator=(const STFTStreamer&) = delete;

  // ProcessChunk(chunk []float64) ([]*SpectrogramFrame, error); the error comes back in *err ("" == nil)
  std::vector<std::shared_ptr<SpectrogramFrame>> ProcessChunk(const std::vector<double>& chunk, std::string* err) {
    if (err) err->clear();
    std::vector<std::shared_ptr<SpectrogramFrame>> frames;
    if (chunk.empty()) return frames;  // :324-326
    const int64_t T = sonar_stft_stream_frames(h_, (int64_t)chunk.size());
    const size_t B = (size_t)freqBins_;
    std::vector<double> mag((size_t)T * B), ph((size_t)T * B), cx((size_t)T * B * 2);
    int64_t got = 0;
    if (sonar_stft_stream_process(h_, chunk.data(), (int64_t)chunk.size(), mag.data(), ph.data(), cx.data(), T, &got) !=
        SONAR_OK) {
      if (err) *err = sonar_last_error();
      return frames;
    }
    for (int64_t t = 0; t < got; t++) {
      auto f = std::make_shared<SpectrogramFrame>();
      f->Magnitude.assign(mag.begin() + t * B, mag.begin() + (t + 1) * B);
      f->Phase.assign(ph.begin() + t * B, ph.begin() + (t + 1) * B);
      f->Complex.resize(B);
      for (size_t k = 0; k < B; k++) f->Complex[k] = {cx[(t * B + k) * 2], cx[(t * B + k) * 2 + 1]};
      frames.push_back(std::move(f));
    }
    return frames;
  }
  int BufferedSamples() const { return (int)sonar_stft_stream_buffered(h_); }

 private:
  friend Result<STFTStreamer> ComputeSTFTStreaming(int, int, const std::string&);
  STFTStreamer(sonar_stft_stream* h, int bins) : h_(h), freqBins_(bins) {}
  sonar_stft_stream* h_ = nullptr;
  int freqBins_ = 0;
};

// SpectralAnalyzer.ComputeSTFTStreaming(windowSize, hopSize, windowType) (*STFTStreamer, error)  (spectral.go:289-310)
inline Result<STFTStreamer> ComputeSTFTStreaming(int windowSize, int hopSize, const std::string& windowType) {
  std::string rerr;
  sonar_ctx* ctx = Runtime::Ctx(&rerr);
  if (!ctx) return Err<STFTStreamer>(rerr);
  sonar_stft_stream* h = nullptr;
  if (sonar_stft_stream_open(ctx, windowSize, hopSize, WindowTypeId(windowType), &h) != SONAR_OK)
    return Err<STFTStreamer>(std::string("failed to generate window: ") + sonar_last_error());
  return Result<STFTStreamer>{std::shared_ptr<STFTStreamer>(new STFTStreamer(h, windowSize / 2 + 1)), ""};
}
}  // namespace analyzers

// ----------------------------------------------------------------------------------------------
namespace stats {  // algorithms/stats
struct CorrelationResult {  // correlation.go:44-71
  std::vector<double> Correlations;
  std::vector<int> Lags;
  double PeakCorrelation = 0, PValue = 0, SNR = 0, Sharpness = 0, SecondPeak = 0, PeakToSidelobe = 0;
  int PeakLag = 0, PeakIndex = 0, MaxLag = 0, OverlapLength = 0;
  bool IsSignificant = false;
};
struct AlignPoint {  // dtw.go:18-22
  int QueryIndex, RefIndex;
  double Cost;
};
struct DTWResult {  // dtw.go:24-34 (CostMatrix is opt-in on the GPU path: SURVEY F7)
  double Distance = 0;
  std::vector<AlignPoint> Path;
  int QueryLength = 0, RefLength = 0;
};
struct AlignmentResult {  // alignment.go:34-58
  std::string Method;
  int Offset = 0;
  double OffsetSeconds = 0, Confidence = 0, Similarity = 0, AlignmentQuality = 0, NoiseLevel = 0, Stability = 0;
  std::shared_ptr<stats::CorrelationResult> CrossCorrResult;
  std::shared_ptr<stats::DTWResult> DTWResult;
  int QueryLength = 0, ReferenceLength = 0, SampleRate = 0;
};
}  // namespace stats

// ----------------------------------------------------------------------------------------------
namespace extractors {  // fingerprint/extractors

struct SpectralFeatures {  // features.go:30-40
  std::vector<double> SpectralCentroid, SpectralRolloff, SpectralBandwidth, SpectralFlatness, SpectralCrest,
      SpectralSlope, SpectralFlux, ZeroCrossingRate;
};
struct EnergyFeatures {  // features.go:97-110
  std::vector<double> ShortTimeEnergy, EnergyEntropy, LowEnergyRatio, HighEnergyRatio;
  double EnergyVariance = 0, LoudnessRange = 0;
};
struct HarmonicFeatures {  // features.go:115-124
  std::vector<double> PitchEstimate, PitchConfidence, VoicingStrength, HarmonicRatio, InharmonicityRatio, TonalCentroid;
};
struct TemporalFeatures {  // features.go:70-92
  std::vector<double> RMSEnergy, AttackTime, EnvelopeShape;
  double DynamicRange = 0, SilenceRatio = 0, PeakAmplitude = 0, AverageAmplitude = 0, OnsetDensity = 0;
};
struct SpeechFeatures {  // features.go:45-65
  std::vector<std::vector<double>> FormantFrequencies;  // host-side LPC analysis in the reference: not produced here
  std::vector<double> VoicingProbability, SpectralTilt, PauseDuration;
  double SpeechRate = 0.0;
  double VocalTractLength = 17.5;  // speech.go:288 default; the formant analyzer's estimate is not produced here
  double Jitter = 0.0, Shimmer = 0.0;  // voice-quality analyzer: not produced here
};

struct ExtractedFeatures {  // features.go:5-27
  std::vector<std::vector<double>> MFCC, ChromaFeatures;
  std::shared_ptr<sonido::extractors::SpeechFeatures> SpeechFeatures;
  std::shared_ptr<extractors::SpectralFeatures> SpectralFeatures;
  std::shared_ptr<extractors::EnergyFeatures> EnergyFeatures;
  std::shared_ptr<extractors::HarmonicFeatures> HarmonicFeatures;
  std::shared_ptr<extractors::TemporalFeatures> TemporalFeatures;
  std::map<std::string, std::string> ExtractionMetadata;
};

class FeatureExtractor {  // feature_extractor.go:10-15 — the reference's plug-in point
 public:
  virtual ~FeatureExtractor() = default;
  virtual Result<ExtractedFeatures> ExtractFeatures(const analyzers::SpectrogramResult* spectrogram,
                                                    const std::vector<double>& pcm, int sampleRate) = 0;
  virtual std::map<std::string, double> GetFeatureWeights() const = 0;
  virtual std::string GetName() const = 0;
  virtual config::ContentType GetContentType() const = 0;
};

// SpeechFeatureExtractor (speech.go) with its compute replaced by sonar_fingerprint_f64.
class SpeechFeatureExtractor : public FeatureExtractor {
 public:
  SpeechFeatureExtractor(const config::FeatureConfig& cfg, bool isNews) : config_(cfg), isNews_(isNews) {}
  std::string GetName() const override { return "SpeechFeatureExtractor"; }  // speech.go:100
  config::ContentType GetContentType() const override { return isNews_ ? config::ContentNews : config::ContentTalk; }
  std::map<std::string, double> GetFeatureWeights() const override {  // speech.go:111-133
    if (!config_.SimilarityWeights.empty()) return config_.SimilarityWeights;
    std::map<std::string, double> w = {{"mfcc", 0.40}, {"speech", 0.35}, {"spectral", 0.15}, {"temporal", 0.10}};
    if (isNews_) w["speech"] = 0.40, w["mfcc"] = 0.35;
    return w;
  }
  // algorithm-construction parameters exactly as NewSpeechFeatureExtractor passes them (speech.go:58-97)
  sonar_fp_params Params(const analyzers::SpectrogramResult& sp, int callSampleRate) const {
    sonar_fp_params p;
    sonar_fp_params_default(&p);
    p.window_size = sp.WindowSize;
    p.hop_size = sp.HopSize;
    p.window_type = analyzers::WindowTypeId(sp.WindowType);
    p.algo_sample_rate = config_.SampleRate;  // 0 through GenerateFingerprint (F2/F3)
    p.call_sample_rate = callSampleRate;
    p.energy_frame = config_.WindowSize;  // temporal.NewEnergy(config.WindowSize, config.HopSize, ...) (F4)
    p.energy_hop = config_.HopSize;
    p.n_mfcc = config_.MFCCCoefficients;  // spectral.NewMFCC(sr, n): n <= 0 -> 13
    p.enable = 0;
    if (config_.EnableMFCC) p.enable |= SONAR_FP_ENABLE_MFCC;
    if (config_.EnableTemporalFeatures) p.enable |= SONAR_FP_ENABLE_TEMPORAL;  // speech.go:201-211
    // EnableSpeechFeatures (news, talk; speech.go:180-190): the frame-level group (IsSpeech gate, voicing sweep with
    // the detector history it leaves behind, spectral tilt, pauses, speech rate) runs on the device; the LPC / formant /
    // voice-quality scalars are host-side analyzers outside this path and keep their defaults.
    if (config_.EnableSpeechFeatures) p.enable |= SONAR_FP_ENABLE_SPEECH;
    return p;
  }
  Result<ExtractedFeatures> ExtractFeatures(const analyzers::SpectrogramResult* spectrogram,
                                            const std::vector<double>& pcm, int sampleRate) override {
    if (!spectrogram) return Err<ExtractedFeatures>("spectrogram cannot be nil");  // speech.go:137
    if (pcm.empty()) return Err<ExtractedFeatures>("PCM data cannot be empty");    // :140
    if (sampleRate <= 0) return Err<ExtractedFeatures>("sample rate must be positive");  // :143
    std::string rerr;
    sonar_ctx* ctx = Runtime::Ctx(&rerr);
    if (!ctx) return Err<ExtractedFeatures>(rerr);
    const sonar_fp_params p = Params(*spectrogram, sampleRate);
    sonar_fp_sizes_t sz;
    if (sonar_fp_sizes(&p, (int64_t)pcm.size(), &sz) != SONAR_OK) return Err<ExtractedFeatures>(sonar_last_error());
    const size_t T = (size_t)sz.n_frames, Te = (size_t)sz.n_energy_frames, Tp = (size_t)sz.n_pitch_frames;
    std::vector<double> mfcc(T * (size_t)sz.n_mfcc);
    auto f = std::make_shared<ExtractedFeatures>();
    auto sf = std::make_shared<extractors::SpectralFeatures>();
    auto ef = std::make_shared<extractors::EnergyFeatures>();
    auto hf = std::make_shared<extractors::HarmonicFeatures>();
    for (auto* v : {&sf->SpectralCentroid, &sf->SpectralRolloff, &sf->SpectralBandwidth, &sf->SpectralFlatness,
                    &sf->SpectralCrest, &sf->SpectralSlope, &sf->ZeroCrossingRate})
      v->assign(T, 0.0);
    sf->SpectralFlux.assign((size_t)sz.n_flux, 0.0);
    for (auto* v : {&ef->ShortTimeEnergy, &ef->EnergyEntropy, &ef->LowEnergyRatio, &ef->HighEnergyRatio}) v->assign(Te, 0.0);
    for (auto* v : {&hf->PitchEstimate, &hf->PitchConfidence, &hf->VoicingStrength, &hf->HarmonicRatio,
                    &hf->InharmonicityRatio, &hf->TonalCentroid})
      v->assign(Tp, 0.0);
    auto tf = std::make_shared<extractors::TemporalFeatures>();
    sonar_fp_out o;
    std::memset(&o, 0, sizeof(o));
    if (config_.EnableTemporalFeatures) {
      tf->RMSEnergy.assign(Te, 0.0);
      tf->EnvelopeShape.assign((size_t)sz.n_envelope, 0.0);
      tf->AttackTime.assign(Te > 0 ? Te : 1, 0.0);
      o.rms_energy = tf->RMSEnergy.data();
      o.envelope_shape = tf->EnvelopeShape.data();
      o.attack_time = tf->AttackTime.data();
      o.attack_time_cap = (int64_t)tf->AttackTime.size();
    }
    o.mfcc = mfcc.data();
    o.spectral_centroid = sf->SpectralCentroid.data();
    o.spectral_rolloff = sf->SpectralRolloff.data();
    o.spectral_bandwidth = sf->SpectralBandwidth.data();
    o.spectral_flatness = sf->SpectralFlatness.data();
    o.spectral_crest = sf->SpectralCrest.data();
    o.spectral_slope = sf->SpectralSlope.data();
    o.spectral_flux = sf->SpectralFlux.data();
    o.zero_crossing_rate = sf->ZeroCrossingRate.data();
    o.short_time_energy = ef->ShortTimeEnergy.data();
    o.energy_entropy = ef->EnergyEntropy.data();
    o.low_energy_ratio = ef->LowEnergyRatio.data();
    o.high_energy_ratio = ef->HighEnergyRatio.data();
    o.pitch_estimate = hf->PitchEstimate.data();
    o.pitch_confidence = hf->PitchConfidence.data();
    o.voicing_strength = hf->VoicingStrength.data();
    o.harmonic_ratio = hf->HarmonicRatio.data();
    o.inharmonicity_ratio = hf->InharmonicityRatio.data();
    o.tonal_centroid = hf->TonalCentroid.data();
    auto spf = std::make_shared<sonido::extractors::SpeechFeatures>();
    if (config_.EnableSpeechFeatures) {
      sonar_speech_out so;
      std::memset(&so, 0, sizeof(so));
      spf->VoicingProbability.assign(Tp, 0.0);
      spf->SpectralTilt.assign(Tp, 0.0);
      spf->PauseDuration.assign(Te / 2 + 1, 0.0);
      so.voicing_probability = spf->VoicingProbability.data();
      so.spectral_tilt = spf->SpectralTilt.data();
      so.pause_duration = spf->PauseDuration.data();
      so.pause_cap = (int64_t)spf->PauseDuration.size();
      if (sonar_fingerprint_speech_f64(ctx, pcm.data(), (int64_t)pcm.size(), &p, &o, &so) != SONAR_OK)
        return Err<ExtractedFeatures>(sonar_last_error());
      spf->VoicingProbability.resize((size_t)so.n_frames);
      spf->SpectralTilt.resize((size_t)so.n_frames);
      spf->PauseDuration.resize((size_t)std::min<int64_t>(so.n_pause, so.pause_cap));
      spf->SpeechRate = so.speech_rate;
      f->SpeechFeatures = spf;
    } else if (sonar_fingerprint_f64(ctx, pcm.data(), (int64_t)pcm.size(), &p, &o) != SONAR_OK) {
      return Err<ExtractedFeatures>(sonar_last_error());
    }
    if (config_.EnableMFCC) {  // speech.go:168-178
      f->MFCC.assign(T, std::vector<double>((size_t)sz.n_mfcc));
      for (size_t t = 0; t < T; t++)
        std::copy(mfcc.begin() + t * sz.n_mfcc, mfcc.begin() + (t + 1) * sz.n_mfcc, f->MFCC[t].begin());
    }
    ef->EnergyVariance = o.energy_variance;
    ef->LoudnessRange = o.loudness_range;
    f->SpectralFeatures = sf;  // unconditional in the reference (speech.go:193,215,224)
    f->EnergyFeatures = ef;
    f->HarmonicFeatures = hf;
    if (config_.EnableTemporalFeatures) {
      tf->AttackTime.resize((size_t)o.n_attack_time);
      tf->DynamicRange = o.dynamic_range, tf->SilenceRatio = o.silence_ratio, tf->PeakAmplitude = o.peak_amplitude;
      tf->AverageAmplitude = o.average_amplitude, tf->OnsetDensity = o.onset_density;
      f->TemporalFeatures = tf;
    }
    f->ExtractionMetadata = {{"extractor_type", "speech"},  // speech.go:233-239
                             {"content_subtype", isNews_ ? "news" : "talk"},
                             {"algorithms_used", "speech,spectral,temporal,filters,tonal"},
                             {"pre_emphasis_applied", "true"},
                             {"sample_rate", std::to_string(sampleRate)},
                             {"spectrogram_frames", std::to_string(spectrogram->TimeFrames)},
                             {"optimization", "speech_optimized"},
                             {"backend", sonar_backend()}};
    if (config_.EnableSpeechFeatures)
      f->ExtractionMetadata["speech_features"] = "frame-level group on the device; formants / jitter / shimmer: host analyzers, not run";
    return Result<ExtractedFeatures>{f, ""};
  }
  const config::FeatureConfig& Config() const { return config_; }

 private:
  config::FeatureConfig config_;
  bool isNews_;
};

struct FeatureExtractorFactory {  // feature_extractor.go:32-63: music/sports/mixed are commented out (F1)
  Result<FeatureExtractor> CreateExtractor(const config::ContentType& ct, const config::FeatureConfig& fc) const {
    const bool isNews = ct != config::ContentTalk;
    return Result<FeatureExtractor>{std::make_shared<SpeechFeatureExtractor>(fc, isNews), ""};
  }
};
inline FeatureExtractorFactory NewFeatureExtractorFactory() { return FeatureExtractorFactory{}; }

// ---- alignment (alignment.go) ------------------------------------------------------------------
struct AlignmentResult {  // alignment.go:62-67
  std::shared_ptr<stats::AlignmentResult> Result;  // embedded *stats.AlignmentResult
  std::string FeatureType, ErrorMsg;
  bool Success = false;
};
struct AlignmentFeatures {  // alignment.go:35-59
  std::shared_ptr<AlignmentResult> BestAlignment, DTWAlignment, CorrAlignment;
  double TemporalOffset = 0, OffsetConfidence = 0, TimeStretch = 0, AlignmentSimilarity = 0, AlignmentQuality = 0;
  std::map<std::string, double> FeatureSimilarity;
  std::string Method;
  double QueryLength = 0, ReferenceLength = 0;
};

class AlignmentExtractor {
 public:
  // NewAlignmentExtractorWithMaxLag (alignment.go:99-136)
  AlignmentExtractor(const config::FeatureConfig* featureConf, const config::AlignmentConfig* alignmentConf,
                     double maxLagSeconds)
      : config_(*featureConf), maxLagSeconds_(maxLagSeconds), confidenceThresh_(alignmentConf->MinConfidence) {
    maxLagSamples_ = (int)(maxLagSeconds * (double)featureConf->SampleRate);
  }
  Result<AlignmentFeatures> ExtractAlignmentFeatures(const ExtractedFeatures* queryFeatures,
                                                     const ExtractedFeatures* referenceFeatures,
                                                     const std::vector<double>& queryPCM,
                                                     const std::vector<double>& referencePCM, int sampleRate) const {
    if (!queryFeatures || !referenceFeatures) return Err<AlignmentFeatures>("feature sets cannot be nil");  // :146
    std::string rerr;
    sonar_ctx* ctx = Runtime::Ctx(&rerr);
    if (!ctx) return Err<AlignmentFeatures>(rerr);
    auto res = std::make_shared<AlignmentFeatures>();
    res->QueryLength = (double)queryPCM.size() / (double)sampleRate;
    res->ReferenceLength = (double)referencePCM.size() / (double)sampleRate;
    std::map<std::string, std::shared_ptr<AlignmentResult>> alignments;  // performMultiFeatureAlignment :300-354
    const auto &qe = queryFeatures->EnergyFeatures, &re = referenceFeatures->EnergyFeatures;
    if (qe && re && !qe->ShortTimeEnergy.empty() && !re->ShortTimeEnergy.empty())
      alignments["corr_energy"] = AlignCorr(ctx, qe->ShortTimeEnergy, re->ShortTimeEnergy, sampleRate);
    if (!queryFeatures->ChromaFeatures.empty() && !referenceFeatures->ChromaFeatures.empty())
      alignments["dtw_chroma"] = AlignDtw(ctx, queryFeatures->ChromaFeatures, referenceFeatures->ChromaFeatures, sampleRate);
    // selectBestAlignment :412-445
    static const std::map<std::string, double> weights = {{"corr_energy", 1.0}, {"dtw_chroma", 0.7}};
    double bestScore = 0.0;
    for (auto& kv : alignments) {
      auto& a = kv.second;
      if (!a->Success || !a->Result) continue;
      auto it = weights.find(kv.first);
      const double w = it == weights.end() ? 0.5 : it->second;
      const double score = w * (0.4 * a->Result->Confidence + 0.4 * a->Result->Similarity + 0.2 * a->Result->AlignmentQuality);
      if (score > bestScore) bestScore = score, res->BestAlignment = a;
    }
    if (res->BestAlignment) {  // :169-176
      res->TemporalOffset = res->BestAlignment->Result->OffsetSeconds;
      res->OffsetConfidence = res->BestAlignment->Result->Confidence;
      res->AlignmentSimilarity = res->BestAlignment->Result->Similarity;
      res->AlignmentQuality = res->BestAlignment->Result->AlignmentQuality;
      res->Method = res->BestAlignment->FeatureType;
    }
    for (auto& kv : alignments) {  // :179-195
      if (kv.first == "corr_energy" && kv.second->Result && kv.second->Result->CrossCorrResult) res->CorrAlignment = kv.second;
      if (kv.second->Success) res->FeatureSimilarity[kv.first] = kv.second->Result->Similarity;
    }
    res->TimeStretch = EstimateTimeStretch(res->BestAlignment.get(), res->QueryLength, res->ReferenceLength);
    return Result<AlignmentFeatures>{res, ""};
  }

 private:
  std::shared_ptr<AlignmentResult> AlignCorr(sonar_ctx* ctx, const std::vector<double>& q, const std::vector<double>& r,
                                             int sampleRate) const {  // alignWithFeatures :357-409
    auto out = std::make_shared<AlignmentResult>();
    out->FeatureType = "corr_energy";
    const int maxLagFrames = config_.HopSize > 0 ? maxLagSamples_ / config_.HopSize : 0;
    std::vector<double> corr((size_t)2 * (size_t)std::max(maxLagFrames, 0) + 1);
    sonar_xcorr_summary xs;
    sonar_align_result ar;
    if (sonar_align_xcorr_f64(ctx, q.data(), (int64_t)q.size(), r.data(), (int64_t)r.size(), maxLagFrames,
                              config_.HopSize, sampleRate, corr.data(), &xs, &ar) != SONAR_OK) {
      out->ErrorMsg = sonar_last_error();  // alignment failures are results, not errors (:389-396)
      return out;
    }
    auto sr = std::make_shared<stats::AlignmentResult>();
    sr->Method = "cross_correlation";
    sr->Offset = ar.offset, sr->OffsetSeconds = ar.offset_seconds, sr->Confidence = ar.confidence;
    sr->Similarity = ar.similarity, sr->AlignmentQuality = ar.alignment_quality, sr->NoiseLevel = ar.noise_level;
    sr->QueryLength = ar.query_length, sr->ReferenceLength = ar.reference_length, sr->SampleRate = sampleRate;
    auto cr = std::make_shared<stats::CorrelationResult>();
    const int nl = 2 * xs.actual_max_lag + 1;
    cr->Correlations.assign(corr.begin(), corr.begin() + nl);
    cr->Lags.resize(nl);
    for (int i = 0; i < nl; i++) cr->Lags[i] = i - xs.actual_max_lag;
    cr->PeakCorrelation = xs.peak_correlation, cr->PValue = xs.p_value, cr->SNR = xs.snr, cr->Sharpness = xs.sharpness;
    cr->SecondPeak = xs.second_peak, cr->PeakToSidelobe = xs.peak_to_sidelobe, cr->PeakLag = xs.peak_lag;
    cr->PeakIndex = xs.peak_index, cr->MaxLag = xs.actual_max_lag, cr->OverlapLength = xs.overlap_length;
    cr->IsSignificant = xs.is_significant != 0;
    sr->CrossCorrResult = cr;
    out->Result = sr;
    out->Success = true;
    return out;
  }
  std::shared_ptr<AlignmentResult> AlignDtw(sonar_ctx* ctx, const std::vector<std::vector<double>>& q,
                                            const std::vector<std::vector<double>>& r, int sampleRate) const {
    auto out = std::make_shared<AlignmentResult>();
    out->FeatureType = "dtw_chroma";
    const int n = (int)q.size(), m = (int)r.size(), dim = (int)q[0].size();
    std::vector<double> fq((size_t)n * dim), fr((size_t)m * dim);  // cgo cannot pass [][]float64: flatten
    for (int i = 0; i < n; i++) std::copy(q[i].begin(), q[i].end(), fq.begin() + (size_t)i * dim);
    for (int i = 0; i < m; i++) std::copy(r[i].begin(), r[i].end(), fr.begin() + (size_t)i * dim);
    std::vector<int32_t> pq((size_t)n + m), pr((size_t)n + m);
    std::vector<double> pc((size_t)n + m);
    sonar_dtw_out d;
    std::memset(&d, 0, sizeof(d));
    d.path_query = pq.data(), d.path_ref = pr.data(), d.path_cost = pc.data(), d.path_cap = n + m;
    sonar_align_result ar;
    if (sonar_dtw_f64(ctx, fq.data(), n, fr.data(), m, dim, -1, SONAR_STEP_SYMMETRIC2, SONAR_METRIC_EUCLIDEAN, &d) != SONAR_OK ||
        sonar_align_dtw_scalars(&d, n, m, sampleRate, &ar) != SONAR_OK) {
      out->ErrorMsg = sonar_last_error();
      return out;
    }
    auto sr = std::make_shared<stats::AlignmentResult>();
    sr->Method = "dtw";
    sr->Offset = ar.offset, sr->OffsetSeconds = ar.offset_seconds, sr->Confidence = ar.confidence;
    sr->Similarity = ar.similarity, sr->AlignmentQuality = ar.alignment_quality, sr->Stability = ar.stability;
    sr->QueryLength = n, sr->ReferenceLength = m, sr->SampleRate = sampleRate;
    auto dr = std::make_shared<stats::DTWResult>();
    dr->Distance = d.distance, dr->QueryLength = n, dr->RefLength = m;
    for (int64_t k = 0; k < d.path_len; k++) dr->Path.push_back({pq[k], pr[k], pc[k]});
    sr->DTWResult = dr;
    out->Result = sr;
    out->Success = true;
    return out;
  }
  static double EstimateTimeStretch(const AlignmentResult* a, double queryLen, double refLen) {  // :448-476
    if (!a || !a->Success || queryLen <= 0 || refLen <= 0) return 1.0;
    const double lengthRatio = queryLen / refLen;
    if (a->Result->DTWResult && a->Result->DTWResult->Path.size() > 1) {
      const auto& p = a->Result->DTWResult->Path;
      const double qs = (double)(p.back().QueryIndex - p.front().QueryIndex + 1);
      const double rs = (double)(p.back().RefIndex - p.front().RefIndex + 1);
      if (rs > 0) return 0.7 * (qs / rs) + 0.3 * lengthRatio;
    }
    return lengthRatio;
  }
  config::FeatureConfig config_;
  int maxLagSamples_ = 0;
  double maxLagSeconds_ = 0, confidenceThresh_ = 0;
};
inline std::shared_ptr<AlignmentExtractor> NewAlignmentExtractorWithMaxLag(const config::FeatureConfig* featureConf,
                                                                            const config::AlignmentConfig* alignmentConf,
                                                                            double maxLagSeconds) {
  return std::make_shared<AlignmentExtractor>(featureConf, alignmentConf, maxLagSeconds);
}

}  // namespace extractors

// ----------------------------------------------------------------------------------------------
namespace fingerprint {  // fingerprint/

struct FingerprintConfig {  // fingerprint.go:29-35
  int WindowSize = 0, HopSize = 0;
  bool EnableContentDetect = false;
  std::shared_ptr<config::FeatureConfig> FeatureConfig;
  std::shared_ptr<config::ContentAwareConfig> ContentConfig;
};

inline std::shared_ptr<FingerprintConfig> DefaultFingerprintConfig() {  // fingerprint.go:70-101
  auto c = std::make_shared<FingerprintConfig>();
  c->WindowSize = 2048, c->HopSize = 512, c->EnableContentDetect = true;
  auto f = std::make_shared<config::FeatureConfig>();
  f->EnableMFCC = f->EnableChroma = f->EnableSpectralContrast = f->EnableTemporalFeatures = true;
  f->MFCCCoefficients = 13, f->ChromaBins = 12, f->WindowType = "hann";
  f->SimilarityWeights = {{"mfcc", 0.40}, {"spectral", 0.25}, {"chroma", 0.20}, {"temporal", 0.15}};
  c->FeatureConfig = f;
  auto cc = std::make_shared<config::ContentAwareConfig>();
  cc->EnableContentDetection = true, cc->DefaultContentType = config::ContentUnknown, cc->AutoDetectThreshold = 2.0;
  c->ContentConfig = cc;
  return c;
}

struct FeatureSettings {  // content_config.go:13-24
  bool EnableMFCC, EnableChroma, EnableSpectralContrast, EnableHarmonicFeatures, EnableSpeechFeatures, EnableTemporalFeatures;
  std::map<std::string, double> SimilarityWeights;
};

class ContentAwareConfigManager {  // content_config.go:36-103
 public:
  explicit ContentAwareConfigManager(std::shared_ptr<FingerprintConfig> base)
      : base_(base ? base : DefaultFingerprintConfig()) {}
  // GetGenerationConfig: copies the base config and REPLACES FeatureConfig with one built from the
  // content table — SampleRate is never set (F2), WindowSize/HopSize come from base FeatureConfig (F4).
  std::shared_ptr<FingerprintConfig> GetGenerationConfig(const config::ContentType& ct) const {
    auto g = std::make_shared<FingerprintConfig>(*base_);
    const FeatureSettings s = Settings(ct);
    auto f = std::make_shared<config::FeatureConfig>();
    f->EnableMFCC = s.EnableMFCC, f->EnableChroma = s.EnableChroma, f->EnableSpectralContrast = s.EnableSpectralContrast;
    f->EnableHarmonicFeatures = s.EnableHarmonicFeatures, f->EnableSpeechFeatures = s.EnableSpeechFeatures;
    f->EnableTemporalFeatures = s.EnableTemporalFeatures;
    f->MFCCCoefficients = 13, f->ChromaBins = 12, f->SimilarityWeights = s.SimilarityWeights, f->WindowType = "hann";
    f->WindowSize = base_->FeatureConfig ? base_->FeatureConfig->WindowSize : 0;
    f->HopSize = base_->FeatureConfig ? base_->FeatureConfig->HopSize : 0;
    g->FeatureConfig = f;
    return g;
  }
  static FeatureSettings Settings(const config::ContentType& ct) {  // getContentConfigs :106-278
    if (ct == config::ContentMusic)
      return {true, true, true, true, false, false, {{"mfcc", 0.35}, {"chroma", 0.30}, {"harmonic", 0.20}, {"spectral", 0.15}}};
    if (ct == config::ContentNews)
      return {true, false, true, false, true, true, {{"mfcc", 0.50}, {"speech", 0.25}, {"spectral", 0.15}, {"temporal", 0.10}}};
    if (ct == config::ContentTalk)
      return {true, false, true, false, true, true, {{"mfcc", 0.45}, {"speech", 0.30}, {"spectral", 0.15}, {"temporal", 0.10}}};
    if (ct == config::ContentMixed)
      return {true, true, true, true, true, true,
              {{"mfcc", 0.30}, {"spectral", 0.20}, {"temporal", 0.20}, {"chroma", 0.15}, {"speech", 0.15}}};
    // ContentSports has no entry -> falls back to ContentUnknown (content_config.go:59-63)
    return {true, true, true, false, false, true, {{"mfcc", 0.40}, {"spectral", 0.25}, {"chroma", 0.20}, {"temporal", 0.15}}};
  }

 private:
  std::shared_ptr<FingerprintConfig> base_;
};

struct AudioFingerprint {  // fingerprint.go:15-26 (ID / Timestamp are time-derived: excluded from parity)
  std::string ID, StreamURL;
  config::ContentType ContentType;
  double DurationSeconds = 0;
  int SampleRate = 0, HopSize = 0, Channels = 0;
  std::shared_ptr<extractors::ExtractedFeatures> Features;
  std::map<std::string, double> FeatureWeights;  // Metadata["feature_weights"]
  std::map<std::string, std::string> Metadata;
};

class FingerprintGenerator {
 public:
  explicit FingerprintGenerator(std::shared_ptr<FingerprintConfig> cfg)
      : config_(cfg ? cfg : DefaultFingerprintConfig()), contentManager_(config_) {}
  Result<AudioFingerprint> GenerateFingerprint(const transcode::AudioData* audioData) const {  // fingerprint.go:137-236
    if (!audioData) return Err<AudioFingerprint>("audio data cannot be nil");  // :139
    // The reference dereferences Metadata unconditionally (:155, F8); the drop-in reports it instead of panicking.
    if (!audioData->Metadata) return Err<AudioFingerprint>("audio metadata cannot be nil");
    config::ContentType contentType = config::ToContentType(audioData->Metadata->ContentType);
    // Content auto-detection (content_detector.go) stays host-side Go and is out of this path's scope:
    // an unknown type keeps the ContentUnknown settings, exactly what the reference does when detection is off.
    auto generationConfig = contentManager_.GetGenerationConfig(contentType);
    auto ex = extractors::NewFeatureExtractorFactory().CreateExtractor(contentType, *generationConfig->FeatureConfig);  // :171 (copy)
    const int windowSize = generationConfig->WindowSize, hopSize = generationConfig->HopSize;
    generationConfig->FeatureConfig->WindowSize = windowSize;  // :177-181 — after the copy: the extractor never sees it (F4)
    generationConfig->FeatureConfig->HopSize = hopSize;
    // ComputeSTFTWithWindow's argument checks (analyzers/spectral.go:387-411); the transform itself is fused
    // into the extractor's kernels, so only the header of the SpectrogramResult is produced here.
    if (audioData->PCM.empty()) return Err<AudioFingerprint>("empty signal");
    if (windowSize <= 0) return Err<AudioFingerprint>("window size must be positive");
    if (hopSize <= 0) return Err<AudioFingerprint>("hop size must be positive");
    const int64_t numFrames = ((int64_t)audioData->PCM.size() - windowSize) / hopSize + 1;
    if (numFrames <= 0) return Err<AudioFingerprint>("signal too short for given window size and hop size");
    analyzers::SpectrogramResult sp;
    sp.TimeFrames = (int)numFrames, sp.FreqBins = windowSize / 2 + 1, sp.SampleRate = audioData->SampleRate;
    sp.WindowSize = windowSize, sp.HopSize = hopSize, sp.WindowType = generationConfig->FeatureConfig->WindowType;
    auto features = ex->ExtractFeatures(&sp, audioData->PCM, audioData->SampleRate);  // :207
    if (!features.ok()) return Err<AudioFingerprint>(features.err);
    auto fp = std::make_shared<AudioFingerprint>();
    fp->StreamURL = audioData->Metadata->URL;
    fp->ContentType = contentType;
    fp->DurationSeconds = audioData->SampleRate > 0  // utils.go:13-19
                              ? (double)audioData->PCM.size() / (double)(audioData->SampleRate * audioData->Channels)
                              : 0.0;
    fp->SampleRate = audioData->SampleRate;
    fp->HopSize = config_->FeatureConfig ? config_->FeatureConfig->HopSize : 0;  // :221 (the BASE FeatureConfig's)
    fp->Channels = audioData->Channels;
    fp->Features = features.value;
    fp->FeatureWeights = ex->GetFeatureWeights();  // utils.go:31-33
    fp->Metadata["extractor_name"] = ex->GetName();
    fp->ID = "gpu-" + std::to_string(audioData->PCM.size()) + "-" + std::to_string(audioData->SampleRate);
    return Result<AudioFingerprint>{fp, ""};
  }

 private:
  std::shared_ptr<FingerprintConfig> config_;
  ContentAwareConfigManager contentManager_;
};
inline std::shared_ptr<FingerprintGenerator> NewFingerprintGenerator(std::shared_ptr<FingerprintConfig> cfg) {
  return std::make_shared<FingerprintGenerator>(cfg);
}

struct SimilarityResult {  // comparison.go:28-39
  double OverallSimilarity = 0, FeatureSimilarity = 0, Confidence = 0;
  bool ContentTypeMatch = false;
  std::map<std::string, double> FeatureDistances;
};

class FingerprintComparator {
 public:
  explicit FingerprintComparator(const config::ComparisonConfig* cfg) : config_(cfg ? *cfg : config::DefaultComparisonConfig()) {}
  Result<SimilarityResult> Compare(const AudioFingerprint* fp1, const AudioFingerprint* fp2) const {  // comparison.go:133-194
    if (!fp1 || !fp2) return Err<SimilarityResult>("fingerprints cannot be nil");  // :135
    auto out = std::make_shared<SimilarityResult>();
    out->ContentTypeMatch = fp1->ContentType == fp2->ContentType;
    if (config_.EnableContentFilter && !out->ContentTypeMatch) {  // :160-166
      out->Confidence = 0.25;
      return Result<SimilarityResult>{out, ""};
    }
    std::string rerr;
    sonar_ctx* ctx = Runtime::Ctx(&rerr);
    if (!ctx) return Err<SimilarityResult>(rerr);
    sonar_cmp_result r;
    std::memset(&r, 0, sizeof(r));
    if (fp1->Features && fp2->Features) {
      std::vector<double> m1, m2;
      sonar_cmp_features f1 = Flatten(*fp1, m1), f2 = Flatten(*fp2, m2);
      const auto w = EffectiveWeights(*fp1);
      sonar_cmp_weights cw;
      const char* order[7] = {"mfcc", "spectral", "chroma", "temporal", "speech", "harmonic", "energy"};
      for (int i = 0; i < 7; i++) {
        auto it = w.find(order[i]);
        cw.w[i] = it == w.end() ? 0.0 : it->second;  // Go map lookup of a missing key yields 0
      }
      if (sonar_compare_f64(ctx, &f1, &f2, &cw, 0, &r) != SONAR_OK) return Err<SimilarityResult>(sonar_last_error());
    } else {  // calculateFeatureSimilarity's error is swallowed into similarity 0 (:170-173)
      r.confidence = 0.5 + (out->ContentTypeMatch ? 0.1 : 0.0);
      r.dist_mfcc = r.dist_spectral = r.dist_temporal = r.dist_harmonic = std::nan("");
    }
    out->FeatureSimilarity = r.feature_similarity;
    out->OverallSimilarity = r.overall_similarity;
    out->Confidence = r.confidence;
    if (!std::isnan(r.dist_mfcc)) out->FeatureDistances["mfcc"] = r.dist_mfcc;
    if (!std::isnan(r.dist_spectral)) out->FeatureDistances["spectral"] = r.dist_spectral;
    if (!std::isnan(r.dist_temporal)) out->FeatureDistances["temporal"] = r.dist_temporal;
    if (!std::isnan(r.dist_harmonic)) out->FeatureDistances["harmonic"] = r.dist_harmonic;
    return Result<SimilarityResult>{out, ""};
  }
  // comparison.go:1107-1151: every non-nil candidate that is not the query itself (same ID) is compared; failures
  // are skipped
  std::vector<std::shared_ptr<SimilarityResult>> BatchCompare(const AudioFingerprint* query,
                                                              const std::vector<const AudioFingerprint*>& candidates,
                                                              std::string* err = nullptr) const {
    std::vector<std::shared_ptr<SimilarityResult>> results;
    if (!query) {
      if (err) *err = "query fingerprint cannot be nil";
      return results;
    }
    for (const AudioFingerprint* c : candidates) {
      if (!c || c->ID == query->ID) continue;
      auto r = Compare(query, c);
      if (r.ok()) results.push_back(r.value);
    }
    return results;
  }
  struct Match {  // comparison.go:52-58
    const AudioFingerprint* Fingerprint = nullptr;
    std::shared_ptr<SimilarityResult> Similarity;
    int Rank = 0;
    std::string MatchType;
  };
  static std::string ClassifyMatch(const SimilarityResult& s) {  // comparison.go:1040-1052
    if (s.OverallSimilarity >= 0.95) return "exact";
    if (s.OverallSimilarity >= 0.85) return "very_similar";
    if (s.OverallSimilarity >= 0.75) return "similar";
    if (s.OverallSimilarity >= 0.6) return "somewhat_similar";
    return "weak";
  }
  // comparison.go:197-263: compare, keep >= SimilarityThreshold, sort descending, cut to MaxCandidates, rank from 1
  std::vector<Match> FindBestMatches(const AudioFingerprint* query, const std::vector<const AudioFingerprint*>& candidates,
                                     std::string* err = nullptr) const {
    std::vector<Match> matches;
    if (!query) {
      if (err) *err = "query fingerprint cannot be nil";
      return matches;
    }
    for (const AudioFingerprint* c : candidates) {
      if (!c || c->ID == query->ID) continue;
      auto r = Compare(query, c);
      if (!r.ok()) continue;
      if (r->OverallSimilarity >= config_.SimilarityThreshold) matches.push_back(Match{c, r.value, 0, ClassifyMatch(*r.value)});
    }
    std::stable_sort(matches.begin(), matches.end(), [](const Match& a, const Match& b) {
      return a.Similarity->OverallSimilarity > b.Similarity->OverallSimilarity;
    });
    if ((int)matches.size() > config_.MaxCandidates) matches.resize((size_t)config_.MaxCandidates);
    for (size_t i = 0; i < matches.size(); i++) matches[i].Rank = (int)i + 1;
    return matches;
  }
  static std::map<std::string, double> EffectiveWeights(const AudioFingerprint& fp) {  // comparison.go:1055-1104
    if (!fp.FeatureWeights.empty()) return fp.FeatureWeights;
    if (fp.ContentType == config::ContentNews || fp.ContentType == config::ContentTalk)
      return {{"mfcc", 0.50}, {"spectral", 0.25}, {"temporal", 0.15}, {"speech", 0.10}, {"chroma", 0.05}, {"harmonic", 0.05}, {"energy", 0.10}};
    if (fp.ContentType == config::ContentMusic)
      return {{"mfcc", 0.30}, {"chroma", 0.25}, {"spectral", 0.20}, {"harmonic", 0.15}, {"temporal", 0.10}, {"speech", 0.05}, {"energy", 0.10}};
    if (fp.ContentType == config::ContentSports)
      return {{"energy", 0.30}, {"temporal", 0.25}, {"mfcc", 0.25}, {"spectral", 0.20}, {"speech", 0.10}, {"chroma", 0.05}, {"harmonic", 0.05}};
    return {{"mfcc", 0.35}, {"spectral", 0.25}, {"temporal", 0.20}, {"energy", 0.15}, {"chroma", 0.10}, {"speech", 0.10}, {"harmonic", 0.10}};
  }

 private:
  static sonar_cmp_features Flatten(const AudioFingerprint& fp, std::vector<double>& mfcc) {
    sonar_cmp_features f;
    std::memset(&f, 0, sizeof(f));
    const auto& x = *fp.Features;
    if (!x.MFCC.empty()) {
      const size_t d = x.MFCC[0].size();
      mfcc.resize(x.MFCC.size() * d);
      for (size_t t = 0; t < x.MFCC.size(); t++) std::copy(x.MFCC[t].begin(), x.MFCC[t].end(), mfcc.begin() + t * d);
      f.mfcc = mfcc.data(), f.mfcc_frames = (int64_t)x.MFCC.size(), f.mfcc_dim = (int32_t)d;
    }
    if (x.SpectralFeatures) {
      f.has_spectral = 1;
      f.spectral_centroid = x.SpectralFeatures->SpectralCentroid.data(), f.n_centroid = (int64_t)x.SpectralFeatures->SpectralCentroid.size();
      f.spectral_rolloff = x.SpectralFeatures->SpectralRolloff.data(), f.n_rolloff = (int64_t)x.SpectralFeatures->SpectralRolloff.size();
      f.spectral_flux = x.SpectralFeatures->SpectralFlux.data(), f.n_flux = (int64_t)x.SpectralFeatures->SpectralFlux.size();
    }
    if (x.TemporalFeatures) {  // comparison.go:688-718
      f.has_temporal = 1;
      f.rms_energy = x.TemporalFeatures->RMSEnergy.data(), f.n_rms = (int64_t)x.TemporalFeatures->RMSEnergy.size();
      f.dynamic_range = x.TemporalFeatures->DynamicRange, f.silence_ratio = x.TemporalFeatures->SilenceRatio;
      f.onset_density = x.TemporalFeatures->OnsetDensity;
    }
    if (x.HarmonicFeatures) {
      f.has_harmonic = 1;
      f.harmonic_ratio = x.HarmonicFeatures->HarmonicRatio.data(), f.n_harmonic_ratio = (int64_t)x.HarmonicFeatures->HarmonicRatio.size();
      f.pitch_estimate = x.HarmonicFeatures->PitchEstimate.data(), f.n_pitch = (int64_t)x.HarmonicFeatures->PitchEstimate.size();
    }
    return f;
  }
  config::ComparisonConfig config_;
};
inline std::shared_ptr<FingerprintComparator> NewFingerprintComparator(const config::ComparisonConfig* cfg) {
  return std::make_shared<FingerprintComparator>(cfg);
}

}  // namespace fingerprint
}  // namespace sonido
