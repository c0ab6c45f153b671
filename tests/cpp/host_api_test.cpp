// Exercises the host-side mirror of the Go API (sonido-sonar_b200/host/sonar_host.hpp) the way a user of
// the reference would: GenerateFingerprint -> ExtractAlignmentFeatures -> Compare, plus the reference's
// error strings and the config-plumbing quirks F1-F4.  Linked against either implementation
// of include/sonar.h: the CUDA product library (GPU test) or the CPU oracle (host-logic test).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../sonido-sonar_b200/host/sonar_host.hpp"

using namespace sonido;

static int failures = 0;
#define CHECK(cond)                                                          \
  do {                                                                       \
    if (!(cond)) {                                                           \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);            \
      failures++;                                                            \
    }                                                                        \
  } while (0)

static std::vector<double> make_stream(int n, int sr, int shift, unsigned seed) {
  // slow envelope x noise (the C2-style process of SURVEY §8d), sample k is process sample k + shift
  std::mt19937_64 rng(12345);
  std::normal_distribution<double> g(0.0, 1.0);
  std::vector<double> base((size_t)n + 200000);
  for (auto& v : base) v = g(rng);
  std::mt19937_64 rng2(seed);
  std::vector<double> x(n);
  for (int k = 0; k < n; k++) {
    const double t = (double)(k + shift) / sr;
    const double env = 0.3 + 0.25 * std::sin(2 * M_PI * 0.37 * t) + 0.2 * std::sin(2 * M_PI * 1.9 * t);
    x[k] = env * base[(size_t)(k + shift)] + 0.01 * g(rng2);
  }
  return x;
}

int main() {
  const int sr = 44100;
  // ---- canonical harness config: both config levels carry WindowSize/HopSize (F4)
  auto cfg = std::make_shared<fingerprint::FingerprintConfig>();
  cfg->WindowSize = 1024, cfg->HopSize = 256, cfg->EnableContentDetect = false;
  cfg->FeatureConfig = std::make_shared<config::FeatureConfig>();
  cfg->FeatureConfig->WindowSize = 1024, cfg->FeatureConfig->HopSize = 256, cfg->FeatureConfig->SampleRate = sr;
  cfg->FeatureConfig->MFCCCoefficients = 13, cfg->FeatureConfig->WindowType = "hann";
  cfg->ContentConfig = std::make_shared<config::ContentAwareConfig>();
  auto gen = fingerprint::NewFingerprintGenerator(cfg);

  // ---- error conventions (fingerprint.go:139, analyzers/spectral.go:387-411)
  CHECK(gen->GenerateFingerprint(nullptr).err == "audio data cannot be nil");
  transcode::AudioData shorty;
  shorty.SampleRate = sr, shorty.PCM.assign(100, 0.1);
  shorty.Metadata = std::make_shared<transcode::StreamMetadata>();
  shorty.Metadata->ContentType = "music";
  CHECK(gen->GenerateFingerprint(&shorty).err == "signal too short for given window size and hop size");
  shorty.PCM.clear();
  CHECK(gen->GenerateFingerprint(&shorty).err == "empty signal");

  // ---- two streams of the same process, the CDN copy 1.25 s later
  const int n = 20 * sr, off = (int)(1.25 * sr);
  transcode::AudioData q, r;
  q.SampleRate = r.SampleRate = sr;
  q.PCM = make_stream(n, sr, off, 1), r.PCM = make_stream(n, sr, 0, 2);
  q.Metadata = std::make_shared<transcode::StreamMetadata>(), r.Metadata = std::make_shared<transcode::StreamMetadata>();
  q.Metadata->ContentType = r.Metadata->ContentType = "music";
  q.Metadata->URL = "source", r.Metadata->URL = "cdn";
  auto fq = gen->GenerateFingerprint(&q), fr = gen->GenerateFingerprint(&r);
  CHECK(fq.ok() && fr.ok());
  if (!fq.ok() || !fr.ok()) {
    std::printf("error: %s %s\n", fq.err.c_str(), fr.err.c_str());
    return 1;
  }
  const int T = (n - 1024) / 256 + 1;
  CHECK((int)fq->Features->MFCC.size() == T && fq->Features->MFCC[0].size() == 13);
  CHECK(fq->ContentType == "music" && fq->HopSize == 256 && fq->StreamURL == "source");
  CHECK(fq->Metadata["extractor_name"] == "SpeechFeatureExtractor");  // F1: music gets the speech extractor
  CHECK(fq->FeatureWeights.count("chroma") == 1);                     // music SimilarityWeights travel with the fingerprint
  // F2/F3: the extractor's algorithms are built with sampleRate 0 -> degenerate MFCC and zeroed centroid
  CHECK(std::fabs(fq->Features->MFCC[T / 2][0] - (-117.40926320884498)) < 1e-3);
  CHECK(fq->Features->SpectralFeatures->SpectralCentroid[T / 2] == 0.0);
  CHECK(fq->Features->SpectralFeatures->SpectralFlatness[T / 2] > 0.0);
  CHECK((int)fq->Features->EnergyFeatures->ShortTimeEnergy.size() == T);
  CHECK(fq->Features->HarmonicFeatures->InharmonicityRatio[3] == 1.0);

  // ---- alignment (extractors/alignment.go:99-219)
  config::FeatureConfig fc;
  fc.SampleRate = sr, fc.WindowSize = 1024, fc.HopSize = 256;
  config::AlignmentConfig ac = config::DefaultAlignmentConfig();
  auto ae = extractors::NewAlignmentExtractorWithMaxLag(&fc, &ac, 5.0);
  CHECK(ae->ExtractAlignmentFeatures(nullptr, fr.value->Features.get(), q.PCM, r.PCM, sr).err == "feature sets cannot be nil");
  auto al = ae->ExtractAlignmentFeatures(fq->Features.get(), fr->Features.get(), q.PCM, r.PCM, sr);
  CHECK(al.ok() && al->BestAlignment && al->BestAlignment->Success);
  if (al.ok() && al->BestAlignment) {
    std::printf("offset %.4f s (true 1.25), lag %d frames, confidence %.3f, similarity %.3f, method %s\n",
                al->TemporalOffset, al->BestAlignment->Result->CrossCorrResult->PeakLag, al->OffsetConfidence,
                al->AlignmentSimilarity, al->Method.c_str());
    CHECK(std::fabs(al->TemporalOffset - 1.25) < 256.0 / sr);  // positive: the CDN is later
    CHECK(al->Method == "corr_energy" && al->CorrAlignment == al->BestAlignment);
    CHECK(al->BestAlignment->Result->CrossCorrResult->Correlations.size() == (size_t)(2 * (5 * sr / 256) + 1));
    CHECK(al->FeatureSimilarity.count("corr_energy") == 1 && al->TimeStretch == 1.0);
    CHECK(std::fabs(al->QueryLength - 20.0) < 1e-12);
  }

  // ---- F4: without FeatureConfig.WindowSize/HopSize the extractor computes no energies -> nothing to align
  auto cfg4 = std::make_shared<fingerprint::FingerprintConfig>(*cfg);
  cfg4->FeatureConfig = std::make_shared<config::FeatureConfig>(*cfg->FeatureConfig);
  cfg4->FeatureConfig->WindowSize = 0, cfg4->FeatureConfig->HopSize = 0;
  auto gen4 = fingerprint::NewFingerprintGenerator(cfg4);
  auto f4 = gen4->GenerateFingerprint(&q);
  CHECK(f4.ok() && f4->Features->EnergyFeatures->ShortTimeEnergy.empty());
  if (f4.ok()) {
    auto al4 = ae->ExtractAlignmentFeatures(f4->Features.get(), fr->Features.get(), q.PCM, r.PCM, sr);
    CHECK(al4.ok() && !al4->BestAlignment && al4->Method.empty());
  }

  // ---- news content: temporal group on, speech group skipped, extractor reports "news"
  {
    transcode::AudioData nw = q;
    nw.Metadata = std::make_shared<transcode::StreamMetadata>();
    nw.Metadata->ContentType = "news";
    auto fn = gen->GenerateFingerprint(&nw);
    CHECK(fn.ok() && fn->ContentType == "news" && fn->Features->TemporalFeatures);
    if (fn.ok() && fn->Features->TemporalFeatures) {
      const auto& tfe = *fn->Features->TemporalFeatures;
      CHECK(tfe.RMSEnergy == fn->Features->EnergyFeatures->ShortTimeEnergy);
      CHECK(tfe.SilenceRatio > 0.09 && tfe.SilenceRatio < 0.12);  // sorted[T/10] threshold -> ~10 % + ties
      CHECK(tfe.PeakAmplitude > tfe.AverageAmplitude && tfe.AverageAmplitude > 0);
      CHECK(tfe.EnvelopeShape.size() == (size_t)((n - 512) / 256 + 1));
      CHECK(fn->Features->ExtractionMetadata["content_subtype"] == "news");
      CHECK(fn->Features->ExtractionMetadata.count("speech_features") == 1);
    }
  }

  // ---- compare (comparison.go:133-194)
  config::ComparisonConfig cc = config::DefaultComparisonConfig();
  auto cmp = fingerprint::NewFingerprintComparator(&cc);
  CHECK(cmp->Compare(nullptr, fr.value.get()).err == "fingerprints cannot be nil");
  auto same = cmp->Compare(fq.value.get(), fq.value.get());
  auto diff = cmp->Compare(fq.value.get(), fr.value.get());
  CHECK(same.ok() && diff.ok());
  if (same.ok() && diff.ok()) {
    std::printf("compare: same %.6f (conf %.2f), source-vs-cdn %.6f\n", same->OverallSimilarity, same->Confidence,
                diff->OverallSimilarity);
    CHECK(same->ContentTypeMatch && same->FeatureDistances.count("mfcc") && same->FeatureDistances.count("spectral"));
    CHECK(std::fabs(same->FeatureDistances["mfcc"]) < 1e-9);  // parity-mode MFCC rows are constant -> cosine 1
    CHECK(diff->OverallSimilarity > 0.5 && diff->OverallSimilarity <= 1.0 + 1e-12);
  }
  {  // BatchCompare / FindBestMatches (comparison.go:1107-1151, 197-263)
    fq.value->ID = "q";
    fr.value->ID = "r";
    fingerprint::AudioFingerprint twin = *fq.value;
    twin.ID = "twin";
    std::vector<const fingerprint::AudioFingerprint*> cands = {fr.value.get(), nullptr, fq.value.get(), &twin};
    std::string berr;
    CHECK(cmp->BatchCompare(nullptr, cands, &berr).empty() && berr == "query fingerprint cannot be nil");
    auto batch = cmp->BatchCompare(fq.value.get(), cands);
    CHECK(batch.size() == 2);  // nil and the query itself (same ID) are skipped
    config::ComparisonConfig loose = cc;
    loose.SimilarityThreshold = 0.0;
    auto best = fingerprint::NewFingerprintComparator(&loose)->FindBestMatches(fq.value.get(), cands);
    CHECK(best.size() == 2 && best[0].Fingerprint == &twin && best[0].Rank == 1 && best[1].Rank == 2);
    CHECK(best[0].Similarity->OverallSimilarity >= best[1].Similarity->OverallSimilarity);
    CHECK(best[0].MatchType == fingerprint::FingerprintComparator::ClassifyMatch(*best[0].Similarity));
    loose.MaxCandidates = 1;
    CHECK(fingerprint::NewFingerprintComparator(&loose)->FindBestMatches(fq.value.get(), cands).size() == 1);
  }
  fr.value->ContentType = "news";
  cc.EnableContentFilter = true;
  auto filt = fingerprint::NewFingerprintComparator(&cc)->Compare(fq.value.get(), fr.value.get());
  CHECK(filt.ok() && filt->OverallSimilarity == 0.0 && filt->Confidence == 0.25 && !filt->ContentTypeMatch);

  {  // ComputeSTFTStreaming / STFTStreamer.ProcessChunk (analyzers/spectral.go:289-374)
    CHECK(!analyzers::ComputeSTFTStreaming(0, 256, "hann").ok());
    auto st = analyzers::ComputeSTFTStreaming(1024, 256, "hann");
    CHECK(st.ok());
    std::string serr = "x";
    CHECK(st->ProcessChunk({}, &serr).empty() && serr.empty());
    std::vector<double> sig(5000);
    for (size_t i = 0; i < sig.size(); i++) sig[i] = std::sin(0.05 * (double)i) + 0.25 * std::sin(0.31 * (double)i);
    auto f1 = st->ProcessChunk(std::vector<double>(sig.begin(), sig.begin() + 1000), &serr);
    CHECK(f1.empty() && serr.empty() && st->BufferedSamples() == 1000);
    auto f2 = st->ProcessChunk(std::vector<double>(sig.begin() + 1000, sig.end()), &serr);
    CHECK(serr.empty() && f2.size() == (5000 - 1024) / 256 + 1);  // 16 frames
    CHECK(st->BufferedSamples() == 5000 - 16 * 256);
    std::vector<double> mag(16 * 513);
    CHECK(sonar_stft_f64(sonido::Runtime::Ctx(), sig.data(), (int64_t)sig.size(), 1024, 256, SONAR_WINDOW_HANN, mag.data(),
                         nullptr, nullptr) == SONAR_OK);
    bool same = f2.size() == 16;
    for (size_t t = 0; same && t < 16; t++)
      for (size_t k = 0; k < 513; k++) same = same && f2[t]->Magnitude[k] == mag[t * 513 + k];
    CHECK(same);
    CHECK(f2[3]->Complex.size() == 513 && std::fabs(std::abs(f2[3]->Complex[40]) - f2[3]->Magnitude[40]) <=
                                              1e-5 * (1.0 + f2[3]->Magnitude[40]));
  }

  std::printf("%s backend=%s failures=%d\n", failures ? "FAILED" : "OK", sonar_backend(), failures);
  return failures ? 1 : 0;
}
