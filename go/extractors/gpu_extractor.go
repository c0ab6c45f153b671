// Drop this file into fingerprint/extractors/ of RyanBlaney/sonido-sonar (package extractors).
// It implements the reference's plug-in interface FeatureExtractor (feature_extractor.go:10-15) on top of
// libsonar.so and is returned by FeatureExtractorFactory.CreateExtractor in place of
// NewSpeechFeatureExtractor (feature_extractor.go:42-61; every content type ends up there, SURVEY F1).
//
// NOT COMPILED HERE (no Go toolchain in the build image); the C++ mirror of this file is
// sonido-sonar_b200/host/sonar_host.hpp::extractors::SpeechFeatureExtractor, which is tested.
package extractors

import (
	"fmt"

	"github.com/RyanBlaney/sonido-sonar/fingerprint/analyzers"
	"github.com/RyanBlaney/sonido-sonar/fingerprint/config"
	"github.com/RyanBlaney/sonido-sonar/sonargpu"
)

// GPUSpeechFeatureExtractor keeps SpeechFeatureExtractor's name, weights and content type.
type GPUSpeechFeatureExtractor struct {
	config *config.FeatureConfig
	isNews bool
}

func NewGPUSpeechFeatureExtractor(cfg *config.FeatureConfig, isNews bool) *GPUSpeechFeatureExtractor {
	return &GPUSpeechFeatureExtractor{config: cfg, isNews: isNews}
}

func (s *GPUSpeechFeatureExtractor) GetName() string { return "SpeechFeatureExtractor" }

func (s *GPUSpeechFeatureExtractor) GetContentType() config.ContentType {
	if s.isNews {
		return config.ContentNews
	}
	return config.ContentTalk
}

func (s *GPUSpeechFeatureExtractor) GetFeatureWeights() map[string]float64 {
	if s.config.SimilarityWeights != nil {
		return s.config.SimilarityWeights
	}
	w := map[string]float64{"mfcc": 0.40, "speech": 0.35, "spectral": 0.15, "temporal": 0.10}
	if s.isNews {
		w["speech"], w["mfcc"] = 0.40, 0.35
	}
	return w
}

// ExtractFeatures: the spectrogram argument only needs its header (TimeFrames, FreqBins, WindowSize,
// HopSize): the STFT is fused into the GPU kernels, so GenerateFingerprint may skip the CPU STFT
// (fingerprint.go:190) and pass &analyzers.SpectrogramResult{WindowSize: w, HopSize: h, ...}.
func (s *GPUSpeechFeatureExtractor) ExtractFeatures(spectrogram *analyzers.SpectrogramResult, pcm []float64, sampleRate int) (*ExtractedFeatures, error) {
	if spectrogram == nil {
		return nil, fmt.Errorf("spectrogram cannot be nil")
	}
	if len(pcm) == 0 {
		return nil, fmt.Errorf("PCM data cannot be empty")
	}
	if sampleRate <= 0 {
		return nil, fmt.Errorf("sample rate must be positive")
	}
	fp, err := sonargpu.GenerateFingerprint(pcm, sonargpu.FpParams{
		WindowSize: spectrogram.WindowSize, HopSize: spectrogram.HopSize, WindowType: 0, // hann
		AlgoSampleRate: s.config.SampleRate, // what NewSpeechFeatureExtractor hands every algorithm (0 via GenerateFingerprint)
		CallSampleRate: sampleRate,
		EnergyFrame:    s.config.WindowSize, EnergyHop: s.config.HopSize, // temporal.NewEnergy(config.WindowSize, config.HopSize, ...)
		MFCCCoefficients: s.config.MFCCCoefficients, EnableMFCC: s.config.EnableMFCC,
	})
	if err != nil {
		return nil, fmt.Errorf("spectral feature extraction failed: %w", err)
	}
	f := &ExtractedFeatures{ExtractionMetadata: map[string]any{}}
	if s.config.EnableMFCC {
		f.MFCC = make([][]float64, fp.Frames) // re-slice one backing array: no per-frame allocation
		for t := range f.MFCC {
			f.MFCC[t] = fp.MFCC[t*fp.NMFCC : (t+1)*fp.NMFCC : (t+1)*fp.NMFCC]
		}
	}
	f.SpectralFeatures = &SpectralFeatures{
		SpectralCentroid: fp.Centroid, SpectralRolloff: fp.Rolloff, SpectralBandwidth: fp.Bandwidth,
		SpectralFlatness: fp.Flatness, SpectralCrest: fp.Crest, SpectralSlope: fp.Slope,
		SpectralFlux: fp.Flux, ZeroCrossingRate: fp.ZCR,
	}
	f.EnergyFeatures = &EnergyFeatures{
		ShortTimeEnergy: fp.ShortTimeEnergy, EnergyEntropy: fp.EnergyEntropy, LowEnergyRatio: fp.LowEnergyRatio,
		HighEnergyRatio: fp.HighEnergyRatio, EnergyVariance: fp.EnergyVariance, LoudnessRange: fp.LoudnessRange,
	}
	f.HarmonicFeatures = &HarmonicFeatures{
		PitchEstimate: fp.Pitch, PitchConfidence: fp.PitchConfidence, VoicingStrength: fp.Voicing,
		HarmonicRatio: fp.HarmonicRatio, InharmonicityRatio: fp.Inharmonicity, TonalCentroid: fp.TonalCentroid,
	}
	f.ExtractionMetadata["extractor_type"] = "speech"
	f.ExtractionMetadata["sample_rate"] = sampleRate
	f.ExtractionMetadata["spectrogram_frames"] = spectrogram.TimeFrames
	f.ExtractionMetadata["backend"] = "cuda-sm100a"
	return f, nil
}
