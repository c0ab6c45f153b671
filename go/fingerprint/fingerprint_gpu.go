// Drop-in replacement for the compute section of FingerprintGenerator.GenerateFingerprint
// (fingerprint/fingerprint.go:190-207 of RyanBlaney/sonido-sonar), see INTEGRATION.md section 4.2.
//
// NOT COMPILED IN THIS REPOSITORY (no Go toolchain in the image).  To apply: copy this file into the reference's
// `fingerprint` package, then in GenerateFingerprint replace the block from
//
//	spectrogram, err := fg.spectralAnalyzer.ComputeSTFTWithWindow(...)     // :190
//	...
//	features, err := extractor.ExtractFeatures(spectrogram, audioData.PCM, audioData.SampleRate)   // :207
//
// by
//
//	features, err := fg.extractFeaturesGPU(audioData, generationConfig, contentType)
//
// Everything before (content detection, GetGenerationConfig, CreateExtractor -- kept for GetFeatureWeights in
// addMetadata) and after (AudioFingerprint assembly, addMetadata, generateID) stays as it is, and so do all signatures.
package fingerprint

import (
	"fmt"

	"github.com/RyanBlaney/sonido-sonar/fingerprint/analyzers"
	"github.com/RyanBlaney/sonido-sonar/fingerprint/config"
	"github.com/RyanBlaney/sonido-sonar/fingerprint/extractors"
	"github.com/RyanBlaney/sonido-sonar/sonargpu"
	"github.com/RyanBlaney/sonido-sonar/transcode"
)

var windowIndex = map[analyzers.WindowType]int{
	analyzers.WindowHann: 0, analyzers.WindowHamming: 1, analyzers.WindowBlackman: 2, analyzers.WindowBlackmanHarris: 3,
	analyzers.WindowKaiser: 4, analyzers.WindowTukey: 5, analyzers.WindowRectangular: 6, analyzers.WindowBartlett: 7,
	analyzers.WindowWelch: 8,
}

// extractFeaturesGPU = ComputeSTFTWithWindow + SpeechFeatureExtractor.ExtractFeatures on the B200 library.  It passes
// the library exactly what the reference's objects were constructed with, including the two plumbing quirks the CPU
// path has (SURVEY.md section 0): the extractor's algorithms see FeatureConfig.SampleRate as buildFeatureConfig left it
// (0 through stock GenerateFingerprint: F2 / F3), and temporal.NewEnergy sees FeatureConfig.WindowSize / HopSize as they
// were BEFORE lines :177-181 patch them (F4) -- which is why both are read from `extractorConfig`, the by-value copy
// CreateExtractor received.
func (fg *FingerprintGenerator) extractFeaturesGPU(audioData *transcode.AudioData, generationConfig *FingerprintConfig,
	extractorConfig config.FeatureConfig) (*extractors.ExtractedFeatures, error) {
	if len(audioData.PCM) == 0 {
		return nil, fmt.Errorf("PCM data cannot be empty") // speech.go:140
	}
	if audioData.SampleRate <= 0 {
		return nil, fmt.Errorf("sample rate must be positive") // speech.go:143
	}
	p := sonargpu.FpParams{
		WindowSize: generationConfig.WindowSize, HopSize: generationConfig.HopSize,
		WindowType:     windowIndex[generationConfig.FeatureConfig.WindowType],
		AlgoSampleRate: extractorConfig.SampleRate, CallSampleRate: audioData.SampleRate,
		EnergyFrame: extractorConfig.WindowSize, EnergyHop: extractorConfig.HopSize,
		MFCCCoefficients: extractorConfig.MFCCCoefficients, EnableMFCC: extractorConfig.EnableMFCC,
	}
	fp, err := sonargpu.GenerateFingerprint(audioData.PCM, p)
	if err != nil {
		return nil, err // carries the reference's own texts ("signal too short for given window size and hop size", ...)
	}
	rows := func(flat []float64, n, d int) [][]float64 {
		out := make([][]float64, n)
		for i := range out {
			out[i] = flat[i*d : (i+1)*d : (i+1)*d]
		}
		return out
	}
	f := &extractors.ExtractedFeatures{ExtractionMetadata: map[string]any{"extractor": "SpeechFeatureExtractor", "backend": "cuda-sm100a"}}
	if extractorConfig.EnableMFCC {
		f.MFCC = rows(fp.MFCC, fp.Frames, fp.NMFCC)
	}
	f.SpectralFeatures = &extractors.SpectralFeatures{
		SpectralCentroid: fp.Centroid, SpectralRolloff: fp.Rolloff, SpectralBandwidth: fp.Bandwidth,
		SpectralFlatness: fp.Flatness, SpectralCrest: fp.Crest, SpectralSlope: fp.Slope, SpectralFlux: fp.Flux,
		ZeroCrossingRate: fp.ZCR,
	}
	f.EnergyFeatures = &extractors.EnergyFeatures{
		ShortTimeEnergy: fp.ShortTimeEnergy, EnergyVariance: fp.EnergyVariance, EnergyEntropy: fp.EnergyEntropy,
		LoudnessRange: fp.LoudnessRange, LowEnergyRatio: fp.LowEnergyRatio, HighEnergyRatio: fp.HighEnergyRatio,
	}
	f.HarmonicFeatures = &extractors.HarmonicFeatures{
		PitchEstimate: fp.Pitch, PitchConfidence: fp.PitchConfidence, VoicingStrength: fp.Voicing,
		HarmonicRatio: fp.HarmonicRatio, InharmonicityRatio: fp.Inharmonicity, TonalCentroid: fp.TonalCentroid,
	}
	return f, nil
}
