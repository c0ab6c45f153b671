// Dumps the outputs of the REAL reference for the hot path, in the layout tests/golden_io.py loads.
// Drop this file into a directory inside a checkout of github.com/RyanBlaney/sonido-sonar (see README.md).
//
// Inputs: <in>/manifest.json (written by tests/golden/make_go_inputs.py) lists the cases; PCM / sequences are raw
// little-endian float64 files next to it.  Outputs: <out>/<case>/<name>.npy (float64 or int32, C order).
package paritydump

import (
	"encoding/binary"
	"encoding/json"
	"flag"
	"fmt"
	"math"
	"os"
	"path/filepath"
	"testing"

	"github.com/RyanBlaney/sonido-sonar/algorithms/stats"
	"github.com/RyanBlaney/sonido-sonar/fingerprint"
	"github.com/RyanBlaney/sonido-sonar/fingerprint/analyzers"
	"github.com/RyanBlaney/sonido-sonar/fingerprint/config"
	"github.com/RyanBlaney/sonido-sonar/fingerprint/extractors"
	"github.com/RyanBlaney/sonido-sonar/logging"
	"github.com/RyanBlaney/sonido-sonar/transcode"
)

var (
	inDir  = flag.String("in", "inputs", "directory with manifest.json and the raw float64 inputs")
	outDir = flag.String("out", "from_go", "directory the .npy outputs are written to")
)

type manifestCase struct {
	Name       string  `json:"name"`
	Kind       string  `json:"kind"` // fingerprint | extract | align | xcorr | dtw | compare
	PCM        string  `json:"pcm,omitempty"`
	PCM2       string  `json:"pcm2,omitempty"`
	A          string  `json:"a,omitempty"`
	B          string  `json:"b,omitempty"`
	Dim        int     `json:"dim,omitempty"`
	SampleRate int     `json:"sample_rate,omitempty"`
	AlgoRate   int     `json:"algo_sample_rate,omitempty"`
	Window     int     `json:"window_size,omitempty"`
	Hop        int     `json:"hop_size,omitempty"`
	Content    string  `json:"content_type,omitempty"`
	MaxLagSec  float64 `json:"max_lag_seconds,omitempty"`
	MaxLag     int     `json:"max_lag,omitempty"`
	Band       int     `json:"band,omitempty"`
	Step       string  `json:"step_pattern,omitempty"`
}

func readF64(t testing.TB, name string) []float64 {
	raw, err := os.ReadFile(filepath.Join(*inDir, name))
	if err != nil {
		t.Fatalf("read %s: %v", name, err)
	}
	out := make([]float64, len(raw)/8)
	for i := range out {
		out[i] = math.Float64frombits(binary.LittleEndian.Uint64(raw[8*i:]))
	}
	return out
}

func readManifest(t testing.TB) []manifestCase {
	raw, err := os.ReadFile(filepath.Join(*inDir, "manifest.json"))
	if err != nil {
		t.Fatalf("manifest: %v", err)
	}
	var cases []manifestCase
	if err := json.Unmarshal(raw, &cases); err != nil {
		t.Fatalf("manifest: %v", err)
	}
	return cases
}

// ---- .npy writers (format 1.0) -------------------------------------------------------------------------------

func npyHeader(descr string, shape []int) []byte {
	dims := ""
	for _, d := range shape {
		dims += fmt.Sprintf("%d,", d)
	}
	h := fmt.Sprintf("{'descr': '%s', 'fortran_order': False, 'shape': (%s), }", descr, dims)
	for (10+len(h)+1)%64 != 0 {
		h += " "
	}
	h += "\n"
	out := []byte("\x93NUMPY\x01\x00")
	out = binary.LittleEndian.AppendUint16(out, uint16(len(h)))
	return append(out, h...)
}

func writeF64(t testing.TB, dir, name string, v []float64, shape ...int) {
	if len(shape) == 0 {
		shape = []int{len(v)}
	}
	buf := npyHeader("<f8", shape)
	for _, x := range v {
		buf = binary.LittleEndian.AppendUint64(buf, math.Float64bits(x))
	}
	if err := os.WriteFile(filepath.Join(dir, name+".npy"), buf, 0o644); err != nil {
		t.Fatal(err)
	}
}

func writeI32(t testing.TB, dir, name string, v []int) {
	buf := npyHeader("<i4", []int{len(v)})
	for _, x := range v {
		buf = binary.LittleEndian.AppendUint32(buf, uint32(int32(x)))
	}
	if err := os.WriteFile(filepath.Join(dir, name+".npy"), buf, 0o644); err != nil {
		t.Fatal(err)
	}
}

func flat(m [][]float64) ([]float64, int, int) {
	if len(m) == 0 {
		return nil, 0, 0
	}
	out := make([]float64, 0, len(m)*len(m[0]))
	for _, r := range m {
		out = append(out, r...)
	}
	return out, len(m), len(m[0])
}

func dumpFeatures(t testing.TB, dir string, f *extractors.ExtractedFeatures) {
	if m, r, c := flat(f.MFCC); r > 0 {
		writeF64(t, dir, "mfcc", m, r, c)
	}
	if s := f.SpectralFeatures; s != nil {
		writeF64(t, dir, "spectral_centroid", s.SpectralCentroid)
		writeF64(t, dir, "spectral_rolloff", s.SpectralRolloff)
		writeF64(t, dir, "spectral_bandwidth", s.SpectralBandwidth)
		writeF64(t, dir, "spectral_flatness", s.SpectralFlatness)
		writeF64(t, dir, "spectral_crest", s.SpectralCrest)
		writeF64(t, dir, "spectral_slope", s.SpectralSlope)
		writeF64(t, dir, "spectral_flux", s.SpectralFlux)
		writeF64(t, dir, "zero_crossing_rate", s.ZeroCrossingRate)
	}
	if e := f.EnergyFeatures; e != nil {
		writeF64(t, dir, "short_time_energy", e.ShortTimeEnergy)
		writeF64(t, dir, "energy_entropy", e.EnergyEntropy)
		writeF64(t, dir, "low_energy_ratio", e.LowEnergyRatio)
		writeF64(t, dir, "high_energy_ratio", e.HighEnergyRatio)
		writeF64(t, dir, "energy_scalars", []float64{e.EnergyVariance, e.LoudnessRange})
	}
	if h := f.HarmonicFeatures; h != nil {
		writeF64(t, dir, "pitch_estimate", h.PitchEstimate)
		writeF64(t, dir, "pitch_confidence", h.PitchConfidence)
		writeF64(t, dir, "voicing_strength", h.VoicingStrength)
		writeF64(t, dir, "harmonic_ratio", h.HarmonicRatio)
		writeF64(t, dir, "inharmonicity_ratio", h.InharmonicityRatio)
		writeF64(t, dir, "tonal_centroid", h.TonalCentroid)
	}
	if tf := f.TemporalFeatures; tf != nil {
		writeF64(t, dir, "rms_energy", tf.RMSEnergy)
		writeF64(t, dir, "envelope_shape", tf.EnvelopeShape)
		writeF64(t, dir, "attack_time", tf.AttackTime)
		writeF64(t, dir, "temporal_scalars", []float64{tf.DynamicRange, tf.SilenceRatio, tf.PeakAmplitude,
			tf.AverageAmplitude, tf.OnsetDensity})
	}
}

// harnessConfig is the canonical configuration of SURVEY.md section 8(d): both levels carry window / hop (finding F4).
func harnessConfig(c manifestCase) *fingerprint.FingerprintConfig {
	return &fingerprint.FingerprintConfig{
		WindowSize: c.Window, HopSize: c.Hop, EnableContentDetect: false,
		FeatureConfig: &config.FeatureConfig{WindowSize: c.Window, HopSize: c.Hop, SampleRate: c.SampleRate,
			MFCCCoefficients: 13, WindowType: analyzers.WindowHann},
		ContentConfig: &config.ContentAwareConfig{},
	}
}

func audio(pcm []float64, c manifestCase) *transcode.AudioData {
	return &transcode.AudioData{PCM: pcm, SampleRate: c.SampleRate, Channels: 1,
		Metadata: &transcode.StreamMetadata{ContentType: c.Content}}
}

// fixedRateFeatures = what GenerateFingerprint would compute if buildFeatureConfig carried the sample rate over
// (findings F2/F3): the same STFT call (fingerprint.go:190) and the same extractor, constructed with SampleRate set.
func fixedRateFeatures(t testing.TB, pcm []float64, c manifestCase) *extractors.ExtractedFeatures {
	fc := config.FeatureConfig{SampleRate: c.AlgoRate, WindowSize: c.Window, HopSize: c.Hop, MFCCCoefficients: 13,
		WindowType: analyzers.WindowHann, EnableMFCC: true, EnableHarmonicFeatures: true,
		EnableTemporalFeatures: c.Content == "news" || c.Content == "talk"}
	ex, err := extractors.NewFeatureExtractorFactory().CreateExtractor(config.ToContentType(c.Content), fc)
	if err != nil {
		t.Fatal(err)
	}
	spec, err := analyzers.NewSpectralAnalyzer(c.SampleRate).ComputeSTFTWithWindow(pcm, c.Window, c.Hop, analyzers.WindowHann)
	if err != nil {
		t.Fatal(err)
	}
	f, err := ex.ExtractFeatures(spec, pcm, c.SampleRate)
	if err != nil {
		t.Fatal(err)
	}
	return f
}

func dumpAlign(t testing.TB, dir, prefix string, r *extractors.AlignmentResult) {
	if r == nil || r.AlignmentResult == nil {
		return
	}
	ok := 0.0
	if r.Success {
		ok = 1
	}
	writeF64(t, dir, prefix+"_scalars", []float64{float64(r.Offset), r.OffsetSeconds, r.Confidence, r.Similarity,
		r.AlignmentQuality, r.NoiseLevel, r.Stability, ok})
	if cr := r.CrossCorrResult; cr != nil {
		writeF64(t, dir, prefix+"_correlations", cr.Correlations)
		writeI32(t, dir, prefix+"_lags", cr.Lags)
		writeF64(t, dir, prefix+"_peak", []float64{cr.PeakCorrelation, float64(cr.PeakLag), float64(cr.PeakIndex), cr.PValue,
			cr.SNR, cr.Sharpness, cr.SecondPeak, cr.PeakToSidelobe, float64(cr.MaxLag), float64(cr.OverlapLength)})
	}
}

func dumpDTW(t testing.TB, dir string, d *stats.DTWResult) {
	q, r := make([]int, len(d.Path)), make([]int, len(d.Path))
	c := make([]float64, len(d.Path))
	for i, p := range d.Path {
		q[i], r[i], c[i] = p.QueryIndex, p.RefIndex, p.Cost
	}
	writeI32(t, dir, "path_query", q)
	writeI32(t, dir, "path_ref", r)
	writeF64(t, dir, "path_cost", c)
	writeF64(t, dir, "distance", []float64{d.Distance})
}

func rows(v []float64, dim int) [][]float64 {
	out := make([][]float64, len(v)/dim)
	for i := range out {
		out[i] = v[i*dim : (i+1)*dim]
	}
	return out
}

func TestParityDump(t *testing.T) {
	logging.SetGlobalLogger(nil)
	for _, c := range readManifest(t) {
		dir := filepath.Join(*outDir, c.Name)
		if err := os.MkdirAll(dir, 0o755); err != nil {
			t.Fatal(err)
		}
		switch c.Kind {
		case "fingerprint": // stock GenerateFingerprint (fingerprint.go:137-236): parity mode, algorithms at sample rate 0
			fp, err := fingerprint.NewFingerprintGenerator(harnessConfig(c)).GenerateFingerprint(audio(readF64(t, c.PCM), c))
			if err != nil {
				t.Fatalf("%s: %v", c.Name, err)
			}
			dumpFeatures(t, dir, fp.Features)
		case "extract": // fixed sample rate mode
			dumpFeatures(t, dir, fixedRateFeatures(t, readF64(t, c.PCM), c))
		case "align": // ExtractAlignmentFeatures (extractors/alignment.go:139-219) on fixed-rate features of a pair
			q, r := readF64(t, c.PCM), readF64(t, c.PCM2)
			fq, fr := fixedRateFeatures(t, q, c), fixedRateFeatures(t, r, c)
			fc := &config.FeatureConfig{SampleRate: c.SampleRate, WindowSize: c.Window, HopSize: c.Hop}
			ae := extractors.NewAlignmentExtractorWithMaxLag(fc, config.AlignmentConfigForContent(config.ContentMusic), c.MaxLagSec)
			af, err := ae.ExtractAlignmentFeatures(fq, fr, q, r, c.SampleRate)
			if err != nil {
				t.Fatalf("%s: %v", c.Name, err)
			}
			writeF64(t, dir, "query_short_time_energy", fq.EnergyFeatures.ShortTimeEnergy)
			writeF64(t, dir, "reference_short_time_energy", fr.EnergyFeatures.ShortTimeEnergy)
			dumpAlign(t, dir, "corr", af.CorrAlignment)
			dumpAlign(t, dir, "best", af.BestAlignment)
			writeF64(t, dir, "summary", []float64{af.TemporalOffset, af.OffsetConfidence, af.TimeStretch, af.AlignmentSimilarity,
				af.AlignmentQuality})
		case "xcorr": // CrossCorrelation.Compute as NewAlignmentAnalyzer configures it (stats/alignment.go:60-81)
			cc := stats.NewCrossCorrelationWithParams(c.MaxLag, stats.NormalizedCrossCorrelation, stats.TimeDomain)
			res, err := cc.Compute(readF64(t, c.A), readF64(t, c.B))
			if err != nil {
				t.Fatalf("%s: %v", c.Name, err)
			}
			writeF64(t, dir, "correlations", res.Correlations)
			writeI32(t, dir, "lags", res.Lags)
			writeF64(t, dir, "peak", []float64{res.PeakCorrelation, float64(res.PeakLag), float64(res.PeakIndex), res.PValue, res.SNR,
				res.Sharpness, res.SecondPeak, res.PeakToSidelobe, float64(res.MaxLag), float64(res.OverlapLength)})
		case "dtw": // DTWAlignment.Align (stats/dtw.go:55-103)
			d, err := stats.NewDTWAlignmentWithParams(c.Band, c.Step, stats.EuclideanDistance).Align(rows(readF64(t, c.A), c.Dim),
				rows(readF64(t, c.B), c.Dim))
			if err != nil {
				t.Fatalf("%s: %v", c.Name, err)
			}
			dumpDTW(t, dir, d)
		case "compare": // FingerprintComparator.Compare (comparison.go:133-194) of two stock fingerprints
			g := fingerprint.NewFingerprintGenerator(harnessConfig(c))
			f1, err1 := g.GenerateFingerprint(audio(readF64(t, c.PCM), c))
			f2, err2 := g.GenerateFingerprint(audio(readF64(t, c.PCM2), c))
			if err1 != nil || err2 != nil {
				t.Fatalf("%s: %v %v", c.Name, err1, err2)
			}
			res, err := fingerprint.NewFingerprintComparator(fingerprint.DefaultComparisonConfig()).Compare(f1, f2)
			if err != nil {
				t.Fatalf("%s: %v", c.Name, err)
			}
			match := 0.0
			if res.ContentTypeMatch {
				match = 1
			}
			writeF64(t, dir, "result", []float64{res.OverallSimilarity, res.FeatureSimilarity, res.Confidence, match})
			keys := []string{"mfcc", "spectral", "temporal", "harmonic", "energy", "chroma", "speech"}
			dist := make([]float64, len(keys))
			for i, k := range keys {
				if v, ok := res.FeatureDistances[k]; ok {
					dist[i] = v
				} else {
					dist[i] = math.NaN()
				}
			}
			writeF64(t, dir, "feature_distances", dist)
			dumpFeatures(t, filepath.Join(dir), f1.Features)
		default:
			t.Fatalf("unknown case kind %q", c.Kind)
		}
	}
}
