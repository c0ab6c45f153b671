// Times the reference's pure-Go hot path on the inputs of tests/golden/make_go_inputs.py (BASELINE.md section 3):
//   go test ./paritydump -run XXX -bench . -benchtime 3x -args -in <repo>/tests/golden/from_go/inputs
// Reports audio-seconds per second for the fingerprint cases and pair alignments per second for the align cases,
// with GOMAXPROCS as Go sets it (the reference's STFT uses runtime.NumCPU() workers, analyzers/spectral.go:217-230).
package paritydump

import (
	"runtime"
	"testing"

	"github.com/RyanBlaney/sonido-sonar/fingerprint"
	"github.com/RyanBlaney/sonido-sonar/fingerprint/config"
	"github.com/RyanBlaney/sonido-sonar/fingerprint/extractors"
	"github.com/RyanBlaney/sonido-sonar/logging"
)

func BenchmarkHotPath(b *testing.B) {
	logging.SetGlobalLogger(nil) // Info logs (feature_extractor.go:44, comparison.go:186) must not pollute the timing
	for _, c := range readManifest(b) {
		c := c
		switch c.Kind {
		case "fingerprint":
			pcm := readF64(b, c.PCM)
			b.Run(c.Name+"/GenerateFingerprint", func(b *testing.B) {
				g := fingerprint.NewFingerprintGenerator(harnessConfig(c))
				b.ResetTimer()
				for i := 0; i < b.N; i++ {
					if _, err := g.GenerateFingerprint(audio(pcm, c)); err != nil {
						b.Fatal(err)
					}
				}
				b.ReportMetric(float64(len(pcm))/float64(c.SampleRate)*float64(b.N)/b.Elapsed().Seconds(), "audio-s/s")
				b.ReportMetric(float64(runtime.GOMAXPROCS(0)), "GOMAXPROCS")
			})
		case "extract":
			pcm := readF64(b, c.PCM)
			b.Run(c.Name+"/ExtractFeaturesFixedRate", func(b *testing.B) {
				for i := 0; i < b.N; i++ {
					fixedRateFeatures(b, pcm, c)
				}
				b.ReportMetric(float64(len(pcm))/float64(c.SampleRate)*float64(b.N)/b.Elapsed().Seconds(), "audio-s/s")
			})
		case "align":
			q, r := readF64(b, c.PCM), readF64(b, c.PCM2)
			b.Run(c.Name+"/FingerprintBoth+ExtractAlignmentFeatures", func(b *testing.B) {
				for i := 0; i < b.N; i++ {
					fq, fr := fixedRateFeatures(b, q, c), fixedRateFeatures(b, r, c)
					fc := &config.FeatureConfig{SampleRate: c.SampleRate, WindowSize: c.Window, HopSize: c.Hop}
					ae := extractors.NewAlignmentExtractorWithMaxLag(fc, config.AlignmentConfigForContent(config.ContentMusic), c.MaxLagSec)
					if _, err := ae.ExtractAlignmentFeatures(fq, fr, q, r, c.SampleRate); err != nil {
						b.Fatal(err)
					}
				}
				b.ReportMetric(float64(b.N)/b.Elapsed().Seconds(), "alignments/s")
				b.ReportMetric(2*float64(len(q))/float64(c.SampleRate)*float64(b.N)/b.Elapsed().Seconds(), "audio-s/s")
			})
		}
	}
}
