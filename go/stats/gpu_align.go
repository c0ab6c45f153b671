// Drop into algorithms/stats/ of RyanBlaney/sonido-sonar (package stats): replaces the body of
// AlignmentAnalyzer.alignWithCrossCorrelation (alignment.go:151-181) and DTWAlignment.Align (dtw.go:55-103)
// with calls into libsonar.so.  NOT COMPILED HERE (no Go toolchain); see INTEGRATION.md.
package stats

import "github.com/RyanBlaney/sonido-sonar/sonargpu"

func (aa *AlignmentAnalyzer) alignWithCrossCorrelationGPU(query, reference [][]float64, result *AlignmentResult) (*AlignmentResult, error) {
	q, r := aa.flatten2DFeatures(query), aa.flatten2DFeatures(reference) // component 0, alignment.go:363-378
	corr, xs, ar, err := sonargpu.AlignCrossCorrelation(q, r, aa.maxLag, aa.hopSize, aa.sampleRate)
	if err != nil {
		return nil, err
	}
	lags := make([]int, len(corr))
	for i := range lags {
		lags[i] = i - xs.MaxLag
	}
	result.CrossCorrResult = &CorrelationResult{
		Correlations: corr, Lags: lags, PeakCorrelation: xs.PeakCorrelation, PeakLag: xs.PeakLag, PeakIndex: xs.PeakIndex,
		PValue: xs.PValue, IsSignificant: xs.IsSignificant, SNR: xs.SNR, Sharpness: xs.Sharpness,
		SecondPeak: xs.SecondPeak, PeakToSidelobe: xs.PeakToSidelobe, OverlapLength: xs.OverlapLength, MaxLag: xs.MaxLag,
	}
	result.Offset, result.OffsetSeconds = ar.Offset, ar.OffsetSeconds
	result.Confidence, result.Similarity = ar.Confidence, ar.Similarity
	result.AlignmentQuality, result.NoiseLevel = ar.AlignmentQuality, ar.NoiseLevel
	return result, nil
}

// AlignGPU keeps DTWAlignment.Align's contract; CostMatrix stays nil unless explicitly requested (the
// reference's full (n+1)x(m+1) matrix is 21 GB for a 5-minute pair: SURVEY F7).
func (dtw *DTWAlignment) AlignGPU(query, reference [][]float64) (*DTWResult, error) {
	n, m := len(query), len(reference)
	if n == 0 || m == 0 {
		return nil, errEmptySequences // "empty sequences provided"
	}
	dim := len(query[0])
	fq, fr := make([]float64, 0, n*dim), make([]float64, 0, m*dim)
	for _, v := range query {
		fq = append(fq, v...)
	}
	for _, v := range reference {
		fr = append(fr, v...)
	}
	pq, pr, pc, dist, err := sonargpu.DTW(fq, n, fr, m, dim, dtw.constraintBand)
	if err != nil {
		return nil, err
	}
	path := make([]AlignPoint, len(pq))
	for i := range path {
		path[i] = AlignPoint{QueryIndex: int(pq[i]), RefIndex: int(pr[i]), Cost: pc[i]}
	}
	return &DTWResult{Distance: dist, Path: path, QueryLength: n, RefLength: m,
		StepPattern: dtw.stepPattern, Constraint: dtw.constraintBand}, nil
}
