// Package sonargpu is the thin cgo layer over libsonar.so (include/sonar.h).
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain (SURVEY.md §8c).  It is the
// binding a maintainer of RyanBlaney/sonido-sonar adds to switch the hot path to the B200 library; the
// same entry points are exercised from C++ (sonido-sonar_b200/host/sonar_host.hpp) and Python
// (sonido-sonar_b200/capi.py) in this repo's tests.
//
// cgo rules honoured here: Go owns every input and output buffer, C never retains a Go pointer after the
// call returns, and no Go pointer to a Go pointer crosses the boundary — every [][]float64 is flattened
// into one backing array first.
package sonargpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../sonido-sonar_b200 -lsonar -Wl,-rpath,${SRCDIR}/../../sonido-sonar_b200
#include <stdlib.h>
#include "sonar.h"
*/
import "C"

import (
	"errors"
	"runtime"
	"sync"
	"unsafe"
)

var (
	once   sync.Once
	ctx    *C.sonar_ctx
	ctxErr error
)

// Ctx returns the process-wide context (one per process; the C layer is re-entrant).
func Ctx() (*C.sonar_ctx, error) {
	once.Do(func() {
		if rc := C.sonar_init(0, nil, &ctx); rc != C.SONAR_OK {
			ctxErr = errors.New(C.GoString(C.sonar_last_error())) // no CPU fallback: surfaces "no usable CUDA device"
		}
	})
	return ctx, ctxErr
}

// lastError fetches the thread-local message of the failing call.  The message lives in a thread_local of the C
// library, and a goroutine may be moved to another OS thread between two cgo calls: every binding below therefore runs
// its call AND this fetch between runtime.LockOSThread / UnlockOSThread (see locked()).
func lastError() error {
	return errors.New(C.GoString(C.sonar_last_error()))
}

// locked runs f (one sonar_* call) on a pinned OS thread and turns a non-zero status into the reference's error text.
func locked(f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := f(); rc != C.SONAR_OK {
		return lastError()
	}
	return nil
}

func ptr(s []float64) *C.double {
	if len(s) == 0 {
		return nil
	}
	return (*C.double)(unsafe.Pointer(&s[0]))
}

// FpParams mirrors sonar_fp_params: exactly what the reference's algorithm objects are constructed with.
type FpParams struct {
	WindowSize, HopSize, WindowType  int
	AlgoSampleRate, CallSampleRate   int // AlgoSampleRate is 0 through stock GenerateFingerprint (F2/F3)
	EnergyFrame, EnergyHop           int // FeatureConfig.WindowSize/HopSize as temporal.NewEnergy sees them (F4)
	MFCCCoefficients                 int
	EnableMFCC                       bool
}

// SpeechFeatures mirrors the frame-level part of extractors.SpeechFeatures (fingerprint/extractors/features.go:45-65) that
// sonar_fingerprint_speech_f64 produces.  FormantFrequencies, VocalTractLength, Jitter and Shimmer stay with the
// reference's host-side analyzers (algorithms/speech): extractSpeechFeatures keeps calling them for those four fields.
type SpeechFeatures struct {
	IsSpeech           bool
	VoicingProbability []float64
	SpectralTilt       []float64
	PauseDuration      []float64
	SpeechRate         float64
}

// Fingerprint holds the flat outputs of sonar_fingerprint_f64 (row-major MFCC).
type Fingerprint struct {
	Frames, EnergyFrames, PitchFrames, NMFCC int
	MFCC                                      []float64
	Centroid, Rolloff, Bandwidth, Flatness    []float64
	Crest, Slope, Flux, ZCR                   []float64
	ShortTimeEnergy, EnergyEntropy            []float64
	LowEnergyRatio, HighEnergyRatio           []float64
	Pitch, PitchConfidence, Voicing           []float64
	HarmonicRatio, Inharmonicity, TonalCentroid []float64
	EnergyVariance, LoudnessRange             float64
}

// GenerateFingerprint replaces ComputeSTFTWithWindow + SpeechFeatureExtractor.ExtractFeatures
// (fingerprint/fingerprint.go:190-207).
func GenerateFingerprint(pcm []float64, p FpParams) (*Fingerprint, error) {
	f, _, err := generate(pcm, p, false)
	return f, err
}

// GenerateFingerprintSpeech is GenerateFingerprint with FeatureConfig.EnableSpeechFeatures (news / talk): the speech
// group runs before the harmonic block on the shared pitch detector, exactly as extractors/speech.go:194-205 orders it.
func GenerateFingerprintSpeech(pcm []float64, p FpParams) (*Fingerprint, *SpeechFeatures, error) {
	return generate(pcm, p, true)
}

func generate(pcm []float64, p FpParams, speech bool) (*Fingerprint, *SpeechFeatures, error) {
	runtime.LockOSThread() // the error message of a failing call is thread-local in the C library (see lastError)
	defer runtime.UnlockOSThread()
	c, err := Ctx()
	if err != nil {
		return nil, nil, err
	}
	var cp C.sonar_fp_params
	C.sonar_fp_params_default(&cp)
	cp.window_size, cp.hop_size, cp.window_type = C.int32_t(p.WindowSize), C.int32_t(p.HopSize), C.int32_t(p.WindowType)
	cp.algo_sample_rate, cp.call_sample_rate = C.int32_t(p.AlgoSampleRate), C.int32_t(p.CallSampleRate)
	cp.energy_frame, cp.energy_hop = C.int32_t(p.EnergyFrame), C.int32_t(p.EnergyHop)
	cp.n_mfcc = C.int32_t(p.MFCCCoefficients)
	cp.enable = 0
	if p.EnableMFCC {
		cp.enable |= C.SONAR_FP_ENABLE_MFCC
	}
	var sz C.sonar_fp_sizes_t
	if rc := C.sonar_fp_sizes(&cp, C.int64_t(len(pcm)), &sz); rc != C.SONAR_OK {
		return nil, nil, lastError() // "empty signal", "signal too short for given window size and hop size", ...
	}
	T, Te, Tp, K := int(sz.n_frames), int(sz.n_energy_frames), int(sz.n_pitch_frames), int(sz.n_mfcc)
	f := &Fingerprint{Frames: T, EnergyFrames: Te, PitchFrames: Tp, NMFCC: K}
	mk := func(n int) []float64 { return make([]float64, n) }
	f.MFCC = mk(T * K)
	f.Centroid, f.Rolloff, f.Bandwidth, f.Flatness = mk(T), mk(T), mk(T), mk(T)
	f.Crest, f.Slope, f.Flux, f.ZCR = mk(T), mk(T), mk(int(sz.n_flux)), mk(T)
	f.ShortTimeEnergy, f.EnergyEntropy, f.LowEnergyRatio, f.HighEnergyRatio = mk(Te), mk(Te), mk(Te), mk(Te)
	f.Pitch, f.PitchConfidence, f.Voicing = mk(Tp), mk(Tp), mk(Tp)
	f.HarmonicRatio, f.Inharmonicity, f.TonalCentroid = mk(Tp), mk(Tp), mk(Tp)
	// sonar_fp_out holds Go pointers, so it must live in C memory for the duration of the call (cgo rule:
	// a Go struct containing Go pointers may not be passed); the pinner keeps the slices in place.
	out := (*C.sonar_fp_out)(C.calloc(1, C.size_t(unsafe.Sizeof(C.sonar_fp_out{}))))
	defer C.free(unsafe.Pointer(out))
	var pin runtime.Pinner
	defer pin.Unpin()
	set := func(dst **C.double, s []float64) {
		if len(s) > 0 {
			pin.Pin(&s[0])
			*dst = ptr(s)
		}
	}
	set(&out.mfcc, f.MFCC)
	set(&out.spectral_centroid, f.Centroid)
	set(&out.spectral_rolloff, f.Rolloff)
	set(&out.spectral_bandwidth, f.Bandwidth)
	set(&out.spectral_flatness, f.Flatness)
	set(&out.spectral_crest, f.Crest)
	set(&out.spectral_slope, f.Slope)
	set(&out.spectral_flux, f.Flux)
	set(&out.zero_crossing_rate, f.ZCR)
	set(&out.short_time_energy, f.ShortTimeEnergy)
	set(&out.energy_entropy, f.EnergyEntropy)
	set(&out.low_energy_ratio, f.LowEnergyRatio)
	set(&out.high_energy_ratio, f.HighEnergyRatio)
	set(&out.pitch_estimate, f.Pitch)
	set(&out.pitch_confidence, f.PitchConfidence)
	set(&out.voicing_strength, f.Voicing)
	set(&out.harmonic_ratio, f.HarmonicRatio)
	set(&out.inharmonicity_ratio, f.Inharmonicity)
	set(&out.tonal_centroid, f.TonalCentroid)
	var sf *SpeechFeatures
	if speech {
		so := (*C.sonar_speech_out)(C.calloc(1, C.size_t(unsafe.Sizeof(C.sonar_speech_out{}))))
		defer C.free(unsafe.Pointer(so))
		voi, tilt, pauses := mk(Tp), mk(Tp), mk(Te/2+1)
		set(&so.voicing_probability, voi)
		set(&so.spectral_tilt, tilt)
		set(&so.pause_duration, pauses)
		so.pause_cap = C.int64_t(len(pauses))
		if rc := C.sonar_fingerprint_speech_f64(c, ptr(pcm), C.int64_t(len(pcm)), &cp, out, so); rc != C.SONAR_OK {
			return nil, nil, lastError()
		}
		nf, np := int(so.n_frames), int(so.n_pause)
		if np > len(pauses) {
			np = len(pauses)
		}
		sf = &SpeechFeatures{IsSpeech: so.is_speech != 0, VoicingProbability: voi[:nf], SpectralTilt: tilt[:nf],
			PauseDuration: pauses[:np], SpeechRate: float64(so.speech_rate)}
	} else if rc := C.sonar_fingerprint_f64(c, ptr(pcm), C.int64_t(len(pcm)), &cp, out); rc != C.SONAR_OK {
		return nil, nil, lastError()
	}
	f.EnergyVariance, f.LoudnessRange = float64(out.energy_variance), float64(out.loudness_range)
	return f, sf, nil
}

// ExactFrameCounts reports how many frames of the last fingerprint batch were re-evaluated in float64 in the reference's
// order instead of taking the FP32 kernels' values (sonar_fp_exact_counts; diagnostic).
func ExactFrameCounts() (spectral, pitch int64, err error) {
	c, err := Ctx()
	if err != nil {
		return 0, 0, err
	}
	var a, b C.int64_t
	err = locked(func() C.int { return C.sonar_fp_exact_counts(c, &a, &b) })
	return int64(a), int64(b), err
}

// NCCLUniqueID / NCCLInit give the library its own communicator over the ranks of a multi-process job (one process per
// GPU): rank 0 draws the id, the host distributes the 128 bytes by any means, every rank calls NCCLInit.
func NCCLUniqueID() ([]byte, error) {
	id := make([]byte, C.SONAR_NCCL_ID_BYTES)
	err := locked(func() C.int { return C.sonar_nccl_unique_id((*C.uchar)(unsafe.Pointer(&id[0])), C.int(len(id))) })
	return id, err
}

func NCCLInit(world, rank int, id []byte) error {
	c, err := Ctx()
	if err != nil {
		return err
	}
	return locked(func() C.int {
		return C.sonar_nccl_init(c, C.int(world), C.int(rank), (*C.uchar)(unsafe.Pointer(&id[0])))
	})
}

// CrossCorrelationSharded is CrossCorrelation.Compute of ONE long pair with the lag range split over the ranks of the
// communicator (every rank calls it with the same sequences and receives the identical result): sonar_xcorr_lag_sharded.
func CrossCorrelationSharded(a, b []float64, maxLag int) (*XcorrSummary, error) {
	c, err := Ctx()
	if err != nil {
		return nil, err
	}
	var s C.sonar_xcorr_summary
	err = locked(func() C.int {
		return C.sonar_xcorr_lag_sharded(c, ptr(a), C.int64_t(len(a)), ptr(b), C.int64_t(len(b)), C.int(maxLag), 0, nil, &s)
	})
	if err != nil {
		return nil, err
	}
	return &XcorrSummary{PeakCorrelation: float64(s.peak_correlation), PValue: float64(s.p_value), SNR: float64(s.snr),
		Sharpness: float64(s.sharpness), SecondPeak: float64(s.second_peak), PeakToSidelobe: float64(s.peak_to_sidelobe),
		PeakLag: int(s.peak_lag), PeakIndex: int(s.peak_index), MaxLag: int(s.actual_max_lag),
		OverlapLength: int(s.overlap_length), IsSignificant: s.is_significant != 0}, nil
}

// XcorrSummary mirrors stats.CorrelationResult without the arrays (algorithms/stats/correlation.go:44-71).
type XcorrSummary struct {
	PeakCorrelation, PValue, SNR, Sharpness, SecondPeak, PeakToSidelobe float64
	PeakLag, PeakIndex, MaxLag, OverlapLength                           int
	IsSignificant                                                       bool
}

// AlignResult mirrors the scalar fields of stats.AlignmentResult (algorithms/stats/alignment.go:34-58).
type AlignResult struct {
	Offset                                                              int
	OffsetSeconds, Confidence, Similarity, AlignmentQuality, NoiseLevel float64
}

// AlignCrossCorrelation replaces AlignmentAnalyzer.AlignFeatures(method = AlignmentCrossCorrelation) as
// extractors.alignWithFeatures calls it for "corr_energy" (fingerprint/extractors/alignment.go:357-409).
func AlignCrossCorrelation(query, reference []float64, maxLagFrames, hopSize, sampleRate int) ([]float64, *XcorrSummary, *AlignResult, error) {
	runtime.LockOSThread() // the error message of a failing call is thread-local in the C library (see lastError)
	defer runtime.UnlockOSThread()
	c, err := Ctx()
	if err != nil {
		return nil, nil, nil, err
	}
	if maxLagFrames < 0 {
		maxLagFrames = 0
	}
	corr := make([]float64, 2*maxLagFrames+1)
	var xs C.sonar_xcorr_summary
	var ar C.sonar_align_result
	rc := C.sonar_align_xcorr_f64(c, ptr(query), C.int64_t(len(query)), ptr(reference), C.int64_t(len(reference)),
		C.int(maxLagFrames), C.int(hopSize), C.int(sampleRate), ptr(corr), &xs, &ar)
	if rc != C.SONAR_OK {
		return nil, nil, nil, lastError() // "empty feature sequences provided"
	}
	s := &XcorrSummary{float64(xs.peak_correlation), float64(xs.p_value), float64(xs.snr), float64(xs.sharpness),
		float64(xs.second_peak), float64(xs.peak_to_sidelobe), int(xs.peak_lag), int(xs.peak_index),
		int(xs.actual_max_lag), int(xs.overlap_length), xs.is_significant != 0}
	a := &AlignResult{int(ar.offset), float64(ar.offset_seconds), float64(ar.confidence), float64(ar.similarity),
		float64(ar.alignment_quality), float64(ar.noise_level)}
	return corr[:2*s.MaxLag+1], s, a, nil
}

// DTW replaces DTWAlignment.Align (algorithms/stats/dtw.go:55-217).  q and r are row-major [n][dim].
func DTW(q []float64, n int, r []float64, m, dim, band int) (pathQ, pathR []int32, pathCost []float64, distance float64, err error) {
	runtime.LockOSThread() // the error message of a failing call is thread-local in the C library (see lastError)
	defer runtime.UnlockOSThread()
	c, err := Ctx()
	if err != nil {
		return nil, nil, nil, 0, err
	}
	pathQ, pathR, pathCost = make([]int32, n+m), make([]int32, n+m), make([]float64, n+m)
	out := (*C.sonar_dtw_out)(C.calloc(1, C.size_t(unsafe.Sizeof(C.sonar_dtw_out{}))))
	defer C.free(unsafe.Pointer(out))
	var pin runtime.Pinner
	defer pin.Unpin()
	pin.Pin(&pathQ[0])
	pin.Pin(&pathR[0])
	pin.Pin(&pathCost[0])
	out.path_query = (*C.int32_t)(unsafe.Pointer(&pathQ[0]))
	out.path_ref = (*C.int32_t)(unsafe.Pointer(&pathR[0]))
	out.path_cost = ptr(pathCost)
	out.path_cap = C.int64_t(n + m)
	if rc := C.sonar_dtw_f64(c, ptr(q), C.int(n), ptr(r), C.int(m), C.int(dim), C.int(band),
		C.SONAR_STEP_SYMMETRIC2, C.SONAR_METRIC_EUCLIDEAN, out); rc != C.SONAR_OK {
		return nil, nil, nil, 0, lastError() // "empty sequences provided"
	}
	l := int(out.path_len)
	return pathQ[:l], pathR[:l], pathCost[:l], float64(out.distance), nil
}

// ColStatsCosine replaces extractMFCCStatistics x2 + cosineSimilarity (fingerprint/comparison.go:774-873).
func ColStatsCosine(x []float64, tx int, y []float64, ty, dim int) (float64, error) {
	runtime.LockOSThread() // the error message of a failing call is thread-local in the C library (see lastError)
	defer runtime.UnlockOSThread()
	c, err := Ctx()
	if err != nil {
		return 0, err
	}
	var sim C.double
	if rc := C.sonar_colstats_cosine_f64(c, ptr(x), C.int64_t(tx), ptr(y), C.int64_t(ty), C.int(dim), &sim); rc != C.SONAR_OK {
		return 0, lastError()
	}
	return float64(sim), nil
}

// PCM sample formats of the *_pcm entry points (include/sonar.h).
const (
	PCMF64 = C.SONAR_PCM_F64
	PCMF32 = C.SONAR_PCM_F32
	PCMS16 = C.SONAR_PCM_S16
)

// PairResult is what one source/CDN pair of AlignPairsS16 / AlignPairsF64 yields: the "corr_energy" lag and the
// banded DTW path of the lag-trimmed short-time energies.
type PairResult struct {
	Xcorr              XcorrSummary
	Align              AlignResult
	PathQ, PathR       []int32
	PathCost           []float64
	Distance           float64
}

// alignPairs drives sonar_align_pairs_pcm: the whole CDN-latency loop of AlignmentExtractor.AlignAudioFiles
// (fingerprint/extractors/alignment.go:489-560) for a batch of equally long pairs, chained on the device.
// The per-pair sample pointers are copied into C memory (no Go pointer to Go pointer crosses the boundary) and
// the sample slices themselves are pinned for the duration of the call.
func alignPairs(q, r []unsafe.Pointer, format C.int, n int, p FpParams, maxLagSeconds float64, dtwBand int) ([]PairResult, error) {
	runtime.LockOSThread() // the error message of a failing call is thread-local in the C library (see lastError)
	defer runtime.UnlockOSThread()
	c, err := Ctx()
	if err != nil {
		return nil, err
	}
	np := len(q)
	if np == 0 || len(r) != np {
		return nil, errors.New("feature sets cannot be nil")
	}
	var cp C.sonar_fp_params
	C.sonar_fp_params_default(&cp)
	cp.window_size, cp.hop_size, cp.window_type = C.int32_t(p.WindowSize), C.int32_t(p.HopSize), C.int32_t(p.WindowType)
	cp.algo_sample_rate, cp.call_sample_rate = C.int32_t(p.AlgoSampleRate), C.int32_t(p.CallSampleRate)
	cp.energy_frame, cp.energy_hop = C.int32_t(p.EnergyFrame), C.int32_t(p.EnergyHop)
	var nLags, dtwLen C.int32_t
	if rc := C.sonar_align_pairs_sizes(&cp, C.int64_t(n), C.double(maxLagSeconds), &nLags, &dtwLen); rc != C.SONAR_OK {
		return nil, lastError()
	}
	ptrBytes := C.size_t(np) * C.size_t(unsafe.Sizeof(uintptr(0)))
	cq := (*[1 << 28]unsafe.Pointer)(C.malloc(ptrBytes))
	cr := (*[1 << 28]unsafe.Pointer)(C.malloc(ptrBytes))
	outs := (*[1 << 20]C.sonar_pair_out)(C.calloc(C.size_t(np), C.size_t(unsafe.Sizeof(C.sonar_pair_out{}))))
	defer C.free(unsafe.Pointer(cq))
	defer C.free(unsafe.Pointer(cr))
	defer C.free(unsafe.Pointer(outs))
	var pin runtime.Pinner
	defer pin.Unpin()
	res := make([]PairResult, np)
	pcap := 2 * int(dtwLen)
	for i := 0; i < np; i++ {
		pin.Pin(q[i])
		pin.Pin(r[i])
		cq[i], cr[i] = q[i], r[i]
		res[i].PathQ, res[i].PathR, res[i].PathCost = make([]int32, pcap), make([]int32, pcap), make([]float64, pcap)
		pin.Pin(&res[i].PathQ[0])
		pin.Pin(&res[i].PathR[0])
		pin.Pin(&res[i].PathCost[0])
		outs[i].dtw.path_query = (*C.int32_t)(unsafe.Pointer(&res[i].PathQ[0]))
		outs[i].dtw.path_ref = (*C.int32_t)(unsafe.Pointer(&res[i].PathR[0]))
		outs[i].dtw.path_cost = ptr(res[i].PathCost)
		outs[i].dtw.path_cap = C.int64_t(pcap)
	}
	rc := C.sonar_align_pairs_pcm(c, (*unsafe.Pointer)(unsafe.Pointer(cq)), (*unsafe.Pointer)(unsafe.Pointer(cr)), format,
		C.int64_t(n), C.int(np), &cp, C.double(maxLagSeconds), C.int(dtwBand), &outs[0])
	if rc != C.SONAR_OK {
		return nil, lastError()
	}
	for i := 0; i < np; i++ {
		xs, ar, d := outs[i].xcorr, outs[i].corr_alignment, outs[i].dtw
		res[i].Xcorr = XcorrSummary{float64(xs.peak_correlation), float64(xs.p_value), float64(xs.snr), float64(xs.sharpness),
			float64(xs.second_peak), float64(xs.peak_to_sidelobe), int(xs.peak_lag), int(xs.peak_index),
			int(xs.actual_max_lag), int(xs.overlap_length), xs.is_significant != 0}
		res[i].Align = AlignResult{int(ar.offset), float64(ar.offset_seconds), float64(ar.confidence), float64(ar.similarity),
			float64(ar.alignment_quality), float64(ar.noise_level)}
		l := int(d.path_len)
		res[i].PathQ, res[i].PathR, res[i].PathCost = res[i].PathQ[:l], res[i].PathR[:l], res[i].PathCost[:l]
		res[i].Distance = float64(d.distance)
	}
	return res, nil
}

// AlignPairsF64 takes the []float64 PCM the reference's decoder produces (transcode/decoder.go:850-870).
func AlignPairsF64(query, reference [][]float64, p FpParams, maxLagSeconds float64, dtwBand int) ([]PairResult, error) {
	q, r := make([]unsafe.Pointer, len(query)), make([]unsafe.Pointer, len(reference))
	for i := range query {
		q[i], r[i] = unsafe.Pointer(&query[i][0]), unsafe.Pointer(&reference[i][0])
	}
	return alignPairs(q, r, PCMF64, len(query[0]), p, maxLagSeconds, dtwBand)
}

// AlignPairsS16 takes the s16le samples ffmpeg holds before the reference asks it for "-f f64le"
// (transcode/decoder.go:707-712): a quarter of the PCIe bytes, bit-identical results (x / 32768 on the device).
func AlignPairsS16(query, reference [][]int16, p FpParams, maxLagSeconds float64, dtwBand int) ([]PairResult, error) {
	q, r := make([]unsafe.Pointer, len(query)), make([]unsafe.Pointer, len(reference))
	for i := range query {
		q[i], r[i] = unsafe.Pointer(&query[i][0]), unsafe.Pointer(&reference[i][0])
	}
	return alignPairs(q, r, PCMS16, len(query[0]), p, maxLagSeconds, dtwBand)
}

// MusicSpectral replaces the three per-frame loops of MusicFeatureExtractor.extractSpectralFeatures
// (fingerprint/extractors/music.go:261-302): SpectralContrast.Compute, the chroma folding of
// ChromaSTFT.convertSTFTToChroma and BarkScale.ComputeBarkSpectrum.  Row-major [T][nBands], [T][12], [T][nBark].
func MusicSpectral(pcm []float64, win, hop, windowType, sampleRate, nBands, nBark int, barkLow, barkHigh float64) (contrast, chroma, bark []float64, frames int, err error) {
	runtime.LockOSThread() // the error message of a failing call is thread-local in the C library (see lastError)
	defer runtime.UnlockOSThread()
	c, err := Ctx()
	if err != nil {
		return nil, nil, nil, 0, err
	}
	if win <= 0 || hop <= 0 || len(pcm) < win {
		return nil, nil, nil, 0, errors.New("signal too short for given window size and hop size")
	}
	frames = (len(pcm)-win)/hop + 1
	contrast, chroma, bark = make([]float64, frames*nBands), make([]float64, frames*12), make([]float64, frames*nBark)
	rc := C.sonar_music_spectral_f64(c, ptr(pcm), C.int64_t(len(pcm)), C.int(win), C.int(hop), C.int(windowType),
		C.int(sampleRate), C.int(nBands), ptr(contrast), ptr(chroma), C.int(nBark), C.double(barkLow), C.double(barkHigh), ptr(bark))
	if rc != C.SONAR_OK {
		return nil, nil, nil, 0, lastError()
	}
	return contrast, chroma, bark, frames, nil
}

// STFTStream replaces analyzers.STFTStreamer (fingerprint/analyzers/spectral.go:312-374): the buffer and its
// advance/empty rule live behind sonar_stft_stream_*, the frames come from the same transform as ComputeSTFTWithWindow.
type STFTStream struct {
	h    *C.sonar_stft_stream
	bins int
}

// NewSTFTStream is what SpectralAnalyzer.ComputeSTFTStreaming (spectral.go:289-310) calls.
func NewSTFTStream(win, hop, windowType int) (*STFTStream, error) {
	runtime.LockOSThread() // the error message of a failing call is thread-local in the C library (see lastError)
	defer runtime.UnlockOSThread()
	c, err := Ctx()
	if err != nil {
		return nil, err
	}
	var h *C.sonar_stft_stream
	if rc := C.sonar_stft_stream_open(c, C.int(win), C.int(hop), C.int(windowType), &h); rc != C.SONAR_OK {
		return nil, lastError()
	}
	return &STFTStream{h: h, bins: win/2 + 1}, nil
}

// ProcessChunk returns row-major magnitude / phase [T][bins] and complex [T][bins][2] for the T frames the chunk
// completes (T may be 0; an empty chunk is not an error, spectral.go:324-326).
func (s *STFTStream) ProcessChunk(chunk []float64) (mag, phase, cplx []float64, frames int, err error) {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if len(chunk) == 0 {
		return nil, nil, nil, 0, nil
	}
	t := int(C.sonar_stft_stream_frames(s.h, C.int64_t(len(chunk))))
	mag, phase, cplx = make([]float64, t*s.bins), make([]float64, t*s.bins), make([]float64, 2*t*s.bins)
	var got C.int64_t
	rc := C.sonar_stft_stream_process(s.h, ptr(chunk), C.int64_t(len(chunk)), ptr(mag), ptr(phase), ptr(cplx), C.int64_t(t), &got)
	if rc != C.SONAR_OK {
		return nil, nil, nil, 0, lastError()
	}
	return mag, phase, cplx, int(got), nil
}

// Close releases the native buffer (the reference's streamer is garbage collected; this one is not).
func (s *STFTStream) Close() {
	if s.h != nil {
		C.sonar_stft_stream_close(s.h)
		s.h = nil
	}
}

// ---- FingerprintComparator.Compare (fingerprint/comparison.go:133-194) ---------------------------------------------

// CmpFeatures is one side of a comparison: the feature arrays Compare reads (comparison.go:266-341, 646-771),
// flattened (MFCC row-major [frames][dim]).  A nil / empty slice means "feature absent".
type CmpFeatures struct {
	MFCC                                       []float64
	MFCCDim                                    int
	ContentType                                int // index of config.ContentType in the order music, news, sports, talk, mixed, unknown
	Centroid, Rolloff, Flux                    []float64
	HasSpectral, HasHarmonic, HasTemporal      bool
	HarmonicRatio, Pitch, RMSEnergy            []float64
	DynamicRange, SilenceRatio, OnsetDensity   float64
}

// CmpResult mirrors the scalars of fingerprint.SimilarityResult (comparison.go:28-39); a NaN distance = not compared.
type CmpResult struct {
	OverallSimilarity, FeatureSimilarity, Confidence     float64
	DistMFCC, DistSpectral, DistTemporal, DistHarmonic   float64
	ContentTypeMatch                                     bool
	NFeatures                                            int
}

// fill writes f into C memory (the struct holds pointers into Go slices, pinned for the duration of the call).
func (f *CmpFeatures) fill(c *C.sonar_cmp_features, pin *runtime.Pinner) {
	set := func(dst **C.double, n *C.int64_t, s []float64, div int) {
		if len(s) > 0 {
			pin.Pin(&s[0])
			*dst = ptr(s)
			*n = C.int64_t(len(s) / div)
		}
	}
	dim := f.MFCCDim
	if dim <= 0 {
		dim = 1
	}
	set(&c.mfcc, &c.mfcc_frames, f.MFCC, dim)
	c.mfcc_dim, c.content_type = C.int32_t(dim), C.int32_t(f.ContentType)
	set(&c.spectral_centroid, &c.n_centroid, f.Centroid, 1)
	set(&c.spectral_rolloff, &c.n_rolloff, f.Rolloff, 1)
	set(&c.spectral_flux, &c.n_flux, f.Flux, 1)
	set(&c.harmonic_ratio, &c.n_harmonic_ratio, f.HarmonicRatio, 1)
	set(&c.pitch_estimate, &c.n_pitch, f.Pitch, 1)
	set(&c.rms_energy, &c.n_rms, f.RMSEnergy, 1)
	b := func(v bool) C.int32_t {
		if v {
			return 1
		}
		return 0
	}
	c.has_spectral, c.has_harmonic, c.has_temporal = b(f.HasSpectral), b(f.HasHarmonic), b(f.HasTemporal)
	c.dynamic_range, c.silence_ratio, c.onset_density = C.double(f.DynamicRange), C.double(f.SilenceRatio), C.double(f.OnsetDensity)
}

func cmpResult(r *C.sonar_cmp_result) CmpResult {
	return CmpResult{
		OverallSimilarity: float64(r.overall_similarity), FeatureSimilarity: float64(r.feature_similarity),
		Confidence: float64(r.confidence), DistMFCC: float64(r.dist_mfcc), DistSpectral: float64(r.dist_spectral),
		DistTemporal: float64(r.dist_temporal), DistHarmonic: float64(r.dist_harmonic),
		ContentTypeMatch: r.content_type_match != 0, NFeatures: int(r.n_features),
	}
}

// Compare replaces the arithmetic of FingerprintComparator.Compare (comparison.go:133-194): weights in the order mfcc,
// spectral, chroma, temporal, speech, harmonic, energy (getEffectiveWeights, comparison.go:1055-1104).
func Compare(f1, f2 *CmpFeatures, weights [7]float64, enableContentFilter bool) (*CmpResult, error) {
	if f1 == nil || f2 == nil {
		return nil, errors.New("fingerprints cannot be nil") // comparison.go:135
	}
	c, err := Ctx()
	if err != nil {
		return nil, err
	}
	mem := (*[2]C.sonar_cmp_features)(C.calloc(2, C.size_t(unsafe.Sizeof(C.sonar_cmp_features{}))))
	defer C.free(unsafe.Pointer(mem))
	var pin runtime.Pinner
	defer pin.Unpin()
	f1.fill(&mem[0], &pin)
	f2.fill(&mem[1], &pin)
	var w C.sonar_cmp_weights
	for i, v := range weights {
		w.w[i] = C.double(v)
	}
	filter := C.int(0)
	if enableContentFilter {
		filter = 1
	}
	var out C.sonar_cmp_result
	if err := locked(func() C.int { return C.sonar_compare_f64(c, &mem[0], &mem[1], &w, filter, &out) }); err != nil {
		return nil, err
	}
	r := cmpResult(&out)
	return &r, nil
}

// CompareBatch replaces the comparison loop of BatchCompare / FindBestMatches (comparison.go:1107-1151, 197-263): one
// query against n candidates (a nil candidate is skipped as in the reference: NFeatures = -1).
func CompareBatch(query *CmpFeatures, candidates []*CmpFeatures, weights [7]float64, enableContentFilter bool) ([]CmpResult, error) {
	if query == nil {
		return nil, errors.New("fingerprints cannot be nil")
	}
	c, err := Ctx()
	if err != nil {
		return nil, err
	}
	n := len(candidates)
	if n == 0 {
		return nil, nil
	}
	sz := C.size_t(unsafe.Sizeof(C.sonar_cmp_features{}))
	mem := C.calloc(C.size_t(n+1), sz)
	defer C.free(mem)
	at := func(i int) *C.sonar_cmp_features { return (*C.sonar_cmp_features)(unsafe.Add(mem, uintptr(i)*uintptr(sz))) }
	ptrs := (*[1 << 28]*C.sonar_cmp_features)(C.calloc(C.size_t(n), C.size_t(unsafe.Sizeof(uintptr(0)))))
	defer C.free(unsafe.Pointer(ptrs))
	var pin runtime.Pinner
	defer pin.Unpin()
	query.fill(at(0), &pin)
	for i, cand := range candidates {
		if cand != nil {
			cand.fill(at(i+1), &pin)
			ptrs[i] = at(i + 1)
		}
	}
	var w C.sonar_cmp_weights
	for i, v := range weights {
		w.w[i] = C.double(v)
	}
	filter := C.int(0)
	if enableContentFilter {
		filter = 1
	}
	res := make([]C.sonar_cmp_result, n)
	if err := locked(func() C.int {
		return C.sonar_compare_batch_f64(c, at(0), (**C.sonar_cmp_features)(unsafe.Pointer(ptrs)), C.int(n), &w, filter, &res[0])
	}); err != nil {
		return nil, err
	}
	out := make([]CmpResult, n)
	for i := range res {
		out[i] = cmpResult(&res[i])
	}
	return out, nil
}

// ---- pinned host memory for Go slices ---------------------------------------------------------------------------------

// RegisterPCM page-locks the backing array of a Go slice for the duration of a batch of calls (cudaHostRegister behind
// sonar_host_register): the H2D copies of the *_f64 entry points then run at full PCIe rate instead of through the
// driver's pageable staging.  The slice must stay alive and must not be resized until UnregisterPCM.
func RegisterPCM(pcm []float64, pin *runtime.Pinner) error {
	if len(pcm) == 0 {
		return nil
	}
	c, err := Ctx()
	if err != nil {
		return err
	}
	pin.Pin(&pcm[0])
	return locked(func() C.int { return C.sonar_host_register(c, unsafe.Pointer(&pcm[0]), C.uint64_t(8*len(pcm))) })
}

// UnregisterPCM undoes RegisterPCM.
func UnregisterPCM(pcm []float64) error {
	if len(pcm) == 0 {
		return nil
	}
	c, err := Ctx()
	if err != nil {
		return err
	}
	return locked(func() C.int { return C.sonar_host_unregister(c, unsafe.Pointer(&pcm[0])) })
}
