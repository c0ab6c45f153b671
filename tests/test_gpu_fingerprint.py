"""GPU parity: GenerateFingerprint's compute (fused STFT + MFCC + spectral, FP64 energy/ZCR, YIN)
against the CPU oracle through the C ABI.

Tolerance (north star: "features within a stated relative tolerance, e.g. 1e-4"): the STFT/MFCC/
spectral kernels compute in FP32, so for every feature array
    |gpu - oracle| <= 1e-4 * max(|oracle|, max|oracle array|)
i.e. 1e-4 relative, floored at 1e-4 of the array's own scale for near-zero entries.  Outputs computed
in FP64 in the reference's summation order (short-time energy, ZCR) must be bit-exact: they feed the
cross-correlation arg-max and the DTW path.  The pitch track's voiced pattern is identical and its values agree to 1e-4
(FP32 transforms for the difference function, float64 re-evaluation in the reference's order of every borderline frame).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FP32_FEATURES = ("mfcc", "spectral_centroid", "spectral_rolloff", "spectral_bandwidth", "spectral_flatness",
                 "spectral_crest", "spectral_slope", "spectral_flux", "low_energy_ratio", "high_energy_ratio")
# Flatness (mean of ln m over ALL bins) and slope (log-log regression over all bins) are dominated by the
# weakest bins of a frame.  The FP32 FFT's noise floor is ~1e-7 of the frame's strongest bin, so on frames
# whose spectrum spans > ~60 dB (a pure tone over a 1e-4 noise floor in the "voiced" cases) those bins carry
# percent-level error; the two features then agree to 2e-3 instead of 1e-4.  Everything else holds 1e-4.
TOL = {"spectral_flatness": 2e-3, "spectral_slope": 2e-3}
# The per-input form of the same statement: both features are linear in ln|X_k|, so an absolute error
# d|X_k| <= ETA * max_j |X_j| of the FP32 transform (ETA = 2^-21: a few ulp of the frame's strongest bin, measured) moves
# them by at most the mean (flatness) / a bounded weighting (slope) of d|X_k| / |X_k|.  log_feature_bound() evaluates that
# bound on the oracle's own magnitudes; 1e-4 + bound is asserted per frame wherever the spectrogram is at hand.
ETA = 2.0 ** -21


def log_feature_bound(oracle, pcm, p):
    """Per-frame bound ETA * max_k|X| * mean_k(1/|X_k|) (bins above the 1e-10 floor) on the relative error of the
    flatness, and on the absolute error of the slope in units of its regression weights (|w_k| <= 4 / B for these grids)."""
    mag, _, _ = oracle.stft(np.asarray(pcm, dtype=np.float64), p.window_size, p.hop_size, p.window_type)
    m = np.where(mag > 1e-10, mag, np.inf)
    return ETA * mag.max(axis=1) * np.mean(1.0 / m, axis=1)


def log_features_close(x, y, bound, name):
    """|x - y| <= (1e-4 + 4 bound_t) * max(|y_t|, max|y|) per frame."""
    scale = np.max(np.abs(y)) if y.size else 0.0
    lim = (1e-4 + 4.0 * bound) * np.maximum(np.abs(y), scale)
    bad = np.abs(x - y) > lim
    assert not bad.any(), f"{name}: {bad.sum()} of {y.size} frames outside 1e-4 + per-input bound; worst {np.max(np.abs(x - y) / lim):.3g}x"
EXACT = ("short_time_energy", "zero_crossing_rate")
FP64_CLOSE = ("energy_entropy",)
PITCH = ("pitch_estimate", "pitch_confidence", "voicing_strength", "harmonic_ratio", "inharmonicity_ratio",
         "tonal_centroid")


def feature_close(x, y, tol=1e-4, name=""):
    assert x.shape == y.shape
    if y.size == 0:
        return
    scale = np.max(np.abs(y))
    bound = tol * np.maximum(np.abs(y), scale)
    bad = np.abs(x - y) > bound
    assert not bad.any(), f"{name}: {bad.sum()} of {y.size} outside tolerance; worst {np.max(np.abs(x - y) / np.maximum(bound, 1e-300)):.3g}x"


# Geometries the third-generation kernel serves (stft_v3.cu): it hands every frame whose FP32 result is not safe -- the
# rolloff bin within the FP32 error of the 85 % threshold, bins or mel bands at the transform's noise floor, magnitudes
# near the 1e-10 validity threshold, silence -- to the float64 re-evaluation in the reference's order
# (spectral_exact.cu).  There the rolloff is BIT-EXACT and flatness / slope hold the plain 1e-4 on every input
# (VERDICT r1 weak #4); the other geometries (first / second generation kernels) keep the stated per-input bound.
STRICT_CASES = {"c1_music_fixed_sr", "c3_speech_512_160_40mel", "hamming", "voiced_yin", "voiced_yin_16k",
                "run_boundaries", "single_frame", "silence"}


def check_fp(a, b, bound=None, strict=False):
    for k in FP32_FEATURES:
        if strict:
            feature_close(a.arrays[k], b.arrays[k], tol=1e-4, name=k)
        elif bound is not None and k in TOL:
            log_features_close(a.arrays[k], b.arrays[k], bound, k)
        else:
            feature_close(a.arrays[k], b.arrays[k], tol=TOL.get(k, 1e-4), name=k)
    if strict:
        assert np.array_equal(a.spectral_rolloff, b.spectral_rolloff), "rolloff bin (discrete) must be bit-exact"
    for k in EXACT:
        assert np.array_equal(a.arrays[k], b.arrays[k]), k
    for k in FP64_CLOSE:
        np.testing.assert_allclose(a.arrays[k], b.arrays[k], rtol=1e-12, atol=1e-15)
    # The pitch track: the difference function runs on FP32 transforms (yin32.cu, values ~1e-6) while every DECISION
    # (voiced or not, which lag) is either outside the FP32 error bound or re-taken in float64 in the reference's order:
    # the voiced pattern must be identical, the values hold the north star's 1e-4.
    assert np.array_equal(a.pitch_estimate > 0, b.pitch_estimate > 0), "voiced / unvoiced pattern"
    for k in PITCH:
        np.testing.assert_allclose(a.arrays[k], b.arrays[k], rtol=1e-4, atol=1e-6, err_msg=k)
    assert a.energy_variance == pytest.approx(b.energy_variance, rel=1e-10)
    assert a.loudness_range == pytest.approx(b.loudness_range, rel=1e-9, abs=1e-12)
    assert a.sizes == b.sizes


def voiced(seconds, sr, f0=440.0, seed=3):
    """Gated tone with vibrato: the reference's windowed YIN (pitch_detection.go:349-420) only dips below
    0.15 on nearly sinusoidal frames, so this is what exercises the pitch track (~60 % voiced frames)."""
    n = int(seconds * sr)
    t = np.arange(n) / sr
    f = f0 * (1.0 + 0.2 * np.sin(2 * np.pi * 0.7 * t))
    ph = 2 * np.pi * np.cumsum(f) / sr
    x = np.sin(ph) + 0.05 * np.sin(2 * ph)
    gate = (np.sin(2 * np.pi * 1.3 * t) > -0.3).astype(float)
    rng = np.random.default_rng(seed)
    return 0.4 * x * gate + 1e-4 * rng.standard_normal(n)


CASES = {
    "c1_music_fixed_sr": (lambda s: s.sweep_noise(4.0, seed=1), dict(algo_sample_rate=44100)),
    "c1_music_parity_sr0": (lambda s: s.sweep_noise(4.0, seed=1), dict(algo_sample_rate=0)),
    "c3_speech_512_160_40mel": (lambda s: s.speech_band_noise(6.0), dict(
        window_size=512, hop_size=160, energy_frame=512, energy_hop=160, algo_sample_rate=16000,
        call_sample_rate=16000, n_mel=40)),
    "c3_speech_stock_26mel_sr0": (lambda s: s.speech_band_noise(3.0), dict(
        window_size=512, hop_size=160, energy_frame=512, energy_hop=160, algo_sample_rate=0,
        call_sample_rate=16000)),
    "w2048_h512": (lambda s: s.sweep_noise(3.0, seed=3), dict(
        window_size=2048, hop_size=512, energy_frame=2048, energy_hop=512, algo_sample_rate=44100)),
    "w256_h64_odd_len": (lambda s: s.sweep_noise(1.0, seed=5)[:40001], dict(
        window_size=256, hop_size=64, energy_frame=256, energy_hop=64, algo_sample_rate=44100)),
    "odd_hop_441": (lambda s: s.sweep_noise(2.0, seed=6), dict(
        window_size=1024, hop_size=441, energy_frame=1024, energy_hop=441, algo_sample_rate=44100)),
    "energy_grid_differs_F4": (lambda s: s.sweep_noise(2.0, seed=7), dict(
        algo_sample_rate=44100, energy_frame=2048, energy_hop=512)),
    "energy_disabled_F4": (lambda s: s.sweep_noise(2.0, seed=7), dict(
        algo_sample_rate=44100, energy_frame=0, energy_hop=0)),
    "no_lifter_20mfcc": (lambda s: s.sweep_noise(2.0, seed=8), dict(
        algo_sample_rate=44100, n_mfcc=20, n_mel=32, use_liftering=0, low_hz=100.0, high_hz=8000.0)),
    "hamming": (lambda s: s.sweep_noise(2.0, seed=9), dict(algo_sample_rate=44100, window_type="hamming")),
    "voiced_yin": (lambda s: voiced(4.0, 44100), dict(algo_sample_rate=44100)),
    "voiced_yin_16k": (lambda s: voiced(4.0, 16000, f0=200.0), dict(
        window_size=512, hop_size=160, energy_frame=512, energy_hop=160, algo_sample_rate=16000,
        call_sample_rate=16000)),
    # the second-generation STFT kernel's other hop instantiations (64 / 128 / 512) and a denser mel bank
    "w1024_h64": (lambda s: s.sweep_noise(1.0, seed=11), dict(
        window_size=1024, hop_size=64, energy_frame=1024, energy_hop=64, algo_sample_rate=44100)),
    "w1024_h128": (lambda s: s.sweep_noise(1.5, seed=12), dict(
        window_size=1024, hop_size=128, energy_frame=1024, energy_hop=128, algo_sample_rate=44100)),
    "w1024_h512": (lambda s: s.sweep_noise(3.0, seed=13), dict(
        window_size=1024, hop_size=512, energy_frame=1024, energy_hop=512, algo_sample_rate=44100)),
    "w1024_h882_generic_hop": (lambda s: s.sweep_noise(3.0, seed=16), dict(
        window_size=1024, hop_size=882, energy_frame=1024, energy_hop=882, algo_sample_rate=44100)),
    "w1024_h1024_no_overlap": (lambda s: s.sweep_noise(3.0, seed=17), dict(
        window_size=1024, hop_size=1024, energy_frame=1024, energy_hop=1024, algo_sample_rate=44100)),
    "w1024_40mel_22k": (lambda s: s.sweep_noise(2.0, seed=14), dict(
        algo_sample_rate=22050, call_sample_rate=22050, n_mel=40)),
    "run_boundaries": (lambda s: s.sweep_noise(1.0, seed=15)[:1024 + 256 * 62], dict(algo_sample_rate=44100)),
    "single_frame": (lambda s: s.sweep_noise(1.0, seed=10)[:1024], dict(algo_sample_rate=44100)),
    "silence": (lambda s: np.zeros(20000), dict(algo_sample_rate=44100)),
    # n in (W - H, W): Go's truncating division yields ONE frame that the STFT worker then skips (spectral.go:409,
    # 472-474) -> a zero spectrum row, ZCR over the n samples, no energy frame; n in (512, 1024): one pitch frame of zeros
    "shorter_than_window_900": (lambda s: s.sweep_noise(1.0, seed=21)[:900], dict(algo_sample_rate=44100)),
    "shorter_than_window_769": (lambda s: s.sweep_noise(1.0, seed=22)[:769], dict(algo_sample_rate=44100)),
    "short_window_ok_pitch_short_700": (lambda s: s.sweep_noise(1.0, seed=23)[:700], dict(
        window_size=512, hop_size=160, energy_frame=512, energy_hop=160, algo_sample_rate=16000, call_sample_rate=16000)),
    "shorter_than_window_energy_grid_fits": (lambda s: s.sweep_noise(1.0, seed=24)[:1000], dict(
        algo_sample_rate=44100, energy_frame=256, energy_hop=128)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_fingerprint_matches_oracle(gpu, oracle, synth, name):
    make, kw = CASES[name]
    pcm = make(synth)
    p = gpu.default_params(**kw)
    bound = log_feature_bound(oracle, pcm, p) if pcm.size >= p.window_size else None
    check_fp(gpu.fingerprint(pcm, p), oracle.fingerprint(pcm, p), bound, strict=name in STRICT_CASES)


def test_weak_bins_take_the_float64_path(gpu, oracle):
    """A tone 74 dB over its noise floor: every frame has bins at the FP32 transform's noise floor, so every frame is
    listed and redone in float64 in the reference's order -- centroid / rolloff / bandwidth / crest / band ratios come
    out BIT-EXACT, flatness / slope / MFCC to the last bits of log() (they held only 2e-3 in FP32).  Broadband input
    (the BASELINE workloads) lists ~1 % of the frames: those whose rolloff threshold falls within the FP32 error."""
    sr = 44100
    t = np.arange(int(3.0 * sr)) / sr
    x = 0.5 * np.sin(2 * np.pi * 440.0 * t) + 1e-4 * np.random.default_rng(5).standard_normal(t.size)
    p = gpu.default_params(algo_sample_rate=sr, call_sample_rate=sr)
    g = gpu.fingerprint(x, p)
    listed, _ = gpu.exact_counts()
    o = oracle.fingerprint(x, p)
    T = g.mfcc.shape[0]
    assert listed == T
    for k in ("spectral_centroid", "spectral_rolloff", "spectral_bandwidth", "spectral_crest", "low_energy_ratio",
              "high_energy_ratio"):
        assert np.array_equal(g.arrays[k], o.arrays[k]), k
    for k in ("mfcc", "spectral_flatness", "spectral_slope"):
        np.testing.assert_allclose(g.arrays[k], o.arrays[k], rtol=1e-11, atol=1e-13, err_msg=k)
    noise = np.random.default_rng(6).standard_normal(int(10.0 * sr)) * 0.2
    gn = gpu.fingerprint(noise, p)
    listed, _ = gpu.exact_counts()
    on = oracle.fingerprint(noise, p)
    assert 0 < listed < 0.03 * gn.mfcc.shape[0]
    assert np.array_equal(gn.spectral_rolloff, on.spectral_rolloff)


ALL_WINDOWS = ("hann", "hamming", "blackman", "blackman_harris", "kaiser", "tukey", "rectangular", "bartlett", "welch")


@pytest.mark.parametrize("wtype", ALL_WINDOWS)
def test_every_window_type_matches_oracle(gpu, oracle, synth, capi, wtype):
    """All nine analyzers.WindowType values (windowing.go:246-371) through the product library: the fused fingerprint
    (1024/256 second-generation kernel and the 512/160 geometry) and the materialised spectrum."""
    assert wtype in capi.WINDOWS
    pcm = synth.sweep_noise(1.5, seed=40)
    for kw in (dict(algo_sample_rate=44100, window_type=wtype),
               dict(window_size=512, hop_size=160, energy_frame=512, energy_hop=160, algo_sample_rate=44100,
                    window_type=wtype)):
        p = gpu.default_params(**kw)
        check_fp(gpu.fingerprint(pcm, p), oracle.fingerprint(pcm, p))
    mg, _, _ = gpu.stft(pcm[:20000], 1024, 256, wtype=wtype)
    mr, _, _ = oracle.stft(pcm[:20000], 1024, 256, wtype=wtype)
    assert np.max(np.abs(mg - mr)) <= 2e-6 * np.max(mr)


def test_short_input_batch_does_not_read_the_neighbour(gpu, oracle, synth):
    """with n in (W - H, W) the lone frame must not be read past the stream's end, i.e. into the next stream
    of a batch.  The neighbour is loud, the short streams are quiet: any leak shows up in the spectral features."""
    p = gpu.default_params(algo_sample_rate=44100)
    quiet = 1e-3 * synth.sweep_noise(1.0, seed=50)[:900]
    loud = 100.0 * synth.sweep_noise(1.0, seed=51)[:900]
    batch = gpu.fingerprint_batch([quiet, loud, quiet], p)
    ref = oracle.fingerprint(quiet, p)
    for fb in (batch[0], batch[2]):
        check_fp(fb, ref)
        assert not fb.spectral_flatness.any() and not fb.spectral_crest.any()
        assert fb.mfcc.shape == (1, 13) and fb.short_time_energy.size == 0
        assert fb.pitch_estimate.size == 1 and fb.pitch_estimate[0] == 0.0 and fb.inharmonicity_ratio[0] == 1.0  # n in (512, 1024)
    mg, ph, cx = gpu.stft(quiet, 1024, 256, phase=True, cplx=True)
    assert mg.shape == (1, 513) and not mg.any() and not ph.any() and not cx.any()


@pytest.mark.gpu
def test_voiced_case_exercises_yin(oracle, synth):
    p = oracle.default_params(algo_sample_rate=44100)
    fp = oracle.fingerprint(voiced(4.0, 44100), p)
    assert (fp.pitch_estimate > 0).mean() > 0.3  # otherwise the YIN parity case above proves nothing


def test_parity_mode_degenerate_constants(gpu, synth):
    """stock GenerateFingerprint builds every algorithm with sampleRate=0."""
    fp = gpu.fingerprint(synth.sweep_noise(2.0, seed=1), gpu.default_params(algo_sample_rate=0))
    assert np.allclose(fp.mfcc[:, 0], -117.40926320884498, rtol=1e-6)
    assert np.all(np.abs(fp.mfcc[:, 1:]) < 1e-3)
    for k in ("spectral_centroid", "spectral_rolloff", "spectral_bandwidth", "spectral_slope", "zero_crossing_rate",
              "pitch_estimate", "harmonic_ratio"):
        assert not fp.arrays[k].any(), k
    assert np.all(fp.inharmonicity_ratio == 1.0)
    assert fp.spectral_flatness.any() and fp.spectral_flux.any() and fp.short_time_energy.any()


def test_batch_ragged_matches_single(gpu, synth):
    p = gpu.default_params(algo_sample_rate=44100)
    pcms = [synth.sweep_noise(s, seed=20 + i) for i, s in enumerate((1.0, 2.5, 1.0, 0.3, 2.5))]
    batch = gpu.fingerprint_batch(pcms, p)
    for x, fb in zip(pcms, batch):
        fs = gpu.fingerprint(x, p)
        for k in fs.arrays:
            assert np.array_equal(fs.arrays[k], fb.arrays[k]), k  # same kernels, same order: identical


def test_seams_between_runs_are_invisible(gpu, oracle, synth):
    """A long stream is cut into per-warp runs of frames (flux needs frame t-1 across the seam)."""
    pcm = synth.sweep_noise(20.0, seed=31)
    p = gpu.default_params(algo_sample_rate=44100)
    a, b = gpu.fingerprint(pcm, p), oracle.fingerprint(pcm, p)
    feature_close(a.spectral_flux, b.spectral_flux)
    feature_close(a.mfcc, b.mfcc)


def test_stft_materialised(gpu, oracle, synth):
    pcm = synth.sweep_noise(1.0, seed=7)
    for win, hop in ((1024, 256), (512, 160), (256, 100), (2048, 512)):
        mg, ph, cx = gpu.stft(pcm, win, hop, phase=True, cplx=True)
        mr, pr, cr = oracle.stft(pcm, win, hop, phase=True, cplx=True)
        scale = np.max(mr)
        assert np.max(np.abs(mg - mr)) <= 2e-6 * scale
        assert np.max(np.abs(cx - cr)) <= 2e-6 * scale
        strong = mr > 1e-3 * scale  # phase is only meaningful where the bin is above the FP32 noise floor
        d = np.angle(np.exp(1j * (ph - pr)))
        assert np.max(np.abs(d[strong])) < 1e-3


def test_fingerprint_errors(gpu, capi):
    p = gpu.default_params()
    for pcm, code, text in ((np.zeros(100), capi.ERR_TOO_SHORT, "signal too short for given window size and hop size"),
                            (np.zeros(0), capi.ERR_EMPTY, "empty signal")):
        with pytest.raises(capi.SonarError) as e:
            gpu.fingerprint(pcm, p)
        assert e.value.code == code and text in e.value.msg
    with pytest.raises(capi.SonarError) as e:
        gpu.fingerprint(np.zeros(5000), gpu.default_params(window_size=3000))  # 1000 is served (float64 route)
    assert e.value.code == capi.ERR_UNSUPPORTED
    with pytest.raises(capi.SonarError) as e:
        gpu.fingerprint(np.zeros(5000), gpu.default_params(call_sample_rate=0))
    assert "sample rate must be positive" in e.value.msg


def test_kernel_launch_counter(gpu, synth):
    before = gpu.kernel_launches()
    gpu.fingerprint(synth.sweep_noise(1.0, seed=1), gpu.default_params(algo_sample_rate=44100))
    assert gpu.kernel_launches() > before


def test_loudness_range_many_windows(gpu, oracle, synth):
    """> 4096 loudness windows (long 16 kHz streams) take the radix-select path; same value as the oracle's sort."""
    sr = 16000
    # (11.2e6 - 6400) / 1600 + 1 = 6997 windows; loud enough for positive loudness units (-0.691 + 10 log10 e^2)
    pcm = 20.0 * synth.speech_band_noise(700.0, sr=sr, seed=12)
    p = gpu.default_params(window_size=512, hop_size=160, energy_frame=512, energy_hop=160, algo_sample_rate=sr,
                           call_sample_rate=sr)
    a, b = gpu.fingerprint(pcm, p), oracle.fingerprint(pcm, p)
    assert b.loudness_range > 0
    assert a.loudness_range == pytest.approx(b.loudness_range, rel=1e-10)
    assert np.array_equal(a.short_time_energy, b.short_time_energy)


@pytest.mark.parametrize("algo_sr", [16000, 0])
def test_temporal_feature_group(gpu, oracle, synth, capi, algo_sr):
    """speech.go:370-408 (news / talk): silence ratio through an exact order statistic instead of the O(T^2)
    bubble sort, onset peak-pick, attack times (NaN / 0.1 artefacts of sampleRate = 0 included), 512/256 envelope."""
    sr = 16000
    rng = np.random.default_rng(5)  # noise bursts with abrupt onsets over a quiet floor
    n = 20 * sr
    pcm = 0.02 * rng.standard_normal(n)
    t = 0
    while t < n:
        on, off = int(rng.uniform(0.05, 0.2) * sr), int(rng.uniform(0.2, 0.6) * sr)
        if t + on <= n:
            pcm[t:t + on] += rng.uniform(0.2, 0.8) * rng.standard_normal(on)
        t += on + off
    kw = dict(window_size=512, hop_size=160, energy_frame=512, energy_hop=160, algo_sample_rate=algo_sr,
              call_sample_rate=sr, enable=capi.FP_ENABLE_MFCC | capi.FP_ENABLE_TEMPORAL)
    a, b = gpu.fingerprint(pcm, gpu.default_params(**kw)), oracle.fingerprint(pcm, oracle.default_params(**kw))
    check_fp(a, b)
    assert b.n_attack_time > 5, "the case must contain onsets"
    assert a.n_attack_time == b.n_attack_time
    assert np.array_equal(a.rms_energy, b.rms_energy) and np.array_equal(a.rms_energy, a.short_time_energy)
    np.testing.assert_allclose(a.envelope_shape, b.envelope_shape, rtol=1e-12)
    assert np.array_equal(a.attack_time, b.attack_time, equal_nan=True)
    for k in ("silence_ratio", "peak_amplitude", "dynamic_range"):
        assert a.scalars[k] == b.scalars[k], k
    for k in ("average_amplitude", "onset_density"):
        assert a.scalars[k] == pytest.approx(b.scalars[k], rel=1e-12), k


@pytest.mark.parametrize("case", [
    dict(make=lambda s: s.sweep_noise(2.0, seed=41), win=1024, hop=256, sr=44100, n_bands=6, n_bark=24, tol_db=1e-3),
    dict(make=lambda s: s.speech_band_noise(3.0), win=512, hop=160, sr=16000, n_bands=6, n_bark=18, tol_db=1e-3),
    dict(make=lambda s: voiced(2.0, 44100), win=2048, hop=512, sr=44100, n_bands=4, n_bark=24, tol_db=0.02),
    dict(make=lambda s: s.sweep_noise(25.0, seed=42), win=1024, hop=256, sr=44100, n_bands=6, n_bark=24,
         tol_db=1e-3),  # more than one 4096-frame chunk
], ids=["music_1024", "speech_512", "voiced_2048", "chunked"])
def test_music_spectral_matches_oracle(gpu, oracle, synth, case):
    """spectral contrast, chroma folding and Bark-band energies of the music extractor on the GPU
    spectrogram.  Chroma and Bark are linear in the FP32 power spectrum: 1e-4 of the array's scale.  The contrast is
    10*log10(peak / valley) where the valley is the mean of the weakest 20 % of a band's bins, i.e. it sits at the
    FP32 FFT's noise floor on frames with > 60 dB of dynamic range (the same effect as spectral flatness):
    0.02 dB absolute there, 1e-3 dB on broadband frames."""
    x = case["make"](synth)
    kw = dict(win=case["win"], hop=case["hop"], sample_rate=case["sr"], n_bands=case["n_bands"], n_bark=case["n_bark"],
              bark_low=50.0, bark_high=0.45 * case["sr"])
    gc, gh, gb = gpu.music_spectral(x, **kw)
    oc, oh, ob = oracle.music_spectral(x, **kw)
    assert gc.shape == oc.shape and gh.shape == oh.shape and gb.shape == ob.shape
    feature_close(gh, oh, tol=1e-4, name="chroma")
    feature_close(gb, ob, tol=1e-4, name="bark")
    assert np.max(np.abs(gc - oc)) <= case["tol_db"], np.max(np.abs(gc - oc))


def test_music_spectral_edge_cases(gpu, oracle):
    c, h, b = gpu.music_spectral(np.zeros(4096), sample_rate=44100)
    assert not c.any() and not h.any() and not b.any()
    for lib in (gpu, oracle):
        with pytest.raises(Exception, match="signal too short"):
            lib.music_spectral(np.zeros(100), sample_rate=44100)
        with pytest.raises(Exception, match="sample rate must be positive"):
            lib.music_spectral(np.zeros(4096), sample_rate=0)


@pytest.mark.parametrize("dtype", [np.int16, np.float32])
def test_batch_pcm_ingest_equals_the_widened_float64_call(gpu, oracle, synth, dtype):
    """sonar_fingerprint_batch_pcm: ragged int16 / float32 batches widened on the device give exactly
    the float64 call's results on the samples the reference's decoder would have produced."""
    p = gpu.default_params(algo_sample_rate=44100)
    raw = [synth.sweep_noise(s, seed=60 + i) for i, s in enumerate((1.0, 2.3, 1.0, 0.4))]
    if dtype == np.int16:
        narrow = [np.clip(np.round(x * 20000.0), -32768, 32767).astype(np.int16) for x in raw]
        wide = [x.astype(np.float64) / 32768.0 for x in narrow]
    else:
        narrow = [x.astype(np.float32) for x in raw]
        wide = [x.astype(np.float64) for x in narrow]
    got = gpu.fingerprint_batch_pcm(narrow, p)
    ref = gpu.fingerprint_batch(wide, p)
    ora = oracle.fingerprint_batch_pcm(narrow, p)
    for g, r, o in zip(got, ref, ora):
        for k in r.arrays:
            assert np.array_equal(g.arrays[k], r.arrays[k]), k
        assert np.array_equal(g.short_time_energy, o.short_time_energy)  # bit-exact against the oracle's widening
        assert g.energy_variance == r.energy_variance and g.loudness_range == r.loudness_range


@pytest.mark.parametrize("sr,n_mel,n_mfcc,low,high", [
    (8000, 20, 13, 0.0, 0.0), (8000, 64, 20, 50.0, 3800.0), (11025, 26, 13, 0.0, 0.0), (16000, 40, 13, 20.0, 7600.0),
    (22050, 64, 32, 0.0, 0.0), (32000, 32, 12, 300.0, 3400.0), (44100, 64, 13, 0.0, 0.0), (44100, 13, 13, 1000.0, 2000.0),
    (48000, 26, 13, 0.0, 0.0), (48000, 48, 24, 100.0, 20000.0), (96000, 26, 13, 0.0, 0.0), (44100, 26, 13, 21000.0, 22000.0),
])
def test_mel_bank_sweep_1024(gpu, oracle, synth, sr, n_mel, n_mfcc, low, high):
    """Mel banks from very dense (64 filters at 8 kHz: zero-width regions, the first-generation kernel's job) to very
    sparse / narrow ones, at N = 1024: whichever kernel the eligibility check picks must match the oracle."""
    kw = dict(algo_sample_rate=sr, call_sample_rate=sr, n_mel=n_mel, n_mfcc=n_mfcc)
    if low or high:
        kw.update(low_hz=low, high_hz=high)
    p = gpu.default_params(**kw)
    x = synth.sweep_noise(1.0, seed=90 + n_mel)[: 1024 + 256 * 70]
    check_fp(gpu.fingerprint(x, p), oracle.fingerprint(x, p))


@pytest.mark.parametrize("win,hop", [(512, 160), (1024, 256), (2048, 3000)])
def test_stft_streamer_equals_the_batch_transform(gpu, oracle, win, hop):
    """sonar_stft_stream_* (SURVEY §8 f3): chunked input gives the frames of the reference's streaming loop; each frame
    equals the GPU batch STFT of the same samples bit for bit and the oracle within the STFT tolerance."""
    from stream_model import frame_starts
    rng = np.random.default_rng(win * 7 + hop)
    x = rng.standard_normal(60000)
    chunks = [5, win - 1, 1, 4000, 0, 17000, hop, 2 * win + 3]
    chunks.append(x.size - sum(chunks))
    starts, left = frame_starts(chunks, win, hop)
    g, o = gpu.stft_stream(win, hop), oracle.stft_stream(win, hop)
    pos = 0
    for c, want in zip(chunks, starts):
        mg, pg, cg = g.process_chunk(x[pos:pos + c])
        mo, po, co = o.process_chunk(x[pos:pos + c])
        pos += c
        assert mg.shape == mo.shape and mg.shape[0] == len(want)
        if want:
            scale = np.max(np.abs(mo))
            assert np.allclose(mg, mo, rtol=1e-4, atol=2e-6 * scale)
            assert np.allclose(cg, co, rtol=1e-4, atol=2e-6 * scale)
            m1, _, _ = gpu.stft(x[want[0]:want[0] + win], win, hop)
            assert np.array_equal(mg[0], m1[0])
    assert g.buffered() == o.buffered() == left
    g.close()
    o.close()


@pytest.mark.parametrize("n", [4410 * 1024 + 30000, 17640 + 4410 * 3, 17640 + 4410 * 3 + 255, 40001, 44100 * 3 + 1])
def test_loudness_from_the_frame_walk_block_sums(gpu, oracle, synth, n):
    """At 44.1 kHz the loudness windows' RMS comes from block sums the frame walk leaves behind (csrc/timedomain.cu,
    WalkBlocks).  Lengths: one that contains sample 4410 * 1024 (a loudness block boundary that is also the first
    sample of a walk thread), a stream that ends exactly with its last window, streams whose tail lies behind the last frame."""
    rng = np.random.default_rng(n % 1000)
    x = 0.3 * rng.standard_normal(n) * (1.0 + 0.8 * np.sin(np.arange(n) * 2e-4))
    p = gpu.default_params(algo_sample_rate=44100)
    a, b = gpu.fingerprint(x, p), oracle.fingerprint(x, p)
    assert b.loudness_range >= 0.0
    assert a.loudness_range == pytest.approx(b.loudness_range, rel=1e-9, abs=1e-12)
    assert np.array_equal(a.short_time_energy, b.short_time_energy)
    assert np.array_equal(a.zero_crossing_rate, b.zero_crossing_rate)


def test_yin_threshold_bisection_decisions_match_the_oracle(gpu, oracle):
    """Adversarial for the 0.15 threshold.  A 440 Hz tone plus a fixed noise realisation scaled by
    `amp`: the oracle's voiced / unvoiced decision for the frame flips at some amplitude, and a bisection on `amp` drives
    the CMNDF dip to within rounding of the threshold (the last iterations differ by single float64 ulps of amp).  At
    EVERY amplitude visited the GPU must take the oracle's decision and reproduce its values: frames whose comparison is
    inside the FP32 error bound are re-evaluated in float64 in the reference's summation order (yin32.cu)."""
    sr, n = 44100, 1024 + 512
    t = np.arange(n) / sr
    tone = 0.5 * np.sin(2 * np.pi * 440.0 * t)
    noise = np.random.default_rng(5).standard_normal(n)
    p = gpu.default_params(algo_sample_rate=sr)

    def both(amp):
        x = tone + amp * noise
        a, b = gpu.fingerprint(x, p), oracle.fingerprint(x, p)
        assert np.array_equal(a.pitch_confidence > 0, b.pitch_confidence > 0), amp
        np.testing.assert_allclose(a.pitch_confidence, b.pitch_confidence, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(a.pitch_estimate, b.pitch_estimate, rtol=1e-4, atol=1e-6)
        return bool(b.pitch_confidence[0] > 0)

    lo, hi = 1e-5, 5e-3
    assert both(lo) and not both(hi)  # voiced with little noise, unvoiced with more
    for _ in range(70):
        mid = 0.5 * (lo + hi)
        if mid == lo or mid == hi:
            break
        if both(mid):
            lo = mid
        else:
            hi = mid
    assert hi - lo <= 4 * np.spacing(lo)  # the search really ended at float64 resolution


def test_yin_noise_sweep_decisions_match_the_oracle(gpu, oracle):
    """The same statistically: 40 s of a 440 Hz tone whose noise level creeps through the range where the dip crosses
    0.15, 3,444 frames; the voiced pattern must equal the oracle's exactly."""
    sr, secs = 44100, 40.0
    n = int(sr * secs)
    t = np.arange(n) / sr
    rng = np.random.default_rng(5)
    amp = np.linspace(1.0e-4, 4.5e-4, n)
    x = 0.5 * np.sin(2 * np.pi * 440.0 * t) + amp * rng.standard_normal(n)
    p = gpu.default_params(algo_sample_rate=sr)
    a, b = gpu.fingerprint(x, p), oracle.fingerprint(x, p)
    va, vb = a.pitch_confidence > 0, b.pitch_confidence > 0
    assert 0.05 < vb.mean() < 0.95  # the sweep really crosses the threshold
    assert np.array_equal(va, vb)
    assert np.array_equal(a.pitch_estimate > 0, b.pitch_estimate > 0)
    np.testing.assert_allclose(a.pitch_confidence, b.pitch_confidence, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(a.pitch_estimate, b.pitch_estimate, rtol=1e-4, atol=1e-6)


def _speechy(seconds=6.0, sr=16000, seed=7):
    n = int(seconds * sr)
    t = np.arange(n) / sr
    f0 = 140.0 * (1.0 + 0.05 * np.sin(2 * np.pi * 0.5 * t))
    ph = 2 * np.pi * np.cumsum(f0) / sr
    x = np.sin(ph) + 0.5 * np.sin(2 * ph) + 0.3 * np.sin(3 * ph)
    gate = ((t % 1.5) < 1.0).astype(float)
    return (0.3 * x + 2e-3 * np.random.default_rng(seed).standard_normal(n)) * gate


@pytest.mark.parametrize("case", ["speech_16k", "speech_44k_1024", "noise_not_speech", "too_quiet", "short_600"])
def test_speech_feature_group_matches_oracle(gpu, oracle, case):
    """SURVEY §8 f1 / VERDICT r1 missing #3: EnableSpeechFeatures (extractors/speech.go:194-205,271-317) -- the IsSpeech
    gate, the voicing sweep with the detector history it leaves to the harmonic block, the spectral tilt (bit-exact sums,
    log10 to the last bits), pause durations and speech rate."""
    sr = 44100 if case == "speech_44k_1024" else 16000
    if case in ("speech_16k", "speech_44k_1024"):
        x = _speechy(5.0, sr)
    elif case == "noise_not_speech":
        x = 0.2 * np.random.default_rng(3).standard_normal(2 * sr)
    elif case == "too_quiet":
        x = 1e-4 * _speechy(3.0, sr)
    else:
        x = _speechy(1.0, sr)[:600]
    kw = dict(algo_sample_rate=sr, call_sample_rate=sr)
    if sr == 16000:
        kw.update(window_size=512, hop_size=160, energy_frame=512, energy_hop=160, n_mel=40)
    p = gpu.default_params(**kw)
    if case == "short_600":
        p = gpu.default_params(window_size=512, hop_size=160, energy_frame=512, energy_hop=160, algo_sample_rate=sr,
                               call_sample_rate=sr)
    g, gs = gpu.fingerprint_speech(x, p)
    o, os_ = oracle.fingerprint_speech(x, p)
    assert gs["is_speech"] == os_["is_speech"] == (case in ("speech_16k", "speech_44k_1024"))
    assert gs["n_pause"] == os_["n_pause"] and gs["speech_rate"] == os_["speech_rate"]
    assert np.array_equal(gs["pause_duration"], os_["pause_duration"])
    assert gs["voicing_probability"].shape == os_["voicing_probability"].shape
    np.testing.assert_allclose(gs["voicing_probability"], os_["voicing_probability"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gs["spectral_tilt"], os_["spectral_tilt"], rtol=1e-13, atol=1e-13)
    # the harmonic block saw the history the sweep left behind: same voiced pattern and values as the oracle's run
    assert np.array_equal(g.pitch_estimate > 0, o.pitch_estimate > 0)
    np.testing.assert_allclose(g.pitch_estimate, o.pitch_estimate, rtol=1e-4, atol=1e-6)
    assert np.array_equal(g.short_time_energy, o.short_time_energy)
    if gs["is_speech"]:
        plain = gpu.fingerprint(x, p)
        assert gs["n_pause"] >= 2
        assert np.array_equal(g.pitch_estimate[3:], plain.pitch_estimate[3:])


@pytest.mark.parametrize("win,hop,sr", [(400, 160, 16000), (1000, 250, 44100), (1102, 441, 44100), (1023, 256, 44100),
                                        (128, 32, 16000), (24, 7, 8000)])
def test_window_lengths_without_a_fused_kernel_take_go_dsps_route(gpu, oracle, synth, win, hop, sr):
    """SURVEY §8 f4 / VERDICT r1 missing #4: any window length (analyzers/spectral.go:125-132 hands the frame to
    fft.FFTReal, which runs Bluestein's algorithm when the length is not a power of two).  These lengths have no fused
    FP32 kernel: every frame goes through the float64 evaluation in the reference's order (spectral_exact.cu) with
    go-dsp's transform -- chirp, two radix-2 transforms of NextPowerOf2(2 n - 1) points, the same butterfly graph as the
    oracle -- so the spectrum, rolloff, centroid, bandwidth, crest and the flux come out bit-exact."""
    x = synth.sweep_noise(1.5, sr=sr, seed=31) if win > 100 else synth.sweep_noise(0.2, sr=sr, seed=32)
    mg, pg, cg = gpu.stft(x, win, hop, "hann", phase=True, cplx=True)
    mo, po, co = oracle.stft(x, win, hop, "hann", phase=True, cplx=True)
    assert mg.shape == mo.shape == ((x.size - win) // hop + 1, win // 2 + 1)
    assert np.array_equal(mg, mo) and np.array_equal(cg, co)
    np.testing.assert_allclose(pg, po, rtol=0, atol=1e-12)
    p = gpu.default_params(window_size=win, hop_size=hop, energy_frame=win, energy_hop=hop, algo_sample_rate=sr,
                           call_sample_rate=sr, n_mel=min(26, max(4, win // 8)))
    g, o = gpu.fingerprint(x, p), oracle.fingerprint(x, p)
    for k in ("spectral_centroid", "spectral_rolloff", "spectral_bandwidth", "spectral_crest", "spectral_flux",
              "low_energy_ratio", "high_energy_ratio", "short_time_energy", "zero_crossing_rate"):
        assert np.array_equal(g.arrays[k], o.arrays[k]), k
    for k in ("mfcc", "spectral_flatness", "spectral_slope"):
        np.testing.assert_allclose(g.arrays[k], o.arrays[k], rtol=1e-11, atol=1e-12, err_msg=k)
    assert np.array_equal(g.pitch_estimate > 0, o.pitch_estimate > 0)


def test_unsupported_window_lengths_say_so(gpu, capi, synth):
    x = synth.sweep_noise(0.5, seed=33)
    for win in (4, 3000, 4096):
        p = gpu.default_params(window_size=win, hop_size=max(1, win // 4), energy_frame=win, energy_hop=max(1, win // 4),
                               algo_sample_rate=44100)
        with pytest.raises(capi.SonarError, match="window size"):
            gpu.fingerprint(x, p)


@pytest.mark.parametrize("switch", ["SONAR_STFT_V3", "SONAR_STFT_V4"])
def test_other_generations_of_the_fused_stft_kernel_match_oracle(switch):
    """The default fused STFT is the transform / scan kernel pair (stft_v5.cu).  The single-role kernel (stft_v3_kernel,
    SONAR_STFT_V3=1: also the fallback when the pair's workspace or a mel bank's slots do not fit) and its warp-specialised
    form (stft_v4_kernel, SONAR_STFT_V4=1: a measurement variant) must hold the same parity cases, the
    seam test and the ragged batch.  The switch is read once per process, hence the subprocess."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env[switch] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_fingerprint.py"), "-m", "gpu",
                        "-x", "-q", "-k", "test_fingerprint_matches_oracle or test_seams_between_runs or "
                        "test_batch_ragged or test_every_window_type or test_weak_bins"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_stft_kernel_pair_in_small_workspace_groups(synth):
    """stft_v5.cu processes the streams of a batch in groups that fit its row workspace (SONAR_STFT_WS_MB): with a 1 MB
    cap every stream is its own group (one transform + scan launch each); the batch must equal the single calls."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SONAR_STFT_WS_MB="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_fingerprint.py"), "-m", "gpu",
                        "-x", "-q", "-k", "test_batch_ragged or test_seams_between_runs or test_short_input_batch"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
