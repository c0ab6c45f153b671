"""BASELINE config[1] at its full size (5-min 44.1 kHz source/CDN pair, known 7.3 s offset, +-60 s lag, band 50)
through the chained C-ABI call, checked with size-independent properties and against the oracle on a prefix."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c2_full_size_known_offset_and_path_properties(gpu, oracle, synth):
    sr, hop, secs, off = 44100, 256, 300.0, 7.3
    q, r = synth.aligned_pair(secs, offset_seconds=off, sr=sr, seed=200)
    p = gpu.default_params(algo_sample_rate=sr, call_sample_rate=sr)
    nl, dl = gpu.align_pairs_sizes(p, q.size, 60.0)
    assert nl == 20671 and dl == 51676 - 10335  # T = 51,676 frames, L = 10,335
    res = gpu.align_pairs([q], [r], p, 60.0, 50)[0]
    # the CDN copy carries the same content 7.3 s later: positive lag of 321,930 / 256 = 1257.5 frames
    assert res["xcorr"].peak_lag in (1257, 1258)
    assert abs(res["corr_alignment"].offset_seconds - off) < 0.01
    corr = res["corr"]
    assert corr.shape == (nl,) and int(np.argmax(np.abs(corr))) == res["xcorr"].peak_index
    assert np.all(np.abs(corr) <= 1.0 + 1e-12)
    # DTW path: starts at (0, 0), ends at (dl-1, dl-1), monotone unit steps, inside the band, costs consistent
    pq, pr, pc = res["path_query"], res["path_ref"], res["path_cost"]
    assert (pq[0], pr[0]) == (0, 0) and (pq[-1], pr[-1]) == (dl - 1, dl - 1)
    dq, dr = np.diff(pq), np.diff(pr)
    assert np.all((dq >= 0) & (dq <= 1) & (dr >= 0) & (dr <= 1) & (dq + dr >= 1))
    assert np.all(np.abs(pq - pr) <= 50) and dl <= pq.size <= 2 * dl
    assert res["distance"] == pytest.approx(res["total_cost"] / pq.size, rel=1e-15)
    # frames are independent of what follows them: the oracle on a 10 s prefix pins the first frames of the full run
    n10 = 10 * sr
    oq = oracle.fingerprint(q[:n10], p)
    t = oq.short_time_energy.size - 4
    fq = res["query"]
    assert np.array_equal(fq.short_time_energy[:t], oq.short_time_energy[:t])
    assert np.array_equal(fq.zero_crossing_rate[:t], oq.zero_crossing_rate[:t])
    scale = np.max(np.abs(oq.mfcc))
    assert np.allclose(fq.mfcc[:t], oq.mfcc[:t], rtol=1e-4, atol=1e-4 * scale)
    tp = oq.pitch_estimate.size - 25  # the pitch tracker looks 20 frames back only
    assert np.allclose(fq.pitch_estimate[:tp], oq.pitch_estimate[:tp], rtol=1e-4, atol=1e-6)
    # frames are also independent of how the kernel cuts the stream into runs of 31 frames / rings of 1024 samples:
    # a fresh fingerprint of the PCM from frame 30,011 on reproduces the long run's frames bit for bit (except the
    # first one, whose pre-emphasis sees x[-1] = 0 in the fresh run; flux additionally needs its predecessor)
    k = 30011
    tail = gpu.fingerprint(q[k * hop: k * hop + 12 * sr], p)
    m = tail.mfcc.shape[0] - 2
    # (the FP64 walks are bit-identical; the FP32 features agree to rounding only: the third-generation kernel transforms
    #  frames in pairs a + i b, so a frame's rounding depends on which neighbour it is paired with in a given run)
    for name in ("short_time_energy", "zero_crossing_rate"):
        assert np.array_equal(fq.arrays[name][k + 1: k + m], tail.arrays[name][1:m]), name
    for name in ("mfcc", "spectral_centroid", "spectral_rolloff", "spectral_bandwidth", "spectral_flatness"):
        x, y = fq.arrays[name][k + 1: k + m], tail.arrays[name][1:m]
        assert np.all(np.abs(x - y) <= 1e-4 * np.maximum(np.abs(y), np.max(np.abs(y)))), name
    x, y = fq.spectral_flux[k + 1: k + m - 1], tail.spectral_flux[1: m - 1]
    assert np.all(np.abs(x - y) <= 1e-4 * np.max(np.abs(y)))


def test_c3_full_size_one_hour_of_speech_band_noise(gpu, oracle, synth):
    """BASELINE config[2] at full size: 1 h @ 16 kHz, 512/160 frames, 40-mel / 13-MFCC.  Frames do not depend on what
    follows them, so the oracle on the first 8 s pins the head of the one-hour run; the rest is checked by shape and
    by the same frames reappearing when the stream is fingerprinted from a later start (hop-aligned shift)."""
    sr, W, H = 16000, 512, 160
    x = synth.speech_band_noise(3600.0, sr=sr)
    p = gpu.default_params(window_size=W, hop_size=H, energy_frame=W, energy_hop=H, algo_sample_rate=sr,
                           call_sample_rate=sr, n_mel=40)
    fp = gpu.fingerprint(x, p)
    T = (x.size - W) // H + 1
    assert T == 359997 and fp.mfcc.shape == (T, 13) and fp.spectral_flux.size == T - 1  # T = 359,997
    head = oracle.fingerprint(x[: 8 * sr], p)
    t = head.short_time_energy.size - 4
    assert np.array_equal(fp.short_time_energy[:t], head.short_time_energy[:t])
    assert np.array_equal(fp.zero_crossing_rate[:t], head.zero_crossing_rate[:t])
    scale = np.max(np.abs(head.mfcc))
    assert np.allclose(fp.mfcc[:t], head.mfcc[:t], rtol=1e-4, atol=1e-4 * scale)
    assert np.allclose(fp.spectral_flux[: t - 1], head.spectral_flux[: t - 1], rtol=1e-4,
                       atol=1e-4 * np.max(head.spectral_flux))
    # shift by 200,000 frames (a multiple of the hop): the tail of the long run equals a fresh run on the shifted PCM,
    # up to the pre-emphasis of the very first sample (x[-1] = 0 in the fresh run) which only touches frame 0
    k = 200000
    tail = gpu.fingerprint(x[k * H: k * H + 20 * sr], p)
    m = tail.mfcc.shape[0] - 2
    assert np.array_equal(fp.short_time_energy[k + 1: k + m], tail.short_time_energy[1:m])
    for x, y in ((fp.mfcc[k + 1: k + m], tail.mfcc[1:m]), (fp.spectral_centroid[k + 1: k + m], tail.spectral_centroid[1:m])):
        assert np.all(np.abs(x - y) <= 1e-4 * np.maximum(np.abs(y), np.max(np.abs(y))))  # FP32: pairing differs per run


def _trim_by_lag(ea, eb, lag, length):
    """TruncateToAlignmentPCM's convention (extractors/alignment.go:239-243): lag > 0 skips the start of stream 2."""
    a, b = (ea, eb[lag:]) if lag >= 0 else (ea[-lag:], eb)
    return np.ascontiguousarray(a[:length]), np.ascontiguousarray(b[:length])


FP32_KEYS = ("mfcc", "spectral_centroid", "spectral_rolloff", "spectral_bandwidth", "spectral_flatness", "spectral_crest",
             "spectral_slope", "spectral_flux", "low_energy_ratio", "high_energy_ratio")


def _pair_bit_exact(gpu, oracle, q, r, sr, hop, max_lag_s, band, check_features):
    """The whole chained pair call against the oracle at full size: every short-time energy and zero-crossing rate, EVERY
    correlation value, the detected lag, the full DTW path and its costs bit for bit; the FP32 features within 1e-4."""
    p = gpu.default_params(algo_sample_rate=sr, call_sample_rate=sr)
    res = gpu.align_pairs([q], [r], p, max_lag_s, band)[0]
    oq, orf = oracle.fingerprint(q, p), oracle.fingerprint(r, p)
    for side, o in (("query", oq), ("reference", orf)):
        g = res[side]
        assert np.array_equal(g.short_time_energy, o.short_time_energy), side
        assert np.array_equal(g.zero_crossing_rate, o.zero_crossing_rate), side
        if check_features:
            for k in FP32_KEYS:  # flatness / slope included: frames with weak bins are redone in float64 (spectral_exact.cu)
                y, x = o.arrays[k], g.arrays[k]
                assert np.all(np.abs(x - y) <= 1e-4 * np.maximum(np.abs(y), np.max(np.abs(y)))), (side, k)
            assert np.array_equal(g.spectral_rolloff, o.spectral_rolloff), (side, "rolloff bin is a discrete choice")
            assert np.allclose(g.pitch_estimate, o.pitch_estimate, rtol=1e-4, atol=1e-6), side
            assert np.allclose(g.pitch_confidence, o.pitch_confidence, rtol=1e-4, atol=1e-6), side
    ea, eb = oq.short_time_energy, orf.short_time_energy
    max_lag = int(max_lag_s * sr) // hop
    corr, xs, al = oracle.align_xcorr(ea, eb, max_lag, hop, sr, want_corr=True)
    assert res["corr"].shape == corr.shape
    assert np.array_equal(res["corr"], corr), "every correlation value must be bit-exact"
    assert (res["xcorr"].peak_lag, res["xcorr"].peak_index) == (xs.peak_lag, xs.peak_index)
    assert res["xcorr"].peak_correlation == xs.peak_correlation and res["xcorr"].second_peak == xs.second_peak
    assert res["corr_alignment"].offset == al.offset and res["corr_alignment"].confidence == pytest.approx(al.confidence, rel=1e-9)
    length = min(ea.size, eb.size) - max_lag
    a, b = _trim_by_lag(ea, eb, xs.peak_lag, length)
    d = oracle.dtw(a, b, band=band)
    assert np.array_equal(res["path_query"], d["path_query"]) and np.array_equal(res["path_ref"], d["path_ref"])
    assert np.array_equal(res["path_cost"], d["path_cost"], equal_nan=True)
    assert res["total_cost"] == d["total_cost"] and res["distance"] == d["distance"]
    return res, xs


def test_c2_full_size_bit_exact_against_the_oracle(gpu, oracle, synth):
    """BASELINE config[1] compared with the oracle in full, not by properties."""
    q, r = synth.aligned_pair(300.0, offset_seconds=7.3, sr=44100, seed=200)
    res, xs = _pair_bit_exact(gpu, oracle, q, r, 44100, 256, 60.0, 50, check_features=True)
    assert res["corr"].size == 20671 and xs.peak_lag in (1257, 1258)
    assert res["path_query"].size >= 41341


def test_c5_one_ten_minute_pair_bit_exact_against_the_oracle(gpu, oracle, synth):
    """One pair of BASELINE config[4]'s shape (10 min, T = 103,356 frames, +-60 s lag, negative true offset)."""
    q, r = synth.aligned_pair(600.0, offset_seconds=-41.9, sr=44100, seed=205)
    res, xs = _pair_bit_exact(gpu, oracle, q, r, 44100, 256, 60.0, 50, check_features=False)
    assert res["query"].short_time_energy.size == 103356 and res["corr"].size == 20671
    assert xs.peak_lag < 0 and abs(xs.peak_lag * 256 / 44100 + 41.9) < 0.01


@pytest.mark.parametrize("algo_sr", [44100, 0])
def test_c1_thirty_seconds_both_sample_rate_modes(gpu, oracle, synth, algo_sr):
    """BASELINE config[0] at its full size (30 s, T = 5,164) in fixed-sr and in parity mode."""
    pcm = synth.sweep_noise(30.0, seed=1)
    p = gpu.default_params(algo_sample_rate=algo_sr, call_sample_rate=44100)
    g, o = gpu.fingerprint(pcm, p), oracle.fingerprint(pcm, p)
    assert g.mfcc.shape == (5164, 13) and g.sizes == o.sizes
    assert np.array_equal(g.short_time_energy, o.short_time_energy) and np.array_equal(g.zero_crossing_rate, o.zero_crossing_rate)
    for k in FP32_KEYS:
        y, x = o.arrays[k], g.arrays[k]
        scale = np.max(np.abs(y))
        assert np.all(np.abs(x - y) <= 1e-4 * np.maximum(np.abs(y), scale)), k
    assert np.array_equal(g.spectral_rolloff, o.spectral_rolloff)
    for k in ("pitch_estimate", "pitch_confidence", "voicing_strength", "harmonic_ratio", "inharmonicity_ratio", "tonal_centroid"):
        assert np.allclose(g.arrays[k], o.arrays[k], rtol=1e-4, atol=1e-6), k
    assert g.energy_variance == pytest.approx(o.energy_variance, rel=1e-10)
    assert g.loudness_range == pytest.approx(o.loudness_range, rel=1e-9, abs=1e-12)


def test_c2_dim13_mfcc_dtw_bit_exact_against_the_oracle(gpu, oracle, synth):
    """SURVEY §8(d) C2 'plus a dim-13 run on MFCC' / VERDICT r1 missing #7: banded DTW (r = 50) of the two 13-dimensional
    MFCC sequences of a 5-min pair (41,005 frames each after the lag trim), Euclidean local distance over the 13
    coefficients in the reference's order (algorithms/stats/distance.go:29-36): path, path costs and totals bit for
    bit.  The inputs are the oracle's float64 MFCCs (the DTW itself is what is compared)."""
    import time
    q, r = synth.aligned_pair(300.0, offset_seconds=7.3, sr=44100, seed=200)
    p = oracle.default_params(algo_sample_rate=44100, call_sample_rate=44100)
    mq, mr = oracle.fingerprint(q, p).mfcc, oracle.fingerprint(r, p).mfcc
    lag, length = 1258, 51676 - 10335
    a, b = np.ascontiguousarray(mq[:length]), np.ascontiguousarray(mr[lag:lag + length])
    t0 = time.perf_counter()
    g = gpu.dtw(a, b, band=50)
    gpu_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    o = oracle.dtw(a, b, band=50)
    cpu_s = time.perf_counter() - t0
    print(f"dim-13 banded DTW {length} x {length}: GPU {gpu_s * 1e3:.1f} ms (host call), oracle {cpu_s * 1e3:.1f} ms")
    assert g["path_query"].size >= length
    assert np.array_equal(g["path_query"], o["path_query"]) and np.array_equal(g["path_ref"], o["path_ref"])
    assert np.array_equal(g["path_cost"], o["path_cost"], equal_nan=True)
    assert g["total_cost"] == o["total_cost"] and g["distance"] == o["distance"]
