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
    assert np.allclose(fq.pitch_estimate[:tp], oq.pitch_estimate[:tp], rtol=1e-7, atol=1e-9)
    # frames are also independent of how the kernel cuts the stream into runs of 31 frames / rings of 1024 samples:
    # a fresh fingerprint of the PCM from frame 30,011 on reproduces the long run's frames bit for bit (except the
    # first one, whose pre-emphasis sees x[-1] = 0 in the fresh run; flux additionally needs its predecessor)
    k = 30011
    tail = gpu.fingerprint(q[k * hop: k * hop + 12 * sr], p)
    m = tail.mfcc.shape[0] - 2
    for name in ("mfcc", "spectral_centroid", "spectral_rolloff", "spectral_bandwidth", "spectral_flatness",
                 "short_time_energy", "zero_crossing_rate"):
        assert np.array_equal(fq.arrays[name][k + 1: k + m], tail.arrays[name][1:m]), name
    assert np.array_equal(fq.spectral_flux[k + 1: k + m - 1], tail.spectral_flux[1: m - 1])


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
    assert np.array_equal(fp.mfcc[k + 1: k + m], tail.mfcc[1:m])
    assert np.array_equal(fp.short_time_energy[k + 1: k + m], tail.short_time_energy[1:m])
    assert np.array_equal(fp.spectral_centroid[k + 1: k + m], tail.spectral_centroid[1:m])
