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
