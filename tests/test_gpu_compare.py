"""GPU parity: column statistics + cosine and FingerprintComparator.Compare (tolerance 1e-9)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_colstats(gpu, oracle):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((5164, 13)) * np.arange(1, 14) + 3.0
    np.testing.assert_allclose(gpu.colstats(x), oracle.colstats(x), rtol=1e-12)
    y = rng.standard_normal((4000, 13))
    assert gpu.colstats_cosine(x, y) == pytest.approx(oracle.colstats_cosine(x, y), rel=1e-10, abs=1e-14)
    s = rng.standard_normal(777)
    np.testing.assert_allclose(gpu.colstats(s), oracle.colstats(s), rtol=1e-12)


def test_compare_fingerprints(gpu, oracle, synth):
    p = gpu.default_params(algo_sample_rate=44100)
    x1 = synth.sweep_noise(4.0, seed=1)
    x2 = synth.sweep_noise(4.0, seed=2, f1=6000.0)
    fa1, fa2 = gpu.fingerprint(x1, p), gpu.fingerprint(x2, p)
    w = [0.5, 0.2, 0.0, 0.1, 0.0, 0.2, 0.0]
    f1, k1 = gpu.cmp_features(fa1)
    f2, k2 = gpu.cmp_features(fa2)
    ra = gpu.compare(f1, f2, w).as_dict()
    rb = oracle.compare(f1, f2, w).as_dict()
    for k in ra:
        assert ra[k] == pytest.approx(rb[k], rel=1e-9, abs=1e-12, nan_ok=True), k
    # identical fingerprints: every present feature scores 1 except all-zero series (cosine of zero
    # vectors is 0, comparison.go:867-869) -- whatever the oracle says, the GPU must say too
    same, same_ref = gpu.compare(f1, f1, w), oracle.compare(f1, f1, w)
    assert same.overall_similarity == pytest.approx(same_ref.overall_similarity, rel=1e-9)
    assert same.dist_mfcc == pytest.approx(0.0, abs=1e-9)
    # content filter short-circuit (comparison.go:160-166)
    f2.content_type = 3
    r = gpu.compare(f1, f2, w, content_filter=True)
    assert r.overall_similarity == 0.0 and r.confidence == 0.25 and r.content_type_match == 0


def test_batch_compare_equals_single_compares(gpu, oracle, synth):
    """sonar_compare_batch_f64 (BatchCompare / FindBestMatches' loop, comparison.go:1107-1151,197-263)."""
    p = gpu.default_params(algo_sample_rate=44100)
    w = [0.5, 0.2, 0.0, 0.1, 0.0, 0.2, 0.0]
    fps = [gpu.fingerprint(synth.sweep_noise(3.0, seed=50 + i, f1=3000.0 + 1500.0 * i), p) for i in range(4)]
    feats = [gpu.cmp_features(f) for f in fps]  # (struct, keep-alive)
    q = feats[0][0]
    cands = [feats[1][0], None, feats[2][0], feats[3][0], feats[0][0]]
    got = gpu.compare_batch(q, cands, w)
    ref = oracle.compare_batch(q, cands, w)
    assert got[1].n_features == -1 and ref[1].n_features == -1  # nil candidates are skipped
    for i, c in enumerate(cands):
        if c is None:
            continue
        one = gpu.compare(q, c, w).as_dict()
        for k, v in got[i].as_dict().items():
            assert v == pytest.approx(one[k], rel=0, abs=0, nan_ok=True), k       # same code path: identical
            assert v == pytest.approx(ref[i].as_dict()[k], rel=1e-9, abs=1e-12, nan_ok=True), k
    order = sorted((i for i, c in enumerate(cands) if c is not None), key=lambda i: -got[i].overall_similarity)
    assert order[0] == 4  # the query itself is its own best match
