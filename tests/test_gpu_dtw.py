"""GPU parity: DTW path bit-exact with the oracle (dtw.go:55-217), all step patterns, band / no band."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def check_dtw(a, b, matrix=False):
    assert np.array_equal(a["path_query"], b["path_query"])
    assert np.array_equal(a["path_ref"], b["path_ref"])
    assert np.array_equal(a["path_cost"], b["path_cost"], equal_nan=True)
    assert a["total_cost"] == b["total_cost"] or (np.isnan(a["total_cost"]) and np.isnan(b["total_cost"]))
    assert a["distance"] == b["distance"] or (np.isnan(a["distance"]) and np.isnan(b["distance"]))
    if matrix:
        assert np.array_equal(a["cost_matrix"], b["cost_matrix"])


@pytest.mark.parametrize("n,m,dim,band,step", [
    (60, 60, 1, -1, 0), (60, 75, 1, -1, 0), (75, 60, 3, -1, 1), (50, 50, 13, -1, 2),
    (300, 300, 1, 50, 0), (300, 310, 1, 20, 0), (300, 340, 2, 20, 0),  # |n-m| > band: +Inf end cell, garbage walk
    (1, 1, 1, -1, 0), (1, 40, 1, -1, 0), (40, 1, 1, 5, 0), (2000, 2000, 1, 50, 0), (700, 650, 13, 64, 2),
])
def test_dtw_random(gpu, oracle, n, m, dim, band, step):
    rng = np.random.default_rng(n * 7 + m * 3 + dim)
    q = np.cumsum(rng.standard_normal((n, dim)), axis=0)
    r = np.cumsum(rng.standard_normal((m, dim)), axis=0)
    small = n * m <= 100000
    a = gpu.dtw(q, r, band=band, step=step, want_matrix=small)
    b = oracle.dtw(q, r, band=band, step=step, want_matrix=small)
    check_dtw(a, b, matrix=small)


def test_dtw_identical_is_diagonal(gpu, oracle):
    x = np.sin(np.arange(500) * 0.05)
    a = gpu.dtw(x, x, band=50)
    b = oracle.dtw(x, x, band=50)
    check_dtw(a, b)
    assert np.array_equal(a["path_query"], np.arange(500)) and np.array_equal(a["path_ref"], np.arange(500))
    assert a["total_cost"] == 0.0


def test_dtw_ties_follow_reference_order(gpu, oracle):
    # integer-valued sequences create exact ties; vertical, horizontal, diagonal scan order decides
    rng = np.random.default_rng(3)
    q = rng.integers(0, 3, 200).astype(float)
    r = rng.integers(0, 3, 220).astype(float)
    for band in (-1, 40):
        check_dtw(gpu.dtw(q, r, band=band), oracle.dtw(q, r, band=band))


def test_dtw_unconstrained_large_line_in_global(gpu, oracle):
    # n + m + 1 offsets exceed the shared-memory line -> global line path
    rng = np.random.default_rng(9)
    q = np.cumsum(rng.standard_normal(13500))
    r = np.cumsum(rng.standard_normal(13400))
    check_dtw(gpu.dtw(q, r, band=-1), oracle.dtw(q, r, band=-1))


def test_dtw_batch_and_scalars(gpu, oracle):
    rng = np.random.default_rng(17)
    qs = [np.cumsum(rng.standard_normal(400)) for _ in range(5)]
    rs = [x + 0.05 * rng.standard_normal(400) for x in qs]
    A = gpu.dtw_batch(qs, rs, band=30)
    B = oracle.dtw_batch(qs, rs, band=30)
    for a, b in zip(A, B):
        check_dtw(a, b)
        sa = gpu.align_dtw_scalars(a, 400, 400, 44100).as_dict()
        sb = oracle.align_dtw_scalars(b, 400, 400, 44100).as_dict()
        for k in sa:
            assert sa[k] == pytest.approx(sb[k], rel=1e-12, abs=1e-300, nan_ok=True), k


def test_dtw_errors(gpu, capi):
    with pytest.raises(capi.SonarError) as e:
        gpu.dtw(np.zeros((0, 1)), np.zeros((4, 1)))
    assert e.value.code == capi.ERR_EMPTY and "empty sequences provided" in e.value.msg


@pytest.mark.parametrize("n,m,dim,band,step", [(300, 300, 1, 50, 0), (120, 140, 1, -1, 0), (200, 200, 3, 30, 2),
                                               (90, 90, 1, 10, 1)])
def test_dtw_nan_inputs_propagate_like_math_min(gpu, oracle, n, m, dim, band, step):
    """Go's math.Min returns NaN when either argument is NaN (dtw.go:137-164), fmin would drop it.  With a
    NaN (and an Inf) inside the sequences the cost matrix turns NaN behind it and the backtrack's strict '<' scan
    (dtw.go:191-217) walks differently: path, costs and matrix must still equal the oracle's."""
    rng = np.random.default_rng(n + m + dim + band)
    q = np.cumsum(rng.standard_normal((n, dim)), axis=0)
    r = np.cumsum(rng.standard_normal((m, dim)), axis=0)
    q[n // 3, 0] = np.nan
    r[2 * m // 3, dim - 1] = np.inf
    small = n * m <= 100000
    a = gpu.dtw(q, r, band=band, step=step, want_matrix=small)
    b = oracle.dtw(q, r, band=band, step=step, want_matrix=small)
    assert np.array_equal(a["path_query"], b["path_query"]) and np.array_equal(a["path_ref"], b["path_ref"])
    assert np.array_equal(a["path_cost"], b["path_cost"], equal_nan=True)
    if small:
        assert np.array_equal(a["cost_matrix"], b["cost_matrix"], equal_nan=True)
