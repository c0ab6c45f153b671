"""GPU path against the committed fixtures (tests/golden/, produced by make_golden.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXACT = ("short_time_energy", "zero_crossing_rate")


@pytest.mark.parametrize("name", ["c1_fixed_sr", "c1_parity", "c3_speech_40mel"])
def test_fingerprint_fixture(gpu, name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    kw = {k[3:]: int(g[k]) for k in g.files if k.startswith("kw_")}
    fp = gpu.fingerprint(g["pcm"], gpu.default_params(**kw))
    for k in g.files:
        if not k.startswith("out_"):
            continue
        x, y = fp.arrays[k[4:]], g[k]
        if k[4:] in EXACT:
            assert np.array_equal(x, y), k
        else:
            scale = np.max(np.abs(y)) if y.size else 0.0
            assert np.all(np.abs(x - y) <= 1e-4 * np.maximum(np.abs(y), scale)), k  # tolerance: test_gpu_fingerprint.py


def test_alignment_fixture(gpu):
    g = np.load(os.path.join(GOLDEN, "c2_alignment.npz"))
    c, s = gpu.xcorr(g["ea"], g["eb"], int(g["max_lag"]))
    assert np.array_equal(c, g["corr"]) and s.peak_lag == int(g["peak_lag"])
    d = gpu.dtw(g["dq"], g["dr"], band=int(g["band"]))
    assert np.array_equal(d["path_query"], g["path_query"]) and np.array_equal(d["path_ref"], g["path_ref"])
    assert np.array_equal(d["path_cost"], g["path_cost"], equal_nan=True)
