"""GPU path against the committed fixtures (tests/golden/, produced by make_golden.py) and -- preferred, when someone has
run the Go dumper of baseline/go/ -- against the outputs of the real reference under tests/golden/from_go/."""
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
            # tolerance as in test_gpu_fingerprint.py: 1e-4, and 2e-3 for the two features that are means of ln|X| over
            # ALL bins (the fixture's tone-over-1e-4-noise half spans > 60 dB: its weakest bins sit at the FP32 FFT's floor)
            tol = 2e-3 if k[4:] in ("spectral_flatness", "spectral_slope") else 1e-4
            assert np.all(np.abs(x - y) <= tol * np.maximum(np.abs(y), scale)), k


def test_alignment_fixture(gpu):
    g = np.load(os.path.join(GOLDEN, "c2_alignment.npz"))
    c, s = gpu.xcorr(g["ea"], g["eb"], int(g["max_lag"]))
    assert np.array_equal(c, g["corr"]) and s.peak_lag == int(g["peak_lag"])
    d = gpu.dtw(g["dq"], g["dr"], band=int(g["band"]))
    assert np.array_equal(d["path_query"], g["path_query"]) and np.array_equal(d["path_ref"], g["path_ref"])
    assert np.array_equal(d["path_cost"], g["path_cost"], equal_nan=True)


# ---- the real reference's outputs, when present (baseline/go/README.md) ------------------------------------------
import golden_io  # noqa: E402

GO_CASES = {c["name"]: c for c in golden_io.manifest()} if golden_io.have_go_outputs() else {}
needs_go = pytest.mark.skipif(not GO_CASES, reason="no Go outputs under tests/golden/from_go (baseline/go/README.md)")


@needs_go
@pytest.mark.parametrize("name", [n for n, c in GO_CASES.items() if c["kind"] in ("fingerprint", "extract")])
def test_fingerprint_matches_go(gpu, name):
    c, g = GO_CASES[name], golden_io.load_case(name)
    algo = 0 if c["kind"] == "fingerprint" else c["algo_sample_rate"]
    p = gpu.default_params(window_size=c["window_size"], hop_size=c["hop_size"], energy_frame=c["window_size"],
                           energy_hop=c["hop_size"], algo_sample_rate=algo, call_sample_rate=c["sample_rate"])
    fp = gpu.fingerprint(golden_io.read_input(c["pcm"]), p)
    for k, y in g.items():
        if k not in fp.arrays:
            continue
        x = fp.arrays[k]
        if k in EXACT:
            assert np.array_equal(x, y), k
        else:
            scale = np.max(np.abs(y)) if y.size else 0.0
            assert np.all(np.abs(x - y) <= 1e-4 * np.maximum(np.abs(y), scale)), k


@needs_go
@pytest.mark.parametrize("name", [n for n, c in GO_CASES.items() if c["kind"] in ("xcorr", "dtw")])
def test_alignment_kernels_match_go_bit_for_bit(gpu, capi, name):
    c, g = GO_CASES[name], golden_io.load_case(name)
    if c["kind"] == "xcorr":
        corr, s = gpu.xcorr(golden_io.read_input(c["a"]), golden_io.read_input(c["b"]), c["max_lag"])
        assert np.array_equal(corr, g["correlations"]) and s.peak_lag == int(g["peak"][1])
    else:
        step = {"symmetric2": capi.STEP_SYMMETRIC2, "symmetric1": capi.STEP_SYMMETRIC1, "asymmetric": capi.STEP_ASYMMETRIC}
        d = gpu.dtw(golden_io.read_input(c["a"]).reshape(-1, c["dim"]), golden_io.read_input(c["b"]).reshape(-1, c["dim"]),
                    band=c["band"], step=step[c["step_pattern"]])
        assert np.array_equal(d["path_query"], g["path_query"]) and np.array_equal(d["path_ref"], g["path_ref"])
        assert np.array_equal(d["path_cost"], g["path_cost"], equal_nan=True)
