"""The oracle against the REAL Go reference -- the only route from "parity unpinned" to pinned.

Skipped until someone with a Go toolchain has run baseline/go/parity_dump_test.go (baseline/go/README.md); the dumper
writes tests/golden/from_go/<case>/*.npy.  Bit-exact where the reference touches neither go-dsp nor gonum (energies,
zero crossings, per-lag NCC, DTW); 1e-9 relative where its FFT / statistics modules are involved.
"""
import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.skipif(not golden_io.have_go_outputs(), reason="no Go outputs under tests/golden/from_go "
                                "(baseline/go/README.md)")
EXACT = ("short_time_energy", "zero_crossing_rate")
CASES = {c["name"]: c for c in golden_io.manifest()}


def _params(lib, c, algo):
    return lib.default_params(window_size=c["window_size"], hop_size=c["hop_size"], energy_frame=c["window_size"],
                              energy_hop=c["hop_size"], algo_sample_rate=algo, call_sample_rate=c["sample_rate"])


def check_features(fp, g, exact=EXACT, rtol=1e-9):
    for k, y in g.items():
        if k in ("energy_scalars", "temporal_scalars", "attack_time", "envelope_shape", "rms_energy", "result",
                 "feature_distances"):
            continue
        x = fp.arrays[k]
        assert x.shape == y.shape, k
        if k in exact:
            assert np.array_equal(x, y), k
        else:
            scale = np.max(np.abs(y)) if y.size else 0.0
            assert np.all(np.abs(x - y) <= rtol * np.maximum(np.abs(y), scale)), k
    if "energy_scalars" in g:
        assert fp.energy_variance == pytest.approx(g["energy_scalars"][0], rel=1e-9)
        assert fp.loudness_range == pytest.approx(g["energy_scalars"][1], rel=1e-9, abs=1e-12)


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c["kind"] in ("fingerprint", "extract")])
def test_oracle_fingerprint_matches_go(oracle, name):
    c = CASES[name]
    algo = 0 if c["kind"] == "fingerprint" else c["algo_sample_rate"]  # stock GenerateFingerprint: SURVEY F2/F3
    fp = oracle.fingerprint(golden_io.read_input(c["pcm"]), _params(oracle, c, algo))
    check_features(fp, golden_io.load_case(name))


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c["kind"] == "xcorr"])
def test_oracle_xcorr_matches_go_bit_for_bit(oracle, name):
    c, g = CASES[name], golden_io.load_case(name)
    corr, s = oracle.xcorr(golden_io.read_input(c["a"]), golden_io.read_input(c["b"]), c["max_lag"])
    assert np.array_equal(corr, g["correlations"])
    assert (s.peak_correlation, s.peak_lag, s.peak_index) == (g["peak"][0], int(g["peak"][1]), int(g["peak"][2]))
    assert s.second_peak == g["peak"][6] and s.sharpness == g["peak"][5]
    assert s.snr == pytest.approx(g["peak"][4], rel=1e-12) and s.peak_to_sidelobe == pytest.approx(g["peak"][7], rel=1e-12)


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c["kind"] == "dtw"])
def test_oracle_dtw_matches_go_bit_for_bit(oracle, capi, name):
    c, g = CASES[name], golden_io.load_case(name)
    q = golden_io.read_input(c["a"]).reshape(-1, c["dim"])
    r = golden_io.read_input(c["b"]).reshape(-1, c["dim"])
    step = {"symmetric2": capi.STEP_SYMMETRIC2, "symmetric1": capi.STEP_SYMMETRIC1, "asymmetric": capi.STEP_ASYMMETRIC}
    d = oracle.dtw(q, r, band=c["band"], step=step[c["step_pattern"]])
    assert np.array_equal(d["path_query"], g["path_query"]) and np.array_equal(d["path_ref"], g["path_ref"])
    assert np.array_equal(d["path_cost"], g["path_cost"], equal_nan=True)
    assert d["distance"] == g["distance"][0]


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c["kind"] == "align"])
def test_oracle_alignment_matches_go(oracle, name):
    c, g = CASES[name], golden_io.load_case(name)
    p = _params(oracle, c, c["algo_sample_rate"])
    ea = oracle.fingerprint(golden_io.read_input(c["pcm"]), p).short_time_energy
    eb = oracle.fingerprint(golden_io.read_input(c["pcm2"]), p).short_time_energy
    assert np.array_equal(ea, g["query_short_time_energy"]) and np.array_equal(eb, g["reference_short_time_energy"])
    max_lag = int(c["max_lag_seconds"] * c["sample_rate"]) // c["hop_size"]
    corr, xs, al = oracle.align_xcorr(ea, eb, max_lag, c["hop_size"], c["sample_rate"], want_corr=True)
    assert np.array_equal(corr, g["corr_correlations"])
    assert xs.peak_lag == int(g["corr_peak"][1]) and al.offset == int(g["corr_scalars"][0])
    for got, want in zip((al.offset_seconds, al.confidence, al.similarity, al.alignment_quality), g["corr_scalars"][1:5]):
        assert got == pytest.approx(want, rel=1e-12)
