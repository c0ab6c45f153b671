"""Writes the seeded inputs the Go parity dumper (baseline/go/parity_dump_test.go) runs the REAL reference on.

    python tests/golden/make_go_inputs.py        ->  tests/golden/from_go/inputs/{manifest.json, *.f64}

Raw little-endian float64 (what Go's math.Float64frombits reads) plus a manifest; the cases mirror the oracle-made
fixtures of make_golden.py and BASELINE.json's configurations at sizes the Go path finishes in seconds.  The dumper's
outputs land in tests/golden/from_go/<case>/*.npy, which tests/test_go_golden.py (oracle vs Go) and
tests/test_gpu_golden.py (CUDA vs Go) prefer over the oracle-made fixtures.
"""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
synth = importlib.import_module("sonido-sonar_b200").synth
OUT = os.path.join(HERE, "from_go", "inputs")
os.makedirs(OUT, exist_ok=True)
cases = []


def put(name, x):
    np.ascontiguousarray(x, dtype="<f8").tofile(os.path.join(OUT, name))
    return name


def tone_mix(seconds, sr, seed):
    n = int(seconds * sr)
    t = np.arange(n) / sr
    rng = np.random.default_rng(seed)
    return 0.4 * np.sin(2 * np.pi * (440 + 60 * np.sin(2 * np.pi * 0.8 * t)) * t) + 1e-4 * rng.standard_normal(n)


# C1: GenerateFingerprint, music, 1024/256 -- stock (parity mode, algorithms at sample rate 0) and fixed-rate mode
c1 = np.concatenate([synth.sweep_noise(3.0, seed=1), tone_mix(1.0, 44100, 2)])
put("c1.f64", c1)
cases.append(dict(name="c1_parity", kind="fingerprint", pcm="c1.f64", sample_rate=44100, window_size=1024, hop_size=256,
                  content_type="music"))
cases.append(dict(name="c1_fixed_sr", kind="extract", pcm="c1.f64", sample_rate=44100, algo_sample_rate=44100,
                  window_size=1024, hop_size=256, content_type="music"))
# C3: news, 16 kHz, 512/160 (stock 26 mel filters: NumMelFilters is not reachable through the public constructors)
c3 = synth.speech_band_noise(4.0)
put("c3.f64", c3)
cases.append(dict(name="c3_news_parity", kind="fingerprint", pcm="c3.f64", sample_rate=16000, window_size=512, hop_size=160,
                  content_type="news"))
cases.append(dict(name="c3_news_fixed_sr", kind="extract", pcm="c3.f64", sample_rate=16000, algo_sample_rate=16000,
                  window_size=512, hop_size=160, content_type="news"))
# C2: ExtractAlignmentFeatures on a 10 s pair, true offset 1.7 s, +-6 s lag
q, r = synth.aligned_pair(10.0, offset_seconds=1.7, seed=2)
put("c2_q.f64", q)
put("c2_r.f64", r)
cases.append(dict(name="c2_align", kind="align", pcm="c2_q.f64", pcm2="c2_r.f64", sample_rate=44100, algo_sample_rate=44100,
                  window_size=1024, hop_size=256, content_type="music", max_lag_seconds=6.0))
# cross-correlation and DTW on plain sequences (no dependence on the FFT / gonum modules: these must be BIT exact)
rng = np.random.default_rng(11)
base = np.convolve(rng.standard_normal(6000), np.ones(16) / 16, mode="same") + 1.0
put("xc_a.f64", base[500:4500])
put("xc_b.f64", base[500 - 137:4500 - 137] + 0.01 * rng.standard_normal(4000))
cases.append(dict(name="xcorr_shift137", kind="xcorr", a="xc_a.f64", b="xc_b.f64", max_lag=600))
put("dtw_q.f64", base[1000:1700])
put("dtw_r.f64", base[1003:1653] * 1.02)
for step in ("symmetric2", "symmetric1", "asymmetric"):
    cases.append(dict(name=f"dtw_band50_{step}", kind="dtw", a="dtw_q.f64", b="dtw_r.f64", dim=1, band=50, step_pattern=step))
cases.append(dict(name="dtw_unconstrained", kind="dtw", a="dtw_q.f64", b="dtw_r.f64", dim=1, band=-1,
                  step_pattern="symmetric2"))
m13 = rng.standard_normal((300, 13))
put("dtw13_q.f64", m13[:280])
put("dtw13_r.f64", m13[5:300] + 0.05 * rng.standard_normal((295, 13)))
cases.append(dict(name="dtw_dim13_band40", kind="dtw", a="dtw13_q.f64", b="dtw13_r.f64", dim=13, band=40,
                  step_pattern="symmetric2"))
# Compare of two stock fingerprints
put("cmp_2.f64", synth.sweep_noise(3.0, seed=5, f1=6000.0))
cases.append(dict(name="compare_stock", kind="compare", pcm="c1.f64", pcm2="cmp_2.f64", sample_rate=44100, window_size=1024,
                  hop_size=256, content_type="music"))

with open(os.path.join(OUT, "manifest.json"), "w") as f:
    json.dump(cases, f, indent=1)
print(f"{len(cases)} cases -> {OUT}")
