"""Generates the committed fixtures in this directory from the CPU oracle.

The reference cannot be run here (pure Go, no toolchain) and has no golden vectors of its own, so
these are *oracle* outputs on seeded inputs: regression pins for the oracle and fixed targets for
the GPU parity tests.  The oracle itself is pinned by the independent numpy/pure-Python
restatements in tests/test_oracle_kat.py.  Run from the repo root:  python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sonido-sonar_b200")
ora = pkg.capi.SonarLib(os.path.join(ROOT, "oracle", "libsonar_oracle.so"))
synth = pkg.synth

KEEP = ("mfcc", "spectral_centroid", "spectral_rolloff", "spectral_bandwidth", "spectral_flatness", "spectral_crest",
        "spectral_slope", "spectral_flux", "zero_crossing_rate", "short_time_energy", "energy_entropy",
        "low_energy_ratio", "high_energy_ratio", "pitch_estimate", "pitch_confidence")


def fp_fixture(name, pcm, **kw):
    fp = ora.fingerprint(pcm, ora.default_params(**kw))
    d = {"pcm": pcm}
    d.update({f"kw_{k}": np.int64(v) for k, v in kw.items()})
    d.update({f"out_{k}": fp.arrays[k] for k in KEEP})
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **d)


def tone_mix(seconds, sr, seed):
    n = int(seconds * sr)
    t = np.arange(n) / sr
    rng = np.random.default_rng(seed)
    return 0.4 * np.sin(2 * np.pi * (440 + 60 * np.sin(2 * np.pi * 0.8 * t)) * t) + 1e-4 * rng.standard_normal(n)


pcm = np.concatenate([synth.sweep_noise(0.6, seed=1), tone_mix(0.6, 44100, 2)])
fp_fixture("c1_fixed_sr", pcm, algo_sample_rate=44100)
fp_fixture("c1_parity", pcm, algo_sample_rate=0)
fp_fixture("c3_speech_40mel", synth.speech_band_noise(1.5), window_size=512, hop_size=160, energy_frame=512,
           energy_hop=160, algo_sample_rate=16000, call_sample_rate=16000, n_mel=40)

q, r = synth.aligned_pair(12.0, offset_seconds=1.7, seed=2)
p = ora.default_params(algo_sample_rate=44100)
ea, eb = ora.fingerprint(q, p).short_time_energy, ora.fingerprint(r, p).short_time_energy
max_lag = 600
corr, s = ora.xcorr(ea, eb, max_lag)
lag = s.peak_lag
n = min(ea.size, eb.size) - max_lag
dq, dr = (ea[:n], eb[lag:lag + n]) if lag >= 0 else (ea[-lag:-lag + n], eb[:n])
d = ora.dtw(dq, dr, band=50)
np.savez_compressed(os.path.join(HERE, "c2_alignment.npz"), ea=ea, eb=eb, max_lag=np.int64(max_lag), corr=corr,
                    peak_lag=np.int64(lag), dq=dq, dr=dr, band=np.int64(50), path_query=d["path_query"],
                    path_ref=d["path_ref"], path_cost=d["path_cost"])
print("fixtures written; peak lag", lag, "path len", len(d["path_query"]))
