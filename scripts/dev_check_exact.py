"""Diagnostic: FP32 fused features vs the oracle per feature, the listed (float64 re-evaluated) frame counts and the
rolloff mismatches, on the BASELINE inputs.  usage: python scripts/dev_check_exact.py [seconds]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sonido-sonar_b200")
capi, synth = pkg.capi, pkg.synth
SEC = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
gpu = capi.SonarLib()
ora = capi.SonarLib(os.path.join(ROOT, "oracle", "libsonar_oracle.so"))
KEYS = ("mfcc", "spectral_centroid", "spectral_rolloff", "spectral_bandwidth", "spectral_flatness", "spectral_crest",
        "spectral_slope", "spectral_flux", "low_energy_ratio", "high_energy_ratio")
def tone_floor(sec, sr=44100, seed=5):
    t = np.arange(int(sec * sr)) / sr
    return 0.5 * np.sin(2 * np.pi * 440.0 * t) + 1e-4 * np.random.default_rng(seed).standard_normal(t.size)
cases = {
    "c2_envelope_noise": (synth.aligned_pair(SEC, 7.3, seed=200)[0], dict(algo_sample_rate=44100, call_sample_rate=44100)),
    "c1_sweep_noise": (synth.sweep_noise(SEC, seed=1), dict(algo_sample_rate=44100, call_sample_rate=44100)),
    "c3_speech_512": (synth.speech_band_noise(SEC, sr=16000), dict(window_size=512, hop_size=160, energy_frame=512, energy_hop=160,
                                                               algo_sample_rate=16000, call_sample_rate=16000, n_mel=40)),
    "tone_1e-4_floor": (tone_floor(min(SEC, 20.0)), dict(algo_sample_rate=44100, call_sample_rate=44100)),
}
for name, (x, kw) in cases.items():
    p = gpu.default_params(**kw)
    g = gpu.fingerprint(x, p)
    cx, cy = gpu.exact_counts()
    o = ora.fingerprint(x, p)
    T = g.mfcc.shape[0]
    print(f"== {name}: T={T} listed spectral={cx} ({100.0*cx/T:.2f} %) pitch={cy}")
    for k in KEYS:
        a, b = g.arrays[k], o.arrays[k]
        scale = np.max(np.abs(b)) if b.size else 1.0
        err = np.abs(a - b) / np.maximum(np.abs(b), scale if scale > 0 else 1.0)
        print(f"   {k:22s} worst rel {err.max():.3e}  mismatched {int((a != b).sum())}/{b.size}")
