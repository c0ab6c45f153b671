"""Diagnostic driver: a tone over a 1e-4 noise floor -> every frame takes the float64 re-evaluation (spectral_exact.cu)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sonido-sonar_b200")
gpu = pkg.capi.SonarLib()
sr = 44100
t = np.arange(int(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0) * sr) / sr
x = 0.5 * np.sin(2 * np.pi * 440.0 * t) + 1e-4 * np.random.default_rng(5).standard_normal(t.size)
p = gpu.default_params(algo_sample_rate=sr, call_sample_rate=sr)
gpu.fingerprint(x, p)
gpu.profile_enable(True)
for _ in range(3):
    gpu.fingerprint(x, p)
print(gpu.exact_counts())
for k, (ms, cnt) in sorted(gpu.profile_read().items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:28s} {ms/cnt:9.3f} ms/launch x{cnt}")
