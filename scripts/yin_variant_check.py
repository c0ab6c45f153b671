"""Pitch arrays of every library variant (build/variants/libsonar_*.so) against the default library: a change of the
pitch kernel's control flow must leave them bit-identical."""
import glob, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sonido-sonar_b200")
sr = 44100
n = int(6.0 * sr); t = np.arange(n) / sr
f = 440.0 * (1.0 + 0.2 * np.sin(2 * np.pi * 0.7 * t)); ph = 2 * np.pi * np.cumsum(f) / sr
voiced = 0.4 * (np.sin(ph) + 0.05 * np.sin(2 * ph)) * (np.sin(2 * np.pi * 1.3 * t) > -0.3) + 1e-4 * np.random.default_rng(3).standard_normal(n)
sigs = {"voiced": voiced, "sweep": pkg.synth.sweep_noise(6.0, seed=1), "odd": voiced[: n - 777]}
base = pkg.capi.SonarLib()
p = base.default_params(algo_sample_rate=sr)
want = {k: base.fingerprint(v, p) for k, v in sigs.items()}
for path in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "libsonar_*.so"))):
    lib = pkg.capi.SonarLib(path)
    for k, v in sigs.items():
        a = lib.fingerprint(v, p)
        same = np.array_equal(a.pitch_estimate, want[k].pitch_estimate) and np.array_equal(a.pitch_confidence, want[k].pitch_confidence)
        print(os.path.basename(path), k, "frames", a.pitch_estimate.size, "voiced", int((a.pitch_confidence > 0).sum()),
              "bit-identical" if same else f"DIFFERENT maxdiff {np.max(np.abs(a.pitch_estimate - want[k].pitch_estimate)):.3g}")
    lib.close()
