"""Development check of the FP32 YIN path alone (SONAR_YIN_NOEXACT=1 keeps the FP32 values of the borderline frames)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sonido-sonar_b200")
g = pkg.capi.SonarLib(); o = pkg.capi.SonarLib(os.path.join(ROOT, "oracle", "libsonar_oracle.so"))
sr = 44100
n = int(6.0 * sr); t = np.arange(n) / sr
f = 440.0 * (1.0 + 0.2 * np.sin(2 * np.pi * 0.7 * t)); ph = 2 * np.pi * np.cumsum(f) / sr
x = 0.4 * (np.sin(ph) + 0.05 * np.sin(2 * ph)) * (np.sin(2 * np.pi * 1.3 * t) > -0.3) + 1e-4 * np.random.default_rng(3).standard_normal(n)
for name, pcm in (("voiced", x), ("sweep", pkg.synth.sweep_noise(6.0, seed=1))):
    p = g.default_params(algo_sample_rate=sr)
    a, b = g.fingerprint(pcm, p), o.fingerprint(pcm, p)
    va, vb = a.pitch_confidence > 0, b.pitch_confidence > 0
    both = va & vb
    print(name, "frames", va.size, "voiced gpu/oracle", int(va.sum()), int(vb.sum()), "pattern equal", bool(np.array_equal(va, vb)))
    if both.any():
        print("  max rel pitch err", float(np.max(np.abs(a.pitch_estimate[both] - b.pitch_estimate[both]) / b.pitch_estimate[both])),
              "max conf err", float(np.max(np.abs(a.pitch_confidence[both] - b.pitch_confidence[both]))))
    bad = np.flatnonzero(va != vb)[:8]
    print("  first mismatches", bad.tolist(), a.pitch_confidence[bad].tolist(), b.pitch_confidence[bad].tolist())
    if name == "voiced":
        err = np.abs(a.pitch_confidence - b.pitch_confidence)
        idx = np.argsort(-err)[:12]
        for i in idx:
            print(f"   frame {i}: conf gpu {a.pitch_confidence[i]:.9f} oracle {b.pitch_confidence[i]:.9f}  pitch gpu {a.pitch_estimate[i]:.6f} oracle {b.pitch_estimate[i]:.6f}")
        print("   median conf err", float(np.median(err[both])), " 90%", float(np.quantile(err[both], 0.9)))
