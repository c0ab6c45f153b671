"""Times the fingerprint kernels of several library variants (build/variants/libsonar_*.so)."""
import glob, importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sonido-sonar_b200")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
SEC = 300.0
n = int(SEC * 44100); stride = (n + 1) & ~1
sig = sys.argv[2] if len(sys.argv) > 2 else "sweep"
x = torch.from_numpy(pkg.synth.sweep_noise(SEC, seed=9) if sig == "sweep" else pkg.synth.envelope_noise(n, seed=2)).cuda()
pcm = torch.zeros((S, stride), dtype=torch.float64, device="cuda"); pcm[:, :n] = x
libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", "libsonar_*.so"))) + [None]
ref = None
for path in libs:
    lib = pkg.capi.SonarLib(path)
    for sr in (44100,):
        p = lib.default_params(algo_sample_rate=sr)
        L = lib.fp_dev_layout(p, n)
        feat = torch.zeros(S * L.total, dtype=torch.float64, device="cuda")
        lib.profile_enable(True)
        for it in range(4):
            lib.fingerprint_batch_dev(pcm.data_ptr(), n, stride, S, p, feat.data_ptr())
            lib.synchronize()
        prof = lib.profile_read()
        ms, cnt = prof["stft_features_kernel"]
        f = feat.view(S, L.total)[0, :L.spectral_flux + 100].cpu().numpy().copy()
        if ref is None: ref = f
        print(f"{os.path.basename(path) if path else 'default':32s} stft {ms/cnt:7.3f} ms  frame_walk {prof['frame_walk_kernel'][0]/cnt:6.3f}  yin {prof['yin_frame_kernel'][0]/cnt:6.3f}  maxdiff_vs_first {np.max(np.abs(f-ref)):.3g}")
    lib.close()
