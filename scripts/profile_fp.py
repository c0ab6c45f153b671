"""Profiling driver: device-resident fingerprint of S streams x SEC seconds (the bench's fingerprint leg only)."""
import ctypes as C, importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sonido-sonar_b200")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
SEC = float(sys.argv[2]) if len(sys.argv) > 2 else 300.0
SR = int(sys.argv[3]) if len(sys.argv) > 3 else 44100
ITERS = int(sys.argv[4]) if len(sys.argv) > 4 else 4
lib = pkg.capi.SonarLib()
n = int(SEC * 44100); stride = (n + 1) & ~1
x = torch.from_numpy(pkg.synth.sweep_noise(SEC, seed=9)).cuda()
pcm = torch.zeros((S, stride), dtype=torch.float64, device="cuda"); pcm[:, :n] = x
p = lib.default_params(algo_sample_rate=SR)
L = lib.fp_dev_layout(p, n)
feat = torch.empty(S * L.total, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
lib.profile_enable(True)
for it in range(ITERS):
    t0 = time.perf_counter()
    lib.fingerprint_batch_dev(pcm.data_ptr(), n, stride, S, p, feat.data_ptr())
    lib.synchronize()
    dt = time.perf_counter() - t0
    print(f"iter {it}: {dt*1e3:.2f} ms -> {S*SEC/dt:.4g} audio-s/s")
for k, (ms, cnt) in sorted(lib.profile_read().items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:28s} {ms/cnt:9.3f} ms/launch x{cnt}")
