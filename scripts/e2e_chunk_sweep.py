"""e2e leg of the bench only, for a sweep of the host staging chunk size (SONAR_PAIR_CHUNK_MB)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("sonido-sonar_b200"); capi, synth = pkg.capi, pkg.synth
lib = capi.SonarLib()
P, seconds = 32, 300.0
n = int(seconds * 44100); stride = (n + 1) & ~1
prm = lib.default_params(algo_sample_rate=44100, call_sample_rate=44100)
host = torch.empty((2 * P, stride), dtype=torch.float64).pin_memory(); hv = host.numpy()
q0, r0 = bench.make_pair(synth, seconds, 0)
for i in range(P):
    hv[2 * i, :n], hv[2 * i + 1, :n] = q0, r0
bufs = lib.alloc_pair_outputs(P, n, prm, 60.0, features=True, corr=True)
qs = [hv[2 * i, :n] for i in range(P)]; rs = [hv[2 * i + 1, :n] for i in range(P)]
lib.align_pairs(qs, rs, prm, 60.0, 50, buffers=bufs)
ts = []
for _ in range(3):
    t0 = time.perf_counter(); lib.align_pairs(qs, rs, prm, 60.0, 50, buffers=bufs); ts.append(time.perf_counter() - t0)
print(os.environ.get("SONAR_PAIR_CHUNK_MB", "default"), "ms", [round(1e3 * t, 1) for t in ts], "GB/s", round(2 * P * n * 8 / min(ts) / 1e9, 1))
