"""int16-ingest e2e leg only, for a sweep of the host staging chunk size (SONAR_PAIR_CHUNK_MB)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("sonido-sonar_b200"); capi, synth = pkg.capi, pkg.synth
lib = capi.SonarLib()
P, seconds = 32, 300.0
n = int(seconds * 44100)
prm = lib.default_params(algo_sample_rate=44100, call_sample_rate=44100)
host = torch.empty((2 * P, n), dtype=torch.int16).pin_memory(); hv = host.numpy()
q0, r0 = bench.make_pair(synth, seconds, 0)
for i in range(P):
    hv[2 * i] = np.clip(np.rint(q0 * 8192.0), -32768, 32767).astype(np.int16)
    hv[2 * i + 1] = np.clip(np.rint(r0 * 8192.0), -32768, 32767).astype(np.int16)
bufs = lib.alloc_pair_outputs(P, n, prm, 60.0, features=True, corr=False)
qs = [hv[2 * i] for i in range(P)]; rs = [hv[2 * i + 1] for i in range(P)]
lib.align_pairs_pcm(qs, rs, prm, 60.0, 50, buffers=bufs)
ts = []
for _ in range(3):
    t0 = time.perf_counter(); lib.align_pairs_pcm(qs, rs, prm, 60.0, 50, buffers=bufs); ts.append(time.perf_counter() - t0)
print(os.environ.get("SONAR_PAIR_CHUNK_MB", "default"), os.environ.get("SONAR_PAIR_CHUNK_MAX", "4"), "ms", [round(1e3 * t, 1) for t in ts])
