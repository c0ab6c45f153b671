"""Profiling driver: banded DTW of P pairs of length N (dim 1)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sonido-sonar_b200")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 41341
lib = pkg.capi.SonarLib()
rng = np.random.default_rng(0)
qs = [np.abs(np.cumsum(rng.standard_normal(N))) * 0.01 for _ in range(P)]
rs = [x + 0.001 * rng.standard_normal(N) for x in qs]
lib.profile_enable(True)
for it in range(3):
    t0 = time.perf_counter()
    out = lib.dtw_batch(qs, rs, band=50)
    print(f"iter {it}: {1e3*(time.perf_counter()-t0):.2f} ms, path {len(out[0]['path_query'])}")
for k, (ms, cnt) in sorted(lib.profile_read().items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:30s} {ms/cnt:9.3f} ms/launch x{cnt}")
