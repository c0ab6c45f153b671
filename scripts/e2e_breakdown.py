"""Host-pointer (e2e) leg breakdown: where do the 120 ms go?"""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("sonido-sonar_b200"); capi, synth = pkg.capi, pkg.synth
lib = capi.SonarLib()
P, seconds = 8, 300.0
n = int(seconds * 44100); stride = (n + 1) & ~1; NS = 2 * P
prm = lib.default_params(algo_sample_rate=44100)
host = torch.empty((NS, stride), dtype=torch.float64).pin_memory(); hv = host.numpy()
for i in range(P):
    q, r = bench.make_pair(synth, seconds, i)
    hv[2 * i, :n], hv[2 * i + 1, :n] = q, r
pcm_list = [hv[i, :n] for i in range(NS)]
max_lag = int(60 * 44100) // 256
bufs = lib.alloc_batch_outputs([n] * NS, prm)
for it in range(4):
    t0 = time.perf_counter()
    fps = lib.fingerprint_batch(pcm_list, prm, buffers=bufs); t1 = time.perf_counter()
    eas = [fps[2 * i].short_time_energy for i in range(P)]; ebs = [fps[2 * i + 1].short_time_energy for i in range(P)]
    _, xs = lib.xcorr_batch(eas, ebs, max_lag); t2 = time.perf_counter()
    dl = eas[0].size - max_lag
    qs, rs = zip(*[bench.trim_by_lag(eas[i], ebs[i], xs[i].peak_lag, dl) for i in range(P)])
    out = lib.dtw_batch(list(qs), list(rs), band=50); t3 = time.perf_counter()
    print(f"it{it}: fingerprint_batch {1e3*(t1-t0):.1f}  xcorr_batch {1e3*(t2-t1):.1f}  dtw_batch {1e3*(t3-t2):.1f}  total {1e3*(t3-t0):.1f}")
# raw H2D rate pinned
x = torch.empty(NS * stride, dtype=torch.float64, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); x.copy_(host.view(-1), non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"torch pinned H2D {host.numel()*8/1e9:.2f} GB in {dt*1e3:.1f} ms = {host.numel()*8/1e9/dt:.1f} GB/s")
# python-side alloc cost of outputs
t0 = time.perf_counter(); [lib._alloc_fp(prm, n) for _ in range(NS)]; print(f"output numpy alloc {1e3*(time.perf_counter()-t0):.1f} ms")
