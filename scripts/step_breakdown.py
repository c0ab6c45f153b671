"""Host-side breakdown of the bench's device-resident step (where does non-kernel time go?)."""
import ctypes as C, importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("sonido-sonar_b200"); capi, synth = pkg.capi, pkg.synth
lib = capi.SonarLib()
ext = torch.cuda.ExternalStream(lib.stream())
P, seconds = 8, 300.0
n = int(seconds * 44100); stride = (n + 1) & ~1; NS = 2 * P
prm = lib.default_params(algo_sample_rate=44100)
sz = lib.fp_sizes(prm, n); L = lib.fp_dev_layout(prm, n); Te = sz.n_energy_frames
max_lag = int(60 * 44100) // 256; dtw_len = Te - max_lag
host = torch.empty((NS, stride), dtype=torch.float64)
for i in range(P):
    q, r = bench.make_pair(synth, seconds, i)
    host[2 * i, :n] = torch.from_numpy(q); host[2 * i + 1, :n] = torch.from_numpy(r)
pcm = host.cuda(); feat = torch.empty(NS * L.total, dtype=torch.float64, device="cuda")
summ = (capi.XcorrSummary * P)()
def T(): torch.cuda.synchronize(); lib.synchronize(); return time.perf_counter()
for it in range(4):
    t0 = T()
    lib.fingerprint_batch_dev(pcm.data_ptr(), n, stride, NS, prm, feat.data_ptr()); t1 = T()
    with torch.cuda.stream(ext):
        e = feat.view(NS, L.total)[:, L.short_time_energy:L.short_time_energy + Te]
        ea, eb = e[0::2].contiguous(), e[1::2].contiguous()
    t2 = T()
    lib._chk(lib.lib.sonar_xcorr_batch_dev(lib.ctx, ea.data_ptr(), Te, eb.data_ptr(), Te, P, max_lag, None, summ)); t3 = T()
    with torch.cuda.stream(ext):
        ea_h, eb_h = ea.cpu().numpy(), eb.cpu().numpy()
    t4 = T()
    qs, rs = [], []
    for i in range(P):
        a, b = bench.trim_by_lag(ea_h[i], eb_h[i], summ[i].peak_lag, dtw_len); qs.append(a); rs.append(b)
    t5 = T()
    out = lib.dtw_batch(qs, rs, band=50); t6 = T()
    print(f"it{it}: fp {1e3*(t1-t0):.1f}  gather {1e3*(t2-t1):.1f}  xcorr {1e3*(t3-t2):.1f}  d2h {1e3*(t4-t3):.1f}  trim {1e3*(t5-t4):.1f}  dtw {1e3*(t6-t5):.1f}  total {1e3*(t6-t0):.1f}")
# dtw internals
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); lib.dtw_batch(qs, rs, band=50); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(8)
