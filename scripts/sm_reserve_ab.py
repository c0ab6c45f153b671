"""A/B of SONAR_STFT_SM_RESERVE inside one process: the bench's device-resident pair step (32 pairs x 2 x 300 s)."""
import ctypes as C, importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("sonido-sonar_b200"); capi, synth = pkg.capi, pkg.synth
lib = capi.SonarLib()
ext = torch.cuda.ExternalStream(lib.stream())
P, seconds = 32, 300.0
n = int(round(seconds * 44100)); stride = (n + 1) & ~1; NS = 2 * P
prm = lib.default_params(algo_sample_rate=44100, call_sample_rate=44100)
host = torch.empty((NS, stride), dtype=torch.float64)
q, r = bench.make_pair(synth, seconds, 0)
for i in range(P):  # one synthetic pair, rolled: the timing does not depend on the content
    host[2 * i, :n] = torch.from_numpy(np.roll(q, 1000 * i)); host[2 * i + 1, :n] = torch.from_numpy(np.roll(r, 1000 * i))
pcm = host.cuda()
bufs = lib.alloc_pair_outputs(P, n, prm, bench.MAX_LAG_S, features=False, corr=False)
def step():
    return lib.align_pairs_dev(pcm.data_ptr(), n, stride, P, prm, bench.MAX_LAG_S, bench.DTW_BAND, buffers=bufs)
for _ in range(3): step()
order = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,8,0,8,16,0,16,4".split(","))]
for rsv in order:
    os.environ["SONAR_STFT_SM_RESERVE"] = str(rsv)
    step(); torch.cuda.synchronize(); lib.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(8): step()
    e1.record(ext); torch.cuda.synchronize(); lib.synchronize()
    print(f"reserve {rsv:3d}: {e0.elapsed_time(e1) / 8:.3f} ms per step", flush=True)
