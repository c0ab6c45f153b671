"""Dev diagnostic (GPU box): fingerprint parity GPU vs oracle, per-array error table + quick timing."""
import importlib, os, sys, time, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sonido-sonar_b200")
capi, synth = pkg.capi, pkg.synth
gpu = capi.SonarLib()
ref = capi.SonarLib(os.path.join(ROOT, "oracle", "libsonar_oracle.so"))

def report(tag, a, b):
    print(f"== {tag}")
    for k in b.arrays:
        x, y = a.arrays[k], b.arrays[k]
        if y.size == 0:
            continue
        d = np.abs(x - y)
        scale = np.max(np.abs(y)) + 1e-300
        rel = d / (np.abs(y) + 1e-300)
        print(f"  {k:22s} n={y.size:8d} max|y|={scale:10.4g} maxabs={np.nanmax(d):9.3g} maxabs/scale={np.nanmax(d)/scale:9.3g} "
              f"medrel={np.nanmedian(rel):9.3g} exact={np.array_equal(x, y)} nan={np.isnan(x).sum()}/{np.isnan(y).sum()}")
    for k in b.scalars:
        print(f"  {k:22s} gpu={a.scalars[k]!r} ref={b.scalars[k]!r}")

cases = [
    ("C1-3s sr=44100 1024/256", synth.sweep_noise(3.0, seed=1), dict(algo_sample_rate=44100)),
    ("C1-3s parity sr=0 1024/256", synth.sweep_noise(3.0, seed=1), dict(algo_sample_rate=0)),
    ("C3-5s 16k 512/160 40mel", synth.speech_band_noise(5.0), dict(window_size=512, hop_size=160, energy_frame=512,
        energy_hop=160, algo_sample_rate=16000, call_sample_rate=16000, n_mel=40)),
    ("2048/512", synth.sweep_noise(2.0, seed=3), dict(window_size=2048, hop_size=512, energy_frame=2048, energy_hop=512, algo_sample_rate=44100)),
    ("256/64 odd n", synth.sweep_noise(1.0, seed=5)[:40001], dict(window_size=256, hop_size=64, energy_frame=256, energy_hop=64, algo_sample_rate=44100)),
]
for tag, pcm, kw in cases:
    p = gpu.default_params(**kw)
    try:
        a = gpu.fingerprint(pcm, p)
        b = ref.fingerprint(pcm, p)
        report(tag, a, b)
    except Exception as e:
        print("FAIL", tag, repr(e))

# stft
pcm = synth.sweep_noise(1.0, seed=7)
try:
    mg, ph, cx = gpu.stft(pcm, 1024, 256, phase=True, cplx=True)
    mr, pr_, cr = ref.stft(pcm, 1024, 256, phase=True, cplx=True)
    print("stft mag maxabs/scale", np.max(np.abs(mg - mr)) / np.max(mr), "cplx", np.max(np.abs(cx - cr)) / np.max(np.abs(cr)))
except Exception as e:
    print("FAIL stft", repr(e))

# timing: device-resident batch
import torch
ns, secs = 64, 60.0
n = int(secs * 44100)
stride = (n + 1) & ~1
x = torch.from_numpy(synth.sweep_noise(secs, seed=9)).cuda()
pcm_dev = x.repeat(ns, 1).contiguous()
for sr in (0, 44100):
    p = gpu.default_params(algo_sample_rate=sr)
    L = gpu.fp_dev_layout(p, n)
    feat = torch.empty(ns * L.total, dtype=torch.float64, device="cuda")
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        gpu.fingerprint_batch_dev(pcm_dev.data_ptr(), n, n, ns, p, feat.data_ptr())
        gpu.synchronize(); dt = time.time() - t0
        print(f"batch_dev sr={sr}: {ns} x {secs}s in {dt*1e3:.2f} ms -> {ns*secs/dt:.3g} audio-s/s")
print("launches", gpu.kernel_launches())
