"""Device-resident fingerprint of BASELINE config[2] (16 kHz, 512/160, 40 mel): per-kernel times for 8 x 1 h streams."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sonido-sonar_b200"); capi, synth = pkg.capi, pkg.synth
lib = capi.SonarLib()
sr, W, H, NS = 16000, 512, 160, 8
x = synth.speech_band_noise(3600.0, sr=sr)
n = x.size; stride = (n + 1) & ~1
p = lib.default_params(window_size=W, hop_size=H, energy_frame=W, energy_hop=H, algo_sample_rate=sr,
                       call_sample_rate=sr, n_mel=40)
dev = torch.zeros((NS, stride), dtype=torch.float64, device="cuda")
dev[:, :n] = torch.from_numpy(x).cuda()
L = lib.fp_dev_layout(p, n)
feat = torch.empty(NS * L.total, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
for _ in range(2):
    lib.fingerprint_batch_dev(dev.data_ptr(), n, stride, NS, p, feat.data_ptr())
lib.synchronize(); lib.profile_enable(True); lib.profile_read()
for _ in range(3):
    lib.fingerprint_batch_dev(dev.data_ptr(), n, stride, NS, p, feat.data_ptr())
prof = lib.profile_read()
T = (n - W) // H + 1
tot = sum(v[0] for v in prof.values()) / 3
print("frames", NS * T, "audio-s/s", NS * 3600.0 / (tot / 1e3))
for k, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:32s} {ms / 3:8.3f} ms  {1e6 * ms / 3 / (NS * T):7.2f} ns/frame")
