"""One process driving several GPUs (sonar_init(n_devices > 1): the batch entry points shard streams / pairs over the
devices with one host thread each).  Needs >= 2 visible GPUs; skipped otherwise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_devices_in_one_process_equal_one(gpu, capi, synth):
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    two = capi.SonarLib(n_devices=2)
    try:
        p = gpu.default_params(algo_sample_rate=44100)
        pcms = [synth.sweep_noise(s, seed=80 + i) for i, s in enumerate((1.0, 1.7, 1.0, 2.2, 0.6))]
        a, b = gpu.fingerprint_batch(pcms, p), two.fingerprint_batch(pcms, p)
        for x, y in zip(a, b):
            for k in x.arrays:
                assert np.array_equal(x.arrays[k], y.arrays[k]), k
        pairs = [synth.aligned_pair(6.0, offset_seconds=o, seed=90 + i) for i, o in enumerate((0.5, -0.3, 1.1))]
        qs, rs = [q for q, _ in pairs], [r for _, r in pairs]
        ra, rb = gpu.align_pairs(qs, rs, p, 1.5, 50), two.align_pairs(qs, rs, p, 1.5, 50)
        q16 = [np.clip(np.round(q * 20000), -32768, 32767).astype(np.int16) for q in qs]
        r16 = [np.clip(np.round(r * 20000), -32768, 32767).astype(np.int16) for r in rs]
        sa, sb = gpu.align_pairs_pcm(q16, r16, p, 1.5, 50), two.align_pairs_pcm(q16, r16, p, 1.5, 50)
        for x, y in list(zip(ra, rb)) + list(zip(sa, sb)):
            assert x["xcorr"].peak_lag == y["xcorr"].peak_lag and np.array_equal(x["corr"], y["corr"])
            assert np.array_equal(x["path_query"], y["path_query"]) and np.array_equal(x["path_ref"], y["path_ref"])
            assert np.array_equal(x["query"].mfcc, y["query"].mfcc)
    finally:
        two.close()
