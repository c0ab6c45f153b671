"""GPU parity: cross-correlation (bit-exact lag + correlations), lag sharding, alignment scalars."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SUMMARY_FLOAT = ("peak_correlation", "p_value", "snr", "sharpness", "second_peak", "peak_to_sidelobe")
SUMMARY_INT = ("peak_lag", "peak_index", "actual_max_lag", "overlap_length", "is_significant")


def same_float(a, b, rtol=1e-9):
    if math.isinf(a) or math.isinf(b) or math.isnan(a) or math.isnan(b):
        return (math.isnan(a) and math.isnan(b)) or a == b
    return abs(a - b) <= rtol * max(1.0, abs(b))


def check_summary(sa, sb):
    for k in SUMMARY_INT:
        assert getattr(sa, k) == getattr(sb, k), k
    # the correlations are bit-exact, so the peak is too; the reductions use another order -> 1e-9
    assert sa.peak_correlation == sb.peak_correlation
    for k in SUMMARY_FLOAT:
        assert same_float(getattr(sa, k), getattr(sb, k)), (k, getattr(sa, k), getattr(sb, k))


def energies(gpu, synth, seconds, offset, seed=2):
    q, r = synth.aligned_pair(seconds, offset_seconds=offset, seed=seed)
    p = gpu.default_params(algo_sample_rate=44100)
    return gpu.fingerprint(q, p).short_time_energy, gpu.fingerprint(r, p).short_time_energy


@pytest.mark.parametrize("na,nb,max_lag", [(500, 500, 100), (777, 512, 300), (64, 300, 1000), (3, 3, 5),
                                           (1, 9, 4), (2000, 2000, 0)])
def test_xcorr_random_bit_exact(gpu, oracle, na, nb, max_lag):
    rng = np.random.default_rng(na * 31 + nb)
    a, b = rng.standard_normal(na), rng.standard_normal(nb)
    ca, sa = gpu.xcorr(a, b, max_lag)
    cb, sb = oracle.xcorr(a, b, max_lag)
    assert ca.shape == cb.shape
    assert np.array_equal(ca, cb)  # sequential f64 sums per lag: bit-exact
    check_summary(sa, sb)


def test_xcorr_degenerate_inputs(gpu, oracle):
    # constant sequences: sigma < 1e-10 -> only de-meaned -> every correlation is 0 -> peak index 0
    a, b = np.full(300, 0.25), np.full(280, -1.5)
    ca, sa = gpu.xcorr(a, b, 50)
    cb, sb = oracle.xcorr(a, b, 50)
    assert np.array_equal(ca, cb)
    check_summary(sa, sb)
    assert sa.peak_index == 0
    # exact ties: periodic signal, the first maximum must win (correlation.go:535-541)
    t = np.arange(400)
    a = np.sign(np.sin(2 * np.pi * t / 20.0) + 1e-9)
    ca, sa = gpu.xcorr(a, a, 100)
    cb, sb = oracle.xcorr(a, a, 100)
    assert np.array_equal(ca, cb)
    check_summary(sa, sb)


def test_xcorr_errors(gpu, capi):
    with pytest.raises(capi.SonarError) as e:
        gpu.xcorr(np.zeros(0), np.zeros(5), 3)
    assert e.value.code == capi.ERR_EMPTY and "empty signals provided" in e.value.msg


def test_alignment_known_offset(gpu, oracle, synth):
    ea, eb = energies(gpu, synth, 40.0, 3.7)
    ca, xa, ra = gpu.align_xcorr(ea, eb, int(30 * 44100) // 256, 256, 44100, want_corr=True)
    cb, xb, rb = oracle.align_xcorr(ea, eb, int(30 * 44100) // 256, 256, 44100, want_corr=True)
    n = 2 * xa.actual_max_lag + 1
    assert np.array_equal(ca[:n], cb[:n])
    check_summary(xa, xb)
    assert abs(xa.peak_lag - 3.7 * 44100 / 256) <= 1.0  # positive lag = CDN later
    assert ra.offset == rb.offset == xa.peak_lag * 256
    for k in ("offset_seconds", "confidence", "similarity", "alignment_quality", "noise_level"):
        assert same_float(getattr(ra, k), getattr(rb, k)), k
    assert ra.method == 1 and ra.query_length == ea.size and ra.reference_length == eb.size


def test_alignment_negative_offset(gpu, oracle, synth):
    ea, eb = energies(gpu, synth, 30.0, -2.1, seed=11)
    _, xa, ra = gpu.align_xcorr(ea, eb, 2000, 256, 44100)
    _, xb, rb = oracle.align_xcorr(ea, eb, 2000, 256, 44100)
    check_summary(xa, xb)
    assert xa.peak_lag < 0 and ra.offset == rb.offset


def test_xcorr_batch_ragged(gpu, oracle):
    rng = np.random.default_rng(5)
    As = [rng.standard_normal(n) for n in (300, 1000, 17, 512, 999)]
    Bs = [rng.standard_normal(n) for n in (280, 1000, 400, 100, 999)]
    ca, sa = gpu.xcorr_batch(As, Bs, 128, want_corr=True)
    cb, sb = oracle.xcorr_batch(As, Bs, 128, want_corr=True)
    for i in range(len(As)):
        n = 2 * sa[i].actual_max_lag + 1
        assert np.array_equal(ca[i][:n], cb[i][:n])
        check_summary(sa[i], sb[i])


def test_xcorr_lag_sharding_matches_whole(gpu, oracle, capi, synth):
    """one correlation split by lag range; merge of per-shard maxima == unsharded result."""
    ea, eb = energies(gpu, synth, 30.0, 5.2, seed=21)
    max_lag = 1500
    _, whole = oracle.xcorr(ea, eb, max_lag)
    nl = 2 * whole.actual_max_lag + 1
    for n_shards in (1, 2, 3, 8):
        step = -(-nl // n_shards)
        shards, peaks = [], []
        for g in range(n_shards):
            sh, pk = gpu.xcorr_shard(ea, eb, max_lag, g * step, min(nl, (g + 1) * step))
            shards.append(sh)
            peaks.append(pk)
        gidx = gpu.xcorr_merge_peaks(peaks)
        assert gidx == whole.peak_index
        parts = [gpu.xcorr_shard_metrics(sh, gidx) for sh in shards]
        merged = gpu.xcorr_merge_metrics(parts, ea.size, eb.size, max_lag, gidx)
        check_summary(merged, whole)
        corr = np.concatenate([gpu.xcorr_shard_corr(sh, min(nl, (g + 1) * step) - g * step)
                               for g, sh in enumerate(shards)])
        cw, _ = oracle.xcorr(ea, eb, max_lag)
        assert np.array_equal(corr, cw)
        for sh in shards:
            gpu.xcorr_shard_close(sh)


def test_chained_pair_pipeline_equals_the_separate_calls(gpu, oracle, synth):
    """sonar_align_pairs_f64 / _dev (fingerprint x2 -> NCC -> trim -> banded DTW, chained on the device) must
    return exactly what the individual entry points return, and match the oracle's composition."""
    import torch
    p = gpu.default_params(algo_sample_rate=44100)
    secs, max_lag_s, band = 12.0, 3.0, 50
    pairs = [synth.aligned_pair(secs, offset_seconds=o, seed=40 + i) for i, o in enumerate((1.1, -0.7, 0.0, 2.4))]
    qs, rs = [a for a, _ in pairs], [b for _, b in pairs]
    got = gpu.align_pairs(qs, rs, p, max_lag_s, band)
    ref = oracle.align_pairs(qs, rs, p, max_lag_s, band)
    n = qs[0].size
    nl, dl = gpu.align_pairs_sizes(p, n, max_lag_s)
    assert (nl, dl) == oracle.align_pairs_sizes(p, n, max_lag_s)
    for g, o, q, r in zip(got, ref, qs, rs):
        fq, fr = gpu.fingerprint(q, p), gpu.fingerprint(r, p)
        for k in fq.arrays:  # the chained call runs the same kernels: identical bits
            assert np.array_equal(g["query"].arrays[k], fq.arrays[k]), k
            assert np.array_equal(g["reference"].arrays[k], fr.arrays[k]), k
        assert np.array_equal(g["corr"], o["corr"])  # bit-exact energies -> bit-exact correlations
        check_summary(g["xcorr"], o["xcorr"])
        assert g["corr_alignment"].offset == o["corr_alignment"].offset
        assert g["dtw_length"] == o["dtw_length"] == dl
        assert np.array_equal(g["path_query"], o["path_query"]) and np.array_equal(g["path_ref"], o["path_ref"])
        assert np.array_equal(g["path_cost"], o["path_cost"], equal_nan=True)
        assert g["total_cost"] == o["total_cost"]
    # device-resident form
    stride = (n + 1) & ~1
    dev = torch.zeros((2 * len(qs), stride), dtype=torch.float64, device="cuda")
    for i, (q, r) in enumerate(zip(qs, rs)):
        dev[2 * i, :n] = torch.from_numpy(q).cuda()
        dev[2 * i + 1, :n] = torch.from_numpy(r).cuda()
    torch.cuda.synchronize()
    got_dev = gpu.align_pairs_dev(dev.data_ptr(), n, stride, len(qs), p, max_lag_s, band)
    for g, d in zip(got, got_dev):
        assert g["xcorr"].peak_lag == d["xcorr"].peak_lag and np.array_equal(g["corr"], d["corr"])
        assert np.array_equal(g["path_query"], d["path_query"]) and np.array_equal(g["path_ref"], d["path_ref"])


@pytest.mark.parametrize("dtype", [np.int16, np.float32])
def test_pair_pipeline_pcm_ingest_equals_the_widened_float64_call(gpu, oracle, synth, dtype):
    """sonar_align_pairs_pcm: int16 / float32 PCM widened on the device must give exactly what the
    float64 entry point gives on the samples the reference's decoder would have produced (x / 32768, exact
    float -> double), and the oracle's own widening must agree bit for bit on the integer / exact outputs."""
    p = gpu.default_params(algo_sample_rate=44100)
    secs, max_lag_s, band = 8.0, 2.0, 50
    pairs = [synth.aligned_pair(secs, offset_seconds=o, seed=70 + i) for i, o in enumerate((0.9, -0.4, 1.7))]
    if dtype == np.int16:
        narrow = [(np.clip(np.round(a * 20000.0), -32768, 32767).astype(np.int16),
                   np.clip(np.round(b * 20000.0), -32768, 32767).astype(np.int16)) for a, b in pairs]
        wide = [(a.astype(np.float64) / 32768.0, b.astype(np.float64) / 32768.0) for a, b in narrow]
    else:
        narrow = [(a.astype(np.float32), b.astype(np.float32)) for a, b in pairs]
        wide = [(a.astype(np.float64), b.astype(np.float64)) for a, b in narrow]
    got = gpu.align_pairs_pcm([a for a, _ in narrow], [b for _, b in narrow], p, max_lag_s, band)
    ref = gpu.align_pairs([a for a, _ in wide], [b for _, b in wide], p, max_lag_s, band)
    ora = oracle.align_pairs_pcm([a for a, _ in narrow], [b for _, b in narrow], p, max_lag_s, band)
    for g, r, o in zip(got, ref, ora):
        for k in r["query"].arrays:
            assert np.array_equal(g["query"].arrays[k], r["query"].arrays[k]), k
            assert np.array_equal(g["reference"].arrays[k], r["reference"].arrays[k]), k
        assert np.array_equal(g["corr"], r["corr"]) and np.array_equal(g["corr"], o["corr"])
        assert g["xcorr"].peak_lag == r["xcorr"].peak_lag == o["xcorr"].peak_lag
        assert np.array_equal(g["path_query"], o["path_query"]) and np.array_equal(g["path_ref"], o["path_ref"])
        assert np.array_equal(g["path_cost"], r["path_cost"], equal_nan=True)


def test_pcm_entry_points_reject_bad_arguments(gpu, oracle, synth):
    """Error behaviour of the ingest entry points: unknown sample format, nil audio, too-short streams."""
    import ctypes as C
    p = gpu.default_params(algo_sample_rate=44100)
    x = np.zeros(44100, np.int16)
    for lib in (gpu, oracle):
        outs, keep = lib.alloc_pair_outputs(1, x.size, p, 0.5)
        ptr = (C.c_void_p * 1)(x.ctypes.data)
        rc = lib.lib.sonar_align_pairs_pcm(lib.ctx, ptr, ptr, 7, x.size, 1, C.byref(p), 0.5, 50, outs)
        assert rc != 0 and b"unknown PCM sample format" in lib.lib.sonar_last_error()
        null = (C.c_void_p * 1)(None)
        rc = lib.lib.sonar_align_pairs_pcm(lib.ctx, ptr, null, 2, x.size, 1, C.byref(p), 0.5, 50, outs)
        assert rc != 0 and b"audio data cannot be nil" in lib.lib.sonar_last_error()
        with pytest.raises(Exception, match="signal too short"):
            lib.align_pairs_pcm([np.zeros(100, np.int16)], [np.zeros(100, np.int16)], p, 0.5, 50)
    # silence in, silence out: all-zero int16 streams give zero energies and an all-zero correlation curve
    res = gpu.align_pairs_pcm([x], [x], p, 0.5, 50)[0]
    assert not res["query"].short_time_energy.any() and not res["corr"].any() and res["xcorr"].peak_lag == \
        oracle.align_pairs_pcm([x], [x], p, 0.5, 50)[0]["xcorr"].peak_lag


def _screen_cases(synth):
    sr, secs = 44100, 12.0
    n = int(secs * sr)
    t = np.arange(n) / sr
    rng = np.random.default_rng(7)
    cases = {}
    cases["offset"] = synth.aligned_pair(secs, offset_seconds=1.9, seed=51)
    cases["negative"] = synth.aligned_pair(secs, offset_seconds=-2.2, seed=52)
    q, _ = synth.aligned_pair(secs, offset_seconds=0.0, seed=53)
    cases["identical"] = (q, q.copy())  # c(lag) == c(-lag) bit for bit: the first index must win every tie
    tone = 0.4 * np.sin(2 * np.pi * 440.0 * t)
    cases["steady_tone"] = (tone, 0.5 * tone)  # near-constant energies: flat, tie-ridden curve
    am = (0.5 + 0.5 * np.sin(2 * np.pi * 1.0 * t)) * np.sin(2 * np.pi * 300.0 * t)
    cases["periodic_envelope"] = (am, np.roll(am, 4410))  # one maximum per envelope period
    cases["unrelated_noise"] = (rng.standard_normal(n) * 0.1, rng.standard_normal(n) * 0.1)
    cases["silence"] = (np.zeros(n), np.zeros(n))
    burst = np.zeros(n)
    burst[1000:1400] = 0.8
    cases["single_burst"] = (burst, np.roll(burst, 30000))
    return cases


def test_screened_pair_pipeline_matches_the_full_curve(gpu, oracle, synth):
    """Without a correlation buffer the pipeline screens the lags with an FFT and evaluates only the candidate blocks
    in reference order (csrc/xcorr_fft.cu).  Lag, peak value and DTW path must still be bit-identical to the oracle's
    full evaluation; the reductions over the rest of the curve agree to 1e-9."""
    p = gpu.default_params(algo_sample_rate=44100)
    max_lag_s, band = 3.0, 50
    cases = _screen_cases(synth)
    qs, rs = [c[0] for c in cases.values()], [c[1] for c in cases.values()]
    bufs = gpu.alloc_pair_outputs(len(qs), qs[0].size, p, max_lag_s, features=False, corr=False)
    got = gpu.align_pairs(qs, rs, p, max_lag_s, band, buffers=bufs)
    ref = oracle.align_pairs(qs, rs, p, max_lag_s, band)
    full = gpu.align_pairs(qs, rs, p, max_lag_s, band)  # with curve: every lag exact
    for name, g, o, f in zip(cases, got, ref, full):
        assert g["corr"] is None
        check_summary(g["xcorr"], o["xcorr"])
        assert g["xcorr"].peak_correlation == f["xcorr"].peak_correlation, name
        assert g["xcorr"].second_peak == f["xcorr"].second_peak, name  # second peak is verified exactly too
        assert g["corr_alignment"].offset == o["corr_alignment"].offset, name
        assert np.array_equal(g["path_query"], o["path_query"]) and np.array_equal(g["path_ref"], o["path_ref"]), name
        assert np.array_equal(g["path_cost"], o["path_cost"], equal_nan=True), name
        assert g["total_cost"] == o["total_cost"], name


def _xcorr_screen_inputs():
    rng = np.random.default_rng(99)
    t = np.arange(400)
    square = np.sign(np.sin(2 * np.pi * t / 20.0) + 1e-9)
    spike = np.zeros(5000)
    spike[[100, 2500, 4100]] = [1e6, -3.0, 2.0]  # huge dynamic range, tiny overlap energies at the ends
    smooth = np.cumsum(rng.standard_normal(30000))  # random walk: broad peak, neighbours within 1e-5
    cases = [(rng.standard_normal(na), rng.standard_normal(nb), ml)
             for na, nb, ml in [(500, 500, 100), (777, 512, 300), (64, 300, 1000), (3, 3, 5), (1, 9, 4), (2000, 2000, 0),
                                (40000, 35000, 9000)]]
    cases += [(np.full(300, 0.25), np.full(280, -1.5), 50), (square, square, 100), (square, -square, 150),
              (spike, np.roll(spike, 37), 4000), (smooth, np.roll(smooth, -211), 3000),
              (np.zeros(100), rng.standard_normal(100), 20)]
    return cases


def test_xcorr_without_curve_is_screened_and_keeps_the_peak_exact(gpu, oracle):
    """want_corr=False: FFT screen + exact candidate blocks (csrc/xcorr_fft.cu) -- the summary must equal the one the
    full reference-order evaluation gives: indices and the peak / second-peak values bit for bit."""
    for a, b, ml in _xcorr_screen_inputs():
        _, ss = gpu.xcorr(a, b, ml, want_corr=False)
        _, sf = gpu.xcorr(a, b, ml, want_corr=True)
        _, so = oracle.xcorr(a, b, ml)
        check_summary(ss, so)
        assert ss.peak_correlation == sf.peak_correlation and ss.second_peak == sf.second_peak, (a.size, b.size, ml)
        assert same_float(ss.sharpness, sf.sharpness, rtol=1e-12), (ss.sharpness, sf.sharpness)
    As, Bs = [c[0] for c in _xcorr_screen_inputs()[:3]], [c[1] for c in _xcorr_screen_inputs()[:3]]
    _, sa = gpu.xcorr_batch(As, Bs, 128, want_corr=False)  # ragged batch through one screen geometry
    _, sb = oracle.xcorr_batch(As, Bs, 128, want_corr=False)
    for x, y in zip(sa, sb):
        check_summary(x, y)


def test_library_side_lag_shard_with_its_own_nccl_communicator(gpu, oracle):
    """sonar_xcorr_lag_sharded on a one-rank communicator (dlopen of libnccl, ncclCommInitRank, the in-place all-gather
    path is skipped at world 1 but the shard geometry, the resident workspace and the summary are the N-rank ones):
    identical to the unsharded evaluation and to the oracle, curve included.  The N > 1 form runs in bench.py --gpus N
    (`lag_sharded.matches_unsharded`)."""
    rng = np.random.default_rng(11)
    base = np.convolve(rng.standard_normal(9000), np.ones(16) / 16, mode="same") + 1.0
    a, b = base[500:8500].copy(), base[500 - 137:8500 - 137] + 0.01 * rng.standard_normal(8000)
    try:
        gpu.nccl_init(1, 0, gpu.nccl_unique_id())
        s, corr = gpu.xcorr_lag_sharded(a, b, 700, want_corr=True)
    finally:
        gpu.nccl_shutdown()
    corr0, s0 = gpu.xcorr(a, b, 700, want_corr=True)
    corr1, s1 = oracle.xcorr(a, b, 700, want_corr=True)
    assert np.array_equal(corr, corr0) and np.array_equal(corr, corr1)
    for k in ("peak_lag", "peak_index", "peak_correlation", "second_peak", "snr", "sharpness", "peak_to_sidelobe"):
        assert getattr(s, k) == getattr(s0, k), k
    for k in ("peak_lag", "peak_index", "peak_correlation", "second_peak"):
        assert getattr(s, k) == getattr(s1, k), k
    for k in ("snr", "sharpness", "peak_to_sidelobe"):
        assert getattr(s, k) == pytest.approx(getattr(s1, k), rel=1e-9), k
    s2, _ = gpu.xcorr_lag_sharded(a, b, 700)  # no communicator: one shard
    assert s2.peak_lag == s.peak_lag and s2.peak_correlation == s.peak_correlation


def test_cdn_latency_loop_align_truncate_refingerprint_compare(gpu, oracle, synth):
    """SURVEY §8 f3 / VERDICT r1 missing #6: the whole loop around the hot path -- ExtractAlignmentFeatures on both
    streams, TruncateToAlignmentPCM by the detected offset (extractors/alignment.go:223-297), GenerateFingerprint of
    the two aligned segments, FingerprintComparator.Compare -- through the C ABI, against the same chain on the oracle."""
    sr, hop = 44100, 256
    q, r = synth.aligned_pair(20.0, offset_seconds=2.37, sr=sr, seed=77)
    p = gpu.default_params(algo_sample_rate=sr, call_sample_rate=sr)
    w = [0.5, 0.2, 0.0, 0.1, 0.0, 0.2, 0.0]

    def chain(lib):
        res = lib.align_pairs([q], [r], p, 5.0, 50)[0]
        off_s = res["corr_alignment"].offset_seconds
        s1, s2, n = lib.truncate_to_alignment(q.size, r.size, sr, off_s)
        f1, f2 = lib.fingerprint_batch([q[s1:s1 + n], r[s2:s2 + n]], p)
        c1, k1 = lib.cmp_features(f1)
        c2, k2 = lib.cmp_features(f2)
        return res, (s1, s2, n), f1, f2, lib.compare(c1, c2, w).as_dict()

    rg, tg, g1, g2, cg = chain(gpu)
    ro, to, o1, o2, co = chain(oracle)
    assert rg["xcorr"].peak_lag == ro["xcorr"].peak_lag and rg["corr_alignment"].offset == ro["corr_alignment"].offset
    assert tg == to
    for a, b in ((g1, o1), (g2, o2)):
        assert np.array_equal(a.short_time_energy, b.short_time_energy)
        assert np.array_equal(a.spectral_rolloff, b.spectral_rolloff)
        np.testing.assert_allclose(a.mfcc, b.mfcc, rtol=1e-4, atol=1e-4 * np.max(np.abs(b.mfcc)))
    for k in cg:
        assert cg[k] == pytest.approx(co[k], rel=1e-6, abs=1e-9, nan_ok=True), k
    assert cg["dist_mfcc"] < 1e-3 and cg["overall_similarity"] > 0.7  # the aligned segments are the same programme


@pytest.mark.parametrize("scale", [1.6e-14, 2.0e-14, 2.6e-14])
def test_screened_ncc_denominators_at_the_1e_10_rule(gpu, oracle, scale):
    """one sequence takes z-scoring's non-normalised branch (sigma < 1e-10, correlation.go:464-501), so the
    per-lag denominators sit around the reference's `den < 1e-10 -> 0` rule (correlation.go:401-405) and change side
    with the overlap length.  The screened form (no curve requested) must land on the reference's peak, second peak and
    lag all the same: lags whose prefix-sum denominator is near the rule are evaluated in the reference's order."""
    rng = np.random.default_rng(int(scale * 1e16))
    n = 5000
    a = rng.standard_normal(n)
    b = 3.0 + scale * np.roll(a, 17) + 0.3 * scale * rng.standard_normal(n)
    assert np.std(b) < 1e-10
    _, s = gpu.xcorr(a, b, 400, want_corr=False)
    corr, so = oracle.xcorr(a, b, 400, want_corr=True)
    if scale == 2.0e-14:  # both sides of the rule occur on one curve (1.6e-14: every lag below it, 2.6e-14: none)
        assert np.count_nonzero(corr == 0.0) > 0 and np.count_nonzero(corr) > 0
    assert (s.peak_lag, s.peak_index) == (so.peak_lag, so.peak_index)
    assert s.peak_correlation == so.peak_correlation and s.second_peak == so.second_peak
    cg, _ = gpu.xcorr(a, b, 400, want_corr=True)
    assert np.array_equal(cg, corr)


def test_pair_pipeline_long_wandering_path_is_fetched_completely(gpu, oracle, synth):
    """The pair pipeline's result copy carries the tail of each path array (sequence length + 12.5 %); two unrelated
    streams make the DTW wander inside its band, so the path is longer than that and its head is fetched on demand:
    path and costs must still be the oracle's, bit for bit."""
    sr = 44100
    q = synth.envelope_noise(int(12 * sr), sr, seed=91)
    r = synth.envelope_noise(int(12 * sr), sr, seed=92)
    p = gpu.default_params(algo_sample_rate=sr, call_sample_rate=sr)
    g = gpu.align_pairs([q], [r], p, 2.0, 50)[0]
    o = oracle.align_pairs([q], [r], p, 2.0, 50)[0]
    assert g["path_query"].size == o["path_query"].size
    assert g["path_query"].size > 1.125 * (o["path_query"].max() + 1) + 64, "the case must exceed the travelling tail"
    assert np.array_equal(g["path_query"], o["path_query"]) and np.array_equal(g["path_ref"], o["path_ref"])
    assert np.array_equal(g["path_cost"], o["path_cost"], equal_nan=True)
