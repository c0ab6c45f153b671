"""Pins for the CPU oracle (oracle/sonar_oracle.cpp).

The reference ships no tests, no golden vectors and cannot be compiled here (no Go toolchain;
SURVEY.md §4, §8c), so parity is *unpinned by the reference itself*.  What pins the oracle instead:
  * known answers derived from the reference's formulas (window end points / power normalisation,
    DCT row 0, the sampleRate=0 degenerate constants of SURVEY F3, pure-tone centroid, shifted-copy lag);
  * independent numpy restatements of the same Go code written from the reference source
    (file:line cited at each), compared with the C++ oracle;
  * the committed fixtures under tests/golden/ (regression pins, see tests/golden/make_golden.py).
"""
import math
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---------------------------------------------------------------- windows / STFT

def test_hann_known_answers(oracle):
    # analyzers/windowing.go:246-256 (symmetric: divide by N-1), :427-437 (power normalisation)
    w = oracle.window("hann", 1024, symmetric=True, normalize=False)
    assert w[0] == 0.0 and abs(w[-1]) < 1e-30 and np.allclose(w, w[::-1], atol=1e-15)
    i = np.arange(1024)
    assert np.allclose(w, 0.5 * (1 - np.cos(2 * np.pi * i / 1023)), rtol=0, atol=1e-16)
    wn = oracle.window("hann", 1024)
    assert abs((wn ** 2).sum() / 1024 - 1.0) < 1e-13
    assert wn[512] / w[512] == pytest.approx(1.63379, rel=1e-5)
    wp = oracle.window("hann", 1024, symmetric=False, normalize=False)
    assert np.allclose(wp, 0.5 * (1 - np.cos(2 * np.pi * i / 1024)), atol=1e-16)


@pytest.mark.parametrize("name,formula", [
    ("hamming", lambda a, i, n: 0.54 - 0.46 * np.cos(a)),
    ("blackman", lambda a, i, n: 0.42 - 0.5 * np.cos(a) + 0.08 * np.cos(2 * a)),
    ("blackman_harris", lambda a, i, n: 0.35875 - 0.48829 * np.cos(a) + 0.14128 * np.cos(2 * a) - 0.01168 * np.cos(3 * a)),
    ("rectangular", lambda a, i, n: np.ones_like(a)),
    ("welch", lambda a, i, n: 1 - ((i - (n - 1) / 2) / ((n - 1) / 2)) ** 2),
    ("bartlett", lambda a, i, n: np.where(i <= n // 2, 2 * i / (n - 1), 2 - 2 * i / (n - 1))),
])
def test_other_windows(oracle, name, formula):
    n = 257
    i = np.arange(n, dtype=np.float64)
    a = 2 * np.pi * i / (n - 1)
    w = oracle.window(name, n, symmetric=True, normalize=False)
    assert np.allclose(w, formula(a, i, n), atol=1e-14)


def np_stft(pcm, win, hop, w):
    T = (pcm.size - win) // hop + 1  # analyzers/spectral.go:409
    idx = np.arange(win)[None, :] + hop * np.arange(T)[:, None]
    return np.fft.rfft(pcm[idx] * w[None, :], axis=1)


def test_stft_against_numpy_rfft(oracle, synth):
    pcm = synth.sweep_noise(0.5, seed=1)
    for win, hop in ((1024, 256), (512, 160)):
        mag, ph, cx = oracle.stft(pcm, win, hop, phase=True, cplx=True)
        X = np_stft(pcm, win, hop, oracle.window("hann", win))
        assert mag.shape == X.shape
        scale = np.abs(X).max()
        assert np.max(np.abs(mag - np.abs(X))) < 1e-11 * scale
        assert np.max(np.abs(cx[..., 0] - X.real)) < 1e-11 * scale and np.max(np.abs(cx[..., 1] - X.imag)) < 1e-11 * scale
        strong = np.abs(X) > 1e-6 * scale
        assert np.max(np.abs(np.angle(np.exp(1j * (ph - np.angle(X))))[strong])) < 1e-8


# ---------------------------------------------------------------- MFCC + spectral descriptors

def np_mel_bank(n_mel, fft_size, sr, lo, hi):
    # algorithms/spectral/mel_scale.go:29-86
    hz2mel = lambda f: 2595.0 * np.log10(1.0 + f / 700.0)
    mel2hz = lambda m: 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    mels = hz2mel(lo) + np.arange(n_mel + 2) * (hz2mel(hi) - hz2mel(lo)) / (n_mel + 1)
    bins = np.minimum(np.floor((fft_size + 1.0) * mel2hz(mels) / sr + 0.5).astype(int), fft_size // 2)
    fb = np.zeros((n_mel, fft_size // 2 + 1))
    for m in range(1, n_mel + 1):
        l, c, r = bins[m - 1], bins[m], bins[m + 1]
        for k in range(l, c):
            fb[m - 1, k] = (k - l) / (c - l)
        for k in range(c, r):
            fb[m - 1, k] = (r - k) / (r - c)
    return fb


def np_mfcc(mag, sr, n_mfcc=13, n_mel=26, lifter=22.0):
    # algorithms/spectral/mfcc.go:113-164 (power, filter bank, ln with 1e-10 floor, DCT-II, lifter)
    B = mag.shape[1]
    fb = np_mel_bank(n_mel, (B - 1) * 2, sr, 0.0, sr / 2.0)
    mel = (mag ** 2) @ fb.T
    lg = np.where(mel > 0, np.log(np.where(mel > 0, mel, 1.0)), math.log(1e-10))
    k = np.arange(n_mfcc)[:, None]
    n = np.arange(n_mel)[None, :]
    D = np.cos(np.pi * k * (n + 0.5) / n_mel) * np.where(k == 0, math.sqrt(1.0 / n_mel), math.sqrt(2.0 / n_mel))
    c = lg @ D.T
    c[:, 1:] *= 1.0 + (lifter / 2.0) * np.sin(np.pi * np.arange(1, n_mfcc) / lifter)
    return c


def test_dct_row0_known_answer():
    # mfcc.go:205-209: D[0][n] = sqrt(1/M)
    assert math.sqrt(1.0 / 26) == pytest.approx(0.19611613513818404)


def test_mfcc_against_numpy_restatement(oracle, synth):
    pcm = synth.sweep_noise(1.0, seed=2)
    p = oracle.default_params(algo_sample_rate=44100)
    fp = oracle.fingerprint(pcm, p)
    mag, _, _ = oracle.stft(pcm, 1024, 256)
    ref = np_mfcc(mag, 44100)
    assert np.allclose(fp.mfcc, ref, rtol=1e-9, atol=1e-9)
    p40 = oracle.default_params(algo_sample_rate=16000, call_sample_rate=16000, window_size=512, hop_size=160,
                                energy_frame=512, energy_hop=160, n_mel=40)
    x = synth.speech_band_noise(1.0)
    mag, _, _ = oracle.stft(x, 512, 160)
    assert np.allclose(oracle.fingerprint(x, p40).mfcc, np_mfcc(mag, 16000, n_mel=40), rtol=1e-9, atol=1e-9)


def test_parity_mode_constants_F3(oracle, synth):
    """sampleRate = 0 -> empty mel bank -> C0 = 26*ln(1e-10)*sqrt(1/26)."""
    fp = oracle.fingerprint(synth.sweep_noise(1.0, seed=1), oracle.default_params(algo_sample_rate=0))
    assert np.allclose(fp.mfcc[:, 0], -117.40926320884498, rtol=1e-14)
    assert np.max(np.abs(fp.mfcc[:, 1:])) < 1e-10
    for k in ("spectral_centroid", "spectral_rolloff", "spectral_bandwidth", "spectral_slope", "zero_crossing_rate",
              "pitch_estimate", "pitch_confidence", "harmonic_ratio"):
        assert not fp.arrays[k].any(), k
    assert np.all(fp.inharmonicity_ratio == 1.0) and fp.loudness_range == 0.0
    assert fp.spectral_flatness.any() and fp.spectral_crest.any() and fp.spectral_flux.any()


def test_spectral_descriptors_against_numpy(oracle, synth):
    sr = 44100
    pcm = synth.sweep_noise(1.0, seed=3)
    fp = oracle.fingerprint(pcm, oracle.default_params(algo_sample_rate=sr))
    mag, _, _ = oracle.stft(pcm, 1024, 256)
    B = mag.shape[1]
    f = np.arange(B) * sr / (2.0 * (B - 1))  # spectral_centroid.go:59-65
    sm = mag.sum(1)
    cen = (mag * f).sum(1) / sm
    assert np.allclose(fp.spectral_centroid, cen, rtol=1e-11)
    bw = np.sqrt((((f[None, :] - cen[:, None]) ** 2) * mag).sum(1) / sm)  # spectral_bandwidth.go:22
    assert np.allclose(fp.spectral_bandwidth, bw, rtol=1e-10)
    p2 = mag ** 2
    cum = np.cumsum(p2, axis=1)
    ro = f[(cum >= 0.85 * cum[:, -1:]).argmax(1)]  # spectral_rolloff.go:19-55
    assert np.array_equal(fp.spectral_rolloff, ro)
    crest = mag.max(1) / np.sqrt(p2.sum(1) / B)  # spectral_crest.go:18
    assert np.allclose(fp.spectral_crest, crest, rtol=1e-11)
    flat = np.exp(np.log(mag).mean(1)) / mag.mean(1)  # spectral_flatness.go:31 (all bins > 1e-10 here)
    assert (mag > 1e-10).all() and np.allclose(fp.spectral_flatness, np.minimum(flat, 1.0), rtol=1e-10)
    x, y = np.log10(f[1:]), np.log10(mag[:, 1:])  # spectral_slope.go:24-67
    n = B - 1
    slope = (n * (x * y).sum(1) - x.sum() * y.sum(1)) / (n * (x * x).sum() - x.sum() ** 2)
    assert np.allclose(fp.spectral_slope, slope, rtol=1e-8)
    d = np.maximum(mag[1:] - mag[:-1], 0.0)  # spectral_flux.go:17-36
    assert np.allclose(fp.spectral_flux, np.sqrt((d * d).sum(1)), rtol=1e-11)
    split = B // 4  # extractors/speech.go:436-456
    low = p2[:, :split].sum(1) / p2.sum(1)
    assert np.allclose(fp.low_energy_ratio, low, rtol=1e-10) and np.allclose(fp.high_energy_ratio, 1 - low, rtol=1e-9)


def test_pure_tone_centroid_and_pitch(oracle):
    sr, f0 = 44100, 3000.0
    t = np.arange(sr) / sr
    fp = oracle.fingerprint(0.5 * np.sin(2 * np.pi * f0 * t), oracle.default_params(algo_sample_rate=sr))
    assert abs(np.median(fp.spectral_centroid) - f0) < 40.0  # leakage of the Hann main lobe only
    fp = oracle.fingerprint(0.5 * np.sin(2 * np.pi * 440.0 * t), oracle.default_params(algo_sample_rate=sr))
    voiced = fp.pitch_estimate[fp.pitch_estimate > 0]
    assert voiced.size > 0.9 * fp.pitch_estimate.size and abs(np.median(voiced) - 440.0) < 3.0


# ---------------------------------------------------------------- time domain (bit-exact restatements)

def seq_sum(x):
    """Left-to-right float64 sum (np.add.accumulate is sequential; np.sum is pairwise)."""
    return np.add.accumulate(x)[-1] if x.size else 0.0


def test_pre_emphasis_energy_zcr_bit_exact(oracle, synth):
    sr = 44100
    pcm = synth.sweep_noise(0.5, seed=4)
    fp = oracle.fingerprint(pcm, oracle.default_params(algo_sample_rate=sr))
    y = pcm.copy()
    y[1:] = pcm[1:] - 0.97 * pcm[:-1]  # filters/pre_emphasis.go:135-155
    T = (pcm.size - 1024) // 256 + 1
    ste = np.array([math.sqrt(seq_sum(y[t * 256:t * 256 + 1024] ** 2) / 1024) for t in range(T)])  # temporal/energy.go:25-50
    assert np.array_equal(fp.short_time_energy, ste)
    neg = y < 0
    zc = np.array([(neg[t * 256 + 1:t * 256 + 1024] != neg[t * 256:t * 256 + 1023]).sum() / (1024 / sr) for t in range(T)])
    assert np.array_equal(fp.zero_crossing_rate, zc)  # spectral/zero_crossing_rate.go:37-53
    ent = np.where(ste > 0, -ste * np.log(ste + 1e-10), 0.0)  # extractors/speech.go:429-434
    assert np.allclose(fp.energy_entropy, ent, rtol=1e-14)
    var = ((ste - ste.mean()) ** 2).sum() / (T - 1)  # temporal/energy.go:97-118
    assert fp.energy_variance == pytest.approx(var, rel=1e-10)


# ---------------------------------------------------------------- cross-correlation

def np_ncc(a, b, max_lag):
    # algorithms/stats/correlation.go:203-228,373-409,421-501 (sequential sums)
    def z(s):
        mean = seq_sum(s) / s.size
        sd = math.sqrt(seq_sum((s - mean) ** 2) / s.size)
        return s - mean if sd < 1e-10 else (s - mean) / sd
    za, zb = z(a), z(b)
    L = max(0, min(max_lag, a.size - 1, b.size - 1))
    out = []
    for lag in range(-L, L + 1):
        if lag >= 0:
            n = min(a.size, b.size - lag)
            x, y = za[:n], zb[lag:lag + n]
        else:
            n = min(a.size + lag, b.size)
            x, y = za[-lag:-lag + n], zb[:n]
        den = math.sqrt(seq_sum(x * x) * seq_sum(y * y))
        out.append(0.0 if den < 1e-10 else seq_sum(x * y) / den)
    return np.array(out), L


def test_xcorr_against_numpy_bit_exact(oracle):
    rng = np.random.default_rng(0)
    for na, nb, ml in ((200, 200, 50), (150, 230, 400), (90, 40, 10)):
        a, b = rng.standard_normal(na), rng.standard_normal(nb)
        c, s = oracle.xcorr(a, b, ml)
        ref, L = np_ncc(a, b, ml)
        assert s.actual_max_lag == L and np.array_equal(c, ref)
        best = int(np.argmax(np.abs(ref)))  # first maximum, strict '>' (correlation.go:535-541)
        assert (s.peak_index, s.peak_lag, s.peak_correlation) == (best, best - L, ref[best])


def test_shifted_copy_returns_the_exact_lag_and_sign(oracle):
    """c(lag) = sum q[i] * r[i + lag]: reference delayed by d  ->  peak at +d."""
    rng = np.random.default_rng(1)
    base = np.convolve(rng.standard_normal(3000), np.ones(20) / 20, mode="same")
    for d in (37, -112, 0):
        q = base[200:2200]
        r = base[200 - d:2200 - d]
        _, s = oracle.xcorr(q, r, 300)
        assert s.peak_lag == d and s.peak_correlation > 0.99


def test_peak_metrics_formulas(oracle):
    rng = np.random.default_rng(2)
    a = rng.standard_normal(400)
    b = np.roll(a, 9) + 0.3 * rng.standard_normal(400)
    c, s = oracle.xcorr(a, b, 60)
    pk = s.peak_index
    idx = np.arange(c.size)
    noise = c[np.abs(idx - pk) > 5]
    assert s.snr == pytest.approx(20 * math.log10(abs(c[pk]) / math.sqrt((noise ** 2).mean())), rel=1e-12)
    assert s.sharpness == pytest.approx(-(c[pk + 1] - 2 * c[pk] + c[pk - 1]), rel=1e-12)
    side = np.abs(c[np.abs(idx - pk) > 10]).max()
    assert s.peak_to_sidelobe == pytest.approx(20 * math.log10(abs(c[pk]) / side), rel=1e-12)
    others = np.where(idx != pk, np.abs(c), -1)
    assert s.second_peak == c[int(np.argmax(others))]
    assert s.overlap_length == min(400, 400 - s.peak_lag) and s.p_value == 0.01 and s.is_significant == 1


def test_confidence_quality_spot_values(oracle, synth):
    """stats/alignment.go:183-305 re-derived by hand for one case."""
    q, r = synth.aligned_pair(20.0, offset_seconds=1.0, seed=5)
    p = oracle.default_params(algo_sample_rate=44100)
    ea, eb = oracle.fingerprint(q, p).short_time_energy, oracle.fingerprint(r, p).short_time_energy
    _, xs, ar = oracle.align_xcorr(ea, eb, 1000, 256, 44100)
    P = abs(xs.peak_correlation)
    peak_score = P + (P - 0.6) * 0.5 if P >= 0.6 else P
    sharp = min(0.9, xs.sharpness * 8)
    side = min(0.8, xs.peak_to_sidelobe / 15) if 0 < xs.peak_to_sidelobe < math.inf else 0.0
    snr = min(0.7, xs.snr / 25) if xs.snr > 0 else 0.0
    ratio = abs(xs.second_peak) / P
    pen = (ratio - 0.7) * 0.25 if (xs.second_peak != 0 and ratio > 0.7) else 0.0
    bonus = 0.12 if P >= 0.75 else (0.08 if P >= 0.6 else 0.0)
    conf = min(0.95, max(0.0, 0.55 * peak_score + 0.22 * sharp + 0.12 * side + 0.06 * snr + 0.05 * 0.15 + bonus - pen))
    assert ar.confidence == pytest.approx(conf, rel=1e-12)
    assert ar.offset == xs.peak_lag * 256 and ar.offset_seconds == ar.offset / 44100
    assert ar.similarity == min(1.0, P) and ar.noise_level == pytest.approx(1 - xs.snr / 20)
    assert abs(xs.peak_lag - 44100 / 256) <= 1


# ---------------------------------------------------------------- DTW

def py_dtw(q, r, band, step=0):
    n, m = len(q), len(r)
    C = np.full((n + 1, m + 1), np.inf)
    C[0, 0] = 0
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            if band > 0 and abs(i - j) > band:
                continue
            ld = math.sqrt(sum((x - y) * (x - y) for x, y in zip(q[i - 1], r[j - 1])))
            v, h, d = C[i - 1, j], C[i, j - 1], C[i - 1, j - 1]
            C[i, j] = ld + (min(min(v, h), d) if step == 0 else min(v, h) if step == 1 else min(v + 1, min(h + 1, d)))
    i, j, path = n, m, []
    while i > 0 or j > 0:
        cost = C[i, j] - C[i - 1, j - 1] if (i > 0 and j > 0) else 0.0
        path.append((i - 1, j - 1, cost))
        if i == 0:
            j -= 1
        elif j == 0:
            i -= 1
        else:
            cs = (C[i - 1, j], C[i, j - 1], C[i - 1, j - 1])
            mi = 0
            for k in range(3):
                if cs[k] < cs[mi]:
                    mi = k
            i, j = (i - 1, j) if mi == 0 else (i, j - 1) if mi == 1 else (i - 1, j - 1)
    return path[::-1], C


@pytest.mark.parametrize("n,m,dim,band,step", [(30, 30, 1, -1, 0), (25, 40, 3, -1, 1), (40, 35, 2, 6, 2),
                                               (30, 50, 1, 5, 0), (1, 7, 1, -1, 0)])
def test_dtw_against_pure_python(oracle, n, m, dim, band, step):
    rng = np.random.default_rng(n + m)
    q = np.cumsum(rng.standard_normal((n, dim)), 0)
    r = np.cumsum(rng.standard_normal((m, dim)), 0)
    with np.errstate(invalid="ignore"):
        path, Cm = py_dtw(q.tolist(), r.tolist(), band, step)
    d = oracle.dtw(q, r, band=band, step=step, want_matrix=True)
    assert [p[0] for p in path] == d["path_query"].tolist() and [p[1] for p in path] == d["path_ref"].tolist()
    assert np.array_equal(np.array([p[2] for p in path]), d["path_cost"], equal_nan=True)
    assert np.array_equal(Cm[1:], d["cost_matrix"])
    tc = Cm[n, m]
    assert d["total_cost"] == tc and (d["distance"] == tc / len(path) or math.isnan(d["distance"]))


def test_dtw_identical_sequences_quirks(oracle):
    """Identical inputs: pure diagonal; the first path point's Cost is C[1][1]-C[0][0] = 0 (dtw.go:171-174)."""
    x = np.sin(np.arange(64) * 0.2)
    d = oracle.dtw(x, x)
    assert d["path_query"].tolist() == list(range(64)) == d["path_ref"].tolist()
    assert d["total_cost"] == 0.0 and not d["path_cost"].any()


# ---------------------------------------------------------------- compare

def test_colstats_and_cosine_against_numpy(oracle):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((500, 13)) + np.arange(13)
    y = rng.standard_normal((400, 13)) * 2
    st = oracle.colstats(x)
    assert np.allclose(st[:13], x.mean(0), rtol=1e-13) and np.allclose(st[13:], x.std(0, ddof=1), rtol=1e-12)
    v1 = np.concatenate([x.mean(0), x.std(0, ddof=1)])
    v2 = np.concatenate([y.mean(0), y.std(0, ddof=1)])
    cos = v1 @ v2 / (np.linalg.norm(v1) * np.linalg.norm(v2))
    assert oracle.colstats_cosine(x, y) == pytest.approx(cos, rel=1e-12)


def test_compare_confidence_steps(oracle, synth):
    p = oracle.default_params(algo_sample_rate=44100)
    fp = oracle.fingerprint(synth.sweep_noise(1.0, seed=6), p)
    f1, _k1 = oracle.cmp_features(fp, harmonic=False)
    r = oracle.compare(f1, f1, [0.35, 0.15, 0.3, 0.0, 0.0, 0.2, 0.0])
    # identical fingerprints: mfcc and spectral both 1 -> similarity 1 -> 0.5 + 0.3 + 0.1 + 2*0.05 (comparison.go:1011-1037)
    assert r.overall_similarity == pytest.approx(1.0, abs=1e-12) and r.n_features == 2
    assert r.confidence == pytest.approx(1.0)


# ---------------------------------------------------------------- committed fixtures

@pytest.mark.parametrize("name", ["c1_fixed_sr", "c1_parity", "c3_speech_40mel"])
def test_golden_fingerprint_fixtures(oracle, name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    kw = {k[3:]: int(g[k]) for k in g.files if k.startswith("kw_")}
    fp = oracle.fingerprint(g["pcm"], oracle.default_params(**kw))
    for k in g.files:
        if k.startswith("out_"):
            assert np.array_equal(fp.arrays[k[4:]], g[k]), k


def test_golden_alignment_fixture(oracle):
    g = np.load(os.path.join(GOLDEN, "c2_alignment.npz"))
    c, s = oracle.xcorr(g["ea"], g["eb"], int(g["max_lag"]))
    assert np.array_equal(c, g["corr"]) and s.peak_lag == int(g["peak_lag"])
    d = oracle.dtw(g["dq"], g["dr"], band=int(g["band"]))
    assert np.array_equal(d["path_query"], g["path_query"]) and np.array_equal(d["path_ref"], g["path_ref"])


# --------------------------------------------------------------------------------------
# music-extractor spectral additions: independent numpy restatements of
# spectral_contrast.go:26-187, chroma_stft.go:63-138 and bark_scale.go:36-128
# --------------------------------------------------------------------------------------

def _np_contrast(mag, sr, n_bands):
    nb = mag.shape[1]
    nyq = sr / 2.0
    lo, hi = math.log10(200.0), math.log10(nyq if nyq > 200.0 else 400.0)
    edges = []
    for i in range(n_bands + 1):
        f = 10.0 ** (lo + i * (hi - lo) / n_bands)
        edges.append(min(max(int(f * (nb - 1) / nyq), 0), nb - 1))
    for i in range(1, n_bands + 1):
        if edges[i] <= edges[i - 1]:
            edges[i] = edges[i - 1] + 1
    out = np.zeros((mag.shape[0], n_bands))
    for t in range(mag.shape[0]):
        for b in range(n_bands):
            s, e = edges[b], min(edges[b + 1], nb)
            if s >= e:
                continue
            p = np.sort(mag[t, s:e] ** 2)
            k = max(int(0.2 * p.size), 1)
            valley, peak = p[:k].sum() / k, p[-k:].sum() / k
            if valley <= 0:
                valley = 1e-10
            out[t, b] = 0.0 if peak <= 0 else 10.0 * math.log10(peak / valley)
    return out, edges


def _np_chroma(mag, sr, win):
    res = sr / win
    out = np.zeros((mag.shape[0], 12))
    for f in range(mag.shape[1]):
        fr = f * res
        if fr < 80.0 or fr > 8000.0:
            continue
        midi = 69.0 + 12.0 * math.log2(fr / 440.0)
        c = int(math.floor(abs(midi) + 0.5) * (1 if midi >= 0 else -1)) % 12  # Go's math.Round
        out[:, c] += mag[:, f] ** 2
    tot = out.sum(axis=1, keepdims=True)
    return np.where(tot > 1e-10, out / np.where(tot > 1e-10, tot, 1.0), out)


def _np_bark(mag, sr, n_filters, low, high):
    nb = mag.shape[1]
    fft = (nb - 1) * 2
    h2b = lambda hz: 26.81 * hz / (1960.0 + hz) - 0.53
    b2h = lambda b: 1960.0 * (b + 0.53) / (26.28 - b)
    pts = [h2b(low) + i * (h2b(high) - h2b(low)) / (n_filters + 1) for i in range(n_filters + 2)]
    bins = [min(int(math.floor((fft + 1.0) * b2h(b) / sr + 0.5)), fft // 2) for b in pts]
    bank = np.zeros((n_filters, nb))
    for m in range(1, n_filters + 1):
        l, c, r = bins[m - 1], bins[m], bins[m + 1]
        for k in range(max(l, 0), min(c, nb)):
            if c != l:
                bank[m - 1, k] = (k - l) / (c - l)
        for k in range(max(c, 0), min(r, nb)):
            if r != c:
                bank[m - 1, k] = (r - k) / (r - c)
    return (mag ** 2) @ bank.T


def test_music_spectral_against_numpy(oracle, synth):
    x = synth.sweep_noise(1.5, seed=31)
    mag, _, _ = oracle.stft(x, 1024, 256)
    contrast, chroma, bark = oracle.music_spectral(x, sample_rate=44100, n_bands=6, n_bark=24, bark_low=50.0,
                                                   bark_high=15000.0)
    ref_c, edges = _np_contrast(mag, 44100, 6)
    assert edges == [4, 10, 22, 48, 106, 233, 512]  # 200 Hz * 110.25^(i/6) in 43.07 Hz bins
    np.testing.assert_allclose(contrast, ref_c, rtol=1e-12, atol=1e-10)
    np.testing.assert_allclose(chroma, _np_chroma(mag, 44100, 1024), rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(bark, _np_bark(mag, 44100, 24, 50.0, 15000.0), rtol=1e-11, atol=1e-12)
    assert np.allclose(chroma.sum(axis=1), 1.0)


def test_music_spectral_known_answers(oracle):
    sr, n = 44100, 44100
    t = np.arange(n) / sr
    a440 = np.sin(2 * np.pi * 440.0 * t)
    contrast, chroma, bark = oracle.music_spectral(a440, sample_rate=sr)
    assert np.all(np.argmax(chroma, axis=1) == 9)          # A4 -> pitch class 9 (MIDI 69 mod 12)
    assert np.all(chroma[:, 9] > 0.5)                       # 43 Hz bins: the Hann main lobe also touches G# and A#
    assert np.all(contrast[:, 0] > 30.0)                    # the 200-438 Hz band holds the tone's skirt: peaky
    # silence: no energy -> valley floored at 1e-10, peak 0 -> contrast 0; chroma stays all-zero (not normalised)
    c0, h0, b0 = oracle.music_spectral(np.zeros(8192), sample_rate=sr)
    assert not c0.any() and not h0.any() and not b0.any()
    for bad in (dict(win=0), dict(hop=0), dict(sample_rate=0)):
        with pytest.raises(Exception):
            oracle.music_spectral(a440, **bad)


@pytest.mark.parametrize("win,hop", [(512, 160), (1024, 256), (256, 256), (256, 700)])
def test_stft_streamer_matches_the_reference_loop(oracle, win, hop):
    """STFTStreamer.ProcessChunk (analyzers/spectral.go:323-374): every frame equals the batch transform of
    the samples the reference's loop would have at the head of its buffer -- including hop > win, where an emptied buffer
    does not skip the rest of the hop."""
    from stream_model import frame_starts
    rng = np.random.default_rng(win + hop)
    x = rng.standard_normal(12000)
    chunks = [1, 100, win - 1, 700, 0, 3000, 2 * win, hop, 37]
    chunks.append(x.size - sum(chunks))
    starts, left = frame_starts(chunks, win, hop)
    st = oracle.stft_stream(win, hop)
    pos = 0
    for c, want in zip(chunks, starts):
        mag, ph, cx = st.process_chunk(x[pos:pos + c])
        pos += c
        assert mag.shape[0] == len(want)
        for k, s0 in enumerate(want):
            m1, p1, c1 = oracle.stft(x[s0:s0 + win], win, hop, phase=True, cplx=True)
            assert np.array_equal(mag[k], m1[0]) and np.array_equal(ph[k], p1[0]) and np.array_equal(cx[k], c1[0])
    assert st.buffered() == left
    st.close()


def test_stft_streamer_arguments(oracle, capi):
    with pytest.raises(capi.SonarError) as e:
        oracle.stft_stream(0, 10)
    assert "window size must be positive" in e.value.msg
    with pytest.raises(capi.SonarError) as e:
        oracle.stft_stream(512, 0)
    assert "hop size must be positive" in e.value.msg
    st = oracle.stft_stream(512, 128)
    mag, _, _ = st.process_chunk(np.zeros(0))
    assert mag.shape == (0, 257) and st.buffered() == 0
    n = capi.C.c_int64()
    rc = oracle.lib.sonar_stft_stream_process(st.h, capi._dp(np.zeros(2000)), 2000, None, None, None, 0, capi.C.byref(n))
    assert rc != 0 and b"frame capacity too small" in oracle.lib.sonar_last_error() and st.buffered() == 0
    st.close()


# ---------------------------------------------------------------- inputs shorter than one window

@pytest.mark.parametrize("n,win,hop", [(900, 1024, 256), (769, 1024, 256), (400, 512, 160)])
def test_incomplete_single_frame_is_skipped_not_read(oracle, n, win, hop):
    """analyzers/spectral.go:409: (n - W)/H + 1 with Go's truncating division is 1 for n in (W - H, W); the worker
    then skips the job because it ends behind the signal (:472-474) and the row stays zero.  The samples behind the
    buffer must never be read: a huge sentinel placed there would otherwise show up in the spectrum."""
    rng = np.random.default_rng(n)
    buf = np.full(n + 2048, 1e30)
    buf[:n] = 0.1 * rng.standard_normal(n)
    x = buf[:n]
    mag, ph, cx = oracle.stft(x, win, hop, phase=True, cplx=True)
    assert mag.shape == (1, win // 2 + 1) and not mag.any() and not ph.any() and not cx.any()
    sr = 16000 if win == 512 else 44100
    p = oracle.default_params(window_size=win, hop_size=hop, energy_frame=win, energy_hop=hop, algo_sample_rate=sr,
                              call_sample_rate=sr)
    fp = oracle.fingerprint(x, p)
    assert fp.mfcc.shape == (1, 13) and np.isfinite(fp.mfcc).all()
    assert fp.mfcc[0, 0] == pytest.approx(-117.40926320884498, rel=1e-12)  # every mel energy 0 -> ln(1e-10)
    for k in ("spectral_centroid", "spectral_rolloff", "spectral_bandwidth", "spectral_flatness", "spectral_crest"):
        assert fp.arrays[k].shape == (1,) and fp.arrays[k][0] == 0.0, k
    assert fp.short_time_energy.size == 0  # energy.go:26-28
    # ZCR over pre[0 : min(W, N)] (speech.go:351-357)
    pre = x - 0.97 * np.concatenate(([0.0], x[:-1]))
    crossings = int(np.sum((pre[:-1] >= 0) != (pre[1:] >= 0)))
    assert fp.zero_crossing_rate[0] == crossings / (n / sr)
    # pitch frames: (n - 1024)/512 + 1 truncates to 1 only for n in (512, 1024); DetectPitch rejects the short frame
    tp = 1 if 512 < n < 1024 else 0
    assert fp.pitch_estimate.size == tp
    if tp:
        assert fp.pitch_estimate[0] == 0.0 and fp.inharmonicity_ratio[0] == 1.0


def test_too_short_still_errors(oracle, capi):
    with pytest.raises(capi.SonarError) as e:
        oracle.stft(np.zeros(768), 1024, 256)  # (768 - 1024)/256 + 1 = 0
    assert "signal too short" in e.value.msg


# ---- speech-specific group: independent numpy restatement of the Go -------------------------------

def _speechy(seconds=6.0, sr=16000, seed=7):
    """Voiced bursts (140 Hz harmonic stack with a little noise) separated by digitally silent gaps (their energies tie
    at the 10th-percentile threshold, so they count as pauses): passes detectSpeech, has pauses longer than 100 ms and a
    pitch the detector tracks."""
    n = int(seconds * sr)
    t = np.arange(n) / sr
    f0 = 140.0 * (1.0 + 0.05 * np.sin(2 * np.pi * 0.5 * t))
    ph = 2 * np.pi * np.cumsum(f0) / sr
    x = np.sin(ph) + 0.5 * np.sin(2 * ph) + 0.3 * np.sin(3 * ph)
    gate = ((t % 1.5) < 1.0).astype(float)
    rng = np.random.default_rng(seed)
    return (0.3 * x + 2e-3 * rng.standard_normal(n)) * gate


def _speech_params(lib, sr=16000):
    return lib.default_params(window_size=512, hop_size=160, energy_frame=512, energy_hop=160, algo_sample_rate=sr,
                              call_sample_rate=sr, n_mel=40)


def test_speech_group_against_numpy(oracle):
    sr = 16000
    x = _speechy(sr=sr)
    p = _speech_params(oracle, sr)
    fp, sp = oracle.fingerprint_speech(x, p)
    plain = oracle.fingerprint(x, p)
    y = x - 0.97 * np.concatenate(([0.0], x[:-1]))  # pre_emphasis.go:184-190
    # detectSpeech (speech_analysis.go:113-207)
    zc = np.count_nonzero((y[:-1] >= 0) != (y[1:] >= 0)) / (y.size - 1)
    f = y[:1024]
    mc = max(max(np.dot(f[: 1024 - lag], f[lag:]) / (1024 - lag) for lag in range(20, 400)), 0.0) / np.mean(f * f)
    assert 0.01 <= zc <= 0.3 and np.sqrt(np.mean(y * y)) >= 1e-3 and mc > 0.1 and sp["is_speech"]
    nf = (x.size - 1024) // 512 + 1
    assert sp["voicing_probability"].size == nf == sp["spectral_tilt"].size
    # the voicing sweep is the detector's voicing on the same frames as the harmonic block (speech.go:529-549)
    assert np.array_equal(sp["voicing_probability"], fp.voicing_strength)
    assert np.array_equal(fp.voicing_strength, plain.voicing_strength)
    # ... and leaves its 20-frame history behind: only the first frames of the pitch track can differ (median of three)
    assert np.array_equal(fp.pitch_estimate[3:], plain.pitch_estimate[3:])
    # extractSpectralTilt (speech.go:551-584)
    tilt = np.zeros(nf)
    for i in range(nf):
        fr = y[i * 512: i * 512 + 1024]
        hi, lo = 0.0, 0.0
        for j in range(1, fr.size):
            d = fr[j] - fr[j - 1]
            hi += d * d
            lo += fr[j] * fr[j]
        tilt[i] = -10 * np.log10(hi / lo) if lo > 0 else 0.0
    np.testing.assert_allclose(sp["spectral_tilt"], tilt, rtol=1e-13, atol=1e-13)
    # extractPauseDurations / estimateSpeechRate (speech.go:586-655,779-797)
    e = fp.short_time_energy
    thr = np.sort(e)[e.size // 10]
    low = np.concatenate((e <= thr, [False]))
    pauses, start = [], None
    for i, v in enumerate(low):
        if v and start is None:
            start = i
        elif not v and start is not None:
            d = (i - start) * (160 / sr)
            if d > 0.1:
                pauses.append(d)
            start = None
    assert sp["n_pause"] == len(pauses) >= 2
    np.testing.assert_allclose(sp["pause_duration"], pauses, rtol=1e-15)
    assert sp["speech_rate"] == pytest.approx(4.0 * (1.0 - np.count_nonzero(e <= thr) / e.size), rel=1e-14)


def test_speech_group_not_speech_returns_empty(oracle, synth):
    """White noise: the zero-crossing rate is ~0.5 (> 0.3), detectSpeech says no -> empty arrays, rate 0 (speech.go:281-291)
    and the fingerprint is the one computed without the group."""
    x = 0.2 * np.random.default_rng(3).standard_normal(16000 * 2)
    p = _speech_params(oracle)
    fp, sp = oracle.fingerprint_speech(x, p)
    assert not sp["is_speech"] and sp["voicing_probability"].size == 0 and sp["n_pause"] == 0 and sp["speech_rate"] == 0.0
    plain = oracle.fingerprint(x, p)
    assert np.array_equal(fp.pitch_estimate, plain.pitch_estimate) and np.array_equal(fp.mfcc, plain.mfcc)


@pytest.mark.parametrize("n1,n2,sr,off", [(441000, 441000, 44100, 7.3), (441000, 400000, 44100, -3.21), (20000, 50000, 16000, 0.0),
                                          (30000, 30000, 44100, 0.1), (50000, 50000, 44100, 0.99999)])
def test_truncate_to_alignment_against_the_go_arithmetic(oracle, n1, n2, sr, off):
    """TruncateToAlignmentPCM (extractors/alignment.go:223-297) restated in Python."""
    o = int(np.floor(abs(off) * sr + 0.5))  # math.Round: half away from zero (the argument is non-negative)
    s1, s2 = (0, o) if off > 0 else ((o, 0) if off < 0 else (0, 0))
    common = min(n1 - s1, n2 - s2)
    pad = int(0.5 * sr)
    if common > 2 * pad:
        s1, s2, common = s1 + pad, s2 + pad, common - 2 * pad
    assert oracle.truncate_to_alignment(n1, n2, sr, off) == (s1, s2, common)


def test_truncate_to_alignment_errors(oracle, capi):
    with pytest.raises(capi.SonarError, match="offset too large: need to skip 88200 samples but pcm2 only has 1000"):
        oracle.truncate_to_alignment(50000, 1000, 44100, 2.0)
    with pytest.raises(capi.SonarError, match="offset too large: need to skip 44100 samples but pcm1 only has 44100"):
        oracle.truncate_to_alignment(44100, 90000, 44100, -1.0)


# ---------------------------------------------------------------- YIN + the detector's temporal post-pass

def py_pitch_track(pcm, sr, alpha=0.97):
    """Independent restatement, written from the Go source, of extractHarmonicFeatures (extractors/speech.go:462-509) over
    the pre-emphasised stream (speech.go:161, filters/pre_emphasis.go:135-155): PitchDetector.DetectPitch per 1024 / 512
    frame = preprocessFrame (pitch_detection.go:282-314) -> detectPitchYin (:349-420, sums sequential in j and tau) ->
    parabolicInterpolation (:743-764) -> postProcessResult (:767-790: octave correction against the median of the last
    five history entries, MinConfidence 0.5) -> updateTemporalTracking (:876-902: history of 20, median-of-three
    smoothing on the history that already includes the frame, exponential blend while the history holds two)."""
    y = pcm.copy()
    y[1:] = pcm[1:] - alpha * pcm[:-1]
    N, H, half = 1024, 512, 512
    hann = np.array([0.5 * (1.0 - math.cos(2.0 * math.pi * i / (N - 1))) for i in range(N)])
    T = (y.size - N) // H + 1
    pitch_out, conf_out = np.zeros(T), np.zeros(T)
    history, prev = [], 0.0
    py_pitch_track.corrected = py_pitch_track.smoothed = 0

    def median_nonzero(vals):
        f = sorted(v for v in vals if v > 0)
        if not f:
            return 0.0
        n = len(f)
        return (f[n // 2 - 1] + f[n // 2]) / 2.0 if n % 2 == 0 else f[n // 2]

    for t in range(T):
        fr = y[t * H:t * H + N]
        p = fr.copy()
        p[1:] = fr[1:] - 0.97 * fr[:-1]
        p = p * hann
        # d[tau] = sum_j (p[j] - p[j + tau])^2, accumulated left to right (np.add.accumulate along j is sequential)
        idx = np.arange(half)[None, :] + np.arange(half)[:, None]          # [tau, j] -> j + tau
        delta = p[None, :half] - p[idx]
        d = np.add.accumulate(delta * delta, axis=1)[:, -1]
        cm = np.ones(half)
        run = 0.0
        for tau in range(1, half):
            run += d[tau]
            cm[tau] = d[tau] / (run / tau) if run != 0.0 else (math.nan if d[tau] == 0.0 else math.inf)
        min_tau = -1
        for tau in range(1, half):
            if cm[tau] < 0.15 and tau + 1 < half and cm[tau] < cm[tau + 1]:
                min_tau = tau
                break
        pitch = conf = 0.0
        if min_tau > 0:
            period = float(min_tau)
            if 0 < min_tau < half - 1:
                y1, y2, y3 = cm[min_tau - 1], cm[min_tau], cm[min_tau + 1]
                a, b = (y1 - 2 * y2 + y3) / 2, (y3 - y1) / 2
                if a != 0:
                    period = min_tau + (-b / (2 * a))
            freq = sr / period
            if 80.0 <= freq <= 1000.0:
                pitch, conf = freq, 1.0 - cm[min_tau]
        # postProcessResult
        if pitch != 0.0 and history:
            recent = history[-5:]
            if len(recent) >= 3:
                med = median_nonzero(recent)
                for ratio in (0.5, 2.0, 1.0 / 3.0, 3.0):
                    expect = med * ratio
                    with np.errstate(divide="ignore", invalid="ignore"):
                        close = np.float64(abs(pitch - expect)) / np.float64(expect) < 0.1
                    if close:
                        if abs(pitch - med) > abs(expect - med):
                            pitch = expect
                            py_pitch_track.corrected += 1
                        break
        if conf < 0.5:
            pitch = conf = 0.0
        # updateTemporalTracking
        history.append(pitch)
        history = history[-20:]
        out = pitch
        if len(history) > 1:
            recent = history[-3:]
            out = median_nonzero(recent) if len(recent) >= 3 else 0.3 * pitch + (1 - 0.3) * prev
        prev = out
        py_pitch_track.smoothed += out != pitch
        pitch_out[t], conf_out[t] = out, conf
    return pitch_out, conf_out


def _voiced_test_signal(sr, seconds, seed):
    n = int(seconds * sr)
    t = np.arange(n) / sr
    f = 300.0 * (1.0 + 0.2 * np.sin(2 * np.pi * 0.9 * t))           # a gliding "voice"
    ph = 2 * np.pi * np.cumsum(f) / sr
    gate = (np.sin(2 * np.pi * 1.7 * t) > -0.4).astype(np.float64)  # voiced stretches and gaps
    rng = np.random.default_rng(seed)
    gate[int(0.42 * n):int(0.75 * n)] = 1.0                          # no gap around the octave jump below
    x = 0.4 * (np.sin(ph) + 0.05 * np.sin(2 * ph)) * gate + 1e-4 * rng.standard_normal(n)
    # an octave jump in the middle: exercises the octave correction against the history
    j0, j1 = int(0.55 * n), int(0.62 * n)
    x[j0:j1] = 0.4 * np.sin(2.0 * ph[j0:j1]) + 1e-4 * rng.standard_normal(j1 - j0)
    return x


@pytest.mark.parametrize("sr,seconds,seed", [(16000, 2.0, 1), (44100, 1.2, 2), (22050, 2.0, 5)])
def test_yin_and_pitch_tracking_against_python_restatement(oracle, sr, seconds, seed):
    """Raw YIN decisions, parabolic refinement, octave correction, confidence gate and the median-of-three smoothing of the
    oracle are bit-identical to a restatement written independently from pitch_detection.go / speech.go."""
    x = _voiced_test_signal(sr, seconds, seed)
    fp = oracle.fingerprint(x, oracle.default_params(algo_sample_rate=sr, call_sample_rate=sr))
    want_p, want_c = py_pitch_track(x, sr)
    assert fp.pitch_estimate.shape == want_p.shape
    voiced = want_c > 0
    assert 20 <= voiced.sum() < want_p.size, "the signal must hold voiced stretches and gaps"
    assert py_pitch_track.smoothed >= 20
    if sr == 22050:
        assert py_pitch_track.corrected >= 1, "this case must exercise the octave correction"
    assert np.array_equal(fp.pitch_confidence, want_c)
    assert np.array_equal(fp.pitch_estimate, want_p)
    assert np.array_equal(fp.voicing_strength, want_c)            # speech.go:484: Voicing = confidence after the gate
    assert np.array_equal(fp.harmonic_ratio, want_c * 10.0)       # :499
    assert np.array_equal(fp.inharmonicity_ratio, 1.0 - want_c)   # :500
    assert np.array_equal(fp.tonal_centroid, np.where(want_p > 0, want_p, 0.0))  # :503-505


# ---------------------------------------------------------------- temporal feature group + loudness range

def py_ste(y, frame, hop):
    """temporal/energy.go:25-50: RMS per frame, squares summed left to right."""
    if y.size < frame or hop <= 0 or frame <= 0:
        return np.zeros(0)
    T = (y.size - frame) // hop + 1
    return np.array([math.sqrt(seq_sum(y[t * hop:t * hop + frame] ** 2) / frame) for t in range(T)])


def py_loudness_range(y, sr):
    """temporal/energy.go:157-225: 400 ms windows, hop = a quarter, 'loudness units', 10th..95th percentile through
    calculatePercentileRange (whose log of a ratio of dB values returns 0 unless the 95th percentile is positive)."""
    if y.size == 0 or sr <= 0:
        return 0.0
    win = int(0.4 * sr)
    hop = win // 4 or 1
    e = py_ste(y, win, hop)
    if e.size == 0:
        return 0.0
    lv = sorted((-0.691 + 10.0 * math.log10(v * v)) if v > 0 else -70.0 for v in e)
    lo, hi = lv[int(0.10 * (len(lv) - 1))], lv[int(0.95 * (len(lv) - 1))]
    if lo <= 0.0:
        lo = 1e-10
    if hi <= 0.0:
        return 0.0
    return 20.0 * math.log10(hi / lo)


def py_temporal_group(pcm, sr, frame, hop, alpha=0.97):
    """extractTemporalFeatures (extractors/speech.go:370-408) and its helpers (:641-777), written from the Go source."""
    y = pcm.copy()
    y[1:] = pcm[1:] - alpha * pcm[:-1]
    rms = py_ste(y, frame, hop)
    out = {"rms": rms, "dynamic_range": py_loudness_range(y, sr)}
    thr = sorted(rms)[len(rms) // 10]                                # :641-668 (the bubble sort is a sort)
    out["silence_ratio"] = float((rms <= thr).sum()) / len(rms)
    out["peak_amplitude"] = float(np.max(np.abs(y)))
    out["average_amplitude"] = seq_sum(np.abs(y)) / y.size
    der = rms[1:] - rms[:-1]                                         # energy.go:122-133
    mean = seq_sum(der) / der.size                                   # :695-716
    sd = math.sqrt(seq_sum((der - mean) ** 2) / der.size)
    gate = mean + 2 * sd
    onsets = [i for i in range(1, der.size - 1) if der[i] > der[i - 1] and der[i] > der[i + 1] and der[i] > gate]
    out["onset_density"] = len(onsets) / (y.size / sr)
    frame_time = hop / sr
    attack = []
    for o in onsets:                                                 # :718-749
        peak, start = rms[o], o
        j = o - 1
        while j >= 0 and j > o - 10:
            if rms[j] < 0.1 * peak:
                start = j
                break
            j -= 1
        attack.append(min((o - start) * frame_time, 0.1))
    out["attack"] = np.array(attack)
    T = (y.size - 512) // 256 + 1                                    # :751-777
    out["envelope"] = np.array([math.sqrt(seq_sum(y[t * 256:t * 256 + 512] ** 2) / 512) for t in range(T)])
    return out


def test_temporal_group_against_python_restatement(oracle, capi):
    sr = 16000
    rng = np.random.default_rng(5)  # noise bursts with abrupt onsets over a quiet floor
    n = 6 * sr
    pcm = 0.02 * rng.standard_normal(n)
    t = 0
    while t < n:
        on, off = int(rng.uniform(0.05, 0.2) * sr), int(rng.uniform(0.2, 0.6) * sr)
        if t + on <= n:
            pcm[t:t + on] += rng.uniform(0.2, 0.8) * rng.standard_normal(on)
        t += on + off
    kw = dict(window_size=512, hop_size=160, energy_frame=512, energy_hop=160, algo_sample_rate=sr, call_sample_rate=sr,
              enable=capi.FP_ENABLE_MFCC | capi.FP_ENABLE_TEMPORAL)
    fp = oracle.fingerprint(pcm, oracle.default_params(**kw))
    want = py_temporal_group(pcm, sr, 512, 160)
    assert len(want["attack"]) >= 3, "the case must contain onsets"
    assert np.array_equal(fp.rms_energy, want["rms"])
    assert np.array_equal(fp.envelope_shape, want["envelope"])
    assert fp.n_attack_time == len(want["attack"])
    assert np.array_equal(fp.attack_time[:fp.n_attack_time], want["attack"])
    assert fp.scalars["silence_ratio"] == want["silence_ratio"]
    assert fp.scalars["peak_amplitude"] == want["peak_amplitude"]
    assert fp.scalars["average_amplitude"] == want["average_amplitude"]
    assert fp.scalars["onset_density"] == pytest.approx(want["onset_density"], rel=1e-15)
    assert fp.scalars["dynamic_range"] == pytest.approx(want["dynamic_range"], rel=1e-12, abs=1e-15)


@pytest.mark.parametrize("gain,sr", [(0.05, 44100), (40.0, 16000), (25.0, 44100)])
def test_loudness_range_against_python_restatement(oracle, synth, gain, sr):
    """Loud inputs reach positive 'loudness units' (the only way calculatePercentileRange returns a non-zero range)."""
    rng = np.random.default_rng(int(gain) + 1)
    n = int(4.0 * sr)
    x = gain * rng.standard_normal(n) * (0.2 + np.abs(np.sin(np.arange(n) * 2 * np.pi / (1.3 * sr))))
    fp = oracle.fingerprint(x, oracle.default_params(algo_sample_rate=sr, call_sample_rate=sr))
    y = x.copy()
    y[1:] = x[1:] - 0.97 * x[:-1]
    want = py_loudness_range(y, sr)
    assert (want > 0.0) == (gain > 1.0), "a quiet input has no positive loudness unit, a loud one has"
    assert fp.loudness_range == pytest.approx(want, rel=1e-12, abs=1e-15)


# ---------------------------------------------------------------- AlignmentAnalyzer.alignWithDTW scalars

def py_dtw_scalars(pq, pr, pc, distance, n, m, sr):
    """algorithms/stats/alignment.go:129-148 and its helpers :380-643, written from the Go source."""
    L = len(pq)

    def mean_cost():
        return seq_sum(np.asarray(pc, dtype=np.float64)) / L if L else 0.0

    def cost_consistency():
        if L <= 1:
            return 0.0
        w = max(min(5, L // 4), 2)
        sm = np.empty(L)
        for i in range(L):
            lo, hi = max(0, i - w // 2), min(L - 1, i + w // 2)
            sm[i] = seq_sum(np.asarray(pc[lo:hi + 1], dtype=np.float64)) / (hi - lo + 1)
        mu = seq_sum(sm) / L
        if mu <= 1e-10:
            return 1.0
        sd = math.sqrt(seq_sum((sm - mu) ** 2) / L)
        return 1.0 / (1.0 + sd / mu)

    def diagonal_bias():
        if L <= 1:
            return 1.0
        diag = sum(1 for i in range(1, L) if pq[i] - pq[i - 1] > 0 and pr[i] - pr[i - 1] > 0)
        return 1.0 / (1.0 + math.exp(-10.0 * (diag / (L - 1) - 0.3)))

    def changes():
        return sum(1 for i in range(2, L) if (pq[i] - pq[i - 1], pr[i] - pr[i - 1]) != (pq[i - 1] - pq[i - 2], pr[i - 1] - pr[i - 2]))

    def smoothness():
        return 1.0 if L <= 2 else max(0.0, 1.0 - changes() / (L - 1))

    def quality():
        if L == 0:
            return 0.0
        eff = min(1.0, max(float(n), float(m)) / L)
        return min(1.0, max(0.0, 0.3 * eff + 0.3 * diagonal_bias() + 0.2 * smoothness() + 0.2 * cost_consistency()))

    avg_len = (n + m) / 2.0
    nd = distance / avg_len
    similarity = min(1.0, max(0.0, 0.5 * (1.0 / (1.0 + nd)) + 0.3 * quality() + 0.2 * (1.0 / (1.0 + mean_cost()))))
    eff = min(1.0, max(float(n), float(m)) / L)
    confidence = min(1.0, max(0.0, 0.4 * math.exp(-nd * 2.0) + 0.25 * eff + 0.2 * cost_consistency() + 0.15 * diagonal_bias()))
    s = sum(int(b) - int(a) for a, b in zip(pq, pr))
    offset = int(s / L) if L else 0                                   # Go integer division truncates toward zero
    stability = 0.0 if L < 3 else max(0.0, 1.0 - changes() / (L - 1))
    return dict(similarity=similarity, confidence=confidence, offset=offset, offset_seconds=offset / sr,
                alignment_quality=quality(), stability=stability)


@pytest.mark.parametrize("n,m,band,seed", [(300, 280, 40, 1), (64, 64, 0, 2), (500, 430, 90, 3), (2, 2, 0, 4)])
def test_dtw_scalars_against_python_restatement(oracle, n, m, band, seed):
    rng = np.random.default_rng(seed)
    base = np.cumsum(rng.standard_normal(max(n, m) + 50))
    q = base[:n] + 0.05 * rng.standard_normal(n)
    r = np.interp(np.linspace(0, n - 1, m) + 6.0 * np.sin(np.linspace(0, 5, m)), np.arange(base.size), base)
    r = r + 0.05 * rng.standard_normal(m)
    res = oracle.dtw(q, r, band=band)
    ar = oracle.align_dtw_scalars(res, n, m, 16000)
    want = py_dtw_scalars(res["path_query"], res["path_ref"], res["path_cost"], res["distance"], n, m, 16000)
    assert ar.offset == want["offset"]
    assert ar.offset_seconds == want["offset_seconds"]
    for k in ("similarity", "confidence", "alignment_quality", "stability"):
        assert getattr(ar, k) == pytest.approx(want[k], rel=1e-12, abs=1e-15), k


# ---------------------------------------------------------------- FingerprintComparator.Compare

def py_compare(a, b, weights, ct_a=0, ct_b=0, content_filter=False, spectral=True, harmonic=True, temporal=False):
    """fingerprint/comparison.go:133-194 (Compare), :266-341 (calculateFeatureSimilarity), :345-402 (compareMFCC = cosine of
    the per-coefficient mean / sample-std vector), :646-770 (spectral / temporal / harmonic groups), :772-882 (helpers),
    :886-889 (overall = feature similarity), :1011-1037 (confidence), written from the Go source; gonum's stat.Mean /
    stat.Variance / floats.Dot / floats.Norm restated as their definitions (unbiased variance, Euclidean norm)."""
    def cos(u, v):
        u, v = np.asarray(u, dtype=np.float64), np.asarray(v, dtype=np.float64)
        if u.size != v.size or u.size == 0:
            return 0.0
        nu, nv = math.sqrt(float(u @ u)), math.sqrt(float(v @ v))
        return 0.0 if nu == 0 or nv == 0 else float(u @ v) / (nu * nv)

    def seqstats(s1, s2):
        if len(s1) == 0 or len(s2) == 0:
            return 0.0
        return cos([np.mean(s1), np.std(s1, ddof=1)], [np.mean(s2), np.std(s2, ddof=1)])

    def scalar(v1, v2):
        if v1 == 0 and v2 == 0:
            return 1.0
        mx = max(abs(v1), abs(v2))
        return 1.0 if mx == 0 else max(0.0, 1.0 - abs(v1 - v2) / mx)

    match = ct_a == ct_b
    if content_filter and not match:
        return dict(overall=0.0, feature=0.0, confidence=0.25, match=match, n=0, dist={})
    sims, ws, dist = [], [], {}
    w = dict(zip(("mfcc", "spectral", "chroma", "temporal", "speech", "harmonic", "energy"), weights))
    if len(a["mfcc"]) and len(b["mfcc"]):
        st = lambda m: np.concatenate([m.mean(0), m.std(0, ddof=1)])
        s = cos(st(a["mfcc"]), st(b["mfcc"]))
        sims.append(s); ws.append(w["mfcc"]); dist["mfcc"] = 1.0 - s
    if spectral:
        parts = [seqstats(a[k], b[k]) for k in ("spectral_centroid", "spectral_rolloff", "spectral_flux") if len(a[k]) and len(b[k])]
        s = float(np.mean(parts)) if parts else 0.0
        sims.append(s); ws.append(w["spectral"]); dist["spectral"] = 1.0 - s if parts else 1.0
    if temporal:
        parts = []
        if a["dynamic_range"] > 0 and b["dynamic_range"] > 0:
            parts.append(scalar(a["dynamic_range"], b["dynamic_range"]))
        parts.append(scalar(a["silence_ratio"], b["silence_ratio"]))
        if a["onset_density"] > 0 and b["onset_density"] > 0:
            parts.append(scalar(a["onset_density"], b["onset_density"]))
        if len(a["rms_energy"]) and len(b["rms_energy"]):
            parts.append(seqstats(a["rms_energy"], b["rms_energy"]))
        s = float(np.mean(parts))
        sims.append(s); ws.append(w["temporal"]); dist["temporal"] = 1.0 - s
    if harmonic:
        parts = [seqstats(a[k], b[k]) for k in ("harmonic_ratio", "pitch_estimate") if len(a[k]) and len(b[k])]
        s = float(np.mean(parts)) if parts else 0.0
        sims.append(s); ws.append(w["harmonic"]); dist["harmonic"] = 1.0 - s if parts else 1.0
    feat = float(np.dot(sims, ws) / np.sum(ws))
    conf = 0.5 + (0.3 if feat > 0.8 else 0.2 if feat > 0.6 else 0.0) + (0.1 if match else 0.0) + 0.05 * len(dist)
    return dict(overall=feat, feature=feat, confidence=max(0.0, min(1.0, conf)), match=match, n=len(dist), dist=dist)


def _cmp_inputs(fp):
    d = {k: fp.arrays[k] for k in ("mfcc", "spectral_centroid", "spectral_rolloff", "spectral_flux", "harmonic_ratio",
                                   "pitch_estimate")}
    if "rms_energy" in fp.arrays:
        d["rms_energy"] = fp.arrays["rms_energy"]
        for k in ("dynamic_range", "silence_ratio", "onset_density"):
            d[k] = fp.scalars[k]
    return d


@pytest.mark.parametrize("temporal,harmonic,ct_b,content_filter", [(False, True, 0, False), (True, True, 0, False),
                                                                   (True, False, 2, False), (False, True, 2, True)])
def test_compare_against_python_restatement(oracle, capi, synth, temporal, harmonic, ct_b, content_filter):
    sr = 16000
    kw = dict(window_size=512, hop_size=160, energy_frame=512, energy_hop=160, algo_sample_rate=sr, call_sample_rate=sr,
              enable=capi.FP_ENABLE_MFCC | (capi.FP_ENABLE_TEMPORAL if temporal else 0))
    rng = np.random.default_rng(11)
    xa = _voiced_test_signal(sr, 2.0, 1) + 0.05 * rng.standard_normal(2 * sr) * (np.arange(2 * sr) % 4000 < 900)
    xb = 0.7 * synth.speech_band_noise(2.5) + 0.3 * _voiced_test_signal(sr, 2.5, 7)
    fa, fb = oracle.fingerprint(xa, oracle.default_params(**kw)), oracle.fingerprint(xb, oracle.default_params(**kw))
    weights = [0.35, 0.25, 0.10, 0.20, 0.10, 0.10, 0.15]  # comparison.go:1092-1103 (default content type)
    ca, _ka = oracle.cmp_features(fa, content_type=0, harmonic=harmonic, temporal=temporal)
    cb, _kb = oracle.cmp_features(fb, content_type=ct_b, harmonic=harmonic, temporal=temporal)
    got = oracle.compare(ca, cb, weights, content_filter=content_filter)
    want = py_compare(_cmp_inputs(fa), _cmp_inputs(fb), weights, 0, ct_b, content_filter, True, harmonic, temporal)
    assert bool(got.content_type_match) == want["match"]
    assert got.confidence == pytest.approx(want["confidence"], abs=1e-12)
    assert got.overall_similarity == pytest.approx(want["overall"], rel=1e-10, abs=1e-12)
    if content_filter and not want["match"]:
        return
    assert 0.05 < want["overall"] < 0.999, "the two fingerprints must differ"
    assert got.feature_similarity == pytest.approx(want["feature"], rel=1e-10)
    assert got.n_features == want["n"]
    for k, v in want["dist"].items():
        assert getattr(got, "dist_" + k) == pytest.approx(v, rel=1e-9, abs=1e-12), k
