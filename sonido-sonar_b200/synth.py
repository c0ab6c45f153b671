"""Seeded synthetic PCM for the BASELINE.json configurations.

All signals are float64 mono in roughly [-1, 1].  numpy's Philox bit generator
is counter based, so the same (seed, length) gives identical samples on every
host; the device-side generator in csrc/ (sonar_synth_*) is only used to fill
HBM for the device-resident bench leg and is not parity-relevant.
"""
from __future__ import annotations

import numpy as np


def _normal(seed: int, n: int) -> np.ndarray:
    return np.random.Generator(np.random.Philox(seed)).standard_normal(n)


def sweep_noise(seconds: float, sr: int = 44100, seed: int = 1, f0: float = 100.0,
                f1: float = 8000.0, amp: float = 0.6, noise: float = 0.1) -> np.ndarray:
    """C1 / C4: linear sine sweep f0->f1 plus white noise."""
    n = int(round(seconds * sr))
    t = np.arange(n, dtype=np.float64) / sr
    dur = n / sr
    phase = f0 * t + 0.5 * (f1 - f0) * t * t / dur
    return amp * np.sin(2 * np.pi * phase) + noise * _normal(seed, n)


def envelope_noise(n: int, sr: int = 44100, seed: int = 2, start: int = 0) -> np.ndarray:
    """C2 / C5 global process S[n] evaluated on [start, start+n): slow envelope x noise."""
    idx = np.arange(start, start + n, dtype=np.float64)
    t = idx / sr
    env = 0.3 + 0.25 * np.sin(2 * np.pi * 0.37 * t) + 0.2 * np.sin(2 * np.pi * 1.9 * t)
    g = _normal(seed, start + n)[start:]
    return env * g


def aligned_pair(seconds: float, offset_seconds: float = 7.3, sr: int = 44100, seed: int = 2,
                 cdn_noise: float = 0.02):
    """C2: source (query) = S[n + offset], CDN (reference) = S[n] + noise.

    With the reference's convention c(lag) = sum q[i] * r[i + lag]
    (algorithms/stats/correlation.go:421-433) the peak sits at lag = +offset.
    """
    n = int(round(seconds * sr))
    off = int(round(offset_seconds * sr))
    base = envelope_noise(n + abs(off), sr, seed)
    if off >= 0:
        query, ref = base[off:off + n].copy(), base[:n].copy()
    else:
        query, ref = base[:n].copy(), base[-off:-off + n].copy()
    ref = ref + cdn_noise * _normal(seed + 1, n)
    return query, ref


def speech_band_noise(seconds: float, sr: int = 16000, seed: int = 4) -> np.ndarray:
    """C3: noise through a 2-pole band-pass (300-3400 Hz) gated by a 4 Hz syllabic envelope."""
    from scipy.signal import butter, lfilter

    n = int(round(seconds * sr))
    g = _normal(seed, n)
    b, a = butter(1, [300.0 / (sr / 2), 3400.0 / (sr / 2)], btype="band")
    y = lfilter(b, a, g)
    t = np.arange(n, dtype=np.float64) / sr
    gate = 0.5 * (1.0 + np.sin(2 * np.pi * 4.0 * t))
    y = y * (0.1 + 0.9 * gate)
    return 0.5 * y / np.max(np.abs(y))
