"""The tensor-pipe question for the STFT, answered on the numerics before any kernel is written:
a 1024-point transform as two stages of 32-point DFTs done as matrix products (the only shape the tensor pipe takes),
with the operands rounded to TF32 (10-bit mantissa), plain and with the 3-product split (hi x hi + hi x lo + lo x hi)
that recovers ~FP32 accuracy.  Reports the worst relative error of |X_k| against float64, relative to the frame's
strongest bin (what the 1e-4 feature tolerance is measured against for the weak bins that dominate flatness / slope),
and the tensor time the step would need at the B200's dense TF32 peak.
usage: python scripts/proto/tf32_fft_experiment.py"""
import numpy as np


def tf32(x):
    """round-to-nearest-even to a 10-bit mantissa (TF32), float32 exponent range"""
    x = np.asarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x0FFF + ((u >> 13) & 1)) & ~np.uint64(0x1FFF)
    return u.astype(np.uint32).view(np.float32)


def mm(a, b, split):
    """a @ b with TF32 operands and FP32 accumulation; split = 1 or 3 products"""
    a, b = a.astype(np.float32), b.astype(np.float32)
    ah, bh = tf32(a), tf32(b)
    out = ah.astype(np.float64) @ bh.astype(np.float64)
    if split == 3:
        al, bl = tf32(a - ah), tf32(b - bh)
        out = out + ah.astype(np.float64) @ bl.astype(np.float64) + al.astype(np.float64) @ bh.astype(np.float64)
    return out.astype(np.float32)


def cmm(wr, wi, xr, xi, split):
    return mm(wr, xr, split) - mm(wi, xi, split), mm(wr, xi, split) + mm(wi, xr, split)


def fft1024_gemm(x, split):
    """x: [frames, 1024] real (windowed).  n = 32 n1 + n2, k = k1 + 32 k2."""
    F = x.shape[0]
    n = np.arange(32)
    w32 = np.exp(-2j * np.pi * np.outer(n, n) / 32)
    tw = np.exp(-2j * np.pi * np.outer(n, n) / 1024)  # [k1, n2]
    X = x.reshape(F, 32, 32)  # [f, n1, n2]
    yr, yi = cmm(w32.real, w32.imag, X.transpose(1, 0, 2).reshape(32, -1), np.zeros((32, F * 32), np.float32), split)
    y = (yr + 1j * yi).reshape(32, F, 32).transpose(1, 0, 2)  # [f, k1, n2]
    y = (y * tw[None]).astype(np.complex64)  # FP32 twiddle
    zr, zi = cmm(w32.real, w32.imag, y.real.transpose(2, 0, 1).reshape(32, -1), y.imag.transpose(2, 0, 1).reshape(32, -1), split)
    z = (zr + 1j * zi).reshape(32, F, 32).transpose(1, 2, 0)  # [f, k1, k2]
    return z.transpose(0, 2, 1).reshape(F, 1024)  # k = k1 + 32 k2


rng = np.random.default_rng(0)
win = np.hanning(1024)
t = np.arange(1024) / 44100
cases = {"broadband noise": rng.standard_normal((64, 1024)),
         "tone 74 dB over its floor": 0.5 * np.sin(2 * np.pi * 440 * t)[None] + 1e-4 * rng.standard_normal((64, 1024))}
for name, x in cases.items():
    xw = x * win
    ref = np.abs(np.fft.fft(xw, axis=1))[:, :513]
    f32 = np.abs(np.fft.fft(xw.astype(np.float32), axis=1))[:, :513]
    print(name)
    for label, got in (("FP32 FFT (numpy, float32 input)", f32), ("TF32 x 1", np.abs(fft1024_gemm(xw, 1))[:, :513]),
                       ("TF32 x 3", np.abs(fft1024_gemm(xw, 3))[:, :513])):
        err = np.abs(got - ref)
        print(f"  {label:34s} max |d|X|| / max|X| = {np.max(err / ref.max(axis=1, keepdims=True)):.2e}   "
              f"worst relative error of a bin = {np.max(err / np.maximum(ref, 1e-300)):.2e}")
frames = 64 * 51676
macs = 2 * (32 * 32 * 32 * 4)  # two stages of complex 32 x 32 x 32 per frame (real first stage: half, ignored)
for split in (1, 3):
    print(f"tensor time per 64 x 300 s step at the dense TF32 peak (1.1 PFLOP/s), {split} product(s): "
          f"{frames * macs * 2 * split / 1.1e15 * 1e3:.2f} ms (+ two operand transposes and the twiddle pass in FP32)")
