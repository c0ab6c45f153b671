"""numpy emulation of the data flow of stft_v3.cu (index math only, float64): one warp transforms TWO frames
A (re) + i B (im) with a 32 x 32 complex FFT, lanes as an array axis.  Checked against numpy.fft.rfft.

N = 1024:  lane l holds c[j] = a_w[l + 32 j] + i b_w[l + 32 j], j = 0..31
  pass 1   C_l[k1] = sum_j c[j] W32^(j k1)                      (in registers)
  twiddle  R[k1][l] = W1024^(l k1) C_l[k1]                      -> exchange tile [k1][l]
  pass 2   lane m = k1:  Z[m + 32 k2] = sum_l R[m][l] W32^(l k2)
  split    k = m + 32 k2 (k2 < 16): partner Z[1024 - k] sits in lane (32 - m) & 31, register 31 - k2
           (lane 0: register (32 - k2) & 31);  XA = (Z + conj Zp)/2,  XB = (Z - conj Zp)/(2i)
N = 512: two packs (A + iB, C + iD) of 16 samples per lane; pass 1 = 2 x DFT-16; lane m: pack m >> 4, row k1 = m & 15;
         Z[k1 + 16 k2]; partner lane (m & 16) | ((16 - k1) & 15), register 31 - k2 (k1 == 0: (32 - k2) & 31).
"""
import numpy as np


def warp_fft_1024(a_w, b_w):
    l = np.arange(32)
    c = np.zeros((32, 32), complex)  # [lane][j]
    for j in range(32):
        c[:, j] = a_w[l + 32 * j] + 1j * b_w[l + 32 * j]
    C = np.fft.fft(c, axis=1)  # [lane][k1]
    R = np.zeros((32, 32), complex)  # tile [k1][l]
    for k1 in range(32):
        R[k1, :] = np.exp(-2j * np.pi * l * k1 / 1024) * C[:, k1]
    Z = np.fft.fft(R, axis=1)  # lane m = k1: Z[m][k2] = Z_full[m + 32 k2]
    XA = np.zeros(513, complex)
    XB = np.zeros(513, complex)
    for m in range(32):
        src = (32 - m) & 31
        for k2 in range(16):
            k = m + 32 * k2
            reg = (32 - k2) & 31 if m == 0 else 31 - k2
            zp = Z[src][reg]
            z = Z[m][k2]
            XA[k] = (z + np.conj(zp)) / 2
            XB[k] = (z - np.conj(zp)) / 2j
    z = Z[0][16]
    XA[512], XB[512] = z.real, z.imag
    return XA, XB


def warp_fft_512(frames_w):
    """frames_w: 4 windowed frames of 512 samples -> 4 spectra of 257 bins."""
    l = np.arange(32)
    R = np.zeros((32, 32), complex)  # tile row m = pack * 16 + k1
    for pack in range(2):
        a_w, b_w = frames_w[2 * pack], frames_w[2 * pack + 1]
        c = np.zeros((32, 16), complex)
        for j in range(16):
            c[:, j] = a_w[l + 32 * j] + 1j * b_w[l + 32 * j]
        C = np.fft.fft(c, axis=1)  # [lane][k1], k1 < 16
        for k1 in range(16):
            R[pack * 16 + k1, :] = np.exp(-2j * np.pi * l * k1 / 512) * C[:, k1]
    Z = np.fft.fft(R, axis=1)  # lane m: Z_pack[k1 + 16 k2]
    X = [np.zeros(257, complex) for _ in range(4)]
    for m in range(32):
        pack, k1 = m >> 4, m & 15
        src = (m & 16) | ((16 - k1) & 15)
        for k2 in range(16):
            k = k1 + 16 * k2
            reg = (32 - k2) & 31 if k1 == 0 else 31 - k2
            z, zp = Z[m][k2], Z[src][reg]
            X[2 * pack][k] = (z + np.conj(zp)) / 2
            X[2 * pack + 1][k] = (z - np.conj(zp)) / 2j
        if k1 == 0:
            z = Z[m][16]
            X[2 * pack][256], X[2 * pack + 1][256] = z.real, z.imag
    return X


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    w = np.hanning(1024)
    a, b = rng.standard_normal(1024) * w, rng.standard_normal(1024) * w
    XA, XB = warp_fft_1024(a, b)
    print("1024:", np.max(np.abs(XA - np.fft.rfft(a))), np.max(np.abs(XB - np.fft.rfft(b))))
    fr = [rng.standard_normal(512) for _ in range(4)]
    X = warp_fft_512(fr)
    print("512:", max(np.max(np.abs(X[i] - np.fft.rfft(fr[i]))) for i in range(4)))
