// CPU run of the screen's FFT passes (sonido-sonar_b200/csrc/xcorr_fft.cuh): the Stockham passes against a naive DFT,
// and the packed two-real-sequence correlation against direct sums.  Build: nvcc -O2 -o t xcorr_fft_selftest.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../sonido-sonar_b200/csrc/xcorr_fft.cuh"

using namespace sonar;

static void transform(std::vector<double2>& a, std::vector<double2>& b, int dir) {
  const int64_t N = (int64_t)a.size();
  int64_t n = N, s = 1;
  double2 *src = a.data(), *dst = b.data();
  while (n >= 4) {
    for (int64_t t = 0; t < N / 4; ++t) xs_radix4(src, dst, t, n, s, dir);
    std::swap(src, dst);
    n /= 4;
    s *= 4;
  }
  if (n == 2) {
    for (int64_t t = 0; t < N / 2; ++t) xs_radix2(src, dst, t, n, s, dir);
    std::swap(src, dst);
  }
  if (src != a.data()) a.swap(b);
}

int main() {
  int bad = 0;
  for (int lg : {2, 3, 6, 7, 10, 11}) {
    const int64_t N = (int64_t)1 << lg;
    std::vector<double2> x(N), w(N), ref(N);
    srand(lg);
    for (auto& v : x) v = make_double2(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
    for (int dir : {-1, 1}) {
      for (int64_t k = 0; k < N; ++k) {
        double re = 0, im = 0;
        for (int64_t n = 0; n < N; ++n) {
          const double2 t = xs_twiddle((k * n) % N, N, dir);
          re += x[n].x * t.x - x[n].y * t.y;
          im += x[n].x * t.y + x[n].y * t.x;
        }
        ref[k] = make_double2(re, im);
      }
      std::vector<double2> a = x;
      transform(a, w, dir);
      double err = 0;
      for (int64_t k = 0; k < N; ++k) err = fmax(err, fmax(fabs(a[k].x - ref[k].x), fabs(a[k].y - ref[k].y)));
      printf("N=%lld dir=%d max err %.3g\n", (long long)N, dir, err);
      if (!(err < 1e-10 * N)) bad++;
    }
  }
  // correlation of two real sequences through one complex transform
  const int64_t na = 700, nb = 650, aml = 300, N = 1024;
  std::vector<double> a(na), b(nb);
  for (auto& v : a) v = rand() / (double)RAND_MAX - 0.5;
  for (auto& v : b) v = rand() / (double)RAND_MAX - 0.5;
  std::vector<double2> z(N), w(N), c(N);
  for (int64_t n = 0; n < N; ++n) z[n] = make_double2(n < na ? a[n] : 0.0, n < nb ? b[n] : 0.0);
  transform(z, w, -1);
  for (int64_t k = 0; k < N; ++k) c[k] = xs_cross_spectrum(z[k], z[(N - k) & (N - 1)]);
  transform(c, w, +1);
  double err = 0;
  for (int64_t lag = -aml; lag <= aml; ++lag) {
    int64_t s1, s2, len;
    xs_overlap(lag, na, nb, &s1, &s2, &len);
    double sum = 0;
    for (int64_t i = 0; i < len; ++i) sum += a[s1 + i] * b[s2 + i];
    err = fmax(err, fabs(sum - c[lag & (N - 1)].x / (double)N));
  }
  printf("correlation max err %.3g\n", err);
  if (!(err < 1e-12)) bad++;
  printf(bad ? "FAIL\n" : "OK\n");
  return bad ? 1 : 0;
}
