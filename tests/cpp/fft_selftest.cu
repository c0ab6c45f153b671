// Host-side self test of the in-register FFT building blocks (no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../../sonido-sonar_b200/csrc/fft_regs.cuh"
using namespace sonar;

template <int R>
double check() {
  float2 v[R];
  double re[R], im[R];
  for (int i = 0; i < R; i++) {
    re[i] = (double)rand() / RAND_MAX - 0.5;
    im[i] = (double)rand() / RAND_MAX - 0.5;
    v[i] = make_float2((float)re[i], (float)im[i]);
  }
  FftReg<R>::run(v);
  double err = 0;
  for (int k = 0; k < R; k++) {
    double sr = 0, si = 0;
    for (int n = 0; n < R; n++) {
      double a = -2 * M_PI * k * n / R;
      sr += re[n] * cos(a) - im[n] * sin(a);
      si += re[n] * sin(a) + im[n] * cos(a);
    }
    err = fmax(err, fmax(fabs(sr - v[k].x), fabs(si - v[k].y)));
  }
  return err;
}
int main() {
  double e[6] = {check<2>(), check<4>(), check<8>(), check<16>(), check<32>(), 0};
  int bad = 0;
  for (int i = 0; i < 5; i++) {
    printf("R=%d max err %.3e\n", 2 << i, e[i]);
    if (!(e[i] < 5e-6)) bad = 1;
  }
  for (int k = 0; k < 64; k++) {
    float2 w = w64(k);
    if (fabs(w.x - cos(-2 * M_PI * k / 64)) > 1e-7 || fabs(w.y - sin(-2 * M_PI * k / 64)) > 1e-7) bad = 1;
  }
  puts(bad ? "FAIL" : "OK");
  return bad;
}
