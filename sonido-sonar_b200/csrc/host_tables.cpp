// Host-side (CPU, float64) pieces of libsonar.so that are not kernels:
//   * table generators the kernels consume (windows, mel bin points, DCT, lifter,
//     FFT twiddles) — built exactly as the reference's constructors build them;
//   * size / layout arithmetic the Go shim needs before it can allocate outputs;
//   * O(1) scalar formulas applied to kernel results (confidence, quality, p-value,
//     DTW path statistics are O(path) integer/f64 scans of an already-finished path).
// No signal-sized arithmetic runs here.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>
#include <sstream>

#include "common.h"

namespace sonar {

namespace {
const double kInf = std::numeric_limits<double>::infinity();

int64_t go_int(double x) {  // Go int(float64) on amd64: CVTTSD2SI, NaN/overflow -> INT64_MIN
  if (!(x > -9.2e18 && x < 9.2e18)) return INT64_MIN;
  return (int64_t)x;
}

double bessel_i0(double x) {  // fingerprint/analyzers/windowing.go:373-390
  double sum = 1.0, term = 1.0;
  for (int k = 1; k < 50; k++) {
    const double h = x / (2.0 * (double)k);
    term *= h * h;
    sum += term;
    if (term < 1e-12) break;
  }
  return sum;
}
}  // namespace

// fingerprint/analyzers/windowing.go:246-371 (coefficients), :393-433 (power normalisation)
int host_window(int type, int n, bool symmetric, bool normalize, double beta, double alpha, double* w) {
  if (n <= 0) return set_error(SONAR_ERR_INVALID, "window size must be positive");
  const double den = symmetric ? (double)(n - 1) : (double)n;
  for (int i = 0; i < n; i++) {
    const double x = (double)i;
    switch (type) {
      case SONAR_WINDOW_HANN: w[i] = 0.5 * (1.0 - std::cos(2 * M_PI * x / den)); break;
      case SONAR_WINDOW_HAMMING: w[i] = 0.54 - 0.46 * std::cos(2 * M_PI * x / den); break;
      case SONAR_WINDOW_BLACKMAN: {
        const double a = 2 * M_PI * x / den;
        w[i] = 0.42 - 0.5 * std::cos(a) + 0.08 * std::cos(2 * a);
        break;
      }
      case SONAR_WINDOW_BLACKMAN_HARRIS: {
        const double a = 2 * M_PI * x / den;
        w[i] = 0.35875 - 0.48829 * std::cos(a) + 0.14128 * std::cos(2 * a) - 0.01168 * std::cos(3 * a);
        break;
      }
      case SONAR_WINDOW_KAISER: {
        const double a = 2.0 * x / den - 1.0;
        w[i] = bessel_i0(beta * std::sqrt(1 - a * a)) / bessel_i0(beta);
        break;
      }
      case SONAR_WINDOW_TUKEY: {
        const int taper = (int)go_int(alpha * (double)n / 2.0);
        if (i < taper)
          w[i] = 0.5 * (1 + std::cos(M_PI * x / (double)taper - M_PI));
        else if (i >= n - taper)
          w[i] = 0.5 * (1 + std::cos(M_PI * (double)(i - (n - taper)) / (double)taper));
        else
          w[i] = 1.0;
        break;
      }
      case SONAR_WINDOW_RECTANGULAR: w[i] = 1.0; break;
      case SONAR_WINDOW_BARTLETT:
        w[i] = (i <= n / 2) ? 2.0 * x / (double)(n - 1) : 2.0 - 2.0 * x / (double)(n - 1);
        break;
      case SONAR_WINDOW_WELCH: {
        const double a = (x - (double)(n - 1) / 2.0) / ((double)(n - 1) / 2.0);
        w[i] = 1.0 - a * a;
        break;
      }
      default: return set_error(SONAR_ERR_INVALID, "unsupported window type");
    }
  }
  if (normalize) {
    double energy = 0.0;
    for (int i = 0; i < n; i++) energy += w[i] * w[i];
    const double nf = 1.0 / std::sqrt(energy / (double)n);
    for (int i = 0; i < n; i++) w[i] *= nf;
  }
  return SONAR_OK;
}

void resolve_mfcc(const sonar_fp_params* p, int* n_mfcc, int* n_mel, double* high, double* lifter) {
  *n_mfcc = p->n_mfcc > 0 ? p->n_mfcc : 13;                                // mfcc.go:59
  *n_mel = p->n_mel > 0 ? p->n_mel : 26;                                   // mfcc.go:62
  *high = p->high_hz > 0 ? p->high_hz : (double)p->algo_sample_rate / 2.0;  // mfcc.go:65
  *lifter = p->lifter > 0 ? p->lifter : 22.0;                              // mfcc.go:68
}

// algorithms/spectral/mel_scale.go:35-56
bool host_mel_bins(int n_mel, int fft_size, int sr, double low, double high, std::vector<int64_t>& bins) {
  const double low_mel = 2595.0 * std::log10(1.0 + low / 700.0);
  const double high_mel = 2595.0 * std::log10(1.0 + high / 700.0);
  const double step = (high_mel - low_mel) / (double)(n_mel + 1);
  bins.resize(n_mel + 2);
  for (int i = 0; i < n_mel + 2; i++) {
    const double mel = low_mel + (double)i * step;
    const double hz = 700.0 * (std::pow(10.0, mel / 2595.0) - 1.0);
    int64_t b = go_int(std::floor(((double)fft_size + 1.0) * hz / (double)sr + 0.5));
    bins[i] = std::min<int64_t>(b, fft_size / 2);
  }
  for (int i = 0; i + 1 < n_mel + 2; i++) {
    if (bins[i + 1] < bins[i]) return false;                 // cannot happen for low <= high
    if (bins[i + 1] > bins[i] && bins[i] < 0) return false;  // reference would index a negative bin
  }
  return true;
}

int host_fp_sizes(const sonar_fp_params* p, int64_t n, sonar_fp_sizes_t* o) {
  if (!p || !o) return set_error(SONAR_ERR_INVALID, "nil argument");
  if (n <= 0) return set_error(SONAR_ERR_EMPTY, "empty signal");  // analyzers/spectral.go:387
  if (p->window_size <= 0) return set_error(SONAR_ERR_INVALID, "window size must be positive");
  if (p->hop_size <= 0) return set_error(SONAR_ERR_INVALID, "hop size must be positive");
  const int64_t T = (n - p->window_size) / p->hop_size + 1;  // :409
  if (T <= 0)
    return set_error(SONAR_ERR_TOO_SHORT, "signal too short for given window size and hop size");
  int n_mfcc, n_mel;
  double high, lifter;
  resolve_mfcc(p, &n_mfcc, &n_mel, &high, &lifter);
  o->n_frames = T;
  o->n_bins = p->window_size / 2 + 1;
  o->n_flux = T > 1 ? T - 1 : 0;
  o->n_energy_frames = (n < p->energy_frame || p->energy_hop <= 0 || p->energy_frame <= 0)
                           ? 0
                           : (n - p->energy_frame) / p->energy_hop + 1;  // temporal/energy.go:26-30
  const int64_t tp = (n - 1024) / 512 + 1;  // extractors/speech.go:468-470
  o->n_pitch_frames = tp > 0 ? tp : 0;
  o->n_mfcc = n_mfcc;
  o->n_envelope = n < 512 ? 0 : (n - 512) / 256 + 1;  // speech.go:751-761
  return SONAR_OK;
}

int host_fp_layout(const sonar_fp_params* p, int64_t n, sonar_fp_dev_layout_t* L) {
  sonar_fp_sizes_t sz;
  int rc = host_fp_sizes(p, n, &sz);
  if (rc) return rc;
  const int64_t T = sz.n_frames, Te = sz.n_energy_frames, Tp = sz.n_pitch_frames;
  int64_t off = 0;
  L->mfcc = off, off += T * sz.n_mfcc;
  L->spectral_centroid = off, off += T;
  L->spectral_rolloff = off, off += T;
  L->spectral_bandwidth = off, off += T;
  L->spectral_flatness = off, off += T;
  L->spectral_crest = off, off += T;
  L->spectral_slope = off, off += T;
  L->spectral_flux = off, off += T;  // T-1 used
  L->zero_crossing_rate = off, off += T;
  L->short_time_energy = off, off += Te;
  L->energy_entropy = off, off += Te;
  L->low_energy_ratio = off, off += Te;
  L->high_energy_ratio = off, off += Te;
  L->pitch_estimate = off, off += Tp;
  L->pitch_confidence = off, off += Tp;
  L->voicing_strength = off, off += Tp;
  L->harmonic_ratio = off, off += Tp;
  L->inharmonicity_ratio = off, off += Tp;
  L->tonal_centroid = off, off += Tp;
  L->scalars = off, off += 8;
  L->total = (off + 1) & ~(int64_t)1;
  return SONAR_OK;
}

// ---- plan -------------------------------------------------------------------

FpPlan::~FpPlan() {
  if (d_blob) cudaFree(d_blob);
}

int build_fp_plan(const sonar_fp_params* p, std::shared_ptr<FpPlan>* out) {
  auto plan = std::make_shared<FpPlan>();
  const int N = p->window_size;
  switch (N) {
    case 256: plan->R1 = 8, plan->R2 = 16; break;
    case 512: plan->R1 = 16, plan->R2 = 16; break;
    case 1024: plan->R1 = 16, plan->R2 = 32; break;
    case 2048: plan->R1 = 32, plan->R2 = 32; break;
    default:
      // any other length goes the way go-dsp sends it (fft.FFT: Bluestein for lengths that are not a power of two;
      // the remaining powers of two take the same float64 kernel with its radix-2 transform): spectral_exact.cu
      if (N < 8 || N > 2048)
        return set_error(SONAR_ERR_UNSUPPORTED,
                         "window size must be 256, 512, 1024, 2048 (fused kernels) or any length in [8, 2048] (float64 route)");
      plan->R1 = 1, plan->R2 = 1;
      plan->exact_only = true;
      break;
  }
  const int R1 = plan->R1, R2 = plan->R2, M = N / 2, B = M + 1;
  plan->N = N;
  plan->hop = p->hop_size;
  plan->B = B;
  plan->algo_sr = p->algo_sample_rate;
  double high, lifter;
  resolve_mfcc(p, &plan->n_mfcc, &plan->n_mel, &high, &lifter);
  if (plan->n_mel > kMaxMel || plan->n_mfcc > kMaxMfcc)
    return set_error(SONAR_ERR_UNSUPPORTED, "at most 64 mel filters and 32 coefficients are supported");
  const int nm = plan->n_mel, nc = plan->n_mfcc;
  plan->n_regions = nm + 3;
  plan->split = B / 4;  // extractors/speech.go:442
  plan->freq_scale = (double)p->algo_sample_rate / (double)N;  // spectral_centroid.go:62
  plan->slope_on = p->algo_sample_rate > 0;

  // window pairs, folded with the 1/2 of the real-FFT split pass
  std::vector<double> w(N);
  // analyzers/spectral.go:415-420 builds WindowConfig{Type,Size,Normalize,Symmetric}: Beta = Alpha = 0
  int rc = host_window(p->window_type, N, true, true, 0.0, 0.0, w.data());
  if (rc) return rc;
  std::vector<float2> win2(M), tw1(M), wn(R1);
  for (int i = 0; i < M; i++) win2[i] = make_float2((float)(0.5 * w[2 * i]), (float)(0.5 * w[2 * i + 1]));
  for (int k1 = 0; k1 < R1; k1++)
    for (int n2 = 0; n2 < R2; n2++) {
      const double a = -2.0 * M_PI * (double)((int64_t)n2 * k1 % M) / (double)M;
      tw1[k1 * R2 + n2] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
  for (int k1 = 0; k1 < R1; k1++) {
    const double a = -2.0 * M_PI * (double)k1 / (double)N;
    wn[k1] = make_float2((float)std::cos(a), (float)std::sin(a));
  }
  // slope abscissae, centred (spectral_slope.go:42-51; the slope is shift invariant in x)
  std::vector<float> xtab(B + 1, 0.f);
  plan->slope_ntot = (double)(B - 1);
  plan->slope_xxtot = 0.0;
  if (plan->slope_on) {
    std::vector<double> x(B, 0.0);
    double mean = 0.0;
    for (int k = 1; k < B; k++) {
      x[k] = std::log10((double)k * (double)p->algo_sample_rate / (double)N);
      mean += x[k];
    }
    mean /= (double)(B - 1);
    for (int k = 1; k < B; k++) {
      xtab[k] = (float)(x[k] - mean);
      plan->slope_xxtot += (double)xtab[k] * (double)xtab[k];
    }
    // residual of the float32 rounding of the centred table: keep sum(x) consistent
    double sx = 0.0;
    for (int k = 1; k < B; k++) sx += (double)xtab[k];
    (void)sx;  // |sx| ~ 1e-6 * B; neglected against n*sxx (documented tolerance 1e-4)
  }
  // DCT-II (mfcc.go:194-212) and lifter (mfcc.go:230-245)
  std::vector<float> dct((size_t)nc * nm), lift(nc);
  for (int k = 0; k < nc; k++)
    for (int n = 0; n < nm; n++) {
      double v = std::cos(M_PI * (double)k * ((double)n + 0.5) / (double)nm);
      v *= (k == 0) ? std::sqrt(1.0 / (double)nm) : std::sqrt(2.0 / (double)nm);
      dct[(size_t)k * nm + n] = (float)v;
    }
  for (int i = 0; i < nc; i++)
    lift[i] = (i == 0 || !p->use_liftering) ? 1.f
                                            : (float)(1.0 + (lifter / 2.0) * std::sin(M_PI * (double)i / lifter));
  // mel regions
  std::vector<MelRegion> reg(nm + 3);
  std::vector<int64_t> bins;
  const bool ok = host_mel_bins(nm, (B - 1) * 2 /* mfcc.go:173-175 */, p->algo_sample_rate, p->low_hz, high, bins);
  if (!ok) return set_error(SONAR_ERR_UNSUPPORTED, "mel bin points are not representable (negative bins)");
  const bool empty_bank = bins[nm + 1] <= 0 || bins[nm + 1] == bins[0];
  for (auto& r : reg) r = MelRegion{INT_MAX, 0.f, 0.f, 0.f, 0.f};
  if (!empty_bank) {
    reg[0].next_b = (int)std::max<int64_t>(bins[0], 0);
    for (int s = 0; s <= nm; s++) {
      const int64_t lo = bins[s], hi = bins[s + 1];
      MelRegion r{(int)hi, (float)lo, (float)hi, 0.f, 0.f};
      if (hi > lo) {
        const float inv = (float)(1.0 / (double)(hi - lo));
        if (s >= 1) r.inv_f = inv;       // falling edge of filter s      (mel_scale.go:79-83)
        if (s <= nm - 1) r.inv_r = inv;  // rising edge of filter s+1     (mel_scale.go:72-76)
      }
      reg[1 + s] = r;
    }
    reg[nm + 2] = MelRegion{INT_MAX, 0.f, 0.f, 0.f, 0.f};
  }
  std::vector<int> chunk(R1);
  for (int j = 0; j < R1; j++) {
    const int k = j * R2;
    int r = 0;
    while (r < nm + 2 && k >= reg[r].next_b) r++;
    chunk[j] = r;
  }

  // pack the blob
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t r = off;
    off += (bytes + 255) & ~(size_t)255;
    return r;
  };
  plan->off_win2 = take(sizeof(float2) * M);
  plan->off_tw1 = take(sizeof(float2) * M);
  plan->off_wn = take(sizeof(float2) * R1);
  plan->off_xtab = take(sizeof(float) * (B + 1));
  plan->off_dct = take(sizeof(float) * dct.size());
  plan->off_lift = take(sizeof(float) * nc);
  plan->off_regions = take(sizeof(MelRegion) * reg.size());
  plan->off_chunk_region = take(sizeof(int) * R1);
  plan->off_hann = take(sizeof(double) * 1024);
  plan->off_zero = take(sizeof(double) * N);  // one frame of silence: the lone incomplete frame of an input shorter than the window
  plan->off_win64 = take(sizeof(double) * N);
  const bool pow2 = (N & (N - 1)) == 0;
  int MF = N;  // length of the radix-2 transforms of the float64 route
  if (!pow2) {
    MF = 1;
    while (MF < 2 * N - 1) MF <<= 1;
  }
  plan->fft_len = MF;
  plan->off_fac64 = take(sizeof(double2) * MF);
  plan->off_chirp = take(sizeof(double2) * N);
  plan->off_bluefb = take(sizeof(double2) * MF);
  plan->off_melbins = take(sizeof(int) * (nm + 2));
  plan->off_dct64 = take(sizeof(double) * (size_t)nc * nm);
  plan->off_lift64 = take(sizeof(double) * nc);
  plan->off_melinvw = take(sizeof(float) * kMaxMel);
  plan->blob_bytes = off;
  std::vector<unsigned char> host(off, 0);
  std::memcpy(host.data() + plan->off_win2, win2.data(), sizeof(float2) * M);
  std::memcpy(host.data() + plan->off_tw1, tw1.data(), sizeof(float2) * M);
  std::memcpy(host.data() + plan->off_wn, wn.data(), sizeof(float2) * R1);
  std::memcpy(host.data() + plan->off_xtab, xtab.data(), sizeof(float) * (B + 1));
  std::memcpy(host.data() + plan->off_dct, dct.data(), sizeof(float) * dct.size());
  std::memcpy(host.data() + plan->off_lift, lift.data(), sizeof(float) * nc);
  std::memcpy(host.data() + plan->off_regions, reg.data(), sizeof(MelRegion) * reg.size());
  plan->h_regions = reg;
  std::memcpy(host.data() + plan->off_chunk_region, chunk.data(), sizeof(int) * R1);
  {  // un-normalised symmetric Hann(1024) of the pitch detector (tonal/pitch_detection.go:318-324)
    double* h = reinterpret_cast<double*>(host.data() + plan->off_hann);
    for (int i = 0; i < 1024; i++) h[i] = 0.5 * (1.0 - std::cos(2.0 * M_PI * (double)i / 1023.0));
  }
  {  // float64 tables of the exact re-evaluation, built the way the reference's constructors build them
    std::memcpy(host.data() + plan->off_win64, w.data(), sizeof(double) * N);
    // go-dsp fft/radix2.go getRadix2Factors: the table of size i takes its even entries from the table of size i / 2
    // and evaluates the odd ones as sincos(-2 pi / i * k); the size-4 table is exact
    std::vector<double2> prev = {make_double2(1, 0), make_double2(0, -1), make_double2(-1, 0), make_double2(0, 1)};
    for (int i = 8; i <= MF; i <<= 1) {
      std::vector<double2> cur(i);
      for (int k = 0; k < i; k += 2) cur[k] = prev[k / 2];
      for (int k = 1; k < i; k += 2) {
        const double ang = -2 * M_PI / (double)i * (double)k;
        cur[k] = make_double2(std::cos(ang), std::sin(ang));
      }
      prev.swap(cur);
    }
    std::memcpy(host.data() + plan->off_fac64, prev.data(), sizeof(double2) * std::min<size_t>(prev.size(), (size_t)MF));
    if (!pow2) {
      // go-dsp fft/bluestein.go: w_i = exp(i pi i^2 / n) (math.Sincos; w_0 = 1 exactly), b[i] = b[MF - i] = w_i, and the
      // spectrum of b by the same radix-2 transform the kernel runs (bit reversal, then log2 MF butterfly stages)
      std::vector<double2> chirp(N), b(MF, make_double2(0, 0));
      double2* cinv = reinterpret_cast<double2*>(host.data() + plan->off_chirp);
      for (int i = 0; i < N; i++) {
        double sn = 0.0, cs = 1.0;
        if (i != 0) {
          const double ang = M_PI / (double)N * (double)((int64_t)i * i);
          sn = std::sin(ang);
          cs = std::cos(ang);
        }
        chirp[i] = make_double2(cs, sn);
        cinv[i] = make_double2(cs, -sn);
        b[i] = chirp[i];
        if (i != 0) b[MF - i] = chirp[i];
      }
      int bits = 0;
      while ((1 << bits) < MF) bits++;
      std::vector<double2> r(MF);
      for (int i = 0; i < MF; i++) {
        int rv = 0;
        for (int q = 0; q < bits; q++)
          if (i & (1 << q)) rv |= 1 << (bits - 1 - q);
        r[i] = b[rv];
      }
      for (int stage = 2; stage <= MF; stage <<= 1) {
        const int blocks = MF / stage, s2 = stage / 2;
        for (int nb = 0; nb < MF; nb += stage)
          for (int j = 0; j < s2; j++) {
            const int i1 = nb + j, i2 = i1 + s2;
            double2 t = r[i2];
            if (stage != 2) {
              const double2 f = prev[(size_t)blocks * j];
              t = make_double2(r[i2].x * f.x - r[i2].y * f.y, r[i2].x * f.y + r[i2].y * f.x);
            }
            const double2 a1 = r[i1];
            r[i1] = make_double2(a1.x + t.x, a1.y + t.y);
            r[i2] = make_double2(a1.x - t.x, a1.y - t.y);
          }
      }
      std::memcpy(host.data() + plan->off_bluefb, r.data(), sizeof(double2) * MF);
    }
    int* mb = reinterpret_cast<int*>(host.data() + plan->off_melbins);
    for (int i = 0; i < nm + 2; i++) mb[i] = (int)bins[i];
    double* d64 = reinterpret_cast<double*>(host.data() + plan->off_dct64);
    for (int k = 0; k < nc; k++)  // mfcc.go:194-212
      for (int n = 0; n < nm; n++) {
        double v = std::cos(M_PI * (double)k * ((double)n + 0.5) / (double)nm);
        if (k == 0)
          v *= std::sqrt(1.0 / (double)nm);
        else
          v *= std::sqrt(2.0 / (double)nm);
        d64[(size_t)k * nm + n] = v;
      }
    double* l64 = reinterpret_cast<double*>(host.data() + plan->off_lift64);
    for (int i = 0; i < nc; i++)  // mfcc.go:230-245; multiplying by 1.0 is exact
      l64[i] = (i == 0 || !p->use_liftering) ? 1.0 : 1.0 + (lifter / 2.0) * std::sin(M_PI * (double)i / lifter);
    float* iw = reinterpret_cast<float*>(host.data() + plan->off_melinvw);
    for (int f = 0; f < kMaxMel; f++) {  // a triangle's weights sum to (r - l) / 2
      const double wsum = (f < nm && !empty_bank) ? 0.5 * (double)(bins[f + 2] - bins[f]) : 0.0;
      iw[f] = wsum > 0.0 ? (float)(1.0 / wsum) : 0.f;
    }
  }
  SONAR_CUDA(cudaMalloc(&plan->d_blob, off));
  SONAR_CUDA(cudaMemcpy(plan->d_blob, host.data(), off, cudaMemcpyHostToDevice));
  *out = plan;
  return SONAR_OK;
}

// ---- music-extractor tables --------------------------------------

// spectral_contrast.go:131-187 initializeBands: log-spaced edges from 200 Hz to Nyquist, made strictly increasing
std::vector<int> host_contrast_edges(int n_bands, int num_bins, int sample_rate) {
  std::vector<int> edges((size_t)n_bands + 1);
  const double nyquist = (double)sample_rate / 2.0;
  const double min_freq = 200.0;
  double max_freq = nyquist;
  if (max_freq <= min_freq) max_freq = min_freq * 2;
  const double log_min = std::log10(min_freq), log_max = std::log10(max_freq);
  const double log_step = (log_max - log_min) / (double)n_bands;
  for (int i = 0; i <= n_bands; i++) {
    const double freq = std::pow(10.0, log_min + (double)i * log_step);
    int bin = (int)(freq * (double)(num_bins - 1) / nyquist);
    bin = std::min(bin, num_bins - 1);
    bin = std::max(bin, 0);
    edges[(size_t)i] = bin;
  }
  for (int i = 1; i <= n_bands; i++)
    if (edges[(size_t)i] <= edges[(size_t)i - 1]) edges[(size_t)i] = edges[(size_t)i - 1] + 1;
  return edges;
}

// chroma_stft.go:92-123: bins between 80 Hz and 8 kHz fold to round(69 + 12 log2(f / 440)) mod 12, others to -1
std::vector<signed char> host_chroma_map(int freq_bins, double freq_resolution) {
  std::vector<signed char> map((size_t)freq_bins);
  for (int f = 0; f < freq_bins; f++) {
    const double frequency = (double)f * freq_resolution;
    if (frequency < 80.0 || frequency > 8000.0) {
      map[(size_t)f] = -1;
      continue;
    }
    const double midi = 69.0 + 12.0 * std::log2(frequency / 440.0);
    map[(size_t)f] = (signed char)((int)std::round(midi) % 12);
  }
  return map;
}

// bark_scale.go:36-93 CreateBarkFilterBank (Traunmueller forward, the reference's own "inverse")
std::vector<double> host_bark_bank(int n_filters, int fft_size, int sample_rate, double low, double high) {
  auto hz2bark = [](double hz) { return (26.81 * hz / (1960.0 + hz)) - 0.53; };
  auto bark2hz = [](double b) { return 1960.0 * (b + 0.53) / (26.28 - b); };
  const int B = fft_size / 2 + 1;
  std::vector<double> bank((size_t)n_filters * B, 0.0);
  const double lo = hz2bark(low), hi = hz2bark(high);
  const double step = (hi - lo) / (double)(n_filters + 1);
  std::vector<int> bins((size_t)n_filters + 2);
  for (int i = 0; i < n_filters + 2; i++) {
    const double hz = bark2hz(lo + (double)i * step);
    const int b = (int)std::floor(((double)fft_size + 1.0) * hz / (double)sample_rate + 0.5);
    bins[(size_t)i] = std::min(b, fft_size / 2);
  }
  for (int m = 1; m <= n_filters; m++) {
    const int left = bins[(size_t)m - 1], center = bins[(size_t)m], right = bins[(size_t)m + 1];
    double* row = bank.data() + (size_t)(m - 1) * B;
    for (int k = std::max(left, 0); k < center && k < B; k++)
      if (center != left) row[k] = (double)(k - left) / (double)(center - left);
    for (int k = std::max(center, 0); k < right && k < B; k++)
      if (right != center) row[k] = (double)(right - k) / (double)(right - center);
  }
  return bank;
}

// ---- cross-correlation scalars ------------------------------------------------

int actual_max_lag(int max_lag, int64_t l1, int64_t l2) {  // stats/correlation.go:452-461
  int64_t m = max_lag;
  m = std::min<int64_t>(m, l1 - 1);
  m = std::min<int64_t>(m, l2 - 1);
  m = std::max<int64_t>(m, 0);
  return (int)m;
}

int64_t overlap_len(int64_t l1, int64_t l2, int64_t lag) {  // correlation.go:421-449,664-667
  if (lag >= 0) return std::min(l1, l2 - lag);
  return std::min(l1 + lag, l2);
}

void xcorr_derive(sonar_xcorr_summary* s, int64_t na, int64_t nb) {
  const int64_t n = std::min(na, nb);  // calculatePValue, correlation.go:547-569
  double pv = 1.0;
  if (n > 2) {
    const double c = s->peak_correlation;
    const double t = std::fabs(c) * std::sqrt((double)(n - 2)) / std::sqrt(1.0 - c * c);
    pv = t > 2.0 ? 0.01 : (t > 1.5 ? 0.05 : (t > 1.0 ? 0.1 : 0.5));
  }
  s->p_value = pv;
  s->is_significant = pv < (1.0 - 0.95);  // correlation.go:168
  s->overlap_length = (int32_t)overlap_len(na, nb, s->peak_lag);
}

double corr_confidence(const sonar_xcorr_summary* c) {  // stats/alignment.go:183-243
  const double pm = std::fabs(c->peak_correlation);
  if (pm < 0.1) return 0.0;
  const double peak_score = pm >= 0.6 ? pm + (pm - 0.6) * 0.5 : pm;
  const double sharp = std::fmin(0.9, c->sharpness * 8.0);
  double side = 0.0;
  if (c->peak_to_sidelobe > 0 && c->peak_to_sidelobe != kInf) side = std::fmin(0.8, c->peak_to_sidelobe / 15.0);
  double snr = 0.0;
  if (c->snr > 0) snr = std::fmin(0.7, c->snr / 25.0);
  double pen = 0.0;
  if (c->second_peak != 0 && pm > 0) {
    const double r = std::fabs(c->second_peak) / pm;
    if (r > 0.7) pen = (r - 0.7) * 0.25;
  }
  const double bonus = pm >= 0.75 ? 0.12 : (pm >= 0.6 ? 0.08 : 0.0);
  const double conf = 0.55 * peak_score + 0.22 * sharp + 0.12 * side + 0.06 * snr + 0.05 * 0.15 + bonus - pen;
  return std::fmin(0.95, std::fmax(0.0, conf));
}

double corr_quality(const sonar_xcorr_summary* c, int max_lag) {  // stats/alignment.go:245-305
  const double pm = std::fabs(c->peak_correlation);
  if (pm < 0.08) return 0.0;
  const double pq = pm >= 0.6 ? pm + (pm - 0.6) * 0.4 : pm;
  const double sharp = std::fmin(0.85, c->sharpness * 5.0);
  double side = 0.0;
  if (c->peak_to_sidelobe > 0 && c->peak_to_sidelobe != kInf) side = std::fmin(0.7, c->peak_to_sidelobe / 20.0);
  double snr = 0.0;
  if (c->snr > 0) snr = std::fmin(0.6, c->snr / 30.0);
  double lagpen = 0.0;
  if (max_lag > 0 && c->peak_lag < 0) {
    const double r = std::fabs((double)c->peak_lag) / (double)max_lag;
    if (r > 0.90) lagpen = (r - 0.90) * 4.0;
  }
  const double bonus = pm >= 0.7 ? 0.10 : (pm >= 0.55 ? 0.06 : 0.0);
  const double q = 0.50 * pq + 0.25 * sharp + 0.15 * side + 0.10 * snr + bonus - lagpen;
  return std::fmin(1.0, std::fmax(0.0, q));
}

// ---- DTW path statistics (stats/alignment.go:129-148,380-643) ----------------

namespace {
struct Path {
  const int32_t* q;
  const int32_t* r;
  const double* c;
  int64_t len;
};

double cost_consistency(const Path& p) {  // :452-500
  if (p.len <= 1) return 0.0;
  int64_t ws = std::max<int64_t>(std::min<int64_t>(5, p.len / 4), 2);
  std::vector<double> sm(p.len);
  for (int64_t i = 0; i < p.len; i++) {
    double sum = 0.0;
    int cnt = 0;
    const int64_t lo = std::max<int64_t>(0, i - ws / 2), hi = std::min<int64_t>(p.len - 1, i + ws / 2);
    for (int64_t j = lo; j <= hi; j++) {
      sum += p.c[j];
      cnt++;
    }
    sm[i] = sum / (double)cnt;
  }
  double mean = 0.0;
  for (double v : sm) mean += v;
  mean /= (double)sm.size();
  if (mean <= 1e-10) return 1.0;
  double var = 0.0;
  for (double v : sm) var += (v - mean) * (v - mean);
  var /= (double)sm.size();
  return 1.0 / (1.0 + std::sqrt(var) / mean);
}

double diagonal_bias(const Path& p) {  // :502-529
  if (p.len <= 1) return 1.0;
  int64_t diag = 0;
  for (int64_t i = 1; i < p.len; i++)
    if (p.q[i] - p.q[i - 1] > 0 && p.r[i] - p.r[i - 1] > 0) diag++;
  const double ratio = (double)diag / (double)(p.len - 1);
  return 1.0 / (1.0 + std::exp(-10.0 * (ratio - 0.3)));
}

int64_t direction_changes(const Path& p) {
  int64_t changes = 0;
  int pdq = 0, pdr = 0;
  for (int64_t i = 1; i < p.len; i++) {
    const int dq = p.q[i] - p.q[i - 1], dr = p.r[i] - p.r[i - 1];
    if (i > 1 && (dq != pdq || dr != pdr)) changes++;
    pdq = dq;
    pdr = dr;
  }
  return changes;
}

double smoothness(const Path& p) {  // :568-601
  if (p.len <= 2) return 1.0;
  return std::fmax(0.0, 1.0 - (double)direction_changes(p) / (double)(p.len - 1));
}

double quality(const Path& p, int n, int m) {  // :543-566
  if (p.len == 0) return 0.0;
  const double eff = std::fmin(1.0, std::fmax((double)n, (double)m) / (double)p.len);
  const double q = 0.3 * eff + 0.3 * diagonal_bias(p) + 0.2 * smoothness(p) + 0.2 * cost_consistency(p);
  return std::fmin(1.0, std::fmax(0.0, q));
}
}  // namespace

int host_align_dtw_scalars(const sonar_dtw_out* d, int n, int m, int sr, sonar_align_result* o) {
  if (!d || !o) return set_error(SONAR_ERR_INVALID, "nil argument");
  std::memset(o, 0, sizeof(*o));
  const Path p{d->path_query, d->path_ref, d->path_cost, d->path_len};
  o->method = 0;
  o->query_length = n;
  o->reference_length = m;
  o->sample_rate = sr;
  const double avg = (double)(n + m) / 2.0;
  if (avg != 0) {  // calculateSimilarityFromDTW :380-405
    const double nd = d->distance / avg;
    double mean_cost = 0.0;
    if (p.len > 0) {
      double tc = 0.0;
      for (int64_t i = 0; i < p.len; i++) tc += p.c[i];
      mean_cost = tc / (double)p.len;
    }
    const double sim = 0.5 * (1.0 / (1.0 + nd)) + 0.3 * quality(p, n, m) + 0.2 * (1.0 / (1.0 + mean_cost));
    o->similarity = std::fmin(1.0, std::fmax(0.0, sim));
    if (p.len > 0) {  // calculateDTWConfidence :420-450
      const double pe = std::fmin(1.0, std::fmax((double)n, (double)m) / (double)p.len);
      const double conf =
          0.4 * std::exp(-nd * 2.0) + 0.25 * pe + 0.2 * cost_consistency(p) + 0.15 * diagonal_bias(p);
      o->confidence = std::fmin(1.0, std::fmax(0.0, conf));
    }
  }
  int64_t so = 0;  // calculateAverageOffset :531-541
  for (int64_t i = 0; i < p.len; i++) so += (int64_t)p.r[i] - (int64_t)p.q[i];
  o->offset = p.len ? (int32_t)(so / p.len) : 0;
  o->offset_seconds = (double)o->offset / (double)sr;
  o->alignment_quality = quality(p, n, m);
  o->stability = p.len < 3 ? 0.0 : std::fmax(0.0, 1.0 - (double)direction_changes(p) / (double)(p.len - 1));
  return SONAR_OK;
}

}  // namespace sonar
