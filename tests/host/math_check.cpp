// Host-compiled window onto csrc/posefit_math.h (the arithmetic the CUDA kernels run per
// object / per hypothesis), so the CPU test-suite can pin it to the oracle without a GPU.
// Not part of the product: built on the fly by tests/test_math_host.py with g++.
#include <cmath>
#include <cstring>
#include <vector>
#include "posefit_math.h"

using namespace posefit;

static void moments_of(const double* src, const double* dst, const int* idx, int n, Moments& m) {
  std::memset(&m, 0, sizeof(m));
  for (int k = 0; k < n; ++k) {
    const int i = idx ? idx[k] : k;
    const double* x = src + 3 * i;
    const double* y = dst + 3 * i;
    m.n += 1.0;
    for (int a = 0; a < 3; ++a) {
      m.sx[a] += x[a];
      m.sy[a] += y[a];
      m.sxx += x[a] * x[a];
      for (int b = 0; b < 3; ++b) m.syx[3 * a + b] += y[a] * x[b];
    }
  }
}

static void pack(const Fit& f, double* out) {
  out[0] = f.s;
  for (int i = 0; i < 9; ++i) out[1 + i] = f.R[i];
  for (int i = 0; i < 3; ++i) out[10 + i] = f.t[i];
  out[13] = f.var;
  for (int i = 0; i < 6; ++i) out[14 + i] = f.H[i];
  for (int i = 0; i < 6; ++i) out[20 + i] = f.Linv[i];
  out[26] = (double)f.status;
}

extern "C" {

// out[27]: s, R(9), t(3), var, H(6), Linv(6), status
void pf_check_fit(const double* src, const double* dst, int n, int precise, double* out) {
  Moments m;
  moments_of(src, dst, nullptr, n, m);
  Fit f;
  if (precise) fit_from_moments<true>(m, f); else fit_from_moments<false>(m, f);
  pack(f, out);
}

// residual^2 of every hypothesis (closed form) -- idx[n_hyp][n_samp] into the n points
void pf_check_hypotheses(const double* src, const double* dst, int n, const int* idx, int n_hyp, int n_samp,
                         int ref_compat, double* res2) {
  Moments all;
  moments_of(src, dst, nullptr, n, all);
  GlobalStats g;
  g.n = all.n;
  for (int a = 0; a < 3; ++a) { g.mux[a] = all.sx[a] / all.n; g.muy[a] = all.sy[a] / all.n; }
  g.Syy = 0.0;
  for (int a = 0; a < 9; ++a) g.Syx[a] = 0.0;
  for (int a = 0; a < 6; ++a) g.Sxx[a] = 0.0;
  for (int i = 0; i < n; ++i) {
    double x[3], y[3];
    for (int a = 0; a < 3; ++a) { x[a] = src[3 * i + a] - g.mux[a]; y[a] = dst[3 * i + a] - g.muy[a]; }
    g.Syy += y[0] * y[0] + y[1] * y[1] + y[2] * y[2];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) g.Syx[3 * a + b] += y[a] * x[b];
    g.Sxx[0] += x[0] * x[0]; g.Sxx[1] += x[0] * x[1]; g.Sxx[2] += x[0] * x[2];
    g.Sxx[3] += x[1] * x[1]; g.Sxx[4] += x[1] * x[2]; g.Sxx[5] += x[2] * x[2];
  }
  for (int h = 0; h < n_hyp; ++h) {
    Moments m;
    moments_of(src, dst, idx + (size_t)h * n_samp, n_samp, m);
    Fit f;
    fit_from_moments<false>(m, f);
    double A[9];
    scoring_transform(f, ref_compat != 0, A);
    res2[h] = residual_sq(g, A, f.t);
  }
}

// adjoint: out[16] = GC(9), gvar, gmux(3), gmuy(3)
void pf_check_adjoint(const double* src, const double* dst, int n, double gs, const double* gR, const double* gt,
                      double* out) {
  Moments m;
  moments_of(src, dst, nullptr, n, m);
  Fit f;
  fit_from_moments<true>(m, f);
  FitAdjoint a;
  fit_adjoint(f, gs, gR, gt, a);
  for (int i = 0; i < 9; ++i) out[i] = a.GC[i];
  out[9] = a.gvar;
  for (int i = 0; i < 3; ++i) { out[10 + i] = a.gmux[i]; out[13 + i] = a.gmuy[i]; }
}

// The float screen of the RANSAC kernel (posefit_math.h: screen_fit32 / screen_interval) next to the all-double
// evaluation, on crop-layout inputs exactly as fit_ransac_kernel<.., SCREEN> forms them: x = noc - 0.5f,
// y = ((float)rx z, -(float)ry z, -z).  rows / cols are FRAME coordinates, kinv4 = {k0, k2, k4, k5} of a pinhole K^-1.
// Outputs per hypothesis: r64 (double fit), lo / hi (interval around the float fit's residual).
void pf_check_screen(const float* noc, const float* z, const int* rows, const int* cols, int n, const double* kinv4,
                     const int* idx, int n_hyp, int n_samp, int ref_compat, double* r64, double* lo, double* hi) {
  std::vector<double> src(3 * (size_t)n), dst(3 * (size_t)n);
  std::vector<float> xs(3 * (size_t)n), ys(3 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    const double rx = kinv4[0] * (double)cols[i] + kinv4[1], ry = kinv4[2] * (double)rows[i] + kinv4[3], zd = z[i];
    for (int a = 0; a < 3; ++a) { src[3 * i + a] = (double)noc[3 * i + a] - 0.5; xs[3 * i + a] = noc[3 * i + a] - 0.5f; }
    dst[3 * i] = rx * zd; dst[3 * i + 1] = -(ry * zd); dst[3 * i + 2] = -zd;
    ys[3 * i] = (float)rx * z[i]; ys[3 * i + 1] = -((float)ry * z[i]); ys[3 * i + 2] = -z[i];
  }
  Moments all;
  moments_of(src.data(), dst.data(), nullptr, n, all);
  GlobalStats g;
  g.n = all.n;
  for (int a = 0; a < 3; ++a) { g.mux[a] = all.sx[a] / all.n; g.muy[a] = all.sy[a] / all.n; }
  g.Syy = 0.0;
  for (int a = 0; a < 9; ++a) g.Syx[a] = 0.0;
  for (int a = 0; a < 6; ++a) g.Sxx[a] = 0.0;
  double xraw = 0.0;
  for (int i = 0; i < n; ++i) {
    double x[3], y[3];
    for (int a = 0; a < 3; ++a) { x[a] = src[3 * i + a] - g.mux[a]; y[a] = dst[3 * i + a] - g.muy[a]; xraw += src[3 * i + a] * src[3 * i + a]; }
    g.Syy += y[0] * y[0] + y[1] * y[1] + y[2] * y[2];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) g.Syx[3 * a + b] += y[a] * x[b];
    g.Sxx[0] += x[0] * x[0]; g.Sxx[1] += x[0] * x[1]; g.Sxx[2] += x[0] * x[2];
    g.Sxx[3] += x[1] * x[1]; g.Sxx[4] += x[1] * x[2]; g.Sxx[5] += x[2] * x[2];
  }
  const double x_rms = sqrt(xraw / n);
  for (int h = 0; h < n_hyp; ++h) {
    const int* id = idx + (size_t)h * n_samp;
    // all-double, shifted by the first sample as the kernel does
    Moments mo;
    std::memset(&mo, 0, sizeof(mo));
    mo.n = n_samp;
    double ox[3], oy[3];
    for (int a = 0; a < 3; ++a) { ox[a] = src[3 * id[0] + a]; oy[a] = dst[3 * id[0] + a]; }
    for (int k = 0; k < n_samp; ++k) {
      double x[3], y[3];
      for (int a = 0; a < 3; ++a) { x[a] = src[3 * id[k] + a] - ox[a]; y[a] = dst[3 * id[k] + a] - oy[a]; }
      for (int a = 0; a < 3; ++a) {
        mo.sx[a] += x[a]; mo.sy[a] += y[a]; mo.sxx = fma(x[a], x[a], mo.sxx);
        for (int b = 0; b < 3; ++b) mo.syx[3 * a + b] = fma(y[a], x[b], mo.syx[3 * a + b]);
      }
    }
    Fit f;
    fit_from_moments<false>(mo, f, ox, oy);
    double A[9];
    scoring_transform(f, ref_compat != 0, A);
    r64[h] = f.status == PF_OK ? residual_sq(g, A, f.t) : NAN;
    // float screen
    float fox[3], foy[3], sx[3] = {0, 0, 0}, sy[3] = {0, 0, 0}, syx[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, sxx = 0, syy = 0;
    for (int a = 0; a < 3; ++a) { fox[a] = xs[3 * id[0] + a]; foy[a] = ys[3 * id[0] + a]; }
    for (int k = 1; k < n_samp; ++k) {
      float x[3], y[3];
      for (int a = 0; a < 3; ++a) { x[a] = xs[3 * id[k] + a] - fox[a]; y[a] = ys[3 * id[k] + a] - foy[a]; }
      for (int a = 0; a < 3; ++a) {
        sx[a] += x[a]; sy[a] += y[a]; sxx = fmaf(x[a], x[a], sxx); syy = fmaf(y[a], y[a], syy);
        for (int b = 0; b < 3; ++b) syx[3 * a + b] = fmaf(y[a], x[b], syx[3 * a + b]);
      }
    }
    ScreenFit sf;
    screen_fit32(n_samp, sx, sy, syx, sxx, syy, fox, foy, ref_compat != 0, sf);
    double Ad[9], td[3];
    for (int i = 0; i < 9; ++i) Ad[i] = sf.A[i];
    for (int i = 0; i < 3; ++i) td[i] = sf.t[i];
    // as the crop kernel evaluates it: isotropic closed form (+ its orthogonality allowance in the interval)
    const double tr = g.Sxx[0] + g.Sxx[3] + g.Sxx[5];
    const double r2 = residual_sq_iso(g.n, g.mux, g.muy, g.Syy, g.Syx, tr, Ad, td, (double)sf.s);
    const double e = screen_interval(sf, r2, (double)n, x_rms, tr);
    lo[h] = r2 - e;
    hi[h] = r2 + e;
  }
}

}  // extern "C"
