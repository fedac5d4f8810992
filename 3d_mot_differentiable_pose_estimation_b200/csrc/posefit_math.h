// posefit_math.h -- per-object 3x3 arithmetic of the pose solver, shared by the CUDA kernels
// (posefit_kernels.cu) and by a host-compiled checker (tests/host/math_check.cpp).
//
// What it implements (reference: PoseEst/pose_utils.py of the upstream repo):
//   * the Umeyama/Procrustes solve of estimateSimilarityUmeyama (pose_utils.py:16-61) from
//     accumulated moments instead of centred point arrays;
//   * the total residual of evaluateModel (pose_utils.py:5-9) in closed form from the global
//     second moments, so RANSAC hypotheses are ranked without a per-point pass;
//   * the adjoint of the fit (no reference exists: the upstream code detaches before the
//     fit, Detection/tracker/postprocess.py:151).
//
// The rotation is NOT taken from a LAPACK-style SVD.  R = U S V^T with S = diag(1,1,det(U)det(V))
// (pose_utils.py:38-44) is the maximiser of tr(R^T C) over SO(3); we get a start from a
// one-sided (Hestenes) Jacobi sweep in registers and polish it with Newton steps on SO(3),
// which also yields H = R^T C and L = tr(H) I - H -- everything the scale and the backward
// pass need -- without ever forming U, V or the singular values.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define PF_HD __host__ __device__ __forceinline__
#define PF_NOINLINE __host__ __device__ __noinline__
#else
#define PF_HD inline
#define PF_NOINLINE inline
#endif

namespace posefit {

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
PF_HD float pf_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
PF_HD double pf_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}
// a / b for the Jacobi rotation angles: the float start only has to land inside Newton's quadratic
// basin, so the device takes the 2-ulp MUFU-based quotient instead of the IEEE sequence
PF_HD float pf_div(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdividef(a, b);
#else
  return a / b;
#endif
}
PF_HD double pf_div(double a, double b) { return a / b; }
PF_HD float pf_abs(float x) { return fabsf(x); }
PF_HD double pf_abs(double x) { return fabs(x); }
PF_HD float pf_sqrt(float x) { return sqrtf(x); }
PF_HD double pf_sqrt(double x) { return sqrt(x); }

template <typename T>
PF_HD void cross3(const T* a, const T* b, T* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// ---------------------------------------------------------------------------------------------
// One-sided Jacobi on a 3x3 matrix held as three columns: on exit the columns of `a` are
// mutually orthogonal (a = A V), `v` holds V.  High relative accuracy for small singular
// values, which a Jacobi on A^T A would lose.
// ---------------------------------------------------------------------------------------------
template <typename T>
PF_HD void jacobi_pair(T* ap, T* aq, T* vp, T* vq) {
  const T alpha = ap[0] * ap[0] + ap[1] * ap[1] + ap[2] * ap[2];
  const T beta = aq[0] * aq[0] + aq[1] * aq[1] + aq[2] * aq[2];
  const T gamma = ap[0] * aq[0] + ap[1] * aq[1] + ap[2] * aq[2];
  const T eps = sizeof(T) == 4 ? (T)1e-14 : (T)1e-30;      // (relative tolerance)^2
  if (!(gamma * gamma > eps * alpha * beta)) return;        // already orthogonal (or NaN / zero)
  const T zeta = pf_div(beta - alpha, (T)2 * gamma);
  const T tt = pf_div(zeta >= (T)0 ? (T)1 : (T)-1, pf_abs(zeta) + pf_sqrt((T)1 + zeta * zeta));
  const T c = pf_rsqrt((T)1 + tt * tt);
  const T s = c * tt;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const T x = ap[i], y = aq[i];
    ap[i] = c * x - s * y;
    aq[i] = s * x + c * y;
    const T vx = vp[i], vy = vq[i];
    vp[i] = c * vx - s * vy;
    vq[i] = s * vx + c * vy;
  }
}

template <typename T>
PF_HD void swap_cols(T* a, T* b) {
#pragma unroll
  for (int i = 0; i < 3; ++i) { const T t = a[i]; a[i] = b[i]; b[i] = t; }
}

// Start rotation from the covariance C (row-major 3x3, any scale).  Returns false when C == 0
// (numpy's SVD of the zero matrix gives U = Vh = I, so the reference rotation is I).
template <typename T, int SWEEPS>
PF_HD bool rotation_start(const double* C, double* R) {
  double m = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) m = fmax(m, fabs(C[i]));
  if (!(m > 0.0) || !(m < 1e300)) {
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return false;
  }
  const double inv = 1.0 / m;
  T a0[3], a1[3], a2[3], v0[3] = {1, 0, 0}, v1[3] = {0, 1, 0}, v2[3] = {0, 0, 1};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    a0[i] = (T)(C[3 * i + 0] * inv);
    a1[i] = (T)(C[3 * i + 1] * inv);
    a2[i] = (T)(C[3 * i + 2] * inv);
  }
#pragma unroll 1
  for (int sweep = 0; sweep < SWEEPS; ++sweep) {
    jacobi_pair(a0, a1, v0, v1);
    jacobi_pair(a0, a2, v0, v2);
    jacobi_pair(a1, a2, v1, v2);
  }
  T n0 = a0[0] * a0[0] + a0[1] * a0[1] + a0[2] * a0[2];
  T n1 = a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2];
  T n2 = a2[0] * a2[0] + a2[1] * a2[1] + a2[2] * a2[2];
  // bring the two largest singular directions to slots 0 and 1
  if (n0 < n1) { swap_cols(a0, a1); swap_cols(v0, v1); const T t = n0; n0 = n1; n1 = t; }
  if (n0 < n2) { swap_cols(a0, a2); swap_cols(v0, v2); const T t = n0; n0 = n2; n2 = t; }
  if (n1 < n2) { swap_cols(a1, a2); swap_cols(v1, v2); const T t = n1; n1 = n2; n2 = t; }
  T u0[3], u1[3], u2[3], w2[3];
  const T r0 = pf_rsqrt(n0);
#pragma unroll
  for (int i = 0; i < 3; ++i) u0[i] = a0[i] * r0;
  T d = u0[0] * a1[0] + u0[1] * a1[1] + u0[2] * a1[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) u1[i] = a1[i] - d * u0[i];
  T l1 = u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2];
  const T tiny = sizeof(T) == 4 ? (T)1e-12 : (T)1e-28;
  if (!(l1 > tiny)) {                       // rank one: any unit vector orthogonal to u0 will do
    const T x = pf_abs(u0[0]), y = pf_abs(u0[1]), z = pf_abs(u0[2]);
    T e[3] = {0, 0, 0};
    if (x <= y && x <= z) e[0] = 1; else if (y <= z) e[1] = 1; else e[2] = 1;
    d = u0[0] * e[0] + u0[1] * e[1] + u0[2] * e[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) u1[i] = e[i] - d * u0[i];
    l1 = u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2];
  }
  const T r1 = pf_rsqrt(l1);
#pragma unroll
  for (int i = 0; i < 3; ++i) u1[i] *= r1;
  cross3(u0, u1, u2);
  cross3(v0, v1, w2);                       // det(+1) on both sides == the reflection fix of :39-42
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      R[3 * i + j] = (double)u0[i] * (double)v0[j] + (double)u1[i] * (double)v1[j] + (double)u2[i] * (double)w2[j];
  return true;
}

// R <- R (1.5 I - 0.5 R^T R): one Newton-Schulz step towards the nearest orthogonal matrix.
PF_HD void orthonormalize_step(double* R) {
  double G[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      G[3 * i + j] = -0.5 * (R[i] * R[j] + R[3 + i] * R[3 + j] + R[6 + i] * R[6 + j]) + (i == j ? 1.5 : 0.0);
  double N[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      N[3 * i + j] = R[3 * i] * G[j] + R[3 * i + 1] * G[3 + j] + R[3 * i + 2] * G[6 + j];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = N[i];
}

// M = R^T C
PF_HD void rt_times(const double* R, const double* C, double* M) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      M[3 * i + j] = R[i] * C[j] + R[3 + i] * C[3 + j] + R[6 + i] * C[6 + j];
}

// inverse of the symmetric 3x3 L = tr(H) I - H given H = (h00,h01,h02,h11,h12,h22).
// Returns det(L); Li is only written when |det| is usable.
PF_HD double inv_trace_minus(const double* H, double* Li) {
  const double tr = H[0] + H[3] + H[5];
  const double l00 = tr - H[0], l11 = tr - H[3], l22 = tr - H[5];
  const double l01 = -H[1], l02 = -H[2], l12 = -H[4];
  const double c00 = l11 * l22 - l12 * l12;
  const double c01 = l02 * l12 - l01 * l22;
  const double c02 = l01 * l12 - l02 * l11;
  const double det = l00 * c00 + l01 * c01 + l02 * c02;
  if (fabs(det) > 1e-200 && fabs(det) < 1e200) {
    const double r = 1.0 / det;
    Li[0] = c00 * r; Li[1] = c01 * r; Li[2] = c02 * r;
    Li[3] = (l00 * l22 - l02 * l02) * r;
    Li[4] = (l01 * l02 - l00 * l12) * r;
    Li[5] = (l00 * l11 - l01 * l01) * r;
  }
  return det;
}

// One Newton step on SO(3) towards skew(R^T C) = 0.  Returns |k|_inf (the skew residual before
// the step) so callers can test convergence.
PF_HD double newton_step(const double* C, double* R) {
  double M[9];
  rt_times(R, C, M);
  const double H[6] = {M[0], 0.5 * (M[1] + M[3]), 0.5 * (M[2] + M[6]), M[4], 0.5 * (M[5] + M[7]), M[8]};
  const double k0 = M[7] - M[5], k1 = M[2] - M[6], k2 = M[3] - M[1];
  double Li[6] = {0, 0, 0, 0, 0, 0};
  const double det = inv_trace_minus(H, Li);
  const double kn = fmax(fabs(k0), fmax(fabs(k1), fabs(k2)));
  if (!(fabs(det) > 1e-200)) return kn;
  double w0 = Li[0] * k0 + Li[1] * k1 + Li[2] * k2;
  double w1 = Li[1] * k0 + Li[3] * k1 + Li[4] * k2;
  double w2 = Li[2] * k0 + Li[4] * k1 + Li[5] * k2;
  double th2 = w0 * w0 + w1 * w1 + w2 * w2;
  if (!(th2 < 1e300)) return kn;
  if (th2 > 0.25) {                           // damp wild steps (ill-conditioned start)
    const double f = 0.5 * pf_rsqrt(th2);
    w0 *= f; w1 *= f; w2 *= f; th2 = 0.25;
  }
  // Rodrigues with series coefficients (|w| <= 0.5): a = sin(t)/t, b = (1-cos t)/t^2
  const double a = 1.0 - th2 * (1.0 / 6.0 - th2 * (1.0 / 120.0 - th2 * (1.0 / 5040.0 - th2 * (1.0 / 362880.0))));
  const double b = 0.5 - th2 * (1.0 / 24.0 - th2 * (1.0 / 720.0 - th2 * (1.0 / 40320.0 - th2 * (1.0 / 3628800.0))));
  // E = I + a [w]x + b [w]x^2,  [w]x^2 = w w^T - |w|^2 I
  double E[9];
  E[0] = 1.0 + b * (w0 * w0 - th2); E[1] = -a * w2 + b * w0 * w1;      E[2] = a * w1 + b * w0 * w2;
  E[3] = a * w2 + b * w0 * w1;      E[4] = 1.0 + b * (w1 * w1 - th2); E[5] = -a * w0 + b * w1 * w2;
  E[6] = -a * w1 + b * w0 * w2;     E[7] = a * w0 + b * w1 * w2;      E[8] = 1.0 + b * (w2 * w2 - th2);
  double N[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      N[3 * i + j] = R[3 * i] * E[j] + R[3 * i + 1] * E[3 + j] + R[3 * i + 2] * E[6 + j];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = N[i];
  return kn;
}

// ---------------------------------------------------------------------------------------------
// The fit.  Sums are UNcentred, over the points that take part (weights 0/1).
// ---------------------------------------------------------------------------------------------
struct Moments {
  double n;        // number of points
  double sx[3];    // sum x           (x = noc - 0.5, "source")
  double sy[3];    // sum y           (y = back-projected depth point, "target")
  double syx[9];   // sum y_i x_j     row-major [i][j]
  double sxx;      // sum |x|^2
};

struct Fit {
  double s;        // scale (Scales[0], pose_utils.py:47-52)
  double R[9];     // TRUE rotation R = U S V^T; the reference reports its transpose (:44)
  double t[3];     // Translation (:55)
  double mux[3], muy[3];
  double var;      // sum of per-axis population variances of x (:46)
  double H[6];     // sym part of R^T C  (h00,h01,h02,h11,h12,h22)
  double Linv[6];  // (tr(H) I - H)^-1, zero when singular
  double n;
  int status;      // 0 ok, 1 empty, 3 NaN covariance (:32-36)
};

enum { PF_OK = 0, PF_EMPTY = 1, PF_LOW_INLIER_RATIO = 2, PF_NAN = 3 };

// Cold path of solve_rotation, kept out of line so the straight-line code a solve walks through
// (and has to fetch: these kernels run once per object, instruction-cache cold) stays short.
struct Mat3 { double m[9]; };
PF_NOINLINE Mat3 rotation_redo(Mat3 C, Mat3 Cn) {
  Mat3 R;
  rotation_start<double, 6>(C.m, R.m);
  newton_step(Cn.m, R.m);
  newton_step(Cn.m, R.m);
  return R;
}

// Rotation maximising tr(R^T C) over SO(3) plus H (and Linv when WANT_LINV).
// Start: one-sided Jacobi in float (3 sweeps, ~1e-6), one Newton-Schulz orthonormalisation, then
// Newton steps on SO(3) in double.  Newton converges quadratically, so the skew residual measured
// BEFORE a step bounds the error after it by its square; if it has not collapsed before the
// second step the solve is redone from a double-precision Jacobi start (near-degenerate C only).
// EXTRA_STEP adds a third Newton step (used for the once-per-object fits; hypotheses skip it).
// `start` (optional): a rotation within ~1e-5 of the solution (the float fit of the RANSAC screen); the Jacobi start is
// skipped, everything else -- orthonormalisation, Newton steps, the convergence test with its double-precision redo --
// is unchanged, so the result is the same fixed point.
template <bool EXTRA_STEP, bool WANT_LINV>
PF_HD void solve_rotation(const double* C, double* R, double* H, double* Linv, const float* start = nullptr,
                          bool use_start = false) {
  double m = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) m = fmax(m, fabs(C[i]));
  bool nonzero;
  if (use_start) {
    nonzero = (m > 0.0) && (m < 1e300);
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = nonzero ? (double)start[i] : ((i % 4 == 0) ? 1.0 : 0.0);
  } else {
    nonzero = rotation_start<float, 3>(C, R);
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) { H[i] = 0.0; if (WANT_LINV) Linv[i] = 0.0; }
  if (!nonzero) return;
  double Cn[9];
  const double inv = 1.0 / m;
#pragma unroll
  for (int i = 0; i < 9; ++i) Cn[i] = C[i] * inv;
  orthonormalize_step(R);
  // Newton steps share ONE loop body (code size); kn = residual BEFORE the second step.  If it has
  // not collapsed the start was outside the quadratic regime: redo from a double-precision start.
  double kn = 0.0;
#pragma unroll 1
  for (int it = 0; it < (EXTRA_STEP ? 3 : 2); ++it) {
    // (the third step is data dependent: a residual below 1e-10 before the second step leaves ~1e-20 after it)
    if (it == 2 && kn < 1e-10) break;
    if (it == 2 && !(kn < 1e-8)) {
      Mat3 c3, cn3;
#pragma unroll
      for (int i = 0; i < 9; ++i) { c3.m[i] = C[i]; cn3.m[i] = Cn[i]; }
      const Mat3 r3 = rotation_redo(c3, cn3);
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = r3.m[i];
    }
    const double k = newton_step(Cn, R);
    if (it == 1) kn = k;
  }
  if (!EXTRA_STEP && !(kn < 1e-8)) {
    Mat3 c3, cn3;
#pragma unroll
    for (int i = 0; i < 9; ++i) { c3.m[i] = C[i]; cn3.m[i] = Cn[i]; }
    const Mat3 r3 = rotation_redo(c3, cn3);
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = r3.m[i];
  }
  orthonormalize_step(R);
  double M[9];
  rt_times(R, Cn, M);
  double Hn[6] = {M[0], 0.5 * (M[1] + M[3]), 0.5 * (M[2] + M[6]), M[4], 0.5 * (M[5] + M[7]), M[8]};
#pragma unroll
  for (int i = 0; i < 6; ++i) H[i] = Hn[i] * m;
  if (WANT_LINV) {
    double Li[6] = {0, 0, 0, 0, 0, 0};
    inv_trace_minus(Hn, Li);
#pragma unroll
    for (int i = 0; i < 6; ++i) Linv[i] = Li[i] * inv;
  }
}

// ox / oy (optional): origin the sums were shifted by (x - ox, y - oy were accumulated); C and
// var are shift invariant, the means are restored before t is formed.
template <bool PRECISE>
PF_HD void fit_from_moments(const Moments& mo, Fit& f, const double* ox = nullptr, const double* oy = nullptr,
                            const float* start = nullptr, bool use_start = false) {
  f.n = mo.n;
  f.s = 1.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) f.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) { f.t[i] = 0.0; f.mux[i] = 0.0; f.muy[i] = 0.0; }
#pragma unroll
  for (int i = 0; i < 6; ++i) { f.H[i] = 0.0; f.Linv[i] = 0.0; }
  f.var = 0.0;
  if (!(mo.n > 0.0)) { f.status = PF_EMPTY; return; }
  const double rn = 1.0 / mo.n;
  double C[9];
  bool nan = false;
#pragma unroll
  for (int i = 0; i < 3; ++i) { f.mux[i] = mo.sx[i] * rn; f.muy[i] = mo.sy[i] * rn; }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      C[3 * i + j] = mo.syx[3 * i + j] * rn - f.muy[i] * f.mux[j];
      nan = nan || (C[3 * i + j] != C[3 * i + j]);
    }
  if (nan) { f.status = PF_NAN; return; }
  f.var = mo.sxx * rn - (f.mux[0] * f.mux[0] + f.mux[1] * f.mux[1] + f.mux[2] * f.mux[2]);
  if (mo.n == 1.0) {
    // a single correspondence: the centred cloud is exactly zero in the reference (U = Vh = I, s = 1,
    // pose_utils.py:27-50); uncentred sums only reach that up to rounding, so say it explicitly
#pragma unroll
    for (int i = 0; i < 9; ++i) C[i] = 0.0;
    f.var = 0.0;
  }
  solve_rotation<PRECISE, PRECISE>(C, f.R, f.H, f.Linv, start, use_start);   // hypotheses never need Linv
  if (ox != nullptr) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { f.mux[i] += ox[i]; f.muy[i] += oy[i]; }
  }
  const double trh = f.H[0] + f.H[3] + f.H[5];           // = sum(D) after the sign fix (:39-42)
  f.s = (f.var * trh != 0.0) ? (1.0 / f.var) * trh : 1.0;   // pose_utils.py:47-50
#pragma unroll
  for (int i = 0; i < 3; ++i)
    f.t[i] = f.muy[i] - f.s * (f.R[3 * i] * f.mux[0] + f.R[3 * i + 1] * f.mux[1] + f.R[3 * i + 2] * f.mux[2]);
  f.status = PF_OK;
}

// ---------------------------------------------------------------------------------------------
// Closed-form total residual of evaluateModel (pose_utils.py:7-9) for the transform [A | t]:
//   sum_i |y_i - A x_i - t|^2 = Syy - 2 <A, Syx> + <A Sxx, A> + n |mu_y - A mu_x - t|^2
// with CENTRED global sums Syy = sum |y~|^2, Syx = sum y~ x~^T, Sxx = sum x~ x~^T.
// ---------------------------------------------------------------------------------------------
struct GlobalStats {
  double n;
  double mux[3], muy[3];
  double Syy;
  double Syx[9];
  double Sxx[6];   // xx, xy, xz, yy, yz, zz
};

PF_HD double residual_sq(const GlobalStats& g, const double* A, const double* t) {
  double acc = g.Syy;
  double lin = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) lin += A[i] * g.Syx[i];
  acc -= 2.0 * lin;
  const double X[9] = {g.Sxx[0], g.Sxx[1], g.Sxx[2], g.Sxx[1], g.Sxx[3], g.Sxx[4], g.Sxx[2], g.Sxx[4], g.Sxx[5]};
  double quad = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      quad += (A[3 * i] * X[j] + A[3 * i + 1] * X[3 + j] + A[3 * i + 2] * X[6 + j]) * A[3 * i + j];
  acc += quad;
  double dd = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double d = g.muy[i] - (A[3 * i] * g.mux[0] + A[3 * i + 1] * g.mux[1] + A[3 * i + 2] * g.mux[2]) - t[i];
    dd += d * d;
  }
  return acc + g.n * dd;
}

// ---------------------------------------------------------------------------------------------
// fp32 SCREEN of a RANSAC hypothesis (fit_ransac_crop_kernel).
//
// The reference ranks the hypotheses by their total residual (pose_utils.py:76); only the winner's
// transform is ever used.  Fitting all of them in double is what kept the RANSAC kernel latency-bound
// (profiles/r01_n_*), so every hypothesis is first fitted in float -- same algorithm, one-sided Jacobi
// start, no Newton polish -- and its residual is evaluated (in double, closed form) for that float
// transform.  The float fit comes with an a-posteriori error estimate `rho` (relative error of the
// transform: Newton residual and rounding of the covariance over the smallest eigenvalue of
// L = tr(H) I - H, plus the rounding of the scale), from which the kernel derives an interval
// [r2 - err, r2 + err] that contains the residual the double fit would give.  Only hypotheses whose
// interval reaches below the smallest upper end (or below the stop threshold) are then fitted in
// double by the very code the v1 kernel ran for all of them, so winner, transform and inlier mask are
// bit-identical to a full double evaluation whenever the intervals hold; rho = +inf (degenerate or
// ill-conditioned sample sets) forces the double fit.  tests/test_math_host.py checks the interval on
// benchmark-like, clean, heavily contaminated, sparse and tiny clouds.
// ---------------------------------------------------------------------------------------------
struct ScreenFit {
  float A[9];      // scoring transform (s * R^T when ref_compat, else s * R), row-major
  float t[3];
  float s;
  float R[9];      // the float rotation (true R), start of the double polish of a candidate
  float mx[3], my[3];   // sample means (unshifted)
  float q;         // max |R^T R - I|: the float rotation's distance from orthogonality
  float rho;       // relative error estimate of A and t's rotation part; +inf = not usable
};

constexpr float kScreenEps = 1.1920929e-7f;       // 2^-23

// sums are over x - ox, y - oy (ox, oy = first sample), n = number of samples (the first included).
// Rotation: Markley's FOAM closed form instead of a Jacobi sweep (about a quarter of the instructions, no dependent
// chain of nine pair rotations): with B = C / |C|_F, the largest root lambda of
//   psi(l) = (l^2 - 1)^2 - 8 l det B - 4 |adj B|_F^2          (roots: +-s1 +-s2 +-s3, even number of minus signs)
// found by Newton from sqrt(3) >= lambda (monotone from above), gives the maximiser of tr(R^T C) over SO(3) as
//   R = [ (kappa + 1) B + lambda adj(B)^T - B B^T B ] / zeta,  kappa = (lambda^2 - 1) / 2,  zeta = kappa lambda - det B,
// where zeta = (s1+s2)(s1+s3)(s2+s3) = det L, so 4 zeta / (lambda + 1)^2 bounds the smallest eigenvalue of L
// from below.  Whatever the float arithmetic loses (small zeta) shows up in the checks that
// follow -- Newton residual of R, its distance from orthogonality -- and widens the interval.
PF_HD void screen_fit32(int n, const float* sx, const float* sy, const float* syx, float sxx, float syy,
                        const float* ox, const float* oy, bool ref_compat, ScreenFit& f) {
  const float inf = __builtin_huge_valf();
  const float rn = 1.0f / (float)n;
  float mux[3], muy[3], C[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) { mux[i] = sx[i] * rn; muy[i] = sy[i] * rn; }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = syx[3 * i + j] * rn - muy[i] * mux[j];
  const float ex = sxx * rn, ey = syy * rn;
  const float var = ex - (mux[0] * mux[0] + mux[1] * mux[1] + mux[2] * mux[2]);
  const float mag = pf_sqrt(ex * ey);               // >= every |syx| * rn (Cauchy-Schwarz)
#pragma unroll
  for (int i = 0; i < 3; ++i) { f.mx[i] = mux[i] + ox[i]; f.my[i] = muy[i] + oy[i]; }
  float nf2 = 0.0f;
#pragma unroll
  for (int i = 0; i < 9; ++i) nf2 = fmaf(C[i], C[i], nf2);
  f.rho = inf;
  f.q = 0.0f;
  f.s = 1.0f;
#pragma unroll
  for (int i = 0; i < 9; ++i) { f.A[i] = (i % 4 == 0) ? 1.0f : 0.0f; f.R[i] = f.A[i]; }
#pragma unroll
  for (int i = 0; i < 3; ++i) f.t[i] = f.my[i] - f.mx[i];
  if (!(nf2 > 1e-10f * mag * mag) || !(mag < 1e18f) || !(var > 0.0f)) return;   // degenerate / non-finite: double decides
  const float inv = pf_rsqrt(nf2);                  // 1 / |C|_F
  float B[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) B[i] = C[i] * inv;
  // adjugate transposed (= cofactor matrix): cof[3 i + j] = cofactor of B[i][j]
  float cof[9];
  cof[0] = B[4] * B[8] - B[5] * B[7]; cof[1] = B[5] * B[6] - B[3] * B[8]; cof[2] = B[3] * B[7] - B[4] * B[6];
  cof[3] = B[2] * B[7] - B[1] * B[8]; cof[4] = B[0] * B[8] - B[2] * B[6]; cof[5] = B[1] * B[6] - B[0] * B[7];
  cof[6] = B[1] * B[5] - B[2] * B[4]; cof[7] = B[2] * B[3] - B[0] * B[5]; cof[8] = B[0] * B[4] - B[1] * B[3];
  const float det = B[0] * cof[0] + B[1] * cof[1] + B[2] * cof[2];
  float c4 = 0.0f;
#pragma unroll
  for (int i = 0; i < 9; ++i) c4 = fmaf(cof[i], cof[i], c4);
  c4 *= 4.0f;
  const float b8 = 8.0f * det;
  float lam = 1.7320508f, step = 1.0f, prev = 2.0f;
#pragma unroll 1
  for (int it = 0; it < 24 && step > 1e-6f * lam && step < prev; ++it) {   // Newton on the quartic, from above: the steps
    prev = step;                                             //  shrink monotonically until rounding takes over (5-6 steps;
    const float t = fmaf(lam, lam, -1.0f);                   //  a near-double root -- s2 + s3 small -- converges linearly
    const float psi = fmaf(t, t, -fmaf(b8, lam, c4));        //  and takes more)
    const float dpsi = fmaf(4.0f * lam, t, -b8);
    step = pf_div(psi, dpsi);
    lam -= step;
  }
  const float kappa = 0.5f * fmaf(lam, lam, -1.0f);
  const float zeta = fmaf(kappa, lam, -det);
  if (!(zeta > 1e-5f) || !(lam > 0.0f)) return;     // ill-conditioned rotation (or NaN): double decides
  // G = B^T B (symmetric), P = B G
  float G[6];
  G[0] = B[0] * B[0] + B[3] * B[3] + B[6] * B[6]; G[1] = B[0] * B[1] + B[3] * B[4] + B[6] * B[7];
  G[2] = B[0] * B[2] + B[3] * B[5] + B[6] * B[8]; G[3] = B[1] * B[1] + B[4] * B[4] + B[7] * B[7];
  G[4] = B[1] * B[2] + B[4] * B[5] + B[7] * B[8]; G[5] = B[2] * B[2] + B[5] * B[5] + B[8] * B[8];
  const float rz = 1.0f / zeta, k1 = kappa + 1.0f;
  float R[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float b0 = B[3 * i], b1 = B[3 * i + 1], b2 = B[3 * i + 2];
    R[3 * i] = (k1 * b0 + lam * cof[3 * i] - (b0 * G[0] + b1 * G[1] + b2 * G[2])) * rz;
    R[3 * i + 1] = (k1 * b1 + lam * cof[3 * i + 1] - (b0 * G[1] + b1 * G[3] + b2 * G[4])) * rz;
    R[3 * i + 2] = (k1 * b2 + lam * cof[3 * i + 2] - (b0 * G[2] + b1 * G[4] + b2 * G[5])) * rz;
  }
  // Newton-Schulz towards O(3), then one float Newton step on SO(3) (the very step newton_step takes in double): FOAM
  // loses digits where zeta is small, the two steps give them back as far as float can hold them
  {
    float Gm[9], N[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        Gm[3 * i + j] = -0.5f * (R[i] * R[j] + R[3 + i] * R[3 + j] + R[6 + i] * R[6 + j]) + (i == j ? 1.5f : 0.0f);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) N[3 * i + j] = R[3 * i] * Gm[j] + R[3 * i + 1] * Gm[3 + j] + R[3 * i + 2] * Gm[6 + j];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = N[i];
  }
  float M[9];
  {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) M[3 * i + j] = R[i] * B[j] + R[3 + i] * B[3 + j] + R[6 + i] * B[6 + j];
    const float h01 = 0.5f * (M[1] + M[3]), h02 = 0.5f * (M[2] + M[6]), h12 = 0.5f * (M[5] + M[7]);
    const float tr = M[0] + M[4] + M[8];
    const float l00 = tr - M[0], l11 = tr - M[4], l22 = tr - M[8], l01 = -h01, l02 = -h02, l12 = -h12;
    const float c00 = l11 * l22 - l12 * l12, c01 = l02 * l12 - l01 * l22, c02 = l01 * l12 - l02 * l11;
    const float dl = l00 * c00 + l01 * c01 + l02 * c02;
    const float k0 = M[7] - M[5], k1s = M[2] - M[6], k2 = M[3] - M[1];
    if (dl > 1e-6f) {
      const float rd = 1.0f / dl;
      const float c11 = l00 * l22 - l02 * l02, c12 = l01 * l02 - l00 * l12, c22 = l00 * l11 - l01 * l01;
      const float w0 = (c00 * k0 + c01 * k1s + c02 * k2) * rd;
      const float w1 = (c01 * k0 + c11 * k1s + c12 * k2) * rd;
      const float w2 = (c02 * k0 + c12 * k1s + c22 * k2) * rd;
      const float th2 = w0 * w0 + w1 * w1 + w2 * w2;
      if (th2 < 0.01f) {                            // a correction, not a rescue
        const float a = 1.0f - th2 * (1.0f / 6.0f), b = 0.5f - th2 * (1.0f / 24.0f);
        float E[9];
        E[0] = 1.0f + b * (w0 * w0 - th2); E[1] = -a * w2 + b * w0 * w1;      E[2] = a * w1 + b * w0 * w2;
        E[3] = a * w2 + b * w0 * w1;      E[4] = 1.0f + b * (w1 * w1 - th2); E[5] = -a * w0 + b * w1 * w2;
        E[6] = -a * w1 + b * w0 * w2;     E[7] = a * w0 + b * w1 * w2;      E[8] = 1.0f + b * (w2 * w2 - th2);
        float N[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) N[3 * i + j] = R[3 * i] * E[j] + R[3 * i + 1] * E[3 + j] + R[3 * i + 2] * E[6 + j];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = N[i];
      }
    }
  }
  // checks: M = R^T B (trace -> scale, skew part -> Newton residual of R), Q = R^T R - I (distance from O(3))
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) M[3 * i + j] = R[i] * B[j] + R[3 + i] * B[3 + j] + R[6 + i] * B[6 + j];
  const float trh = M[0] + M[4] + M[8];
  const float kn = fmaxf(pf_abs(M[7] - M[5]), fmaxf(pf_abs(M[2] - M[6]), pf_abs(M[3] - M[1])));
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j)
      q = fmaxf(q, pf_abs(R[i] * R[j] + R[3 + i] * R[3 + j] + R[6 + i] * R[6 + j] - (i == j ? 1.0f : 0.0f)));
  if (!(trh > 0.0f) || !(q < 1e-2f)) return;
  const float dc = 8.0f * kScreenEps * mag * inv;   // rounding of the (normalised) covariance entries
  const float s = trh / (inv * var);                // pose_utils.py:47-50: tr(R^T C) / var
  // smallest eigenvalue of L: s2 + s3 = zeta / ((s1+s2)(s1+s3)) and (s1+s2)(s1+s3) <= ((lambda + s1) / 2)^2 <= ((lambda + 1) / 2)^2
  const float gap = 4.0f * zeta / ((lam + 1.0f) * (lam + 1.0f));
  f.rho = (kn + dc) / gap + 2.0f * q + 3.0f * dc / trh + 8.0f * kScreenEps * ex / var;
  f.s = s;
  f.q = q;
#pragma unroll
  for (int i = 0; i < 9; ++i) f.R[i] = R[i];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) f.A[3 * i + j] = s * (ref_compat ? R[3 * j + i] : R[3 * i + j]);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    f.t[i] = f.my[i] - s * (R[3 * i] * f.mx[0] + R[3 * i + 1] * f.mx[1] + R[3 * i + 2] * f.mx[2]);
}

// Half-width of the interval around r2 (the residual^2 of the float transform, evaluated in double) that
// contains the residual^2 of the double fit: |d r2| <= 2 sqrt(r2) (|dA| sqrt(sum |x|^2) + |dt| sqrt(n)) to first
// order; x_rms = sqrt(sum_i |x_i|^2 / n) over ALL correspondences.  The leading factor 2 is a safety margin on top of the
// worst-case rounding constants (tests/test_math_host.py: no error above 5 % of the half-width on any regime).
PF_HD double screen_interval(const ScreenFit& f, double r2, double n_all, double x_rms, double tr_sxx = 0.0) {
  if (!(f.rho < 1e30f) || !(r2 >= 0.0) || !(r2 < 1e300)) return __builtin_huge_val();
  const double s = (double)pf_abs(f.s), rho = (double)f.rho;
  const double amx = (double)(pf_abs(f.mx[0]) + pf_abs(f.mx[1]) + pf_abs(f.mx[2]));
  const double amy = (double)(pf_abs(f.my[0]) + pf_abs(f.my[1]) + pf_abs(f.my[2]));
  const double e8 = 8.0 * (double)kScreenEps;
  const double lin = s * rho * (x_rms + amx) + e8 * (amy + s * amx);
  // tr_sxx > 0: r2 came from residual_sq_iso, whose quadratic term assumes an exactly orthogonal rotation
  return 2.0 * (2.0 * sqrt(r2 * n_all) * lin + n_all * lin * lin) + 3.0 * (double)f.q * s * s * tr_sxx;
}

// The same closed form for a SCALED ROTATION A = s Q (Q orthogonal): <A Sxx, A> = s^2 tr(Sxx), so only the trace of
// the centred source scatter is needed (the crop kernel accumulates 18 sums per pixel instead of 23).  For a Q that is
// orthogonal only up to q = max |Q^T Q - I| the quadratic term is off by at most 3 q s^2 tr(Sxx): callers add that to
// their interval (the double fits end in a Newton-Schulz step: q ~ 1e-16).
PF_HD double residual_sq_iso(double n, const double* mux, const double* muy, double Syy, const double* Syx,
                             double trSxx, const double* A, const double* t, double s) {
  double lin = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) lin = fma(A[i], Syx[i], lin);
  double dd = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double d = muy[i] - (A[3 * i] * mux[0] + A[3 * i + 1] * mux[1] + A[3 * i + 2] * mux[2]) - t[i];
    dd = fma(d, d, dd);
  }
  return fma(n, dd, fma(s * s, trSxx, fma(-2.0, lin, Syy)));
}

// Scoring transform of a fit.  ref_compat: A = s * R^T, the block the reference really builds
// (pose_utils.py:58); otherwise the geometrically correct A = s * R.
PF_HD void scoring_transform(const Fit& f, bool ref_compat, double* A) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A[3 * i + j] = f.s * (ref_compat ? f.R[3 * j + i] : f.R[3 * i + j]);
}

// ---------------------------------------------------------------------------------------------
// Adjoint of the fit.  Given dL/ds, dL/dR (w.r.t. the TRUE rotation), dL/dt, produce the
// per-point coefficients:  dL/dx_i = (w_i/n) (GC^T y~_i + 2 gvar x~_i + gmux),
//                          dL/dy_i = (w_i/n) (GC x~_i + gmuy)
// Rotation part: dR = R [dw]x with (tr(H) I - H) dw = axial(R^T dC - dC^T R)  =>  GC_rot = R [p]x,
// p = Linv q, q = axial(R^T G_R) (same linear system as the forward Newton step).
// ---------------------------------------------------------------------------------------------
struct FitAdjoint {
  double GC[9];
  double gvar;
  double gmux[3], gmuy[3];
};

PF_HD void fit_adjoint(const Fit& f, double gs, const double* gR, const double* gt, FitAdjoint& a) {
  const bool scale_live = (f.var * (f.H[0] + f.H[3] + f.H[5]) != 0.0);
  // t = muy - s R mux
  double Rmu[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) Rmu[i] = f.R[3 * i] * f.mux[0] + f.R[3 * i + 1] * f.mux[1] + f.R[3 * i + 2] * f.mux[2];
  double gs2 = gs - (gt[0] * Rmu[0] + gt[1] * Rmu[1] + gt[2] * Rmu[2]);
  double G[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) G[3 * i + j] = gR[3 * i + j] - f.s * gt[i] * f.mux[j];
  double Q[9];
  rt_times(f.R, G, Q);                                    // Q = R^T G
  const double q0 = Q[7] - Q[5], q1 = Q[2] - Q[6], q2 = Q[3] - Q[1];
  const double p0 = f.Linv[0] * q0 + f.Linv[1] * q1 + f.Linv[2] * q2;
  const double p1 = f.Linv[1] * q0 + f.Linv[3] * q1 + f.Linv[4] * q2;
  const double p2 = f.Linv[2] * q0 + f.Linv[4] * q1 + f.Linv[5] * q2;
  const double P[9] = {0.0, -p2, p1, p2, 0.0, -p0, -p1, p0, 0.0};
  if (!scale_live) gs2 = 0.0;                             // s is the constant 1 (pose_utils.py:50)
  const double ginv = (f.var != 0.0) ? gs2 / f.var : 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      a.GC[3 * i + j] = f.R[3 * i] * P[j] + f.R[3 * i + 1] * P[3 + j] + f.R[3 * i + 2] * P[6 + j] + ginv * f.R[3 * i + j];
  a.gvar = -ginv * f.s;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    a.gmux[i] = -f.s * (f.R[i] * gt[0] + f.R[3 + i] * gt[1] + f.R[6 + i] * gt[2]);
    a.gmuy[i] = gt[i];
  }
}

}  // namespace posefit
