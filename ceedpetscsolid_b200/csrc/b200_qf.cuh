// Device bodies of the recognised QFunctions (hand-written; the user headers
// /root/reference/qfunctions/*.h stay untouched and are what the oracle runs).
//
// Conventions:
//   H[c][m] = d u_c / d X_m          reference-space gradient of component c   (QFunction in0[m][c])
//   A[m][k] = dXdx[m][k]             qdata[1 + 3m + k]
//   g[c][k] = sum_m A[m][k] H[c][m]  physical gradient
//   W[c][k]                          QFunction out0[k][c] = w * sum_m A[k][m] T[c][m]
#pragma once
#include "b200_common.cuh"

namespace b200 {

// host + device: tests/c/qf_host_check.cu runs the SAME point functions on the CPU against the reference QFunctions
#define B200_DI __host__ __device__ __forceinline__

B200_DI void phys_grad(const double (&A)[3][3], const double (&H)[3][3], double (&g)[3][3]) {
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int k = 0; k < 3; k++) g[c][k] = A[0][k] * H[c][0] + A[1][k] * H[c][1] + A[2][k] * H[c][2];
}

// W[c][k] = sum_m A[k][m] * T[c][m]   (T already scaled by w*detJ)
B200_DI void pull_back(const double (&A)[3][3], const double (&T)[3][3], double (&W)[3][3]) {
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int k = 0; k < 3; k++) W[c][k] = A[k][0] * T[c][0] + A[k][1] * T[c][1] + A[k][2] * T[c][2];
}

// ------------------------------------------------------------------ linear elasticity
// qfunctions/linElas.h:97-153 (F) == :221-275 (dF).  The shear entries keep the
// reference's extra factor 0.5 (linElas.h:137-139).
B200_DI void linelas_point(const Material &mt, double w, const double (&A)[3][3],
                           const double (&H)[3][3], double (&W)[3][3]) {
  double g[3][3], T[3][3];
  phys_grad(A, H, g);
  const double c1 = mt.le_c1 * w, c2 = mt.le_c2 * w, c3 = mt.le_c3 * w;
  T[0][0] = c1 * g[0][0] + c2 * (g[1][1] + g[2][2]);
  T[1][1] = c1 * g[1][1] + c2 * (g[0][0] + g[2][2]);
  T[2][2] = c1 * g[2][2] + c2 * (g[0][0] + g[1][1]);
  T[0][1] = T[1][0] = c3 * (g[0][1] + g[1][0]);
  T[0][2] = T[2][0] = c3 * (g[0][2] + g[2][0]);
  T[1][2] = T[2][1] = c3 * (g[1][2] + g[2][1]);
  pull_back(A, T, W);
}

// ------------------------------------------------------------------ hyperSS
// qfunctions/hyperSS.h:43-55
B200_DI double log1p_series(double x) {
  // odd reciprocals multiplied instead of divided (see log1p_series_in_y below)
  double y = x / (2. + x);
  const double y2 = y * y;
  double sum = y;
  y *= y2; sum += y * (1. / 3);
  y *= y2; sum += y * (1. / 5);
  y *= y2; sum += y * (1. / 7);
  return 2 * sum;
}

// residual, qfunctions/hyperSS.h:112-177; g (= gradu to store) is returned
B200_DI void hyperss_f_point(const Material &mt, double w, const double (&A)[3][3],
                             const double (&H)[3][3], double (&g)[3][3], double (&W)[3][3]) {
  double T[3][3];
  phys_grad(A, H, g);
  const double llv = mt.lambda * log1p_series(g[0][0] + g[1][1] + g[2][2]) * w;
  const double mw = mt.mu * w;
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) T[a][b] = mw * (g[a][b] + g[b][a]) + (a == b ? llv : 0.);
  pull_back(A, T, W);
}

// Jacobian, qfunctions/hyperSS.h:239-316; s = 1 / (1 + tr gradu)
B200_DI void hyperss_df_point(const Material &mt, double w, const double (&A)[3][3], double s,
                              const double (&H)[3][3], double (&W)[3][3]) {
  double g[3][3], T[3][3];
  phys_grad(A, H, g);
  const double ltr = (mt.lambda * s * w) * (g[0][0] + g[1][1] + g[2][2]);
  const double mw = mt.mu * w;
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) T[a][b] = mw * (g[a][b] + g[b][a]) + (a == b ? ltr : 0.);
  pull_back(A, T, W);
}

// ------------------------------------------------------------------ hyperFS
// qfunctions/hyperFS.h:45-67
// range reduction of qfunctions/hyperFS.h:45-67: log1p(x) = shift + log1p(xr), xr in [1/sqrt2 - 1, sqrt2 - 1]
B200_DI double log1p_shift(double x, double &xr) {
  const double left = 0.70710678118654752440 - 1, right = 1.41421356237309504880 - 1;
  const double half_ln2 = 0.34657359027997265471;
  double shift = 0;
  if (x < left) {
    shift = -half_ln2;
    x = 1 + 2 * x;
  } else if (right < x) {
    shift = half_ln2;
    x = (x - 1) / 2;
  }
  xr = x;
  return shift;
}
// the reference's series in y = x / (2 + x): 2 (y + y^3/3 + y^5/5 + y^7/7).  The odd reciprocals are multiplied
// (the reference divides: differs by <= 1 ulp of a term that is itself < 2e-3 of the sum) -- an FP64 division costs
// ~15 instructions on the device and there are three of them per quadrature point otherwise.
B200_DI double log1p_series_in_y(double y) {
  const double y2 = y * y;
  double sum = y;
  y *= y2; sum += y * (1. / 3);
  y *= y2; sum += y * (1. / 5);
  y *= y2; sum += y * (1. / 7);
  return 2 * sum;
}
B200_DI double log1p_series_shifted(double x) {
  double xr;
  const double shift = log1p_shift(x, xr);
  return 2 * shift + log1p_series_in_y(xr / (2. + xr));
}

// 2E in Voigt order (00,11,22,12,02,01) and det(C) - 1 (hyperFS.h:72-80, :91-97)
B200_DI double green_lagrange2(const double (&g)[3][3], double (&e)[6]) {
  const int vj[6] = {0, 1, 2, 1, 0, 0}, vk[6] = {0, 1, 2, 2, 2, 1};
#pragma unroll
  for (int m = 0; m < 6; m++) {
    const int j = vj[m], k = vk[m];
    e[m] = g[j][k] + g[k][j] + g[0][j] * g[0][k] + g[1][j] * g[1][k] + g[2][j] * g[2][k];
  }
  return e[0] * (e[1] * e[2] - e[3] * e[3]) + e[5] * (e[4] * e[3] - e[5] * e[2]) +
         e[4] * (e[5] * e[3] - e[4] * e[1]) + e[0] + e[1] + e[2] + e[0] * e[1] + e[0] * e[2] +
         e[1] * e[2] - e[5] * e[5] - e[4] * e[4] - e[3] * e[3];
}

B200_DI void voigt_sym(const double (&v)[6], double (&m)[3][3]) {
  m[0][0] = v[0]; m[1][1] = v[1]; m[2][2] = v[2];
  m[1][2] = m[2][1] = v[3];
  m[0][2] = m[2][0] = v[4];
  m[0][1] = m[1][0] = v[5];
}

// commonFS, hyperFS.h:85-142: S, C^-1 (full symmetric) and llnj = lambda log J
B200_DI void fs_common(const Material &mt, const double (&g)[3][3], double (&S)[3][3],
                       double (&Ci)[3][3], double &llnj) {
  double e[6], E2[3][3], C[3][3], civ[6], sv[6];
  const double detC_m1 = green_lagrange2(g, e);
  voigt_sym(e, E2);
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) C[a][b] = E2[a][b] + (a == b ? 1. : 0.);
  const double Adj[6] = {C[1][1] * C[2][2] - C[1][2] * C[2][1], C[0][0] * C[2][2] - C[0][2] * C[2][0],
                         C[0][0] * C[1][1] - C[0][1] * C[1][0], C[0][2] * C[1][0] - C[0][0] * C[1][2],
                         C[0][1] * C[1][2] - C[0][2] * C[1][1], C[0][2] * C[2][1] - C[0][1] * C[2][2]};
  const double rdet = 1. / (detC_m1 + 1.);
#pragma unroll
  for (int m = 0; m < 6; m++) civ[m] = Adj[m] * rdet;
  voigt_sym(civ, Ci);
  llnj = mt.lambda * log1p_series_shifted(detC_m1) / 2.;
  const int vj[6] = {0, 1, 2, 1, 0, 0}, vk[6] = {0, 1, 2, 2, 2, 1};
#pragma unroll
  for (int m = 0; m < 6; m++) {
    double s = llnj * civ[m];
#pragma unroll
    for (int n = 0; n < 3; n++) s += mt.mu * Ci[vj[m]][n] * E2[n][vk[m]];
    sv[m] = s;
  }
  voigt_sym(sv, S);
}

// det(I + s) - 1 of a symmetric s in Voigt order (00,11,22,12,02,01), without cancellation for small s
// (the reference's computeDetCM1, hyperFS.h:72-80, applied to s = 2E)
B200_DI double det_sym_m1(const double (&e)[6]) {
  return e[0] * (e[1] * e[2] - e[3] * e[3]) + e[5] * (e[4] * e[3] - e[5] * e[2]) +
         e[4] * (e[5] * e[3] - e[4] * e[1]) + e[0] + e[1] + e[2] + e[0] * e[1] + e[0] * e[2] +
         e[1] * e[2] - e[5] * e[5] - e[4] * e[4] - e[3] * e[3];
}

// residual, hyperFS.h:212-276, in spatial form.  The reference evaluates S = lambda lnJ C^-1 + 2 mu C^-1 E and
// P = F S; with C^-1 = F^-1 F^-T this is exactly
//     P = tau F^-T,   tau = mu (b - I) + lambda lnJ I,   b - I = g + g^T + g g^T
// (Kirchhoff stress): no C^-1, no Voigt round trip, one reciprocal.  b - I is formed from g directly and
// lnJ = log1p(det b - 1) / 2 with det b - 1 = det C - 1 from the reference's cancellation-free polynomial and its
// own series (hyperFS.h:45-80), so small strains keep full relative accuracy as in the reference.
B200_DI void hyperfs_f_point(const Material &mt, double w, const double (&A)[3][3],
                             const double (&H)[3][3], double (&g)[3][3], double (&W)[3][3]) {
  phys_grad(A, H, g);
  const int vj[6] = {0, 1, 2, 1, 0, 0}, vk[6] = {0, 1, 2, 2, 2, 1};
  double s[6], tau[3][3];
#pragma unroll
  for (int m = 0; m < 6; m++) {
    const int j = vj[m], k = vk[m];
    s[m] = g[j][k] + g[k][j] + g[j][0] * g[k][0] + g[j][1] * g[k][1] + g[j][2] * g[k][2];
  }
  // cofactors of F first: det F shares ONE reciprocal with the series variable y = xr / (2 + xr)
  const double F00 = g[0][0] + 1., F11 = g[1][1] + 1., F22 = g[2][2] + 1.;
  double cof[3][3];
  cof[0][0] = F11 * F22 - g[1][2] * g[2][1];
  cof[0][1] = g[1][2] * g[2][0] - g[1][0] * F22;
  cof[0][2] = g[1][0] * g[2][1] - F11 * g[2][0];
  cof[1][0] = g[0][2] * g[2][1] - g[0][1] * F22;
  cof[1][1] = F00 * F22 - g[0][2] * g[2][0];
  cof[1][2] = g[0][1] * g[2][0] - F00 * g[2][1];
  cof[2][0] = g[0][1] * g[1][2] - g[0][2] * F11;
  cof[2][1] = g[0][2] * g[1][0] - F00 * g[1][2];
  cof[2][2] = F00 * F11 - g[0][1] * g[1][0];
  const double detF = F00 * cof[0][0] + g[0][1] * cof[0][1] + g[0][2] * cof[0][2];
  double xr;
  const double shift = log1p_shift(det_sym_m1(s), xr);
  const double den = 2. + xr;
  const double rboth = 1. / (den * detF);
  const double llnj = mt.lambda * (shift + 0.5 * log1p_series_in_y(xr * (rboth * detF)));
  const double wr = w * (rboth * den);
  const double tv[6] = {mt.mu * s[0] + llnj, mt.mu * s[1] + llnj, mt.mu * s[2] + llnj, mt.mu * s[3], mt.mu * s[4], mt.mu * s[5]};
  voigt_sym(tv, tau);
  // F^-T = cof(F) / det F, scaled by w:  Kt[n][k] = w / detF * sum_m cof[n][m] A[k][m];   W = tau Kt
  double Kt[3][3];
#pragma unroll
  for (int n = 0; n < 3; n++)
#pragma unroll
    for (int k = 0; k < 3; k++) Kt[n][k] = wr * (cof[n][0] * A[k][0] + cof[n][1] * A[k][1] + cof[n][2] * A[k][2]);
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int k = 0; k < 3; k++) W[c][k] = tau[c][0] * Kt[0][k] + tau[c][1] * Kt[1][k] + tau[c][2] * Kt[2][k];
}

// Jacobian in the reference's own algebra (hyperFS.h:339-459), used by the generic
// (non-fused) operator path and to cross-check the cached form on the device.
B200_DI void hyperfs_df_point_faithful(const Material &mt, double w, const double (&A)[3][3],
                                       const double (&g)[3][3], const double (&H)[3][3],
                                       double (&W)[3][3]) {
  double S[3][3], Ci[3][3], llnj, F[3][3], gd[3][3], dE[3][3], dECi[3][3], dS[3][3], dP[3][3];
  phys_grad(A, H, gd);
  fs_common(mt, g, S, Ci, llnj);
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) F[a][b] = g[a][b] + (a == b ? 1. : 0.);
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) {
      double s = 0;
#pragma unroll
      for (int n = 0; n < 3; n++) s += (gd[n][a] * F[n][b] + F[n][a] * gd[n][b]) / 2.;
      dE[a][b] = s;
    }
  double CiE = 0;
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) CiE += Ci[a][b] * dE[a][b];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) dECi[a][b] = dE[a][0] * Ci[0][b] + dE[a][1] * Ci[1][b] + dE[a][2] * Ci[2][b];
  const double llnj_m = llnj - mt.mu;
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++)
      dS[a][b] = mt.lambda * CiE * Ci[a][b] -
                 2. * llnj_m * (Ci[a][0] * dECi[0][b] + Ci[a][1] * dECi[1][b] + Ci[a][2] * dECi[2][b]);
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) {
      double s = 0;
#pragma unroll
      for (int m = 0; m < 3; m++) s += gd[a][m] * S[m][b] + F[a][m] * dS[m][b];
      dP[a][b] = s * w;
    }
  pull_back(A, dP, W);
}

// ------------------------------------------------------------------ post-processing
// Strain energy density and nodal diagnostics (one-shot operators, setuplibceed.c:650-737):
//   p[0] pressure, p[1] first strain invariant, p[2] second invariant, p[3] volume ratio, p[4] energy density
// linElas.h:285-478, hyperSS.h:326-528, hyperFS.h:469-668.  The small-strain energies keep the reference's
// `strain_vol*mu` term as written.
template <int PROB>
B200_DI void post_point(const Material &mt, const double (&A)[3][3], const double (&H)[3][3], double (&p)[5]) {
  double g[3][3];
  phys_grad(A, H, g);
  if (PROB == B200_PROB_HYPERFS) {
    double e[6];
    const double detC_m1 = green_lagrange2(g, e);  // e = 2E, Voigt (00,11,22,12,02,01)
    const double logj = log1p_series_shifted(detC_m1) / 2.;
    const double tr2 = e[0] + e[1] + e[2];
    p[0] = -mt.lambda * logj;
    p[1] = tr2 / 2.;
    p[2] = (e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + 2. * (e[3] * e[3] + e[4] * e[4] + e[5] * e[5])) / 4.;
    p[3] = sqrt(detC_m1 + 1.);
    p[4] = mt.lambda * logj * logj / 2. - mt.mu * logj + mt.mu * tr2 / 2.;
  } else {
    const double e01 = (g[0][1] + g[1][0]) / 2., e02 = (g[0][2] + g[2][0]) / 2., e12 = (g[1][2] + g[2][1]) / 2.;
    const double tr = g[0][0] + g[1][1] + g[2][2];
    const double shear = (e01 * e01 + e02 * e02 + e12 * e12) * 2. * mt.mu;
    if (PROB == B200_PROB_HYPERSS) {
      const double llv = log1p_series(tr);
      p[0] = -mt.lambda * llv;
      p[4] = mt.lambda * (1. + tr) * (llv - 1.) + tr * mt.mu + shear;
    } else {
      p[0] = -mt.lambda * tr;
      p[4] = mt.lambda * tr * tr / 2. + tr * mt.mu + shear;
    }
    p[1] = tr;
    p[2] = g[0][0] * g[0][0] + g[1][1] * g[1][1] + g[2][2] * g[2][2] + 2. * (e01 * e01 + e02 * e02 + e12 * e12);
    p[3] = 1. + tr;
  }
}

// ------------------------------------------------------------------ Jacobian cache
// Per quadrature point, the Jacobian action in its cheapest exact algebraic form.  Every Jacobian here has the
// shape  W = w L(H A) A^T  with L linear, so the weight is folded into the geometry: A' = sqrt(w) A gives
// W = L(H A') A'^T and one double less to stream per point (w = weight x det J > 0).
//
//   linElas  ( 9): A'                         sigma(H A') A'^T, as the reference
//   hyperSS  (10): A', s = 1/(1+tr gradu)     dsigma = lambda s tr(g) I + mu (g + g^T)
//   hyperFS  (16): K' = sqrt(w) A F^-1, b = F F^T (Voigt), lnJ
//       With gt = H K (spatial gradient of the increment) the reference's
//       dP = grad(du) S + F dS, S = mu I + (lambda lnJ - mu) C^-1, collapses to
//         W = w [ mu gt b + lambda tr(gt) I + (mu - lambda lnJ) gt^T ] K^T
//       (the two (lambda lnJ - mu) gt K K^T terms cancel), ~105 FP64 ops instead of ~580.
//       Material constants are NOT baked in, so GetDiag_Ceed's smoother-context swap
//       (matops.c:215-217) keeps working.
template <int PROB> struct JCache;
template <> struct JCache<B200_PROB_LINELAS> { static constexpr int N = 9; };
template <> struct JCache<B200_PROB_HYPERSS> { static constexpr int N = 10; };
template <> struct JCache<B200_PROB_HYPERFS> { static constexpr int N = 16; };

// qd[10] = qdata, gu[9] = gradu [c][k]  ->  jc[N]
template <int PROB>
B200_DI void jcache_point(const double *qd, const double *gu, double *jc) {
  const double sw = sqrt(qd[0]);
  if (PROB == B200_PROB_LINELAS) {
#pragma unroll
    for (int i = 0; i < 9; i++) jc[i] = sw * qd[1 + i];
  } else if (PROB == B200_PROB_HYPERSS) {
#pragma unroll
    for (int i = 0; i < 9; i++) jc[i] = sw * qd[1 + i];
    jc[9] = 1. / (1. + (gu[0] + gu[4] + gu[8]));
  } else {
    double g[3][3], F[3][3], Fi[3][3], e[6];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        g[c][k] = gu[c * 3 + k];
        F[c][k] = g[c][k] + (c == k ? 1. : 0.);
      }
    const double c00 = F[1][1] * F[2][2] - F[1][2] * F[2][1];
    const double c01 = F[1][2] * F[2][0] - F[1][0] * F[2][2];
    const double c02 = F[1][0] * F[2][1] - F[1][1] * F[2][0];
    const double rdet = sw / (F[0][0] * c00 + F[0][1] * c01 + F[0][2] * c02);  // sqrt(w) folded into F^-1
    Fi[0][0] = c00 * rdet;
    Fi[1][0] = c01 * rdet;
    Fi[2][0] = c02 * rdet;
    Fi[0][1] = (F[0][2] * F[2][1] - F[0][1] * F[2][2]) * rdet;
    Fi[1][1] = (F[0][0] * F[2][2] - F[0][2] * F[2][0]) * rdet;
    Fi[2][1] = (F[0][1] * F[2][0] - F[0][0] * F[2][1]) * rdet;
    Fi[0][2] = (F[0][1] * F[1][2] - F[0][2] * F[1][1]) * rdet;
    Fi[1][2] = (F[0][2] * F[1][0] - F[0][0] * F[1][2]) * rdet;
    Fi[2][2] = (F[0][0] * F[1][1] - F[0][1] * F[1][0]) * rdet;
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
      for (int j = 0; j < 3; j++)
        jc[3 * k + j] = qd[1 + 3 * k] * Fi[0][j] + qd[2 + 3 * k] * Fi[1][j] + qd[3 + 3 * k] * Fi[2][j];
    const int vj[6] = {0, 1, 2, 1, 0, 0}, vk[6] = {0, 1, 2, 2, 2, 1};
#pragma unroll
    for (int m = 0; m < 6; m++)
      jc[9 + m] = F[vj[m]][0] * F[vk[m]][0] + F[vj[m]][1] * F[vk[m]][1] + F[vj[m]][2] * F[vk[m]][2];
    jc[15] = log1p_series_shifted(green_lagrange2(g, e)) / 2.;
  }
}

// Jacobian action from the cache:  W = J_q(H)
template <int PROB>
B200_DI void jacobian_point(const Material &mt, const double *jc, const double (&H)[3][3],
                            double (&W)[3][3]) {
  double A[3][3];
#pragma unroll
  for (int m = 0; m < 3; m++)
#pragma unroll
    for (int k = 0; k < 3; k++) A[m][k] = jc[3 * m + k];
  if (PROB == B200_PROB_LINELAS) {
    linelas_point(mt, 1., A, H, W);
  } else if (PROB == B200_PROB_HYPERSS) {
    hyperss_df_point(mt, 1., A, jc[9], H, W);
  } else {
    // A holds K' here
#ifdef B200_OLD_JPOINT
    double gt[3][3], Z[3][3], bm[3][3];
    const double bv[6] = {jc[9], jc[10], jc[11], jc[12], jc[13], jc[14]};
    voigt_sym(bv, bm);
    phys_grad(A, H, gt);
    const double cw = mt.mu, bw = mt.mu - mt.lambda * jc[15];
    const double aw = mt.lambda * (gt[0][0] + gt[1][1] + gt[2][2]);
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int j = 0; j < 3; j++)
        Z[c][j] = cw * (gt[c][0] * bm[0][j] + gt[c][1] * bm[1][j] + gt[c][2] * bm[2][j]) + bw * gt[j][c] +
                  (c == j ? aw : 0.);
    pull_back(A, Z, W);
#else
    double gt[3][3], Z[3][3], bm[3][3];
    // mu folded into b once (6 multiplies) so that every entry of Z is one multiply-or-fma start + 3 fma
    const double bv[6] = {mt.mu * jc[9], mt.mu * jc[10], mt.mu * jc[11], mt.mu * jc[12], mt.mu * jc[13], mt.mu * jc[14]};
    voigt_sym(bv, bm);
    phys_grad(A, H, gt);  // gt[c][j] = sum_m H[c][m] K'[m][j]
    const double bw = mt.mu - mt.lambda * jc[15];
    const double aw = mt.lambda * (gt[0][0] + gt[1][1] + gt[2][2]);
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        double z = c == j ? fma(bw, gt[c][c], aw) : bw * gt[j][c];
        z = fma(gt[c][0], bm[0][j], z);
        z = fma(gt[c][1], bm[1][j], z);
        Z[c][j] = fma(gt[c][2], bm[2][j], z);
      }
    pull_back(A, Z, W);  // W[c][k] = sum_m K'[k][m] Z[c][m]
#endif
  }
}

// Diagonal point blocks in closed form.  Every Jacobian of this path has the shape
//     W = [ m1 gt b + k1 gt^T + k2 tr(gt) I ] K^T,   gt = H K      (K = cached geometry, b symmetric)
// so the 3x3 block of component c, A_c[d'][d] = dW[c][d] / dH[c][d'], is
//     A_c = m1 K b K^T + (k1 + k2) k_c k_c^T,     k_c[d] = K[d][c]
// (symmetric; the first term is the same for all three components).  Returns M = m1 K b K^T in Voigt order
// (00,11,22,12,02,01) and kappa = k1 + k2:
//   linElas  m1 = c3, b = I, kappa = c1 - c3      (the reference's shear factor, linElas.h:137-139, included)
//   hyperSS  m1 = mu, b = I, kappa = mu + lambda s
//   hyperFS  m1 = mu, b = F F^T, kappa = lambda + mu - lambda ln J
// ~70 FP64 operations per point for all three components, instead of 9 unit inputs through jacobian_point.
template <int PROB>
B200_DI void diag_blocks_point(const Material &mt, const double *jc, double (&M)[6], double &kappa) {
  const int vj[6] = {0, 1, 2, 1, 0, 0}, vk[6] = {0, 1, 2, 2, 2, 1};
  double Kb[3][3];
  if (PROB == B200_PROB_HYPERFS) {
    double bm[3][3];
    const double bv[6] = {jc[9], jc[10], jc[11], jc[12], jc[13], jc[14]};
    voigt_sym(bv, bm);
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
      for (int j = 0; j < 3; j++) Kb[d][j] = jc[3 * d] * bm[0][j] + jc[3 * d + 1] * bm[1][j] + jc[3 * d + 2] * bm[2][j];
    kappa = mt.lambda + mt.mu - mt.lambda * jc[15];
  } else {
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
      for (int j = 0; j < 3; j++) Kb[d][j] = jc[3 * d + j];
    kappa = PROB == B200_PROB_HYPERSS ? mt.mu + mt.lambda * jc[9] : mt.le_c1 - mt.le_c3;
  }
  const double m1 = PROB == B200_PROB_LINELAS ? mt.le_c3 : mt.mu;
#pragma unroll
  for (int t = 0; t < 6; t++) {
    const int d = vj[t], e = vk[t];
    M[t] = m1 * (Kb[d][0] * jc[3 * e] + Kb[d][1] * jc[3 * e + 1] + Kb[d][2] * jc[3 * e + 2]);
  }
}

}  // namespace b200
