/* TEST INFRASTRUCTURE ONLY -- see oracle/README.md.
 *
 * CPU restatement ("port") of the reference's hot-path QFunctions, written from
 * the mathematics so that the oracle is self-contained on a box where
 * /root/reference (and a prebuilt oracle/_ref/libref_qf.so) is absent.  It is
 * pinned against the REAL reference functions (oracle/_ref, built from
 * /root/reference/qfunctions/ by oracle/Makefile) and against the committed
 * golden vectors in tests/golden/ by tests/test_oracle_qfunctions.py.
 *
 * Reference sources restated here:
 *   SetupGeo   qfunctions/common.h:47-101
 *   LinElasF   qfunctions/linElas.h:39-158     LinElasdF  qfunctions/linElas.h:163-280
 *   HyperSSF   qfunctions/hyperSS.h:60-182     HyperSSdF  qfunctions/hyperSS.h:187-321
 *   HyperFSF   qfunctions/hyperFS.h:147-281    HyperFSdF  qfunctions/hyperFS.h:286-464
 *   (helpers: log1p_series hyperSS.h:43-55, log1p_series_shifted hyperFS.h:45-67,
 *    computeDetCM1 hyperFS.h:72-80, commonFS hyperFS.h:85-142)
 *
 * Array conventions: a GRAD field is [deriv d][comp c][Q]
 * (d slowest); qdata is [10][Q] = {w*detJ, dXdx row-major}; stored gradu is
 * [comp][deriv][Q].  All functions have the libCEED user-QFunction signature.
 */
#include <math.h>

typedef struct {
  double nu, E;
} PortPhysics; /* elasticity.h:30-37 Physics_private */

typedef double M3[3][3];

/* Voigt ordering used by the reference: (00,11,22,12,02,01) */
static const int VJ[6] = {0, 1, 2, 1, 0, 0}, VK[6] = {0, 1, 2, 2, 2, 1};

static inline void voigt_to_sym(const double v[6], M3 m) {
  m[0][0] = v[0]; m[1][1] = v[1]; m[2][2] = v[2];
  m[1][2] = m[2][1] = v[3];
  m[0][2] = m[2][0] = v[4];
  m[0][1] = m[1][0] = v[5];
}

/* physical gradient  g[c][k] = sum_m dXdx[m][k] * du_ref[c][m]   with the GRAD
 * input laid out as in[(m*3 + c)*Q + i] */
static inline void phys_grad(const double *ug, const M3 dXdx, int Q, int i, M3 g) {
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < 3; k++) {
      double s = 0;
      for (int m = 0; m < 3; m++) s += dXdx[m][k] * ug[(m * 3 + c) * Q + i];
      g[c][k] = s;
    }
}

/* out[(k*3 + c)*Q + i] = sum_m dXdx[k][m] * T[c][m] * wdetJ */
static inline void pull_back(const M3 T, const M3 dXdx, double wdetJ, int Q, int i,
                             double *out) {
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < 3; k++) {
      double s = 0;
      for (int m = 0; m < 3; m++) s += dXdx[k][m] * T[c][m] * wdetJ;
      out[(k * 3 + c) * Q + i] = s;
    }
}

static inline void load_qdata(const double *qd, int Q, int i, double *wdetJ, M3 dXdx) {
  *wdetJ = qd[i];
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) dXdx[a][b] = qd[(1 + 3 * a + b) * Q + i];
}

/* ---------------------------------------------------------------- SetupGeo */
int port_SetupGeo(void *ctx, int Q, const double *const *in, double *const *out) {
  (void)ctx;
  const double *J = in[0], *w = in[1];
  double *qd = out[0];
  for (int i = 0; i < Q; i++) {
    /* J[d][c] = d x_c / d X_d  at in[(d*3+c)*Q+i]; the reference names J(c+1)(d+1) */
    double Jm[3][3]; /* Jm[c][d] */
    for (int d = 0; d < 3; d++)
      for (int c = 0; c < 3; c++) Jm[c][d] = J[(d * 3 + c) * Q + i];
    double A[3][3];
    A[0][0] = Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1];
    A[0][1] = Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2];
    A[0][2] = Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1];
    A[1][0] = Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2];
    A[1][1] = Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0];
    A[1][2] = Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2];
    A[2][0] = Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0];
    A[2][1] = Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1];
    A[2][2] = Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0];
    const double detJ = Jm[0][0] * A[0][0] + Jm[1][0] * A[0][1] + Jm[2][0] * A[0][2];
    qd[i] = w[i] * detJ;
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) qd[(1 + 3 * a + b) * Q + i] = A[a][b] / detJ;
  }
  return 0;
}

/* ---------------------------------------------------------------- linElas */
static inline void linelas_stress(const M3 g, double E, double nu, M3 sig) {
  double e[3][3];
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) e[a][b] = (g[a][b] + g[b][a]) / 2.;
  const double ss = E / ((1 + nu) * (1 - 2 * nu));
  sig[0][0] = ss * ((1 - nu) * e[0][0] + nu * e[1][1] + nu * e[2][2]);
  sig[1][1] = ss * (nu * e[0][0] + (1 - nu) * e[1][1] + nu * e[2][2]);
  sig[2][2] = ss * (nu * e[0][0] + nu * e[1][1] + (1 - nu) * e[2][2]);
  /* the reference's shear term carries an extra 0.5 (linElas.h:137-139); kept */
  sig[1][2] = sig[2][1] = ss * (1 - 2 * nu) * e[1][2] * 0.5;
  sig[0][2] = sig[2][0] = ss * (1 - 2 * nu) * e[0][2] * 0.5;
  sig[0][1] = sig[1][0] = ss * (1 - 2 * nu) * e[0][1] * 0.5;
}

static int linelas_common(void *ctx, int Q, const double *const *in, double *const *out) {
  const PortPhysics *p = (const PortPhysics *)ctx;
  for (int i = 0; i < Q; i++) {
    double wdetJ;
    M3 dXdx, g, sig;
    load_qdata(in[1], Q, i, &wdetJ, dXdx);
    phys_grad(in[0], dXdx, Q, i, g);
    linelas_stress(g, p->E, p->nu, sig);
    pull_back(sig, dXdx, wdetJ, Q, i, out[0]);
  }
  return 0;
}
int port_LinElasF(void *ctx, int Q, const double *const *in, double *const *out) {
  return linelas_common(ctx, Q, in, out);
}
int port_LinElasdF(void *ctx, int Q, const double *const *in, double *const *out) {
  return linelas_common(ctx, Q, in, out); /* linear: Jacobian action == residual action */
}

/* ---------------------------------------------------------------- hyperSS */
static inline double series_log1p(double x) {
  double y = x / (2. + x);
  const double y2 = y * y;
  double sum = y;
  y *= y2; sum += y / 3;
  y *= y2; sum += y / 5;
  y *= y2; sum += y / 7;
  return 2 * sum;
}

static inline void lame(const PortPhysics *p, double *TwoMu, double *mu, double *lambda) {
  *TwoMu = p->E / (1 + p->nu);
  *mu = *TwoMu / 2;
  const double Kbulk = p->E / (3 * (1 - 2 * p->nu));
  *lambda = (3 * Kbulk - *TwoMu) / 3;
}

int port_HyperSSF(void *ctx, int Q, const double *const *in, double *const *out) {
  double TwoMu, mu, lambda;
  lame((const PortPhysics *)ctx, &TwoMu, &mu, &lambda);
  double *gradu = out[1];
  for (int i = 0; i < Q; i++) {
    double wdetJ;
    M3 dXdx, g, sig;
    load_qdata(in[1], Q, i, &wdetJ, dXdx);
    phys_grad(in[0], dXdx, Q, i, g);
    for (int c = 0; c < 3; c++)
      for (int k = 0; k < 3; k++) gradu[(c * 3 + k) * Q + i] = g[c][k];
    double e[3][3];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) e[a][b] = (g[a][b] + g[b][a]) / 2.;
    const double llv = series_log1p(e[0][0] + e[1][1] + e[2][2]);
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) sig[a][b] = (a == b ? lambda * llv : 0.) + TwoMu * e[a][b];
    pull_back(sig, dXdx, wdetJ, Q, i, out[0]);
  }
  return 0;
}

int port_HyperSSdF(void *ctx, int Q, const double *const *in, double *const *out) {
  double TwoMu, mu, lambda;
  lame((const PortPhysics *)ctx, &TwoMu, &mu, &lambda);
  const double *gradu = in[2];
  for (int i = 0; i < Q; i++) {
    double wdetJ;
    M3 dXdx, g, dsig;
    load_qdata(in[1], Q, i, &wdetJ, dXdx);
    phys_grad(in[0], dXdx, Q, i, g);
    double de[3][3];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) de[a][b] = (g[a][b] + g[b][a]) / 2.;
    const double strain_vol = gradu[0 * Q + i] + gradu[4 * Q + i] + gradu[8 * Q + i];
    const double lambda_bar = lambda / (1 + strain_vol);
    const double ldt = lambda_bar * (de[0][0] + de[1][1] + de[2][2]);
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) dsig[a][b] = (a == b ? ldt : 0.) + TwoMu * de[a][b];
    pull_back(dsig, dXdx, wdetJ, Q, i, out[0]);
  }
  return 0;
}

/* ---------------------------------------------------------------- hyperFS */
static inline double series_log1p_shifted(double x) {
  const double left = sqrt(2.) / 2 - 1, right = sqrt(2.) - 1;
  double sum = 0;
  if (x < left) {
    sum -= log(2.) / 2;
    x = 1 + 2 * x;
  } else if (right < x) {
    sum += log(2.) / 2;
    x = (x - 1) / 2;
  }
  double y = x / (2. + x);
  const double y2 = y * y;
  sum += y;
  y *= y2; sum += y / 3;
  y *= y2; sum += y / 5;
  y *= y2; sum += y / 7;
  return 2 * sum;
}

/* det(C) - 1 from 2E in Voigt form, cancellation-free polynomial */
static inline double det_c_minus_1(const double e[6]) {
  return e[0] * (e[1] * e[2] - e[3] * e[3]) + e[5] * (e[4] * e[3] - e[5] * e[2]) +
         e[4] * (e[5] * e[3] - e[4] * e[1]) + e[0] + e[1] + e[2] + e[0] * e[1] +
         e[0] * e[2] + e[1] * e[2] - e[5] * e[5] - e[4] * e[4] - e[3] * e[3];
}

/* S (2nd Piola-Kirchhoff, Voigt), C^-1 (Voigt), lambda*log(J) from grad u */
static void fs_kinematics(double lambda, double mu, const M3 g, double Sv[6],
                          double Civ[6], double *llnj) {
  double e2v[6];
  for (int m = 0; m < 6; m++) {
    const int j = VJ[m], k = VK[m];
    double s = g[j][k] + g[k][j];
    for (int n = 0; n < 3; n++) s += g[n][j] * g[n][k];
    e2v[m] = s;
  }
  M3 E2, C;
  voigt_to_sym(e2v, E2);
  const double detC_m1 = det_c_minus_1(e2v);
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) C[a][b] = E2[a][b] + (a == b ? 1. : 0.);
  const double A[6] = {C[1][1] * C[2][2] - C[1][2] * C[2][1],
                       C[0][0] * C[2][2] - C[0][2] * C[2][0],
                       C[0][0] * C[1][1] - C[0][1] * C[1][0],
                       C[0][2] * C[1][0] - C[0][0] * C[1][2],
                       C[0][1] * C[1][2] - C[0][2] * C[1][1],
                       C[0][2] * C[2][1] - C[0][1] * C[2][2]};
  for (int m = 0; m < 6; m++) Civ[m] = A[m] / (detC_m1 + 1.);
  M3 Ci;
  voigt_to_sym(Civ, Ci);
  *llnj = lambda * series_log1p_shifted(detC_m1) / 2.;
  for (int m = 0; m < 6; m++) {
    double s = (*llnj) * Civ[m];
    for (int n = 0; n < 3; n++) s += mu * Ci[VJ[m]][n] * E2[n][VK[m]];
    Sv[m] = s;
  }
}

int port_HyperFSF(void *ctx, int Q, const double *const *in, double *const *out) {
  double TwoMu, mu, lambda;
  lame((const PortPhysics *)ctx, &TwoMu, &mu, &lambda);
  double *gradu = out[1];
  for (int i = 0; i < Q; i++) {
    double wdetJ, Sv[6], Civ[6], llnj;
    M3 dXdx, g, F, S, P;
    load_qdata(in[1], Q, i, &wdetJ, dXdx);
    phys_grad(in[0], dXdx, Q, i, g);
    for (int c = 0; c < 3; c++)
      for (int k = 0; k < 3; k++) {
        gradu[(c * 3 + k) * Q + i] = g[c][k];
        F[c][k] = g[c][k] + (c == k ? 1. : 0.);
      }
    fs_kinematics(lambda, mu, g, Sv, Civ, &llnj);
    voigt_to_sym(Sv, S);
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {
        double s = 0;
        for (int m = 0; m < 3; m++) s += F[a][m] * S[m][b];
        P[a][b] = s;
      }
    pull_back(P, dXdx, wdetJ, Q, i, out[0]);
  }
  return 0;
}

int port_HyperFSdF(void *ctx, int Q, const double *const *in, double *const *out) {
  double TwoMu, mu, lambda;
  lame((const PortPhysics *)ctx, &TwoMu, &mu, &lambda);
  const double *gradu = in[2];
  for (int i = 0; i < Q; i++) {
    double wdetJ, Sv[6], Civ[6], llnj, dEv[6];
    M3 dXdx, gd, g, F, S, Ci, dE, dECi, dS, dP;
    load_qdata(in[1], Q, i, &wdetJ, dXdx);
    phys_grad(in[0], dXdx, Q, i, gd);
    for (int c = 0; c < 3; c++)
      for (int k = 0; k < 3; k++) {
        g[c][k] = gradu[(c * 3 + k) * Q + i];
        F[c][k] = g[c][k] + (c == k ? 1. : 0.);
      }
    fs_kinematics(lambda, mu, g, Sv, Civ, &llnj);
    voigt_to_sym(Sv, S);
    voigt_to_sym(Civ, Ci);
    for (int m = 0; m < 6; m++) {
      double s = 0;
      for (int n = 0; n < 3; n++)
        s += (gd[n][VJ[m]] * F[n][VK[m]] + F[n][VJ[m]] * gd[n][VK[m]]) / 2.;
      dEv[m] = s;
    }
    voigt_to_sym(dEv, dE);
    double CiE = 0;
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) CiE += Ci[a][b] * dE[a][b];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {
        double s = 0;
        for (int m = 0; m < 3; m++) s += dE[a][m] * Ci[m][b];
        dECi[a][b] = s;
      }
    const double llnj_m = llnj - mu;
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {
        double s = 0;
        for (int m = 0; m < 3; m++) s += Ci[a][m] * dECi[m][b];
        dS[a][b] = lambda * CiE * Ci[a][b] - 2. * llnj_m * s;
      }
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {
        double s = 0;
        for (int m = 0; m < 3; m++) s += gd[a][m] * S[m][b] + F[a][m] * dS[m][b];
        dP[a][b] = s;
      }
    pull_back(dP, dXdx, wdetJ, Q, i, out[0]);
  }
  return 0;
}

/* ---------------------------------------------------------------- forcing / MMS */
/* SetupConstantForce, qfunctions/constantForce.h:39-70: in = {x[3] (unused), qdata[10]}, ctx = vector[3] */
int port_SetupConstantForce(void *ctx, int Q, const double *const *in, double *const *out) {
  const double *v = (const double *)ctx, *qd = in[1];
  for (int i = 0; i < Q; i++)
    for (int c = 0; c < 3; c++) out[0][c * Q + i] = v[c] * qd[i];
  return 0;
}

/* manufactured solution u = (e^2x sin3y cos4z, e^3y sin4z cos2x, e^4z sin2x cos3y) / 1e8,
 * qfunctions/manufacturedTrue.h:30-56 */
int port_MMSTrueSoln(void *ctx, int Q, const double *const *in, double *const *out) {
  (void)ctx;
  for (int i = 0; i < Q; i++) {
    const double x = in[0][i], y = in[0][Q + i], z = in[0][2 * Q + i];
    out[0][i] = exp(2 * x) * sin(3 * y) * cos(4 * z) / 1e8;
    out[0][Q + i] = exp(3 * y) * sin(4 * z) * cos(2 * x) / 1e8;
    out[0][2 * Q + i] = exp(4 * z) * sin(2 * x) * cos(3 * y) / 1e8;
  }
  return 0;
}

/* SetupMMSForce, qfunctions/manufacturedForce.h:39-103, restated from the mathematics:
 * f = -div sigma(u) * w detJ with the reference's linear-elastic stress law (linElas.h:127-139,
 * sigma_ii = lambda tr(e) + 2 mu e_ii, sigma_ij = mu e_ij) and the closed-form second derivatives of u. */
int port_SetupMMSForce(void *ctx, int Q, const double *const *in, double *const *out) {
  const PortPhysics *p = (const PortPhysics *)ctx;
  const double lam = p->E * p->nu / ((1 + p->nu) * (1 - 2 * p->nu)), mu = p->E / (2 * (1 + p->nu));
  for (int i = 0; i < Q; i++) {
    const double x = in[0][i], y = in[0][Q + i], z = in[0][2 * Q + i], w = in[1][i] / 1e8;
    const double ex = exp(2 * x), ey = exp(3 * y), ez = exp(4 * z);
    const double s2x = sin(2 * x), c2x = cos(2 * x), s3y = sin(3 * y), c3y = cos(3 * y), s4z = sin(4 * z), c4z = cos(4 * z);
    const double u1 = ex * s3y * c4z, u2 = ey * s4z * c2x, u3 = ez * s2x * c3y;
    const double u1xx = 4 * u1, u1yy = -9 * u1, u1zz = -16 * u1, u1xy = 6 * ex * c3y * c4z, u1xz = -8 * ex * s3y * s4z;
    const double u2yy = 9 * u2, u2xx = -4 * u2, u2zz = -16 * u2, u2xy = -6 * ey * s4z * s2x, u2yz = 12 * ey * c4z * c2x;
    const double u3zz = 16 * u3, u3xx = -4 * u3, u3yy = -9 * u3, u3xz = 8 * ez * c2x * c3y, u3yz = -12 * ez * s2x * s3y;
    out[0][i] = -(lam * (u1xx + u2xy + u3xz) + 2 * mu * u1xx + 0.5 * mu * (u1yy + u2xy + u1zz + u3xz)) * w;
    out[0][Q + i] = -(lam * (u1xy + u2yy + u3yz) + 2 * mu * u2yy + 0.5 * mu * (u2xx + u1xy + u2zz + u3yz)) * w;
    out[0][2 * Q + i] = -(lam * (u1xz + u2yz + u3zz) + 2 * mu * u3zz + 0.5 * mu * (u3xx + u1xz + u3yy + u2yz)) * w;
  }
  return 0;
}

/* ---------------------------------------------------------------- strain energy and nodal diagnostics
 * (one-shot post-processing operators, setuplibceed.c:650-737).  Point quantities
 *   p[0] pressure, p[1] first strain invariant, p[2] second invariant, p[3] volume ratio, p[4] energy density
 * restated from linElas.h:285-478, hyperSS.h:326-528, hyperFS.h:469-668; the small-strain energy keeps the
 * reference's `strain_vol*mu` term as written. */
static void post_small_strain(int hyper, double lambda, double mu, const M3 g, double p[5]) {
  M3 e;
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) e[a][b] = (g[a][b] + g[b][a]) / 2.;
  const double tr = e[0][0] + e[1][1] + e[2][2];
  const double shear = (e[0][1] * e[0][1] + e[0][2] * e[0][2] + e[1][2] * e[1][2]) * 2 * mu;
  double ee = 0;
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) ee += e[a][b] * e[b][a];
  if (hyper) {
    const double llv = series_log1p(tr);
    p[0] = -lambda * llv;
    p[4] = lambda * (1 + tr) * (llv - 1) + tr * mu + shear;
  } else {
    p[0] = -lambda * tr;
    p[4] = lambda * tr * tr / 2. + tr * mu + shear;
  }
  p[1] = tr;
  p[2] = ee;
  p[3] = 1 + tr;
}

static void post_finite_strain(double lambda, double mu, const M3 g, double p[5]) {
  double e2v[6];
  for (int m = 0; m < 6; m++) {
    const int j = VJ[m], k = VK[m];
    double s = g[j][k] + g[k][j];
    for (int n = 0; n < 3; n++) s += g[n][j] * g[n][k];
    e2v[m] = s;
  }
  M3 E2;
  voigt_to_sym(e2v, E2);
  const double detC_m1 = det_c_minus_1(e2v);
  const double logj = series_log1p_shifted(detC_m1) / 2.;
  const double trE2 = E2[0][0] + E2[1][1] + E2[2][2];
  double ee = 0;
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) ee += E2[a][b] * E2[b][a] / 4.;
  p[0] = -lambda * logj;
  p[1] = trE2 / 2.;
  p[2] = ee;
  p[3] = sqrt(detC_m1 + 1);
  p[4] = lambda * logj * logj / 2. - mu * logj + mu * trE2 / 2.;
}

/* kind: 0 linElas, 1 hyperSS, 2 hyperFS;  du = GRAD input, qd = qdata */
static void post_point(int kind, const PortPhysics *ph, const double *du, const double *qd, int Q, int i,
                       double *wdetJ, double p[5]) {
  double TwoMu, mu, lambda;
  lame(ph, &TwoMu, &mu, &lambda);
  M3 dXdx, g;
  load_qdata(qd, Q, i, wdetJ, dXdx);
  phys_grad(du, dXdx, Q, i, g);
  if (kind == 2) post_finite_strain(lambda, mu, g, p);
  else post_small_strain(kind, lambda, mu, g, p);
}

static int energy_common(int kind, void *ctx, int Q, const double *const *in, double *const *out) {
  for (int i = 0; i < Q; i++) {
    double w, p[5];
    post_point(kind, (const PortPhysics *)ctx, in[0], in[1], Q, i, &w, p);
    out[0][i] = p[4] * w;
  }
  return 0;
}

static int diagnostic_common(int kind, void *ctx, int Q, const double *const *in, double *const *out) {
  for (int i = 0; i < Q; i++) {
    double w, p[5];
    post_point(kind, (const PortPhysics *)ctx, in[1], in[2], Q, i, &w, p);
    for (int c = 0; c < 3; c++) out[0][c * Q + i] = in[0][c * Q + i];
    for (int c = 0; c < 5; c++) out[0][(3 + c) * Q + i] = p[c];
  }
  return 0;
}

int port_LinElasEnergy(void *ctx, int Q, const double *const *in, double *const *out) { return energy_common(0, ctx, Q, in, out); }
int port_HyperSSEnergy(void *ctx, int Q, const double *const *in, double *const *out) { return energy_common(1, ctx, Q, in, out); }
int port_HyperFSEnergy(void *ctx, int Q, const double *const *in, double *const *out) { return energy_common(2, ctx, Q, in, out); }
int port_LinElasDiagnostic(void *ctx, int Q, const double *const *in, double *const *out) { return diagnostic_common(0, ctx, Q, in, out); }
int port_HyperSSDiagnostic(void *ctx, int Q, const double *const *in, double *const *out) { return diagnostic_common(1, ctx, Q, in, out); }
int port_HyperFSDiagnostic(void *ctx, int Q, const double *const *in, double *const *out) { return diagnostic_common(2, ctx, Q, in, out); }
