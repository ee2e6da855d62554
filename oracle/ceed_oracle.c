/* TEST INFRASTRUCTURE ONLY -- see oracle/README.md.
 *
 * CPU oracle: a plain-C restatement of what the reference's hot path executes
 * when it runs with `-ceed /cpu/self`:
 *
 *   ApplyLocalCeedOp            /root/reference/src/matops.c:26-60
 *     -> CeedOperatorApply      (libCEED, NOT vendored in the reference; see below)
 *   GetDiag_Ceed                /root/reference/src/matops.c:206-244
 *     -> CeedOperatorLinearAssembleDiagonal
 *   operator wiring             /root/reference/src/setuplibceed.c:370-393 (SetupGeo),
 *                               :518-542 (residual), :818-839 (Jacobian)
 *
 * The arithmetic below CeedOperatorApply lives in libCEED (dev snapshot between
 * v0.6 and v0.7, un-pinned: reference Makefile:20-21 `CEED_DIR ?= ../..`), which is
 * absent from /root/reference.  Its published /cpu/self algorithm is restated here
 * (SURVEY.md Appendix B): Gauss / Gauss-Lobatto rules by Newton iteration on the
 * Legendre recurrence, Lagrange interp/grad matrices by Fornberg's recurrence,
 * offset restriction gather / transpose scatter-add (compstride interlaced),
 * tensor-product contraction with the x index fastest, per-element operator loop,
 * and the diagonal assembly from the point-wise assembled QFunction.
 *
 * PARITY PINNING.  The reference repository holds no golden vectors for this path
 * (SURVEY.md 8(c)).  The QFunction layer is pinned against the reference's real
 * functions compiled from /root/reference/qfunctions (oracle/_ref/libref_qf.so) and
 * the committed vectors in tests/golden/.  The libCEED layer restated in THIS file
 * is "parity unpinned" by reference-owned vectors; it is validated through
 * properties only (polynomial exactness, J.d vs finite differences of F, symmetry,
 * rigid-body null space, diagonal vs unit-vector probing): tests/test_oracle_*.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may call into this file.  The product never does.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int (*OracleQFn)(void *ctx, int Q, const double *const *in, double *const *out);

/* ------------------------------------------------------------------------- */
/* B.2  one-dimensional rules                                                 */
/* ------------------------------------------------------------------------- */

/* Legendre P_n(x) and P_{n-1}(x) by the three-term recurrence */
static void legendre(int n, double x, double *Pn, double *Pnm1) {
  double p0 = 1.0, p1 = x;
  if (n == 0) { *Pn = 1.0; *Pnm1 = 0.0; return; }
  for (int j = 2; j <= n; j++) {
    const double p2 = ((2 * j - 1) * x * p1 - (j - 1) * p0) / j;
    p0 = p1;
    p1 = p2;
  }
  *Pn = p1;
  *Pnm1 = p0;
}

/* Gauss-Legendre, Q points on [-1,1], ascending */
int oracle_gauss(int Q, double *x, double *w) {
  const double PI = 4.0 * atan(1.0);
  for (int i = 0; i <= Q / 2; i++) {
    double xi = cos(PI * (2 * i + 1) / (2.0 * Q));
    double PQ, PQm1, dP = 1;
    for (int it = 0; it < 100; it++) {
      legendre(Q, xi, &PQ, &PQm1);
      dP = (xi * PQ - PQm1) * Q / (xi * xi - 1.0);
      const double dx = PQ / dP;
      xi -= dx;
      if (fabs(dx) < 1e-16 || fabs(PQ) < 1e-15) break;
    }
    legendre(Q, xi, &PQ, &PQm1);
    dP = (xi * PQ - PQm1) * Q / (xi * xi - 1.0);
    const double wi = 2.0 / ((1.0 - xi * xi) * dP * dP);
    w[i] = w[Q - 1 - i] = wi;
    x[i] = -xi;
    x[Q - 1 - i] = xi;
  }
  return 0;
}

/* Gauss-Legendre-Lobatto, Q points on [-1,1], ascending; interior points are the
 * roots of P'_{Q-1}; weights 2/(Q(Q-1) P_{Q-1}(x)^2) */
int oracle_lobatto(int Q, double *x, double *w) {
  const double PI = 4.0 * atan(1.0);
  const int n = Q - 1;
  if (Q < 2) return 1;
  x[0] = -1.0;
  x[Q - 1] = 1.0;
  w[0] = w[Q - 1] = 2.0 / (Q * (double)n);
  for (int i = 1; i <= n / 2; i++) {
    double xi = cos(PI * i / n);
    double Pn, Pnm1;
    for (int it = 0; it < 100; it++) {
      legendre(n, xi, &Pn, &Pnm1);
      /* P'_n = n (x P_n - P_{n-1}) / (x^2-1);  P''_n from the Legendre ODE */
      const double dP = (xi * Pn - Pnm1) * n / (xi * xi - 1.0);
      const double d2P = (2 * xi * dP - n * (n + 1.0) * Pn) / (1.0 - xi * xi);
      const double dx = dP / d2P;
      xi -= dx;
      if (fabs(dx) < 1e-16) break;
    }
    legendre(n, xi, &Pn, &Pnm1);
    const double wi = 2.0 / (Q * (double)n * Pn * Pn);
    w[i] = w[Q - 1 - i] = wi;
    x[i] = -xi;
    x[Q - 1 - i] = xi;
  }
  return 0;
}

/* Lagrange basis on `nodes[P]` evaluated at `q[Q]`: interp[Q*P], grad[Q*P]
 * (row = quadrature point) by Fornberg's recurrence */
int oracle_lagrange_at(int P, const double *nodes, int Q, const double *q,
                       double *interp, double *grad) {
  for (int i = 0; i < Q; i++) {
    double *B = interp + (size_t)i * P, *D = grad + (size_t)i * P;
    for (int j = 0; j < P; j++) B[j] = D[j] = 0.0;
    double c1 = 1.0, c3 = nodes[0] - q[i];
    B[0] = 1.0;
    for (int j = 1; j < P; j++) {
      double c2 = 1.0;
      const double c4 = c3;
      c3 = nodes[j] - q[i];
      for (int k = 0; k < j; k++) {
        const double dx = nodes[j] - nodes[k];
        c2 *= dx;
        if (k == j - 1) {
          D[j] = c1 * (B[k] - c4 * D[k]) / c2;
          B[j] = -c1 * c4 * B[k] / c2;
        }
        D[k] = (c3 * D[k] - B[k]) / dx;
        B[k] = c3 * B[k] / dx;
      }
      c1 = c2;
    }
  }
  return 0;
}

/* CeedBasisCreateTensorH1Lagrange(dim, ncomp, P, Q, qmode) -- 1-D pieces.
 * qmode 0 = CEED_GAUSS, 1 = CEED_GAUSS_LOBATTO.  Outputs interp1d[Q*P], grad1d[Q*P],
 * qref1d[Q], qweight1d[Q].  (setuplibceed.c:335-348, :782-803) */
int oracle_basis_1d(int P, int Q, int qmode, double *interp1d, double *grad1d,
                    double *qref1d, double *qweight1d) {
  double *nodes = (double *)malloc(sizeof(double) * P * 2);
  if (!nodes) return 1;
  oracle_lobatto(P, nodes, nodes + P);
  if (qmode == 0) oracle_gauss(Q, qref1d, qweight1d);
  else oracle_lobatto(Q, qref1d, qweight1d);
  oracle_lagrange_at(P, nodes, Q, qref1d, interp1d, grad1d);
  free(nodes);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* B.1  restrictions                                                          */
/* ------------------------------------------------------------------------- */

/* E[e][c][n] = L[offsets[e*elemsize+n] + c*compstride] */
int oracle_restrict_gather(int nelem, int elemsize, int ncomp, int compstride,
                           const int *offsets, const double *L, double *E) {
  for (int e = 0; e < nelem; e++)
    for (int c = 0; c < ncomp; c++)
      for (int n = 0; n < elemsize; n++)
        E[((size_t)e * ncomp + c) * elemsize + n] =
            L[(size_t)offsets[(size_t)e * elemsize + n] + (size_t)c * compstride];
  return 0;
}

/* L[offsets[...] + c*compstride] += E[e][c][n], serial order (e, c, n) */
int oracle_restrict_scatter_add(int nelem, int elemsize, int ncomp, int compstride,
                                const int *offsets, const double *E, double *L) {
  for (int e = 0; e < nelem; e++)
    for (int c = 0; c < ncomp; c++)
      for (int n = 0; n < elemsize; n++)
        L[(size_t)offsets[(size_t)e * elemsize + n] + (size_t)c * compstride] +=
            E[((size_t)e * ncomp + c) * elemsize + n];
  return 0;
}

/* CeedElemRestrictionGetMultiplicity: transpose-apply of ones (misc.c:117-123) */
int oracle_multiplicity(int nelem, int elemsize, int ncomp, int compstride, int lsize,
                        const int *offsets, double *mult) {
  memset(mult, 0, sizeof(double) * (size_t)lsize);
  for (int e = 0; e < nelem; e++)
    for (int c = 0; c < ncomp; c++)
      for (int n = 0; n < elemsize; n++)
        mult[(size_t)offsets[(size_t)e * elemsize + n] + (size_t)c * compstride] += 1.0;
  return 0;
}

/* ------------------------------------------------------------------------- */
/* B.3  tensor basis, one element                                             */
/* ------------------------------------------------------------------------- */

/* v[a][j][c] (+)= sum_b t[j][b] u[a][b][c];  t is J x Bd (or its transpose) */
static void contract(int A, int Bd, int C, int J, const double *t, int tmode, int add,
                     const double *u, double *v) {
  const int tsj = tmode ? 1 : Bd, tsb = tmode ? J : 1;
  if (!add) memset(v, 0, sizeof(double) * (size_t)A * J * C);
  for (int a = 0; a < A; a++)
    for (int b = 0; b < Bd; b++)
      for (int j = 0; j < J; j++) {
        const double tq = t[j * tsj + b * tsb];
        for (int c = 0; c < C; c++) v[((size_t)a * J + j) * C + c] += tq * u[((size_t)a * Bd + b) * C + c];
      }
}

static int ipow(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

/* Apply (mat_x in slot 0 ... ) to all ncomp components: arrays are [comp][z][y][x].
 * Contracting dimension by dimension starting with the fastest (x) index:
 * view as [pre][n_in][post] with post the already-untouched faster dims. */
static void tensor_apply3(int ncomp, int P, int Q, const double *m0, const double *m1,
                          const double *m2, int transpose, int add, const double *u,
                          double *v, double *tmp1, double *tmp2) {
  /* forward: in sizes P -> out sizes Q (matrices Q x P); transpose swaps roles */
  const int nin = transpose ? Q : P, nout = transpose ? P : Q;
  /* x (fastest): pre = ncomp*nin*nin, post = 1 */
  contract(ncomp * nin * nin, nin, 1, nout, m0, transpose, 0, u, tmp1);
  /* y: pre = ncomp*nin, post = nout */
  contract(ncomp * nin, nin, nout, nout, m1, transpose, 0, tmp1, tmp2);
  /* z: pre = ncomp, post = nout*nout */
  contract(ncomp, nin, nout * nout, nout, m2, transpose, add, tmp2, v);
}

/* CeedBasisApply for one element.
 *  emode: 1 = INTERP, 2 = GRAD, 4 = WEIGHT (libCEED CeedEvalMode values)
 *  tmode: 0 = NOTRANSPOSE (nodes -> quadrature), 1 = TRANSPOSE (overwrites v)
 *  INTERP: u[ncomp][P^3] -> v[ncomp][Q^3]
 *  GRAD:   u[ncomp][P^3] -> v[dim][ncomp][Q^3]       (d slowest)
 *  WEIGHT: v[Q^3] */
int oracle_basis_apply_elem(int ncomp, int P, int Q, const double *interp1d,
                            const double *grad1d, const double *qweight1d, int tmode,
                            int emode, const double *u, double *v) {
  const int P3 = P * P * P, Q3 = Q * Q * Q;
  const int mx = P > Q ? P : Q;
  if (emode == 4) {
    for (int k = 0; k < Q; k++)
      for (int j = 0; j < Q; j++)
        for (int i = 0; i < Q; i++)
          v[(k * Q + j) * Q + i] = qweight1d[i] * qweight1d[j] * qweight1d[k];
    return 0;
  }
  double *tmp1 = (double *)malloc(sizeof(double) * 2 * (size_t)ncomp * mx * mx * mx);
  if (!tmp1) return 1;
  double *tmp2 = tmp1 + (size_t)ncomp * mx * mx * mx;
  if (emode == 1) {
    tensor_apply3(ncomp, P, Q, interp1d, interp1d, interp1d, tmode, 0, u, v, tmp1, tmp2);
  } else if (emode == 2) {
    for (int d = 0; d < 3; d++) {
      const double *m0 = d == 0 ? grad1d : interp1d;
      const double *m1 = d == 1 ? grad1d : interp1d;
      const double *m2 = d == 2 ? grad1d : interp1d;
      if (!tmode)
        tensor_apply3(ncomp, P, Q, m0, m1, m2, 0, 0, u, v + (size_t)d * ncomp * Q3, tmp1, tmp2);
      else
        tensor_apply3(ncomp, P, Q, m0, m1, m2, 1, d > 0, u + (size_t)d * ncomp * Q3, v, tmp1, tmp2);
    }
  } else {
    free(tmp1);
    return 2;
  }
  (void)P3;
  free(tmp1);
  return 0;
}

/* Batched CeedBasisApply over nelem elements: E-vector [elem][comp][P^3],
 * Q-vector per element contiguous ([elem][...] with the per-element layout above) */
int oracle_basis_apply(int nelem, int ncomp, int P, int Q, const double *interp1d,
                       const double *grad1d, const double *qweight1d, int tmode, int emode,
                       const double *u, double *v) {
  const size_t P3 = (size_t)P * P * P, Q3 = (size_t)Q * Q * Q;
  const size_t nsz = ncomp * P3;
  const size_t qsz = emode == 2 ? 3 * ncomp * Q3 : (emode == 4 ? Q3 : ncomp * Q3);
  int err = 0;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < nelem; e++) {
    int r;
    if (emode == 4) r = oracle_basis_apply_elem(ncomp, P, Q, interp1d, grad1d, qweight1d, 0, 4, NULL, v + e * qsz);
    else if (!tmode) r = oracle_basis_apply_elem(ncomp, P, Q, interp1d, grad1d, qweight1d, 0, emode, u + e * nsz, v + e * qsz);
    else r = oracle_basis_apply_elem(ncomp, P, Q, interp1d, grad1d, qweight1d, 1, emode, u + e * qsz, v + e * nsz);
    if (r) err = r;
  }
  return err;
}

/* ------------------------------------------------------------------------- */
/* geometric factors: SetupGeo operator (setuplibceed.c:370-393)              */
/* ------------------------------------------------------------------------- */

/* xoffsets[nelem][8] into the interlaced coordinate L-vector xcoord (ncompx = 3,
 * compstride 1), coordinate basis P=2 -> Q.  qdata out: [elem][10][Q^3]
 * (the /cpu/self meaning of CEED_STRIDES_BACKEND, App. B.1). */
int oracle_setup_geo(OracleQFn setupgeo, int nelem, int Q, const int *xoffsets,
                     const double *xcoord, double *qdata) {
  const int Q3 = Q * Q * Q;
  double bx[16 * 2], dx[16 * 2], qr[16], qw[16];
  if (Q > 16) return 1;
  oracle_basis_1d(2, Q, 0, bx, dx, qr, qw);
  int err = 0;
#pragma omp parallel
  {
    double *xe = (double *)malloc(sizeof(double) * (3 * 8 + 9 * (size_t)Q3 + Q3));
    double *J = xe + 24, *w = J + 9 * (size_t)Q3;
#pragma omp for schedule(static)
    for (int e = 0; e < nelem; e++) {
      for (int c = 0; c < 3; c++)
        for (int n = 0; n < 8; n++) xe[c * 8 + n] = xcoord[(size_t)xoffsets[(size_t)e * 8 + n] + c];
      oracle_basis_apply_elem(3, 2, Q, bx, dx, qw, 0, 2, xe, J);
      oracle_basis_apply_elem(1, 2, Q, bx, dx, qw, 0, 4, NULL, w);
      const double *in[2] = {J, w};
      double *out[1] = {qdata + (size_t)e * 10 * Q3};
      if (setupgeo(NULL, Q3, in, out)) err = 1;
    }
    free(xe);
  }
  return err;
}

/* ------------------------------------------------------------------------- */
/* B.4  operator apply (residual / Jacobian shape)                            */
/* ------------------------------------------------------------------------- */

/* y_L += E^T G^T D(qdata[,gradu]) G E x_L  for the reference's solid-mechanics
 * operators.
 *   qf        user QFunction (reference or port)
 *   ctx       Physics {nu, E}
 *   gradu_mode 0: no gradu field (linElas); 1: gradu is a passive OUTPUT written by
 *             the QFunction (residual HyperSSF/HyperFSF); 2: gradu is a passive INPUT
 *             (Jacobian HyperSSdF/HyperFSdF)
 *   offsets   [nelem][P^3] node offsets into the interlaced L-vector (compstride 1)
 *   qdata     [elem][10][Q^3];  gradu [elem][9][Q^3]
 * Parallel over elements into an E-vector, then the serial (e,c,n)-ordered
 * scatter-add of /cpu/self.  The caller zeroes y (CeedOperatorApply semantics). */
int oracle_operator_apply_add(OracleQFn qf, void *ctx, int gradu_mode, int nelem, int P,
                              int Q, const double *interp1d, const double *grad1d,
                              const int *offsets, const double *qdata, double *gradu,
                              const double *x, double *y) {
  const int P3 = P * P * P, Q3 = Q * Q * Q;
  double *Ey = (double *)malloc(sizeof(double) * (size_t)nelem * 3 * P3);
  if (!Ey) return 1;
  int err = 0;
#pragma omp parallel
  {
    double *ue = (double *)malloc(sizeof(double) * (3 * (size_t)P3 + 18 * (size_t)Q3));
    double *du = ue + 3 * (size_t)P3, *dv = du + 9 * (size_t)Q3;
#pragma omp for schedule(static)
    for (int e = 0; e < nelem; e++) {
      for (int c = 0; c < 3; c++)
        for (int n = 0; n < P3; n++) ue[c * P3 + n] = x[(size_t)offsets[(size_t)e * P3 + n] + c];
      oracle_basis_apply_elem(3, P, Q, interp1d, grad1d, NULL, 0, 2, ue, du);
      const double *qd = qdata + (size_t)e * 10 * Q3;
      double *ge = gradu ? gradu + (size_t)e * 9 * Q3 : NULL;
      int r;
      if (gradu_mode == 1) {
        const double *in[2] = {du, qd};
        double *out[2] = {dv, ge};
        r = qf(ctx, Q3, in, out);
      } else if (gradu_mode == 2) {
        const double *in[3] = {du, qd, ge};
        double *out[1] = {dv};
        r = qf(ctx, Q3, in, out);
      } else {
        const double *in[2] = {du, qd};
        double *out[1] = {dv};
        r = qf(ctx, Q3, in, out);
      }
      if (r) err = r;
      oracle_basis_apply_elem(3, P, Q, interp1d, grad1d, NULL, 1, 2, dv, Ey + (size_t)e * 3 * P3);
    }
    free(ue);
  }
  oracle_restrict_scatter_add(nelem, P3, 3, 1, offsets, Ey, y);
  free(Ey);
  return err;
}

/* ------------------------------------------------------------------------- */
/* B.5  operator diagonal                                                     */
/* ------------------------------------------------------------------------- */

/* diag_L += E^T diag_e, diag_e[c][n] = sum_q sum_{dout,din} G_dout[q,n] Dq[(din,c)][(dout,c)][q] G_din[q,n]
 * with Dq obtained by feeding unit fields through the (linear) Jacobian QFunction.
 * Entries with |Dq| <= 1e-12 * max|Dq| are skipped as upstream does.
 * gradu_mode as above (0 or 2).  The caller zeroes diag. */
int oracle_operator_diagonal_add(OracleQFn qf, void *ctx, int gradu_mode, int nelem, int P,
                                 int Q, const double *interp1d, const double *grad1d,
                                 const int *offsets, const double *qdata, double *gradu,
                                 double *diag) {
  const int P3 = P * P * P, Q3 = Q * Q * Q;
  double *Ed = (double *)calloc((size_t)nelem * 3 * P3, sizeof(double));
  /* dense 3-D gradient matrices G[d][q][n] */
  double *G = (double *)malloc(sizeof(double) * 3 * (size_t)Q3 * P3);
  if (!Ed || !G) return 1;
  for (int d = 0; d < 3; d++)
    for (int qz = 0; qz < Q; qz++)
      for (int qy = 0; qy < Q; qy++)
        for (int qx = 0; qx < Q; qx++)
          for (int k = 0; k < P; k++)
            for (int j = 0; j < P; j++)
              for (int i = 0; i < P; i++) {
                const double fx = (d == 0 ? grad1d : interp1d)[qx * P + i];
                const double fy = (d == 1 ? grad1d : interp1d)[qy * P + j];
                const double fz = (d == 2 ? grad1d : interp1d)[qz * P + k];
                G[((size_t)d * Q3 + (qz * Q + qy) * Q + qx) * P3 + (k * P + j) * P + i] = fx * fy * fz;
              }
  int err = 0;
#pragma omp parallel
  {
    double *unit = (double *)malloc(sizeof(double) * (9 * (size_t)Q3 * 2 + 81 * (size_t)Q3));
    double *dv = unit + 9 * (size_t)Q3, *Dq = dv + 9 * (size_t)Q3;
#pragma omp for schedule(static)
    for (int e = 0; e < nelem; e++) {
      const double *qd = qdata + (size_t)e * 10 * Q3;
      double *ge = gradu ? gradu + (size_t)e * 9 * Q3 : NULL;
      double dmax = 0;
      for (int s = 0; s < 9; s++) { /* s = din*3 + cin */
        memset(unit, 0, sizeof(double) * 9 * (size_t)Q3);
        for (int q = 0; q < Q3; q++) unit[(size_t)s * Q3 + q] = 1.0;
        int r;
        if (gradu_mode == 2) {
          const double *in[3] = {unit, qd, ge};
          double *out[1] = {dv};
          r = qf(ctx, Q3, in, out);
        } else {
          const double *in[2] = {unit, qd};
          double *out[1] = {dv};
          r = qf(ctx, Q3, in, out);
        }
        if (r) err = r;
        memcpy(Dq + (size_t)s * 9 * Q3, dv, sizeof(double) * 9 * (size_t)Q3);
        for (size_t t = 0; t < 9 * (size_t)Q3; t++)
          if (fabs(dv[t]) > dmax) dmax = fabs(dv[t]);
      }
      const double thresh = 1e-12 * dmax;
      for (int c = 0; c < 3; c++)
        for (int din = 0; din < 3; din++)
          for (int dout = 0; dout < 3; dout++) {
            const double *D = Dq + ((size_t)(din * 3 + c) * 9 + (dout * 3 + c)) * Q3;
            for (int q = 0; q < Q3; q++) {
              if (fabs(D[q]) <= thresh) continue;
              const double *go = G + ((size_t)dout * Q3 + q) * P3, *gi = G + ((size_t)din * Q3 + q) * P3;
              double *de = Ed + ((size_t)e * 3 + c) * P3;
              for (int n = 0; n < P3; n++) de[n] += go[n] * D[q] * gi[n];
            }
          }
    }
    free(unit);
  }
  oracle_restrict_scatter_add(nelem, P3, 3, 1, offsets, Ed, diag);
  free(Ed);
  free(G);
  return err;
}

/* ------------------------------------------------------------------------- */
/* p-multigrid transfer (matops.c:115-203; setuplibceed.c:799-803,847-863)     */
/* ------------------------------------------------------------------------- */

/* Prolongation: f_L += Ef^T ( I_{c->f} Ec c_L ), identity QFunction, interpolation
 * basis Pc -> Pf at the fine GLL points.  transpose=1 gives the restriction
 * c_L += Ec^T I^T Ef f_L.  Multiplicity scaling is the caller's (matops.c:149,176). */
int oracle_transfer_add(int transpose, int nelem, int Pc, int Pf, const double *interpCtoF,
                        const int *offc, const int *offf, const double *in, double *out) {
  const int Pc3 = Pc * Pc * Pc, Pf3 = Pf * Pf * Pf;
  const int nin = transpose ? Pf3 : Pc3, nout = transpose ? Pc3 : Pf3;
  const int *oin = transpose ? offf : offc, *oout = transpose ? offc : offf;
  double *Eo = (double *)malloc(sizeof(double) * (size_t)nelem * 3 * nout);
  if (!Eo) return 1;
#pragma omp parallel
  {
    double *ue = (double *)malloc(sizeof(double) * 3 * (size_t)nin);
#pragma omp for schedule(static)
    for (int e = 0; e < nelem; e++) {
      for (int c = 0; c < 3; c++)
        for (int n = 0; n < nin; n++) ue[c * nin + n] = in[(size_t)oin[(size_t)e * nin + n] + c];
      oracle_basis_apply_elem(3, Pc, Pf, interpCtoF, interpCtoF, NULL, transpose, 1, ue,
                              Eo + (size_t)e * 3 * nout);
    }
    free(ue);
  }
  oracle_restrict_scatter_add(nelem, nout, 3, 1, oout, Eo, out);
  free(Eo);
  return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the benchmark's CPU legs set the
 * thread count explicitly (all host cores on rank 0) instead of inheriting that. */
int oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}
