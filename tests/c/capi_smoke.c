/* C-side use of the drop-in boundary: the call sequence of the reference's
 * SetupLibceedFineLevel / SetupLibceedLevel (/root/reference/src/setuplibceed.c:278-393,518-542,
 * 818-839) and ApplyLocalCeedOp (/root/reference/src/matops.c:40-50) for linear elasticity on a
 * 2x2x2 box of degree 2, written against <ceed.h> only.  A QFunction is declared with the
 * CEED_QFUNCTION macro exactly as the reference's qfunctions headers do.  The backend dispatches on
 * the ":Name" locator and, at operator set-up, runs the host pointer on a few known points to make sure
 * it computes what its device body computes: the bodies here forward to the oracle's restated
 * QFunctions (oracle/qf_port.c, test infrastructure), standing in for the reference's headers.
 * Exit code 0 and "capi_smoke OK" = operators built, applied, diagonal assembled, results finite. */
#include <ceed.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

typedef struct { CeedScalar nu, E; } Physics_s;

int port_SetupGeo(void *ctx, int Q, const double *const *in, double *const *out);
int port_LinElasF(void *ctx, int Q, const double *const *in, double *const *out);
int port_LinElasdF(void *ctx, int Q, const double *const *in, double *const *out);
CEED_QFUNCTION(SetupGeo)(void *ctx, CeedInt Q, const CeedScalar *const *in, CeedScalar *const *out) { return port_SetupGeo(ctx, Q, in, out); }
CEED_QFUNCTION(LinElasF)(void *ctx, CeedInt Q, const CeedScalar *const *in, CeedScalar *const *out) { return port_LinElasF(ctx, Q, in, out); }
CEED_QFUNCTION(LinElasdF)(void *ctx, CeedInt Q, const CeedScalar *const *in, CeedScalar *const *out) { return port_LinElasdF(ctx, Q, in, out); }

static CeedInt *box_offsets(int n, int p, int ncomp) {
  const int P = p + 1, N = n * p + 1;
  CeedInt *off = malloc(sizeof(CeedInt) * n * n * n * P * P * P);
  size_t k = 0;
  for (int ez = 0; ez < n; ez++) for (int ey = 0; ey < n; ey++) for (int ex = 0; ex < n; ex++)
    for (int c = 0; c < P; c++) for (int b = 0; b < P; b++) for (int a = 0; a < P; a++)
      off[k++] = ncomp * ((ex * p + a) + N * ((ey * p + b) + N * (ez * p + c)));
  return off;
}

int main(int argc, char **argv) {
  const char *resource = argc > 1 ? argv[1] : "/gpu/b200";
  const int n = 2, p = 2, P = p + 1, Q = p + 1, dim = 3, ncompu = 3, qdatasize = 10;
  const CeedInt nelem = n * n * n, Nu = n * p + 1, Nx = n + 1, Ulocsz = ncompu * Nu * Nu * Nu;
  Ceed ceed;
  CeedMemType memtype;
  Physics_s phys = {0.3, 1.0};
  CeedInit(resource, &ceed);
  CeedGetPreferredMemType(ceed, &memtype);
  printf("resource %s, preferred memtype %s\n", resource, CeedMemTypes[memtype]);

  CeedElemRestriction Erestrictx, Erestrictu, Erestrictqdi;
  CeedInt *ox = box_offsets(n, 1, 3), *ou = box_offsets(n, p, ncompu);
  CeedElemRestrictionCreate(ceed, nelem, 8, 3, 1, 3 * Nx * Nx * Nx, CEED_MEM_HOST, CEED_COPY_VALUES, ox, &Erestrictx);
  CeedElemRestrictionCreate(ceed, nelem, P * P * P, ncompu, 1, Ulocsz, CEED_MEM_HOST, CEED_COPY_VALUES, ou, &Erestrictu);
  free(ox); free(ou);
  CeedElemRestrictionCreateStrided(ceed, nelem, Q * Q * Q, qdatasize, qdatasize * nelem * Q * Q * Q, CEED_STRIDES_BACKEND,
                                   &Erestrictqdi);
  CeedVector xcoord, qdata, xceed, yceed;
  CeedElemRestrictionCreateVector(Erestrictx, &xcoord, NULL);
  CeedScalar *coords = malloc(sizeof(CeedScalar) * 3 * Nx * Nx * Nx);
  for (int k = 0, i = 0; k < Nx; k++) for (int j = 0; j < Nx; j++) for (int ii = 0; ii < Nx; ii++) {
    coords[i++] = (double)ii / n; coords[i++] = (double)j / n; coords[i++] = (double)k / n;
  }
  CeedVectorSetArray(xcoord, CEED_MEM_HOST, CEED_COPY_VALUES, coords);
  free(coords);
  CeedBasis basisu, basisx;
  CeedBasisCreateTensorH1Lagrange(ceed, dim, ncompu, P, Q, CEED_GAUSS, &basisu);
  CeedBasisCreateTensorH1Lagrange(ceed, dim, 3, 2, Q, CEED_GAUSS, &basisx);
  CeedInt nqpts;
  CeedBasisGetNumQuadraturePoints(basisu, &nqpts);
  CeedVectorCreate(ceed, qdatasize * nelem * nqpts, &qdata);

  CeedQFunction qfSetupGeo, qfApply, qfJacob;
  CeedOperator opSetupGeo, opApply, opJacob;
  CeedQFunctionCreateInterior(ceed, 1, SetupGeo, SetupGeo_loc, &qfSetupGeo);
  CeedQFunctionAddInput(qfSetupGeo, "dx", 9, CEED_EVAL_GRAD);
  CeedQFunctionAddInput(qfSetupGeo, "weight", 1, CEED_EVAL_WEIGHT);
  CeedQFunctionAddOutput(qfSetupGeo, "qdata", qdatasize, CEED_EVAL_NONE);
  CeedOperatorCreate(ceed, qfSetupGeo, CEED_QFUNCTION_NONE, CEED_QFUNCTION_NONE, &opSetupGeo);
  CeedOperatorSetField(opSetupGeo, "dx", Erestrictx, basisx, CEED_VECTOR_ACTIVE);
  CeedOperatorSetField(opSetupGeo, "weight", CEED_ELEMRESTRICTION_NONE, basisx, CEED_VECTOR_NONE);
  CeedOperatorSetField(opSetupGeo, "qdata", Erestrictqdi, CEED_BASIS_COLLOCATED, CEED_VECTOR_ACTIVE);
  CeedOperatorApply(opSetupGeo, xcoord, qdata, CEED_REQUEST_IMMEDIATE);
  CeedQFunctionDestroy(&qfSetupGeo);
  CeedOperatorDestroy(&opSetupGeo);

  CeedQFunctionCreateInterior(ceed, 1, LinElasF, LinElasF_loc, &qfApply);
  CeedQFunctionAddInput(qfApply, "du", 9, CEED_EVAL_GRAD);
  CeedQFunctionAddInput(qfApply, "qdata", qdatasize, CEED_EVAL_NONE);
  CeedQFunctionAddOutput(qfApply, "dv", 9, CEED_EVAL_GRAD);
  CeedQFunctionSetContext(qfApply, &phys, sizeof(phys));
  CeedOperatorCreate(ceed, qfApply, CEED_QFUNCTION_NONE, CEED_QFUNCTION_NONE, &opApply);
  CeedOperatorSetField(opApply, "du", Erestrictu, basisu, CEED_VECTOR_ACTIVE);
  CeedOperatorSetField(opApply, "qdata", Erestrictqdi, CEED_BASIS_COLLOCATED, qdata);
  CeedOperatorSetField(opApply, "dv", Erestrictu, basisu, CEED_VECTOR_ACTIVE);

  CeedQFunctionCreateInterior(ceed, 1, LinElasdF, LinElasdF_loc, &qfJacob);
  CeedQFunctionAddInput(qfJacob, "deltadu", 9, CEED_EVAL_GRAD);
  CeedQFunctionAddInput(qfJacob, "qdata", qdatasize, CEED_EVAL_NONE);
  CeedQFunctionAddOutput(qfJacob, "deltadv", 9, CEED_EVAL_GRAD);
  CeedQFunctionSetContext(qfJacob, &phys, sizeof(&phys)); /* the reference's sizeof(pointer) quirk */
  CeedOperatorCreate(ceed, qfJacob, CEED_QFUNCTION_NONE, CEED_QFUNCTION_NONE, &opJacob);
  CeedOperatorSetField(opJacob, "deltadu", Erestrictu, basisu, CEED_VECTOR_ACTIVE);
  CeedOperatorSetField(opJacob, "qdata", Erestrictqdi, CEED_BASIS_COLLOCATED, qdata);
  CeedOperatorSetField(opJacob, "deltadv", Erestrictu, basisu, CEED_VECTOR_ACTIVE);

  /* ApplyLocalCeedOp with -memtype host: borrow host arrays, apply, take them back */
  CeedVectorCreate(ceed, Ulocsz, &xceed);
  CeedVectorCreate(ceed, Ulocsz, &yceed);
  CeedScalar *x = malloc(sizeof(CeedScalar) * Ulocsz), *y = calloc(Ulocsz, sizeof(CeedScalar)), *y2 = calloc(Ulocsz, sizeof(CeedScalar));
  for (int i = 0; i < Ulocsz; i++) x[i] = sin(0.1 * i);
  CeedVectorSetArray(xceed, CEED_MEM_HOST, CEED_USE_POINTER, x);
  CeedVectorSetArray(yceed, CEED_MEM_HOST, CEED_USE_POINTER, y);
  CeedOperatorApply(opApply, xceed, yceed, CEED_REQUEST_IMMEDIATE);
  CeedVectorTakeArray(xceed, CEED_MEM_HOST, NULL);
  CeedVectorTakeArray(yceed, CEED_MEM_HOST, NULL);
  CeedVectorSetArray(xceed, CEED_MEM_HOST, CEED_USE_POINTER, x);
  CeedVectorSetArray(yceed, CEED_MEM_HOST, CEED_USE_POINTER, y2);
  CeedOperatorApply(opJacob, xceed, yceed, CEED_REQUEST_IMMEDIATE);
  CeedVectorTakeArray(xceed, CEED_MEM_HOST, NULL);
  CeedVectorTakeArray(yceed, CEED_MEM_HOST, NULL);
  double n1 = 0, diff = 0, sum = 0;
  for (int i = 0; i < Ulocsz; i++) { n1 += y[i] * y[i]; diff += (y[i] - y2[i]) * (y[i] - y2[i]); sum += y[i]; }
  /* diagonal into a borrowed host array (GetDiag_Ceed, matops.c:223-235) */
  CeedVectorSetArray(xceed, CEED_MEM_HOST, CEED_USE_POINTER, x);
  CeedOperatorLinearAssembleDiagonal(opJacob, xceed, CEED_REQUEST_IMMEDIATE);
  CeedVectorTakeArray(xceed, CEED_MEM_HOST, NULL);
  double dmin = 1e300;
  for (int i = 0; i < Ulocsz; i++) dmin = x[i] < dmin ? x[i] : dmin;
  printf("|F(x)| = %.15e, |F(x) - J x| = %.3e (linear problem), sum F = %.3e (rigid modes), min diag = %.3e\n",
         sqrt(n1), sqrt(diff), sum, dmin);
  const int ok = isfinite(n1) && n1 > 0 && sqrt(diff) < 1e-13 * sqrt(n1) && fabs(sum) < 1e-12 * sqrt(n1) && dmin > 0;
  free(x); free(y); free(y2);
  CeedVectorDestroy(&xceed); CeedVectorDestroy(&yceed); CeedVectorDestroy(&qdata); CeedVectorDestroy(&xcoord);
  CeedOperatorDestroy(&opApply); CeedOperatorDestroy(&opJacob);
  CeedQFunctionDestroy(&qfApply); CeedQFunctionDestroy(&qfJacob);
  CeedBasisDestroy(&basisu); CeedBasisDestroy(&basisx);
  CeedElemRestrictionDestroy(&Erestrictx); CeedElemRestrictionDestroy(&Erestrictu); CeedElemRestrictionDestroy(&Erestrictqdi);
  CeedDestroy(&ceed);
  printf(ok ? "capi_smoke OK\n" : "capi_smoke FAILED\n");
  return ok ? 0 : 1;
}
