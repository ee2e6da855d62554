/* TEST INFRASTRUCTURE ONLY (see oracle/README.md) -- NOT part of the product, never linked into libceed_b200.so.
 *
 * A generic CPU restatement of the libCEED user API subset that the reference's host code calls
 * (prototypes: include/ceed/ceed.h), with the `/cpu/self` semantics of SURVEY.md
 * Appendix B: object model (vectors, offset / strided restrictions, tensor H1 Lagrange bases, QFunctions given as
 * HOST FUNCTION POINTERS, operators wired field by field with CeedOperatorSetField), element-by-element apply
 * (restrict -> basis -> user QFunction -> basis^T -> restrict^T, serial (e,c,n) summation order) and
 * CeedOperatorLinearAssembleDiagonal by unit inputs through the QFunction.
 *
 * libCEED itself is a third-party dependency of the reference that is not vendored in /root/reference (no version
 * pinned: Makefile:20-21 `CEED_DIR ?= ../..`; API level = upstream main between v0.6 and v0.7).  This file restates
 * its published algorithm for exactly the calls at
 *     /root/reference/src/setuplibceed.c:243-939     (every object created and wired)
 *     /root/reference/src/matops.c:26-300            (Apply, LinearAssembleDiagonal, context swap)
 *     /root/reference/src/misc.c:115-143, 217-300    (GetMultiplicity, diagnostic operator)
 * so that those files -- compiled unchanged from where they lie -- can run on the CPU with the reference's own
 * QFunctions: oracle/_ref/ref_driver_cpu (oracle/Makefile target `refdriver_cpu`).  What the reference's
 * FormResidual_Ceed / ApplyJacobian_Ceed / GetDiag_Ceed / Prolong_Ceed / Restrict_Ceed / ComputeStrainEnergy /
 * ViewDiagnosticQuantities produce through THIS generic path is compared with the fixed-shape restatement of
 * oracle/ceed_oracle.c in tests/test_reference_host_code_on_cpu.py: two independent wirings of the same semantics,
 * one of them dictated by the reference's call sites.  The numerical kernels (quadrature rules, Lagrange matrices,
 * tensor contraction) are shared with ceed_oracle.c.
 *
 * Host memory only: CEED_MEM_DEVICE requests abort.  Errors print and abort (the reference never checks return codes).
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ceed/ceed.h"

/* from ceed_oracle.c */
int oracle_basis_1d(int P, int Q, int qmode, double *interp1d, double *grad1d, double *qref1d, double *qweight1d);
int oracle_gauss(int Q, double *x, double *w);
int oracle_lobatto(int Q, double *x, double *w);
int oracle_basis_apply_elem(int ncomp, int P, int Q, const double *interp1d, const double *grad1d,
                            const double *qweight1d, int tmode, int emode, const double *u, double *v);

#define MAXF 16

struct Ceed_private { char resource[128]; };
struct CeedRequest_private { int unused; };
struct CeedVector_private { Ceed ceed; CeedInt n; double *a; int owned, valid, refcount; };
struct CeedElemRestriction_private {
  Ceed ceed;
  CeedInt nelem, elemsize, ncomp, compstride, lsize;
  CeedInt *offsets; /* NULL: strided */
  CeedInt strides[3];
  int refcount;
};
struct CeedBasis_private {
  Ceed ceed;
  CeedInt dim, ncomp, P, Q;
  double *interp, *grad, *qref, *qweight;
  int refcount;
};
typedef struct { char name[64]; CeedInt size; CeedEvalMode emode; } QField;
struct CeedQFunction_private {
  Ceed ceed;
  CeedQFunctionUser f;
  void *ctx;
  int nin, nout, identity, refcount;
  QField in[MAXF], out[MAXF];
};
typedef struct { CeedElemRestriction r; CeedBasis b; CeedVector v; int set; } OpField;
struct CeedOperator_private {
  Ceed ceed;
  CeedQFunction qf;
  OpField in[MAXF], out[MAXF];
  int composite, nsubs, refcount;
  CeedOperator subs[MAXF];
};

static struct CeedVector_private vector_active_, vector_none_;
static struct CeedBasis_private basis_collocated_;
static struct CeedElemRestriction_private rstr_none_;
static struct CeedQFunction_private qf_none_;
static struct CeedRequest_private req_immediate_, req_ordered_;
static CeedRequest req_immediate_p_ = &req_immediate_, req_ordered_p_ = &req_ordered_;
const CeedInt CEED_STRIDES_BACKEND[3] = {0, 0, 0};
const CeedBasis CEED_BASIS_COLLOCATED = &basis_collocated_;
const CeedVector CEED_VECTOR_ACTIVE = &vector_active_;
const CeedVector CEED_VECTOR_NONE = &vector_none_;
const CeedElemRestriction CEED_ELEMRESTRICTION_NONE = &rstr_none_;
const CeedQFunction CEED_QFUNCTION_NONE = &qf_none_;
CeedRequest *const CEED_REQUEST_IMMEDIATE = &req_immediate_p_;
CeedRequest *const CEED_REQUEST_ORDERED = &req_ordered_p_;
const char *const CeedMemTypes[] = {"host", "device", 0};
const char *const CeedEvalModes[] = {"none", "interpolation", "gradient", "", "divergence", "", "", "", "curl", "", "", "", "", "", "", "", "quadrature weights", 0};

static int fail(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "ceed_cpu (test-only CPU restatement of libCEED): ");
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
  abort();
  return 1;
}
static void *xcalloc(size_t n, size_t sz) {
  void *p = calloc(n ? n : 1, sz);
  if (!p) fail("out of memory");
  return p;
}

/* ------------------------------------------------------------------------------------------------ Ceed */
int CeedInit(const char *resource, Ceed *ceed) {
  *ceed = (Ceed)xcalloc(1, sizeof **ceed);
  snprintf((*ceed)->resource, sizeof (*ceed)->resource, "%s", resource ? resource : "/cpu/self");
  return 0;
}
int CeedDestroy(Ceed *ceed) { if (ceed && *ceed) { free(*ceed); *ceed = NULL; } return 0; }
int CeedGetResource(Ceed ceed, const char **resource) { *resource = ceed->resource; return 0; }
int CeedGetPreferredMemType(Ceed ceed, CeedMemType *type) { (void)ceed; *type = CEED_MEM_HOST; return 0; }
int CeedIsDeterministic(Ceed ceed, int *isDeterministic) { (void)ceed; *isDeterministic = 1; return 0; }

/* ------------------------------------------------------------------------------------------------ CeedVector */
static void need_host(CeedMemType m) { if (m != CEED_MEM_HOST) fail("CEED_MEM_DEVICE requested from the CPU restatement"); }
int CeedVectorCreate(Ceed ceed, CeedInt len, CeedVector *vec) {
  *vec = (CeedVector)xcalloc(1, sizeof **vec);
  (*vec)->ceed = ceed; (*vec)->n = len; (*vec)->refcount = 1;
  return 0;
}
static void vec_drop_array(CeedVector v) {
  if (v->a && v->owned) free(v->a);
  v->a = NULL; v->owned = 0; v->valid = 0;
}
static double *vec_alloc(CeedVector v) {
  if (!v->a) { v->a = (double *)xcalloc((size_t)v->n, sizeof(double)); v->owned = 1; }
  return v->a;
}
int CeedVectorSetArray(CeedVector vec, CeedMemType mtype, CeedCopyMode cmode, CeedScalar *array) {
  need_host(mtype);
  if (cmode == CEED_COPY_VALUES) {
    if (vec->a && !vec->owned) vec->a = NULL;
    memcpy(vec_alloc(vec), array, sizeof(double) * (size_t)vec->n);
  } else {
    vec_drop_array(vec);
    vec->a = array;
    vec->owned = cmode == CEED_OWN_POINTER;
  }
  vec->valid = 1;
  return 0;
}
int CeedVectorTakeArray(CeedVector vec, CeedMemType mtype, CeedScalar **array) {
  need_host(mtype);
  if (array) *array = vec->a;
  else if (vec->owned) free(vec->a);
  vec->a = NULL; vec->owned = 0; vec->valid = 0;
  return 0;
}
int CeedVectorSetValue(CeedVector vec, CeedScalar value) {
  double *a = vec_alloc(vec);
  for (CeedInt i = 0; i < vec->n; i++) a[i] = value;
  vec->valid = 1;
  return 0;
}
int CeedVectorSyncArray(CeedVector vec, CeedMemType mtype) { (void)vec; need_host(mtype); return 0; }
static double *vec_data(CeedVector v, const char *who) {
  if (!v || v == CEED_VECTOR_ACTIVE || v == CEED_VECTOR_NONE) fail("%s: not a data vector", who);
  if (!v->a || !v->valid) fail("%s: CeedVector has no valid data", who);
  return v->a;
}
int CeedVectorGetArray(CeedVector vec, CeedMemType mtype, CeedScalar **array) { need_host(mtype); *array = vec_data(vec, "CeedVectorGetArray"); return 0; }
int CeedVectorGetArrayRead(CeedVector vec, CeedMemType mtype, const CeedScalar **array) { need_host(mtype); *array = vec_data(vec, "CeedVectorGetArrayRead"); return 0; }
int CeedVectorRestoreArray(CeedVector vec, CeedScalar **array) { (void)vec; if (array) *array = NULL; return 0; }
int CeedVectorRestoreArrayRead(CeedVector vec, const CeedScalar **array) { (void)vec; if (array) *array = NULL; return 0; }
int CeedVectorNorm(CeedVector vec, CeedNormType type, CeedScalar *norm) {
  const double *a = vec_data(vec, "CeedVectorNorm");
  double s = 0;
  for (CeedInt i = 0; i < vec->n; i++) {
    if (type == CEED_NORM_1) s += fabs(a[i]);
    else if (type == CEED_NORM_2) s += a[i] * a[i];
    else if (fabs(a[i]) > s) s = fabs(a[i]);
  }
  *norm = type == CEED_NORM_2 ? sqrt(s) : s;
  return 0;
}
int CeedVectorReciprocal(CeedVector vec) {
  double *a = vec_data(vec, "CeedVectorReciprocal");
  for (CeedInt i = 0; i < vec->n; i++)
    if (fabs(a[i]) > 1e-14) a[i] = 1.0 / a[i];   /* upstream: entries with |x| <= CEED_EPSILON stay */
  return 0;
}
int CeedVectorGetLength(CeedVector vec, CeedInt *length) { *length = vec->n; return 0; }
int CeedVectorDestroy(CeedVector *vec) {
  if (!vec || !*vec || *vec == CEED_VECTOR_ACTIVE || *vec == CEED_VECTOR_NONE) return 0;
  if (--(*vec)->refcount == 0) { vec_drop_array(*vec); free(*vec); }
  *vec = NULL;
  return 0;
}

/* ------------------------------------------------------------------------------------------------ CeedElemRestriction (B.1) */
int CeedElemRestrictionCreate(Ceed ceed, CeedInt nelem, CeedInt elemsize, CeedInt ncomp, CeedInt compstride, CeedInt lsize,
                              CeedMemType mtype, CeedCopyMode cmode, const CeedInt *offsets, CeedElemRestriction *rstr) {
  need_host(mtype);
  CeedElemRestriction r = (CeedElemRestriction)xcalloc(1, sizeof *r);
  r->ceed = ceed; r->nelem = nelem; r->elemsize = elemsize; r->ncomp = ncomp; r->compstride = compstride; r->lsize = lsize;
  r->refcount = 1;
  r->offsets = (CeedInt *)xcalloc((size_t)nelem * elemsize, sizeof(CeedInt));
  memcpy(r->offsets, offsets, sizeof(CeedInt) * (size_t)nelem * elemsize);
  for (size_t i = 0; i < (size_t)nelem * elemsize; i++)
    if (offsets[i] < 0 || (long long)offsets[i] + (long long)(ncomp - 1) * compstride >= lsize)
      fail("CeedElemRestrictionCreate: offset %d outside the L-vector (lsize %d)", (int)offsets[i], (int)lsize);
  if (cmode == CEED_OWN_POINTER) free((void *)offsets);
  *rstr = r;
  return 0;
}
int CeedElemRestrictionCreateStrided(Ceed ceed, CeedInt nelem, CeedInt elemsize, CeedInt ncomp, CeedInt lsize,
                                     const CeedInt strides[3], CeedElemRestriction *rstr) {
  CeedElemRestriction r = (CeedElemRestriction)xcalloc(1, sizeof *r);
  r->ceed = ceed; r->nelem = nelem; r->elemsize = elemsize; r->ncomp = ncomp; r->compstride = 0; r->lsize = lsize;
  r->refcount = 1;
  if (strides == CEED_STRIDES_BACKEND || (strides[0] == 0 && strides[1] == 0 && strides[2] == 0)) {
    /* the CPU backends' choice: the L-vector IS the E-vector, [elem][comp][node] */
    r->strides[0] = 1; r->strides[1] = elemsize; r->strides[2] = elemsize * ncomp;
  } else {
    memcpy(r->strides, strides, sizeof r->strides);
  }
  if ((long long)nelem * elemsize * ncomp > lsize) fail("CeedElemRestrictionCreateStrided: L-vector too short");
  *rstr = r;
  return 0;
}
int CeedElemRestrictionCreateVector(CeedElemRestriction rstr, CeedVector *lvec, CeedVector *evec) {
  if (lvec) CeedVectorCreate(rstr->ceed, rstr->lsize, lvec);
  if (evec) CeedVectorCreate(rstr->ceed, rstr->nelem * rstr->elemsize * rstr->ncomp, evec);
  return 0;
}
static size_t l_index(CeedElemRestriction r, int e, int c, int n) {
  if (r->offsets) return (size_t)r->offsets[(size_t)e * r->elemsize + n] + (size_t)c * r->compstride;
  return (size_t)n * r->strides[0] + (size_t)c * r->strides[1] + (size_t)e * r->strides[2];
}
/* E[e][c][n] = L[...]   /   L[...] += E[e][c][n] in the serial order (e, c, n) */
static void rstr_apply(CeedElemRestriction r, int transpose, const double *u, double *v) {
  for (int e = 0; e < r->nelem; e++)
    for (int c = 0; c < r->ncomp; c++)
      for (int n = 0; n < r->elemsize; n++) {
        const size_t ei = ((size_t)e * r->ncomp + c) * r->elemsize + n, li = l_index(r, e, c, n);
        if (!transpose) v[ei] = u[li];
        else v[li] += u[ei];
      }
}
int CeedElemRestrictionApply(CeedElemRestriction rstr, CeedTransposeMode tmode, CeedVector u, CeedVector ru, CeedRequest *request) {
  (void)request;
  const double *x = vec_data(u, "CeedElemRestrictionApply");
  if (tmode == CEED_NOTRANSPOSE) { vec_alloc(ru); ru->valid = 1; }
  rstr_apply(rstr, tmode == CEED_TRANSPOSE, x, vec_data(ru, "CeedElemRestrictionApply"));
  return 0;
}
/* transpose-apply of an E-vector of ones (misc.c:117-123 zeroes the target first; so does upstream) */
int CeedElemRestrictionGetMultiplicity(CeedElemRestriction rstr, CeedVector mult) {
  const size_t n = (size_t)rstr->nelem * rstr->elemsize * rstr->ncomp;
  double *ones = (double *)xcalloc(n, sizeof(double));
  for (size_t i = 0; i < n; i++) ones[i] = 1.0;
  CeedVectorSetValue(mult, 0.0);
  rstr_apply(rstr, 1, ones, mult->a);
  free(ones);
  return 0;
}
int CeedElemRestrictionGetNumElements(CeedElemRestriction rstr, CeedInt *numelem) { *numelem = rstr->nelem; return 0; }
int CeedElemRestrictionGetElementSize(CeedElemRestriction rstr, CeedInt *elemsize) { *elemsize = rstr->elemsize; return 0; }
int CeedElemRestrictionGetLVectorSize(CeedElemRestriction rstr, CeedInt *lsize) { *lsize = rstr->lsize; return 0; }
int CeedElemRestrictionGetNumComponents(CeedElemRestriction rstr, CeedInt *ncomp) { *ncomp = rstr->ncomp; return 0; }
int CeedElemRestrictionDestroy(CeedElemRestriction *rstr) {
  if (!rstr || !*rstr || *rstr == CEED_ELEMRESTRICTION_NONE) return 0;
  if (--(*rstr)->refcount == 0) { free((*rstr)->offsets); free(*rstr); }
  *rstr = NULL;
  return 0;
}

/* ------------------------------------------------------------------------------------------------ CeedBasis (B.2, B.3) */
int CeedGaussQuadrature(CeedInt Q, CeedScalar *qref1d, CeedScalar *qweight1d) { return oracle_gauss(Q, qref1d, qweight1d); }
int CeedLobattoQuadrature(CeedInt Q, CeedScalar *qref1d, CeedScalar *qweight1d) { return oracle_lobatto(Q, qref1d, qweight1d); }
int CeedBasisCreateTensorH1(Ceed ceed, CeedInt dim, CeedInt ncomp, CeedInt P1d, CeedInt Q1d, const CeedScalar *interp1d,
                            const CeedScalar *grad1d, const CeedScalar *qref1d, const CeedScalar *qweight1d, CeedBasis *basis) {
  if (dim != 3) fail("CeedBasisCreateTensorH1: dim %d (the reference is 3-D only)", (int)dim);
  CeedBasis b = (CeedBasis)xcalloc(1, sizeof *b);
  b->ceed = ceed; b->dim = dim; b->ncomp = ncomp; b->P = P1d; b->Q = Q1d; b->refcount = 1;
  b->interp = (double *)xcalloc((size_t)P1d * Q1d, sizeof(double));
  b->grad = (double *)xcalloc((size_t)P1d * Q1d, sizeof(double));
  b->qref = (double *)xcalloc((size_t)Q1d, sizeof(double));
  b->qweight = (double *)xcalloc((size_t)Q1d, sizeof(double));
  memcpy(b->interp, interp1d, sizeof(double) * (size_t)P1d * Q1d);
  memcpy(b->grad, grad1d, sizeof(double) * (size_t)P1d * Q1d);
  memcpy(b->qref, qref1d, sizeof(double) * (size_t)Q1d);
  memcpy(b->qweight, qweight1d, sizeof(double) * (size_t)Q1d);
  *basis = b;
  return 0;
}
int CeedBasisCreateTensorH1Lagrange(Ceed ceed, CeedInt dim, CeedInt ncomp, CeedInt P, CeedInt Q, CeedQuadMode qmode, CeedBasis *basis) {
  double *w = (double *)xcalloc(2 * (size_t)P * Q + 2 * (size_t)Q, sizeof(double));
  double *B = w, *D = B + (size_t)P * Q, *qr = D + (size_t)P * Q, *qw = qr + Q;
  if (oracle_basis_1d(P, Q, qmode == CEED_GAUSS_LOBATTO, B, D, qr, qw)) fail("CeedBasisCreateTensorH1Lagrange(%d,%d) failed", (int)P, (int)Q);
  CeedBasisCreateTensorH1(ceed, dim, ncomp, P, Q, B, D, qr, qw, basis);
  free(w);
  return 0;
}
static size_t qsize(CeedBasis b, CeedEvalMode emode) {
  const size_t Q3 = (size_t)b->Q * b->Q * b->Q;
  return emode == CEED_EVAL_GRAD ? 3 * b->ncomp * Q3 : (emode == CEED_EVAL_WEIGHT ? Q3 : b->ncomp * Q3);
}
/* E-vector [elem][comp][P^3]; Q-vector [elem][per-element layout of B.3] */
int CeedBasisApply(CeedBasis basis, CeedInt nelem, CeedTransposeMode tmode, CeedEvalMode emode, CeedVector u, CeedVector v) {
  const size_t nsz = (size_t)basis->ncomp * basis->P * basis->P * basis->P, qsz = qsize(basis, emode);
  const double *x = emode == CEED_EVAL_WEIGHT ? NULL : vec_data(u, "CeedBasisApply");
  vec_alloc(v); v->valid = 1;
  for (int e = 0; e < nelem; e++) {
    int r;
    if (emode == CEED_EVAL_WEIGHT) r = oracle_basis_apply_elem(basis->ncomp, basis->P, basis->Q, basis->interp, basis->grad, basis->qweight, 0, 4, NULL, v->a + e * qsz);
    else if (tmode == CEED_NOTRANSPOSE) r = oracle_basis_apply_elem(basis->ncomp, basis->P, basis->Q, basis->interp, basis->grad, basis->qweight, 0, (int)emode, x + e * nsz, v->a + e * qsz);
    else r = oracle_basis_apply_elem(basis->ncomp, basis->P, basis->Q, basis->interp, basis->grad, basis->qweight, 1, (int)emode, x + e * qsz, v->a + e * nsz);
    if (r) fail("CeedBasisApply: eval mode %d not restated", (int)emode);
  }
  return 0;
}
int CeedBasisGetNumNodes(CeedBasis basis, CeedInt *P) { *P = basis->P * basis->P * basis->P; return 0; }
int CeedBasisGetNumQuadraturePoints(CeedBasis basis, CeedInt *Q) { *Q = basis->Q * basis->Q * basis->Q; return 0; }
int CeedBasisGetInterp1D(CeedBasis basis, const CeedScalar **interp1d) { *interp1d = basis->interp; return 0; }
int CeedBasisGetGrad1D(CeedBasis basis, const CeedScalar **grad1d) { *grad1d = basis->grad; return 0; }
int CeedBasisGetQRef(CeedBasis basis, const CeedScalar **qref) { *qref = basis->qref; return 0; }
int CeedBasisGetQWeights(CeedBasis basis, const CeedScalar **qweight) { *qweight = basis->qweight; return 0; }
int CeedBasisDestroy(CeedBasis *basis) {
  if (!basis || !*basis || *basis == CEED_BASIS_COLLOCATED) return 0;
  if (--(*basis)->refcount == 0) { free((*basis)->interp); free((*basis)->grad); free((*basis)->qref); free((*basis)->qweight); free(*basis); }
  *basis = NULL;
  return 0;
}

/* ------------------------------------------------------------------------------------------------ CeedQFunction */
int CeedQFunctionCreateInterior(Ceed ceed, CeedInt vlength, CeedQFunctionUser f, const char *source, CeedQFunction *qf) {
  (void)vlength; (void)source;
  if (!f) fail("CeedQFunctionCreateInterior: NULL function pointer");
  *qf = (CeedQFunction)xcalloc(1, sizeof **qf);
  (*qf)->ceed = ceed; (*qf)->f = f; (*qf)->refcount = 1;
  return 0;
}
int CeedQFunctionAddInput(CeedQFunction qf, const char *fieldname, CeedInt size, CeedEvalMode emode) {
  if (qf->nin == MAXF) fail("too many QFunction inputs");
  QField *fl = &qf->in[qf->nin++];
  snprintf(fl->name, sizeof fl->name, "%s", fieldname); fl->size = size; fl->emode = emode;
  return 0;
}
int CeedQFunctionAddOutput(CeedQFunction qf, const char *fieldname, CeedInt size, CeedEvalMode emode) {
  if (qf->nout == MAXF) fail("too many QFunction outputs");
  QField *fl = &qf->out[qf->nout++];
  snprintf(fl->name, sizeof fl->name, "%s", fieldname); fl->size = size; fl->emode = emode;
  return 0;
}
/* gallery "Identity": fields "input" / "output" of `size` values per point (elasticity.c:249-252) */
int CeedQFunctionCreateIdentity(Ceed ceed, CeedInt size, CeedEvalMode inmode, CeedEvalMode outmode, CeedQFunction *qf) {
  *qf = (CeedQFunction)xcalloc(1, sizeof **qf);
  (*qf)->ceed = ceed; (*qf)->identity = 1; (*qf)->refcount = 1;
  CeedQFunctionAddInput(*qf, "input", size, inmode);
  CeedQFunctionAddOutput(*qf, "output", size, outmode);
  return 0;
}
/* the pointer is kept, not the bytes (GetDiag_Ceed swaps it, matops.c:216-235) */
int CeedQFunctionSetContext(CeedQFunction qf, void *ctx, size_t ctxsize) { (void)ctxsize; qf->ctx = ctx; return 0; }
int CeedQFunctionDestroy(CeedQFunction *qf) {
  if (!qf || !*qf || *qf == CEED_QFUNCTION_NONE) return 0;
  if (--(*qf)->refcount == 0) free(*qf);
  *qf = NULL;
  return 0;
}

/* ------------------------------------------------------------------------------------------------ CeedOperator (B.4, B.5) */
int CeedOperatorCreate(Ceed ceed, CeedQFunction qf, CeedQFunction dqf, CeedQFunction dqfT, CeedOperator *op) {
  (void)dqf; (void)dqfT;
  *op = (CeedOperator)xcalloc(1, sizeof **op);
  (*op)->ceed = ceed; (*op)->qf = qf; (*op)->refcount = 1;
  qf->refcount++;
  return 0;
}
int CeedCompositeOperatorCreate(Ceed ceed, CeedOperator *op) {
  *op = (CeedOperator)xcalloc(1, sizeof **op);
  (*op)->ceed = ceed; (*op)->composite = 1; (*op)->refcount = 1;
  return 0;
}
int CeedCompositeOperatorAddSub(CeedOperator compositeop, CeedOperator subop) {
  if (!compositeop->composite || compositeop->nsubs == MAXF) fail("CeedCompositeOperatorAddSub: not a composite / too many");
  compositeop->subs[compositeop->nsubs++] = subop;
  subop->refcount++;
  return 0;
}
int CeedOperatorSetField(CeedOperator op, const char *fieldname, CeedElemRestriction r, CeedBasis b, CeedVector v) {
  CeedQFunction qf = op->qf;
  OpField *f = NULL;
  for (int i = 0; i < qf->nin && !f; i++) if (!strcmp(qf->in[i].name, fieldname)) f = &op->in[i];
  for (int i = 0; i < qf->nout && !f; i++) if (!strcmp(qf->out[i].name, fieldname)) f = &op->out[i];
  if (!f) fail("CeedOperatorSetField: the QFunction has no field \"%s\"", fieldname);
  if (f->set) fail("CeedOperatorSetField: field \"%s\" set twice", fieldname);
  f->r = r; f->b = b; f->v = v; f->set = 1;
  if (r && r != CEED_ELEMRESTRICTION_NONE) r->refcount++;
  if (b && b != CEED_BASIS_COLLOCATED) b->refcount++;
  if (v && v != CEED_VECTOR_ACTIVE && v != CEED_VECTOR_NONE) v->refcount++;
  return 0;
}

typedef struct { int nelem, Q3; } OpShape;
static OpShape op_shape(CeedOperator op) {
  CeedQFunction qf = op->qf;
  OpShape s = {-1, -1};
  for (int pass = 0; pass < 2; pass++) {
    const int n = pass ? qf->nout : qf->nin;
    for (int i = 0; i < n; i++) {
      OpField *f = pass ? &op->out[i] : &op->in[i];
      const QField *q = pass ? &qf->out[i] : &qf->in[i];
      if (!f->set) fail("CeedOperatorApply: field \"%s\" was never set", q->name);
      int Q3 = -1, ne = -1;
      if (f->b && f->b != CEED_BASIS_COLLOCATED) Q3 = f->b->Q * f->b->Q * f->b->Q;
      else if (f->r && f->r != CEED_ELEMRESTRICTION_NONE) Q3 = f->r->elemsize;
      if (f->r && f->r != CEED_ELEMRESTRICTION_NONE) ne = f->r->nelem;
      if (Q3 >= 0) { if (s.Q3 < 0) s.Q3 = Q3; else if (s.Q3 != Q3) fail("field \"%s\": %d quadrature points, other fields %d", q->name, Q3, s.Q3); }
      if (ne >= 0) { if (s.nelem < 0) s.nelem = ne; else if (s.nelem != ne) fail("field \"%s\": %d elements, other fields %d", q->name, ne, s.nelem); }
      if (q->emode == CEED_EVAL_WEIGHT) continue;
      /* size bookkeeping exactly as upstream checks it */
      const int ncomp = f->r->ncomp;
      const int expect = q->emode == CEED_EVAL_GRAD ? 3 * ncomp : ncomp;
      if (expect != q->size) fail("field \"%s\": QFunction size %d, restriction/basis give %d", q->name, (int)q->size, expect);
      if (q->emode != CEED_EVAL_NONE) {
        if (!f->b || f->b == CEED_BASIS_COLLOCATED) fail("field \"%s\": eval mode needs a basis", q->name);
        if (f->b->P * f->b->P * f->b->P != f->r->elemsize || f->b->ncomp != ncomp) fail("field \"%s\": basis and restriction disagree", q->name);
      }
    }
  }
  if (s.nelem < 0 || s.Q3 < 0) fail("CeedOperatorApply: cannot determine the element / point counts");
  return s;
}

static CeedVector field_vec(OpField *f, CeedVector active) { return f->v == CEED_VECTOR_ACTIVE ? active : f->v; }

/* out += A(in): restrict every input, loop over elements (basis, user QFunction with Q = all points of the element
 * and arrays in field declaration order, transposed basis), transposed-restrict every output */
static void op_apply_add(CeedOperator op, CeedVector in, CeedVector out) {
  if (op->composite) {
    for (int i = 0; i < op->nsubs; i++) op_apply_add(op->subs[i], in, out);
    return;
  }
  CeedQFunction qf = op->qf;
  const OpShape s = op_shape(op);
  double *ein[MAXF] = {0}, *eout[MAXF] = {0}, *qin[MAXF] = {0}, *qout[MAXF] = {0};
  for (int i = 0; i < qf->nin; i++) {
    OpField *f = &op->in[i];
    if (qf->in[i].emode == CEED_EVAL_WEIGHT) { qin[i] = (double *)xcalloc((size_t)s.Q3, sizeof(double)); continue; }
    ein[i] = (double *)xcalloc((size_t)s.nelem * f->r->elemsize * f->r->ncomp, sizeof(double));
    rstr_apply(f->r, 0, vec_data(field_vec(f, in), qf->in[i].name), ein[i]);
    if (qf->in[i].emode != CEED_EVAL_NONE) qin[i] = (double *)xcalloc(qsize(f->b, qf->in[i].emode), sizeof(double));
  }
  for (int i = 0; i < qf->nout; i++) {
    OpField *f = &op->out[i];
    eout[i] = (double *)xcalloc((size_t)s.nelem * f->r->elemsize * f->r->ncomp, sizeof(double));
    if (qf->out[i].emode != CEED_EVAL_NONE) qout[i] = (double *)xcalloc(qsize(f->b, qf->out[i].emode), sizeof(double));
  }
  for (int e = 0; e < s.nelem; e++) {
    const double *qfi[MAXF];
    double *qfo[MAXF];
    for (int i = 0; i < qf->nin; i++) {
      OpField *f = &op->in[i];
      const CeedEvalMode em = qf->in[i].emode;
      if (em == CEED_EVAL_NONE) { qfi[i] = ein[i] + (size_t)e * f->r->elemsize * f->r->ncomp; continue; }
      CeedBasis b = f->b;
      const double *ue = em == CEED_EVAL_WEIGHT ? NULL : ein[i] + (size_t)e * f->r->elemsize * f->r->ncomp;
      if (oracle_basis_apply_elem(b->ncomp, b->P, b->Q, b->interp, b->grad, b->qweight, 0, (int)em == 16 ? 4 : (int)em, ue, qin[i]))
        fail("field \"%s\": eval mode %d not restated", qf->in[i].name, (int)em);
      qfi[i] = qin[i];
    }
    for (int i = 0; i < qf->nout; i++) {
      OpField *f = &op->out[i];
      qfo[i] = qf->out[i].emode == CEED_EVAL_NONE ? eout[i] + (size_t)e * f->r->elemsize * f->r->ncomp : qout[i];
    }
    if (qf->identity) memcpy(qfo[0], qfi[0], sizeof(double) * (size_t)qf->in[0].size * s.Q3);
    else if (qf->f(qf->ctx, s.Q3, qfi, qfo)) fail("user QFunction returned an error");
    for (int i = 0; i < qf->nout; i++) {
      OpField *f = &op->out[i];
      const CeedEvalMode em = qf->out[i].emode;
      if (em == CEED_EVAL_NONE) continue;
      CeedBasis b = f->b;
      if (oracle_basis_apply_elem(b->ncomp, b->P, b->Q, b->interp, b->grad, b->qweight, 1, (int)em, qout[i],
                                  eout[i] + (size_t)e * f->r->elemsize * f->r->ncomp))
        fail("field \"%s\": eval mode %d not restated", qf->out[i].name, (int)em);
    }
  }
  for (int i = 0; i < qf->nout; i++) {
    OpField *f = &op->out[i];
    CeedVector v = field_vec(f, out);
    if (!v->a) CeedVectorSetValue(v, 0.0);
    v->valid = 1;
    rstr_apply(f->r, 1, eout[i], v->a);
  }
  for (int i = 0; i < MAXF; i++) { free(ein[i]); free(eout[i]); free(qin[i]); free(qout[i]); }
}
/* every output vector (active and passive) is zeroed, then ApplyAdd */
static void op_zero_outputs(CeedOperator op, CeedVector out) {
  if (op->composite) {
    if (out && out != CEED_VECTOR_NONE) CeedVectorSetValue(out, 0.0);
    for (int i = 0; i < op->nsubs; i++)
      for (int k = 0; k < op->subs[i]->qf->nout; k++) {
        CeedVector v = op->subs[i]->out[k].v;
        if (v && v != CEED_VECTOR_ACTIVE && v != CEED_VECTOR_NONE) CeedVectorSetValue(v, 0.0);
      }
    return;
  }
  for (int k = 0; k < op->qf->nout; k++) {
    CeedVector v = field_vec(&op->out[k], out);
    if (v && v != CEED_VECTOR_NONE) CeedVectorSetValue(v, 0.0);
  }
}
int CeedOperatorApply(CeedOperator op, CeedVector in, CeedVector out, CeedRequest *request) {
  (void)request;
  op_zero_outputs(op, out);
  op_apply_add(op, in, out);
  return 0;
}
int CeedOperatorApplyAdd(CeedOperator op, CeedVector in, CeedVector out, CeedRequest *request) {
  (void)request;
  op_apply_add(op, in, out);
  return 0;
}

/* B.5: unit fields through the QFunction give Dq[(slot in)][(slot out)][q]; diag_e[c][n] = sum_q sum_{din,dout}
 * G_dout[q,n] Dq[(din,c)][(dout,c)][q] G_din[q,n] with G_d the dense 3-D matrices of the evaluation modes of the
 * active fields (GRAD: three derivative directions; INTERP: one); entries with |Dq| <= 1e-12 max|Dq| are skipped as
 * upstream does; scatter-add with the active restriction. */
static void dense_eval_matrix(CeedBasis b, int d /* -1: interp */, double *G) {
  const int P = b->P, Q = b->Q, P3 = P * P * P;
  for (int qz = 0; qz < Q; qz++)
    for (int qy = 0; qy < Q; qy++)
      for (int qx = 0; qx < Q; qx++)
        for (int k = 0; k < P; k++)
          for (int j = 0; j < P; j++)
            for (int i = 0; i < P; i++)
              G[((size_t)(qz * Q + qy) * Q + qx) * P3 + (k * P + j) * P + i] =
                  (d == 0 ? b->grad : b->interp)[qx * P + i] * (d == 1 ? b->grad : b->interp)[qy * P + j] * (d == 2 ? b->grad : b->interp)[qz * P + k];
}
static void op_diagonal_add(CeedOperator op, CeedVector assembled) {
  if (op->composite) {
    for (int i = 0; i < op->nsubs; i++) op_diagonal_add(op->subs[i], assembled);
    return;
  }
  CeedQFunction qf = op->qf;
  const OpShape s = op_shape(op);
  /* active slots: (field, direction, component) for inputs and outputs */
  typedef struct { int field, d, c; } Slot;
  Slot slin[64], slout[64];
  int nsin = 0, nsout = 0;
  CeedElemRestriction ract = NULL;
  CeedBasis bact = NULL;
  for (int pass = 0; pass < 2; pass++)
    for (int i = 0; i < (pass ? qf->nout : qf->nin); i++) {
      OpField *f = pass ? &op->out[i] : &op->in[i];
      const QField *q = pass ? &qf->out[i] : &qf->in[i];
      if (f->v != CEED_VECTOR_ACTIVE) continue;
      if (q->emode != CEED_EVAL_GRAD && q->emode != CEED_EVAL_INTERP) fail("diagonal: active field \"%s\" must be INTERP or GRAD", q->name);
      if (ract && (ract != f->r || bact != f->b)) fail("diagonal: active fields on different restrictions / bases");
      ract = f->r; bact = f->b;
      const int nd = q->emode == CEED_EVAL_GRAD ? 3 : 1;
      for (int d = 0; d < nd; d++)
        for (int c = 0; c < f->r->ncomp; c++) {
          Slot sl = {i, q->emode == CEED_EVAL_GRAD ? d : -1, c};
          if ((pass ? nsout : nsin) == 64) fail("diagonal: too many active slots");
          if (pass) slout[nsout++] = sl; else slin[nsin++] = sl;
        }
    }
  if (!ract) fail("diagonal: no active field");
  const int P3 = ract->elemsize, ncomp = ract->ncomp, Q3 = s.Q3;
  double *G[4];
  for (int d = 0; d < 4; d++) { G[d] = (double *)xcalloc((size_t)Q3 * P3, sizeof(double)); dense_eval_matrix(bact, d - 1, G[d]); }
  /* passive inputs restricted once; active inputs are the unit fields */
  double *ein[MAXF] = {0}, *qin[MAXF] = {0}, *qout[MAXF] = {0};
  for (int i = 0; i < qf->nin; i++) {
    OpField *f = &op->in[i];
    const CeedEvalMode em = qf->in[i].emode;
    if (f->v == CEED_VECTOR_ACTIVE) { qin[i] = (double *)xcalloc((size_t)qf->in[i].size * Q3, sizeof(double)); continue; }
    if (em == CEED_EVAL_WEIGHT) { qin[i] = (double *)xcalloc((size_t)Q3, sizeof(double)); continue; }
    ein[i] = (double *)xcalloc((size_t)s.nelem * f->r->elemsize * f->r->ncomp, sizeof(double));
    rstr_apply(f->r, 0, vec_data(f->v, qf->in[i].name), ein[i]);
    if (em != CEED_EVAL_NONE) qin[i] = (double *)xcalloc(qsize(f->b, em), sizeof(double));
  }
  for (int i = 0; i < qf->nout; i++) qout[i] = (double *)xcalloc((size_t)qf->out[i].size * Q3, sizeof(double));
  double *Dq = (double *)xcalloc((size_t)nsin * nsout * Q3, sizeof(double));
  double *Ed = (double *)xcalloc((size_t)s.nelem * ncomp * P3, sizeof(double));
  for (int e = 0; e < s.nelem; e++) {
    const double *qfi[MAXF];
    double *qfo[MAXF];
    for (int i = 0; i < qf->nin; i++) {
      OpField *f = &op->in[i];
      const CeedEvalMode em = qf->in[i].emode;
      if (f->v == CEED_VECTOR_ACTIVE) { qfi[i] = qin[i]; continue; }
      if (em == CEED_EVAL_NONE) { qfi[i] = ein[i] + (size_t)e * f->r->elemsize * f->r->ncomp; continue; }
      const double *ue = em == CEED_EVAL_WEIGHT ? NULL : ein[i] + (size_t)e * f->r->elemsize * f->r->ncomp;
      oracle_basis_apply_elem(f->b->ncomp, f->b->P, f->b->Q, f->b->interp, f->b->grad, f->b->qweight, 0, (int)em == 16 ? 4 : (int)em, ue, qin[i]);
      qfi[i] = qin[i];
    }
    for (int i = 0; i < qf->nout; i++) qfo[i] = qout[i];
    double dmax = 0;
    for (int a = 0; a < nsin; a++) {
      for (int i = 0; i < qf->nin; i++)
        if (op->in[i].v == CEED_VECTOR_ACTIVE) memset(qin[i], 0, sizeof(double) * (size_t)qf->in[i].size * Q3);
      const int slot = (slin[a].d < 0 ? 0 : slin[a].d) * ncomp + slin[a].c;          /* [d][c][q] */
      for (int q = 0; q < Q3; q++) qin[slin[a].field][(size_t)slot * Q3 + q] = 1.0;
      if (qf->identity) memcpy(qfo[0], qfi[0], sizeof(double) * (size_t)qf->in[0].size * Q3);
      else if (qf->f(qf->ctx, Q3, qfi, qfo)) fail("user QFunction returned an error");
      for (int b = 0; b < nsout; b++) {
        const int oslot = (slout[b].d < 0 ? 0 : slout[b].d) * ncomp + slout[b].c;
        const double *src = qout[slout[b].field] + (size_t)oslot * Q3;
        double *dst = Dq + ((size_t)a * nsout + b) * Q3;
        for (int q = 0; q < Q3; q++) { dst[q] = src[q]; if (fabs(src[q]) > dmax) dmax = fabs(src[q]); }
      }
    }
    const double thresh = 1e-12 * dmax;
    for (int a = 0; a < nsin; a++)
      for (int b = 0; b < nsout; b++) {
        if (slin[a].c != slout[b].c) continue;
        const double *D = Dq + ((size_t)a * nsout + b) * Q3, *gi = G[slin[a].d + 1], *go = G[slout[b].d + 1];
        double *de = Ed + ((size_t)e * ncomp + slin[a].c) * P3;
        for (int q = 0; q < Q3; q++) {
          if (fabs(D[q]) <= thresh) continue;
          for (int n = 0; n < P3; n++) de[n] += go[(size_t)q * P3 + n] * D[q] * gi[(size_t)q * P3 + n];
        }
      }
  }
  if (!assembled->a) CeedVectorSetValue(assembled, 0.0);
  assembled->valid = 1;
  rstr_apply(ract, 1, Ed, assembled->a);
  for (int i = 0; i < MAXF; i++) { free(ein[i]); free(qin[i]); free(qout[i]); }
  for (int d = 0; d < 4; d++) free(G[d]);
  free(Dq); free(Ed);
}
int CeedOperatorLinearAssembleAddDiagonal(CeedOperator op, CeedVector assembled, CeedRequest *request) {
  (void)request;
  op_diagonal_add(op, assembled);
  return 0;
}
int CeedOperatorLinearAssembleDiagonal(CeedOperator op, CeedVector assembled, CeedRequest *request) {
  (void)request;
  CeedVectorSetValue(assembled, 0.0);
  op_diagonal_add(op, assembled);
  return 0;
}
int CeedOperatorDestroy(CeedOperator *op) {
  if (!op || !*op) return 0;
  CeedOperator o = *op;
  *op = NULL;
  if (--o->refcount) return 0;
  if (o->composite) {
    for (int i = 0; i < o->nsubs; i++) CeedOperatorDestroy(&o->subs[i]);
  } else {
    for (int pass = 0; pass < 2; pass++)
      for (int i = 0; i < (pass ? o->qf->nout : o->qf->nin); i++) {
        OpField *f = pass ? &o->out[i] : &o->in[i];
        if (!f->set) continue;
        CeedElemRestrictionDestroy(&f->r);
        CeedBasisDestroy(&f->b);
        CeedVectorDestroy(&f->v);
      }
    CeedQFunctionDestroy(&o->qf);
  }
  free(o);
  return 0;
}
