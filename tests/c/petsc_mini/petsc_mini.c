/* TEST-ONLY functional miniature of PETSc for tests/c/ref_driver.c -- see petsc.h in this directory.
 * Device-resident Vecs (the VECCUDA path of -memtype device) are moved with the backend's thin C-ABI CUDA layer
 * (include/b200_kernels.h): this file contains no CUDA. */
#include "petsc.h"

#include <stdarg.h>

#include "b200_kernels.h"

#define B2CHK(call) do { int rc_ = (call); if (rc_) { fprintf(stderr, "petsc_mini: %s: %s\n", #call, b200_last_error()); return 77; } } while (0)

static PetscErrorCode vec_new(PetscInt n, int device, DM dm, Vec *out) {
  Vec v = (Vec)calloc(1, sizeof *v);
  if (!v) return 55;
  v->n = n; v->device = device; v->dm = dm;
  if (device) {
    B2CHK(b200_malloc((void **)&v->a, sizeof(double) * (size_t)(n > 0 ? n : 1)));
    B2CHK(b200_memset(v->a, 0, sizeof(double) * (size_t)n));
  } else {
    v->a = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    if (!v->a) return 55;
  }
  *out = v;
  return 0;
}

/* ---------------------------------------------------------------- Vec */
PetscErrorCode VecZeroEntries(Vec v) {
  if (v->device) B2CHK(b200_memset(v->a, 0, sizeof(double) * (size_t)v->n));
  else memset(v->a, 0, sizeof(double) * (size_t)v->n);
  return 0;
}
PetscErrorCode VecDuplicate(Vec v, Vec *out) { return vec_new(v->n, v->device, v->dm, out); }
PetscErrorCode VecDestroy(Vec *v) {
  if (!v || !*v) return 0;
  if ((*v)->device) b200_free((*v)->a);
  else free((*v)->a);
  free(*v);
  *v = NULL;
  return 0;
}
PetscErrorCode VecGetSize(Vec v, PetscInt *n) { *n = v->n; return 0; }
PetscErrorCode VecGetLocalSize(Vec v, PetscInt *n) { *n = v->n; return 0; }
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y) {
  if (w->device) B2CHK(b200_vec_pointwise_mult(w->a, x->a, y->a, (size_t)w->n));
  else
    for (PetscInt i = 0; i < w->n; i++) w->a[i] = x->a[i] * y->a[i];
  return 0;
}
/* PETSc: x[i] = 1/x[i] where x[i] != 0 */
PetscErrorCode VecReciprocal(Vec v) {
  if (v->device) B2CHK(b200_vec_reciprocal(v->a, (size_t)v->n));
  else
    for (PetscInt i = 0; i < v->n; i++)
      if (v->a[i] != 0.0) v->a[i] = 1.0 / v->a[i];
  return 0;
}
static PetscErrorCode get_host(Vec v, PetscScalar **a) {
  if (v->device) { fprintf(stderr, "petsc_mini: VecGetArray on a device Vec\n"); return 56; }
  *a = v->a;
  return 0;
}
static PetscErrorCode get_dev(Vec v, PetscScalar **a) {
  if (!v->device) { fprintf(stderr, "petsc_mini: VecCUDAGetArray on a host Vec\n"); return 56; }
  *a = v->a;
  return 0;
}
PetscErrorCode VecGetArray(Vec v, PetscScalar **a) { return get_host(v, a); }
PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **a) { return get_host(v, (PetscScalar **)a); }
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a) { (void)v; if (a) *a = NULL; return 0; }
PetscErrorCode VecRestoreArrayRead(Vec v, const PetscScalar **a) { (void)v; if (a) *a = NULL; return 0; }
PetscErrorCode VecCUDAGetArray(Vec v, PetscScalar **a) { return get_dev(v, a); }
PetscErrorCode VecCUDAGetArrayRead(Vec v, const PetscScalar **a) { return get_dev(v, (PetscScalar **)a); }
PetscErrorCode VecCUDARestoreArray(Vec v, PetscScalar **a) { (void)v; if (a) *a = NULL; return 0; }
PetscErrorCode VecCUDARestoreArrayRead(Vec v, const PetscScalar **a) { (void)v; if (a) *a = NULL; return 0; }
PetscErrorCode VecGetDM(Vec v, DM *dm) { *dm = v->dm; return 0; }
/* the "VTK viewer": keeps a host copy of the last vector viewed, for the test driver to write out */
double *petsc_mini_viewed = NULL;
PetscInt petsc_mini_viewed_n = 0;
PetscErrorCode VecView(Vec v, PetscViewer viewer) {
  (void)viewer;
  free(petsc_mini_viewed);
  petsc_mini_viewed = (double *)malloc(sizeof(double) * (size_t)(v->n ? v->n : 1));
  petsc_mini_viewed_n = v->n;
  if (v->device) B2CHK(b200_memcpy_d2h(petsc_mini_viewed, v->a, sizeof(double) * (size_t)v->n));
  else memcpy(petsc_mini_viewed, v->a, sizeof(double) * (size_t)v->n);
  return 0;
}

/* ---------------------------------------------------------------- DM */
PetscErrorCode DMGetDimension(DM dm, PetscInt *dim) { *dim = dm->dim; return 0; }
PetscErrorCode DMGetSection(DM dm, PetscSection *s) { (void)dm; *s = NULL; return 0; }
PetscErrorCode DMGetCoordinateDM(DM dm, DM *cdm) { *cdm = dm->coordDM; return 0; }
PetscErrorCode DMGetCoordinatesLocal(DM dm, Vec *c) { *c = dm->coords; return 0; }
PetscErrorCode DMGetLocalVector(DM dm, Vec *v) { return vec_new(dm->lsize, 0, dm, v); }
PetscErrorCode DMRestoreLocalVector(DM dm, Vec *v) { (void)dm; return VecDestroy(v); }
PetscErrorCode DMCreateLocalVector(DM dm, Vec *v) { return vec_new(dm->lsize, dm->device, dm, v); }
PetscErrorCode DMCreateGlobalVector(DM dm, Vec *v) { return vec_new(dm->gsize, dm->device, dm, v); }
PetscErrorCode DMSetOutputSequenceNumber(DM dm, PetscInt n, PetscReal t) { (void)dm; (void)n; (void)t; return 0; }
PetscErrorCode DMPlexGetHeightStratum(DM dm, PetscInt h, PetscInt *s, PetscInt *e) { (void)h; *s = 0; *e = dm->nelem; return 0; }
PetscErrorCode DMPlexSetClosurePermutationTensor(DM dm, PetscInt p, PetscSection s) { (void)dm; (void)p; (void)s; return 0; }
PetscErrorCode DMPlexGetClosureIndices(DM dm, PetscSection s0, PetscSection s1, PetscInt cell, PetscBool useClPerm,
                                       PetscInt *numindices, PetscInt **indices, PetscInt *outOffsets, PetscScalar **values) {
  (void)s0; (void)s1; (void)useClPerm; (void)outOffsets; (void)values;
  const PetscInt n = dm->P * dm->P * dm->P * dm->ncomp;
  *numindices = n;
  *indices = dm->closure + (size_t)cell * n;
  return 0;
}
PetscErrorCode DMPlexRestoreClosureIndices(DM dm, PetscSection s0, PetscSection s1, PetscInt cell, PetscBool useClPerm,
                                           PetscInt *numindices, PetscInt **indices, PetscInt *outOffsets, PetscScalar **values) {
  (void)dm; (void)s0; (void)s1; (void)cell; (void)useClPerm; (void)numindices; (void)outOffsets; (void)values;
  *indices = NULL;
  return 0;
}
/* INSERT_VALUES: the owned unconstrained dofs of the global vector into the local one; other local entries untouched */
PetscErrorCode DMGlobalToLocal(DM dm, Vec g, InsertMode mode, Vec l) {
  if (mode != INSERT_VALUES) return 56;
  if (l->device) B2CHK(b200_scatter_set(l->a, dm->d_l2g_loc, g->a, (size_t)dm->gsize));
  else
    for (PetscInt i = 0; i < dm->gsize; i++) l->a[dm->g2l[i]] = g->a[i];
  return 0;
}
/* ADD_VALUES: local contributions summed into the global vector, constrained dofs dropped (one rank: no ghosts) */
PetscErrorCode DMLocalToGlobal(DM dm, Vec l, InsertMode mode, Vec g) {
  if (mode != ADD_VALUES) return 56;
  if (l->device) {
    Vec tmp;
    PetscErrorCode ierr = vec_new(dm->gsize, 1, dm, &tmp);
    if (ierr) return ierr;
    B2CHK(b200_gather(tmp->a, l->a, dm->d_l2g_loc, (size_t)dm->gsize));
    B2CHK(b200_vec_axpy(g->a, 1.0, tmp->a, (size_t)dm->gsize));
    B2CHK(b200_sync());
    VecDestroy(&tmp);
  } else {
    for (PetscInt i = 0; i < dm->gsize; i++) g->a[i] += l->a[dm->g2l[i]];
  }
  return 0;
}
/* essential boundary values at `time` (= load increment) into the local vector */
PetscErrorCode DMPlexInsertBoundaryValues(DM dm, PetscBool insertEssential, Vec l, PetscReal time, Vec a, Vec b, Vec c) {
  (void)insertEssential; (void)a; (void)b; (void)c;
  if (!dm->nbc) return 0;
  if (l->device) {
    double *scaled = (double *)malloc(sizeof(double) * (size_t)dm->nbc), *d;
    for (PetscInt i = 0; i < dm->nbc; i++) scaled[i] = dm->bc_val[i] * time;
    B2CHK(b200_malloc((void **)&d, sizeof(double) * (size_t)dm->nbc));
    B2CHK(b200_memcpy_h2d(d, scaled, sizeof(double) * (size_t)dm->nbc));
    B2CHK(b200_scatter_set(l->a, dm->d_bc_idx, d, (size_t)dm->nbc));
    B2CHK(b200_sync());
    b200_free(d);
    free(scaled);
  } else {
    for (PetscInt i = 0; i < dm->nbc; i++) l->a[dm->bc_idx[i]] = dm->bc_val[i] * time;
  }
  return 0;
}

/* ---------------------------------------------------------------- Mat / misc */
PetscErrorCode MatShellGetContext(Mat A, void *ctx) { *(void **)ctx = A->ctx; return 0; }
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }
PetscErrorCode SNESComputeJacobianDefaultColor(SNES s, Vec u, Mat J, Mat P, void *ctx) {
  (void)s; (void)u; (void)J; (void)P; (void)ctx;
  return 0;
}
PetscErrorCode PetscViewerVTKOpen(MPI_Comm c, const char *name, PetscFileMode m, PetscViewer *v) {
  (void)c; (void)name; (void)m; *v = NULL;
  return 0;
}
PetscErrorCode PetscViewerDestroy(PetscViewer *v) { *v = NULL; return 0; }
PetscErrorCode PetscObjectSetName(PetscObject o, const char *name) { (void)o; (void)name; return 0; }
PetscErrorCode PetscSNPrintf(char *buf, size_t len, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, len, fmt, ap);
  va_end(ap);
  return 0;
}
int MPI_Allreduce(const void *in, void *out, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  (void)in; (void)out; (void)n; (void)t; (void)op; (void)c;   /* one rank, MPI_IN_PLACE: nothing to do */
  return 0;
}
