/* TEST-ONLY: the reference's OWN host code driving /gpu/b200.
 *
 * This program links the reference's  src/setuplibceed.c, src/matops.c and src/misc.c  -- compiled UNCHANGED from
 * /root/reference where they lie (oracle/Makefile target `refdriver`; nothing is copied into this repository) --
 * against this repository's <ceed.h> / libceed_b200.so and a functional miniature of the PETSc pieces they use
 * (tests/c/petsc_mini).  It follows the set-up order of the reference's main() (/root/reference/elasticity.c:225-452):
 *
 *   SetupLibceedFineLevel, SetupLibceedLevel (every level)           setuplibceed.c:243-745, 748-939
 *   SetupJacobianCtx, SetupProlongRestrictCtx                         misc.c:26-146
 *   FormResidual_Ceed, ApplyJacobian_Ceed, GetDiag_Ceed,
 *   Prolong_Ceed, Restrict_Ceed, ComputeStrainEnergy                  matops.c:63-300
 *   ViewDiagnosticQuantities (its VecView captured instead of written) misc.c:217-300
 *   optionally (forcing type in the second byte of the problem word): the forcing operator and, for MMS, the nodal
 *   true solution that SetupLibceedFineLevel computes                  setuplibceed.c:550-640
 *
 * so the libCEED objects are created, wired and applied by the reference's code, the QFunction pointers are the
 * reference's own (which also exercises the backend's QFunction guard), and only the mesh (box, lexicographic
 * numbering, closure indices with essential-BC dofs encoded as -(loc+1)) and the vectors come from the input file
 * written by tests/test_reference_host_code_on_gpu.py, which compares the outputs with the CPU oracle.
 *
 *   ref_driver <input.bin> <output.bin> [resource]
 */
#include "elasticity.h" /* the reference's header: -I/root/reference */

#include "b200_kernels.h"

extern double *petsc_mini_viewed;   /* last vector passed to VecView (petsc_mini.c) */
extern PetscInt petsc_mini_viewed_n;
static FILE *fin;
static int rd_i(void) { int v; if (fread(&v, sizeof v, 1, fin) != 1) { fprintf(stderr, "ref_driver: short input\n"); exit(2); } return v; }
static int *rd_iv(size_t n) {
  int *p = (int *)malloc(sizeof(int) * (n ? n : 1));
  if (fread(p, sizeof(int), n, fin) != n) { fprintf(stderr, "ref_driver: short input\n"); exit(2); }
  return p;
}
static double *rd_dv(size_t n) {
  double *p = (double *)malloc(sizeof(double) * (n ? n : 1));
  if (fread(p, sizeof(double), n, fin) != n) { fprintf(stderr, "ref_driver: short input\n"); exit(2); }
  return p;
}
static int *to_device_i(const int *h, size_t n) {
  int *d = NULL;
  if (b200_malloc((void **)&d, sizeof(int) * (n ? n : 1)) || b200_memcpy_h2d(d, h, sizeof(int) * n) || b200_sync()) {
    fprintf(stderr, "ref_driver: %s\n", b200_last_error());
    exit(3);
  }
  return d;
}

/* a DM without constraints: global = local */
static void identity_maps(DM dm) {
  dm->l2g = (int *)malloc(sizeof(int) * (size_t)dm->lsize);
  for (int i = 0; i < dm->lsize; i++) dm->l2g[i] = i;
  dm->g2l = dm->l2g;
  if (dm->device) dm->d_l2g_loc = to_device_i(dm->g2l, (size_t)dm->gsize);
}

/* solution-space DM of one level from the input stream */
static DM read_dm(int nelem, int P, int ncomp, int device, int with_maps) {
  DM dm = (DM)calloc(1, sizeof *dm);
  dm->dim = 3; dm->nelem = nelem; dm->P = P; dm->ncomp = ncomp; dm->device = device;
  dm->lsize = rd_i();
  dm->gsize = rd_i();
  dm->closure = rd_iv((size_t)nelem * P * P * P * ncomp);
  if (with_maps) {
    dm->l2g = rd_iv((size_t)dm->lsize);
    dm->g2l = (int *)malloc(sizeof(int) * (size_t)(dm->gsize ? dm->gsize : 1));
    for (int i = 0; i < dm->lsize; i++)
      if (dm->l2g[i] >= 0) dm->g2l[dm->l2g[i]] = i;
    if (device) dm->d_l2g_loc = to_device_i(dm->g2l, (size_t)dm->gsize);
  }
  return dm;
}

static void set_global(Vec v, const double *h) {
  if (v->device) { b200_memcpy_h2d(v->a, h, sizeof(double) * (size_t)v->n); b200_sync(); }
  else memcpy(v->a, h, sizeof(double) * (size_t)v->n);
}
static void write_vec(FILE *f, Vec v) {
  double *h = (double *)malloc(sizeof(double) * (size_t)(v->n ? v->n : 1));
  if (v->device) b200_memcpy_d2h(h, v->a, sizeof(double) * (size_t)v->n);
  else memcpy(h, v->a, sizeof(double) * (size_t)v->n);
  fwrite(h, sizeof(double), (size_t)v->n, f);
  free(h);
}
#define CHK(call) do { PetscErrorCode e_ = (call); if (e_) { fprintf(stderr, "ref_driver: %s failed (%d)\n", #call, (int)e_); return 4; } } while (0)

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: ref_driver input.bin output.bin [resource]\n"); return 1; }
  const char *resource = argc > 3 ? argv[3] : "/gpu/b200";
  fin = fopen(argv[1], "rb");
  if (!fin) { perror(argv[1]); return 1; }
  if (rd_i() != 0x42323030) { fprintf(stderr, "ref_driver: bad magic\n"); return 2; }
  const int numLevels = rd_i(), nelem = rd_i(), problemWord = rd_i(), memtype = rd_i();
  /* low byte: problemType; next byte: forcingType (0 = none, the default of the GPU test; elasticity.h:64-66) */
  const int problem = problemWord & 0xff, forcing = (problemWord >> 8) & 0xff;
  const double nu = 0.3, E = 1.0;
  int *degrees = rd_iv((size_t)numLevels);
  const int fineLevel = numLevels - 1, device = memtype == CEED_MEM_DEVICE, ncompu = 3;

  /* ---- application context (elasticity.h:117-141), as ProcessCommandLineOptions would fill it */
  AppCtx appCtx = (AppCtx)calloc(1, sizeof *appCtx);
  snprintf(appCtx->ceedResource, sizeof appCtx->ceedResource, "%s", resource);
  appCtx->problemChoice = (problemType)problem;
  appCtx->forcingChoice = (forcingType)forcing;
  appCtx->forcingVector[0] = 0.3; appCtx->forcingVector[1] = -1.0; appCtx->forcingVector[2] = 2.5;
  appCtx->multigridChoice = MULTIGRID_LOGARITHMIC;
  appCtx->degree = degrees[fineLevel];
  appCtx->qextra = 0;
  appCtx->numLevels = numLevels;
  appCtx->levelDegrees = degrees;
  appCtx->numIncrements = 1;
  appCtx->memTypeRequested = (CeedMemType)memtype;
  Physics phys = (Physics)calloc(1, sizeof *phys);
  phys->nu = nu; phys->E = E;

  Ceed ceed;
  CeedInit(appCtx->ceedResource, &ceed);

  /* ---- DMs: coordinates, one per level, energy, diagnostic */
  DM dmcoord = read_dm(nelem, 2, 3, 0, 0);
  double *coords = rd_dv((size_t)dmcoord->lsize);
  Vec coordVec = (Vec)calloc(1, sizeof *coordVec);
  coordVec->n = dmcoord->lsize; coordVec->a = coords; coordVec->dm = dmcoord;
  DM *levelDMs = (DM *)calloc((size_t)numLevels, sizeof(DM));
  for (int l = 0; l < numLevels; l++) {
    levelDMs[l] = read_dm(nelem, degrees[l] + 1, ncompu, device, 1);
    levelDMs[l]->coordDM = dmcoord;
    levelDMs[l]->coords = coordVec;
  }
  DM dmEnergy = read_dm(nelem, degrees[fineLevel] + 1, 1, device, 0);
  DM dmDiagnostic = read_dm(nelem, degrees[fineLevel] + 1, ncompu + 5, device, 0);
  DM fine = levelDMs[fineLevel];
  fine->nbc = rd_i();
  fine->bc_idx = rd_iv((size_t)fine->nbc);
  fine->bc_val = rd_dv((size_t)fine->nbc);
  if (device) fine->d_bc_idx = to_device_i(fine->bc_idx, (size_t)fine->nbc);

  /* ---- libCEED objects, built by the reference (elasticity.c:243-279) */
  CeedQFunction qfRestrict = NULL, qfProlong = NULL;
  CeedQFunctionCreateIdentity(ceed, ncompu, CEED_EVAL_NONE, CEED_EVAL_INTERP, &qfRestrict);
  CeedQFunctionCreateIdentity(ceed, ncompu, CEED_EVAL_INTERP, CEED_EVAL_NONE, &qfProlong);
  CeedData *ceedData = (CeedData *)calloc((size_t)numLevels, sizeof(CeedData));
  for (int l = 0; l < numLevels; l++) ceedData[l] = (CeedData)calloc(1, sizeof **ceedData);
  CeedVector forceCeed = NULL;       /* elasticity.c:289-294: the forcing L-vector, filled by the set-up */
  if (forcing != FORCE_NONE) CeedVectorCreate(ceed, fine->lsize, &forceCeed);
  CHK(SetupLibceedFineLevel(fine, dmEnergy, dmDiagnostic, ceed, appCtx, phys, ceedData, fineLevel, ncompu, fine->gsize,
                            fine->lsize, forceCeed, qfRestrict, qfProlong));
  for (int l = 0; l < numLevels; l++)
    CHK(SetupLibceedLevel(levelDMs[l], ceed, appCtx, phys, ceedData, l, ncompu, levelDMs[l]->gsize, levelDMs[l]->lsize, NULL,
                          qfRestrict, qfProlong));

  /* ---- MatShell contexts (elasticity.c:386-452) */
  Vec *Ug = (Vec *)calloc((size_t)numLevels, sizeof(Vec)), *Uloc = (Vec *)calloc((size_t)numLevels, sizeof(Vec));
  UserMult *jacobCtx = (UserMult *)calloc((size_t)numLevels, sizeof(UserMult));
  struct _p_Mat *jacobMat = (struct _p_Mat *)calloc((size_t)numLevels, sizeof *jacobMat);
  for (int l = 0; l < numLevels; l++) {
    CHK(DMCreateGlobalVector(levelDMs[l], &Ug[l]));
    CHK(DMCreateLocalVector(levelDMs[l], &Uloc[l]));
    jacobCtx[l] = (UserMult)calloc(1, sizeof **jacobCtx);
    CHK(SetupJacobianCtx(0, appCtx, levelDMs[l], Ug[l], Uloc[l], ceedData[l], ceed, phys, NULL, jacobCtx[l]));
    jacobMat[l].ctx = jacobCtx[l];
  }
  UserMult resCtx = (UserMult)calloc(1, sizeof *resCtx);
  memcpy(resCtx, jacobCtx[fineLevel], sizeof *resCtx);
  resCtx->op = ceedData[fineLevel]->opApply;
  resCtx->qf = ceedData[fineLevel]->qfApply;
  resCtx->loadIncrement = 1.0;
  UserMultProlongRestr *prCtx = (UserMultProlongRestr *)calloc((size_t)numLevels, sizeof(UserMultProlongRestr));
  struct _p_Mat *prMat = (struct _p_Mat *)calloc((size_t)numLevels, sizeof *prMat);
  for (int l = 1; l < numLevels; l++) {
    prCtx[l] = (UserMultProlongRestr)calloc(1, sizeof **prCtx);
    CHK(SetupProlongRestrictCtx(0, appCtx, levelDMs[l - 1], levelDMs[l], Ug[l], Uloc[l - 1], Uloc[l], ceedData[l - 1],
                                ceedData[l], ceed, prCtx[l]));
    prMat[l].ctx = prCtx[l];
  }

  /* ---- run the reference's MatShell callbacks */
  FILE *fout = fopen(argv[2], "wb");
  if (!fout) { perror(argv[2]); return 1; }
  double *hU = rd_dv((size_t)fine->gsize);
  Vec U, R;
  CHK(DMCreateGlobalVector(fine, &U));
  CHK(DMCreateGlobalVector(fine, &R));
  set_global(U, hU);
  CHK(FormResidual_Ceed(NULL, U, R, resCtx));           /* also writes gradu for the Jacobians below */
  write_vec(fout, R);
  Vec *X = (Vec *)calloc((size_t)numLevels, sizeof(Vec));
  for (int l = 0; l < numLevels; l++) {
    Vec Y, D;
    double *hX = rd_dv((size_t)levelDMs[l]->gsize);
    CHK(DMCreateGlobalVector(levelDMs[l], &X[l]));
    CHK(DMCreateGlobalVector(levelDMs[l], &Y));
    CHK(DMCreateGlobalVector(levelDMs[l], &D));
    set_global(X[l], hX);
    CHK(ApplyJacobian_Ceed(&jacobMat[l], X[l], Y));
    write_vec(fout, Y);
    CHK(GetDiag_Ceed(&jacobMat[l], D));
    write_vec(fout, D);
    free(hX);
    VecDestroy(&Y); VecDestroy(&D);
  }
  for (int l = 1; l < numLevels; l++) {
    Vec Yf, Yc;
    CHK(DMCreateGlobalVector(levelDMs[l], &Yf));
    CHK(DMCreateGlobalVector(levelDMs[l - 1], &Yc));
    CHK(Prolong_Ceed(&prMat[l], X[l - 1], Yf));
    write_vec(fout, Yf);
    CHK(Restrict_Ceed(&prMat[l], X[l], Yc));
    write_vec(fout, Yc);
    VecDestroy(&Yf); VecDestroy(&Yc);
  }
  PetscReal energy = 0;
  CHK(ComputeStrainEnergy(dmEnergy, resCtx, ceedData[fineLevel]->opEnergy, U, &energy));
  fwrite(&energy, sizeof energy, 1, fout);
  { /* nodal diagnostic quantities, as elasticity.c:835-850 sets them up */
    identity_maps(dmDiagnostic);
    UserMult diagnosticCtx = (UserMult)calloc(1, sizeof *diagnosticCtx);
    memcpy(diagnosticCtx, resCtx, sizeof *resCtx);
    diagnosticCtx->dm = dmDiagnostic;
    diagnosticCtx->op = ceedData[fineLevel]->opDiagnostic;
    CHK(ViewDiagnosticQuantities(0, fine, diagnosticCtx, U, ceedData[fineLevel]->ErestrictDiagnostic));
    if (petsc_mini_viewed_n != dmDiagnostic->gsize) { fprintf(stderr, "ref_driver: diagnostic vector was not viewed\n"); return 5; }
    fwrite(petsc_mini_viewed, sizeof(double), (size_t)petsc_mini_viewed_n, fout);
    free(diagnosticCtx);
  }
  if (forcing != FORCE_NONE) { /* appended after everything else: the forcing L-vector and, for MMS, the nodal true solution */
    const CeedScalar *f;
    CeedVectorGetArrayRead(forceCeed, CEED_MEM_HOST, &f);
    fwrite(f, sizeof(double), (size_t)fine->lsize, fout);
    CeedVectorRestoreArrayRead(forceCeed, &f);
    if (forcing == FORCE_MMS) {
      CeedVectorGetArrayRead(ceedData[fineLevel]->truesoln, CEED_MEM_HOST, &f);
      fwrite(f, sizeof(double), (size_t)fine->lsize, fout);
      CeedVectorRestoreArrayRead(ceedData[fineLevel]->truesoln, &f);
    }
    CeedVectorDestroy(&forceCeed);
  }
  fclose(fout);
  fclose(fin);
  int det = 0;
  CeedIsDeterministic(ceed, &det);
  printf("ref_driver OK: %d levels, %d elements, problem %d, memtype %s, resource %s%s, strain energy %.15e\n", numLevels, nelem,
         problem, device ? "device" : "host", resource, det ? " (deterministic)" : "", energy);
  for (int l = 0; l < numLevels; l++) CeedDataDestroy(l, ceedData[l]);
  CeedQFunctionDestroy(&qfRestrict);
  CeedQFunctionDestroy(&qfProlong);
  CeedDestroy(&ceed);
  return 0;
}
