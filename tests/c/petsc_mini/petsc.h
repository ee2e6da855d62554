/* TEST-ONLY functional miniature of the PETSc pieces that the reference's libCEED-facing sources use
 * (/root/reference/src/setuplibceed.c, src/matops.c, src/misc.c): enough to COMPILE THEM UNCHANGED and RUN them
 * against libceed_b200.so (tests/c/ref_driver.c, tests/test_reference_host_code_on_gpu.py).
 *
 * PETSc is not in the image.  This is not PETSc: a Vec is an array (host memory, or device memory
 * for the -memtype device path, moved with the thin CUDA layer of the backend), a DM is a box-mesh description with
 * pre-computed closure indices, a MatShell holds its context.  Implementations: petsc_mini.c.
 * (tests/c/petsc_stub is the types-only sibling used for the syntax-only compile test.) */
#ifndef PETSC_MINI_H
#define PETSC_MINI_H
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int PetscErrorCode;
typedef int PetscInt;
typedef int PetscMPIInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef double PetscLogDouble;
typedef int PetscLogStage;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int InsertMode;
typedef const char *VecType;
typedef const char *MatType;
typedef int MatAssemblyType;
typedef int PetscFileMode;

typedef struct _p_DM *DM;
typedef struct _p_Vec *Vec;
typedef struct _p_Mat *Mat;
typedef struct _p_SNES *SNES;
typedef struct _p_KSP *KSP;
typedef struct _p_PC *PC;
typedef struct _p_PetscSection *PetscSection;
typedef struct _p_PetscFE *PetscFE;
typedef struct _p_PetscViewer *PetscViewer;
typedef struct _p_PetscObject *PetscObject;

struct _p_Vec {
  PetscInt n;
  int device;     /* 0: a is host memory, 1: a is device memory (VECCUDA stand-in) */
  double *a;
  DM dm;
};
struct _p_DM {
  PetscInt dim, nelem, P, ncomp;
  PetscInt lsize, gsize;
  PetscInt *closure;   /* [nelem][P^3 * ncomp] interlaced local dof indices in tensor order; essential-BC dofs as -(loc+1) */
  PetscInt *l2g;       /* [lsize] global (unconstrained, owned) index of a local dof, or -1 */
  PetscInt *d_l2g_loc; /* device: local index of every global dof [gsize] (for the device Vec path) */
  PetscInt *g2l;       /* host: the same map */
  DM coordDM;
  Vec coords;
  PetscInt nbc, *bc_idx, *d_bc_idx;   /* essential boundary dofs and their values at load increment 1 */
  double *bc_val, *d_bc_val;
  int device;          /* Vecs created from this DM live in device memory */
};
struct _p_Mat { void *ctx; };

#define PETSC_COMM_WORLD 0
#define PETSC_COMM_SELF 0
#define PETSC_MAX_PATH_LEN 4096
#define PETSC_DEFAULT (-2)
#define PETSC_DECIDE (-1)
#define PETSC_DETERMINE (-1)
#define PETSC_ERR_ARG_WRONG 62
#define PETSC_ERR_SUP 56
#define PETSC_ERR_ARG_INCOMP 75
#define PETSC_STATIC_INLINE static inline
#define PetscFunctionBeginUser do { } while (0)
#define PetscFunctionBegin do { } while (0)
#define PetscFunctionReturn(x) return (x)
#define CHKERRQ(ierr) do { if (ierr) { fprintf(stderr, "petsc_mini: error %d at %s:%d\n", (int)(ierr), __FILE__, __LINE__); return (ierr); } } while (0)
#define SETERRQ(comm, code, msg) do { fprintf(stderr, "petsc_mini: %s\n", msg); return (code); } while (0)
#define SETERRQ1(comm, code, msg, a) do { fprintf(stderr, "petsc_mini: %s\n", msg); return (code); } while (0)
#define SETERRQ2(comm, code, msg, a, b) do { fprintf(stderr, "petsc_mini: %s\n", msg); return (code); } while (0)
#define PetscMax(a, b) (((a) < (b)) ? (b) : (a))
#define PetscMin(a, b) (((a) < (b)) ? (a) : (b))
#define PetscMalloc1(n, p) ((*(p) = malloc(sizeof(**(p)) * (size_t)((n) > 0 ? (n) : 1))) ? 0 : 55)
#define PetscCalloc1(n, p) ((*(p) = calloc((size_t)((n) > 0 ? (n) : 1), sizeof(**(p)))) ? 0 : 55)
#define PetscFree(p) (free(p), (p) = NULL, 0)
#define PetscMemcpy(d, s, n) (memcpy((d), (s), (n)), 0)
#define PETSC_VERSION_LT(a, b, c) 0
#define PETSC_VERSION_GE(a, b, c) 1
#define INSERT_VALUES 1
#define ADD_VALUES 2
#define MAT_FINAL_ASSEMBLY 0
#define FILE_MODE_WRITE 1
#define MPI_IN_PLACE ((void *)1)
#define MPIU_REAL 0
#define MPIU_SUM 0
#define VECCUDA "cuda"
#define VECSTANDARD "standard"

/* Vec */
PetscErrorCode VecZeroEntries(Vec);
PetscErrorCode VecDuplicate(Vec, Vec *);
PetscErrorCode VecDestroy(Vec *);
PetscErrorCode VecGetSize(Vec, PetscInt *);
PetscErrorCode VecGetLocalSize(Vec, PetscInt *);
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y);
PetscErrorCode VecReciprocal(Vec);
PetscErrorCode VecGetArray(Vec, PetscScalar **);
PetscErrorCode VecGetArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecRestoreArray(Vec, PetscScalar **);
PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecCUDAGetArray(Vec, PetscScalar **);
PetscErrorCode VecCUDAGetArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecCUDARestoreArray(Vec, PetscScalar **);
PetscErrorCode VecCUDARestoreArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecGetDM(Vec, DM *);
PetscErrorCode VecView(Vec, PetscViewer);
/* DM */
PetscErrorCode DMGetDimension(DM, PetscInt *);
PetscErrorCode DMGetSection(DM, PetscSection *);
PetscErrorCode DMGetCoordinateDM(DM, DM *);
PetscErrorCode DMGetCoordinatesLocal(DM, Vec *);
PetscErrorCode DMGetLocalVector(DM, Vec *);
PetscErrorCode DMRestoreLocalVector(DM, Vec *);
PetscErrorCode DMCreateLocalVector(DM, Vec *);
PetscErrorCode DMCreateGlobalVector(DM, Vec *);
PetscErrorCode DMGlobalToLocal(DM, Vec, InsertMode, Vec);
PetscErrorCode DMLocalToGlobal(DM, Vec, InsertMode, Vec);
PetscErrorCode DMSetOutputSequenceNumber(DM, PetscInt, PetscReal);
PetscErrorCode DMPlexGetHeightStratum(DM, PetscInt, PetscInt *, PetscInt *);
PetscErrorCode DMPlexGetClosureIndices(DM, PetscSection, PetscSection, PetscInt, PetscBool, PetscInt *, PetscInt **, PetscInt *,
                                       PetscScalar **);
PetscErrorCode DMPlexRestoreClosureIndices(DM, PetscSection, PetscSection, PetscInt, PetscBool, PetscInt *, PetscInt **,
                                           PetscInt *, PetscScalar **);
PetscErrorCode DMPlexSetClosurePermutationTensor(DM, PetscInt, PetscSection);
PetscErrorCode DMPlexInsertBoundaryValues(DM, PetscBool, Vec, PetscReal, Vec, Vec, Vec);
/* Mat / SNES / misc */
PetscErrorCode MatShellGetContext(Mat, void *);
PetscErrorCode MatAssemblyBegin(Mat, MatAssemblyType);
PetscErrorCode MatAssemblyEnd(Mat, MatAssemblyType);
PetscErrorCode SNESComputeJacobianDefaultColor(SNES, Vec, Mat, Mat, void *);
PetscErrorCode PetscViewerVTKOpen(MPI_Comm, const char *, PetscFileMode, PetscViewer *);
PetscErrorCode PetscViewerDestroy(PetscViewer *);
PetscErrorCode PetscObjectSetName(PetscObject, const char *);
PetscErrorCode PetscSNPrintf(char *, size_t, const char *, ...);
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
#endif
