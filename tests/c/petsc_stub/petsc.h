/* TEST-ONLY stand-in for <petsc.h>: just enough TYPES and MACROS for the reference's libCEED-facing sources
 * (/root/reference/src/setuplibceed.c, src/matops.c, src/misc.c and elasticity.h) to pass `gcc -fsyntax-only`
 * against THIS repository's <ceed.h>.  PETSc is not available in the image; PETSc FUNCTIONS are
 * deliberately left undeclared (C implicit declarations) -- the test only insists that every Ceed* identifier the
 * reference uses is declared by include/ceed/ceed.h with a compatible prototype.
 * tests/test_reference_sources_compile.py */
#ifndef PETSC_STUB_H
#define PETSC_STUB_H
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>

typedef int PetscErrorCode;
typedef int PetscInt;
typedef int PetscMPIInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef double PetscLogDouble;
typedef int PetscClassId;
typedef int PetscLogStage;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef int MPI_Comm;
typedef int InsertMode;
typedef int PetscCopyMode;
typedef int PetscDataType;
typedef int DMLabelValue;
typedef int NormType;
typedef int VecType_;
typedef const char *VecType;
typedef const char *MatType;
typedef int MatOperation;
typedef int PetscMemType;
typedef int DMBoundaryConditionType;
typedef int PetscViewerFormat;
typedef struct _p_DM *DM;
typedef struct _p_Vec *Vec;
typedef struct _p_Mat *Mat;
typedef struct _p_SNES *SNES;
typedef struct _p_KSP *KSP;
typedef struct _p_PC *PC;
typedef struct _p_PetscSection *PetscSection;
typedef struct _p_PetscFE *PetscFE;
typedef struct _p_PetscDS *PetscDS;
typedef struct _p_DMLabel *DMLabel;
typedef struct _p_IS *IS;
typedef struct _p_PetscViewer *PetscViewer;
typedef struct _p_PetscSF *PetscSF;
typedef struct _p_PetscSpace *PetscSpace;
typedef struct _p_PetscDualSpace *PetscDualSpace;
typedef struct _p_PetscQuadrature *PetscQuadrature;
typedef struct _p_PetscPartitioner *PetscPartitioner;
typedef struct _p_ISColoring *ISColoring;
typedef struct _p_MatFDColoring *MatFDColoring;
typedef struct _p_MatColoring *MatColoring;

typedef struct _p_PetscObject *PetscObject;
/* PETSc functions the reference stores in function-pointer members (elasticity.h UserMult, misc.c:58-66) */
PetscErrorCode VecGetArray(Vec, PetscScalar **);
PetscErrorCode VecGetArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecRestoreArray(Vec, PetscScalar **);
PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecCUDAGetArray(Vec, PetscScalar **);
PetscErrorCode VecCUDAGetArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecCUDARestoreArray(Vec, PetscScalar **);
PetscErrorCode VecCUDARestoreArrayRead(Vec, const PetscScalar **);
/* elasticity.c (main driver): solver-configuration names */
typedef const char *KSPType;
typedef const char *PCType;
typedef const char *SNESType;
typedef int PCMGCycleType;
typedef int PCMGType;
typedef int SNESConvergedReason;
typedef struct _p_SNESLineSearch *SNESLineSearch;
extern const char *const PCMGCycleTypes[], *const PCMGTypes[], *const *SNESConvergedReasons;
PetscErrorCode SNESComputeJacobianDefaultColor(SNES, Vec, Mat, Mat, void *);
#define KSPCG "cg"
#define KSPCHEBYSHEV "chebyshev"
#define KSPPREONLY "preonly"
#define KSP_NORM_NATURAL 3
#define MATAIJ "aij"
#define MATOP_MULT_TRANSPOSE 5
#define MPI_DOUBLE 0
#define PCGAMG "gamg"
#define PCJACOBI "jacobi"
#define PCMG "mg"
#define PC_JACOBI_DIAGONAL 0
#define PC_MG_CYCLE_V 1
#define PC_MG_MULTIPLICATIVE 0
#define PETSC_ERR_SUP_SYS 57
#define SNESLINESEARCHCP "cp"
#define PETSC_VERSION_LT(a, b, c) 0
#define PETSC_VERSION_GE(a, b, c) 1
#define PETSC_ERR_ARG_INCOMP 75
#define MPI_IN_PLACE ((void *)1)
#define MPIU_SUM 0
#define MAT_FINAL_ASSEMBLY 0
#define FILE_MODE_WRITE 1
#define PETSC_COMM_WORLD 0
#define PETSC_COMM_SELF 0
#define PETSC_MAX_PATH_LEN 4096
#define PETSC_DEFAULT (-2)
#define PETSC_DECIDE (-1)
#define PETSC_DETERMINE (-1)
#define PETSC_ERR_ARG_WRONG 62
#define PETSC_ERR_SUP 56
#define PETSC_ERR_ARG_OUTOFRANGE 63
#define PETSC_ERR_LIB 76
#define PETSC_STATIC_INLINE static inline
#define PetscFunctionBeginUser do { } while (0)
#define PetscFunctionBegin do { } while (0)
#define PetscFunctionReturn(x) return (x)
#define CHKERRQ(ierr) do { if (ierr) return (ierr); } while (0)
#define SETERRQ(comm, code, msg) return (code)
#define SETERRQ1(comm, code, msg, a) return (code)
#define SETERRQ2(comm, code, msg, a, b) return (code)
#define SETERRQ3(comm, code, msg, a, b, c) return (code)
#define PetscMax(a, b) (((a) < (b)) ? (b) : (a))
#define PetscMin(a, b) (((a) < (b)) ? (a) : (b))
#define PetscSqr(a) ((a) * (a))
#define PetscPowInt(b, e) ((PetscInt)pow((double)(b), (double)(e)))
#define PETSC_PI M_PI
#define INSERT_VALUES 1
#define ADD_VALUES 2
#define INSERT_ALL_VALUES 3
#define NORM_2 1
#define NORM_1 0
#define NORM_MAX 3
#define PETSC_COPY_VALUES 0
#define PETSC_OWN_POINTER 1
#define PETSC_USE_POINTER 2
#define MATOP_MULT 3
#define MATOP_GET_DIAGONAL 17
#define MATSHELL "shell"
#define VECCUDA "cuda"
#define VECSTANDARD "standard"
#define DM_BC_ESSENTIAL 1
#define PETSC_VIEWER_ASCII_INFO_DETAIL 2
#define PETSC_SCALAR 1
#define MPIU_INT 0
#define MPIU_REAL 0
#define MPIU_SCALAR 0
#define MPI_SUM 0
#define MPI_MAX 0
#define MPI_MIN 0
#endif
