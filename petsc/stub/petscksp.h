/*
 * petsc/stub/petscksp.h -- COMPILE-CHECK STUB, NOT PETSc.
 *
 * Declares exactly the PETSc 3.22 / MPI names petsc/wb_petsc_plugin.c uses, with the argument lists of the PETSc 3.22
 * manual pages, so that the plug-in can be syntax- and type-checked in an image that has neither PETSc nor MPI
 * (tests/test_petsc_plugin.py).  Nothing here is implemented; a real build uses the real headers (see the plug-in).
 */
#ifndef WB_PETSC_STUB_H
#define WB_PETSC_STUB_H
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef int PetscErrorCode;
typedef int PetscInt;
typedef int PetscMPIInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
#define PETSC_SUCCESS 0
#define PETSC_ERR_SUP 56
#define PETSC_ERR_LIB 76
#define PETSC_EXTERN extern
typedef int MPI_Comm;
typedef int MPI_Datatype;
#define PETSC_COMM_SELF 1
#define MPI_INT 2
#define MPI_BYTE 3
#define MPIU_INT MPI_INT
int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Comm_size(MPI_Comm, int *);
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Alltoall(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm);
int MPI_Allgather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm);
int MPI_Alltoallv(const void *, const int *, const int *, MPI_Datatype, void *, const int *, const int *, MPI_Datatype, MPI_Comm);

typedef struct _p_PetscObject *PetscObject;
typedef struct _p_Mat *Mat;
typedef struct _p_Vec *Vec;
typedef struct _p_PC *PC;
typedef struct _p_KSP *KSP;
typedef struct _n_PetscOptions *PetscOptions;
typedef struct _p_PetscOptionItems PetscOptionItems;
typedef const char *MatType;
#define MATMPIBAIJ "mpibaij"
typedef enum { PC_LEFT, PC_RIGHT, PC_SYMMETRIC } PCSide;
typedef enum { KSP_NORM_NONE, KSP_NORM_PRECONDITIONED, KSP_NORM_UNPRECONDITIONED, KSP_NORM_NATURAL } KSPNormType;
typedef enum { KSP_CONVERGED_ITERATING = 0, KSP_CONVERGED_RTOL = 2, KSP_CONVERGED_ATOL = 3, KSP_DIVERGED_ITS = -3,
               KSP_DIVERGED_DTOL = -4, KSP_DIVERGED_BREAKDOWN = -5, KSP_DIVERGED_NANORINF = -9 } KSPConvergedReason;
typedef enum { PC_NOERROR, PC_FACTOR_STRUCT_ZEROPIVOT, PC_FACTOR_NUMERIC_ZEROPIVOT } PCFailedReason;

#define PetscFunctionBegin
#define PetscFunctionReturn(x) return (x)
#define PetscCall(call)               \
  do {                                \
    PetscErrorCode ierr_ = (call);    \
    if (ierr_) return ierr_;          \
  } while (0)
#define PetscCallMPI(call) PetscCall(call)
PetscErrorCode PetscStubError(MPI_Comm, PetscErrorCode, const char *, ...);
#define PetscCheck(cond, comm, code, ...)                         \
  do {                                                            \
    if (!(cond)) return PetscStubError(comm, code, __VA_ARGS__);  \
  } while (0)
MPI_Comm PetscObjectComm(PetscObject);
PetscErrorCode PetscObjectGetComm(PetscObject, MPI_Comm *);
PetscErrorCode PetscObjectTypeCompare(PetscObject, const char[], PetscBool *);
PetscErrorCode PetscMemzero(void *, size_t);
PetscErrorCode PetscStubMalloc(size_t, void *);
PetscErrorCode PetscStubFree(void *);
#define PetscNew(p) PetscStubMalloc(sizeof(**(p)), (void *)(p))
#define PetscMalloc1(n, p) PetscStubMalloc((size_t)(n) * sizeof(**(p)), (void *)(p))
#define PetscMalloc2(n1, p1, n2, p2) (PetscMalloc1(n1, p1) || PetscMalloc1(n2, p2))
#define PetscMalloc3(n1, p1, n2, p2, n3, p3) (PetscMalloc2(n1, p1, n2, p2) || PetscMalloc1(n3, p3))
#define PetscMalloc5(n1, p1, n2, p2, n3, p3, n4, p4, n5, p5) (PetscMalloc3(n1, p1, n2, p2, n3, p3) || PetscMalloc2(n4, p4, n5, p5))
#define PetscCalloc4(n1, p1, n2, p2, n3, p3, n4, p4) (PetscMalloc2(n1, p1, n2, p2) || PetscMalloc2(n3, p3, n4, p4))
#define PetscFree(p) PetscStubFree((void *)(p))
#define PetscFree2(a, b) (PetscFree(a) || PetscFree(b))
#define PetscFree3(a, b, c) (PetscFree2(a, b) || PetscFree(c))
#define PetscFree4(a, b, c, d) (PetscFree2(a, b) || PetscFree2(c, d))
#define PetscFree5(a, b, c, d, e) (PetscFree3(a, b, c) || PetscFree2(d, e))
#define PetscArraycpy(dst, src, n) ((void)memcpy((dst), (src), (size_t)(n) * sizeof(*(dst))), PETSC_SUCCESS)

PetscErrorCode PetscOptionsGetInt(PetscOptions, const char[], const char[], PetscInt *, PetscBool *);
void PetscStubOptionsHead(PetscOptionItems *, const char *);
#define PetscOptionsHeadBegin(items, title) PetscStubOptionsHead(items, title)
#define PetscOptionsHeadEnd() (void)0
PetscErrorCode PetscStubOptionsInt(PetscOptionItems *, const char[], const char[], const char[], PetscInt, PetscInt *, PetscBool *);
PetscErrorCode PetscStubOptionsBool(PetscOptionItems *, const char[], const char[], const char[], PetscBool, PetscBool *, PetscBool *);
#define PetscOptionsInt(opt, text, man, cur, val, set) PetscStubOptionsInt(items, opt, text, man, cur, val, set)
#define PetscOptionsBool(opt, text, man, cur, val, set) PetscStubOptionsBool(items, opt, text, man, cur, val, set)

PetscErrorCode MatGetBlockSize(Mat, PetscInt *);
PetscErrorCode MatGetSize(Mat, PetscInt *, PetscInt *);
PetscErrorCode MatGetOwnershipRanges(Mat, const PetscInt **);
PetscErrorCode MatGetRowIJ(Mat, PetscInt, PetscBool, PetscBool, PetscInt *, const PetscInt *[], const PetscInt *[], PetscBool *);
PetscErrorCode MatRestoreRowIJ(Mat, PetscInt, PetscBool, PetscBool, PetscInt *, const PetscInt *[], const PetscInt *[], PetscBool *);
PetscErrorCode MatSeqBAIJGetArray(Mat, PetscScalar *[]);
PetscErrorCode MatSeqBAIJRestoreArray(Mat, PetscScalar *[]);
PetscErrorCode MatMPIBAIJGetSeqBAIJ(Mat, Mat *, Mat *, const PetscInt *[]);
PetscErrorCode VecGetArray(Vec, PetscScalar **);
PetscErrorCode VecRestoreArray(Vec, PetscScalar **);
PetscErrorCode VecGetArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar **);
PetscErrorCode PCGetOperators(PC, Mat *, Mat *);
PetscErrorCode PCRegister(const char[], PetscErrorCode (*)(PC));
PetscErrorCode KSPRegister(const char[], PetscErrorCode (*)(KSP));
PetscErrorCode KSPGetPC(KSP, PC *);
PetscErrorCode KSPSetSupportedNorm(KSP, KSPNormType, PCSide, PetscInt);
#endif
