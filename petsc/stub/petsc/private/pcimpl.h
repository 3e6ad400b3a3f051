/* COMPILE-CHECK STUB, NOT PETSc: the fields of struct _p_PC (petsc/private/pcimpl.h, PETSc 3.22) the plug-in touches */
#ifndef WB_PETSC_STUB_PCIMPL_H
#define WB_PETSC_STUB_PCIMPL_H
#include <petscksp.h>
struct _PCOps {
  PetscErrorCode (*setup)(PC);
  PetscErrorCode (*apply)(PC, Vec, Vec);
  PetscErrorCode (*setfromoptions)(PC, PetscOptionItems *);
  PetscErrorCode (*destroy)(PC);
};
struct _p_PC {
  struct _PCOps ops[1];
  void *data;
  PCFailedReason failedreason;
};
#endif
