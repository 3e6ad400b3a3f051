/* COMPILE-CHECK STUB, NOT PETSc: the fields of struct _p_KSP (petsc/private/kspimpl.h, PETSc 3.22) the plug-in touches */
#ifndef WB_PETSC_STUB_KSPIMPL_H
#define WB_PETSC_STUB_KSPIMPL_H
#include <petscksp.h>
struct _KSPOps {
  PetscErrorCode (*setup)(KSP);
  PetscErrorCode (*solve)(KSP);
  PetscErrorCode (*setfromoptions)(KSP, PetscOptionItems *);
  PetscErrorCode (*destroy)(KSP);
};
struct _p_KSP {
  struct _KSPOps ops[1];
  void *data;
  PetscInt max_it, its;
  PetscReal rtol, abstol, divtol, rnorm;
  KSPConvergedReason reason;
  Vec vec_rhs, vec_sol;
};
#endif
