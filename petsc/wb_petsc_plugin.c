/*
 * wb_petsc_plugin.c -- seam B1 of INTEGRATION.md: PETSc run-time types "wbpc" (PC) and "wbksp" (KSP) that hand the
 * linear solve of Waiwera's Newton iteration to the B200 engine through the C ABI of include/waiwera_b200.h.
 *
 * Waiwera configures its KSP / PC from the JSON input and then calls KSPSetFromOptions / PCSetFromOptions
 * (src/timestepper.F90:1708, 1783), and it forwards command-line PETSc options (src/waiwera.F90:70-86).  So
 *
 *     mpiexec -np <gpus> waiwera model.json -dll_append libwb_petsc_plugin.so -ksp_type wbksp -pc_type wbpc
 *
 * replaces KSPSolve (GMRES / BiCGStab), PCSetUp (block-Jacobi ILU(0) factorisation) and PCApply for the flow
 * Jacobian with zero change to the Fortran sources: one MPI rank per GPU, the rank's rows of the (MPI)BAIJ matrix
 * Waiwera assembled (src/ode.F90:266-287) are read with the public Mat accessors every time PCSetUp runs (once per
 * Newton iteration), the solve runs on the device, KSPConvergedReason / iteration count / residual norm go back
 * where src/timestepper.F90:1872, 2397-2406 read them.
 *
 * Build against PETSc >= 3.22 (the version Waiwera pins, config.py:21):
 *     mpicc -shared -fPIC -I$PETSC_DIR/include -I$PETSC_DIR/$PETSC_ARCH/include -Iinclude \
 *           petsc/wb_petsc_plugin.c -Lwaiwera_b200 -lwaiwera_b200 -lpetsc -o libwb_petsc_plugin.so
 * This image has no PETSc and no MPI: the file is compile-checked against petsc/stub/ (declarations of exactly the
 * PETSc / MPI names used here, written from the PETSc 3.22 manual pages -- NOT PETSc) by tests/test_petsc_plugin.py.
 */
#include <petscksp.h>
#include <petsc/private/kspimpl.h>
#include <petsc/private/pcimpl.h>

#include "waiwera_b200.h"

typedef struct {
  wb_ctx *ctx;        /* one engine context per rank = per GPU */
  wb_mat *A;          /* the rank's rows of the Jacobian, ghost columns >= nb */
  wb_pc *pc;
  PetscInt nb, ncolb, bs, nnzb;
  PetscInt local_blocks; /* -pc_wb_local_blocks: block-Jacobi sub-domains per GPU (PETSc -pc_bjacobi_local_blocks) */
  PetscInt cube;         /* -pc_wb_subdomain_rows: rows per sub-domain (contiguous ranges), 0 = use local_blocks */
  PetscInt overlap;      /* -pc_wb_overlap: 0 block Jacobi, 1 restricted additive Schwarz (PCASM default) on the same sub-domains */
  int32_t *rowptr, *colidx;
  double *vals;
} WbPC;

static PetscErrorCode WbCheck(int rc, const char *what) {
  PetscFunctionBegin;
  PetscCheck(rc >= 0, PETSC_COMM_SELF, PETSC_ERR_LIB, "%s: %s", what, wb_last_error());
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* one engine context per process: the parameters of the flow path are irrelevant for the linear-algebra seam */
static PetscErrorCode WbContext(MPI_Comm comm, wb_ctx **out) {
  static wb_ctx *ctx = NULL;
  PetscMPIInt rank, size;
  PetscFunctionBegin;
  if (!ctx) {
    wb_params prm;
    int ndev = 1, device;
    unsigned char id[128];
    PetscCall(PetscMemzero(&prm, sizeof(prm)));
    prm.eos = WB_EOS_WE;
    PetscCallMPI(MPI_Comm_rank(comm, &rank));
    PetscCallMPI(MPI_Comm_size(comm, &size));
    PetscCall(PetscOptionsGetInt(NULL, NULL, "-wb_devices_per_node", &ndev, NULL));
    device = rank % (ndev > 0 ? ndev : 1);
    PetscCall(WbCheck(wb_create(&prm, device, &ctx), "wb_create"));
    if (size > 1) { /* NCCL communicator over the ranks of the matrix: id made on rank 0, broadcast over MPI */
      if (rank == 0) PetscCall(WbCheck(wb_comm_unique_id(id), "wb_comm_unique_id"));
      PetscCallMPI(MPI_Bcast(id, 128, MPI_BYTE, 0, comm));
      PetscCall(WbCheck(wb_comm_init(ctx, rank, size, id), "wb_comm_init"));
    }
  }
  *out = ctx;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ghost exchange plan of MatMult_MPIBAIJ's VecScatter, rebuilt from garray (global block column of every ghost
   column, ascending, hence grouped by owner): receive lists follow directly, send lists come from one all-to-all */
static PetscErrorCode WbHaloFromGarray(Mat P, WbPC *w, const PetscInt *garray, PetscInt nghost) {
  MPI_Comm comm;
  PetscMPIInt rank, size;
  const PetscInt *ranges;
  PetscInt bs = w->bs, k, r, nneigh = 0, nsend = 0;
  int *rcount, *scount, *rdisp, *sdisp;
  PetscInt *want, *give;
  int32_t *neigh, *send_ptr, *send_idx, *recv_ptr, *recv_idx;
  PetscFunctionBegin;
  PetscCall(PetscObjectGetComm((PetscObject)P, &comm));
  PetscCallMPI(MPI_Comm_rank(comm, &rank));
  PetscCallMPI(MPI_Comm_size(comm, &size));
  PetscCall(MatGetOwnershipRanges(P, &ranges)); /* scalar rows: block row = row / bs */
  PetscCall(PetscCalloc4(size, &rcount, size, &scount, size + 1, &rdisp, size + 1, &sdisp));
  for (k = 0, r = 0; k < nghost; k++) {
    while (garray[k] * bs >= ranges[r + 1]) r++;
    rcount[r]++;
  }
  PetscCallMPI(MPI_Alltoall(rcount, 1, MPI_INT, scount, 1, MPI_INT, comm));
  for (r = 0; r < size; r++) {
    rdisp[r + 1] = rdisp[r] + rcount[r];
    sdisp[r + 1] = sdisp[r] + scount[r];
    if (rcount[r] || scount[r]) nneigh++;
  }
  nsend = sdisp[size];
  PetscCall(PetscMalloc2(nghost, &want, nsend, &give));
  for (k = 0; k < nghost; k++) want[k] = garray[k];
  PetscCallMPI(MPI_Alltoallv(want, rcount, rdisp, MPIU_INT, give, scount, sdisp, MPIU_INT, comm));
  PetscCall(PetscMalloc5(nneigh, &neigh, nneigh + 1, &send_ptr, nsend, &send_idx, nneigh + 1, &recv_ptr, nghost, &recv_idx));
  send_ptr[0] = recv_ptr[0] = 0;
  for (r = 0, k = 0; r < size; r++) {
    PetscInt q;
    if (!rcount[r] && !scount[r]) continue;
    neigh[k] = (int32_t)r;
    for (q = sdisp[r]; q < sdisp[r + 1]; q++) send_idx[q] = (int32_t)(give[q] - ranges[rank] / bs); /* my local block row */
    for (q = rdisp[r]; q < rdisp[r + 1]; q++) recv_idx[q] = (int32_t)(w->nb + q);                    /* ghost column */
    send_ptr[k + 1] = (int32_t)sdisp[r + 1];
    recv_ptr[k + 1] = (int32_t)rdisp[r + 1];
    k++;
  }
  PetscCall(WbCheck(wb_set_halo(w->ctx, (int)nneigh, neigh, send_ptr, send_idx, recv_ptr, recv_idx), "wb_set_halo"));
  PetscCall(WbCheck(wb_set_global_offset(w->ctx, ranges[rank] / bs, ranges[size] / bs), "wb_set_global_offset"));
  { /* NVLink peer memory for the per-iteration exchanges (halo entries, Gram-Schmidt sums): every rank exports its
       CUDA-IPC blob, all ranks open all of them; on failure (no peer access, ranks on different nodes) the context
       stays on NCCL and the launch-per-operation solver */
    int blob = wb_comm_p2p_blob_size();
    unsigned char *mine, *all;
    PetscCall(PetscMalloc2(blob, &mine, (size_t)blob * size, &all));
    if (wb_comm_p2p_export(w->ctx, mine) == 0) {
      PetscCallMPI(MPI_Allgather(mine, blob, MPI_BYTE, all, blob, MPI_BYTE, comm));
      if (wb_comm_p2p_open(w->ctx, all) != 0) PetscCall(WbCheck(wb_comm_p2p_disable(w->ctx), "wb_comm_p2p_disable"));
    } else {
      PetscCall(PetscMemzero(mine, blob)); /* keep the collective matched: an all-zero blob makes every rank fall back */
      PetscCallMPI(MPI_Allgather(mine, blob, MPI_BYTE, all, blob, MPI_BYTE, comm));
      PetscCall(WbCheck(wb_comm_p2p_disable(w->ctx), "wb_comm_p2p_disable"));
    }
    PetscCall(PetscFree2(mine, all));
  }
  PetscCall(PetscFree5(neigh, send_ptr, send_idx, recv_ptr, recv_idx));
  PetscCall(PetscFree2(want, give));
  PetscCall(PetscFree4(rcount, scount, rdisp, sdisp));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* the rank's rows in BAIJ form with local column numbering: diagonal part (columns < nb) followed by the off-diagonal
   part (ghost column k -> nb + k); both parts have sorted columns, so the merged rows are sorted as well */
static PetscErrorCode WbGatherRows(Mat P, WbPC *w, PetscBool pattern) {
  Mat Ad = P, Ao = NULL;
  const PetscInt *garray = NULL, *ia, *ja, *io = NULL, *jo = NULL;
  PetscInt nb, nbo = 0, nghost = 0, i, k, bs2;
  PetscScalar *va, *vo = NULL;
  PetscBool done, mpi;
  PetscFunctionBegin;
  PetscCall(PetscObjectTypeCompare((PetscObject)P, MATMPIBAIJ, &mpi));
  if (mpi) PetscCall(MatMPIBAIJGetSeqBAIJ(P, &Ad, &Ao, &garray));
  PetscCall(MatGetBlockSize(P, &w->bs));
  bs2 = w->bs * w->bs;
  PetscCall(MatGetRowIJ(Ad, 0, PETSC_FALSE, PETSC_TRUE, &nb, &ia, &ja, &done));
  PetscCheck(done, PETSC_COMM_SELF, PETSC_ERR_SUP, "wbpc needs a (MPI)BAIJ matrix");
  if (Ao) {
    PetscCall(MatGetRowIJ(Ao, 0, PETSC_FALSE, PETSC_TRUE, &nbo, &io, &jo, &done));
    PetscCall(MatGetSize(Ao, NULL, &nghost));
    nghost /= w->bs;
  }
  PetscCall(MatSeqBAIJGetArray(Ad, &va));
  if (Ao) PetscCall(MatSeqBAIJGetArray(Ao, &vo));
  if (pattern) {
    w->nb = nb;
    w->ncolb = nb + nghost;
    w->nnzb = ia[nb] + (Ao ? io[nb] : 0);
    PetscCall(PetscMalloc3(nb + 1, &w->rowptr, w->nnzb, &w->colidx, w->nnzb * bs2, &w->vals));
    if (Ao) PetscCall(WbHaloFromGarray(P, w, garray, nghost));
  }
  for (i = 0, k = 0; i < nb; i++) {
    PetscInt q;
    w->rowptr[i] = (int32_t)k;
    for (q = ia[i]; q < ia[i + 1]; q++, k++) {
      w->colidx[k] = (int32_t)ja[q];
      PetscCall(PetscArraycpy(w->vals + k * bs2, va + q * bs2, bs2)); /* BAIJ blocks are column-major on both sides */
    }
    if (Ao)
      for (q = io[i]; q < io[i + 1]; q++, k++) {
        w->colidx[k] = (int32_t)(nb + jo[q]);
        PetscCall(PetscArraycpy(w->vals + k * bs2, vo + q * bs2, bs2));
      }
  }
  w->rowptr[nb] = (int32_t)k;
  PetscCall(MatSeqBAIJRestoreArray(Ad, &va));
  if (Ao) PetscCall(MatSeqBAIJRestoreArray(Ao, &vo));
  PetscCall(MatRestoreRowIJ(Ad, 0, PETSC_FALSE, PETSC_TRUE, &nb, &ia, &ja, &done));
  if (Ao) PetscCall(MatRestoreRowIJ(Ao, 0, PETSC_FALSE, PETSC_TRUE, &nbo, &io, &jo, &done));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* PCSetUp: symbolic part once per pattern, numeric part (values + ILU(0) factorisation) every Newton iteration */
static PetscErrorCode PCSetUp_WB(PC pc) {
  WbPC *w = (WbPC *)pc->data;
  Mat P;
  PetscFunctionBegin;
  PetscCall(PCGetOperators(pc, NULL, &P));
  if (!w->A) {
    MPI_Comm comm;
    int32_t *block_of_row = NULL;
    PetscInt nblocks = w->local_blocks, i;
    PetscCall(PetscObjectGetComm((PetscObject)P, &comm));
    PetscCall(WbContext(comm, &w->ctx));
    PetscCall(WbGatherRows(P, w, PETSC_TRUE));
    PetscCall(WbCheck(wb_mat_create(w->ctx, (int)w->nb, (int)w->ncolb, (int)w->bs, (int)w->nnzb, w->rowptr, w->colidx, w->vals, &w->A),
                      "wb_mat_create"));
    if (w->cube > 0) { /* sub-domains of a fixed number of consecutive rows */
      PetscCall(PetscMalloc1(w->nb, &block_of_row));
      for (i = 0; i < w->nb; i++) block_of_row[i] = (int32_t)(i / w->cube);
      nblocks = (w->nb + w->cube - 1) / w->cube;
    }
    PetscCall(WbCheck(wb_pc_setup(w->A, w->overlap > 0 ? WB_PC_ASM_ILU0 : WB_PC_BJACOBI_ILU0, (int)nblocks, block_of_row, &w->pc), "wb_pc_setup"));
    PetscCall(PetscFree(block_of_row));
  } else {
    int rc;
    PetscCall(WbGatherRows(P, w, PETSC_FALSE));
    PetscCall(WbCheck(wb_mat_set_values(w->A, w->vals), "wb_mat_set_values"));
    rc = wb_pc_refactor(w->pc);
    PetscCall(WbCheck(rc, "wb_pc_refactor"));
    if (rc > 0) pc->failedreason = PC_FACTOR_NUMERIC_ZEROPIVOT; /* singular diagonal block: KSP reports PC failure */
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode PCApply_WB(PC pc, Vec r, Vec z) {
  WbPC *w = (WbPC *)pc->data;
  const PetscScalar *ra;
  PetscScalar *za;
  PetscFunctionBegin;
  PetscCall(VecGetArrayRead(r, &ra));
  PetscCall(VecGetArray(z, &za));
  PetscCall(WbCheck(wb_pc_apply(w->pc, ra, za), "wb_pc_apply"));
  PetscCall(VecRestoreArray(z, &za));
  PetscCall(VecRestoreArrayRead(r, &ra));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode PCSetFromOptions_WB(PC pc, PetscOptionItems *items) {
  WbPC *w = (WbPC *)pc->data;
  PetscFunctionBegin;
  PetscOptionsHeadBegin(items, "waiwera_b200 block-Jacobi / ILU(0) options");
  PetscCall(PetscOptionsInt("-pc_wb_local_blocks", "block-Jacobi sub-domains per GPU", "PCBJacobiSetLocalBlocks", w->local_blocks,
                            &w->local_blocks, NULL));
  PetscCall(PetscOptionsInt("-pc_wb_subdomain_rows", "rows per sub-domain (0: use -pc_wb_local_blocks)", "", w->cube, &w->cube, NULL));
  PetscCall(PetscOptionsInt("-pc_wb_overlap", "0: block Jacobi; 1: restricted additive Schwarz, overlap 1", "PCASMSetOverlap", w->overlap,
                            &w->overlap, NULL));
  PetscOptionsHeadEnd();
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode PCDestroy_WB(PC pc) {
  WbPC *w = (WbPC *)pc->data;
  PetscFunctionBegin;
  if (w->pc) wb_pc_destroy(w->pc);
  if (w->A) wb_mat_destroy(w->A);
  PetscCall(PetscFree3(w->rowptr, w->colidx, w->vals));
  PetscCall(PetscFree(pc->data));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PETSC_EXTERN PetscErrorCode PCCreate_WB(PC pc) {
  WbPC *w;
  PetscFunctionBegin;
  PetscCall(PetscNew(&w));
  w->local_blocks = 1;
  pc->data = (void *)w;
  pc->ops->setup = PCSetUp_WB;
  pc->ops->apply = PCApply_WB;
  pc->ops->setfromoptions = PCSetFromOptions_WB;
  pc->ops->destroy = PCDestroy_WB;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ---- KSP: the whole Krylov solve on the device, one persistent kernel (GMRES; -ksp_wb_bcgs: BiCGStab) ---- */
typedef struct {
  PetscInt type, restart;
} WbKSP;

static PetscErrorCode KSPSetUp_WB(KSP ksp) {
  PetscFunctionBegin;
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode KSPSolve_WB(KSP ksp) {
  WbKSP *k = (WbKSP *)ksp->data;
  PC pc;
  WbPC *w;
  PetscBool iswb;
  wb_ksp_opts o;
  const PetscScalar *b;
  PetscScalar *x;
  int its = 0, reason = 0;
  double rnorm = 0.0;
  PetscFunctionBegin;
  PetscCall(KSPGetPC(ksp, &pc));
  PetscCall(PetscObjectTypeCompare((PetscObject)pc, "wbpc", &iswb));
  PetscCheck(iswb, PetscObjectComm((PetscObject)ksp), PETSC_ERR_SUP, "-ksp_type wbksp needs -pc_type wbpc (the factors live on the GPU)");
  w = (WbPC *)pc->data;
  o.type = (int)k->type;
  o.restart = (int)k->restart;
  o.maxit = (int)ksp->max_it;
  o.rtol = ksp->rtol;
  o.atol = ksp->abstol;
  o.dtol = ksp->divtol;
  PetscCall(VecGetArrayRead(ksp->vec_rhs, &b));
  PetscCall(VecGetArray(ksp->vec_sol, &x));
  /* zero initial guess, left preconditioning, convergence on the preconditioned residual norm: Waiwera's settings */
  PetscCall(WbCheck(wb_ksp_solve(w->A, w->pc, &o, b, x, &its, &reason, &rnorm), "wb_ksp_solve"));
  PetscCall(VecRestoreArray(ksp->vec_sol, &x));
  PetscCall(VecRestoreArrayRead(ksp->vec_rhs, &b));
  ksp->its = its;
  ksp->rnorm = rnorm;
  ksp->reason = (KSPConvergedReason)reason; /* PETSc's own values: 2 / 3 rtol / atol, -3 its, -4 dtol, -5 breakdown, -9 NaN */
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode KSPSetFromOptions_WB(KSP ksp, PetscOptionItems *items) {
  WbKSP *k = (WbKSP *)ksp->data;
  PetscBool bcgs = PETSC_FALSE;
  PetscFunctionBegin;
  PetscOptionsHeadBegin(items, "waiwera_b200 Krylov options");
  PetscCall(PetscOptionsBool("-ksp_wb_bcgs", "BiCGStab instead of GMRES", "KSPBCGS", bcgs, &bcgs, NULL));
  PetscCall(PetscOptionsInt("-ksp_gmres_restart", "GMRES restart", "KSPGMRESSetRestart", k->restart, &k->restart, NULL));
  PetscOptionsHeadEnd();
  if (bcgs) k->type = WB_KSP_BCGS;
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode KSPDestroy_WB(KSP ksp) {
  PetscFunctionBegin;
  PetscCall(PetscFree(ksp->data));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PETSC_EXTERN PetscErrorCode KSPCreate_WB(KSP ksp) {
  WbKSP *k;
  PetscFunctionBegin;
  PetscCall(PetscNew(&k));
  k->type = WB_KSP_GMRES;
  k->restart = 30;
  ksp->data = (void *)k;
  PetscCall(KSPSetSupportedNorm(ksp, KSP_NORM_PRECONDITIONED, PC_LEFT, 3));
  ksp->ops->setup = KSPSetUp_WB;
  ksp->ops->solve = KSPSolve_WB;
  ksp->ops->setfromoptions = KSPSetFromOptions_WB;
  ksp->ops->destroy = KSPDestroy_WB;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* entry point PETSc calls for a library named with -dll_append / -dll_prepend */
PETSC_EXTERN PetscErrorCode PetscDLLibraryRegister_wb_petsc_plugin(void) {
  PetscFunctionBegin;
  PetscCall(PCRegister("wbpc", PCCreate_WB));
  PetscCall(KSPRegister("wbksp", KSPCreate_WB));
  PetscFunctionReturn(PETSC_SUCCESS);
}
