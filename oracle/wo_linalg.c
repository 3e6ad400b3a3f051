/*
 * wo_linalg.c -- oracle (TEST INFRASTRUCTURE): the PETSc-side operators of the
 * Newton step, written to PETSc 3.22's documented semantics (PETSc is an
 * external dependency of the reference and is not in its tree; PARITY UNPINNED
 * at the operator level -- see SURVEY.md 8(c)):
 *   - BAIJ sparsity from FV adjacency        (src/dm_utils.F90:1041-1051, src/ode.F90:266-287)
 *   - MatMult_SeqBAIJ_N                      (called from KSP, src/timestepper.F90:1645-1836)
 *   - distance-2 colouring + MatFDColoringApply, MATMFFD_DS step
 *                                            (src/timestepper.F90:1584-1611, doc/user/setup_time.rst:434-452)
 *   - PCPBJACOBI, PCBJACOBI/ILU(0) natural ordering, inverted diagonal blocks
 *   - KSPGMRES (restart, classical Gram-Schmidt, left PC), KSPBCGS (left PC)
 * Blocks are stored column-major (PETSc BAIJ): val[b*bs*bs + j*bs + i] = A(i,j).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------- sparsity ---------------- */

static int cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}

wo_bsr *wo_bsr_from_mesh(const wo_mesh *mesh, int bs) {
  int nb = mesh->nowned;
  wo_bsr *A = (wo_bsr *)calloc(1, sizeof(wo_bsr));
  A->nb = nb;
  A->bs = bs;
  int32_t *cnt = (int32_t *)calloc(nb + 1, sizeof(int32_t));
  for (int i = 0; i < nb; i++) cnt[i + 1] = 1;
  for (int f = 0; f < mesh->nface; f++) {
    int c1 = mesh->face_cells[2 * f], c2 = mesh->face_cells[2 * f + 1];
    if (c1 < nb && c2 < nb) {
      cnt[c1 + 1]++;
      cnt[c2 + 1]++;
    }
  }
  A->rowptr = (int32_t *)malloc((nb + 1) * sizeof(int32_t));
  A->rowptr[0] = 0;
  for (int i = 0; i < nb; i++) A->rowptr[i + 1] = A->rowptr[i] + cnt[i + 1];
  A->nnzb = A->rowptr[nb];
  A->colidx = (int32_t *)malloc((size_t)A->nnzb * sizeof(int32_t));
  int32_t *fill = (int32_t *)malloc(nb * sizeof(int32_t));
  for (int i = 0; i < nb; i++) {
    fill[i] = A->rowptr[i];
    A->colidx[fill[i]++] = i;
  }
  for (int f = 0; f < mesh->nface; f++) {
    int c1 = mesh->face_cells[2 * f], c2 = mesh->face_cells[2 * f + 1];
    if (c1 < nb && c2 < nb) {
      A->colidx[fill[c1]++] = c2;
      A->colidx[fill[c2]++] = c1;
    }
  }
  for (int i = 0; i < nb; i++)
    qsort(A->colidx + A->rowptr[i], A->rowptr[i + 1] - A->rowptr[i], sizeof(int32_t), cmp_i32);
  A->val = (double *)calloc((size_t)A->nnzb * bs * bs, sizeof(double));
  free(cnt);
  free(fill);
  return A;
}

void wo_bsr_destroy(wo_bsr *A) {
  if (!A) return;
  free(A->rowptr);
  free(A->colidx);
  free(A->val);
  free(A);
}

static int bsr_find(const wo_bsr *A, int row, int col) {
  int lo = A->rowptr[row], hi = A->rowptr[row + 1] - 1;
  while (lo <= hi) {
    int mid = (lo + hi) / 2;
    if (A->colidx[mid] == col) return mid;
    if (A->colidx[mid] < col) lo = mid + 1;
    else hi = mid - 1;
  }
  return -1;
}

/* MatMult_SeqBAIJ_N: y_i = sum_j A_ij x_j, blocks in row order */
void wo_bsr_spmv(const wo_bsr *A, const double *x, double *y) {
  int bs = A->bs, bs2 = bs * bs;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < A->nb; i++) {
    double acc[WO_MAX_NP] = {0, 0, 0};
    for (int k = A->rowptr[i]; k < A->rowptr[i + 1]; k++) {
      const double *v = A->val + (size_t)k * bs2;
      const double *xb = x + (size_t)A->colidx[k] * bs;
      for (int jj = 0; jj < bs; jj++)
        for (int ii = 0; ii < bs; ii++) acc[ii] += v[jj * bs + ii] * xb[jj];
    }
    for (int ii = 0; ii < bs; ii++) y[(size_t)i * bs + ii] = acc[ii];
  }
}

/* greedy distance-2 colouring of block columns in natural order (pattern symmetric) */
int wo_bsr_coloring(const wo_bsr *A, int32_t *color) {
  int nb = A->nb, ncolor = 0;
  int cap = 64;
  int32_t *mark = (int32_t *)malloc(cap * sizeof(int32_t));
  for (int i = 0; i < cap; i++) mark[i] = -1;
  for (int c = 0; c < nb; c++) color[c] = -1;
  for (int c = 0; c < nb; c++) {
    /* columns at distance <= 2: share a row with c. rows containing c = neighbours of c */
    for (int k = A->rowptr[c]; k < A->rowptr[c + 1]; k++) {
      int r = A->colidx[k];
      for (int k2 = A->rowptr[r]; k2 < A->rowptr[r + 1]; k2++) {
        int c2 = A->colidx[k2];
        if (color[c2] >= 0) mark[color[c2]] = c;
      }
    }
    int col = 0;
    while (col < ncolor && mark[col] == c) col++;
    if (col == ncolor) {
      ncolor++;
      if (ncolor >= cap) {
        mark = (int32_t *)realloc(mark, 2 * cap * sizeof(int32_t));
        for (int i = cap; i < 2 * cap; i++) mark[i] = -1;
        cap *= 2;
      }
    }
    color[c] = col;
  }
  free(mark);
  return ncolor;
}

/* MatFDColoringApply_BAIJ, htype "ds" */
int wo_fd_jacobian(wo_flow *f, const double *y, const double *lhs_last, double dt, const double *F0,
                   const int32_t *color, int ncolor, double fd_err, double fd_umin, wo_bsr *J) {
  int nb = J->nb, bs = J->bs, bs2 = bs * bs;
  size_t n = (size_t)nb * bs;
  double *w3 = (double *)malloc(n * sizeof(double));
  double *w2 = (double *)malloc(n * sizeof(double));
  double *lhs = (double *)malloc(n * sizeof(double));
  double *rhs = (double *)malloc(n * sizeof(double));
  double *vscale = (double *)malloc(n * sizeof(double));
  int32_t *cols = (int32_t *)malloc(nb * sizeof(int32_t));
  int err = 0;
  memset(J->val, 0, (size_t)J->nnzb * bs2 * sizeof(double));
  for (int k = 0; k < ncolor && !err; k++) {
    int ncols = 0;
    for (int c = 0; c < nb; c++)
      if (color[c] == k) cols[ncols++] = c;
    for (int i = 0; i < bs && !err; i++) {
      memcpy(w3, y, n * sizeof(double));
      for (int l = 0; l < ncols; l++) {
        size_t col = (size_t)i + (size_t)bs * cols[l];
        double dx = y[col];
        if (dx == 0.0) dx = 1.0;
        if (fabs(dx) < fd_umin && dx >= 0.0) dx = fd_umin;
        else if (dx < 0.0 && fabs(dx) < fd_umin) dx = -fd_umin;
        dx *= fd_err;
        vscale[col] = 1.0 / dx;
        w3[col] += dx;
      }
      err = wo_residual_be(f, w3, lhs_last, dt, cols, ncols, lhs, rhs, w2);
      if (err) break;
#pragma omp parallel for schedule(static) if (n >= 20000)
      for (size_t q = 0; q < n; q++) w2[q] = w2[q] + (-1.0) * F0[q];
      /* columns of one colour touch disjoint rows: independent */
#pragma omp parallel for schedule(static) if (ncols >= 1000)
      for (int l = 0; l < ncols; l++) {
        int c = cols[l];
        size_t col = (size_t)i + (size_t)bs * c;
        for (int kk = J->rowptr[c]; kk < J->rowptr[c + 1]; kk++) {
          int r = J->colidx[kk]; /* symmetric pattern: rows containing column c */
          int pos = bsr_find(J, r, c);
          if (pos < 0) continue;
          double *blk = J->val + (size_t)pos * bs2;
          for (int ii = 0; ii < bs; ii++) blk[i * bs + ii] = w2[(size_t)r * bs + ii] * vscale[col];
        }
      }
    }
  }
  free(w3);
  free(w2);
  free(lhs);
  free(rhs);
  free(vscale);
  free(cols);
  return err;
}

/* ---------------- small dense block kernels (column-major) ---------------- */

static int blk_invert(const double *a, double *inv, int bs) {
  /* Gauss-Jordan with partial pivoting */
  double m[WO_MAX_NP][2 * WO_MAX_NP];
  for (int i = 0; i < bs; i++)
    for (int j = 0; j < bs; j++) {
      m[i][j] = a[j * bs + i];
      m[i][bs + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < bs; c++) {
    int piv = c;
    for (int r = c + 1; r < bs; r++)
      if (fabs(m[r][c]) > fabs(m[piv][c])) piv = r;
    if (m[piv][c] == 0.0) return 1;
    if (piv != c)
      for (int j = 0; j < 2 * bs; j++) {
        double t = m[c][j];
        m[c][j] = m[piv][j];
        m[piv][j] = t;
      }
    double d = 1.0 / m[c][c];
    for (int j = 0; j < 2 * bs; j++) m[c][j] *= d;
    for (int r = 0; r < bs; r++)
      if (r != c) {
        double fct = m[r][c];
        if (fct != 0.0)
          for (int j = 0; j < 2 * bs; j++) m[r][j] -= fct * m[c][j];
      }
  }
  for (int i = 0; i < bs; i++)
    for (int j = 0; j < bs; j++) inv[j * bs + i] = m[i][bs + j];
  return 0;
}

/* C = A*B */
static void blk_mul(const double *a, const double *b, double *c, int bs) {
  for (int j = 0; j < bs; j++)
    for (int i = 0; i < bs; i++) {
      double s = 0.0;
      for (int k = 0; k < bs; k++) s += a[k * bs + i] * b[j * bs + k];
      c[j * bs + i] = s;
    }
}
/* C -= A*B */
static void blk_mulsub(const double *a, const double *b, double *c, int bs) {
  for (int j = 0; j < bs; j++)
    for (int i = 0; i < bs; i++) {
      double s = 0.0;
      for (int k = 0; k < bs; k++) s += a[k * bs + i] * b[j * bs + k];
      c[j * bs + i] -= s;
    }
}

/* ---------------- preconditioners ---------------- */

struct wo_pc {
  int type, nb, bs;
  /* factor storage: same pattern as the (sub-domain restricted) matrix */
  int32_t *rowptr, *colidx, *diag;
  double *val; /* L (multipliers), inverted diagonal, U */
  /* sub-domains (block Jacobi): rows of block b are blk_rows[blk_ptr[b]..blk_ptr[b+1]), ascending.
     Sub-domains are independent, so they are factored / solved on separate host threads
     (one MPI rank per block in the reference); per row the arithmetic is the sequential one. */
  int nblocks;
  int32_t *blk_ptr, *blk_rows;
  /* additive Schwarz: the extended sub-domain matrices side by side (one block-diagonal matrix `ext`), the block-Jacobi
     ILU(0) of it, and the maps between the two numberings */
  wo_bsr *ext;
  wo_pc *inner;
  int32_t *ext_row; /* [ext->nb] global row of each extended row */
  int32_t *own_pos; /* [nb] extended row that holds the owned copy of a global row */
  double *ext_r, *ext_z;
};

/* PCASM with overlap 1 (PCSetUp_ASM: MatIncreaseOverlap by one layer of matrix neighbours, MatCreateSubMatrices on the
   sorted index sets, sub-KSP preonly + ILU(0); PCApply_ASM with PC_ASM_RESTRICT: the residual is restricted to the
   extended sub-domain, only the rows the sub-domain owns are written back). */
static int asm_build(wo_pc *pc, const wo_bsr *A, const int32_t *block_of_row) {
  int nb = A->nb, bs2 = A->bs * A->bs;
  int nblk = 1;
  if (block_of_row)
    for (int i = 0; i < nb; i++)
      if (block_of_row[i] + 1 > nblk) nblk = block_of_row[i] + 1;
  /* rows of the extended sub-domains, ascending inside each: a row belongs to E_b when it or one of its matrix
     neighbours is owned by b (symmetric pattern: the blocks that reach row i through one matrix entry are the blocks
     of the columns of row i) */
  int32_t *start = (int32_t *)calloc(nblk + 1, sizeof(int32_t));
  int32_t *cnt = (int32_t *)malloc(nblk * sizeof(int32_t));
  int32_t *seen = (int32_t *)malloc(nblk * sizeof(int32_t));
  for (int pass = 0; pass < 2; pass++) {
    for (int b = 0; b < nblk; b++) seen[b] = -1;
    for (int i = 0; i < nb; i++)
      for (int k = A->rowptr[i]; k < A->rowptr[i + 1]; k++) {
        int b = block_of_row ? block_of_row[A->colidx[k]] : 0;
        if (seen[b] == i) continue;
        seen[b] = i;
        if (pass == 0) start[b + 1]++;
        else pc->ext_row[cnt[b]++] = i;
      }
    if (pass == 0) {
      for (int b = 0; b < nblk; b++) start[b + 1] += start[b];
      for (int b = 0; b < nblk; b++) cnt[b] = start[b];
      pc->ext_row = (int32_t *)malloc((size_t)start[nblk] * sizeof(int32_t));
    }
  }
  int ne = start[nblk];
  wo_bsr *E = (wo_bsr *)calloc(1, sizeof(wo_bsr));
  E->nb = ne;
  E->bs = A->bs;
  E->rowptr = (int32_t *)malloc((size_t)(ne + 1) * sizeof(int32_t));
  pc->own_pos = (int32_t *)malloc((size_t)nb * sizeof(int32_t));
  int32_t *blk_ext = (int32_t *)malloc((size_t)ne * sizeof(int32_t));
  int32_t *loc = (int32_t *)malloc((size_t)nb * sizeof(int32_t)); /* global row -> extended row of the current block */
  for (int i = 0; i < nb; i++) loc[i] = -1;
  for (int pass = 0; pass < 2; pass++) {
    int q = 0;
    for (int b = 0; b < nblk; b++) {
      for (int e = start[b]; e < start[b + 1]; e++) loc[pc->ext_row[e]] = e;
      for (int e = start[b]; e < start[b + 1]; e++) {
        int i = pc->ext_row[e];
        if (pass == 0) {
          E->rowptr[e] = q;
          blk_ext[e] = b;
          if ((block_of_row ? block_of_row[i] : 0) == b) pc->own_pos[i] = e;
        }
        for (int k = A->rowptr[i]; k < A->rowptr[i + 1]; k++) {
          int le = loc[A->colidx[k]];
          if (le < 0) continue;
          if (pass == 1) {
            E->colidx[q] = le;
            memcpy(E->val + (size_t)q * bs2, A->val + (size_t)k * bs2, bs2 * sizeof(double));
          }
          q++;
        }
      }
      for (int e = start[b]; e < start[b + 1]; e++) loc[pc->ext_row[e]] = -1;
    }
    if (pass == 0) {
      E->rowptr[ne] = q;
      E->nnzb = q;
      E->colidx = (int32_t *)malloc((size_t)q * sizeof(int32_t));
      E->val = (double *)malloc((size_t)q * bs2 * sizeof(double));
    }
  }
  pc->ext = E;
  pc->inner = wo_pc_create(E, WO_PC_BJACOBI_ILU0, blk_ext);
  pc->ext_r = (double *)malloc((size_t)ne * A->bs * sizeof(double));
  pc->ext_z = (double *)malloc((size_t)ne * A->bs * sizeof(double));
  free(cnt);
  free(seen);
  free(start);
  free(blk_ext);
  free(loc);
  return pc->inner ? 0 : 1;
}

wo_pc *wo_pc_create(const wo_bsr *A, int type, const int32_t *block_of_row) {
  int nb = A->nb, bs = A->bs, bs2 = bs * bs;
  wo_pc *pc = (wo_pc *)calloc(1, sizeof(wo_pc));
  pc->type = type;
  pc->nb = nb;
  pc->bs = bs;
  if (type == WO_PC_NONE) return pc;
  if (type == WO_PC_PBJACOBI) {
    pc->val = (double *)malloc((size_t)nb * bs2 * sizeof(double));
    for (int i = 0; i < nb; i++) {
      int d = bsr_find(A, i, i);
      if (blk_invert(A->val + (size_t)d * bs2, pc->val + (size_t)i * bs2, bs)) {
        wo_pc_destroy(pc);
        return NULL;
      }
    }
    return pc;
  }
  if (type == WO_PC_ASM_ILU0) {
    if (asm_build(pc, A, block_of_row)) {
      wo_pc_destroy(pc);
      return NULL;
    }
    return pc;
  }
  /* ILU(0) on the block-diagonal (sub-domain) restriction, natural ordering */
  pc->rowptr = (int32_t *)malloc((nb + 1) * sizeof(int32_t));
  pc->diag = (int32_t *)malloc(nb * sizeof(int32_t));
  pc->rowptr[0] = 0;
  for (int i = 0; i < nb; i++) {
    int cnt = 0;
    for (int k = A->rowptr[i]; k < A->rowptr[i + 1]; k++)
      if (!block_of_row || block_of_row[A->colidx[k]] == block_of_row[i]) cnt++;
    pc->rowptr[i + 1] = pc->rowptr[i] + cnt;
  }
  int nnzb = pc->rowptr[nb];
  pc->colidx = (int32_t *)malloc((size_t)nnzb * sizeof(int32_t));
  pc->val = (double *)malloc((size_t)nnzb * bs2 * sizeof(double));
  for (int i = 0; i < nb; i++) {
    int q = pc->rowptr[i];
    for (int k = A->rowptr[i]; k < A->rowptr[i + 1]; k++)
      if (!block_of_row || block_of_row[A->colidx[k]] == block_of_row[i]) {
        pc->colidx[q] = A->colidx[k];
        memcpy(pc->val + (size_t)q * bs2, A->val + (size_t)k * bs2, bs2 * sizeof(double));
        if (A->colidx[k] == i) pc->diag[i] = q;
        q++;
      }
  }
  /* sub-domain row lists */
  {
    int nblk = 1;
    if (block_of_row)
      for (int i = 0; i < nb; i++)
        if (block_of_row[i] + 1 > nblk) nblk = block_of_row[i] + 1;
    pc->nblocks = nblk;
    pc->blk_ptr = (int32_t *)calloc(nblk + 1, sizeof(int32_t));
    pc->blk_rows = (int32_t *)malloc(nb * sizeof(int32_t));
    for (int i = 0; i < nb; i++) pc->blk_ptr[(block_of_row ? block_of_row[i] : 0) + 1]++;
    for (int b = 0; b < nblk; b++) pc->blk_ptr[b + 1] += pc->blk_ptr[b];
    int32_t *fill = (int32_t *)malloc(nblk * sizeof(int32_t));
    for (int b = 0; b < nblk; b++) fill[b] = pc->blk_ptr[b];
    for (int i = 0; i < nb; i++) pc->blk_rows[fill[block_of_row ? block_of_row[i] : 0]++] = i;
    free(fill);
  }
  /* IKJ block ILU(0): MatILUFactorNumeric_SeqBAIJ_N_NaturalOrdering semantics, sub-domain by sub-domain */
  int failed = 0;
#pragma omp parallel
  {
    int32_t *pos = (int32_t *)malloc(nb * sizeof(int32_t));
    for (int i = 0; i < nb; i++) pos[i] = -1;
    double mult[WO_MAX_NP * WO_MAX_NP], inv[WO_MAX_NP * WO_MAX_NP];
#pragma omp for schedule(dynamic, 1)
    for (int b = 0; b < pc->nblocks; b++) {
      for (int q = pc->blk_ptr[b]; q < pc->blk_ptr[b + 1]; q++) {
        int i = pc->blk_rows[q];
        for (int k = pc->rowptr[i]; k < pc->rowptr[i + 1]; k++) pos[pc->colidx[k]] = k;
        for (int k = pc->rowptr[i]; k < pc->diag[i]; k++) {
          int kr = pc->colidx[k];
          double *aik = pc->val + (size_t)k * bs2;
          /* multiplier = A_ik * inv(A_kk) (diag of row kr already inverted) */
          blk_mul(aik, pc->val + (size_t)pc->diag[kr] * bs2, mult, bs);
          memcpy(aik, mult, bs2 * sizeof(double));
          for (int qq = pc->diag[kr] + 1; qq < pc->rowptr[kr + 1]; qq++) {
            int p = pos[pc->colidx[qq]];
            if (p >= 0) blk_mulsub(mult, pc->val + (size_t)qq * bs2, pc->val + (size_t)p * bs2, bs);
          }
        }
        if (blk_invert(pc->val + (size_t)pc->diag[i] * bs2, inv, bs)) {
#pragma omp atomic write
          failed = 1;
        } else
          memcpy(pc->val + (size_t)pc->diag[i] * bs2, inv, bs2 * sizeof(double));
        for (int k = pc->rowptr[i]; k < pc->rowptr[i + 1]; k++) pos[pc->colidx[k]] = -1;
      }
    }
    free(pos);
  }
  if (failed) {
    wo_pc_destroy(pc);
    return NULL;
  }
  return pc;
}

void wo_pc_apply(const wo_pc *pc, const double *r, double *z) {
  int nb = pc->nb, bs = pc->bs, bs2 = bs * bs;
  if (pc->type == WO_PC_NONE) {
    memcpy(z, r, (size_t)nb * bs * sizeof(double));
    return;
  }
  if (pc->type == WO_PC_PBJACOBI) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nb; i++) {
      const double *d = pc->val + (size_t)i * bs2;
      for (int ii = 0; ii < bs; ii++) {
        double s = 0.0;
        for (int jj = 0; jj < bs; jj++) s += d[jj * bs + ii] * r[(size_t)i * bs + jj];
        z[(size_t)i * bs + ii] = s;
      }
    }
    return;
  }
  if (pc->type == WO_PC_ASM_ILU0) {
    int ne = pc->ext->nb;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < ne; e++)
      for (int ii = 0; ii < bs; ii++) pc->ext_r[(size_t)e * bs + ii] = r[(size_t)pc->ext_row[e] * bs + ii];
    wo_pc_apply(pc->inner, pc->ext_r, pc->ext_z);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nb; i++)
      for (int ii = 0; ii < bs; ii++) z[(size_t)i * bs + ii] = pc->ext_z[(size_t)pc->own_pos[i] * bs + ii];
    return;
  }
  /* MatSolve_SeqBAIJ_N_NaturalOrdering: forward (unit L), backward with inverted diagonal,
     one sub-domain per host thread */
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < pc->nblocks; b++) {
    for (int q = pc->blk_ptr[b]; q < pc->blk_ptr[b + 1]; q++) {
      int i = pc->blk_rows[q];
      double s[WO_MAX_NP];
      for (int ii = 0; ii < bs; ii++) s[ii] = r[(size_t)i * bs + ii];
      for (int k = pc->rowptr[i]; k < pc->diag[i]; k++) {
        const double *v = pc->val + (size_t)k * bs2;
        const double *xb = z + (size_t)pc->colidx[k] * bs;
        for (int jj = 0; jj < bs; jj++)
          for (int ii = 0; ii < bs; ii++) s[ii] -= v[jj * bs + ii] * xb[jj];
      }
      for (int ii = 0; ii < bs; ii++) z[(size_t)i * bs + ii] = s[ii];
    }
    for (int q = pc->blk_ptr[b + 1] - 1; q >= pc->blk_ptr[b]; q--) {
      int i = pc->blk_rows[q];
      double s[WO_MAX_NP], t[WO_MAX_NP];
      for (int ii = 0; ii < bs; ii++) s[ii] = z[(size_t)i * bs + ii];
      for (int k = pc->diag[i] + 1; k < pc->rowptr[i + 1]; k++) {
        const double *v = pc->val + (size_t)k * bs2;
        const double *xb = z + (size_t)pc->colidx[k] * bs;
        for (int jj = 0; jj < bs; jj++)
          for (int ii = 0; ii < bs; ii++) s[ii] -= v[jj * bs + ii] * xb[jj];
      }
      const double *d = pc->val + (size_t)pc->diag[i] * bs2;
      for (int ii = 0; ii < bs; ii++) {
        double acc = 0.0;
        for (int jj = 0; jj < bs; jj++) acc += d[jj * bs + ii] * s[jj];
        t[ii] = acc;
      }
      for (int ii = 0; ii < bs; ii++) z[(size_t)i * bs + ii] = t[ii];
    }
  }
}

void wo_pc_destroy(wo_pc *pc) {
  if (!pc) return;
  free(pc->rowptr);
  free(pc->colidx);
  free(pc->diag);
  free(pc->val);
  free(pc->blk_ptr);
  free(pc->blk_rows);
  if (pc->inner) wo_pc_destroy(pc->inner);
  if (pc->ext) wo_bsr_destroy(pc->ext);
  free(pc->ext_row);
  free(pc->own_pos);
  free(pc->ext_r);
  free(pc->ext_z);
  free(pc);
}

/* ---------------- Krylov solvers ---------------- */

static double vdot(const double *a, const double *b, size_t n) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (size_t i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}
static void vaxpy(double *y, double a, const double *x, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) y[i] += a * x[i];
}

/* reason codes follow KSPConvergedReason */
#define KSP_CONVERGED_RTOL 2
#define KSP_CONVERGED_ATOL 3
#define KSP_CONVERGED_HAPPY_BREAKDOWN 5
#define KSP_DIVERGED_ITS (-3)
#define KSP_DIVERGED_DTOL (-4)
#define KSP_DIVERGED_BREAKDOWN (-5)

static int ksp_converged(double rnorm, double rnorm0, const wo_ksp_opts *o) {
  double ttol = fmax(o->rtol * rnorm0, o->atol);
  if (rnorm != rnorm) return -9;
  if (rnorm <= ttol) return (rnorm < o->atol) ? KSP_CONVERGED_ATOL : KSP_CONVERGED_RTOL;
  if (rnorm >= o->dtol * rnorm0) return KSP_DIVERGED_DTOL;
  return 0;
}

static void pc_amul(const wo_bsr *A, const wo_pc *pc, const double *v, double *tmp, double *out) {
  wo_bsr_spmv(A, v, tmp);
  wo_pc_apply(pc, tmp, out);
}

static int gmres_solve(const wo_bsr *A, const wo_pc *pc, const wo_ksp_opts *o, const double *b, double *x,
                       int *its_out, double *rnorm_out) {
  size_t n = (size_t)A->nb * A->bs;
  int m = o->restart > 0 ? o->restart : 30;
  double *V = (double *)malloc((size_t)(m + 1) * n * sizeof(double));
  double *H = (double *)calloc((size_t)(m + 1) * m, sizeof(double)); /* H[j + (m+1)*col] */
  double *cs = (double *)calloc(m + 1, sizeof(double)), *sn = (double *)calloc(m + 1, sizeof(double));
  double *rs = (double *)calloc(m + 2, sizeof(double)), *yv = (double *)calloc(m + 1, sizeof(double));
  double *tmp = (double *)malloc(n * sizeof(double)), *w = (double *)malloc(n * sizeof(double));
  int its = 0, reason = 0;
  double rnorm0 = -1.0, res = 0.0;
  memset(x, 0, n * sizeof(double));
  int first = 1;
  while (!reason) {
    /* r = M^-1 (b - A x) */
    if (first) {
      wo_pc_apply(pc, b, V);
    } else {
      wo_bsr_spmv(A, x, tmp);
      for (size_t i = 0; i < n; i++) tmp[i] = b[i] - tmp[i];
      wo_pc_apply(pc, tmp, V);
    }
    res = sqrt(vdot(V, V, n));
    if (first) {
      rnorm0 = res;
      first = 0;
      reason = ksp_converged(res, rnorm0, o);
      if (reason || res == 0.0) {
        if (!reason) reason = KSP_CONVERGED_ATOL;
        break;
      }
    }
    double inv = 1.0 / res;
    for (size_t i = 0; i < n; i++) V[i] *= inv;
    rs[0] = res;
    int it = 0;
    while (it < m && !reason) {
      double *vn = V + (size_t)(it + 1) * n;
      pc_amul(A, pc, V + (size_t)it * n, tmp, vn);
      double *hcol = H + (size_t)(m + 1) * it;
      /* classical Gram-Schmidt: all dots first, then one MAXPY */
      for (int j = 0; j <= it; j++) hcol[j] = vdot(vn, V + (size_t)j * n, n);
      for (int j = 0; j <= it; j++) vaxpy(vn, -hcol[j], V + (size_t)j * n, n);
      double tt = sqrt(vdot(vn, vn, n));
      hcol[it + 1] = tt;
      int happy = (tt < 1.e-30 * fmax(res, 1e-300)) || tt == 0.0;
      if (!happy) {
        double s = 1.0 / tt;
        for (size_t i = 0; i < n; i++) vn[i] *= s;
      }
      /* apply previous rotations, then a new one */
      for (int j = 0; j < it; j++) {
        double t1 = hcol[j], t2 = hcol[j + 1];
        hcol[j] = cs[j] * t1 + sn[j] * t2;
        hcol[j + 1] = -sn[j] * t1 + cs[j] * t2;
      }
      double hh = hcol[it], hp = hcol[it + 1];
      double den = sqrt(hh * hh + hp * hp);
      if (den == 0.0) {
        reason = KSP_DIVERGED_BREAKDOWN;
        break;
      }
      cs[it] = hh / den;
      sn[it] = hp / den;
      rs[it + 1] = -sn[it] * rs[it];
      rs[it] = cs[it] * rs[it];
      hcol[it] = cs[it] * hh + sn[it] * hp;
      hcol[it + 1] = 0.0;
      res = fabs(rs[it + 1]);
      it++;
      its++;
      reason = ksp_converged(res, rnorm0, o);
      if (!reason && happy) reason = KSP_CONVERGED_HAPPY_BREAKDOWN;
      if (!reason && its >= o->maxit) reason = KSP_DIVERGED_ITS;
    }
    /* form solution update */
    for (int k = it - 1; k >= 0; k--) {
      double s = rs[k];
      for (int j = k + 1; j < it; j++) s -= H[(size_t)(m + 1) * j + k] * yv[j];
      yv[k] = s / H[(size_t)(m + 1) * k + k];
    }
    for (int j = 0; j < it; j++) vaxpy(x, yv[j], V + (size_t)j * n, n);
  }
  *its_out = its;
  *rnorm_out = res;
  free(V);
  free(H);
  free(cs);
  free(sn);
  free(rs);
  free(yv);
  free(tmp);
  free(w);
  return reason;
}

static int bcgs_solve(const wo_bsr *A, const wo_pc *pc, const wo_ksp_opts *o, const double *b, double *x,
                      int *its_out, double *rnorm_out) {
  size_t n = (size_t)A->nb * A->bs;
  double *R = (double *)malloc(n * sizeof(double)), *RP = (double *)malloc(n * sizeof(double));
  double *P = (double *)calloc(n, sizeof(double)), *V = (double *)calloc(n, sizeof(double));
  double *S = (double *)malloc(n * sizeof(double)), *T = (double *)malloc(n * sizeof(double));
  double *tmp = (double *)malloc(n * sizeof(double));
  memset(x, 0, n * sizeof(double));
  wo_pc_apply(pc, b, R);
  double dp = sqrt(vdot(R, R, n)), rnorm0 = dp;
  int reason = ksp_converged(dp, rnorm0, o), i = 0;
  if (!reason && dp == 0.0) reason = KSP_CONVERGED_ATOL;
  memcpy(RP, R, n * sizeof(double));
  double rhoold = 1.0, alpha = 1.0, omegaold = 1.0;
  while (!reason) {
    double rho = vdot(R, RP, n);
    if (rho == 0.0) {
      reason = KSP_DIVERGED_BREAKDOWN;
      break;
    }
    double beta = (rho / rhoold) * (alpha / omegaold);
    for (size_t q = 0; q < n; q++) P[q] = R[q] + beta * (P[q] - omegaold * V[q]);
    pc_amul(A, pc, P, tmp, V);
    double d1 = vdot(V, RP, n);
    if (d1 == 0.0) {
      reason = KSP_DIVERGED_BREAKDOWN;
      break;
    }
    alpha = rho / d1;
    for (size_t q = 0; q < n; q++) S[q] = R[q] - alpha * V[q];
    pc_amul(A, pc, S, tmp, T);
    d1 = vdot(S, T, n);
    double d2 = vdot(T, T, n);
    if (d2 == 0.0) {
      vaxpy(x, alpha, P, n);
      i++;
      dp = 0.0;
      reason = KSP_CONVERGED_ATOL;
      break;
    }
    double omega = d1 / d2;
    for (size_t q = 0; q < n; q++) x[q] += alpha * P[q] + omega * S[q];
    for (size_t q = 0; q < n; q++) R[q] = S[q] - omega * T[q];
    dp = sqrt(vdot(R, R, n));
    rhoold = rho;
    omegaold = omega;
    i++;
    reason = ksp_converged(dp, rnorm0, o);
    if (!reason && i >= o->maxit) reason = KSP_DIVERGED_ITS;
  }
  *its_out = i;
  *rnorm_out = dp;
  free(R);
  free(RP);
  free(P);
  free(V);
  free(S);
  free(T);
  free(tmp);
  return reason;
}

int wo_ksp_solve(const wo_bsr *A, const wo_pc *pc, const wo_ksp_opts *o, const double *b, double *x, int *its,
                 double *rnorm) {
  if (o->type == WO_KSP_BCGS) return bcgs_solve(A, pc, o, b, x, its, rnorm);
  return gmres_solve(A, pc, o, b, x, its, rnorm);
}
