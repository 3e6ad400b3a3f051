/*
 * wo_tracer.c -- oracle (TEST INFRASTRUCTURE): the passive-tracer auxiliary linear problem
 * (SURVEY.md section 8 row f-4).  Restated from
 *   src/tracer.F90:48-61                      Arrhenius decay rate
 *   src/cell.F90:146-201                      tracer balance coefficients, diffusion factor, tortuosity
 *   src/face.F90:519-536                      harmonic diffusion factor of a face
 *   src/flow_simulation.F90:1489-1556         aux_lhs  (tracer_cell_balances, diagonal Al)
 *   src/flow_simulation.F90:1560-1833         aux_rhs  (tracer_cell_inflows: Ar, br; sources; decay)
 *   src/flow_simulation.F90:1837-1959         aux_pre_solve (absent phases and boundary rows)
 *   src/timestepper.F90:458-581               setup_linear of backward Euler / BDF2 / direct steady state
 * The linear system is solved with the same Krylov code as the Jacobian system (wo_linalg.c).
 *
 * Unknowns: nt tracer mass fractions per cell, rows of the owned cells [0, nowned) followed by the
 * boundary (Dirichlet) ghost cells [ninterior, ncell) -- the reference builds A_aux on mesh%dm, which
 * holds the boundary ghost cells (src/ode.F90:301-321).  Serial meshes only here (nowned == ninterior).
 * A is kept in the BAIJ layout (bs = nt, column-major blocks); only block diagonals are non-zero
 * because tracers do not couple.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "wo_flow_priv.h"

static inline int nint_(double x) { return (int)lround(x); }

#define WO_TC_K 273.15          /* src/thermodynamics.F90: tc_k */
#define WO_GAS_CONSTANT 8.3144598 /* src/thermodynamics.F90: gas_constant */

void wo_flow_set_tracers(wo_flow *f, int nt, const wo_tracer *tracers) {
  free(f->tracers);
  f->nt = nt;
  f->tracers = (wo_tracer *)malloc((nt + 1) * sizeof(wo_tracer));
  memcpy(f->tracers, tracers, nt * sizeof(wo_tracer));
  free(f->tracer_injection);
  f->tracer_injection = NULL;
}

void wo_flow_set_tracer_injection(wo_flow *f, const double *rate) {
  size_t n = (size_t)f->nsrc * f->nt;
  free(f->tracer_injection);
  f->tracer_injection = (double *)malloc((n + 1) * sizeof(double));
  memcpy(f->tracer_injection, rate, n * sizeof(double));
}

/* tracer_decay: src/tracer.F90:48-61 */
double wo_tracer_decay(const wo_tracer *t, double temperature) {
  double Tk = temperature + WO_TC_K;
  return t->decay * exp(-t->activation / (WO_GAS_CONSTANT * Tk));
}

static const double *phase_rec(const wo_flow *f, const double *fluid, int p) {
  return fluid + (7 + f->nc - 1) + p * (8 + f->nc - 1); /* density 0, viscosity 1, saturation 2, ... */
}

/* cell%tracer_balance_coefs: src/cell.F90:146-164 */
static double balance_coef(const wo_flow *f, const double *rock, const double *fluid, int phase) {
  const double *ph = phase_rec(f, fluid, phase - 1);
  return rock[5] * ph[2] * ph[0]; /* porosity * saturation * density */
}

/* cell%diffusion_factor: src/cell.F90:168-180 with cell%tortuosity :184-201 (rock tortuosity 1,
   fluid tortuosity = phase saturation) */
static double cell_diffusion_factor(const wo_flow *f, const double *rock, const double *fluid, int phase) {
  const double *ph = phase_rec(f, fluid, phase - 1);
  const double rock_tortuosity = 1.0;
  double tortuosity = rock_tortuosity * ph[2];
  return rock[5] * ph[0] * tortuosity;
}

/* row of cell c in the tracer system */
static int row_of(const wo_mesh *m, int c) { return c < m->nowned ? c : m->nowned + (c - m->ninterior); }

/* pattern of A_aux: FV adjacency over owned + boundary cells (DMCreateMatrix on the tracer DM) */
wo_bsr *wo_tracer_pattern(const wo_mesh *m, int nt) {
  int nb = m->nowned + (m->ncell - m->ninterior);
  int *deg = (int *)calloc(nb + 1, sizeof(int));
  for (int i = 0; i < nb; i++) deg[i] = 1;
  for (int fc = 0; fc < m->nface; fc++) {
    int r0 = row_of(m, m->face_cells[2 * fc]), r1 = row_of(m, m->face_cells[2 * fc + 1]);
    deg[r0]++;
    deg[r1]++;
  }
  wo_bsr *A = (wo_bsr *)calloc(1, sizeof(wo_bsr));
  A->nb = nb;
  A->bs = nt;
  A->rowptr = (int32_t *)malloc((nb + 1) * sizeof(int32_t));
  A->rowptr[0] = 0;
  for (int i = 0; i < nb; i++) A->rowptr[i + 1] = A->rowptr[i] + deg[i];
  int32_t *col = (int32_t *)malloc((size_t)A->rowptr[nb] * sizeof(int32_t));
  int *fill = (int *)calloc(nb + 1, sizeof(int));
  for (int i = 0; i < nb; i++) col[A->rowptr[i] + fill[i]++] = i;
  for (int fc = 0; fc < m->nface; fc++) {
    int r0 = row_of(m, m->face_cells[2 * fc]), r1 = row_of(m, m->face_cells[2 * fc + 1]);
    col[A->rowptr[r0] + fill[r0]++] = r1;
    col[A->rowptr[r1] + fill[r1]++] = r0;
  }
  /* sort + unique per row */
  int32_t *rp = (int32_t *)malloc((nb + 1) * sizeof(int32_t));
  int nnz = 0;
  rp[0] = 0;
  for (int i = 0; i < nb; i++) {
    int32_t *r = col + A->rowptr[i];
    int n = fill[i];
    for (int a = 1; a < n; a++) {
      int32_t v = r[a];
      int b = a - 1;
      while (b >= 0 && r[b] > v) { r[b + 1] = r[b]; b--; }
      r[b + 1] = v;
    }
    int32_t last = -1;
    for (int a = 0; a < n; a++)
      if (r[a] != last) { col[nnz++] = r[a]; last = r[a]; }
    rp[i + 1] = nnz;
  }
  free(A->rowptr);
  A->rowptr = rp;
  A->colidx = (int32_t *)realloc(col, (size_t)(nnz + 1) * sizeof(int32_t));
  A->nnzb = nnz;
  A->val = (double *)calloc((size_t)nnz * nt * nt + 1, sizeof(double));
  free(deg);
  free(fill);
  return A;
}

/* MatSetValuesLocal(A, irow, icol, v, ADD_VALUES) on entry (it, it) of block (row, col) */
static void add_value(wo_bsr *A, int row, int col, int it, double v) {
  int nt = A->bs;
  for (int k = A->rowptr[row]; k < A->rowptr[row + 1]; k++)
    if (A->colidx[k] == col) {
      A->val[(size_t)k * nt * nt + it * nt + it] += v;
      return;
    }
}

/* aux_lhs: src/flow_simulation.F90:1489-1556.  Al: nrows*nt (zero for rows the loop does not visit) */
void wo_tracer_cell_balances(wo_flow *f, double *Al) {
  const wo_mesh *m = &f->mesh;
  int nt = f->nt;
  int nrows = m->nowned + (m->ncell - m->ninterior);
  for (size_t i = 0; i < (size_t)nrows * nt; i++) Al[i] = 0.0;
  for (int c = 0; c < m->ncell; c++) {
    if (c >= m->nowned && c < m->ninterior) continue; /* partition ghost */
    double *coefs = Al + (size_t)row_of(m, c) * nt;
    for (int it = 0; it < nt; it++)
      coefs[it] = balance_coef(f, f->rock + 8 * (size_t)c, f->fluid + (size_t)c * f->dof, f->tracers[it].phase);
  }
}

/* aux_rhs: src/flow_simulation.F90:1560-1833.  Ar: pattern of wo_tracer_pattern; br: nrows*nt */
void wo_tracer_cell_inflows(wo_flow *f, wo_bsr *Ar, double *br) {
  const wo_mesh *m = &f->mesh;
  int np = f->np, nf = f->nflux, nt = f->nt;
  const double flux_sign[2] = {-1.0, 1.0};
  int nrows = Ar->nb;
  memset(Ar->val, 0, (size_t)Ar->nnzb * nt * nt * sizeof(double)); /* MatZeroEntries :1603 */
  for (size_t i = 0; i < (size_t)nrows * nt; i++) br[i] = 0.0;

  for (int iface = 0; iface < m->nface; iface++) {
    const int32_t *cells = m->face_cells + 2 * (size_t)iface;
    const double *g = m->face_geom + 12 * (size_t)iface;
    const double *phase_flux = f->flux + (size_t)iface * nf + np;
    double vol[2];
    for (int i = 0; i < 2; i++) vol[i] = m->cell_geom[4 * (size_t)cells[i] + 3];
    for (int it = 0; it < nt; it++) {
      const wo_tracer *tr = &f->tracers[it];
      double tracer_phase_flux = phase_flux[tr->phase - 1];
      int up = (tracer_phase_flux >= 0.0) ? 0 : 1;
      double tracer_flow = tracer_phase_flux * g[0];
      double cell_factor[2];
      for (int i = 0; i < 2; i++)
        cell_factor[i] = cell_diffusion_factor(f, f->rock + 8 * (size_t)cells[i],
                                               f->current_fluid + (size_t)cells[i] * f->dof, tr->phase);
      double diffusion_factor = wo_face_harmonic_average(g, cell_factor);
      for (int i = 0; i < 2; i++) {
        if (cells[i] < m->nowned) { /* ghost_cell < 0 and cell < end_interior_cell */
          int irow = row_of(m, cells[i]);
          double Ft = flux_sign[i] * tracer_flow / vol[i];
          add_value(Ar, irow, row_of(m, cells[up]), it, Ft);
          for (int j = 0; j < 2; j++) {
            Ft = -flux_sign[i] * flux_sign[j] * g[0] * diffusion_factor * tr->diffusion / (g[3] * vol[i]);
            add_value(Ar, irow, row_of(m, cells[j]), it, Ft);
          }
        }
      }
    }
  }

  /* tracer_source_iterator :1717-1771 */
  for (int s = 0; s < f->nsrc; s++) {
    int c = f->src_cell[s];
    if (c < 0 || c >= m->nowned) continue;
    double volume = m->cell_geom[4 * (size_t)c + 3];
    double rate = wo_flow_source_rate(f, s);
    if (wo_flow_source_component(f, s, rate) < np) {
      if (rate < 0.0) {
        double frac[WO_MAX_NP];
        wo_flow_source_phase_fractions(f, s, frac);
        for (int it = 0; it < nt; it++) {
          double q = frac[f->tracers[it].phase - 1] * rate / volume;
          add_value(Ar, c, c, it, q);
        }
      } else if (f->tracer_injection) {
        for (int it = 0; it < nt; it++) br[(size_t)c * nt + it] += f->tracer_injection[(size_t)s * nt + it] / volume;
      }
    }
  }

  /* apply_tracer_decay :1775-1831 (uses the stored fluid, self%fluid) */
  for (int c = 0; c < m->ncell; c++) {
    if (c >= m->nowned && c < m->ninterior) continue;
    const double *fl = f->fluid + (size_t)c * f->dof;
    int row = row_of(m, c);
    for (int it = 0; it < nt; it++) {
      double coef = balance_coef(f, f->rock + 8 * (size_t)c, fl, f->tracers[it].phase);
      double a = -wo_tracer_decay(&f->tracers[it], fl[1]) * coef;
      add_value(Ar, row, row, it, a);
    }
  }
}

/* aux_pre_solve: src/flow_simulation.F90:1837-1959.  x_prev: previous solution (nrows*nt), whose
   boundary rows hold the boundary-condition mass fractions */
void wo_tracer_pre_solve(wo_flow *f, wo_bsr *A, double *b, const double *x_prev) {
  const wo_mesh *m = &f->mesh;
  int nt = f->nt;
  for (int c = 0; c < m->ncell; c++) {
    if (c >= m->nowned && c < m->ninterior) continue;
    int row = row_of(m, c);
    int boundary = (c >= m->ninterior);
    int phases = nint_(f->fluid[(size_t)c * f->dof + 4]);
    for (int it = 0; it < nt; it++) {
      int absent = !(phases & (1 << (f->tracers[it].phase - 1)));
      if (!boundary && !absent) continue;
      /* MatZeroRowsLocal(A, rows, diag = 1) + VecSetValuesLocal(b, rows, mass_fraction) */
      for (int k = A->rowptr[row]; k < A->rowptr[row + 1]; k++) {
        double *blk = A->val + (size_t)k * nt * nt;
        for (int jt = 0; jt < nt; jt++) blk[jt * nt + it] = 0.0; /* row `it` of the column-major block */
        if (A->colidx[k] == row) blk[it * nt + it] = 1.0;
      }
      b[(size_t)row * nt + it] = boundary ? x_prev[(size_t)row * nt + it] : 0.0;
    }
  }
}

/* method%setup_linear (src/timestepper.F90:458-581) followed by aux_pre_solve.
   method 0: backward Euler  A = Al - dt Ar,  b = Al_last x_last + dt br
   method 1: BDF2  A = (1+2r) Al - (r+1) dt Ar,  b = (r+1)^2 Al_last x_last - r^2 Al_last2 x_last2 + dt (r+1) br,
             r = dt / dt_last
   method 2: direct steady state  A = Ar, b = -br
   al (out): the new Al.  x_last's boundary rows give the Dirichlet values. */
void wo_tracer_setup_linear(wo_flow *f, int method, double dt, double dt_last, const double *al_last,
                            const double *x_last, const double *al_last2, const double *x_last2, wo_bsr *A,
                            double *b, double *al) {
  int nt = f->nt;
  size_t n = (size_t)A->nb * nt;
  size_t nv = (size_t)A->nnzb * nt * nt;
  double *br = (double *)malloc((n + 1) * sizeof(double));
  if (method == 2) {
    wo_tracer_cell_inflows(f, A, b);
    for (size_t i = 0; i < n; i++) b[i] = -1.0 * b[i]; /* VecScale :576 */
    if (al) wo_tracer_cell_balances(f, al);
  } else {
    wo_tracer_cell_balances(f, al);
    wo_tracer_cell_inflows(f, A, br);
    if (method == 0) {
      for (size_t i = 0; i < nv; i++) A->val[i] = -dt * A->val[i]; /* MatScale :484 */
      for (int i = 0; i < A->nb; i++)                               /* MatDiagonalSet ADD_VALUES :485 */
        for (int it = 0; it < nt; it++) add_value(A, i, i, it, al[(size_t)i * nt + it]);
      for (size_t i = 0; i < n; i++) b[i] = al_last[i] * x_last[i]; /* VecPointwiseMult :487 */
      for (size_t i = 0; i < n; i++) b[i] = b[i] + dt * br[i];      /* VecAXPY :489 */
    } else {
      double r = dt / dt_last, r1 = r + 1.0;
      double sa = -dt * r1, sd = 1.0 + 2.0 * r, s0 = r1 * r1, s2 = -r * r, sb = dt * r1;
      for (size_t i = 0; i < nv; i++) A->val[i] = sa * A->val[i];   /* MatScale :533 */
      for (int i = 0; i < A->nb; i++)
        for (int it = 0; it < nt; it++) add_value(A, i, i, it, al[(size_t)i * nt + it] * sd); /* :535-538 */
      for (size_t i = 0; i < n; i++) b[i] = al_last[i] * x_last[i];
      for (size_t i = 0; i < n; i++) b[i] = b[i] * s0;              /* VecScale :541 */
      for (size_t i = 0; i < n; i++) b[i] = b[i] + s2 * (al_last2[i] * x_last2[i]); /* :542-544 */
      for (size_t i = 0; i < n; i++) b[i] = b[i] + sb * br[i];      /* :546 */
    }
  }
  free(br);
  wo_tracer_pre_solve(f, A, b, x_last);
}
