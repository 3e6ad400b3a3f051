/*
 * wo_newton.c -- oracle (TEST INFRASTRUCTURE): one SNES newtonls solve of a
 * backward-Euler time step, replaying the callback order Waiwera registers
 * with PETSc (src/timestepper.F90:1552-1641) -- residual (:587), update hook
 * (:628), FD-coloured Jacobian (:1584-1611), KSP solve, shell line search
 * (:673-735) with post-check -> fluid_transitions (:649), convergence test
 * (:1898-1951).  PETSc's SNESSolve_NEWTONLS / SNESConvergedDefault control
 * flow is restated from PETSc 3.22 documentation (PARITY UNPINNED).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* SNESConvergedReason values used */
#define SNES_CONVERGED_FNORM_ABS 2
#define SNES_CONVERGED_FNORM_RELATIVE 3
#define SNES_CONVERGED_SNORM_RELATIVE 4
#define SNES_DIVERGED_FUNCTION_DOMAIN (-1)
#define SNES_DIVERGED_LINEAR_SOLVE (-3)
#define SNES_DIVERGED_FNORM_NAN (-4)
#define SNES_DIVERGED_MAX_IT (-5)
#define SNES_DIVERGED_LINE_SEARCH (-6)
#define SNES_DIVERGED_DTOL (-9)

static double norm2(const double *v, size_t n) {
  double s = 0.0;
  for (size_t i = 0; i < n; i++) s += v[i] * v[i];
  return sqrt(s);
}

typedef struct {
  double ttol, rnorm0;
} snes_state;

/* SNES_convergence: timestepper.F90:1898-1951 over SNESConvergedDefault */
static int converged(const wo_newton_opts *o, snes_state *st, int it, double xnorm, double snorm, double fnorm,
                     const double *F, const double *lhs_last, const double *update, const double *solution,
                     size_t n, double *max_residual) {
  const double snes_rtol = 1.e-8, snes_abstol = 1.e-50, snes_stol = 1.e-99, snes_divtol = 1.e8;
  int reason = 0, loc;
  wo_vec_max_pointwise_abs_scale(F, lhs_last, o->abs_tol, (int)n, max_residual, &loc);
  if (!it) {
    st->ttol = fnorm * snes_rtol;
    st->rnorm0 = fnorm;
  }
  if (fnorm != fnorm || isinf(fnorm)) reason = SNES_DIVERGED_FNORM_NAN;
  else if (fnorm < snes_abstol) reason = SNES_CONVERGED_FNORM_ABS;
  if (it && !reason) {
    if (fnorm <= st->ttol) reason = SNES_CONVERGED_FNORM_RELATIVE;
    else if (snorm < snes_stol * xnorm) reason = SNES_CONVERGED_SNORM_RELATIVE;
    else if (fnorm > snes_divtol * st->rnorm0) reason = SNES_DIVERGED_DTOL;
  }
  if (it < o->min_iterations) {
    reason = 0;
  } else if (*max_residual < o->rel_tol) {
    reason = SNES_CONVERGED_FNORM_RELATIVE;
  } else if (it > 0) {
    double max_update;
    wo_vec_max_pointwise_abs_scale(update, solution, o->update_abs_tol, (int)n, &max_update, &loc);
    if (max_update <= o->update_rel_tol) reason = SNES_CONVERGED_SNORM_RELATIVE;
  }
  return reason;
}

int wo_newton_solve_be(wo_flow *f, wo_bsr *J, const int32_t *color, int ncolor, const int32_t *block_of_row,
                       const wo_newton_opts *o, double dt, const double *lhs_last, double *y,
                       wo_newton_result *res) {
  size_t n = (size_t)J->nb * J->bs;
  double *F = (double *)malloc(n * sizeof(double)), *Y = (double *)calloc(n, sizeof(double));
  double *W = (double *)malloc(n * sizeof(double));
  double *lhs = (double *)malloc(n * sizeof(double)), *rhs = (double *)malloc(n * sizeof(double));
  snes_state st = {0, 0};
  memset(res, 0, sizeof(*res));
  int reason = 0, it = 0;
  int err = wo_residual_be(f, y, lhs_last, dt, NULL, 0, lhs, rhs, F);
  if (err) {
    reason = SNES_DIVERGED_FUNCTION_DOMAIN;
    goto done;
  }
  double fnorm = norm2(F, n), xnorm = 0.0, ynorm = 0.0;
  reason = converged(o, &st, 0, 0.0, 0.0, fnorm, F, lhs_last, Y, y, n, &res->max_residual[0]);
  while (!reason && it < o->max_iterations) {
    wo_flow_pre_iteration(f); /* SNESSetUpdate hook */
    err = wo_fd_jacobian(f, y, lhs_last, dt, F, color, ncolor, o->fd_err, o->fd_umin, J);
    if (err) {
      reason = SNES_DIVERGED_FUNCTION_DOMAIN;
      break;
    }
    wo_pc *pc = wo_pc_create(J, o->pc_type, block_of_row);
    if (!pc) {
      reason = SNES_DIVERGED_LINEAR_SOLVE;
      break;
    }
    int lits = 0;
    double lres = 0.0;
    int kreason = wo_ksp_solve(J, pc, &o->ksp, F, Y, &lits, &lres);
    wo_pc_destroy(pc);
    res->lin_its[it < 32 ? it : 31] = lits;
    res->lin_reason[it < 32 ? it : 31] = kreason;
    res->lin_rnorm[it < 32 ? it : 31] = lres;
    res->linear_iterations += lits;
    if (kreason < 0) {
      reason = SNES_DIVERGED_LINEAR_SOLVE;
      break;
    }
    /* shell line search: timestepper.F90:673-735, lambda = 1 */
#pragma omp parallel for schedule(static) if (n >= 20000)
    for (size_t i = 0; i < n; i++) W[i] = -1.0 * Y[i] + y[i]; /* VecWAXPY(w,-lambda,y,x) */
    int changed_search = 0, changed_w = 0;
    err = wo_flow_fluid_transitions(f, y, Y, W, &changed_search, &changed_w);
    if (err) {
      reason = SNES_DIVERGED_FUNCTION_DOMAIN;
      break;
    }
    if (changed_search && !changed_w)
      for (size_t i = 0; i < n; i++) W[i] = -1.0 * Y[i] + y[i];
    memcpy(y, W, n * sizeof(double));
    /* the line search evaluates F at the new iterate in every iteration, the last one included */
    err = wo_residual_be(f, y, lhs_last, dt, NULL, 0, lhs, rhs, F);
    if (err) {
      reason = SNES_DIVERGED_LINE_SEARCH; /* line search failed: function domain */
      break;
    }
    fnorm = norm2(F, n);
    xnorm = norm2(y, n);
    ynorm = norm2(Y, n);
    it++;
    reason = converged(o, &st, it, xnorm, ynorm, fnorm, F, lhs_last, Y, y, n, &res->max_residual[it < 32 ? it : 31]);
  }
  if (!reason && it >= o->max_iterations) reason = SNES_DIVERGED_MAX_IT;
done:
  res->reason = reason;
  res->iterations = it;
  free(F);
  free(Y);
  free(W);
  free(lhs);
  free(rhs);
  return reason;
}
