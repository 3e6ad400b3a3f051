// wb_newton.cu -- one SNES newtonls solve of a backward-Euler step, host control
// flow in C++ replaying the callbacks Waiwera registers with PETSc
// (src/timestepper.F90:1552-1641): residual (:587), SNESSetUpdate hook (:628),
// Jacobian (:1584-1611), KSPSolve, shell line search with post-check ->
// fluid_transitions (:649-735), convergence test (:1898-1951).  All vectors stay
// on the GPU; the host only sees a handful of scalars per iteration.
#include <math.h>

#include "wb_common.cuh"

// SNESConvergedReason values used
#define SNES_CONVERGED_FNORM_ABS 2
#define SNES_CONVERGED_FNORM_RELATIVE 3
#define SNES_CONVERGED_SNORM_RELATIVE 4
#define SNES_DIVERGED_FUNCTION_DOMAIN (-1)
#define SNES_DIVERGED_LINEAR_SOLVE (-3)
#define SNES_DIVERGED_FNORM_NAN (-4)
#define SNES_DIVERGED_MAX_IT (-5)
#define SNES_DIVERGED_LINE_SEARCH (-6)
#define SNES_DIVERGED_DTOL (-9)

struct NewtonWork {
  size_t n = 0;
  double *F = nullptr, *Y = nullptr, *W = nullptr, *y = nullptr, *lhs_last = nullptr;
  wb_pc *pc = nullptr;
  int pc_type = -1, pc_nblocks = -1;
};
static std::map<wb_ctx *, NewtonWork> g_newton;

void wb_newton_release(wb_ctx *c) {
  std::lock_guard<std::mutex> lk(wb_registry_mutex());
  auto it = g_newton.find(c);
  if (it == g_newton.end()) return;
  NewtonWork &w = it->second;
  if (w.pc) wb_pc_destroy(w.pc);
  cudaFree(w.F);
  g_newton.erase(it);
}

static int ensure(wb_ctx *c, size_t n, NewtonWork **out) {
  NewtonWork *wq;
  {
    std::lock_guard<std::mutex> lk(wb_registry_mutex());
    wq = &g_newton[c];
  }
  NewtonWork &w = *wq;
  if (w.n != n) {
    if (w.pc) wb_pc_destroy(w.pc);
    cudaFree(w.F);
    w = NewtonWork();
    w.n = n;
    WB_CUDA(cudaMalloc(&w.F, sizeof(double) * n * 5));
    w.Y = w.F + n;
    w.W = w.Y + n;
    w.y = w.W + n;
    w.lhs_last = w.y + n;
  }
  *out = &w;
  return 0;
}

// the PC's pattern belongs to the mesh: drop it when the mesh changes
void wb_newton_invalidate_pc(wb_ctx *c) {
  std::lock_guard<std::mutex> lk(wb_registry_mutex());
  auto it = g_newton.find(c);
  if (it == g_newton.end()) return;
  if (it->second.pc) wb_pc_destroy(it->second.pc);
  it->second.pc = nullptr;
  it->second.pc_type = -1;
}

extern "C" int wb_set_pc_blocks(wb_ctx *c, const int32_t *block_of_row) {
  if (block_of_row) {
    WB_CHECK(c->nowned > 0, "wb_set_pc_blocks: no mesh");
    c->pc_blocks.assign(block_of_row, block_of_row + c->nowned);
  } else {
    c->pc_blocks.clear();
  }
  wb_newton_invalidate_pc(c);
  return 0;
}

struct SnesState {
  double ttol, rnorm0;
};

// SNES_convergence (src/timestepper.F90:1898-1951) on top of SNESConvergedDefault
static int converged(wb_ctx *c, const wb_newton_opts *o, SnesState *st, int it, double xnorm, double snorm,
                     double fnorm, const double *d_F, const double *d_lhs_last, const double *d_update,
                     const double *d_solution, int n, double *max_residual, int *reason_out) {
  const double snes_rtol = 1.e-8, snes_abstol = 1.e-50, snes_stol = 1.e-99, snes_divtol = 1.e8;
  int reason = 0;
  int64_t loc;
  WB_TRY(wb_max_scaled_core(c, d_F, d_lhs_last, o->abs_tol, n, max_residual, &loc));
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
    WB_TRY(wb_max_scaled_core(c, d_update, d_solution, o->update_abs_tol, n, &max_update, &loc));
    if (max_update <= o->update_rel_tol) reason = SNES_CONVERGED_SNORM_RELATIVE;
  }
  *reason_out = reason;
  return 0;
}

static int norm2(wb_ctx *c, const double *d_v, size_t n, double *out) {
  double s;
  WB_TRY(wb_vec_dot_host(c, d_v, d_v, n, &s));
  *out = sqrt(s);
  return 0;
}

static int flags_err(wb_ctx *c, int *err) {
  WB_TRY(wb_reduce_flags(c, 4));
  *err = c->h_flags[0] ? 1 : 0;
  return 0;
}

extern "C" int wb_newton_solve_be(wb_ctx *c, const wb_newton_opts *o, double dt, const double *lhs_last, double *y,
                                  wb_newton_result *res) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(c->ncell > 0, "wb_newton_solve_be: no mesh");
  const size_t n = (size_t)c->nowned * c->np;
  NewtonWork *wp;
  WB_TRY(ensure(c, n, &wp));
  NewtonWork &w = *wp;
  memset(res, 0, sizeof(*res));
  const bool y_dev = wb_is_device_ptr(y), l_dev = wb_is_device_ptr(lhs_last);
  double *d_y = w.y;
  const double *d_ll = l_dev ? lhs_last : w.lhs_last;
  WB_CUDA(cudaMemcpyAsync(d_y, y, sizeof(double) * n, cudaMemcpyDefault, c->stream));
  if (!l_dev) WB_CUDA(cudaMemcpyAsync(w.lhs_last, lhs_last, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  (void)y_dev;

  wb_mat *J = &c->J;
  SnesState st = {0, 0};
  int reason = 0, it = 0, err = 0;
  double fnorm = 0.0, xnorm = 0.0, ynorm = 0.0;
  WB_CUDA(cudaMemsetAsync(w.Y, 0, sizeof(double) * n, c->stream));
  WB_TRY(wb_residual_be_dev(c, d_y, d_ll, dt, true, nullptr, nullptr, w.F));
  WB_TRY(flags_err(c, &err));
  if (err) {
    reason = SNES_DIVERGED_FUNCTION_DOMAIN;
    goto done;
  }
  WB_TRY(norm2(c, w.F, n, &fnorm));
  WB_TRY(converged(c, o, &st, 0, 0.0, 0.0, fnorm, w.F, d_ll, w.Y, d_y, (int)n, &res->max_residual[0], &reason));
  while (!reason && it < o->max_iterations) {
    WB_TRY(wb_pre_iteration(c));  // SNESSetUpdate hook
    WB_TRY(wb_jacobian_be_dev(c, d_y, d_ll, dt, o->fd_err, o->fd_umin, true));
    WB_TRY(flags_err(c, &err));
    if (err) {
      reason = SNES_DIVERGED_FUNCTION_DOMAIN;
      break;
    }
    // PCSetUp: symbolic once per mesh, numeric every Newton iteration
    int prc;
    if (!w.pc || w.pc_type != o->pc_type || w.pc_nblocks != o->pc_nblocks) {
      if (w.pc) wb_pc_destroy(w.pc);
      w.pc = nullptr;
      prc = wb_pc_setup(J, o->pc_type, o->pc_nblocks, c->pc_blocks.empty() ? nullptr : c->pc_blocks.data(), &w.pc);
      w.pc_type = o->pc_type;
      w.pc_nblocks = o->pc_nblocks;
    } else {
      prc = wb_pc_refactor(w.pc);
    }
    if (prc < 0) return prc;
    if (prc > 0) {
      reason = SNES_DIVERGED_LINEAR_SOLVE;
      break;
    }
    int lits = 0, kreason = 0;
    double lres = 0.0;
    WB_TRY(wb_ksp_solve_dev(J, w.pc, &o->ksp, w.F, w.Y, &lits, &kreason, &lres));
    res->lin_its[it < 32 ? it : 31] = lits;
    res->lin_reason[it < 32 ? it : 31] = kreason;
    res->lin_rnorm[it < 32 ? it : 31] = lres;
    res->linear_iterations += lits;
    if (kreason < 0) {
      reason = SNES_DIVERGED_LINEAR_SOLVE;
      break;
    }
    // shell line search (src/timestepper.F90:673-735), lambda = 1: w = y - search
    WB_TRY(wb_vec_axpby_dev(c, w.W, -1.0, w.Y, 1.0, d_y, n));
    WB_TRY(wb_fluid_transitions_dev(c, d_y, w.Y, w.W));
    WB_TRY(wb_reduce_flags(c, 4));
    if (c->h_flags[0]) {
      reason = SNES_DIVERGED_FUNCTION_DOMAIN;
      break;
    }
    if (c->h_flags[1] && !c->h_flags[2]) WB_TRY(wb_vec_axpby_dev(c, w.W, -1.0, w.Y, 1.0, d_y, n));
    WB_CUDA(cudaMemcpyAsync(d_y, w.W, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    // the line search evaluates F at the new iterate in every iteration, the last one included
    // (SNESLineSearchApply -> SNESComputeFunction): the norms below are those of the new state
    WB_TRY(wb_residual_be_dev(c, d_y, d_ll, dt, true, nullptr, nullptr, w.F));
    WB_TRY(flags_err(c, &err));
    if (err) {
      reason = SNES_DIVERGED_LINE_SEARCH;
      break;
    }
    WB_TRY(norm2(c, w.F, n, &fnorm));
    WB_TRY(norm2(c, d_y, n, &xnorm));
    WB_TRY(norm2(c, w.Y, n, &ynorm));
    it++;
    WB_TRY(converged(c, o, &st, it, xnorm, ynorm, fnorm, w.F, d_ll, w.Y, d_y, (int)n,
                     &res->max_residual[it < 32 ? it : 31], &reason));
  }
  if (!reason && it >= o->max_iterations) reason = SNES_DIVERGED_MAX_IT;
done:
  res->reason = reason;
  res->iterations = it;
  WB_CUDA(cudaMemcpyAsync(y, d_y, sizeof(double) * n, cudaMemcpyDefault, c->stream));
  WB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
