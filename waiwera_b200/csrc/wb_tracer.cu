// wb_tracer.cu -- the passive-tracer auxiliary linear problem:
// after a converged Newton solve the reference assembles ONE linear system for the tracer mass fractions
// (src/timestepper.F90:2347-2353, setup_linear :458-581) from aux_lhs (src/flow_simulation.F90:1489-1556),
// aux_rhs (:1560-1833) and aux_pre_solve (:1837-1959) and hands it to KSP.  Here one kernel builds the whole
// system -- balance coefficients Al, the advection / diffusion / production / decay matrix Ar, the method's
// scaling, the right-hand side and the pre_solve row fixes -- row by row on the Jacobian's block pattern
// (BAIJ, bs = number of tracers, diagonal blocks because tracers do not couple), and the Krylov kernels of
// wb_linalg.cu solve it.  Compiled without FMA contraction like wb_flow.cu so that the entries round as the
// reference's operation order does.
#include <algorithm>

#include "wb_common.cuh"

#include "wb_tracer.cuh"

// ================================================================ host side

static WbTracerDev tracer_dev(const wb_ctx *c) {
  WbTracerDev t;
  for (int k = 0; k < WB_MAX_TRACERS; k++) {
    t.phase[k] = std::max(c->trc_phase[k] - 1, 0);
    t.diffusion[k] = c->trc_diffusion[k];
    t.decay[k] = c->trc_decay[k];
    t.activation[k] = c->trc_activation[k];
  }
  return t;
}

void wb_tracer_release(wb_ctx *c) {
  if (c->trc_pc) wb_pc_destroy(c->trc_pc);
  c->trc_pc = nullptr;
  c->trc_pc_type = -1;
  if (c->A_aux) {
    // rowptr / colidx are the Jacobian's
    cudaFree(c->A_aux->d_val);
    cudaFree(c->A_aux->d_xloc);
    cudaFree(c->A_aux->d_tile_e0);
    wb_sell_free(c->A_aux);
    delete c->A_aux;
  }
  c->A_aux = nullptr;
  cudaFree(c->d_trc_b); cudaFree(c->d_trc_x); cudaFree(c->d_trc_al);
  c->d_trc_b = c->d_trc_x = c->d_trc_al = nullptr;
}

extern "C" int wb_set_tracers(wb_ctx *c, int nt, const int32_t *phase, const double *diffusion, const double *decay,
                              const double *activation) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(nt >= 0 && nt <= WB_MAX_TRACERS, "wb_set_tracers: %d tracers (at most %d)", nt, WB_MAX_TRACERS);
  WB_CHECK(nt == 0 || c->ncell > 0, "wb_set_tracers: no mesh");
  WB_CUDA(cudaStreamSynchronize(c->stream));
  wb_tracer_release(c);
  cudaFree(c->d_trc_inj);
  c->d_trc_inj = nullptr;
  c->nt = nt;
  for (int k = 0; k < nt; k++) {
    WB_CHECK(phase[k] >= 1 && phase[k] <= c->nph, "wb_set_tracers: tracer %d: phase %d out of range", k, phase[k]);
    c->trc_phase[k] = phase[k];
    c->trc_diffusion[k] = diffusion ? diffusion[k] : 0.0;
    c->trc_decay[k] = decay ? decay[k] : 0.0;
    c->trc_activation[k] = activation ? activation[k] : 0.0;
  }
  if (nt == 0) return 0;
  // A_aux: DMCreateMatrix on the tracer DM (src/ode.F90:301-321) -- the Jacobian's block pattern with bs = nt
  const wb_mat &J = c->J;
  wb_mat *A = new wb_mat();
  A->ctx = c; A->nb = J.nb; A->ncolb = J.ncolb; A->bs = nt; A->nnzb = J.nnzb; A->owns = false;
  A->d_rowptr = J.d_rowptr; A->d_colidx = J.d_colidx;
  A->h_rowptr = J.h_rowptr; A->h_colidx = J.h_colidx;
  c->A_aux = A;
  const size_t nv = (size_t)J.nnzb * nt * nt;
  WB_CUDA(cudaMalloc(&A->d_val, sizeof(double) * std::max<size_t>(nv, 1) + WB_PAD_BYTES));
  WB_CUDA(wb_memset_sync(A->d_val, 0, sizeof(double) * std::max<size_t>(nv, 1) + WB_PAD_BYTES));
  WB_CUDA(cudaMalloc(&A->d_xloc, sizeof(double) * (size_t)(J.ncolb - J.nb + 1) * nt));
  WB_CUDA(wb_memset_sync(A->d_xloc, 0, sizeof(double) * (size_t)(J.ncolb - J.nb + 1) * nt));
  WB_TRY(wb_mat_build_tiles(A));
  const size_t n = (size_t)c->nowned * nt;
  WB_CUDA(cudaMalloc(&c->d_trc_b, sizeof(double) * std::max<size_t>(n, 1)));
  WB_CUDA(cudaMalloc(&c->d_trc_x, sizeof(double) * std::max<size_t>(n, 1)));
  WB_CUDA(cudaMalloc(&c->d_trc_al, sizeof(double) * std::max<size_t>(n, 1)));
  return 0;
}

extern "C" int wb_set_tracer_injection(wb_ctx *c, const double *rate) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(c->d_trc_inj);
  c->d_trc_inj = nullptr;
  if (!rate || c->nsrc == 0 || c->nt == 0) return 0;
  const int nt = c->nt;
  std::vector<double> host((size_t)c->nsrc * nt), sorted((size_t)c->nsrc * nt);
  WB_CUDA(wb_memcpy_sync(host.data(), rate, sizeof(double) * host.size(), cudaMemcpyDefault));
  for (int k = 0; k < c->nsrc; k++)
    for (int t = 0; t < nt; t++) sorted[(size_t)k * nt + t] = host[(size_t)c->h_src_order[k] * nt + t];
  WB_CUDA(cudaMalloc(&c->d_trc_inj, sizeof(double) * sorted.size()));
  WB_CUDA(wb_memcpy_sync(c->d_trc_inj, sorted.data(), sizeof(double) * sorted.size(), cudaMemcpyHostToDevice));
  return 0;
}

#define DISPATCH_EOS_NT(ctx, CALL)                                               \
  do {                                                                           \
    const int nt_ = (ctx)->nt;                                                   \
    if ((ctx)->eos.eos == WB_EOS_WE) {                                           \
      if (nt_ == 1) { CALL(WB_EOS_WE, 1); } else if (nt_ == 2) { CALL(WB_EOS_WE, 2); } else { CALL(WB_EOS_WE, 3); }    \
    } else if ((ctx)->eos.eos == WB_EOS_WCE) {                                   \
      if (nt_ == 1) { CALL(WB_EOS_WCE, 1); } else if (nt_ == 2) { CALL(WB_EOS_WCE, 2); } else { CALL(WB_EOS_WCE, 3); } \
    } else {                                                                     \
      if (nt_ == 1) { CALL(WB_EOS_W, 1); } else if (nt_ == 2) { CALL(WB_EOS_W, 2); } else { CALL(WB_EOS_W, 3); }       \
    }                                                                            \
  } while (0)

// device-pointer core: d_b / d_al may be null; assemble = false computes the balance coefficients only
static int tracer_assemble_dev(wb_ctx *c, bool assemble, double dt, const double *d_al_last, const double *d_x_last,
                               const double *d_al_last2, const double *d_x_last2, const double *d_xb, double *d_al,
                               double *d_b) {
  WB_CHECK(c->nt > 0 && c->A_aux, "tracers: call wb_set_tracers first");
  TracerArgs a;
  a.state = c->d_state;  // slot 0: the last unperturbed evaluation
  a.face = c->d_face; a.vol = c->d_vol; a.rockp = c->d_rockp;
  a.cf_ptr = c->d_cf_ptr; a.cf_face = c->d_cf_face; a.cf_other = c->d_cf_other; a.cf_bpos = c->d_cf_bpos;
  a.diagpos = c->d_diagpos; a.rowptr = c->J.d_rowptr;
  a.src = wb_sources_args(c);
  a.inj = c->d_trc_inj;
  a.trc = tracer_dev(c);
  a.method = c->method;
  if (c->method == WB_METHOD_BDF2) {
    WB_CHECK(!assemble || (d_al_last2 && d_x_last2 && c->dt_last > 0.0), "tracers: BDF2 needs al_last2, x_last2 and dt_last");
    const double r = dt / c->dt_last, r1 = r + 1.0;
    a.sA = -dt * r1; a.sD = 1.0 + 2.0 * r; a.s0 = r1 * r1; a.s2 = -r * r; a.sb = dt * r1;
  } else if (c->method == WB_METHOD_DIRECTSS) {
    a.sA = 1.0; a.sD = 0.0; a.s0 = 0.0; a.s2 = 0.0; a.sb = 0.0;
  } else {
    a.sA = -dt; a.sD = 1.0; a.s0 = 1.0; a.s2 = 0.0; a.sb = dt;
  }
  WB_CHECK(!assemble || c->method == WB_METHOD_DIRECTSS || (d_al_last && d_x_last), "tracers: al_last and x_last are required");
  a.al_last = d_al_last; a.x_last = d_x_last; a.al_last2 = d_al_last2; a.x_last2 = d_x_last2;
  a.xb = d_xb;
  a.val = assemble ? c->A_aux->d_val : nullptr;
  if (assemble) c->A_aux->version++;
  a.b = d_b; a.al = d_al;
  a.ncell = c->ncell; a.ninterior = c->ninterior; a.nowned = c->nowned; a.nface = c->nface;
  const int grid = wb_grid(c->nowned, 128);
#define CALL(E, T) k_tracer_assemble<E, T><<<grid, 128, 0, c->stream>>>(a)
  DISPATCH_EOS_NT(c, CALL);
#undef CALL
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int wb_tracer_cell_balances(wb_ctx *c, double *al) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(c->nt > 0, "wb_tracer_cell_balances: no tracers");
  int rc = 0;
  WbStage st(c);
  double *d_al = st.out(al, (size_t)c->nowned * c->nt, &rc);
  if (rc) return rc;
  WB_TRY(tracer_assemble_dev(c, false, 0.0, nullptr, nullptr, nullptr, nullptr, nullptr, d_al, nullptr));
  return st.finish();
}

extern "C" int wb_tracer_setup_linear(wb_ctx *c, double dt, const double *al_last, const double *x_last,
                                      const double *al_last2, const double *x_last2, const double *x_boundary,
                                      double *al, double *b, wb_mat **A) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(c->nt > 0, "wb_tracer_setup_linear: no tracers");
  const size_t n = (size_t)c->nowned * c->nt, nbdy = (size_t)(c->ncell - c->ninterior) * c->nt;
  int rc = 0;
  WbStage st(c);
  const double *d_al0 = st.in(al_last, n, &rc), *d_x0 = st.in(x_last, n, &rc);
  const double *d_al2 = st.in(al_last2, n, &rc), *d_x2 = st.in(x_last2, n, &rc);
  const double *d_xb = nbdy ? st.in(x_boundary, nbdy, &rc) : nullptr;
  double *d_al = st.out(al, n, &rc), *d_b = st.out(b, n, &rc);
  if (rc) return rc;
  {
    WbScopedTimer tm(c, "tracer_setup");
    WB_TRY(tracer_assemble_dev(c, true, dt, d_al0, d_x0, d_al2, d_x2, d_xb, d_al, d_b ? d_b : c->d_trc_b));
  }
  if (A) *A = c->A_aux;
  return st.finish();
}

extern "C" int wb_tracer_solve(wb_ctx *c, const wb_ksp_opts *ksp, int pc_type, int pc_nblocks, double dt,
                               const double *al_last, const double *x_last, const double *al_last2,
                               const double *x_last2, const double *x_boundary, double *al, double *x, int *its,
                               int *reason) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(c->nt > 0, "wb_tracer_solve: no tracers");
  const size_t n = (size_t)c->nowned * c->nt, nbdy = (size_t)(c->ncell - c->ninterior) * c->nt;
  int rc = 0;
  WbStage st(c);
  const double *d_al0 = st.in(al_last, n, &rc), *d_x0 = st.in(x_last, n, &rc);
  const double *d_al2 = st.in(al_last2, n, &rc), *d_x2 = st.in(x_last2, n, &rc);
  const double *d_xb = nbdy ? st.in(x_boundary, nbdy, &rc) : nullptr;
  double *d_al = st.out(al, n, &rc), *d_x = st.out(x, n, &rc);
  if (rc) return rc;
  {
    WbScopedTimer tm(c, "tracer_setup");
    WB_TRY(tracer_assemble_dev(c, true, dt, d_al0, d_x0, d_al2, d_x2, d_xb, d_al, c->d_trc_b));
  }
  // PCSetUp: symbolic once per pattern, numeric every step
  int prc;
  if (!c->trc_pc || c->trc_pc_type != pc_type || c->trc_pc_nblocks != pc_nblocks) {
    if (c->trc_pc) wb_pc_destroy(c->trc_pc);
    c->trc_pc = nullptr;
    prc = wb_pc_setup(c->A_aux, pc_type, pc_nblocks, c->pc_blocks.empty() ? nullptr : c->pc_blocks.data(), &c->trc_pc);
    c->trc_pc_type = pc_type;
    c->trc_pc_nblocks = pc_nblocks;
  } else {
    prc = wb_pc_refactor(c->trc_pc);
  }
  if (prc < 0) return prc;
  int its_ = 0, reason_ = 0;
  double rn = 0.0;
  if (prc > 0) {
    reason_ = -11;  // KSP_DIVERGED_PC_FAILED
  } else {
    WbScopedTimer tm(c, "tracer_solve");
    WB_TRY(wb_ksp_solve_dev(c->A_aux, c->trc_pc, ksp, c->d_trc_b, d_x, &its_, &reason_, &rn));
  }
  if (its) *its = its_;
  if (reason) *reason = reason_;
  return st.finish();
}
