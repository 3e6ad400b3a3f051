// wb_state.cuh -- device-side accessors shared by the assembly kernels (wb_flow.cu, wb_tracer.cu):
// the SoA cell state written by k_eos, the SoA face geometry and the sorted source list.
// Host-compilable (tests/hostcheck builds the same source with g++).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "wb_eos.cuh"

// ---------------------------------------------------------------- state SoA

template <int NC, int NPH>
WB_HD void load_state(const double *st, size_t ncell, int c,
                                           WbCellState<NC, NPH> &s) {
  const double *p = st + c;
  s.P = p[0];
  s.T = p[ncell];
  s.cond = p[2 * ncell];
  s.phases = (int)p[3 * ncell];
#pragma unroll
  for (int q = 0; q < NPH; q++) {
    const double *pp = p + (size_t)(4 + q * (5 + (NC > 1 ? NC : 0))) * ncell;
    s.rho[q] = pp[0];
    s.sat[q] = pp[ncell];
    s.pc[q] = pp[2 * ncell];
    s.mob[q] = pp[3 * ncell];
    s.h[q] = pp[4 * ncell];
    if (NC == 1) {
      s.X[q][0] = (s.phases & (1 << q)) ? 1.0 : 0.0;  // single component
    } else {
#pragma unroll
      for (int k = 0; k < NC; k++) s.X[q][k] = pp[(size_t)(5 + k) * ncell];
    }
  }
}

template <int NC, int NPH>
WB_HD void store_state(double *st, size_t ncell, int c,
                                            const WbCellState<NC, NPH> &s) {
  double *p = st + c;
  p[0] = s.P;
  p[ncell] = s.T;
  p[2 * ncell] = s.cond;
  p[3 * ncell] = (double)s.phases;
#pragma unroll
  for (int q = 0; q < NPH; q++) {
    double *pp = p + (size_t)(4 + q * (5 + (NC > 1 ? NC : 0))) * ncell;
    pp[0] = s.rho[q];
    pp[ncell] = s.sat[q];
    pp[2 * ncell] = s.pc[q];
    pp[3 * ncell] = s.mob[q];
    pp[4 * ncell] = s.h[q];
    if (NC > 1) {
#pragma unroll
      for (int k = 0; k < NC; k++) pp[(size_t)(5 + k) * ncell] = s.X[q][k];
    }
  }
}

WB_HD WbFaceGeom load_face(const double *face, size_t nface, int f) {
  WbFaceGeom g;
  const double *p = face + f;
  g.area = p[0];
  g.d1 = p[nface];
  g.d2 = p[2 * nface];
  g.d12 = p[3 * nface];
  g.gravn = p[4 * nface];
  g.k = p[5 * nface];
  return g;
}

// fixed-rate sources sorted by cell (stable: input order inside a cell); head[c] = first source of owned cell c or -1
// component word of a source: injection component | production component << 8 (source%update_flow picks by the
// sign of the current rate, src/source.F90:372-380, 469-476)
WB_HD int wb_source_component(int word, double rate) {
  return rate > 0.0 ? (word & 0xff) : (word >> 8);
}
struct WbSources {
  const int32_t *head, *cell, *comp;
  const double *rate, *enth;
  int n;
  // source controls (null: none), sorted like the sources: ctrl bit 0 = on deliverability, bits 1-2 = direction
  // (0 both, 1 production, 2 injection); productivity index, reference pressure, total-flow limit (<= 0: none)
  const int32_t *ctrl;
  const double *pi, *pref, *limit;
};

// source_network%update for one source (src/source_network.F90:90-292): its rate after the source controls
// (deliverability_source_control_flow_rate src/source_control.F90:359-403 with constant productivity and reference
// pressure and permeability factor 1; direction_source_control_iterator :596-620) and the network controls
// ("total" limiter, src/source_network_node.F90:245-315), evaluated from the state of the source's cell -- at
// every function evaluation, perturbed ones included, so that the finite-difference Jacobian sees it
template <int NC, int NPH>
WB_HD double wb_source_rate(const WbSources &S, int k, const WbCellState<NC, NPH> &s) {
  double rate = S.rate[k];
  if (!S.ctrl) return rate;
  const int ctrl = S.ctrl[k];
  if (ctrl & 1) {
    const double effective_productivity = S.pi[k] * 1.0;
    const double pressure_difference = s.P - S.pref[k];
    rate = 0.0;
#pragma unroll
    for (int p = 0; p < NPH; p++)
      if (s.phases & (1 << p)) rate = rate - effective_productivity * s.mob[p] * pressure_difference;
  }
  const int direction = (ctrl >> 1) & 3;
  if (direction == 1 && !(rate < 0.0)) rate = 0.0;
  if (direction == 2 && !(rate > 0.0)) rate = 0.0;
  const double limit = S.limit[k];
  if (limit > 0.0) {
    const double abs_rate = fabs(rate);
    if (abs_rate > limit) {
      double scale = 1.0;
      if (abs_rate > 1.e-6) scale = fmin(scale, limit / abs_rate);
      rate = rate * scale;
    }
  }
  return rate;
}

