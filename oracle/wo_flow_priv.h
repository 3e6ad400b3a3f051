/*
 * wo_flow_priv.h -- oracle (TEST INFRASTRUCTURE): the flow-simulation object shared by
 * wo_flow.c and wo_tracer.c (the fields of flow_simulation_type this path uses,
 * src/flow_simulation.F90:43-130).
 */
#ifndef WO_FLOW_PRIV_H
#define WO_FLOW_PRIV_H
#include "oracle.h"

struct wo_flow {
  wo_params prm;
  wo_mesh mesh;
  wo_eos *eos;
  int np, nc, nphase, nmobile, dof, nflux, isothermal;
  double *fluid, *current_fluid, *last_iteration_fluid, *last_timestep_fluid;
  double *balances; /* nowned*np: last unperturbed lhs */
  double *flux;     /* nface*nflux */
  double *update;   /* ncell: +1 / -1 */
  double *rock;     /* private copy so boundary ghost rock can be set */
  int unperturbed;
  /* OpenMP owner-computes plan of the face loop: cell chunk bounds and per-chunk face lists (wo_flow.c) */
  int eos_nthr;      /* per-thread EOS instances (the EOS carries mutable power-table scratch) */
  wo_eos **eos_thr;
  int plan_nthr;
  int *plan_c0, *plan_nfaces;
  int32_t **plan_faces;
  /* time-stepping method: 0 backward Euler, 1 BDF2, 2 direct steady state (timestepper.F90:345-452) */
  int method;
  double dt_last;
  double *lhs_last2;
  /* fixed-rate sources / sinks (src/source.F90:375-480), in input order */
  int nsrc;
  int32_t *src_cell, *src_component;
  int32_t *src_pcomponent; /* production component (src_component is the injection component) */
  double *src_rate, *src_enthalpy;
  /* source controls (src/source_control.F90): deliverability (:322-507), direction (:596-620), total limiter
     (src/source_network_node.F90:245-315); all NULL when no control is set */
  int32_t *src_ctrl, *src_direction;   /* per source: 1 = on deliverability; 0 both / 1 production / 2 injection */
  double *src_pi, *src_pref, *src_limit; /* productivity index, reference pressure, total rate limit (<= 0: none) */
  /* separators (src/separator.F90) and separated-flow limiters: stages per source (0: none), reference water and steam
     enthalpies of up to two stages [4 per source: hw0, hs0, hw1, hs1], limits on the separated water and steam rates */
  int32_t *src_sep_n;
  double *src_sep_h, *src_limit_water, *src_limit_steam;
  /* reference pressure of a source on deliverability tabulated against the flowing enthalpy or the pressure
     (deliverability.pressure: {"enthalpy": [[h, P], ...]} / {"pressure": ...}, src/source_control.F90:376-388): points per
     source (0: none), coordinate (0 enthalpy, 1 pressure), step interpolation flag, WO_PTAB_MAX (x, y) pairs per source */
  int32_t *src_ptab_n, *src_ptab_coord, *src_ptab_step;
  double *src_ptab;
  double *src_rate_eval;               /* rate every source had at the last unperturbed evaluation */
  /* passive tracers (src/tracer.F90:25-41): auxiliary linear problem, wo_tracer.c */
  int nt;
  wo_tracer *tracers;
  double *tracer_injection; /* nsrc*nt: source%tracer_injection_rate */
};

/* source%fluid%phase_flow_fractions (src/fluid.F90:374-398) of the current fluid of source s's cell */
void wo_flow_source_phase_fractions(const wo_flow *f, int s, double *frac);
/* rate of source s after its controls, evaluated from the current fluid of its cell */
double wo_flow_source_rate(const wo_flow *f, int s);
/* injection or production component of source s for a rate of this sign */
int wo_flow_source_component(const wo_flow *f, int s, double rate);

#endif
