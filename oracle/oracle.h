/*
 * oracle.h -- CPU restatement of the Waiwera Newton-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * the CPU baseline.  The product (waiwera_b200/) never links or calls it.
 *
 * Every function cites the reference file:line (relative to the Waiwera
 * source tree, v1.5.1) whose arithmetic and operation order it follows.
 *
 * Parity pinning: the physics (thermodynamics, EOS, curves, face flux, cell
 * balance, transitions) is pinned by the reference's own known-answer unit
 * tests, transcribed in tests/test_oracle_*.py.  The PETSc-side operators
 * (FD-coloured Jacobian, BAIJ SpMV, ILU(0), GMRES/BCGS, SNES newtonls) have
 * no reference test vectors and PETSc is not in the reference tree:
 * for those the oracle is "parity unpinned" (written to PETSc 3.22's
 * documented semantics).
 */
#ifndef WAIWERA_ORACLE_H
#define WAIWERA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants (src/thermodynamics.F90:36-40) ---- */
#define WO_RCONST 0.461526e3
#define WO_TC_K 273.15

/* thermodynamic formulation ids */
#define WO_THERMO_IAPWS 0
#define WO_THERMO_IFC67 1

/* ---- power table (src/powertable.F90) ---- */
typedef struct {
  int lower, upper;
  double *power;   /* indexed power[i - lower] */
  int *product;    /* 2 entries per power */
  int *required;
  int nlist;
  int *list;       /* 3 ints per entry: fac1, fac2, prod */
} wo_powertable;

void wo_powertable_init(wo_powertable *t);
void wo_powertable_configure(wo_powertable *t, const int *powers, int n);
void wo_powertable_compute(wo_powertable *t, double val);
double wo_powertable_get(const wo_powertable *t, int i);
void wo_powertable_destroy(wo_powertable *t);
/* test helper: configure with `powers`, compute for val, return power[q] for each query */
void wo_powertable_eval(const int *powers, int n, double val, const int *query, int nq, double *out);

/* ---- thermodynamics (src/IAPWS.F90, src/IFC67.F90) ---- */
typedef struct wo_thermo wo_thermo;
wo_thermo *wo_thermo_create(int id, int extrapolate);
void wo_thermo_destroy(wo_thermo *th);
int wo_thermo_id(const wo_thermo *th);
double wo_thermo_tcritical(const wo_thermo *th);
double wo_thermo_pcritical(const wo_thermo *th);
/* region = 1 (water), 2 (steam), 3 (supercritical; IAPWS only). param = (p|rho, t degC). returns err */
int wo_region_properties(wo_thermo *th, int region, const double param[2], double props[2]);
double wo_region_viscosity(wo_thermo *th, int region, double temperature, double pressure, double density);
int wo_saturation_pressure(const wo_thermo *th, double t, double *p);
int wo_saturation_temperature(const wo_thermo *th, double p, double *t);
int wo_phase_composition(const wo_thermo *th, int region, double pressure, double temperature);
double wo_boundary23_pressure(double t);
double wo_boundary23_temperature(double p);

/* ---- interpolation table, linear (src/interpolation.F90:202-581) ---- */
typedef struct {
  int n, dim, index;
  double *x;   /* n */
  double *val; /* dim * n, val[d + dim*i] */
} wo_table;
void wo_table_init(wo_table *t, const double *x, const double *v, int n, int dim);
void wo_table_destroy(wo_table *t);
void wo_table_find(wo_table *t, double x);
void wo_table_interpolate_at_index(const wo_table *t, double x, double *y);
void wo_table_interpolate(wo_table *t, double x, double *y);
int wo_table_find_component_at_index(const wo_table *t, double yi, int component, double *x);

/* ---- Brent root finder (src/root_finder.F90:127-248) ---- */
typedef double (*wo_root_fn)(double x, void *ctx);
typedef struct {
  double interval[2], root_tolerance, function_tolerance, root;
  int max_iterations, iterations, err;
} wo_root_finder;
void wo_root_finder_init(wo_root_finder *r);
void wo_root_finder_find(wo_root_finder *r, wo_root_fn f, void *ctx);

/* ---- relative permeability / capillary pressure curves ---- */
#define WO_RP_FULLY_MOBILE 0
#define WO_RP_LINEAR 1
#define WO_RP_PICKENS 2
#define WO_RP_COREY 3
#define WO_RP_GRANT 4
#define WO_RP_VAN_GENUCHTEN 5
#define WO_RP_TABLE 6
#define WO_CP_ZERO 0
#define WO_CP_LINEAR 1
#define WO_CP_VAN_GENUCHTEN 2
#define WO_CP_TABLE 3

#define WO_MAX_TABLE 16
typedef struct {
  int type;
  /* linear: liquid limits p[0..1], vapour limits p[2..3];
     pickens: p[0]=power; corey/grant: p[0]=slr,p[1]=ssr;
     van genuchten: p[0]=lambda,p[1]=slr,p[2]=sls,p[3]=sum_unity(0/1),p[4]=ssr;
     table: nl, nv points in tl*/
  double p[8];
  int nl, nv;
  double lx[WO_MAX_TABLE], ly[WO_MAX_TABLE], vx[WO_MAX_TABLE], vy[WO_MAX_TABLE];
} wo_relperm;
typedef struct {
  int type;
  /* linear: p[0..1] saturation limits, p[2]=pressure;
     vG: p[0]=P0,p[1]=lambda,p[2]=slr,p[3]=sls,p[4]=Pmax,p[5]=apply_Pmax;
     table: n points */
  double p[8];
  int n;
  double x[WO_MAX_TABLE], y[WO_MAX_TABLE];
} wo_cappress;
void wo_relperm_values(const wo_relperm *rp, double sl, double out[2]);
double wo_cappress_value(const wo_cappress *cp, double sl, double t);

/* ---- EOS ---- */
#define WO_EOS_WE 0
#define WO_EOS_W 1
#define WO_EOS_WCE 2
#define WO_EOS_WAE 3 /* water + air + energy: eos_wge with the air NCG (src/eos_wae.F90, src/ncg_air_thermodynamics.F90) */
#define WO_EOS_WAE 3

#define WO_MAX_NP 3  /* max primaries */
#define WO_MAX_NC 2  /* max mass components */

typedef struct {
  int eos;                /* WO_EOS_* */
  int thermo;             /* WO_THERMO_* */
  int extrapolate;
  double pressure_scale, temperature_scale; /* eos.primary.scale.* */
  double partial_pressure_scale;            /* eos_wge: <= 0 adaptive Pg/P scaling (the reference default, eos_wge.F90:96-100) */
  double eos_w_temperature;                 /* eos_w fixed temperature */
  wo_relperm relperm;
  wo_cappress cappress;
  double gravity[3];
} wo_params;

typedef struct wo_eos wo_eos;
wo_eos *wo_eos_create(const wo_params *prm);
void wo_eos_destroy(wo_eos *e);
int wo_eos_num_primary(const wo_eos *e);
int wo_eos_num_components(const wo_eos *e);
int wo_eos_num_phases(const wo_eos *e);
int wo_eos_fluid_dof(const wo_eos *e);
wo_thermo *wo_eos_thermo(wo_eos *e);
void wo_eos_scale(const wo_eos *e, const double *primary, int region, double *scaled);
void wo_eos_unscale(const wo_eos *e, const double *scaled, int region, double *primary);
/* fluid: pointer to one fluid record (Appendix A layout); rock: 8-double record */
int wo_eos_bulk_properties(wo_eos *e, const double *primary, double *fluid);
int wo_eos_phase_properties(wo_eos *e, const double *primary, const double *rock, double *fluid);
int wo_eos_transition(wo_eos *e, const double *old_primary, double *primary,
                      const double *old_fluid, double *fluid, int *transition);
int wo_eos_check_primary_variables(const wo_eos *e, const double *fluid, double *primary, int *changed);
double wo_eos_conductivity(const double *rock, const double *fluid, int nc);

/* ---- CO2 non-condensible gas (src/ncg_co2_thermodynamics.F90, src/ncg_thermodynamics.F90) ---- */
void wo_co2_properties(double partial_pressure, double temperature, double props[2]); /* density, enthalpy */
double wo_co2_henrys_constant(double temperature);
double wo_co2_energy_solution(double temperature, double henrys_constant);
int wo_co2_viscosity(double partial_pressure, double temperature, double *viscosity);

/* ---- air non-condensible gas (src/ncg_air_thermodynamics.F90) ---- */
void wo_air_properties(double partial_pressure, double temperature, double props[2]); /* density, enthalpy */
double wo_air_henrys_constant(double temperature, double constituent[2]);
double wo_air_energy_solution(double temperature, const double constituent_henrys_constant[2]);
double wo_air_mixture_viscosity(double water_viscosity, double temperature, double xg, int phase);

/* ---- local cell / face objects (src/cell.F90:114, src/face.F90:443) ---- */
void wo_cell_balance(const double *rock, const double *fluid, int nc, int nphase, int np, double *balance);
/* face_geom: 12 doubles; rock/fluid: records of the two support cells; flux: np + nmobile */
void wo_face_flux(const double *face_geom, const double *rock1, const double *rock2,
                  const double *fluid1, const double *fluid2,
                  int nc, int np, int nphase, int nmobile, int isothermal, double *flux);
void wo_face_calculate_distances(const double *c1, const double *c2, const double *fc,
                                 const double *normal, double dist[2], double *dist12);
double wo_face_harmonic_average(const double *face_geom, const double x[2]);

/* ---- flow simulation (src/flow_simulation.F90) ---- */
typedef struct {
  int ncell;          /* local cells incl. partition + boundary ghosts */
  int ninterior;      /* end_interior_cell: cells [0, ninterior) have dofs or are partition ghosts */
  int nowned;         /* owned cells (ghost_cell<0) are [0,nowned) */
  int nface;
  const int32_t *face_cells;   /* 2*nface, local cell indices */
  const double *face_geom;     /* 12*nface */
  const double *cell_geom;     /* 4*ncell */
  const double *rock;          /* 8*ncell */
} wo_mesh;

/* OpenMP threads used by the oracle's loops (0: leave the runtime's choice); returns the thread count in effect.
   Every result is bit-identical for every thread count (see wo_flow.c). */
int wo_set_num_threads(int n);

typedef struct wo_flow wo_flow;
wo_flow *wo_flow_create(const wo_params *prm, const wo_mesh *mesh);
void wo_flow_destroy(wo_flow *f);
wo_eos *wo_flow_eos(wo_flow *f);
double *wo_flow_fluid(wo_flow *f);          /* ncell * fluid_dof, "fluid" vector (local incl. ghosts) */
double *wo_flow_current_fluid(wo_flow *f);
double *wo_flow_flux(wo_flow *f);           /* nface * (np+nmobile) */
/* set regions + fluid records of cells from primaries (fluid_init, flow_simulation.F90:2171) */
int wo_flow_fluid_init(wo_flow *f, const double *y, const int32_t *region);
/* Dirichlet boundary ghost cell (mesh.F90:1185-1202): rock copied from interior cell, fluid from unscaled primary */
int wo_flow_set_boundary(wo_flow *f, int ghost_cell, int interior_cell, const double *primary, int region);
int wo_flow_set_rock(wo_flow *f, const double *rock);
/* residual form of wo_residual_be: 0 backward Euler (timestepper.F90:345), 1 BDF2 (:378), 2 direct steady state (:431) */
void wo_flow_set_method(wo_flow *f, int method, double dt_last, const double *lhs_last2);
/* fixed-rate sources / sinks (src/source.F90:375-480): local owned cell, component (1-based; 0 = all mass
   components for production; np = heat), rate (< 0 production), injection enthalpy */
void wo_flow_set_sources(wo_flow *f, int n, const int32_t *cell, const int32_t *component, const double *rate,
                         const double *enthalpy);
/* injection / production component of every source; which applies follows the sign of the current rate */
void wo_flow_set_source_components(wo_flow *f, int n, const int32_t *injection, const int32_t *production);
/* source controls for n of those sources (source: index into the wo_flow_set_sources arrays): deliverability
   (src/source_control.F90:322-507; pi <= 0: none), direction (0 both, 1 production, 2 injection; :596-620), total-flow
   limiter (limit <= 0: none; src/source_network_node.F90:245-315); re-evaluated at every function evaluation */
void wo_flow_set_source_controls(wo_flow *f, int n, const int32_t *source, const double *pi, const double *pref,
                                 const int32_t *direction, const double *limit);
/* rate of every source at the last unperturbed evaluation */
void wo_flow_set_source_recharge(wo_flow *f, int n, const int32_t *source, const double *coefficient, const double *pref);
void wo_flow_get_source_rates(const wo_flow *f, double *rate);
/* separators (src/separator.F90) and limiters on the separated water / steam flows (src/source_network_node.F90:245-315) */
int wo_separator_stage(wo_thermo *th, double pressure, double *ref_water_enthalpy, double *ref_steam_enthalpy);
void wo_separate(int nstage, const double *stage_h, double rate, double enthalpy, double out[5]);
int wo_flow_set_source_separators(wo_flow *f, int n, const int32_t *source, const int32_t *nstage, const double *pressure,
                                  const double *limit_water, const double *limit_steam);
/* reference pressure of sources on deliverability tabulated against the flowing enthalpy (coordinate 0) or the pressure (1)
   of the cell: npts[k] <= WO_PTAB_MAX points (x, y) at table[2 WO_PTAB_MAX k ...] (src/source_control.F90:376-388) */
#define WO_PTAB_MAX 8
int wo_flow_set_source_pressure_table(wo_flow *f, int n, const int32_t *source, const int32_t *coordinate, const int32_t *step,
                                      const int32_t *npts, const double *table);
void wo_flow_source_separated(const wo_flow *f, int s, double rate, double out[5]);
/* pre_eval: update mask from perturbed block columns (NULL/0 => unperturbed) + fluid_properties */
int wo_flow_pre_eval(wo_flow *f, const double *y, const int32_t *perturbed, int nperturbed);
int wo_flow_cell_balances(wo_flow *f, double *lhs);
int wo_flow_cell_inflows(wo_flow *f, double *rhs);
void wo_flow_pre_iteration(wo_flow *f);
void wo_flow_pre_timestep(wo_flow *f);
void wo_flow_pre_retry_timestep(wo_flow *f);
int wo_flow_fluid_transitions(wo_flow *f, const double *y_old, double *search, double *y,
                              int *changed_search, int *changed_y);
void wo_flow_get_regions(wo_flow *f, int32_t *region);

/* ---- passive tracers: auxiliary linear problem (wo_tracer.c; src/tracer.F90, src/flow_simulation.F90:1489-1959,
   src/timestepper.F90:458-581) ---- */
typedef struct {
  int phase;          /* 1-based phase index (tracer%phase_index) */
  double diffusion;   /* diffusion coefficient (m2/s) */
  double decay;       /* decay constant (1/s) */
  double activation;  /* activation energy (J/mol) */
} wo_tracer;
void wo_flow_set_tracers(wo_flow *f, int nt, const wo_tracer *tracers);
/* source%tracer_injection_rate, rate[nsources*nt] in the order of wo_flow_set_sources */
void wo_flow_set_tracer_injection(wo_flow *f, const double *rate);
double wo_tracer_decay(const wo_tracer *t, double temperature);

/* BE residual r = L(y) - L_last - dt*R(y)  (src/timestepper.F90:345-374), includes pre_eval */
int wo_residual_be(wo_flow *f, const double *y, const double *lhs_last, double dt,
                   const int32_t *perturbed, int nperturbed, double *lhs, double *rhs, double *r);
/* max_i |v_i| / max(|scale_i|, tol), first argmax (src/dm_utils.F90:644-685) */
void wo_vec_max_pointwise_abs_scale(const double *v, const double *scale, double tol, int n,
                                    double *maxval, int *maxloc);

/* ---- PETSc-side operators (parity unpinned; PETSc 3.22 semantics) ---- */
typedef struct {
  int nb, bs, nnzb;
  int32_t *rowptr;  /* nb+1 */
  int32_t *colidx;  /* nnzb, sorted per row */
  double *val;      /* nnzb*bs*bs, blocks column-major (PETSc BAIJ) */
} wo_bsr;

/* FV adjacency pattern: row i = {i} U face neighbours among owned cells (dm_utils.F90:1041) */
wo_bsr *wo_bsr_from_mesh(const wo_mesh *mesh, int bs);
void wo_bsr_destroy(wo_bsr *A);
void wo_bsr_spmv(const wo_bsr *A, const double *x, double *y);
/* distance-2 greedy colouring of block columns; returns ncolors, colour per block column */
int wo_bsr_coloring(const wo_bsr *A, int32_t *color);
/* MatFDColoringApply with MATMFFD_DS step rule; F0 = residual at y (already computed) */
int wo_fd_jacobian(wo_flow *f, const double *y, const double *lhs_last, double dt,
                   const double *F0, const int32_t *color, int ncolor,
                   double fd_err, double fd_umin, wo_bsr *J);

/* tracer system (bs = nt): rows = owned cells, then boundary ghost cells (identity rows after pre_solve) */
wo_bsr *wo_tracer_pattern(const wo_mesh *mesh, int nt);
void wo_tracer_cell_balances(wo_flow *f, double *Al);
void wo_tracer_cell_inflows(wo_flow *f, wo_bsr *Ar, double *br);
void wo_tracer_pre_solve(wo_flow *f, wo_bsr *A, double *b, const double *x_prev);
/* method 0 backward Euler, 1 BDF2, 2 direct steady state, then aux_pre_solve */
void wo_tracer_setup_linear(wo_flow *f, int method, double dt, double dt_last, const double *al_last,
                            const double *x_last, const double *al_last2, const double *x_last2, wo_bsr *A,
                            double *b, double *al);

typedef struct wo_pc wo_pc;
#define WO_PC_NONE 0
#define WO_PC_PBJACOBI 1
#define WO_PC_BJACOBI_ILU0 2   /* nblocks sub-domains, ILU(0) each (1 block = global ILU(0)) */
#define WO_PC_ASM_ILU0 3       /* restricted additive Schwarz (PCASM default, PC_ASM_RESTRICT), overlap 1, ILU(0) on each
                                  extended sub-domain */
wo_pc *wo_pc_create(const wo_bsr *A, int type, const int32_t *block_of_row /* may be NULL */);
void wo_pc_apply(const wo_pc *pc, const double *r, double *z);
void wo_pc_destroy(wo_pc *pc);

#define WO_KSP_GMRES 0
#define WO_KSP_BCGS 1
typedef struct {
  int type, restart, maxit;
  double rtol, atol, dtol;
} wo_ksp_opts;
/* returns KSPConvergedReason-like code (>0 converged, <0 diverged); zero initial guess */
int wo_ksp_solve(const wo_bsr *A, const wo_pc *pc, const wo_ksp_opts *o, const double *b,
                 double *x, int *its, double *rnorm);

/* one SNES newtonls solve of a BE step following timestepper.F90 callbacks (Appendix C) */
typedef struct {
  int max_iterations, min_iterations;
  double rel_tol, abs_tol, update_rel_tol, update_abs_tol;
  double fd_err, fd_umin;
  int pc_type;
  wo_ksp_opts ksp;
} wo_newton_opts;
typedef struct {
  int reason, iterations, linear_iterations;
  double max_residual[32];
  int lin_its[32];
  int lin_reason[32];
  double lin_rnorm[32];
} wo_newton_result;
int wo_newton_solve_be(wo_flow *f, wo_bsr *J, const int32_t *color, int ncolor,
                       const int32_t *block_of_row,
                       const wo_newton_opts *o, double dt, const double *lhs_last,
                       double *y, wo_newton_result *res);

#ifdef __cplusplus
}
#endif
#endif
