/*
 * wo_flow.c -- oracle (TEST INFRASTRUCTURE): the flow-simulation ODE loops.
 * Restated from src/flow_simulation.F90:1102-1137 (update mask), :1242-1330
 * (cell_balances), :1334-1485 (cell_inflows), :2022-2147 (pre_* hooks),
 * :2171-2287 (fluid_init), :2291-2415 (fluid_properties), :2419-2576
 * (fluid_transitions); src/timestepper.F90:345-374 (BE residual);
 * src/dm_utils.F90:644-685 (max pointwise scaled abs).
 *
 * Local cell numbering: owned cells [0,nowned), partition ghost cells
 * [nowned,ninterior), boundary (Dirichlet) ghost cells [ninterior,ncell).
 * The reference loops visit owned cells in local index order; the face loop
 * visits flux faces in flux_face order (here: face index order).
 */
#include "wo_flow_priv.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Host parallelism (OpenMP): the cell loops are split into contiguous chunks of owned cells, one per thread, and
   every thread visits the faces that touch its chunk in ascending face order and adds only to its own cells --
   "owner computes", the way the reference's MPI ranks each loop over their local faces and own cells
   (flow_simulation.F90:1291-1316, 1410-1458).  Per cell the additions happen in the same order as in the serial
   face loop, so the results are bit-identical for every thread count. */
/* problems smaller than this many cells run serially (thread start-up costs more than the loop) */
#define WO_PAR_MIN 20000
static int wo_nthreads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int wo_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
  return wo_nthreads();
}

static void par_memcpy(void *dst, const void *src, size_t bytes) {
  const size_t chunk = (size_t)1 << 22;
  if (bytes < 2 * chunk) {
    memcpy(dst, src, bytes);
    return;
  }
  const long nchunk = (long)((bytes + chunk - 1) / chunk);
#pragma omp parallel for schedule(static)
  for (long k = 0; k < nchunk; k++) {
    size_t o = (size_t)k * chunk, len = bytes - o < chunk ? bytes - o : chunk;
    memcpy((char *)dst + o, (const char *)src + o, len);
  }
}

/* The EOS object carries mutable scratch (the IAPWS power tables, powertable.F90:261-278), as the reference's does:
   every thread evaluates with its own instance, the way every MPI rank of the reference owns one. */
static void eos_pool_build(wo_flow *f) {
  int T = wo_nthreads();
  if (f->eos_nthr >= T) return;
  f->eos_thr = (wo_eos **)realloc(f->eos_thr, T * sizeof(wo_eos *));
  for (int t = f->eos_nthr; t < T; t++) f->eos_thr[t] = t == 0 ? f->eos : wo_eos_create(&f->prm);
  f->eos_nthr = T;
}
static inline wo_eos *thread_eos(const wo_flow *f) {
#ifdef _OPENMP
  int t = omp_get_thread_num();
  return t < f->eos_nthr ? f->eos_thr[t] : f->eos;
#else
  return f->eos;
#endif
}

/* face lists of the thread chunks (built on first use for the current thread count) */
static void face_plan_free(wo_flow *f) {
  if (f->plan_faces) {
    for (int t = 0; t < f->plan_nthr; t++) free(f->plan_faces[t]);
  }
  free(f->plan_faces);
  free(f->plan_nfaces);
  free(f->plan_c0);
  f->plan_faces = NULL;
  f->plan_nfaces = NULL;
  f->plan_c0 = NULL;
  f->plan_nthr = 0;
}

static void face_plan_build(wo_flow *f) {
  const wo_mesh *m = &f->mesh;
  int T = m->nowned >= WO_PAR_MIN ? wo_nthreads() : 1;
  if (f->plan_nthr == T) return;
  face_plan_free(f);
  f->plan_nthr = T;
  f->plan_c0 = (int *)malloc((T + 1) * sizeof(int));
  for (int t = 0; t <= T; t++) f->plan_c0[t] = (int)(((long long)m->nowned * t) / T);
  f->plan_nfaces = (int *)calloc(T, sizeof(int));
  f->plan_faces = (int32_t **)calloc(T, sizeof(int32_t *));
  /* chunk of an owned cell: chunks are equal-sized up to rounding, found by division + correction */
#define CHUNK_OF(c, out)                                          \
  do {                                                            \
    int t_ = (int)(((long long)(c)*T) / (m->nowned > 0 ? m->nowned : 1)); \
    if (t_ >= T) t_ = T - 1;                                      \
    while (t_ > 0 && (c) < f->plan_c0[t_]) t_--;                  \
    while (t_ < T - 1 && (c) >= f->plan_c0[t_ + 1]) t_++;         \
    (out) = t_;                                                   \
  } while (0)
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 1)
      for (int t = 0; t < T; t++) {
        f->plan_faces[t] = (int32_t *)malloc((f->plan_nfaces[t] + 1) * sizeof(int32_t));
        f->plan_nfaces[t] = 0;
      }
    for (int iface = 0; iface < m->nface; iface++) {
      const int32_t *cells = m->face_cells + 2 * (size_t)iface;
      int t0 = -1, t1 = -1;
      if (cells[0] < m->nowned) CHUNK_OF(cells[0], t0);
      if (cells[1] < m->nowned) CHUNK_OF(cells[1], t1);
      if (t0 < 0 && t1 < 0) t0 = 0; /* no owned cell: evaluated (and stored) once, added nowhere */
      if (t0 >= 0) {
        if (pass == 1) f->plan_faces[t0][f->plan_nfaces[t0]] = iface;
        f->plan_nfaces[t0]++;
      }
      if (t1 >= 0 && t1 != t0) {
        if (pass == 1) f->plan_faces[t1][f->plan_nfaces[t1]] = iface;
        f->plan_nfaces[t1]++;
      }
    }
  }
#undef CHUNK_OF
}


static inline int nint_(double x) { return (int)lround(x); }

wo_flow *wo_flow_create(const wo_params *prm, const wo_mesh *mesh) {
  wo_flow *f = (wo_flow *)calloc(1, sizeof(wo_flow));
  f->prm = *prm;
  f->mesh = *mesh;
  f->eos = wo_eos_create(prm);
  if (!f->eos) {
    free(f);
    return NULL;
  }
  f->np = wo_eos_num_primary(f->eos);
  f->nc = wo_eos_num_components(f->eos);
  f->nphase = wo_eos_num_phases(f->eos);
  f->nmobile = f->nphase; /* we / w / wge: all phases mobile */
  f->isothermal = (f->np == f->nc);
  f->dof = wo_eos_fluid_dof(f->eos);
  f->nflux = f->np + f->nmobile; /* flow_simulation.F90:174 */
  size_t nfl = (size_t)mesh->ncell * f->dof;
  f->fluid = (double *)calloc(nfl, sizeof(double));
  f->current_fluid = (double *)calloc(nfl, sizeof(double));
  f->last_iteration_fluid = (double *)calloc(nfl, sizeof(double));
  f->last_timestep_fluid = (double *)calloc(nfl, sizeof(double));
  f->balances = (double *)calloc((size_t)mesh->nowned * f->np, sizeof(double));
  f->flux = (double *)calloc((size_t)mesh->nface * f->nflux, sizeof(double));
  f->update = (double *)calloc(mesh->ncell, sizeof(double));
  f->rock = (double *)malloc((size_t)mesh->ncell * 8 * sizeof(double));
  memcpy(f->rock, mesh->rock, (size_t)mesh->ncell * 8 * sizeof(double));
  f->mesh.rock = f->rock;
  f->unperturbed = 1;
  return f;
}

void wo_flow_set_method(wo_flow *f, int method, double dt_last, const double *lhs_last2) {
  size_t n = (size_t)f->mesh.nowned * f->np;
  f->method = method;
  f->dt_last = dt_last;
  if (method == 1) {
    if (!f->lhs_last2) f->lhs_last2 = (double *)malloc(n * sizeof(double));
    memcpy(f->lhs_last2, lhs_last2, n * sizeof(double));
  }
}

/* Fixed-rate sources: cell (local owned index), component (1-based; 0 = all mass components, production
   only; np = heat), rate (kg/s or W; < 0 production), injection enthalpy (J/kg). */
void wo_flow_set_sources(wo_flow *f, int n, const int32_t *cell, const int32_t *component, const double *rate,
                         const double *enthalpy) {
  free(f->src_cell); free(f->src_component); free(f->src_rate); free(f->src_enthalpy); free(f->src_pcomponent);
  free(f->src_ctrl); free(f->src_direction); free(f->src_pi); free(f->src_pref); free(f->src_limit); free(f->src_rate_eval);
  free(f->src_sep_n); free(f->src_sep_h); free(f->src_limit_water); free(f->src_limit_steam);
  free(f->src_ptab_n); free(f->src_ptab_coord); free(f->src_ptab_step); free(f->src_ptab);
  f->src_ptab_n = f->src_ptab_coord = f->src_ptab_step = NULL;
  f->src_ptab = NULL;
  f->src_sep_n = NULL;
  f->src_sep_h = f->src_limit_water = f->src_limit_steam = NULL;
  f->src_ctrl = f->src_direction = NULL;
  f->src_pi = f->src_pref = f->src_limit = NULL;
  f->src_rate_eval = (double *)calloc(n + 1, sizeof(double));
  f->nsrc = n;
  f->src_cell = (int32_t *)malloc((n + 1) * sizeof(int32_t));
  f->src_component = (int32_t *)malloc((n + 1) * sizeof(int32_t));
  f->src_rate = (double *)malloc((n + 1) * sizeof(double));
  f->src_enthalpy = (double *)malloc((n + 1) * sizeof(double));
  memcpy(f->src_cell, cell, n * sizeof(int32_t));
  memcpy(f->src_component, component, n * sizeof(int32_t));
  f->src_pcomponent = (int32_t *)malloc((n + 1) * sizeof(int32_t));
  memcpy(f->src_pcomponent, component, n * sizeof(int32_t));
  memcpy(f->src_rate, rate, n * sizeof(double));
  memcpy(f->src_enthalpy, enthalpy, n * sizeof(double));
  memcpy(f->src_rate_eval, rate, n * sizeof(double));
}

/* injection / production component of every source (get_components, src/source_setup.F90:2052-2083); source%update_flow
   picks by the sign of the current rate (src/source.F90:372-380, 469-476) */
void wo_flow_set_source_components(wo_flow *f, int n, const int32_t *injection, const int32_t *production) {
  for (int k = 0; k < n && k < f->nsrc; k++) {
    f->src_component[k] = injection[k];
    f->src_pcomponent[k] = production[k];
  }
}
int wo_flow_source_component(const wo_flow *f, int s, double rate) {
  return rate > 0.0 ? f->src_component[s] : f->src_pcomponent[s];
}

/* Source controls for n of the sources of the last wo_flow_set_sources (source: index into those arrays):
   deliverability with productivity index pi and reference pressure pref (pi <= 0: no deliverability control),
   flow direction (0 both, 1 production only, 2 injection only), total-flow limiter (limit <= 0: none). */
void wo_flow_set_source_controls(wo_flow *f, int n, const int32_t *source, const double *pi, const double *pref,
                                 const int32_t *direction, const double *limit) {
  int ns = f->nsrc;
  free(f->src_ctrl); free(f->src_direction); free(f->src_pi); free(f->src_pref); free(f->src_limit);
  f->src_ctrl = (int32_t *)calloc(ns + 1, sizeof(int32_t));
  f->src_direction = (int32_t *)calloc(ns + 1, sizeof(int32_t));
  f->src_pi = (double *)calloc(ns + 1, sizeof(double));
  f->src_pref = (double *)calloc(ns + 1, sizeof(double));
  f->src_limit = (double *)calloc(ns + 1, sizeof(double));
  for (int k = 0; k < n; k++) {
    int s = source[k];
    f->src_ctrl[s] = pi[k] > 0.0;
    f->src_pi[s] = pi[k];
    f->src_pref[s] = pref[k];
    f->src_direction[s] = direction ? direction[k] : 0;
    f->src_limit[s] = limit ? limit[k] : 0.0;
  }
}

/* recharge / injectivity controls (recharge_source_control_iterator, src/source_control.F90:554-577; both input keys set
   up this control, src/source_setup.F90:2984-3092): rate = -coefficient (P - reference pressure); replaces a
   deliverability control of the same source, keeps its direction and limiter.  After wo_flow_set_source_controls. */
void wo_flow_set_source_recharge(wo_flow *f, int n, const int32_t *source, const double *coefficient, const double *pref) {
  if (!f->src_ctrl) wo_flow_set_source_controls(f, 0, NULL, NULL, NULL, NULL, NULL);
  for (int k = 0; k < n; k++) {
    int s = source[k];
    f->src_ctrl[s] = 2;
    f->src_pi[s] = coefficient[k];
    f->src_pref[s] = pref[k];
  }
}

void wo_flow_get_source_rates(const wo_flow *f, double *rate) { memcpy(rate, f->src_rate_eval, f->nsrc * sizeof(double)); }

/* separator_stage_init (src/separator.F90:108-136): reference water and steam enthalpies u + P / rho on the saturation
   line at the stage's pressure */
int wo_separator_stage(wo_thermo *th, double pressure, double *ref_water_enthalpy, double *ref_steam_enthalpy) {
  double saturation_temperature, params[2], water_props[2], steam_props[2];
  int err = wo_saturation_temperature(th, pressure, &saturation_temperature);
  if (err) return err;
  params[0] = pressure;
  params[1] = saturation_temperature;
  err = wo_region_properties(th, 1, params, water_props);
  if (err) return err;
  err = wo_region_properties(th, 2, params, steam_props);
  if (err) return err;
  *ref_water_enthalpy = water_props[1] + pressure / water_props[0];
  *ref_steam_enthalpy = steam_props[1] + pressure / steam_props[0];
  return 0;
}

/* separator_separate (src/separator.F90:212-260) over stages of separator_stage_separate (:140-166); stage_h holds
   (reference water enthalpy, reference steam enthalpy) per stage; out = water rate, water enthalpy, steam rate, steam
   enthalpy, steam fraction */
void wo_separate(int nstage, const double *stage_h, double rate, double enthalpy, double out[5]) {
  const double tol = 1.e-9;
  double q = rate, h = enthalpy, total_steam_mass_rate = 0.0, total_steam_energy_rate = 0.0;
  for (int i = 0; i < nstage; i++) {
    double ref_water_enthalpy = stage_h[2 * i], ref_steam_enthalpy = stage_h[2 * i + 1];
    double steam_fraction, water_enthalpy, steam_enthalpy;
    if (h <= ref_water_enthalpy) {
      steam_fraction = 0.0;
      water_enthalpy = h;
      steam_enthalpy = 0.0;
    } else if (h <= ref_steam_enthalpy) {
      steam_fraction = (h - ref_water_enthalpy) / (ref_steam_enthalpy - ref_water_enthalpy);
      water_enthalpy = ref_water_enthalpy;
      steam_enthalpy = ref_steam_enthalpy;
    } else {
      steam_fraction = 1.0;
      water_enthalpy = 0.0;
      steam_enthalpy = h;
    }
    double water_rate = (1.0 - steam_fraction) * q, steam_rate = steam_fraction * q;
    total_steam_mass_rate = total_steam_mass_rate + steam_rate;
    total_steam_energy_rate = total_steam_energy_rate + steam_rate * steam_enthalpy;
    q = water_rate;
    h = water_enthalpy;
  }
  out[0] = q;
  out[1] = h;
  out[2] = total_steam_mass_rate;
  out[3] = fabs(total_steam_mass_rate) > tol ? total_steam_energy_rate / total_steam_mass_rate : 0.0;
  out[4] = fabs(rate) > tol ? total_steam_mass_rate / rate : 0.0;
}

/* Separators and separated-flow limiters of n sources (after wo_flow_set_source_controls): nstage[k] <= 2 separator stages
   at pressure[2 k], pressure[2 k + 1]; limits on the separated water and steam mass rates (<= 0: none).  Source input
   "separator": {"pressure": ...} / "limiter": {"type": "water" | "steam", "limit": ..., "separator_pressure": ...} or
   {"total": ..., "water": ..., "steam": ...} (src/source_setup.F90:2255-2330, 3117-3276). */
int wo_flow_set_source_separators(wo_flow *f, int n, const int32_t *source, const int32_t *nstage, const double *pressure,
                                  const double *limit_water, const double *limit_steam) {
  int ns = f->nsrc;
  free(f->src_sep_n); free(f->src_sep_h); free(f->src_limit_water); free(f->src_limit_steam);
  f->src_sep_n = (int32_t *)calloc(ns + 1, sizeof(int32_t));
  f->src_sep_h = (double *)calloc(4 * (size_t)ns + 4, sizeof(double));
  f->src_limit_water = (double *)calloc(ns + 1, sizeof(double));
  f->src_limit_steam = (double *)calloc(ns + 1, sizeof(double));
  if (!f->src_ctrl) wo_flow_set_source_controls(f, 0, NULL, NULL, NULL, NULL, NULL);
  for (int k = 0; k < n; k++) {
    int s = source[k];
    if (nstage[k] < 0 || nstage[k] > 2) return 1;
    f->src_sep_n[s] = nstage[k];
    for (int i = 0; i < nstage[k]; i++) {
      int err = wo_separator_stage(wo_eos_thermo(f->eos), pressure[2 * k + i], &f->src_sep_h[4 * s + 2 * i], &f->src_sep_h[4 * s + 2 * i + 1]);
      if (err) return err;
    }
    f->src_limit_water[s] = limit_water ? limit_water[k] : 0.0;
    f->src_limit_steam[s] = limit_steam ? limit_steam[k] : 0.0;
  }
  return 0;
}

/* enthalpy of the fluid a producing source takes from its cell: fluid%specific_enthalpy(fluid%phase_flow_fractions())
   (src/fluid.F90:394-436), what source_separator_iterator (src/source_network.F90:197-216) gives the separator */
static double source_fluid_enthalpy(const wo_flow *f, int s) {
  const double *fl = f->current_fluid + (size_t)f->src_cell[s] * f->dof;
  int phases = nint_(fl[4]);
  double frac[2] = {0.0, 0.0}, sum = 0.0, h = 0.0;
  for (int p = 0; p < f->nphase; p++) {
    const double *ph = fl + (7 + f->nc - 1) + p * (8 + f->nc - 1);
    if (phases & (1 << p)) frac[p] = ph[3] * ph[0] / ph[1];
    sum += frac[p];
  }
  for (int p = 0; p < f->nphase; p++) {
    const double *ph = fl + (7 + f->nc - 1) + p * (8 + f->nc - 1);
    frac[p] = frac[p] / sum;
    if (phases & (1 << p)) h = h + frac[p] * ph[5];
  }
  return h;
}

/* Reference pressure tables of n sources on deliverability (after wo_flow_set_source_controls): npts[k] <= WO_PTAB_MAX
   points (x, y) at table[2 WO_PTAB_MAX k ...], x the flowing enthalpy (coordinate[k] = 0) or the pressure (1) of the
   source's cell, linear or step (step[k] != 0) interpolation, constant beyond the ends -- source input
   "deliverability": {"pressure": {"enthalpy": [[h, P], ...]}} (src/source_setup.F90, src/source_control.F90:376-388). */
int wo_flow_set_source_pressure_table(wo_flow *f, int n, const int32_t *source, const int32_t *coordinate, const int32_t *step,
                                      const int32_t *npts, const double *table) {
  int ns = f->nsrc;
  free(f->src_ptab_n); free(f->src_ptab_coord); free(f->src_ptab_step); free(f->src_ptab);
  f->src_ptab_n = (int32_t *)calloc(ns + 1, sizeof(int32_t));
  f->src_ptab_coord = (int32_t *)calloc(ns + 1, sizeof(int32_t));
  f->src_ptab_step = (int32_t *)calloc(ns + 1, sizeof(int32_t));
  f->src_ptab = (double *)calloc(2 * WO_PTAB_MAX * ((size_t)ns + 1), sizeof(double));
  for (int k = 0; k < n; k++) {
    int s = source[k];
    if (npts[k] < 1 || npts[k] > WO_PTAB_MAX) return 1;
    f->src_ptab_n[s] = npts[k];
    f->src_ptab_coord[s] = coordinate[k];
    f->src_ptab_step[s] = step ? step[k] : 0;
    for (int i = 0; i < 2 * npts[k]; i++) f->src_ptab[2 * WO_PTAB_MAX * (size_t)s + i] = table[2 * WO_PTAB_MAX * (size_t)k + i];
  }
  return 0;
}

/* interpolation_table%interpolate (src/interpolation.F90:300-420): find the interval, linear or step; constant outside */
static double table_interpolate(const double *tab, int n, int step, double x) {
  if (x <= tab[0]) return tab[1];
  if (x >= tab[2 * (n - 1)]) return tab[2 * (n - 1) + 1];
  int i = 0;
  while (i + 2 < n && x >= tab[2 * (i + 1)]) i++;
  if (step) return tab[2 * i + 1];
  double xi = (x - tab[2 * i]) / (tab[2 * i + 2] - tab[2 * i]);
  return tab[2 * i + 1] + (tab[2 * i + 3] - tab[2 * i + 1]) * xi;
}

/* separated flows of source s at the given rate (source_network_node_get_separated_flows, src/source_network_node.F90:
   116-131): zero unless the source produces and has a separator */
void wo_flow_source_separated(const wo_flow *f, int s, double rate, double out[5]) {
  for (int i = 0; i < 5; i++) out[i] = 0.0;
  if (f->src_sep_n && f->src_sep_n[s] > 0 && rate < 0.0)
    wo_separate(f->src_sep_n[s], f->src_sep_h + 4 * s, rate, source_fluid_enthalpy(f, s), out);
}

/* source_network%update for one source (src/source_network.F90:90-292): source controls, then network controls.
   deliverability_source_control_flow_rate (src/source_control.F90:359-403, constant productivity and reference
   pressure), direction_source_control_iterator (:596-620), limit_rate with a "total" limiter
   (src/source_network_node.F90:245-315) */
double wo_flow_source_rate(const wo_flow *f, int s) {
  double rate = f->src_rate[s];
  if (!f->src_ctrl) return rate;
  const double *fl = f->current_fluid + (size_t)f->src_cell[s] * f->dof;
  if (f->src_ctrl[s] == 2) {
    double pressure_difference = fl[0] - f->src_pref[s];
    rate = -f->src_pi[s] * pressure_difference;
  } else if (f->src_ctrl[s]) {
    int phases = nint_(fl[4]);
    double effective_productivity = f->src_pi[s] * fl[5]; /* permeability_factor */
    double reference_pressure = f->src_pref[s];
    if (f->src_ptab_n && f->src_ptab_n[s] > 0) {
      double x = f->src_ptab_coord[s] ? fl[0] : source_fluid_enthalpy(f, s);
      reference_pressure = table_interpolate(f->src_ptab + 2 * WO_PTAB_MAX * (size_t)s, f->src_ptab_n[s], f->src_ptab_step[s], x);
    }
    double pressure_difference = fl[0] - reference_pressure;
    rate = 0.0;
    for (int p = 0; p < f->nphase; p++) {
      const double *ph = fl + (7 + f->nc - 1) + p * (8 + f->nc - 1);
      if (phases & (1 << p)) rate = rate - effective_productivity * (ph[3] * ph[0] / ph[1]) * pressure_difference;
    }
  }
  if (f->src_direction[s] == 1 && !(rate < 0.0)) rate = 0.0;
  if (f->src_direction[s] == 2 && !(rate > 0.0)) rate = 0.0;
  /* limiter: source_network_node_limit_rate (src/source_network_node.F90:245-315) over the limited flow types (total,
     separated water, separated steam): the smallest scale that brings every rate over its limit back to it */
  {
    const double small = 1.e-6;
    double scale = 1.0, sep[5];
    int over = 0;
    int has_sep_limit = f->src_sep_n && (f->src_limit_water[s] > 0.0 || f->src_limit_steam[s] > 0.0);
    if (has_sep_limit) wo_flow_source_separated(f, s, rate, sep);
    for (int type = 0; type < 3; type++) {
      double limit = type == 0 ? f->src_limit[s] : (has_sep_limit ? (type == 1 ? f->src_limit_water[s] : f->src_limit_steam[s]) : 0.0);
      if (!(limit > 0.0)) continue;
      double abs_rate = fabs(type == 0 ? rate : (type == 1 ? sep[0] : sep[2]));
      if (abs_rate > limit) {
        over = 1;
        if (abs_rate > small) scale = fmin(scale, limit / abs_rate);
      }
    }
    if (over) rate = rate * scale;
  }
  return rate;
}

/* fluid%phase_flow_fractions (src/fluid.F90:394-411) of the current fluid in the cell of source s */
void wo_flow_source_phase_fractions(const wo_flow *f, int s, double *frac) {
  const double *fl = f->current_fluid + (size_t)f->src_cell[s] * f->dof;
  int phases = nint_(fl[4]);
  double sum = 0.0;
  for (int p = 0; p < f->nphase; p++) {
    const double *ph = fl + (7 + f->nc - 1) + p * (8 + f->nc - 1);
    frac[p] = 0.0;
    if (phases & (1 << p)) frac[p] = ph[3] * ph[0] / ph[1];
    sum += frac[p];
  }
  for (int p = 0; p < f->nphase; p++) frac[p] = frac[p] / sum;
}

/* source%update_flow: src/source.F90:457-480 with update_injection_mass_flow :385-399,
   update_production_mass_flow :403-438 (fluid%phase_flow_fractions / component_flow_fractions /
   specific_enthalpy src/fluid.F90:374-456) and update_energy_flow :442-453 */
static void source_flow(const wo_flow *f, int s, double *flow) {
  int np = f->np, nc = f->nc;
  double rate = wo_flow_source_rate(f, s), enthalpy = 0.0;
  int component = wo_flow_source_component(f, s, rate);
  if (f->unperturbed) f->src_rate_eval[s] = rate;
  for (int k = 0; k < np; k++) flow[k] = 0.0;
  if (rate > 0.0) {
    if (component > 0) {
      enthalpy = f->src_enthalpy[s];
      flow[component - 1] = rate;
    }
  } else {
    const double *fl = f->current_fluid + (size_t)f->src_cell[s] * f->dof;
    int phases = nint_(fl[4]);
    double frac[2] = {0.0, 0.0};
    if (component < np) {
      double sum = 0.0;
      for (int p = 0; p < f->nphase; p++) {
        const double *ph = fl + (7 + nc - 1) + p * (8 + nc - 1);
        if (phases & (1 << p)) frac[p] = ph[3] * ph[0] / ph[1]; /* mobility kr*rho/mu */
        sum += frac[p];
      }
      for (int p = 0; p < f->nphase; p++) frac[p] = frac[p] / sum;
      if (!f->isothermal) {
        for (int p = 0; p < f->nphase; p++) {
          const double *ph = fl + (7 + nc - 1) + p * (8 + nc - 1);
          if (phases & (1 << p)) enthalpy = enthalpy + frac[p] * ph[5];
        }
      }
    }
    if (component <= 0) {
      double cf[WO_MAX_NC], csum = 0.0;
      for (int c = 0; c < nc; c++) {
        cf[c] = 0.0;
        for (int p = 0; p < f->nphase; p++) {
          const double *ph = fl + (7 + nc - 1) + p * (8 + nc - 1);
          if (phases & (1 << p)) cf[c] = cf[c] + frac[p] * ph[7 + c];
        }
        csum += cf[c];
      }
      for (int c = 0; c < nc; c++) flow[c] = rate * (cf[c] / csum);
    } else {
      flow[component - 1] = rate;
    }
  }
  if (!f->isothermal && component < np) flow[np - 1] = flow[np - 1] + enthalpy * rate;
}

void wo_flow_destroy(wo_flow *f) {
  if (!f) return;
  free(f->src_cell); free(f->src_component); free(f->src_rate); free(f->src_enthalpy); free(f->src_pcomponent);
  free(f->src_ctrl); free(f->src_direction); free(f->src_pi); free(f->src_pref); free(f->src_limit); free(f->src_rate_eval);
  free(f->src_sep_n); free(f->src_sep_h); free(f->src_limit_water); free(f->src_limit_steam);
  free(f->src_ptab_n); free(f->src_ptab_coord); free(f->src_ptab_step); free(f->src_ptab);
  free(f->lhs_last2);
  face_plan_free(f);
  for (int t = 1; t < f->eos_nthr; t++) wo_eos_destroy(f->eos_thr[t]);
  free(f->eos_thr);
  free(f->tracers);
  free(f->tracer_injection);
  wo_eos_destroy(f->eos);
  free(f->fluid);
  free(f->current_fluid);
  free(f->last_iteration_fluid);
  free(f->last_timestep_fluid);
  free(f->balances);
  free(f->flux);
  free(f->update);
  free(f->rock);
  free(f);
}

wo_eos *wo_flow_eos(wo_flow *f) { return f->eos; }
double *wo_flow_fluid(wo_flow *f) { return f->fluid; }
double *wo_flow_current_fluid(wo_flow *f) { return f->current_fluid; }
double *wo_flow_flux(wo_flow *f) { return f->flux; }

void wo_flow_get_regions(wo_flow *f, int32_t *region) {
  for (int c = 0; c < f->mesh.ncell; c++) region[c] = nint_(f->fluid[(size_t)c * f->dof + 2]);
}

/* boundary ghost cell: mesh.F90:1185-1202 (rock copied from interior cell, fluid from BC primary) */
int wo_flow_set_boundary(wo_flow *f, int ghost_cell, int interior_cell, const double *primary, int region) {
  double *rk = f->rock + (size_t)ghost_cell * 8;
  memcpy(rk, f->rock + (size_t)interior_cell * 8, 8 * sizeof(double));
  double *fl = f->fluid + (size_t)ghost_cell * f->dof;
  fl[2] = (double)region;
  int err = wo_eos_bulk_properties(f->eos, primary, fl);
  if (err == 0) err = wo_eos_phase_properties(f->eos, primary, rk, fl);
  memcpy(f->current_fluid + (size_t)ghost_cell * f->dof, fl, f->dof * sizeof(double));
  return err;
}

/* update_rock_properties: flow_simulation.F90:2051-2089 (rock records of the interior cells replaced between time
   steps by the table controls of rock_control.F90:49-116; boundary ghost cells keep their copies) */
int wo_flow_set_rock(wo_flow *f, const double *rock) {
  memcpy(f->rock, rock, (size_t)f->mesh.ninterior * 8 * sizeof(double));
  return 0;
}

/* fluid_init: flow_simulation.F90:2171-2287 (regions supplied with the initial conditions) */
int wo_flow_fluid_init(wo_flow *f, const double *y, const int32_t *region) {
  int err = 0;
  double primary[WO_MAX_NP];
  for (int c = 0; c < f->mesh.nowned; c++) {
    double *fl = f->fluid + (size_t)c * f->dof;
    fl[2] = (double)region[c];
    wo_eos_unscale(f->eos, y + (size_t)c * f->np, region[c], primary);
    err = wo_eos_bulk_properties(f->eos, primary, fl);
    if (err == 0) err = wo_eos_phase_properties(f->eos, primary, f->rock + (size_t)c * 8, fl);
    if (err) break;
  }
  memcpy(f->current_fluid, f->fluid, (size_t)f->mesh.ncell * f->dof * sizeof(double));
  return err;
}

/* identify_update_cells: flow_simulation.F90:1102-1137 */
static void identify_update_cells(wo_flow *f, const int32_t *perturbed, int nperturbed) {
  f->unperturbed = (nperturbed == 0);
  const int ncell = f->mesh.ncell;
  double *update = f->update;
  if (f->unperturbed) {
#pragma omp parallel for schedule(static) if (ncell >= WO_PAR_MIN)
    for (int c = 0; c < ncell; c++) update[c] = 1.0;
  } else {
#pragma omp parallel for schedule(static) if (ncell >= WO_PAR_MIN)
    for (int c = 0; c < ncell; c++) update[c] = -1.0;
#pragma omp parallel for schedule(static) if (nperturbed >= WO_PAR_MIN)
    for (int k = 0; k < nperturbed; k++) update[perturbed[k]] = 1.0;
  }
}

/* fluid_properties: flow_simulation.F90:2291-2415 */
static int fluid_properties(wo_flow *f, const double *y) {
  int err = 0, first_bad = f->mesh.nowned;
  eos_pool_build(f);
  par_memcpy(f->current_fluid, f->fluid, (size_t)f->mesh.ncell * f->dof * sizeof(double)); /* :2331 */
  /* the serial loop stops at the first cell in error; here every cell is evaluated and the error of the
     lowest-numbered failing cell is returned (the evaluation is abandoned by the caller either way) */
#pragma omp parallel for schedule(static) if (f->mesh.nowned >= WO_PAR_MIN)
  for (int c = 0; c < f->mesh.nowned; c++) {
    if (f->update[c] > 0) {
      double primary[WO_MAX_NP];
      wo_eos *eos = thread_eos(f);
      double *fl = f->current_fluid + (size_t)c * f->dof;
      wo_eos_unscale(eos, y + (size_t)c * f->np, nint_(fl[2]), primary);
      int e = wo_eos_bulk_properties(eos, primary, fl);
      if (e == 0) e = wo_eos_phase_properties(eos, primary, f->rock + (size_t)c * 8, fl);
      if (e) {
#pragma omp critical(wo_fluid_err)
        if (c < first_bad) {
          first_bad = c;
          err = e;
        }
      }
    }
  }
  return err;
}

/* pre_eval: flow_simulation.F90:2126-2147 */
int wo_flow_pre_eval(wo_flow *f, const double *y, const int32_t *perturbed, int nperturbed) {
  identify_update_cells(f, perturbed, nperturbed);
  int err = fluid_properties(f, y);
  if (f->unperturbed) par_memcpy(f->fluid, f->current_fluid, (size_t)f->mesh.ncell * f->dof * sizeof(double));
  return err;
}

/* cell_balances: flow_simulation.F90:1242-1330 */
int wo_flow_cell_balances(wo_flow *f, double *lhs) {
  size_t n = (size_t)f->mesh.nowned * f->np;
  par_memcpy(lhs, f->balances, n * sizeof(double)); /* :1273 */
#pragma omp parallel for schedule(static) if (f->mesh.nowned >= WO_PAR_MIN)
  for (int c = 0; c < f->mesh.nowned; c++) {
    if (f->update[c] > 0)
      wo_cell_balance(f->rock + (size_t)c * 8, f->current_fluid + (size_t)c * f->dof, f->nc, f->nphase, f->np,
                      lhs + (size_t)c * f->np);
  }
  if (f->unperturbed) par_memcpy(f->balances, lhs, n * sizeof(double));
  return 0;
}

/* cell_inflows: flow_simulation.F90:1334-1485, then the source terms (:1468-1473,
   source_network.F90:296-355: inflow = inflow + flow / volume, in source order) */
int wo_flow_cell_inflows(wo_flow *f, double *rhs) {
  const wo_mesh *m = &f->mesh;
  int np = f->np, nf = f->nflux;
  const double flux_sign[2] = {-1.0, 1.0};
  double flow[WO_MAX_NP];
  face_plan_build(f);
  const int T = f->plan_nthr;
#pragma omp parallel for schedule(static, 1) if (T > 1)
  for (int t = 0; t < T; t++) {
    const int c0 = f->plan_c0[t], c1 = f->plan_c0[t + 1];
    double face_flux[WO_MAX_NP + 3], fl[WO_MAX_NP];
    for (size_t i = (size_t)c0 * np; i < (size_t)c1 * np; i++) rhs[i] = 0.0;
    for (int q = 0; q < f->plan_nfaces[t]; q++) {
      const int iface = f->plan_faces[t][q];
      const int32_t *cells = m->face_cells + 2 * (size_t)iface;
      const double *g = m->face_geom + 12 * (size_t)iface;
      int update_flux = 0;
      for (int i = 0; i < 2; i++)
        if (cells[i] < m->ninterior) update_flux = update_flux || (f->update[cells[i]] > 0);
      double *stored = f->flux + (size_t)iface * nf;
      if (update_flux) {
        wo_face_flux(g, f->rock + 8 * (size_t)cells[0], f->rock + 8 * (size_t)cells[1],
                     f->current_fluid + (size_t)cells[0] * f->dof, f->current_fluid + (size_t)cells[1] * f->dof,
                     f->nc, np, f->nphase, f->nmobile, f->isothermal, face_flux);
        /* a face shared by two chunks is evaluated by both (as by two ranks in the reference); the chunk of its
           first owned cell keeps the stored copy */
        const int writer_cell = cells[0] < m->nowned ? cells[0] : cells[1];
        const int mine = (writer_cell >= c0 && writer_cell < c1) || (writer_cell >= m->nowned && t == 0);
        if (f->unperturbed && mine) memcpy(stored, face_flux, nf * sizeof(double));
      } else {
        memcpy(face_flux, stored, nf * sizeof(double));
      }
      for (int k = 0; k < np; k++) fl[k] = face_flux[k] * g[0];
      for (int i = 0; i < 2; i++) {
        int c = cells[i];
        if (c >= c0 && c < c1) { /* ghost_cell(c) < 0 and c < end_interior_cell, and the cell is this chunk's */
          double vol = m->cell_geom[4 * (size_t)c + 3];
          double *inflow = rhs + (size_t)c * np;
          for (int k = 0; k < np; k++) inflow[k] = inflow[k] + flux_sign[i] * fl[k] / vol;
        }
      }
    }
  }
  for (int s = 0; s < f->nsrc; s++) {
    int c = f->src_cell[s];
    if (c >= 0 && c < m->nowned) {
      double vol = m->cell_geom[4 * (size_t)c + 3];
      double *inflow = rhs + (size_t)c * np;
      source_flow(f, s, flow);
      for (int k = 0; k < np; k++) inflow[k] = inflow[k] + flow[k] / vol;
    }
  }
  return 0;
}

void wo_flow_pre_iteration(wo_flow *f) { /* :2108-2122 */
  memcpy(f->last_iteration_fluid, f->fluid, (size_t)f->mesh.ncell * f->dof * sizeof(double));
}
void wo_flow_pre_timestep(wo_flow *f) { /* :2022-2035 */
  memcpy(f->last_timestep_fluid, f->fluid, (size_t)f->mesh.ncell * f->dof * sizeof(double));
}
void wo_flow_pre_retry_timestep(wo_flow *f) { /* :2093-2104 */
  memcpy(f->fluid, f->last_timestep_fluid, (size_t)f->mesh.ncell * f->dof * sizeof(double));
}

/* fluid_transitions: flow_simulation.F90:2419-2576.  changed_y keeps the reference's
   semantics: it is overwritten per cell by check_primary_variables, so its final value
   is the last visited cell's; changed_search is sticky. */
int wo_flow_fluid_transitions(wo_flow *f, const double *y_old, double *search, double *y, int *changed_search,
                              int *changed_y) {
  int err = 0, np = f->np, first_bad = f->mesh.nowned;
  int any_search = 0, last_changed = 0;
  const int nowned = f->mesh.nowned;
  eos_pool_build(f);
  /* cells are independent; changed_y ends up as the value of the last cell visited, changed_search as the OR over
     all cells, exactly as in the serial loop (which stops at the first cell in error: here the error of the
     lowest-numbered failing cell is returned and the caller abandons the step) */
#pragma omp parallel for schedule(static) reduction(| : any_search) if (nowned >= WO_PAR_MIN)
  for (int c = 0; c < nowned; c++) {
    double primary[WO_MAX_NP], old_primary[WO_MAX_NP];
    double *fl = f->fluid + (size_t)c * f->dof;
    const double *ofl = f->last_iteration_fluid + (size_t)c * f->dof;
    double *yc = y + (size_t)c * np;
    const double *yo = y_old + (size_t)c * np;
    double *sc = search + (size_t)c * np;
    int transition = 0, changed = 0;
    wo_eos *eos = thread_eos(f);
    wo_eos_unscale(eos, yc, nint_(fl[2]), primary);
    wo_eos_unscale(eos, yo, nint_(ofl[2]), old_primary);
    fl[3] = fl[2]; /* old_region = region :2502 */
    int e = wo_eos_transition(eos, old_primary, primary, ofl, fl, &transition);
    if (e == 0) {
      e = wo_eos_check_primary_variables(eos, fl, primary, &changed);
      if (e == 0) {
        if (transition) changed = 1;
        if (changed) {
          any_search = 1;
          wo_eos_scale(eos, primary, nint_(fl[2]), yc);
          for (int k = 0; k < np; k++) sc[k] = yo[k] - yc[k];
        }
      }
    }
    if (c == nowned - 1) last_changed = changed;
    if (e) {
#pragma omp critical(wo_trans_err)
      if (c < first_bad) {
        first_bad = c;
        err = e;
      }
    }
  }
  *changed_search = any_search;
  *changed_y = last_changed;
  return err;
}

/* SNES_residual + backwards_Euler_residual: timestepper.F90:345-374, 587-624 */
int wo_residual_be(wo_flow *f, const double *y, const double *lhs_last, double dt, const int32_t *perturbed,
                   int nperturbed, double *lhs, double *rhs, double *r) {
  size_t n = (size_t)f->mesh.nowned * f->np;
  int err = wo_flow_pre_eval(f, y, perturbed, nperturbed);
  if (err) return err;
  if (f->method == 2) { /* direct_ss_residual: timestepper.F90:431-452 */
    err = wo_flow_cell_balances(f, lhs);
    if (err) return err;
    err = wo_flow_cell_inflows(f, rhs);
    if (err) return err;
    par_memcpy(r, rhs, n * sizeof(double)); /* VecCopy */
    return 0;
  }
  err = wo_flow_cell_balances(f, lhs);
  if (err) return err;
  if (f->method == 1) { /* BDF2_residual: timestepper.F90:378-427 */
    double q = dt / f->dt_last, q1 = q + 1.0;
#pragma omp parallel for schedule(static) if (n >= WO_PAR_MIN)
    for (size_t i = 0; i < n; i++) r[i] = lhs[i];                            /* VecCopy */
#pragma omp parallel for schedule(static) if (n >= WO_PAR_MIN)
    for (size_t i = 0; i < n; i++) r[i] = r[i] * (1.0 + 2.0 * q);            /* VecScale */
#pragma omp parallel for schedule(static) if (n >= WO_PAR_MIN)
    for (size_t i = 0; i < n; i++) r[i] = r[i] + (-q1 * q1) * lhs_last[i];   /* VecAXPY */
#pragma omp parallel for schedule(static) if (n >= WO_PAR_MIN)
    for (size_t i = 0; i < n; i++) r[i] = r[i] + (q * q) * f->lhs_last2[i];  /* VecAXPY */
    err = wo_flow_cell_inflows(f, rhs);
    if (err) return err;
#pragma omp parallel for schedule(static) if (n >= WO_PAR_MIN)
    for (size_t i = 0; i < n; i++) r[i] = r[i] + (-dt * q1) * rhs[i];        /* VecAXPY */
    return 0;
  }
#pragma omp parallel for schedule(static) if (n >= WO_PAR_MIN)
  for (size_t i = 0; i < n; i++) r[i] = lhs[i];                 /* VecCopy */
#pragma omp parallel for schedule(static) if (n >= WO_PAR_MIN)
  for (size_t i = 0; i < n; i++) r[i] = r[i] + (-1.0) * lhs_last[i]; /* VecAXPY */
  err = wo_flow_cell_inflows(f, rhs);
  if (err) return err;
#pragma omp parallel for schedule(static) if (n >= WO_PAR_MIN)
  for (size_t i = 0; i < n; i++) r[i] = r[i] + (-dt) * rhs[i];   /* VecAXPY */
  return 0;
}

/* dm_utils.F90:644-685; VecMax returns the first location of the maximum */
void wo_vec_max_pointwise_abs_scale(const double *v, const double *scale, double tol, int n, double *maxval,
                                    int *maxloc) {
  double best = -INFINITY;
  int loc = -1;
  for (int i = 0; i < n; i++) {
    double s = fmax(fabs(scale[i]), tol);
    double q = fabs(v[i]) / s;
    if (q > best) {
      best = q;
      loc = i;
    }
  }
  *maxval = best;
  *maxloc = loc;
}
