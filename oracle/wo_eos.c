/*
 * wo_eos.c -- oracle (TEST INFRASTRUCTURE): equations of state (we, w, wce) and
 * the local cell / face objects.  Restated from src/eos.F90:186-257,
 * src/eos_we.F90:149-526, src/eos_w.F90, src/eos_wge.F90:40-705, src/eos_wce.F90,
 * src/ncg_thermodynamics.F90:145-340, src/ncg_co2_thermodynamics.F90:14-292, src/fluid.F90:197-370,
 * src/rock.F90:142, src/cell.F90:114-142, src/face.F90:230-515.
 *
 * Fluid record layout (src/fluid.F90:232-267), nc components, nph phases:
 *   0 pressure, 1 temperature, 2 region, 3 old_region, 4 phase_composition,
 *   5 permeability_factor, 6..6+nc-1 partial_pressure, then per phase
 *   (stride 8+nc-1): density, viscosity, saturation, relative_permeability,
 *   capillary_pressure, specific_enthalpy, internal_energy, mass_fraction(nc).
 * Rock record (src/rock.F90:97-112): permeability(3), wet_conductivity,
 *   dry_conductivity, porosity, density, specific_heat.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct wo_eos {
  wo_params prm;
  wo_thermo *thermo;
  int np, nc, nphase, nmobile, isothermal;
  double primary_scale[WO_MAX_NP][5]; /* [var][region 1..4] */
  int adaptive_pp_scale;              /* eos_wge: partial pressure scaled by the cell's total pressure */
  int gas;                            /* eos_wge: 0 CO2 (eos_wce), 1 air (eos_wae) */
};

enum { F_P = 0, F_T = 1, F_REGION = 2, F_OLD_REGION = 3, F_PHASES = 4, F_PERMFAC = 5, F_PARTIAL = 6 };
enum { PH_RHO = 0, PH_MU = 1, PH_SAT = 2, PH_KR = 3, PH_PC = 4, PH_H = 5, PH_U = 6, PH_X = 7 };
enum { R_PERM = 0, R_WET = 3, R_DRY = 4, R_POR = 5, R_RHO = 6, R_CP = 7 };

static inline int bulk_dof(int nc) { return 7 + nc - 1; }
static inline int phase_dof(int nc) { return 8 + nc - 1; }
static inline double *phase_ptr(double *fluid, int nc, int p) { return fluid + bulk_dof(nc) + p * phase_dof(nc); }
static inline const double *cphase_ptr(const double *fluid, int nc, int p) {
  return fluid + bulk_dof(nc) + p * phase_dof(nc);
}
static inline int nint_(double x) { return (int)lround(x); }

wo_eos *wo_eos_create(const wo_params *prm) {
  wo_eos *e = (wo_eos *)calloc(1, sizeof(wo_eos));
  e->prm = *prm;
  if (prm->eos == WO_EOS_WAE) { /* eos_wae.F90:27-64: eos_wge with the air NCG */
    e->gas = 1;
    e->prm.eos = WO_EOS_WCE;
  }
  e->thermo = wo_thermo_create(prm->thermo, prm->extrapolate);
  double ps = prm->pressure_scale > 0 ? prm->pressure_scale : 1.e6;      /* eos_we.F90:75-76 */
  double ts = prm->temperature_scale > 0 ? prm->temperature_scale : 1.e2;
  switch (e->prm.eos) {
    case WO_EOS_WE: /* eos_we.F90:78-109 */
      e->np = 2; e->nc = 1; e->nphase = 2; e->nmobile = 2; e->isothermal = 0;
      e->primary_scale[0][1] = ps; e->primary_scale[1][1] = ts;
      e->primary_scale[0][2] = ps; e->primary_scale[1][2] = ts;
      e->primary_scale[0][3] = 0.0; e->primary_scale[1][3] = 0.0;
      e->primary_scale[0][4] = ps; e->primary_scale[1][4] = 1.0;
      break;
    case WO_EOS_W: /* eos_w.F90:67-96 */
      e->np = 1; e->nc = 1; e->nphase = 1; e->nmobile = 1; e->isothermal = 1;
      e->primary_scale[0][1] = ps; e->primary_scale[0][2] = ps;
      break;
    case WO_EOS_WCE: { /* eos_wge.F90:40-131 + eos_wce.F90:23-53 */
      e->np = 3; e->nc = 2; e->nphase = 2; e->nmobile = 2; e->isothermal = 0;
      double pps = prm->partial_pressure_scale;
      e->adaptive_pp_scale = !(pps > 0.0);
      if (e->adaptive_pp_scale) pps = 0.0;
      e->primary_scale[0][1] = ps; e->primary_scale[1][1] = ts; e->primary_scale[2][1] = pps;
      e->primary_scale[0][2] = ps; e->primary_scale[1][2] = ts; e->primary_scale[2][2] = pps;
      e->primary_scale[0][4] = ps; e->primary_scale[1][4] = 1.0; e->primary_scale[2][4] = pps;
      break;
    }
    default:
      wo_thermo_destroy(e->thermo);
      free(e);
      return NULL;
  }
  return e;
}

void wo_eos_destroy(wo_eos *e) {
  if (!e) return;
  wo_thermo_destroy(e->thermo);
  free(e);
}
int wo_eos_num_primary(const wo_eos *e) { return e->np; }
int wo_eos_num_components(const wo_eos *e) { return e->nc; }
int wo_eos_num_phases(const wo_eos *e) { return e->nphase; }
int wo_eos_fluid_dof(const wo_eos *e) { return bulk_dof(e->nc) + e->nphase * phase_dof(e->nc); }
wo_thermo *wo_eos_thermo(wo_eos *e) { return e->thermo; }

/* eos.F90:186-210 ; adaptive: eos_wge.F90:639-674 */
void wo_eos_scale(const wo_eos *e, const double *primary, int region, double *scaled) {
  if (e->adaptive_pp_scale) {
    for (int i = 0; i < 2; i++) scaled[i] = primary[i] / e->primary_scale[i][region];
    scaled[2] = primary[2] / primary[0];
    return;
  }
  for (int i = 0; i < e->np; i++) scaled[i] = primary[i] / e->primary_scale[i][region];
}
void wo_eos_unscale(const wo_eos *e, const double *scaled, int region, double *primary) {
  if (e->adaptive_pp_scale) {
    for (int i = 0; i < 2; i++) primary[i] = scaled[i] * e->primary_scale[i][region];
    primary[2] = scaled[2] * primary[0];
    return;
  }
  for (int i = 0; i < e->np; i++) primary[i] = scaled[i] * e->primary_scale[i][region];
}

/* ---- CO2 (src/ncg_co2_thermodynamics.F90) and the NCG base class (src/ncg_thermodynamics.F90) ---- */
static const double co2_molecular_weight = 44.01;     /* ncg_co2_thermodynamics.F90:14 */
static const double water_molecular_weight = 18.01528; /* thermodynamics.F90:38 */
static const double gas_constant = 8.3144598;          /* thermodynamics.F90:39 */
static const double henry_data[6] = {0.783666, 1.96025, 8.20574, -7.40674, 2.18380, -0.220999};
static const double co2_tscale = 100.0;
/* viscosity_data(5,6), column-major: pressures (MPa), then 5 polynomial coefficients per pressure (:22-30) */
static const double co2_visc_p[5] = {0.0, 10.0, 15.0, 20.0, 30.0};
static const double co2_visc_c[5][5] = {
    {1.3578, 3.9189, 9.6607, 13.1566, 14.7968},
    {4.9227e-3, -35.984e-3, -135.479e-3, -179.352e-3, -160.731e-3},
    {-2.9661e-6, 0.25825e-3, 0.90087e-3, 1.12474e-3, 0.850257e-3},
    {2.8529e-9, -7.1178e-7, -2.4727e-6, -2.98864e-6, -1.99076e-6},
    {-2.1829e-12, 6.9578e-10, 2.4156e-9, 2.85911e-9, 1.73423e-9}};

/* utils.F90:224-241 (Horner) */
static double polynomial(const double *a, int n, double x) {
  double p = a[n - 1];
  for (int i = n - 2; i >= 0; i--) p = a[i] + x * p;
  return p;
}

/* ncg_co2_thermodynamics.F90:84-111 */
void wo_co2_properties(double partial_pressure, double temperature, double props[2]) {
  double tk = temperature + WO_TC_K;
  double pp = partial_pressure * 1.0e-6;
  double tc = pow(0.01 * tk, 3.3333333333);
  double hci = 1.667 + 0.001542 * tk - 0.7948 * log10(tk) - 41.35 / tk;
  props[1] = 1.e6 * (hci - 0.3571 * pp * (1.0 + 0.07576 * pp) / tc);
  double vc = 0.00018882 * tk - pp * (0.0824 + 0.01249 * pp) / tc;
  props[0] = pp / vc;
}

/* ncg_co2_thermodynamics.F90:115-135 */
double wo_co2_henrys_constant(double temperature) {
  return 1.e8 * polynomial(henry_data, 6, temperature / co2_tscale);
}

/* henrys_derivative (:172-197) with polynomial_derivative (utils.F90:291-310: da(i) = i*a(i+1)), then
   energy_solution (ncg_thermodynamics.F90:196-223, 176-192) */
double wo_co2_energy_solution(double temperature, double henrys_constant) {
  double da[5];
  for (int i = 0; i < 5; i++) da[i] = (double)(i + 1) * henry_data[i + 1];
  double henrys_derivative = 1.e8 * polynomial(da, 5, temperature / co2_tscale) / (henrys_constant * co2_tscale);
  double tk = temperature + WO_TC_K;
  return -1.e3 * gas_constant * tk * tk * henrys_derivative / co2_molecular_weight;
}

/* ncg_co2_thermodynamics.F90:237-263: coefficients interpolated linearly in pressure (MPa), polynomial in T */
int wo_co2_viscosity(double partial_pressure, double temperature, double *viscosity) {
  if (!(partial_pressure <= 300.e5)) return 1;
  double vals[25], coefs[5];
  for (int i = 0; i < 5; i++)
    for (int d = 0; d < 5; d++) vals[d + 5 * i] = co2_visc_c[d][i];
  wo_table tbl;
  wo_table_init(&tbl, co2_visc_p, vals, 5, 5);
  wo_table_interpolate(&tbl, partial_pressure / 1.e6, coefs);
  wo_table_destroy(&tbl);
  *viscosity = 1.e-5 * polynomial(coefs, 5, temperature);
  return 0;
}

/* ncg_thermodynamics.F90:145-157 */
static double co2_mole_to_mass_fraction(double xmole) {
  double w = xmole * co2_molecular_weight;
  return w / (w + (1.0 - xmole) * water_molecular_weight);
}

/* ---- air (src/ncg_air_thermodynamics.F90) ---- */
static const double air_molecular_weight = 28.96;                           /* :14 */
static const double air_enthalpy_data[4] = {1.20740, 9.24502, 0.115984, -5.63568e-4};
static const double air_constituent_weight[2] = {0.79, 0.21};
static const double air_henry_p0[2] = {1.01325e5, 1.e5};
/* henry_data(2,7): coefficient k of constituent c at [c][k] */
static const double air_henry_data[2][7] = {
    {0.513726, 1.58603, -5.9378e-1, -6.98282e-1, 5.10330e-1, -1.21388e-1, 1.00041e-2},
    {0.26234, 0.610628, 7.00732e-1, -0.139299e1, 7.13850e-1, -1.54216e-1, 1.23190e-2}};
static const double air_tscale = 100.0;
static const double air_fair = 97.0, air_fwat = 363.0, air_cair = 3.617, air_cwat = 2.655;

/* ncg_air_properties (:97-121): ideal gas with deviation factor 1, enthalpy relative to the triple point */
void wo_air_properties(double partial_pressure, double temperature, double props[2]) {
  const double ttriple = 0.01;                                              /* thermodynamics.F90 */
  double tk = temperature + WO_TC_K;
  double enthalpy_shift = polynomial(air_enthalpy_data, 4, (ttriple + WO_TC_K) / air_tscale); /* ncg_air_init :84-86 */
  props[0] = partial_pressure * air_molecular_weight / (1.e3 * gas_constant * 1.0 * tk);
  props[1] = 1.e4 * (polynomial(air_enthalpy_data, 4, tk / air_tscale) - enthalpy_shift);
}

/* ncg_air_henrys_constant (:125-143): weighted sum over N2 and O2 */
double wo_air_henrys_constant(double temperature, double constituent[2]) {
  double hc = 0.0;
  for (int c = 0; c < 2; c++) {
    constituent[c] = 1.e5 * air_henry_p0[c] * polynomial(air_henry_data[c], 7, temperature / air_tscale);
    hc += air_constituent_weight[c] * constituent[c];
  }
  return hc;
}

/* ncg_air_henrys_derivative (:176-198) then ncg_energy_solution (ncg_thermodynamics.F90:187-231) */
double wo_air_energy_solution(double temperature, const double constituent_henrys_constant[2]) {
  double henrys_derivative = 0.0;
  for (int c = 0; c < 2; c++) {
    double da[6];
    for (int i = 0; i < 6; i++) da[i] = (double)(i + 1) * air_henry_data[c][i + 1];
    double dhinv = 1.e5 * polynomial(da, 6, temperature / air_tscale);
    double d = air_henry_p0[c] * dhinv / (constituent_henrys_constant[c] * air_tscale);
    henrys_derivative += air_constituent_weight[c] * d;
  }
  double tk = temperature + WO_TC_K;
  return -1.e3 * gas_constant * tk * tk * henrys_derivative / air_molecular_weight;
}

static double air_mass_to_mole_fraction(double xg) { /* ncg_thermodynamics.F90:171-183 */
  double w = xg / air_molecular_weight;
  return w / (w + (1.0 - xg) / water_molecular_weight);
}
static double air_mole_to_mass_fraction(double xmole) { /* :155-167 */
  double w = xmole * air_molecular_weight;
  return w / (w + (1.0 - xmole) * water_molecular_weight);
}
static double covis(double trd, double c, double ome, double rm, double f) {
  return 266.93e-7 * sqrt(rm * trd * f) / (c * c * ome * trd);
}

/* ncg_air_mixture_viscosity (:256-312); phase 1-based */
double wo_air_mixture_viscosity(double water_viscosity, double temperature, double xg, int phase) {
  if (phase == 1) return water_viscosity;
  double rm1 = air_molecular_weight, rm2 = water_molecular_weight;
  double fmix = sqrt(air_fair * air_fwat), cmix = 0.5 * (air_cair + air_cwat);
  double x1 = air_mass_to_mole_fraction(xg), x2 = 1.0 - x1;
  double tk = temperature + WO_TC_K;
  double trd1 = tk / air_fair, trd3 = tk / fmix;
  double ome1 = (1.188 - 0.051 * trd1) / trd1;
  double ome3 = (1.48 - 0.412 * log(trd3)) / trd3;
  double ard = 1.095 / trd3;
  double rm3 = 2.0 * rm1 * rm2 / (rm1 + rm2);
  double vis1 = covis(trd1, air_cair, ome1, rm1, air_fair);
  double vis2 = 10.0 * water_viscosity;
  double vis3 = covis(trd3, cmix, ome3, rm3, fmix);
  double z1 = x1 * x1 / vis1 + 2.0 * x2 * x1 / vis3 + x2 * x2 / vis2;
  double g = x1 * x1 * rm1 / rm2;
  double h = x2 * x2 * rm2 / rm1;
  double e = (2.0 * x1 * x2 * rm1 * rm2 / (rm3 * rm3)) * vis3 / (vis1 * vis2);
  double z2 = 0.6 * ard * (g / vis1 + e + h / vis2);
  double z3 = 0.6 * ard * (g + e * (vis1 + vis2) - 2.0 * x1 * x2 + h);
  return 0.1 * (1.0 + z3) / (z1 + z2);
}

/* eos.F90:214-236 */
static int eos_phase_composition(wo_eos *e, double *fluid) {
  int region = nint_(fluid[F_REGION]);
  int phases = wo_phase_composition(e->thermo, region, fluid[F_P], fluid[F_T]);
  if (phases > 0) {
    fluid[F_PHASES] = (double)phases;
    return 0;
  }
  return 1;
}

/* eos_we.F90:327-390 ; eos_w.F90 bulk_properties */
int wo_eos_bulk_properties(wo_eos *e, const double *primary, double *fluid) {
  int err = 0;
  int nc = e->nc;
  if (e->prm.eos == WO_EOS_W) {
    fluid[F_P] = primary[0];
    fluid[F_T] = e->prm.eos_w_temperature;
    phase_ptr(fluid, nc, 0)[PH_SAT] = 1.0;
    err = eos_phase_composition(e, fluid);
    fluid[F_PERMFAC] = 1.0;
    fluid[F_PARTIAL] = fluid[F_P];
    return err;
  }
  fluid[F_P] = primary[0];
  int region = nint_(fluid[F_REGION]);
  if (e->prm.eos == WO_EOS_WCE) { /* eos_wge.F90:350-389 */
    fluid[F_PARTIAL] = fluid[F_P] - primary[2];
    fluid[F_PARTIAL + 1] = primary[2];
    if (region == 4) err = wo_saturation_temperature(e->thermo, fluid[F_PARTIAL], &fluid[F_T]);
    else fluid[F_T] = primary[1];
    if (err == 0) {
      fluid[F_PERMFAC] = 1.0;
      err = eos_phase_composition(e, fluid);
      if (err == 0) { /* phase_saturations: eos_wge.F90:393-417 */
        double *l = phase_ptr(fluid, nc, 0), *v = phase_ptr(fluid, nc, 1);
        switch (region) {
          case 1: l[PH_SAT] = 1.0; v[PH_SAT] = 0.0; break;
          case 2: l[PH_SAT] = 0.0; v[PH_SAT] = 1.0; break;
          case 4: l[PH_SAT] = 1.0 - primary[1]; v[PH_SAT] = primary[1]; break;
          default: break;
        }
      }
    }
    return err;
  }
  if (region == 4) err = wo_saturation_temperature(e->thermo, fluid[F_P], &fluid[F_T]);
  else fluid[F_T] = primary[1];
  if (err == 0) {
    fluid[F_PERMFAC] = 1.0;
    err = eos_phase_composition(e, fluid);
    if (err == 0) {
      /* phase_saturations: eos_we.F90:366-390 */
      double *l = phase_ptr(fluid, nc, 0), *v = phase_ptr(fluid, nc, 1);
      switch (region) {
        case 1: l[PH_SAT] = 1.0; v[PH_SAT] = 0.0; break;
        case 2: l[PH_SAT] = 0.0; v[PH_SAT] = 1.0; break;
        case 4: l[PH_SAT] = 1.0 - primary[1]; v[PH_SAT] = primary[1]; break;
        default: break;
      }
      fluid[F_PARTIAL] = fluid[F_P];
    }
  }
  return err;
}

/* eos_we.F90:394-458 ; eos_w.F90 phase_properties */
int wo_eos_phase_properties(wo_eos *e, const double *primary, const double *rock, double *fluid) {
  (void)primary;
  (void)rock;
  int nc = e->nc, err = 0;
  double properties[2];
  if (e->prm.eos == WO_EOS_W) {
    int p = nint_(fluid[F_REGION]);
    double param[2] = {fluid[F_P], fluid[F_T]};
    err = wo_region_properties(e->thermo, p, param, properties);
    if (err == 0) {
      double *ph = phase_ptr(fluid, nc, p - 1);
      ph[PH_RHO] = properties[0];
      ph[PH_U] = properties[1];
      ph[PH_H] = ph[PH_U] + fluid[F_P] / ph[PH_RHO];
      ph[PH_KR] = 1.0;
      ph[PH_PC] = 0.0;
      ph[PH_X] = 1.0;
      ph[PH_MU] = wo_region_viscosity(e->thermo, p, fluid[F_T], fluid[F_P], ph[PH_RHO]);
    }
    return err;
  }
  int phases = nint_(fluid[F_PHASES]);
  double sl = phase_ptr(fluid, nc, 0)[PH_SAT];
  double relperm[2], cap[2];
  wo_relperm_values(&e->prm.relperm, sl, relperm);
  if (e->prm.eos == WO_EOS_WCE) { /* eos_wge.F90:421-543 */
    double gas_properties[2], constituent_henrys[2];
    if (e->gas == 1) wo_air_properties(fluid[F_PARTIAL + 1], fluid[F_T], gas_properties);
    else wo_co2_properties(fluid[F_PARTIAL + 1], fluid[F_T], gas_properties);
    for (int p = 0; p < e->nphase; p++) {
      double *ph = phase_ptr(fluid, nc, p);
      if (phases & (1 << p)) {
        double water_pressure, capillary_pressure, henrys_constant, energy_solution;
        if (p == 0) {
          water_pressure = fluid[F_P];
          capillary_pressure = wo_cappress_value(&e->prm.cappress, sl, fluid[F_T]);
          if (e->gas == 1) {
            henrys_constant = wo_air_henrys_constant(fluid[F_T], constituent_henrys);
            energy_solution = wo_air_energy_solution(fluid[F_T], constituent_henrys);
          } else {
            henrys_constant = wo_co2_henrys_constant(fluid[F_T]);
            energy_solution = wo_co2_energy_solution(fluid[F_T], henrys_constant);
          }
        } else {
          water_pressure = fluid[F_PARTIAL];
          capillary_pressure = 0.0;
          henrys_constant = 0.0;
          energy_solution = 0.0;
        }
        double param[2] = {water_pressure, fluid[F_T]};
        err = wo_region_properties(e->thermo, p + 1, param, properties);
        if (err) break;
        /* effective_properties: no free gas density in the liquid phase (ncg_thermodynamics.F90:315-340) */
        double gas_density = (p == 0) ? 0.0 : gas_properties[0], gas_enthalpy = gas_properties[1];
        double water_density = properties[0], water_internal_energy = properties[1];
        /* mass_fraction: ncg_thermodynamics.F90:279-311 */
        double xg;
        if (p == 0) {
          double xmole = fluid[F_PARTIAL + 1] / henrys_constant;
          xg = (e->gas == 1) ? air_mole_to_mass_fraction(xmole) : co2_mole_to_mass_fraction(xmole);
        } else {
          double total_density = gas_density + water_density;
          xg = (total_density < 1.e-30) ? 0.0 : gas_density / total_density;
        }
        double water_viscosity = wo_region_viscosity(e->thermo, p + 1, fluid[F_T], fluid[F_P], water_density);
        /* mixture_viscosity: ncg_co2_thermodynamics.F90:267-292 / ncg_air_thermodynamics.F90:256-312 */
        if (e->gas == 1) {
          ph[PH_MU] = wo_air_mixture_viscosity(water_viscosity, fluid[F_T], xg, p + 1);
        } else if (p == 0) {
          ph[PH_MU] = water_viscosity;
        } else {
          double gas_viscosity;
          err = wo_co2_viscosity(fluid[F_PARTIAL + 1], fluid[F_T], &gas_viscosity);
          if (err) break;
          ph[PH_MU] = water_viscosity * (1.0 - xg) + gas_viscosity * xg;
        }
        ph[PH_RHO] = water_density + gas_density;
        ph[PH_X] = 1.0 - xg;
        ph[PH_X + 1] = xg;
        ph[PH_KR] = relperm[p];
        ph[PH_PC] = capillary_pressure;
        double water_enthalpy = water_internal_energy + water_pressure / water_density;
        ph[PH_H] = water_enthalpy * (1.0 - xg) + (gas_enthalpy + energy_solution) * xg;
        ph[PH_U] = ph[PH_H] - fluid[F_P] / ph[PH_RHO];
      } else {
        ph[PH_RHO] = 0.0; ph[PH_U] = 0.0; ph[PH_H] = 0.0; ph[PH_KR] = 0.0;
        ph[PH_PC] = 0.0; ph[PH_MU] = 0.0; ph[PH_X] = 0.0; ph[PH_X + 1] = 0.0;
      }
    }
    return err;
  }
  cap[0] = wo_cappress_value(&e->prm.cappress, sl, fluid[F_T]);
  cap[1] = 0.0;
  for (int p = 0; p < e->nphase; p++) {
    double *ph = phase_ptr(fluid, nc, p);
    if (phases & (1 << p)) {
      double param[2] = {fluid[F_P], fluid[F_T]};
      err = wo_region_properties(e->thermo, p + 1, param, properties);
      if (err == 0) {
        ph[PH_RHO] = properties[0];
        ph[PH_U] = properties[1];
        ph[PH_H] = ph[PH_U] + fluid[F_P] / ph[PH_RHO];
        ph[PH_X] = 1.0;
        ph[PH_KR] = relperm[p];
        ph[PH_PC] = cap[p];
        ph[PH_MU] = wo_region_viscosity(e->thermo, p + 1, fluid[F_T], fluid[F_P], ph[PH_RHO]);
      } else
        break;
    } else {
      ph[PH_RHO] = 0.0;
      ph[PH_U] = 0.0;
      ph[PH_H] = 0.0;
      ph[PH_KR] = 0.0;
      ph[PH_PC] = 0.0;
      ph[PH_MU] = 0.0;
      ph[PH_X] = 0.0;
    }
  }
  return err;
}

/* ---- transitions: eos_we.F90:149-323 ---- */

typedef struct {
  const wo_thermo *thermo;
  double v0[WO_MAX_NP], v1[WO_MAX_NP]; /* interpolator val(:,1), val(:,2) on coord [0,1] */
} satline_ctx;

/* interpolate_at_index with index fixed to 1 on the 2-point table x=[0,1]
   (interpolation.F90:388-403,494-510; eos_we.F90:100-104,181-183) */
static void pv_interp(const satline_ctx *c, int np, double x, double *y) {
  double xi = (x - 0.0) / (1.0 - 0.0);
  for (int i = 0; i < np; i++) y[i] = (1.0 - xi) * c->v0[i] + xi * c->v1[i];
}

/* eos_wge.F90:678-701 */
static double wge_saturation_difference(double x, void *ctx) {
  satline_ctx *c = (satline_ctx *)ctx;
  double var[WO_MAX_NP], Ps = 0.0;
  pv_interp(c, 3, x, var);
  wo_saturation_pressure(c->thermo, var[1], &Ps);
  return var[0] - var[2] - Ps;
}

/* eos_wge.F90:149-227 */
static int wge_transition_to_single_phase(wo_eos *e, const double *old_primary, const double *old_fluid,
                                          int new_region, double *primary, double *fluid, int *transition) {
  const double small = 1.e-6;
  int err = 0;
  *transition = 0;
  double saturation_bound = (new_region == 1) ? 0.0 : 1.0;
  double pressure_factor = (new_region == 1) ? 1.0 + small : 1.0 - small;
  primary[2] = fmax(0.0, fmin(primary[2], primary[0]));
  double xs[2] = {0.0, 1.0}, vals[2 * WO_MAX_NP];
  for (int i = 0; i < 3; i++) {
    vals[i] = old_primary[i];
    vals[3 + i] = primary[i];
  }
  wo_table tbl;
  wo_table_init(&tbl, xs, vals, 2, 3);
  tbl.index = 1;
  double xi = 0.0;
  err = wo_table_find_component_at_index(&tbl, saturation_bound, 2, &xi);
  if (err == 0) {
    double ip[WO_MAX_NP];
    wo_table_interpolate(&tbl, xi, ip);
    double interpolated_water_pressure = ip[0] - ip[2];
    primary[0] = pressure_factor * interpolated_water_pressure + ip[2];
    primary[2] = ip[2];
    err = wo_saturation_temperature(e->thermo, interpolated_water_pressure, &primary[1]);
    if (err == 0) {
      fluid[F_REGION] = (double)new_region;
      *transition = 1;
    }
  } else {
    double old_ps;
    err = wo_saturation_pressure(e->thermo, old_fluid[F_T], &old_ps);
    if (err == 0) {
      primary[0] = pressure_factor * old_ps + primary[2];
      primary[1] = old_fluid[F_T];
      fluid[F_REGION] = (double)new_region;
      *transition = 1;
    }
  }
  wo_table_destroy(&tbl);
  return err;
}

/* eos_wge.F90:231-285 */
static int wge_transition_to_two_phase(wo_eos *e, double saturation_pressure, const double *old_primary,
                                       const double *old_fluid, double *primary, double *fluid, int *transition) {
  const double small = 1.e-6;
  primary[2] = fmax(0.0, fmin(primary[2], primary[0]));
  satline_ctx c;
  c.thermo = e->thermo;
  for (int i = 0; i < 3; i++) {
    c.v0[i] = old_primary[i];
    c.v1[i] = primary[i];
  }
  wo_root_finder rf;
  wo_root_finder_init(&rf);
  wo_root_finder_find(&rf, wge_saturation_difference, &c);
  if (rf.err == 0) {
    double xs[2] = {0.0, 1.0}, vals[6] = {c.v0[0], c.v0[1], c.v0[2], c.v1[0], c.v1[1], c.v1[2]}, ip[3];
    wo_table tbl;
    wo_table_init(&tbl, xs, vals, 2, 3);
    wo_table_interpolate(&tbl, rf.root, ip);
    wo_table_destroy(&tbl);
    primary[0] = ip[0];
    primary[2] = ip[2];
  } else {
    primary[0] = saturation_pressure + primary[2];
  }
  int old_region = nint_(old_fluid[F_REGION]);
  primary[1] = (old_region == 1) ? small : 1.0 - small;
  fluid[F_REGION] = 4.0;
  *transition = 1;
  return 0;
}

/* eos_we.F90:530-553 */
static double saturation_difference(double x, void *ctx) {
  satline_ctx *c = (satline_ctx *)ctx;
  double var[WO_MAX_NP], Ps = 0.0;
  pv_interp(c, 2, x, var);
  wo_saturation_pressure(c->thermo, var[1], &Ps);
  return var[0] - Ps;
}

/* eos_we.F90:149-216 */
static int we_transition_to_single_phase(wo_eos *e, const double *old_primary, const double *old_fluid,
                                         int new_region, double *primary, double *fluid, int *transition) {
  const double small = 1.e-6;
  int err = 0;
  *transition = 0;
  double saturation_bound, pressure_factor;
  if (new_region == 1) {
    saturation_bound = 0.0;
    pressure_factor = 1.0 + small;
  } else {
    saturation_bound = 1.0;
    pressure_factor = 1.0 - small;
  }
  /* 2-point table on x = [0, 1], index set to 1; find xi where component 2 == bound */
  double xs[2] = {0.0, 1.0};
  double vals[2 * WO_MAX_NP];
  for (int i = 0; i < 2; i++) {
    vals[i] = old_primary[i];
    vals[2 + i] = primary[i];
  }
  wo_table tbl;
  wo_table_init(&tbl, xs, vals, 2, 2);
  tbl.index = 1;
  double xi = 0.0;
  err = wo_table_find_component_at_index(&tbl, saturation_bound, 2, &xi);
  if (err == 0) {
    /* interpolate(xi): find + interpolate_at_index (interpolation.F90:533-545) */
    double interpolated[WO_MAX_NP];
    wo_table_interpolate(&tbl, xi, interpolated);
    primary[0] = pressure_factor * interpolated[0];
    err = wo_saturation_temperature(e->thermo, interpolated[0], &primary[1]);
    if (err == 0) {
      fluid[F_REGION] = (double)new_region;
      *transition = 1;
    }
  } else {
    double old_ps;
    err = wo_saturation_pressure(e->thermo, old_fluid[F_T], &old_ps);
    if (err == 0) {
      primary[0] = pressure_factor * old_ps;
      primary[1] = old_fluid[F_T];
      fluid[F_REGION] = (double)new_region;
      *transition = 1;
    }
  }
  wo_table_destroy(&tbl);
  return err;
}

/* eos_we.F90:220-268 */
static int we_transition_to_two_phase(wo_eos *e, double saturation_pressure, const double *old_primary,
                                      const double *old_fluid, double *primary, double *fluid,
                                      int *transition) {
  const double small = 1.e-6;
  satline_ctx c;
  c.thermo = e->thermo;
  for (int i = 0; i < 2; i++) {
    c.v0[i] = old_primary[i];
    c.v1[i] = primary[i];
  }
  wo_root_finder rf;
  wo_root_finder_init(&rf);
  wo_root_finder_find(&rf, saturation_difference, &c);
  if (rf.err == 0) {
    /* interpolate(xi) = find + interpolate_at_index: clamps outside [0,1] */
    double xs[2] = {0.0, 1.0}, vals[4] = {c.v0[0], c.v0[1], c.v1[0], c.v1[1]}, ip[2];
    wo_table tbl;
    wo_table_init(&tbl, xs, vals, 2, 2);
    wo_table_interpolate(&tbl, rf.root, ip);
    wo_table_destroy(&tbl);
    primary[0] = ip[0];
  } else {
    primary[0] = saturation_pressure;
  }
  int old_region = nint_(old_fluid[F_REGION]);
  primary[1] = (old_region == 1) ? small : 1.0 - small;
  fluid[F_REGION] = 4.0;
  *transition = 1;
  return 0;
}

/* eos_we.F90:272-323 */
int wo_eos_transition(wo_eos *e, const double *old_primary, double *primary, const double *old_fluid,
                      double *fluid, int *transition) {
  int err = 0;
  *transition = 0;
  if (e->prm.eos == WO_EOS_W) return 0;
  int old_region = nint_(old_fluid[F_REGION]);
  if (e->prm.eos == WO_EOS_WCE) { /* eos_wge.F90:289-346 */
    if (old_region == 4) {
      double sv = primary[1];
      if (sv < 0.0) err = wge_transition_to_single_phase(e, old_primary, old_fluid, 1, primary, fluid, transition);
      else if (sv > 1.0) err = wge_transition_to_single_phase(e, old_primary, old_fluid, 2, primary, fluid, transition);
    } else {
      double ps;
      err = wo_saturation_pressure(e->thermo, primary[1], &ps);
      if (err == 0) {
        double water_pressure = primary[0] - primary[2];
        if ((old_region == 1 && water_pressure < ps) || (old_region == 2 && water_pressure > ps))
          err = wge_transition_to_two_phase(e, ps, old_primary, old_fluid, primary, fluid, transition);
      }
    }
    return err;
  }
  if (old_region == 4) {
    double sv = primary[1];
    if (sv < 0.0) err = we_transition_to_single_phase(e, old_primary, old_fluid, 1, primary, fluid, transition);
    else if (sv > 1.0) err = we_transition_to_single_phase(e, old_primary, old_fluid, 2, primary, fluid, transition);
  } else {
    double ps;
    err = wo_saturation_pressure(e->thermo, primary[1], &ps);
    if (err == 0) {
      if ((old_region == 1 && primary[0] < ps) || (old_region == 2 && primary[0] > ps))
        err = we_transition_to_two_phase(e, ps, old_primary, old_fluid, primary, fluid, transition);
    }
  }
  return err;
}

/* eos_we.F90:486-526 ; eos_w.F90 check_primary_variables */
int wo_eos_check_primary_variables(const wo_eos *e, const double *fluid, double *primary, int *changed) {
  *changed = 0;
  if (e->prm.eos == WO_EOS_WCE) { /* eos_wge.F90:573-635: clamps the gas partial pressure */
    const double small = 1.e-6;
    if (!(primary[0] > 0.0)) return 1;
    double max_pp = (1.0 - small) * primary[0];
    if (primary[2] > max_pp) {
      primary[2] = max_pp;
      *changed = 1;
    } else if (primary[2] < 0.0) {
      primary[2] = 0.0;
      *changed = 1;
    }
    double pw = primary[0] - primary[2];
    if (pw > 100.e6) return 1;
    if (nint_(fluid[F_REGION]) == 4) {
      if (primary[1] < -1.0 || primary[1] > 2.0) return 1;
    } else {
      if (primary[1] < 0.0 || primary[1] > 800.0) return 1;
    }
    return 0;
  }
  double p = primary[0];
  if (p < 0.0 || p > 100.e6) return 1;
  if (e->prm.eos == WO_EOS_W) return 0;
  int region = nint_(fluid[F_REGION]);
  if (region == 4) {
    double sv = primary[1];
    if (sv < -1.0 || sv > 2.0) return 1;
  } else {
    double t = primary[1];
    if (t < 0.0 || t > 800.0) return 1;
  }
  return 0;
}

/* eos.F90:240-257 */
double wo_eos_conductivity(const double *rock, const double *fluid, int nc) {
  double sl = cphase_ptr(fluid, nc, 0)[PH_SAT];
  return rock[R_DRY] + sqrt(sl) * (rock[R_WET] - rock[R_DRY]);
}

/* ---- cell balance: cell.F90:114-142, fluid.F90:295-370, rock.F90:142 ---- */
void wo_cell_balance(const double *rock, const double *fluid, int nc, int nphase, int np, double *balance) {
  int isothermal = (np == nc);
  double d[WO_MAX_NC];
  for (int c = 0; c < nc; c++) d[c] = 0.0;
  for (int p = 0; p < nphase; p++) {
    const double *ph = cphase_ptr(fluid, nc, p);
    double ds = ph[PH_RHO] * ph[PH_SAT];
    for (int c = 0; c < nc; c++) d[c] = d[c] + ds * ph[PH_X + c];
  }
  for (int c = 0; c < nc; c++) balance[c] = rock[R_POR] * d[c];
  if (!isothermal) {
    double er = rock[R_RHO] * rock[R_CP] * fluid[F_T];
    double ef = 0.0;
    for (int p = 0; p < nphase; p++) {
      const double *ph = cphase_ptr(fluid, nc, p);
      double ds = ph[PH_RHO] * ph[PH_SAT];
      ef = ef + ds * ph[PH_U];
    }
    balance[np - 1] = rock[R_POR] * ef + (1.0 - rock[R_POR]) * er;
  }
}

/* ---- face: face.F90 ---- */
enum { G_AREA = 0, G_DIST = 1, G_DIST12 = 3, G_NORMAL = 4, G_GRAVN = 7, G_CENTROID = 8, G_PERMDIR = 11 };

/* face.F90:230-249 */
void wo_face_calculate_distances(const double *c1, const double *c2, const double *fc, const double *normal,
                                 double dist[2], double *dist12) {
  double d1 = 0, d2 = 0, d12 = 0;
  for (int i = 0; i < 3; i++) {
    d1 += (fc[i] - c1[i]) * normal[i];
    d2 += (c2[i] - fc[i]) * normal[i];
    d12 += (c2[i] - c1[i]) * normal[i];
  }
  double correction = d12 / (d1 + d2);
  dist[0] = d1 * correction;
  dist[1] = d2 * correction;
  *dist12 = d12;
}

/* face.F90:358-377 */
double wo_face_harmonic_average(const double *g, const double x[2]) {
  const double tol = 1.e-30;
  double wx = (g[G_DIST] * x[1] + g[G_DIST + 1] * x[0]) / g[G_DIST12];
  if (fabs(wx) > tol) return x[0] * x[1] / wx;
  return 0.0;
}

/* face.F90:443-515 */
void wo_face_flux(const double *g, const double *rock1, const double *rock2, const double *fluid1,
                  const double *fluid2, int nc, int np, int nphase, int nmobile, int isothermal,
                  double *flux) {
  (void)nphase;
  const double *rock[2] = {rock1, rock2};
  const double *fluid[2] = {fluid1, fluid2};
  int nf = np + nmobile;
  for (int i = 0; i < nf; i++) flux[i] = 0.0;
  /* permeability: face.F90:381-398 */
  int direction = nint_(g[G_PERMDIR]);
  double perm[2];
  for (int i = 0; i < 2; i++) perm[i] = rock[i][R_PERM + direction - 1] * fluid[i][F_PERMFAC];
  double k = wo_face_harmonic_average(g, perm);
  if (!isothermal) {
    double kcell[2], t[2];
    for (int i = 0; i < 2; i++) {
      kcell[i] = wo_eos_conductivity(rock[i], fluid[i], nc);
      t[i] = fluid[i][F_T];
    }
    double cond = wo_face_harmonic_average(g, kcell);
    double dtdn = (t[1] - t[0]) / g[G_DIST12];
    flux[np - 1] = -cond * dtdn;
  }
  int phases[2];
  for (int i = 0; i < 2; i++) phases[i] = nint_(fluid[i][F_PHASES]);
  int phase_present = phases[0] | phases[1];
  double *phase_flux = flux + np;
  for (int p = 0; p < nmobile; p++) {
    if (phase_present & (1 << p)) {
      /* phase_density: face.F90:334-354 */
      double rho = 0.0, weight = 0.0;
      for (int i = 0; i < 2; i++) {
        const double *ph = cphase_ptr(fluid[i], nc, p);
        rho = rho + ph[PH_SAT] * ph[PH_RHO];
        weight = weight + ph[PH_SAT];
      }
      rho = rho / weight;
      /* pressure_gradient: face.F90:296-313 */
      double pr[2];
      for (int i = 0; i < 2; i++) pr[i] = fluid[i][F_P] + cphase_ptr(fluid[i], nc, p)[PH_PC];
      double dpdn = (pr[1] - pr[0]) / g[G_DIST12];
      double G = dpdn - rho * g[G_GRAVN];
      int up = (G <= 0.0) ? 0 : 1; /* face.F90:426-439 */
      if (phases[up] & (1 << p)) {
        const double *ups = cphase_ptr(fluid[up], nc, p);
        double mobility = ups[PH_KR] * ups[PH_RHO] / ups[PH_MU]; /* fluid.F90:197-207 */
        double F = -k * mobility * G;
        double sum = 0.0;
        for (int c = 0; c < nc; c++) {
          double pcf = F * ups[PH_X + c];
          flux[c] = flux[c] + pcf;
          sum += pcf;
        }
        if (!isothermal) {
          double h = ups[PH_H];
          flux[np - 1] = flux[np - 1] + h * F;
        }
        phase_flux[p] = sum;
      }
    }
  }
}
