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
  // (0 both, 1 production, 2 injection), bit 3 = recharge / injectivity (pi = its coefficient); productivity index,
  // reference pressure, total-flow limit (<= 0: none)
  const int32_t *ctrl;
  const double *pi, *pref, *limit;
  // separators and limiters on the separated flows (null: none): stages per source (0..2), reference water / steam
  // enthalpies of the stages [4 per source], limits on the separated water and steam rates (<= 0: none)
  const int32_t *sep_n;
  const double *sep_h, *limit_w, *limit_s;
  // reference pressure of a source on deliverability as a table against the flowing enthalpy or the pressure of its cell
  // (null: none): per source a word (bits 0-7 number of points, 0: no table; bit 8: the coordinate is the pressure;
  // bit 9: step interpolation) and WB_PTAB_MAX (coordinate, value) pairs
  const int32_t *ptab_n;
  const double *ptab;
};

#define WB_PTAB_MAX 8

// interpolation_table%interpolate (src/interpolation.F90): linear between the data points or the value of the point at or
// before x (step); constant beyond both ends
WB_HD double wb_table_interpolate(const double *tab, int n, bool step, double x) {
  if (!(x > tab[0])) return tab[1];
  if (!(x < tab[2 * (n - 1)])) return tab[2 * (n - 1) + 1];
  int i = 0;
  while (i + 2 < n && !(x < tab[2 * (i + 1)])) i++;
  const double x0 = tab[2 * i], y0 = tab[2 * i + 1], x1 = tab[2 * i + 2], y1 = tab[2 * i + 3];
  if (step) return y0;
  return y0 + (y1 - y0) * ((x - x0) / (x1 - x0));
}

// separator_separate (src/separator.F90:212-260) over separator_stage_separate (:140-166): separated water and steam
// mass rates of a flow of `rate` at `enthalpy` through nstage <= 2 flash stages with the reference enthalpies stage_h
WB_HD void wb_separate(int nstage, const double *stage_h, double rate, double enthalpy, double &water_rate,
                       double &water_enthalpy, double &steam_rate, double &steam_enthalpy) {
  double q = rate, h = enthalpy, total_steam_mass_rate = 0.0, total_steam_energy_rate = 0.0;
  for (int i = 0; i < nstage; i++) {
    const double ref_water_enthalpy = stage_h[2 * i], ref_steam_enthalpy = stage_h[2 * i + 1];
    double steam_fraction, stage_water_enthalpy, stage_steam_enthalpy;
    if (h <= ref_water_enthalpy) {
      steam_fraction = 0.0;
      stage_water_enthalpy = h;
      stage_steam_enthalpy = 0.0;
    } else if (h <= ref_steam_enthalpy) {
      steam_fraction = (h - ref_water_enthalpy) / (ref_steam_enthalpy - ref_water_enthalpy);
      stage_water_enthalpy = ref_water_enthalpy;
      stage_steam_enthalpy = ref_steam_enthalpy;
    } else {
      steam_fraction = 1.0;
      stage_water_enthalpy = 0.0;
      stage_steam_enthalpy = h;
    }
    const double stage_water_rate = (1.0 - steam_fraction) * q, stage_steam_rate = steam_fraction * q;
    total_steam_mass_rate = total_steam_mass_rate + stage_steam_rate;
    total_steam_energy_rate = total_steam_energy_rate + stage_steam_rate * stage_steam_enthalpy;
    q = stage_water_rate;
    h = stage_water_enthalpy;
  }
  water_rate = q;
  water_enthalpy = h;
  steam_rate = total_steam_mass_rate;
  steam_enthalpy = fabs(total_steam_mass_rate) > 1.e-9 ? total_steam_energy_rate / total_steam_mass_rate : 0.0;
}

// separated flows of source k at the given rate (source_network_node_get_separated_flows, src/source_network_node.F90:
// 116-131; the enthalpy is that of the fluid the source takes from its cell, src/source_network.F90:197-216): zero
// unless the source produces and has a separator.  out: water rate, water enthalpy, steam rate, steam enthalpy, fraction
template <int NC, int NPH>
WB_HD void wb_source_separated(const WbSources &S, int k, const WbCellState<NC, NPH> &s, double rate, double *out) {
  for (int i = 0; i < 5; i++) out[i] = 0.0;
  if (!S.sep_n || S.sep_n[k] <= 0 || !(rate < 0.0)) return;
  double frac[NPH], sum = 0.0, h = 0.0;
#pragma unroll
  for (int p = 0; p < NPH; p++) {
    frac[p] = 0.0;
    if (s.phases & (1 << p)) frac[p] = s.mob[p];
    sum += frac[p];
  }
#pragma unroll
  for (int p = 0; p < NPH; p++) {
    frac[p] = frac[p] / sum;
    if (s.phases & (1 << p)) h = h + frac[p] * s.h[p];
  }
  wb_separate(S.sep_n[k], S.sep_h + 4 * (size_t)k, rate, h, out[0], out[1], out[2], out[3]);
  out[4] = fabs(rate) > 1.e-9 ? out[2] / rate : 0.0;
}

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
    double reference_pressure = S.pref[k];
    if (S.ptab_n && (S.ptab_n[k] & 255) > 0) {
      // SRC_PRESSURE_TABLE_COORD_ENTHALPY / _PRESSURE (src/source_control.F90:376-388): the table is looked up at the
      // enthalpy of the fluid the source takes (phase enthalpies weighted by the flow fractions) or at the pressure
      const int word = S.ptab_n[k];
      double x = s.P;
      if (!(word & 256)) {
        double sum = 0.0;
        x = 0.0;
#pragma unroll
        for (int p = 0; p < NPH; p++)
          if (s.phases & (1 << p)) sum += s.mob[p];
#pragma unroll
        for (int p = 0; p < NPH; p++)
          if (s.phases & (1 << p)) x = x + (s.mob[p] / sum) * s.h[p];
      }
      reference_pressure = wb_table_interpolate(S.ptab + 2 * WB_PTAB_MAX * (size_t)k, word & 255, (word & 512) != 0, x);
    }
    const double pressure_difference = s.P - reference_pressure;
    rate = 0.0;
#pragma unroll
    for (int p = 0; p < NPH; p++)
      if (s.phases & (1 << p)) rate = rate - effective_productivity * s.mob[p] * pressure_difference;
  }
  if (ctrl & 8) {  // recharge / injectivity (recharge_source_control_iterator, src/source_control.F90:554-577)
    const double pressure_difference = s.P - S.pref[k];
    rate = -S.pi[k] * pressure_difference;
  }
  const int direction = (ctrl >> 1) & 3;
  if (direction == 1 && !(rate < 0.0)) rate = 0.0;
  if (direction == 2 && !(rate > 0.0)) rate = 0.0;
  // limiter (source_network_node_limit_rate, src/source_network_node.F90:245-315) over the limited flow types -- total,
  // separated water, separated steam: the smallest scale that brings every rate over its limit back to it
  double scale = 1.0, sep[5];
  bool over = false;
  const bool sep_limits = S.sep_n && (S.limit_w[k] > 0.0 || S.limit_s[k] > 0.0);
  if (sep_limits) wb_source_separated(S, k, s, rate, sep);
  for (int type = 0; type < 3; type++) {
    const double limit = type == 0 ? S.limit[k] : (sep_limits ? (type == 1 ? S.limit_w[k] : S.limit_s[k]) : 0.0);
    if (!(limit > 0.0)) continue;
    const double abs_rate = fabs(type == 0 ? rate : (type == 1 ? sep[0] : sep[2]));
    if (abs_rate > limit) {
      over = true;
      if (abs_rate > 1.e-6) scale = fmin(scale, limit / abs_rate);
    }
  }
  if (over) rate = rate * scale;
  return rate;
}

