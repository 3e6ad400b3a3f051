// wb_eos.cuh -- per-cell equation of state, curves, cell balance, face flux
// and phase transitions as stateless device functions.
//
// What is computed follows the reference (file:line cited per function); how
// it is organised does not: the reference walks AoS fluid records through
// class(eos_type) pointers, one cell at a time.  Here a cell's properties are
// a register-resident struct (WbFluid) produced from (scaled primaries, region)
// alone, and the face flux consumes a compact WbCellState -- only the ten or so
// numbers a two-point flux needs -- so the hot kernels never touch the
// 23-double record unless the caller asks for it.
#pragma once
#include "../../include/waiwera_b200.h"
#include "wb_thermo.cuh"

// ---------------------------------------------------------------- parameters

struct WbEosParams {
  int eos, np, nc, nphase;
  WbThermo thermo;
  double scale[WB_MAX_NP][5];  // primary_scale(var, region 1..4), src/eos_we.F90:104-109
  int adaptive_pp;             // eos_wce: gas partial pressure scaled by the cell's total pressure
  int gas;                     // eos_wge family: 0 CO2 (eos_wce), 1 air (eos_wae)
  double eos_w_temperature;
  wb_relperm relperm;
  wb_cappress cappress;
};

inline int wb_eos_params_make(const wb_params &prm, WbEosParams &e) {
  e = WbEosParams();
  e.eos = prm.eos;
  if (prm.eos == WB_EOS_WAE) {  // src/eos_wae.F90:27-64: eos_wge with the air NCG; same kernels as eos_wce
    e.eos = WB_EOS_WCE;
    e.gas = 1;
  }
  e.thermo = wb_thermo_make(prm.thermo, prm.extrapolate);
  e.eos_w_temperature = prm.eos_w_temperature;
  e.relperm = prm.relperm;
  e.cappress = prm.cappress;
  const double ps = prm.pressure_scale > 0 ? prm.pressure_scale : 1.e6;  // src/eos_we.F90:75-76
  const double ts = prm.temperature_scale > 0 ? prm.temperature_scale : 1.e2;
  if (e.eos == WB_EOS_WE) {  // src/eos_we.F90:78-109
    e.np = 2; e.nc = 1; e.nphase = 2;
    e.scale[0][1] = ps; e.scale[1][1] = ts;
    e.scale[0][2] = ps; e.scale[1][2] = ts;
    e.scale[0][4] = ps; e.scale[1][4] = 1.0;
  } else if (e.eos == WB_EOS_W) {  // src/eos_w.F90:67-96
    e.np = 1; e.nc = 1; e.nphase = 1;
    e.scale[0][1] = ps; e.scale[0][2] = ps;
  } else if (e.eos == WB_EOS_WCE) {  // src/eos_wge.F90:40-131, src/eos_wce.F90:23-53
    e.np = 3; e.nc = 2; e.nphase = 2;
    double pps = prm.partial_pressure_scale;
    e.adaptive_pp = !(pps > 0.0);
    if (e.adaptive_pp) pps = 0.0;
    e.scale[0][1] = ps; e.scale[1][1] = ts; e.scale[2][1] = pps;
    e.scale[0][2] = ps; e.scale[1][2] = ts; e.scale[2][2] = pps;
    e.scale[0][4] = ps; e.scale[1][4] = 1.0; e.scale[2][4] = pps;
  } else {
    return 1;
  }
  return 0;
}

template <int EOS> struct WbEosTraits;
template <> struct WbEosTraits<WB_EOS_WE> { static constexpr int NP = 2, NC = 1, NPH = 2; };
template <> struct WbEosTraits<WB_EOS_W> { static constexpr int NP = 1, NC = 1, NPH = 1; };
template <> struct WbEosTraits<WB_EOS_WCE> { static constexpr int NP = 3, NC = 2, NPH = 2; };

// ---------------------------------------------------------------- curves

// 2-point / n-point linear table with end clamping: find + interpolate_at_index
// (src/interpolation.F90:202-306, 388-403, 494-510).  The reference's hunt cache
// only changes how the bracket is found, not which bracket: a plain scan returns
// the same interval.
WB_HD double wb_table_interp(const double *xs, const double *ys, int n, double x) {
  if (x <= xs[0]) return ys[0];
  if (x >= xs[n - 1]) return ys[n - 1];
  int i = 0;
  while (i + 2 < n && x >= xs[i + 1]) i++;
  const double xi = (x - xs[i]) / (xs[i + 1] - xs[i]);
  return (1.0 - xi) * ys[i] + xi * ys[i + 1];
}

WB_HD double wb_table2(double x0, double x1, double y0, double y1, double x) {
  if (x <= x0) return y0;
  if (x >= x1) return y1;
  const double xi = (x - x0) / (x1 - x0);
  return (1.0 - xi) * y0 + xi * y1;
}

// src/relative_permeability.F90:197-558
WB_HD void wb_relperm_values(const wb_relperm &rp, double sl, double &krl, double &krv) {
  switch (rp.type) {
    case WB_RP_FULLY_MOBILE:
      krl = 1.0; krv = 1.0;
      break;
    case WB_RP_LINEAR:  // :213-259
      krl = wb_table2(rp.p[0], rp.p[1], 0.0, 1.0, sl);
      krv = wb_table2(rp.p[2], rp.p[3], 0.0, 1.0, 1.0 - sl);
      break;
    case WB_RP_PICKENS:  // :297-308
      krl = pow(sl, rp.p[0]);
      krv = 1.0;
      break;
    case WB_RP_COREY:
    case WB_RP_GRANT: {  // :349-370, :399-420
      const double slr = rp.p[0], ssr = rp.p[1];
      const double sv = 1.0 - sl;
      if (sv < ssr) {
        krl = 1.0; krv = 0.0;
      } else if (sv > 1.0 - slr) {
        krl = 0.0; krv = 1.0;
      } else {
        const double sstar = (sl - slr) / (1.0 - slr - ssr);
        const double sstar2 = sstar * sstar;
        krl = sstar2 * sstar2;
        krv = rp.type == WB_RP_COREY ? (1.0 - 2.0 * sstar + sstar2) * (1.0 - sstar2) : 1.0 - krl;
      }
      break;
    }
    case WB_RP_VAN_GENUCHTEN: {  // :461-491
      const double lambda = rp.p[0], slr = rp.p[1], sls = rp.p[2], ssr = rp.p[4];
      const double sstar = (sl - slr) / (sls - slr);
      if (sstar < 0.0) krl = 0.0;
      else if (sstar < 1.0) {
        const double b = 1.0 - pow(1.0 - pow(sstar, 1.0 / lambda), lambda);
        krl = sqrt(sstar) * (b * b);
      } else
        krl = 1.0;
      if (rp.p[3] != 0.0) krv = 1.0 - krl;
      else {
        const double s_hat = (sl - slr) / (1.0 - slr - ssr);
        const double s_hat2 = s_hat * s_hat;
        krv = fmin(1.0, (1.0 - 2.0 * s_hat + s_hat2) * (1.0 - s_hat2));
      }
      break;
    }
    case WB_RP_TABLE:  // :547-558
      krl = wb_table_interp(rp.lx, rp.ly, rp.nl, sl);
      krv = wb_table_interp(rp.vx, rp.vy, rp.nv, 1.0 - sl);
      break;
    default:
      krl = 0.0; krv = 0.0;
  }
}

// src/capillary_pressure.F90:159-358
WB_HD double wb_cappress_value(const wb_cappress &cp, double sl, double t) {
  (void)t;
  switch (cp.type) {
    case WB_CP_LINEAR:  // :176-218
      return wb_table2(cp.p[0], cp.p[1], -fabs(cp.p[2]), 0.0, sl);
    case WB_CP_VAN_GENUCHTEN: {  // :273-305
      const double eps = 1.e-3;
      const double P0 = fabs(cp.p[0]), lambda = cp.p[1], slr = cp.p[2], sls = cp.p[3], Pmax = fabs(cp.p[4]);
      double c;
      if (sl < 1.0) {
        const double sstar = (sl - slr) / (sls - slr);
        if (sstar < 0.0) c = -Pmax;
        else if (sstar < 1.0) c = -P0 * pow(pow(sstar, -1.0 / lambda) - 1.0, 1.0 - lambda);
        else c = 0.0;
        c = fmin(0.0, c);
        if (cp.p[5] != 0.0) c = fmax(-Pmax, c);
        if (sl > 1.0 - eps) c = c * (1.0 - sl) / eps;
      } else
        c = 0.0;
      return c;
    }
    case WB_CP_TABLE:  // :349-358
      return wb_table_interp(cp.x, cp.y, cp.n, sl);
    default:  // zero :159
      return 0.0;
  }
}

// ---------------------------------------------------------------- fluid

template <int NC, int NPH> struct WbFluid {
  double P, T;
  int region, phases;
  double pp[NC];  // partial pressures
  struct Phase { double rho, mu, sat, kr, pc, h, u, X[NC]; } ph[NPH];
};

// what a two-point flux and a cell balance need from a cell (WE: 14 numbers)
template <int NC, int NPH> struct WbCellState {
  double P, T, cond;
  int phases;
  double rho[NPH], sat[NPH], pc[NPH], mob[NPH], h[NPH];
  double X[NPH][NC];
};
template <int NC, int NPH> struct WbStateLayout {
  // SoA field count of WbCellState in global memory (phases stored as a double)
  static constexpr int NF = 4 + NPH * (5 + (NC > 1 ? NC : 0));
};

WB_HD int wb_nint(double x) { return (int)(x < 0 ? x - 0.5 : x + 0.5); }

// eos%unscale (src/eos.F90:200-210; adaptive partial pressure: src/eos_wge.F90:659-674)
template <int NP> WB_HD void wb_unscale(const WbEosParams &e, const double *y, int region, double *primary) {
#pragma unroll
  for (int i = 0; i < NP; i++) primary[i] = y[i] * e.scale[i][region];
  if (NP == 3 && e.adaptive_pp) primary[NP - 1] = y[NP - 1] * primary[0];
}
// eos%scale (src/eos.F90:186-196; adaptive: src/eos_wge.F90:639-655)
template <int NP> WB_HD void wb_scale(const WbEosParams &e, const double *primary, int region, double *y) {
#pragma unroll
  for (int i = 0; i < NP; i++) y[i] = primary[i] / e.scale[i][region];
  if (NP == 3 && e.adaptive_pp) y[NP - 1] = primary[NP - 1] / primary[0];
}

// ---------------------------------------------------------------- CO2 (non-condensible gas)
// src/ncg_co2_thermodynamics.F90:14-292, src/ncg_thermodynamics.F90:145-340

#define WB_CO2_MW 44.01          // ncg_co2_thermodynamics.F90:14
#define WB_WATER_MW 18.01528     // thermodynamics.F90:38
#define WB_GAS_CONSTANT 8.3144598  // thermodynamics.F90:39

// utils.F90:224-241 (Horner), coefficients a[0..n-1]
template <int N> WB_HD double wb_polynomial(const double *a, double x) {
  double p = a[N - 1];
#pragma unroll
  for (int i = N - 2; i >= 0; i--) p = a[i] + x * p;
  return p;
}

// density and enthalpy of CO2 at (partial pressure, temperature): ncg_co2_thermodynamics.F90:84-111
WB_HD void wb_co2_properties(double partial_pressure, double temperature, double &density, double &enthalpy) {
  const double tk = temperature + 273.15;
  const double pp = partial_pressure * 1.0e-6;
  const double tc = pow(0.01 * tk, 3.3333333333);
  const double hci = 1.667 + 0.001542 * tk - 0.7948 * log10(tk) - 41.35 / tk;
  enthalpy = 1.e6 * (hci - 0.3571 * pp * (1.0 + 0.07576 * pp) / tc);
  const double vc = 0.00018882 * tk - pp * (0.0824 + 0.01249 * pp) / tc;
  density = pp / vc;
}

// Henry's constant (:115-135) and the energy of solution from its temperature derivative
// (:172-197 with polynomial_derivative utils.F90:291-310; ncg_thermodynamics.F90:176-223)
WB_HD double wb_co2_henrys_constant(double temperature) {
  const double a[6] = {0.783666, 1.96025, 8.20574, -7.40674, 2.18380, -0.220999};
  return 1.e8 * wb_polynomial<6>(a, temperature / 100.0);
}
WB_HD double wb_co2_energy_solution(double temperature, double henrys_constant) {
  const double a[6] = {0.783666, 1.96025, 8.20574, -7.40674, 2.18380, -0.220999};
  double da[5];
#pragma unroll
  for (int i = 0; i < 5; i++) da[i] = (double)(i + 1) * a[i + 1];
  const double henrys_derivative = 1.e8 * wb_polynomial<5>(da, temperature / 100.0) / (henrys_constant * 100.0);
  const double tk = temperature + 273.15;
  return -1.e3 * WB_GAS_CONSTANT * tk * tk * henrys_derivative / WB_CO2_MW;
}

// viscosity: coefficients interpolated linearly in pressure (MPa) on the table of :22-30 (end clamping of
// interpolation.F90:202-306, 494-510), polynomial in temperature (:237-263)
WB_HD int wb_co2_viscosity(double partial_pressure, double temperature, double &viscosity) {
  if (!(partial_pressure <= 300.e5)) return 1;
  const double xp[5] = {0.0, 10.0, 15.0, 20.0, 30.0};
  const double c[5][5] = {{1.3578, 3.9189, 9.6607, 13.1566, 14.7968},
                          {4.9227e-3, -35.984e-3, -135.479e-3, -179.352e-3, -160.731e-3},
                          {-2.9661e-6, 0.25825e-3, 0.90087e-3, 1.12474e-3, 0.850257e-3},
                          {2.8529e-9, -7.1178e-7, -2.4727e-6, -2.98864e-6, -1.99076e-6},
                          {-2.1829e-12, 6.9578e-10, 2.4156e-9, 2.85911e-9, 1.73423e-9}};
  const double x = partial_pressure / 1.e6;
  double coefs[5];
  if (x <= xp[0]) {
#pragma unroll
    for (int d = 0; d < 5; d++) coefs[d] = c[d][0];
  } else if (x >= xp[4]) {
#pragma unroll
    for (int d = 0; d < 5; d++) coefs[d] = c[d][4];
  } else {
    int i = 0;
    while (i + 2 < 5 && x >= xp[i + 1]) i++;
    const double xi = (x - xp[i]) / (xp[i + 1] - xp[i]);
#pragma unroll
    for (int d = 0; d < 5; d++) coefs[d] = (1.0 - xi) * c[d][i] + xi * c[d][i + 1];
  }
  viscosity = 1.e-5 * wb_polynomial<5>(coefs, temperature);
  return 0;
}

// ncg_thermodynamics.F90:145-157
WB_HD double wb_co2_mole_to_mass_fraction(double xmole) {
  const double w = xmole * WB_CO2_MW;
  return w / (w + (1.0 - xmole) * WB_WATER_MW);
}

// ---------------------------------------------------------------- air (non-condensible gas of eos_wae)
// src/ncg_air_thermodynamics.F90

#define WB_AIR_MW 28.96  // :14

// ncg_air_properties (:97-121): ideal gas (deviation factor 1), enthalpy relative to the triple point of water
WB_HD void wb_air_properties(double partial_pressure, double temperature, double &density, double &enthalpy) {
  const double a[4] = {1.20740, 9.24502, 0.115984, -5.63568e-4};
  const double tk = temperature + 273.15;
  const double enthalpy_shift = wb_polynomial<4>(a, (0.01 + 273.15) / 100.0);  // ncg_air_init :84-86
  density = partial_pressure * WB_AIR_MW / (1.e3 * WB_GAS_CONSTANT * 1.0 * tk);
  enthalpy = 1.e4 * (wb_polynomial<4>(a, tk / 100.0) - enthalpy_shift);
}

// Henry's constant of the two constituents (N2, O2) and their weighted sum (:125-143); energy of solution from the
// temperature derivatives (:176-198, ncg_thermodynamics.F90:187-231)
WB_HD void wb_air_henry(double temperature, double &henrys_constant, double &energy_solution) {
  const double w[2] = {0.79, 0.21}, p0[2] = {1.01325e5, 1.e5};
  const double a[2][7] = {{0.513726, 1.58603, -5.9378e-1, -6.98282e-1, 5.10330e-1, -1.21388e-1, 1.00041e-2},
                          {0.26234, 0.610628, 7.00732e-1, -0.139299e1, 7.13850e-1, -1.54216e-1, 1.23190e-2}};
  double hc = 0.0, henrys_derivative = 0.0;
#pragma unroll
  for (int c = 0; c < 2; c++) {
    const double chc = 1.e5 * p0[c] * wb_polynomial<7>(a[c], temperature / 100.0);
    hc += w[c] * chc;
    double da[6];
#pragma unroll
    for (int i = 0; i < 6; i++) da[i] = (double)(i + 1) * a[c][i + 1];
    const double dhinv = 1.e5 * wb_polynomial<6>(da, temperature / 100.0);
    const double d = p0[c] * dhinv / (chc * 100.0);
    henrys_derivative += w[c] * d;
  }
  henrys_constant = hc;
  const double tk = temperature + 273.15;
  energy_solution = -1.e3 * WB_GAS_CONSTANT * tk * tk * henrys_derivative / WB_AIR_MW;
}

WB_HD double wb_air_mole_to_mass_fraction(double xmole) {  // ncg_thermodynamics.F90:155-167
  const double w = xmole * WB_AIR_MW;
  return w / (w + (1.0 - xmole) * WB_WATER_MW);
}

WB_HD double wb_air_covis(double trd, double c, double ome, double rm, double f) {
  return 266.93e-7 * sqrt(rm * trd * f) / (c * c * ome * trd);
}
// ncg_air_mixture_viscosity (:256-312) for the vapour phase
WB_HD double wb_air_vapour_viscosity(double water_viscosity, double temperature, double xg) {
  const double fair = 97.0, fwat = 363.0, cair = 3.617, cwat = 2.655;
  const double rm1 = WB_AIR_MW, rm2 = WB_WATER_MW;
  const double fmix = sqrt(fair * fwat), cmix = 0.5 * (cair + cwat);
  const double wm = xg / WB_AIR_MW;  // mass_to_mole_fraction, ncg_thermodynamics.F90:171-183
  const double x1 = wm / (wm + (1.0 - xg) / WB_WATER_MW), x2 = 1.0 - x1;
  const double tk = temperature + 273.15;
  const double trd1 = tk / fair, trd3 = tk / fmix;
  const double ome1 = (1.188 - 0.051 * trd1) / trd1;
  const double ome3 = (1.48 - 0.412 * log(trd3)) / trd3;
  const double ard = 1.095 / trd3;
  const double rm3 = 2.0 * rm1 * rm2 / (rm1 + rm2);
  const double vis1 = wb_air_covis(trd1, cair, ome1, rm1, fair);
  const double vis2 = 10.0 * water_viscosity;
  const double vis3 = wb_air_covis(trd3, cmix, ome3, rm3, fmix);
  const double z1 = x1 * x1 / vis1 + 2.0 * x2 * x1 / vis3 + x2 * x2 / vis2;
  const double g = x1 * x1 * rm1 / rm2;
  const double h = x2 * x2 * rm2 / rm1;
  const double ee = (2.0 * x1 * x2 * rm1 * rm2 / (rm3 * rm3)) * vis3 / (vis1 * vis2);
  const double z2 = 0.6 * ard * (g / vis1 + ee + h / vis2);
  const double z3 = 0.6 * ard * (g + ee * (vis1 + vis2) - 2.0 * x1 * x2 + h);
  return 0.1 * (1.0 + z3) / (z1 + z2);
}

// bulk_properties + phase_saturations + phase_properties for one cell:
// eos_we: src/eos_we.F90:327-390, 394-458; eos_w: src/eos_w.F90.
// `fl.region` must be set on entry.  Returns the reference's err (0 / 1).
template <int EOS>
WB_HD int wb_eos_properties(const WbEosParams &e, const double *primary,
                            WbFluid<WbEosTraits<EOS>::NC, WbEosTraits<EOS>::NPH> &fl) {
  const WbThermo &th = e.thermo;
  int err = 0;
  if (EOS == WB_EOS_W) {
    fl.P = primary[0];
    fl.T = e.eos_w_temperature;
    fl.ph[0].sat = 1.0;
    const int phases = wb_phase_composition(th, fl.region, fl.P, fl.T);
    if (phases <= 0) return 1;
    fl.phases = phases;
    fl.pp[0] = fl.P;
    const int p = fl.region;
    double rho, u;
    err = wb_region_properties(th, p, fl.P, fl.T, rho, u);
    if (err) return err;
    // eos_w has a single phase slot; region 2 (steam) would index phase 2 in the reference,
    // which isothermal-water models never reach
    fl.ph[0].rho = rho;
    fl.ph[0].u = u;
    fl.ph[0].h = u + fl.P / rho;
    fl.ph[0].kr = 1.0;
    fl.ph[0].pc = 0.0;
    fl.ph[0].X[0] = 1.0;
    fl.ph[0].mu = wb_region_viscosity(th, p, fl.T, fl.P, rho);
    return 0;
  } else if (EOS == WB_EOS_WCE) {
    // bulk_properties src/eos_wge.F90:350-389, phase_saturations :393-417, phase_properties :421-543
    constexpr int NPH = WbEosTraits<EOS>::NPH, XG = WbEosTraits<EOS>::NC - 1;
    fl.P = primary[0];
    const int region = fl.region;
    const double pg = primary[WbEosTraits<EOS>::NP - 1];
    fl.pp[0] = fl.P - pg;
    fl.pp[XG] = pg;
    if (region == 4) err = wb_saturation_temperature(th, fl.pp[0], fl.T);
    else fl.T = primary[1];
    if (err) return err;
    const int phases = wb_phase_composition(th, region, fl.P, fl.T);
    if (phases <= 0) return 1;
    fl.phases = phases;
    double sl = fl.ph[0].sat, sv = fl.ph[1].sat;
    if (region == 1) { sl = 1.0; sv = 0.0; }
    else if (region == 2) { sl = 0.0; sv = 1.0; }
    else if (region == 4) { sl = 1.0 - primary[1]; sv = primary[1]; }
    fl.ph[0].sat = sl;
    fl.ph[1].sat = sv;
    double kr[2];
    wb_relperm_values(e.relperm, sl, kr[0], kr[1]);
    double gas_density_free, gas_enthalpy;
    const bool air = e.gas == 1;
    if (air) wb_air_properties(pg, fl.T, gas_density_free, gas_enthalpy);
    else wb_co2_properties(pg, fl.T, gas_density_free, gas_enthalpy);
#pragma unroll
    for (int p = 0; p < NPH; p++) {
      if (phases & (1 << p)) {
        double water_pressure, capillary_pressure, henrys_constant = 0.0, energy_solution = 0.0;
        if (p == 0) {
          water_pressure = fl.P;
          capillary_pressure = wb_cappress_value(e.cappress, sl, fl.T);
          if (air) {
            wb_air_henry(fl.T, henrys_constant, energy_solution);
          } else {
            henrys_constant = wb_co2_henrys_constant(fl.T);
            energy_solution = wb_co2_energy_solution(fl.T, henrys_constant);
          }
        } else {
          water_pressure = fl.pp[0];
          capillary_pressure = 0.0;
        }
        double water_density, water_u;
        err = wb_region_properties(th, p + 1, water_pressure, fl.T, water_density, water_u);
        if (err) return err;
        // effective_properties: no free gas density in the liquid phase (ncg_thermodynamics.F90:315-340)
        const double gas_density = (p == 0) ? 0.0 : gas_density_free;
        // mass_fraction: ncg_thermodynamics.F90:279-311
        double xg;
        if (p == 0) {
          xg = air ? wb_air_mole_to_mass_fraction(pg / henrys_constant) : wb_co2_mole_to_mass_fraction(pg / henrys_constant);
        } else {
          const double total_density = gas_density + water_density;
          xg = (total_density < 1.e-30) ? 0.0 : gas_density / total_density;
        }
        const double water_viscosity = wb_region_viscosity(th, p + 1, fl.T, fl.P, water_density);
        if (p == 0) {
          fl.ph[p].mu = water_viscosity;  // mixture_viscosity: ncg_co2_thermodynamics.F90:267-292, ncg_air :256-312
        } else if (air) {
          fl.ph[p].mu = wb_air_vapour_viscosity(water_viscosity, fl.T, xg);
        } else {
          double gas_viscosity;
          if (wb_co2_viscosity(pg, fl.T, gas_viscosity)) return 1;
          fl.ph[p].mu = water_viscosity * (1.0 - xg) + gas_viscosity * xg;
        }
        fl.ph[p].rho = water_density + gas_density;
        fl.ph[p].X[0] = 1.0 - xg;
        fl.ph[p].X[XG] = xg;
        fl.ph[p].kr = kr[p];
        fl.ph[p].pc = capillary_pressure;
        const double water_enthalpy = water_u + water_pressure / water_density;
        fl.ph[p].h = water_enthalpy * (1.0 - xg) + (gas_enthalpy + energy_solution) * xg;
        fl.ph[p].u = fl.ph[p].h - fl.P / fl.ph[p].rho;
      } else {
        fl.ph[p].rho = 0.0; fl.ph[p].u = 0.0; fl.ph[p].h = 0.0; fl.ph[p].kr = 0.0;
        fl.ph[p].pc = 0.0; fl.ph[p].mu = 0.0; fl.ph[p].X[0] = 0.0; fl.ph[p].X[XG] = 0.0;
      }
    }
    return 0;
  } else {
    constexpr int NPH = WbEosTraits<EOS>::NPH;
    fl.P = primary[0];
    const int region = fl.region;
    if (region == 4) err = wb_saturation_temperature(th, fl.P, fl.T);
    else fl.T = primary[1];
    if (err) return err;
    const int phases = wb_phase_composition(th, region, fl.P, fl.T);
    if (phases <= 0) return 1;
    fl.phases = phases;
    // phase_saturations: src/eos_we.F90:366-390 (other regions leave saturations untouched)
    double sl = fl.ph[0].sat, sv = fl.ph[1].sat;
    if (region == 1) { sl = 1.0; sv = 0.0; }
    else if (region == 2) { sl = 0.0; sv = 1.0; }
    else if (region == 4) { sl = 1.0 - primary[1]; sv = primary[1]; }
    fl.ph[0].sat = sl;
    fl.ph[1].sat = sv;
    fl.pp[0] = fl.P;
    double kr[2], pc[2];
    wb_relperm_values(e.relperm, sl, kr[0], kr[1]);
    pc[0] = wb_cappress_value(e.cappress, sl, fl.T);
    pc[1] = 0.0;
#pragma unroll
    for (int p = 0; p < NPH; p++) {
      if (phases & (1 << p)) {
        double rho, u;
        err = wb_region_properties(th, p + 1, fl.P, fl.T, rho, u);
        if (err) return err;
        fl.ph[p].rho = rho;
        fl.ph[p].u = u;
        fl.ph[p].h = u + fl.P / rho;
        fl.ph[p].X[0] = 1.0;
        fl.ph[p].kr = kr[p];
        fl.ph[p].pc = pc[p];
        fl.ph[p].mu = wb_region_viscosity(th, p + 1, fl.T, fl.P, rho);
      } else {
        fl.ph[p].rho = 0.0; fl.ph[p].u = 0.0; fl.ph[p].h = 0.0; fl.ph[p].kr = 0.0;
        fl.ph[p].pc = 0.0; fl.ph[p].mu = 0.0; fl.ph[p].X[0] = 0.0;
      }
    }
    return 0;
  }
}

// rock record offsets (src/rock.F90:97-112)
enum { WB_R_PERM = 0, WB_R_WET = 3, WB_R_DRY = 4, WB_R_POR = 5, WB_R_RHO = 6, WB_R_CP = 7 };

// cell%balance: src/cell.F90:114-142, src/fluid.F90:295-370, src/rock.F90:142
template <int NP, int NC, int NPH>
WB_HD void wb_cell_balance(const WbFluid<NC, NPH> &fl, double porosity, double rock_density, double rock_cp,
                           double *balance) {
  double d[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) d[c] = 0.0;
#pragma unroll
  for (int p = 0; p < NPH; p++) {
    const double ds = fl.ph[p].rho * fl.ph[p].sat;
#pragma unroll
    for (int c = 0; c < NC; c++) d[c] = d[c] + ds * fl.ph[p].X[c];
  }
#pragma unroll
  for (int c = 0; c < NC; c++) balance[c] = porosity * d[c];
  if (NP != NC) {
    const double er = rock_density * rock_cp * fl.T;
    double ef = 0.0;
#pragma unroll
    for (int p = 0; p < NPH; p++) {
      const double ds = fl.ph[p].rho * fl.ph[p].sat;
      ef = ef + ds * fl.ph[p].u;
    }
    balance[NP - 1] = porosity * ef + (1.0 - porosity) * er;
  }
}

// compact state from the full property set; mobility = kr*rho/mu (src/fluid.F90:197-207),
// conductivity = dry + sqrt(Sl)*(wet-dry) (src/eos.F90:240-257)
template <int NC, int NPH>
WB_HD void wb_state_from_fluid(const WbFluid<NC, NPH> &fl, double wet, double dry, WbCellState<NC, NPH> &s) {
  s.P = fl.P;
  s.T = fl.T;
  s.phases = fl.phases;
  s.cond = dry + sqrt(fl.ph[0].sat) * (wet - dry);
#pragma unroll
  for (int p = 0; p < NPH; p++) {
    s.rho[p] = fl.ph[p].rho;
    s.sat[p] = fl.ph[p].sat;
    s.pc[p] = fl.ph[p].pc;
    s.h[p] = fl.ph[p].h;
    s.mob[p] = (fl.phases & (1 << p)) ? fl.ph[p].kr * fl.ph[p].rho / fl.ph[p].mu : 0.0;
#pragma unroll
    for (int c = 0; c < NC; c++) s.X[p][c] = fl.ph[p].X[c];
  }
}

// face%harmonic_average: src/face.F90:358-377
WB_HD double wb_harmonic(double d1, double d2, double d12, double x1, double x2) {
  const double wx = (d1 * x2 + d2 * x1) / d12;
  return (fabs(wx) > 1.e-30) ? x1 * x2 / wx : 0.0;
}

// face geometry as the flux needs it (subset of the 12-double record, src/face.F90:127-133)
struct WbFaceGeom {
  double area, d1, d2, d12, gravn, k;  // k: harmonic-averaged permeability along the face normal
};

// face%flux: src/face.F90:443-515.  flux[0..NC-1] component mass fluxes, flux[NP-1] energy
// flux (if non-isothermal), accumulated in the reference's order: conduction, then phase 1, 2.
template <int NP, int NC, int NPH>
WB_HD void wb_face_flux(const WbFaceGeom &g, const WbCellState<NC, NPH> &s1, const WbCellState<NC, NPH> &s2,
                        double *flux, double *phase_flux) {
#pragma unroll
  for (int i = 0; i < NP; i++) flux[i] = 0.0;
  if (NP != NC) {
    const double cond = wb_harmonic(g.d1, g.d2, g.d12, s1.cond, s2.cond);
    const double dtdn = (s2.T - s1.T) / g.d12;
    flux[NP - 1] = -cond * dtdn;
  }
  const int present = s1.phases | s2.phases;
#pragma unroll
  for (int p = 0; p < NPH; p++) {
    double pf = 0.0;
    if (present & (1 << p)) {
      // phase_density: src/face.F90:334-354
      double rho = 0.0, weight = 0.0;
      rho = rho + s1.sat[p] * s1.rho[p];
      weight = weight + s1.sat[p];
      rho = rho + s2.sat[p] * s2.rho[p];
      weight = weight + s2.sat[p];
      rho = rho / weight;
      // pressure_gradient: src/face.F90:296-313
      const double dpdn = ((s2.P + s2.pc[p]) - (s1.P + s1.pc[p])) / g.d12;
      const double G = dpdn - rho * g.gravn;
      const bool up1 = (G <= 0.0);  // src/face.F90:426-439
      const int up_phases = up1 ? s1.phases : s2.phases;
      if (up_phases & (1 << p)) {
        const double mob = up1 ? s1.mob[p] : s2.mob[p];
        const double F = -g.k * mob * G;
        double sum = 0.0;
#pragma unroll
        for (int c = 0; c < NC; c++) {
          const double pcf = F * (up1 ? s1.X[p][c] : s2.X[p][c]);
          flux[c] = flux[c] + pcf;
          sum += pcf;
        }
        if (NP != NC) {
          const double h = up1 ? s1.h[p] : s2.h[p];
          flux[NP - 1] = flux[NP - 1] + h * F;
        }
        pf = sum;
      }
    }
    if (phase_flux) phase_flux[p] = pf;
  }
}

// ---------------------------------------------------------------- transitions

// Brent on f(x) = P(x) - Psat(T(x)) along the segment old -> new primaries
// (src/root_finder.F90:127-248 with the defaults of :76-79; src/eos_we.F90:530-553)
struct WbSatLine {
  double p0, t0, p1, t1;
  double g0, g1;  // gas partial pressure at both ends (eos_wge: f = P - Pg - Psat(T), src/eos_wge.F90:678-701)
  bool gas;
};
WB_HD double wb_satline_f(const WbThermo &th, const WbSatLine &c, double x) {
  const double xi = (x - 0.0) / (1.0 - 0.0);
  const double P = (1.0 - xi) * c.p0 + xi * c.p1;
  const double T = (1.0 - xi) * c.t0 + xi * c.t1;
  double Ps = 0.0;
  wb_saturation_pressure(th, T, Ps);
  if (c.gas) {
    const double Pg = (1.0 - xi) * c.g0 + xi * c.g1;
    return P - Pg - Ps;
  }
  return P - Ps;
}

WB_HD int wb_brent_satline(const WbThermo &th, const WbSatLine &ctx, double &root) {
  const double rtol = 1.e-8, ftol = 1.e-8, small = 1.e-16;
  const int max_iterations = 100;
  double a = 0.0, b = 1.0, c, d = 0.0, e = 0.0, fa, fb, fc, dx, p, pc, q, r, s;
  bool found = false;
  root = 0.0;
  fa = wb_satline_f(th, ctx, a);
  fb = wb_satline_f(th, ctx, b);
  if (fa * fb > 0.0) return 1;
  c = b;
  fc = fb;
  for (int iter = 1; iter <= max_iterations; iter++) {
    if (fb * fc > 0.0) {
      c = a; fc = fa; d = b - a; e = d;
    }
    if (fabs(fc) < fabs(fb)) {
      a = b; b = c; c = a;
      fa = fb; fb = fc; fc = fa;
    }
    dx = 0.5 * (c - b);
    if (fabs(dx) <= rtol || fabs(fb) <= ftol) {
      found = true;
      break;
    }
    if (fabs(e) >= rtol && fabs(fa) > fabs(fb)) {
      s = fb / fa;
      if (fabs(a - c) <= small) {
        p = 2.0 * dx * s;
        q = 1.0 - s;
      } else {
        q = fa / fc;
        r = fb / fc;
        p = s * (2.0 * dx * q * (q - r) - (b - a) * (r - 1.0));
        q = (q - 1.0) * (r - 1.0) * (s - 1.0);
      }
      if (p > 0.0) q = -q;
      else p = -p;
      pc = fmin(3.0 * dx * q - fabs(rtol * q), fabs(e * q));
      if (2.0 * p < pc) {
        e = d;
        d = p / q;
      } else {
        d = dx;
        e = d;
      }
    } else {
      d = dx;
      e = d;
    }
    a = b;
    fa = fb;
    if (fabs(d) > rtol) b = b + d;
    else b = b + copysign(rtol, dx);
    fb = wb_satline_f(th, ctx, b);
  }
  root = b;
  return found ? 0 : 2;
}

// eos_we%transition: src/eos_we.F90:149-323.  primary in/out (unscaled), region in/out.
// old_region / old_T are the cell's values at the start of the Newton iteration
// (last_iteration_fluid).  Returns err; transition set when the region changed.
WB_HD int wb_we_transition(const WbThermo &th, const double *old_primary, double *primary, int old_region,
                           double old_T, int &region, bool &transition) {
  const double small = 1.e-6;
  int err = 0;
  transition = false;
  if (old_region == 4) {
    const double sv = primary[1];
    int new_region = 0;
    if (sv < 0.0) new_region = 1;
    else if (sv > 1.0) new_region = 2;
    if (new_region) {
      // transition_to_single_phase: src/eos_we.F90:149-216
      const double bound = (new_region == 1) ? 0.0 : 1.0;
      const double pfac = (new_region == 1) ? 1.0 + small : 1.0 - small;
      // inverse interpolation of the saturation component on x = [0,1]
      // (src/interpolation.F90:407-438), tolerance 1e-8
      const double v1 = old_primary[1], v2 = primary[1];
      const double vmax = fmax(fabs(v1), fabs(v2));
      if (fabs(v2 - v1) >= 1.e-8 * vmax) {
        const double vs1 = v1 / vmax, vs2 = v2 / vmax, ys = bound / vmax;
        const double xq = (ys - vs1) / (vs2 - vs1);
        const double xi = (1.0 - xq) * 0.0 + xq * 1.0;
        // interpolate(xi): find + interpolate_at_index, clamped outside [0,1]
        double pint;
        if (xi <= 0.0) pint = old_primary[0];
        else if (xi >= 1.0) pint = primary[0];
        else {
          const double w = (xi - 0.0) / (1.0 - 0.0);
          pint = (1.0 - w) * old_primary[0] + w * primary[0];
        }
        primary[0] = pfac * pint;
        err = wb_saturation_temperature(th, pint, primary[1]);
        if (err == 0) {
          region = new_region;
          transition = true;
        }
      } else {
        double old_ps;
        err = wb_saturation_pressure(th, old_T, old_ps);
        if (err == 0) {
          primary[0] = pfac * old_ps;
          primary[1] = old_T;
          region = new_region;
          transition = true;
        }
      }
    }
  } else {
    double ps;
    err = wb_saturation_pressure(th, primary[1], ps);
    if (err == 0) {
      if ((old_region == 1 && primary[0] < ps) || (old_region == 2 && primary[0] > ps)) {
        // transition_to_two_phase: src/eos_we.F90:220-268
        WbSatLine c = {old_primary[0], old_primary[1], primary[0], primary[1], 0.0, 0.0, false};
        double root;
        if (wb_brent_satline(th, c, root) == 0) {
          double pint;
          if (root <= 0.0) pint = c.p0;
          else if (root >= 1.0) pint = c.p1;
          else {
            const double w = (root - 0.0) / (1.0 - 0.0);
            pint = (1.0 - w) * c.p0 + w * c.p1;
          }
          primary[0] = pint;
        } else {
          primary[0] = ps;
        }
        primary[1] = (old_region == 1) ? small : 1.0 - small;
        region = 4;
        transition = true;
      }
    }
  }
  return err;
}

// eos_we%check_primary_variables: src/eos_we.F90:486-526
WB_HD int wb_we_check_primary(const double *primary, int region) {
  const double p = primary[0];
  if (p < 0.0 || p > 100.e6) return 1;
  if (region == 4) {
    const double sv = primary[1];
    if (sv < -1.0 || sv > 2.0) return 1;
  } else {
    const double t = primary[1];
    if (t < 0.0 || t > 800.0) return 1;
  }
  return 0;
}


// 2-point table on x = [0, 1]: interpolate(xi) = find + interpolate_at_index, clamped outside [0, 1]
// (src/interpolation.F90:202-306, 388-403)
WB_HD double wb_interp01(double v0, double v1, double xi) {
  if (xi <= 0.0) return v0;
  if (xi >= 1.0) return v1;
  const double w = (xi - 0.0) / (1.0 - 0.0);
  return (1.0 - w) * v0 + w * v1;
}

// eos_wge%transition: src/eos_wge.F90:149-346.  primary = (P, T | Sv, Pg) unscaled, in/out.
WB_HD int wb_wge_transition(const WbThermo &th, const double *old_primary, double *primary, int old_region,
                            double old_T, int &region, bool &transition) {
  const double small = 1.e-6;
  int err = 0;
  transition = false;
  if (old_region == 4) {
    const double sv = primary[1];
    int new_region = 0;
    if (sv < 0.0) new_region = 1;
    else if (sv > 1.0) new_region = 2;
    if (new_region) {
      // transition_to_single_phase: :149-227
      const double bound = (new_region == 1) ? 0.0 : 1.0;
      const double pfac = (new_region == 1) ? 1.0 + small : 1.0 - small;
      primary[2] = fmax(0.0, fmin(primary[2], primary[0]));
      const double v1 = old_primary[1], v2 = primary[1];
      const double vmax = fmax(fabs(v1), fabs(v2));
      if (fabs(v2 - v1) >= 1.e-8 * vmax) {
        const double vs1 = v1 / vmax, vs2 = v2 / vmax, ys = bound / vmax;
        const double xq = (ys - vs1) / (vs2 - vs1);
        const double xi = (1.0 - xq) * 0.0 + xq * 1.0;
        const double pint = wb_interp01(old_primary[0], primary[0], xi);
        const double gint = wb_interp01(old_primary[2], primary[2], xi);
        const double water_pressure = pint - gint;
        primary[0] = pfac * water_pressure + gint;
        primary[2] = gint;
        err = wb_saturation_temperature(th, water_pressure, primary[1]);
        if (err == 0) {
          region = new_region;
          transition = true;
        }
      } else {
        double old_ps;
        err = wb_saturation_pressure(th, old_T, old_ps);
        if (err == 0) {
          primary[0] = pfac * old_ps + primary[2];
          primary[1] = old_T;
          region = new_region;
          transition = true;
        }
      }
    }
  } else {
    double ps;
    err = wb_saturation_pressure(th, primary[1], ps);
    if (err == 0) {
      const double water_pressure = primary[0] - primary[2];
      if ((old_region == 1 && water_pressure < ps) || (old_region == 2 && water_pressure > ps)) {
        // transition_to_two_phase: :231-285
        primary[2] = fmax(0.0, fmin(primary[2], primary[0]));
        WbSatLine c = {old_primary[0], old_primary[1], primary[0], primary[1], old_primary[2], primary[2], true};
        double root;
        if (wb_brent_satline(th, c, root) == 0) {
          primary[0] = wb_interp01(c.p0, c.p1, root);
          primary[2] = wb_interp01(c.g0, c.g1, root);
        } else {
          primary[0] = ps + primary[2];
        }
        primary[1] = (old_region == 1) ? small : 1.0 - small;
        region = 4;
        transition = true;
      }
    }
  }
  return err;
}

// eos_wge%check_primary_variables: src/eos_wge.F90:573-635 (clamps the gas partial pressure; `changed` set)
WB_HD int wb_wge_check_primary(double *primary, int region, bool &changed) {
  const double small = 1.e-6;
  changed = false;
  if (!(primary[0] > 0.0)) return 1;
  const double max_pp = (1.0 - small) * primary[0];
  if (primary[2] > max_pp) {
    primary[2] = max_pp;
    changed = true;
  } else if (primary[2] < 0.0) {
    primary[2] = 0.0;
    changed = true;
  }
  const double pw = primary[0] - primary[2];
  if (pw > 100.e6) return 1;
  if (region == 4) {
    if (primary[1] < -1.0 || primary[1] > 2.0) return 1;
  } else {
    if (primary[1] < 0.0 || primary[1] > 800.0) return 1;
  }
  return 0;
}
