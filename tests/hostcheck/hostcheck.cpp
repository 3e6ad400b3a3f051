// hostcheck.cpp -- TEST-ONLY: compiles the device physics headers (wb_thermo.cuh,
// wb_eos.cuh, wb_iapws_gen.cuh) with the host compiler so that the -m "not gpu"
// suite can compare the exact source the CUDA kernels inline against the oracle,
// cell by cell, without a GPU.  This library lives under tests/, is never built
// by the package and is never loaded by the product: waiwera_b200 has no CPU path.
#include <string.h>

#include "../../waiwera_b200/csrc/wb_eos.cuh"

extern "C" {

int hc_region_properties(int thermo, int extrapolate, int region, double p, double t, double *props) {
  WbThermo th = wb_thermo_make(thermo, extrapolate);
  return wb_region_properties(th, region, p, t, props[0], props[1]);
}
double hc_region_viscosity(int thermo, int region, double t, double p, double rho) {
  WbThermo th = wb_thermo_make(thermo, 0);
  return wb_region_viscosity(th, region, t, p, rho);
}
int hc_sat_pressure(int thermo, double t, double *p) {
  WbThermo th = wb_thermo_make(thermo, 0);
  return wb_saturation_pressure(th, t, *p);
}
int hc_sat_temperature(int thermo, double p, double *t) {
  WbThermo th = wb_thermo_make(thermo, 0);
  return wb_saturation_temperature(th, p, *t);
}
void hc_relperm(const wb_relperm *rp, double sl, double *out) { wb_relperm_values(*rp, sl, out[0], out[1]); }
double hc_cappress(const wb_cappress *cp, double sl, double t) { return wb_cappress_value(*cp, sl, t); }

// fluid record (reference AoS layout) of one eos_we cell from unscaled primaries
int hc_we_fluid(const wb_params *prm, const double *primary, int region, double *r) {
  WbEosParams e;
  if (wb_eos_params_make(*prm, e)) return -1;
  WbFluid<1, 2> fl = {};
  fl.region = region;
  int err = wb_eos_properties<WB_EOS_WE>(e, primary, fl);
  r[0] = fl.P; r[1] = fl.T; r[2] = region; r[3] = 0; r[4] = fl.phases; r[5] = 1.0; r[6] = fl.pp[0];
  for (int p = 0; p < 2; p++) {
    double *q = r + 7 + 8 * p;
    q[0] = fl.ph[p].rho; q[1] = fl.ph[p].mu; q[2] = fl.ph[p].sat; q[3] = fl.ph[p].kr; q[4] = fl.ph[p].pc;
    q[5] = fl.ph[p].h; q[6] = fl.ph[p].u; q[7] = fl.ph[p].X[0];
  }
  return err;
}

// balance + flux between two eos_we cells given unscaled primaries, regions, rock records and
// the 12-double face record (permeability harmonic-averaged as k_face_perm does)
int hc_we_flux(const wb_params *prm, const double *face12, const double *rock1, const double *rock2,
               const double *prim1, int reg1, const double *prim2, int reg2, double *flux4, double *bal1) {
  WbEosParams e;
  if (wb_eos_params_make(*prm, e)) return -1;
  WbFluid<1, 2> f1 = {}, f2 = {};
  f1.region = reg1;
  f2.region = reg2;
  int err = wb_eos_properties<WB_EOS_WE>(e, prim1, f1);
  err |= wb_eos_properties<WB_EOS_WE>(e, prim2, f2);
  WbCellState<1, 2> s1, s2;
  wb_state_from_fluid(f1, rock1[WB_R_WET], rock1[WB_R_DRY], s1);
  wb_state_from_fluid(f2, rock2[WB_R_WET], rock2[WB_R_DRY], s2);
  for (int p = 0; p < 2; p++) {
    s1.X[p][0] = (s1.phases >> p) & 1 ? 1.0 : 0.0;
    s2.X[p][0] = (s2.phases >> p) & 1 ? 1.0 : 0.0;
  }
  WbFaceGeom g;
  g.area = face12[0]; g.d1 = face12[1]; g.d2 = face12[2]; g.d12 = face12[3]; g.gravn = face12[7];
  const int d = (int)(face12[11] + 0.5) - 1;
  g.k = wb_harmonic(g.d1, g.d2, g.d12, rock1[d] * 1.0, rock2[d] * 1.0);
  wb_face_flux<2, 1, 2>(g, s1, s2, flux4, flux4 + 2);
  wb_cell_balance<2, 1, 2>(f1, rock1[WB_R_POR], rock1[WB_R_RHO], rock1[WB_R_CP], bal1);
  return err;
}

// transition of one eos_we cell: unscaled primaries in/out
int hc_we_transition(const wb_params *prm, const double *old_primary, double *primary, int old_region, double old_T,
                     int *region, int *transition) {
  WbEosParams e;
  if (wb_eos_params_make(*prm, e)) return -1;
  bool tr = false;
  int reg = *region;
  int err = wb_we_transition(e.thermo, old_primary, primary, old_region, old_T, reg, tr);
  if (err == 0) err = wb_we_check_primary(primary, reg);
  *region = reg;
  *transition = tr ? 1 : 0;
  return err;
}

// fluid record (reference AoS layout, 26 doubles) of one eos_wce cell from unscaled primaries
int hc_wce_fluid(const wb_params *prm, const double *primary, int region, double *r) {
  WbEosParams e;
  if (wb_eos_params_make(*prm, e)) return -1;
  WbFluid<2, 2> fl = {};
  fl.region = region;
  int err = wb_eos_properties<WB_EOS_WCE>(e, primary, fl);
  r[0] = fl.P; r[1] = fl.T; r[2] = region; r[3] = 0; r[4] = fl.phases; r[5] = 1.0; r[6] = fl.pp[0]; r[7] = fl.pp[1];
  for (int p = 0; p < 2; p++) {
    double *q = r + 8 + 9 * p;
    q[0] = fl.ph[p].rho; q[1] = fl.ph[p].mu; q[2] = fl.ph[p].sat; q[3] = fl.ph[p].kr; q[4] = fl.ph[p].pc;
    q[5] = fl.ph[p].h; q[6] = fl.ph[p].u; q[7] = fl.ph[p].X[0]; q[8] = fl.ph[p].X[1];
  }
  return err;
}

// balance + flux between two eos_wce cells (3 component/energy fluxes + 2 phase fluxes; 3 balances)
int hc_wce_flux(const wb_params *prm, const double *face12, const double *rock1, const double *rock2,
                const double *prim1, int reg1, const double *prim2, int reg2, double *flux5, double *bal1) {
  WbEosParams e;
  if (wb_eos_params_make(*prm, e)) return -1;
  WbFluid<2, 2> f1 = {}, f2 = {};
  f1.region = reg1;
  f2.region = reg2;
  int err = wb_eos_properties<WB_EOS_WCE>(e, prim1, f1);
  err |= wb_eos_properties<WB_EOS_WCE>(e, prim2, f2);
  WbCellState<2, 2> s1, s2;
  wb_state_from_fluid(f1, rock1[WB_R_WET], rock1[WB_R_DRY], s1);
  wb_state_from_fluid(f2, rock2[WB_R_WET], rock2[WB_R_DRY], s2);
  WbFaceGeom g;
  g.area = face12[0]; g.d1 = face12[1]; g.d2 = face12[2]; g.d12 = face12[3]; g.gravn = face12[7];
  const int d = (int)(face12[11] + 0.5) - 1;
  g.k = wb_harmonic(g.d1, g.d2, g.d12, rock1[d] * 1.0, rock2[d] * 1.0);
  wb_face_flux<3, 2, 2>(g, s1, s2, flux5, flux5 + 3);
  wb_cell_balance<3, 2, 2>(f1, rock1[WB_R_POR], rock1[WB_R_RHO], rock1[WB_R_CP], bal1);
  return err;
}

// transition + check of one eos_wce cell: unscaled primaries in/out
int hc_wce_transition(const wb_params *prm, const double *old_primary, double *primary, int old_region, double old_T,
                      int *region, int *transition, int *changed) {
  WbEosParams e;
  if (wb_eos_params_make(*prm, e)) return -1;
  bool tr = false, ch = false;
  int reg = *region;
  int err = wb_wge_transition(e.thermo, old_primary, primary, old_region, old_T, reg, tr);
  if (err == 0) err = wb_wge_check_primary(primary, reg, ch);
  *region = reg;
  *transition = tr ? 1 : 0;
  *changed = ch ? 1 : 0;
  return err;
}

// scale / unscale round trip through the device functions
void hc_wce_scale(const wb_params *prm, const double *primary, int region, double *y, double *back) {
  WbEosParams e;
  wb_eos_params_make(*prm, e);
  wb_scale<3>(e, primary, region, y);
  wb_unscale<3>(e, y, region, back);
}
}
