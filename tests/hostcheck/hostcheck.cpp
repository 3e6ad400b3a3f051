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

// ---------------------------------------------------------------- tracer row assembly (wb_tracer.cuh)
#include <vector>

#include "../../waiwera_b200/csrc/wb_tracer.cuh"

// Runs wb_tracer_row -- the body of k_tracer_assemble -- over all owned rows on the host.  The harness builds
// what wb_set_mesh / k_eos / k_face_perm hold on the device: SoA cell states from unscaled primaries, SoA faces
// with the harmonic permeability, the cell -> face lists in ascending face order and the block positions in
// the given BSR pattern (rowptr / colidx over owned cells, sorted columns).
template <int EOS, int NT>
static int tracer_assemble_host(const wb_params *prm, int ncell, int nowned, int nface, const int32_t *face_cells,
                                const double *face12, const double *cell4, const double *rock8,
                                const double *primary, const int32_t *region, const int32_t *phase,
                                const double *diffusion, const double *decay, const double *activation, int nsrc,
                                const int32_t *src_cell, const int32_t *src_comp, const double *src_rate,
                                const int32_t *src_ctrl, const double *src_pi, const double *src_pref,
                                const double *src_limit, const double *inj, int method, double dt, double dt_last,
                                const double *al_last,
                                const double *x_last, const double *al_last2, const double *x_last2,
                                const double *xb, const int32_t *rowptr, const int32_t *colidx, double *val,
                                double *b, double *al) {
  constexpr int NP = WbEosTraits<EOS>::NP, NC = WbEosTraits<EOS>::NC, NPH = WbEosTraits<EOS>::NPH;
  constexpr int NF = WbStateLayout<NC, NPH>::NF;
  WbEosParams e;
  if (wb_eos_params_make(*prm, e)) return -1;
  std::vector<double> state((size_t)NF * ncell), rockp((size_t)5 * ncell), vol(ncell), face((size_t)6 * nface);
  for (int c = 0; c < ncell; c++) {
    const double *rk = rock8 + 8 * (size_t)c;
    WbFluid<NC, NPH> fl = {};
    fl.region = region[c];
    if (wb_eos_properties<EOS>(e, primary + (size_t)c * NP, fl)) return 1;
    WbCellState<NC, NPH> s;
    wb_state_from_fluid(fl, rk[WB_R_WET], rk[WB_R_DRY], s);
    store_state(state.data(), (size_t)ncell, c, s);
    rockp[c] = rk[WB_R_POR];
    vol[c] = cell4[4 * (size_t)c + 3];
  }
  for (int f = 0; f < nface; f++) {
    const double *g = face12 + 12 * (size_t)f;
    const int c1 = face_cells[2 * f], c2 = face_cells[2 * f + 1];
    const int d = (int)(g[11] + 0.5) - 1;
    face[f] = g[0];
    face[(size_t)nface + f] = g[1];
    face[(size_t)2 * nface + f] = g[2];
    face[(size_t)3 * nface + f] = g[3];
    face[(size_t)4 * nface + f] = g[7];
    face[(size_t)5 * nface + f] = wb_harmonic(g[1], g[2], g[3], rock8[8 * (size_t)c1 + d] * 1.0, rock8[8 * (size_t)c2 + d] * 1.0);
  }
  std::vector<int32_t> cf_ptr(nowned + 1, 0), cf_face, cf_other, cf_bpos, diagpos(nowned);
  for (int f = 0; f < nface; f++)
    for (int s = 0; s < 2; s++)
      if (face_cells[2 * f + s] < nowned) cf_ptr[face_cells[2 * f + s] + 1]++;
  for (int i = 0; i < nowned; i++) cf_ptr[i + 1] += cf_ptr[i];
  cf_face.resize(cf_ptr[nowned]); cf_other.resize(cf_ptr[nowned]); cf_bpos.resize(cf_ptr[nowned]);
  std::vector<int32_t> fill(cf_ptr.begin(), cf_ptr.end() - 1);
  auto find = [&](int row, int col) {
    for (int k = rowptr[row]; k < rowptr[row + 1]; k++)
      if (colidx[k] == col) return k;
    return -1;
  };
  for (int f = 0; f < nface; f++)
    for (int s = 0; s < 2; s++) {
      const int c = face_cells[2 * f + s], o = face_cells[2 * f + 1 - s];
      if (c >= nowned) continue;
      const int e = fill[c]++;
      cf_face[e] = 2 * f + s;
      cf_other[e] = o;
      cf_bpos[e] = o < nowned ? find(c, o) : -1;
    }
  for (int i = 0; i < nowned; i++) diagpos[i] = find(i, i);
  // sources sorted by cell (stable), as wb_set_sources keeps them
  std::vector<int> order(nsrc);
  for (int k = 0; k < nsrc; k++) order[k] = k;
  for (int a = 1; a < nsrc; a++) {
    const int v = order[a];
    int q = a - 1;
    while (q >= 0 && src_cell[order[q]] > src_cell[v]) { order[q + 1] = order[q]; q--; }
    order[q + 1] = v;
  }
  std::vector<int32_t> head(nowned, -1), sc(nsrc + 1), sk(nsrc + 1), sctrl(nsrc + 1);
  std::vector<double> sr(nsrc + 1), sinj((size_t)nsrc * NT + 1), spi(nsrc + 1), spref(nsrc + 1), slim(nsrc + 1);
  for (int k = 0; k < nsrc; k++) {
    sc[k] = src_cell[order[k]]; sk[k] = src_comp[order[k]] | (src_comp[order[k]] << 8);  // as wb_set_sources packs it
    sr[k] = src_rate[order[k]];
    if (src_ctrl) {
      sctrl[k] = src_ctrl[order[k]]; spi[k] = src_pi[order[k]]; spref[k] = src_pref[order[k]]; slim[k] = src_limit[order[k]];
    }
    for (int t = 0; t < NT; t++) sinj[(size_t)k * NT + t] = inj ? inj[(size_t)order[k] * NT + t] : 0.0;
    if (head[sc[k]] < 0) head[sc[k]] = k;
  }
  TracerArgs a = {};
  a.state = state.data(); a.face = face.data(); a.vol = vol.data(); a.rockp = rockp.data();
  a.cf_ptr = cf_ptr.data(); a.cf_face = cf_face.data(); a.cf_other = cf_other.data(); a.cf_bpos = cf_bpos.data();
  a.diagpos = diagpos.data(); a.rowptr = rowptr;
  a.src.head = nsrc ? head.data() : nullptr; a.src.cell = sc.data(); a.src.comp = sk.data(); a.src.rate = sr.data();
  a.src.enth = nullptr; a.src.n = nsrc;
  a.src.ctrl = src_ctrl ? sctrl.data() : nullptr; a.src.pi = spi.data(); a.src.pref = spref.data(); a.src.limit = slim.data();
  a.inj = inj ? sinj.data() : nullptr;
  for (int t = 0; t < NT; t++) {
    a.trc.phase[t] = phase[t] - 1; a.trc.diffusion[t] = diffusion[t]; a.trc.decay[t] = decay[t];
    a.trc.activation[t] = activation[t];
  }
  a.method = method;
  if (method == WB_METHOD_BDF2) {
    const double r = dt / dt_last, r1 = r + 1.0;
    a.sA = -dt * r1; a.sD = 1.0 + 2.0 * r; a.s0 = r1 * r1; a.s2 = -r * r; a.sb = dt * r1;
  } else if (method == WB_METHOD_DIRECTSS) {
    a.sA = 1.0; a.sD = 0.0;
  } else {
    a.sA = -dt; a.sD = 1.0; a.s0 = 1.0; a.sb = dt;
  }
  a.al_last = al_last; a.x_last = x_last; a.al_last2 = al_last2; a.x_last2 = x_last2; a.xb = xb;
  a.val = val; a.b = b; a.al = al;
  a.ncell = ncell; a.ninterior = nowned; a.nowned = nowned; a.nface = nface;
  for (int i = 0; i < nowned; i++) wb_tracer_row<EOS, NT>(a, i);
  return 0;
}

extern "C" int hc_tracer_assemble(const wb_params *prm, int nt, int ncell, int nowned, int nface,
                                  const int32_t *face_cells, const double *face12, const double *cell4,
                                  const double *rock8, const double *primary, const int32_t *region,
                                  const int32_t *phase, const double *diffusion, const double *decay,
                                  const double *activation, int nsrc, const int32_t *src_cell,
                                  const int32_t *src_comp, const double *src_rate, const int32_t *src_ctrl,
                                  const double *src_pi, const double *src_pref, const double *src_limit,
                                  const double *inj, int method,
                                  double dt, double dt_last, const double *al_last, const double *x_last,
                                  const double *al_last2, const double *x_last2, const double *xb,
                                  const int32_t *rowptr, const int32_t *colidx, double *val, double *b, double *al) {
#define HC_TRACER(E, T)                                                                                              \
  return tracer_assemble_host<E, T>(prm, ncell, nowned, nface, face_cells, face12, cell4, rock8, primary, region,    \
                                    phase, diffusion, decay, activation, nsrc, src_cell, src_comp, src_rate, src_ctrl, \
                                    src_pi, src_pref, src_limit, inj,                                               \
                                    method, dt, dt_last, al_last, x_last, al_last2, x_last2, xb, rowptr, colidx, val, \
                                    b, al)
  if (prm->eos == WB_EOS_WE) {
    if (nt == 1) HC_TRACER(WB_EOS_WE, 1);
    if (nt == 2) HC_TRACER(WB_EOS_WE, 2);
    if (nt == 3) HC_TRACER(WB_EOS_WE, 3);
  } else if (prm->eos == WB_EOS_WCE || prm->eos == WB_EOS_WAE) {
    if (nt == 1) HC_TRACER(WB_EOS_WCE, 1);
    if (nt == 2) HC_TRACER(WB_EOS_WCE, 2);
  }
#undef HC_TRACER
  return -2;
}


// wb_source_rate / wb_source_separated (wb_state.cuh: deliverability, recharge, direction, total / water / steam
// limiters, separators) for ONE source in a cell of the given primaries: what k_residual / k_jacobian evaluate per
// source.  out: rate, then water rate, water enthalpy, steam rate, steam enthalpy, steam fraction at that rate.
template <int EOS>
static int source_rate_host(const wb_params *prm, const double *primary, int region, const double *rock8, int ctrl, double pi,
                            double pref, double limit, double rate, int sep_n, const double *sep_h, double limit_w,
                            double limit_s, double *out) {
  constexpr int NC = WbEosTraits<EOS>::NC, NPH = WbEosTraits<EOS>::NPH;
  WbEosParams e;
  if (wb_eos_params_make(*prm, e)) return -1;
  WbFluid<NC, NPH> fl = {};
  fl.region = region;
  if (wb_eos_properties<EOS>(e, primary, fl)) return 1;
  WbCellState<NC, NPH> s;
  wb_state_from_fluid(fl, rock8[WB_R_WET], rock8[WB_R_DRY], s);
  int32_t c_ctrl = ctrl, c_sep = sep_n, c_cell = 0, c_comp = 0, c_head = 0;
  double c_enth = 0.0;
  WbSources S = {};
  S.head = &c_head; S.cell = &c_cell; S.comp = &c_comp; S.rate = &rate; S.enth = &c_enth; S.n = 1;
  S.ctrl = &c_ctrl; S.pi = &pi; S.pref = &pref; S.limit = &limit;
  S.sep_n = sep_h ? &c_sep : nullptr; S.sep_h = sep_h; S.limit_w = &limit_w; S.limit_s = &limit_s;
  out[0] = wb_source_rate(S, 0, s);
  wb_source_separated(S, 0, s, out[0], out + 1);
  return 0;
}
extern "C" int hc_source_rate(const wb_params *prm, const double *primary, int region, const double *rock8, int ctrl,
                              double pi, double pref, double limit, double rate, int sep_n, const double *sep_h,
                              double limit_w, double limit_s, double *out) {
  if (prm->eos == WB_EOS_WE)
    return source_rate_host<WB_EOS_WE>(prm, primary, region, rock8, ctrl, pi, pref, limit, rate, sep_n, sep_h, limit_w, limit_s, out);
  if (prm->eos == WB_EOS_WCE || prm->eos == WB_EOS_WAE)
    return source_rate_host<WB_EOS_WCE>(prm, primary, region, rock8, ctrl, pi, pref, limit, rate, sep_n, sep_h, limit_w, limit_s, out);
  return -2;
}
// wb_source_rate for one source on deliverability whose reference pressure is a table against the flowing enthalpy or
// the pressure of its cell (WbSources::ptab_n / ptab)
template <int EOS>
static int source_rate_ptab_host(const wb_params *prm, const double *primary, int region, const double *rock8, int ctrl, double pi,
                                 double pref, int word, const double *table, double *out) {
  constexpr int NC = WbEosTraits<EOS>::NC, NPH = WbEosTraits<EOS>::NPH;
  WbEosParams e;
  if (wb_eos_params_make(*prm, e)) return -1;
  WbFluid<NC, NPH> fl = {};
  fl.region = region;
  if (wb_eos_properties<EOS>(e, primary, fl)) return 1;
  WbCellState<NC, NPH> s;
  wb_state_from_fluid(fl, rock8[WB_R_WET], rock8[WB_R_DRY], s);
  int32_t c_ctrl = ctrl, c_word = word, c_cell = 0, c_comp = 0, c_head = 0;
  double c_enth = 0.0, rate = -1.0, limit = 0.0;
  WbSources S = {};
  S.head = &c_head; S.cell = &c_cell; S.comp = &c_comp; S.rate = &rate; S.enth = &c_enth; S.n = 1;
  S.ctrl = &c_ctrl; S.pi = &pi; S.pref = &pref; S.limit = &limit;
  S.ptab_n = &c_word; S.ptab = table;
  out[0] = wb_source_rate(S, 0, s);
  return 0;
}
extern "C" int hc_source_rate_ptab(const wb_params *prm, const double *primary, int region, const double *rock8, int ctrl,
                                   double pi, double pref, int word, const double *table, double *out) {
  if (prm->eos == WB_EOS_WE) return source_rate_ptab_host<WB_EOS_WE>(prm, primary, region, rock8, ctrl, pi, pref, word, table, out);
  if (prm->eos == WB_EOS_WCE || prm->eos == WB_EOS_WAE)
    return source_rate_ptab_host<WB_EOS_WCE>(prm, primary, region, rock8, ctrl, pi, pref, word, table, out);
  return -2;
}
// separator_stage_init with the device thermodynamics (what wb_separator_stage runs on the host side of the library)
extern "C" int hc_separator_stage(int thermo, double pressure, double *hw, double *hs) {
  WbThermo th = wb_thermo_make(thermo, 0);
  double ts = 0.0, rho = 0.0, u = 0.0;
  if (wb_saturation_temperature(th, pressure, ts)) return 1;
  if (wb_region_properties(th, 1, pressure, ts, rho, u)) return 1;
  *hw = u + pressure / rho;
  if (wb_region_properties(th, 2, pressure, ts, rho, u)) return 1;
  *hs = u + pressure / rho;
  return 0;
}
