// wb_flow.cu -- function evaluation and Jacobian assembly kernels:
// fluid properties (K1), cell balances (K2), cell inflows (K3), BE residual,
// local finite-difference BAIJ Jacobian (K4), phase transitions (K9), scaled
// max norm (K8).  See DESIGN.md for the data layout and the roofline of each.
#include <algorithm>

#include "wb_common.cuh"
#include "wb_state.cuh"

void wb_free_mesh(wb_ctx *c);
void wb_newton_invalidate_pc(wb_ctx *c);

// MATMFFD_DS step (PETSc MatFDColoringApply, doc/user/setup_time.rst:434-452)
__device__ __forceinline__ double fd_step(double yj, double err, double umin) {
  double dx = yj;
  if (dx == 0.0) dx = 1.0;
  if (fabs(dx) < umin && dx >= 0.0) dx = umin;
  else if (dx < 0.0 && fabs(dx) < umin) dx = -umin;
  return dx * err;
}

// ---------------------------------------------------------------- K1+K2: properties and balances

struct EosArgs {
  const double *y;        // [ninterior*np] scaled primaries incl. partition ghosts
  const int32_t *region;  // [ncell]
  const double *rockp;    // SoA [5][ncell]
  double *state;          // [(np+1)][nf][ncell]
  double *Lvar;           // [(np+1)][np][nowned]
  double *dx;             // [np][ninterior]
  int *flags;
  int ncell, ninterior, nowned;
  int slot0;      // destination slot of variant 0
  int variant0;   // first variant handled by blockIdx.y == 0
  double fd_err, fd_umin;
};

// thread per (cell, variant): variant 0 evaluates y, variant v > 0 evaluates y + h e_{v-1}
// (src/flow_simulation.F90:2291-2415 fluid_properties + :1242-1330 cell_balances)
template <int EOS>
__global__ void __launch_bounds__(128) k_eos(const __grid_constant__ WbEosParams e, const EosArgs a) {
  constexpr int NP = WbEosTraits<EOS>::NP, NC = WbEosTraits<EOS>::NC, NPH = WbEosTraits<EOS>::NPH;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.ninterior) return;
  const int v = a.variant0 + blockIdx.y;
  const int slot = a.slot0 + blockIdx.y;
  double yv[NP], primary[NP];
#pragma unroll
  for (int k = 0; k < NP; k++) yv[k] = a.y[(size_t)c * NP + k];
  if (v > 0) {
    const double dx = fd_step(yv[v - 1], a.fd_err, a.fd_umin);
    yv[v - 1] += dx;
    a.dx[(size_t)(v - 1) * a.ninterior + c] = dx;
  }
  const int region = a.region[c];
  wb_unscale<NP>(e, yv, region, primary);
  WbFluid<NC, NPH> fl = {};
  fl.region = region;
  const int err = wb_eos_properties<EOS>(e, primary, fl);
  if (err) {
    atomicMax(&a.flags[0], 1);
    return;
  }
  const size_t nc = a.ncell;
  WbCellState<NC, NPH> s;
  wb_state_from_fluid(fl, a.rockp[3 * nc + c], a.rockp[4 * nc + c], s);
  store_state(a.state + (size_t)slot * WbStateLayout<NC, NPH>::NF * nc, nc, c, s);
  if (c < a.nowned) {
    double bal[NP];
    wb_cell_balance<NP, NC, NPH>(fl, a.rockp[c], a.rockp[nc + c], a.rockp[2 * nc + c], bal);
#pragma unroll
    for (int k = 0; k < NP; k++) a.Lvar[((size_t)slot * NP + k) * a.nowned + c] = bal[k];
  }
}

// full fluid records in the reference's AoS layout (src/fluid.F90:232-267), for output / parity
struct RecordArgs {
  const double *y;          // scaled primaries [ninterior*np]
  const double *bprimary;   // unscaled primaries of boundary ghosts [(ncell-ninterior)*np]
  const int32_t *region, *old_region;
  double *fluid;            // [ncell*dof]
  int ncell, ninterior, dof;
};
template <int EOS>
__global__ void k_fluid_record(const __grid_constant__ WbEosParams e, const RecordArgs a) {
  constexpr int NP = WbEosTraits<EOS>::NP, NC = WbEosTraits<EOS>::NC, NPH = WbEosTraits<EOS>::NPH;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.ncell) return;
  double primary[NP];
  const int region = a.region[c];
  if (c < a.ninterior) {
    double yv[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) yv[k] = a.y[(size_t)c * NP + k];
    wb_unscale<NP>(e, yv, region, primary);
  } else {
#pragma unroll
    for (int k = 0; k < NP; k++) primary[k] = a.bprimary[(size_t)(c - a.ninterior) * NP + k];
  }
  WbFluid<NC, NPH> fl = {};
  fl.region = region;
  wb_eos_properties<EOS>(e, primary, fl);
  double *r = a.fluid + (size_t)c * a.dof;
  r[0] = fl.P;
  r[1] = fl.T;
  r[2] = (double)region;
  r[3] = (double)a.old_region[c];
  r[4] = (double)fl.phases;
  r[5] = 1.0;  // permeability_factor (src/eos_we.F90:354)
#pragma unroll
  for (int k = 0; k < NC; k++) r[6 + k] = fl.pp[k];
#pragma unroll
  for (int p = 0; p < NPH; p++) {
    double *q = r + 6 + NC + p * (7 + NC);
    q[0] = fl.ph[p].rho;
    q[1] = fl.ph[p].mu;
    q[2] = fl.ph[p].sat;
    q[3] = fl.ph[p].kr;
    q[4] = fl.ph[p].pc;
    q[5] = fl.ph[p].h;
    q[6] = fl.ph[p].u;
#pragma unroll
    for (int k = 0; k < NC; k++) q[7 + k] = fl.ph[p].X[k];
  }
}

// boundary ghost cells: state from unscaled primaries (src/mesh.F90:1185-1202)
struct BoundaryArgs {
  const int32_t *cells;    // ghost cell indices
  const double *primary;   // [n*np] unscaled
  const int32_t *region;   // [n]
  const double *rockp;
  double *state;
  int32_t *region_out;
  double *bprimary;
  int *flags;
  int n, ncell, ninterior, nslots;
};
template <int EOS>
__global__ void k_boundary(const __grid_constant__ WbEosParams e, const BoundaryArgs a) {
  constexpr int NP = WbEosTraits<EOS>::NP, NC = WbEosTraits<EOS>::NC, NPH = WbEosTraits<EOS>::NPH;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const int c = a.cells[i];
  double primary[NP];
#pragma unroll
  for (int k = 0; k < NP; k++) {
    primary[k] = a.primary[(size_t)i * NP + k];
    a.bprimary[(size_t)(c - a.ninterior) * NP + k] = primary[k];
  }
  WbFluid<NC, NPH> fl = {};
  fl.region = a.region[i];
  a.region_out[c] = fl.region;
  if (wb_eos_properties<EOS>(e, primary, fl)) {
    atomicMax(&a.flags[0], 1);
    return;
  }
  const size_t nc = a.ncell;
  WbCellState<NC, NPH> s;
  wb_state_from_fluid(fl, a.rockp[3 * nc + c], a.rockp[4 * nc + c], s);
  for (int slot = 0; slot < a.nslots; slot++)
    store_state(a.state + (size_t)slot * WbStateLayout<NC, NPH>::NF * nc, nc, c, s);
}

// ---------------------------------------------------------------- time-stepping residual forms

// residual of the method timestepper.F90 selects (context%residual):
//   backward Euler  (:345-374)  r = L - L_last - dt R                        [VecCopy, VecAXPY, VecAXPY]
//   BDF2            (:378-427)  r = (1+2q) L - (q+1)^2 L_last + q^2 L_last2 - dt (q+1) R,  q = dt / dt_last
//                                                                            [VecCopy, VecScale, 3 x VecAXPY]
//   direct steady state (:431-452)  r = R
// evaluated entry by entry in the reference's operation order
struct WbResForm {
  int method;               // WB_METHOD_*
  double a0, a1, a2, cR;    // scale of L (BDF2 only), of L_last, of L_last2, of R
  const double *l1, *l2;    // L_last, L_last2 (AoS [nowned*np])
};
__device__ __forceinline__ double wb_form_residual(const WbResForm &f, double L, double R, size_t idx) {
  if (f.method == WB_METHOD_DIRECTSS) return R;
  double r = L;
  if (f.method == WB_METHOD_BDF2) r = r * f.a0;
  r = r + f.a1 * f.l1[idx];
  if (f.method == WB_METHOD_BDF2) r = r + f.a2 * f.l2[idx];
  r = r + f.cR * R;
  return r;
}

// ---------------------------------------------------------------- sources / sinks

// adds the inflow of every source of cell i, in source order (src/source_network.F90:296-355: inflow += flow / V);
// source%update_flow src/source.F90:457-480: injection :385-399, production by mobility-weighted phase flow
// fractions :403-438 (src/fluid.F90:374-456), energy :442-453
template <int NP, int NC, int NPH>
__device__ __forceinline__ void source_terms(const WbSources &S, int i, const WbCellState<NC, NPH> &s, double vol,
                                             double *acc) {
  if (!S.head) return;
  int k = S.head[i];
  if (k < 0) return;
  for (; k < S.n && S.cell[k] == i; k++) {
    const double rate = wb_source_rate(S, k, s);
    const int component = wb_source_component(S.comp[k], rate);
    double flow[NP], enthalpy = 0.0;
#pragma unroll
    for (int q = 0; q < NP; q++) flow[q] = 0.0;
    if (rate > 0.0) {
      if (component > 0) {
        enthalpy = S.enth[k];
#pragma unroll
        for (int q = 0; q < NP; q++)
          if (q == component - 1) flow[q] = rate;
      }
    } else {
      double frac[NPH];
#pragma unroll
      for (int p = 0; p < NPH; p++) frac[p] = 0.0;
      if (component < NP) {
        double sum = 0.0;
#pragma unroll
        for (int p = 0; p < NPH; p++) {
          if (s.phases & (1 << p)) frac[p] = s.mob[p];
          sum += frac[p];
        }
#pragma unroll
        for (int p = 0; p < NPH; p++) frac[p] = frac[p] / sum;
        if (NP != NC) {
#pragma unroll
          for (int p = 0; p < NPH; p++)
            if (s.phases & (1 << p)) enthalpy = enthalpy + frac[p] * s.h[p];
        }
      }
      if (component <= 0) {
        double cf[NC], csum = 0.0;
#pragma unroll
        for (int c = 0; c < NC; c++) {
          cf[c] = 0.0;
#pragma unroll
          for (int p = 0; p < NPH; p++)
            if (s.phases & (1 << p)) cf[c] = cf[c] + frac[p] * s.X[p][c];
          csum += cf[c];
        }
#pragma unroll
        for (int c = 0; c < NC; c++) flow[c] = rate * (cf[c] / csum);
      } else {
#pragma unroll
        for (int q = 0; q < NP; q++)
          if (q == component - 1) flow[q] = rate;
      }
    }
    if (NP != NC && component < NP) flow[NP - 1] = flow[NP - 1] + enthalpy * rate;
#pragma unroll
    for (int q = 0; q < NP; q++) acc[q] = acc[q] + flow[q] / vol;
  }
}

// ---------------------------------------------------------------- K3: inflows + BE residual

struct ResidualArgs {
  const double *state;  // slot base
  const double *Lvar;   // slot base [np][nowned]
  const double *face;   // SoA [6][nface]
  const double *vol;
  const int32_t *cf_ptr, *cf_face, *cf_other;
  WbResForm form;          // r is formed when r != null
  double *lhs, *rhs, *r;   // any may be null; AoS [nowned*np]
  double dt;
  int ncell, nowned, nface;
  WbSources src;
};

// inflow term of one face seen from cell i (src/flow_simulation.F90:1445-1455):
// sign * (flux * area) / volume with sign -1 when i is the face's first cell
template <int NP, int NC, int NPH>
__device__ __forceinline__ void face_term(const WbFaceGeom &g, int side, const WbCellState<NC, NPH> &si,
                                          const WbCellState<NC, NPH> &so, double vol, double *term) {
  double flux[NP];
  if (side == 0) wb_face_flux<NP, NC, NPH>(g, si, so, flux, nullptr);
  else wb_face_flux<NP, NC, NPH>(g, so, si, flux, nullptr);
  const double sign = side == 0 ? -1.0 : 1.0;
#pragma unroll
  for (int k = 0; k < NP; k++) {
    const double flow = flux[k] * g.area;
    term[k] = sign * flow / vol;
  }
}

// thread per owned cell: gathers its faces in ascending face order, which reproduces the
// summation order of the reference's sequential face loop without atomics
template <int EOS>
__global__ void __launch_bounds__(128) k_residual(const ResidualArgs a) {
  constexpr int NP = WbEosTraits<EOS>::NP, NC = WbEosTraits<EOS>::NC, NPH = WbEosTraits<EOS>::NPH;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.nowned) return;
  WbCellState<NC, NPH> si, so;
  load_state(a.state, a.ncell, i, si);
  const double vol = a.vol[i];
  double acc[NP];
#pragma unroll
  for (int k = 0; k < NP; k++) acc[k] = 0.0;
  const int e0 = a.cf_ptr[i], e1 = a.cf_ptr[i + 1];
  for (int e = e0; e < e1; e++) {
    const int fs = a.cf_face[e];
    const WbFaceGeom g = load_face(a.face, a.nface, fs >> 1);
    load_state(a.state, a.ncell, a.cf_other[e], so);
    double term[NP];
    face_term<NP, NC, NPH>(g, fs & 1, si, so, vol, term);
#pragma unroll
    for (int k = 0; k < NP; k++) acc[k] = acc[k] + term[k];
  }
  source_terms<NP, NC, NPH>(a.src, i, si, vol, acc);
#pragma unroll
  for (int k = 0; k < NP; k++) {
    const double L = a.Lvar[(size_t)k * a.nowned + i];
    if (a.lhs) a.lhs[(size_t)i * NP + k] = L;
    if (a.rhs) a.rhs[(size_t)i * NP + k] = acc[k];
    if (a.r) a.r[(size_t)i * NP + k] = wb_form_residual(a.form, L, acc[k], (size_t)i * NP + k);
  }
}

// ---------------------------------------------------------------- K4: local FD Jacobian

struct JacArgs {
  const double *state;  // [(np+1)][nf][ncell]
  const double *Lvar;   // [(np+1)][np][nowned]
  const double *dx;     // [np][ninterior]
  const double *face, *vol;
  WbResForm form;
  const int32_t *cf_ptr, *cf_face, *cf_other, *cf_bpos, *diagpos;
  double *val;  // BAIJ blocks, column-major bs x bs
  double dt;
  int ncell, ninterior, nowned, nface;
  WbSources src;
};

// Thread per owned cell (= block row).  Row i of the FD-coloured Jacobian only ever sees one
// perturbed column per colour (distance-2 colouring), so F_i(y + h_j e_j) can be formed
// locally: perturbing cell i's own variable re-evaluates its balance and all its faces,
// perturbing neighbour j re-evaluates the one shared face.  Sums are re-run in face order so
// F' and F carry the same rounding, as in the colouring loop.
template <int EOS, int MAXDEG>
__global__ void __launch_bounds__(128) k_jacobian(const JacArgs a) {
  constexpr int NP = WbEosTraits<EOS>::NP, NC = WbEosTraits<EOS>::NC, NPH = WbEosTraits<EOS>::NPH;
  constexpr int NF = WbStateLayout<NC, NPH>::NF;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.nowned) return;
  const size_t nc = a.ncell, slot_sz = (size_t)NF * nc;
  const int e0 = a.cf_ptr[i];
  const int deg = a.cf_ptr[i + 1] - e0;
  const double vol = a.vol[i];
  WbCellState<NC, NPH> s0, sv, so;
  load_state(a.state, nc, i, s0);

  double t[MAXDEG][NP];  // base inflow terms, face order
  double L0[NP], F0[NP];
#pragma unroll
  for (int k = 0; k < NP; k++) L0[k] = a.Lvar[(size_t)k * a.nowned + i];
  {
    double acc[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) acc[k] = 0.0;
#pragma unroll
    for (int m = 0; m < MAXDEG; m++) {
      if (m < deg) {
        const int fs = a.cf_face[e0 + m];
        const WbFaceGeom g = load_face(a.face, a.nface, fs >> 1);
        load_state(a.state, nc, a.cf_other[e0 + m], so);
        face_term<NP, NC, NPH>(g, fs & 1, s0, so, vol, t[m]);
#pragma unroll
        for (int k = 0; k < NP; k++) acc[k] = acc[k] + t[m][k];
      }
    }
    source_terms<NP, NC, NPH>(a.src, i, s0, vol, acc);
#pragma unroll
    for (int k = 0; k < NP; k++) F0[k] = wb_form_residual(a.form, L0[k], acc[k], (size_t)i * NP + k);
  }

  // diagonal block: own variable v perturbed
  {
    double blk[NP * NP];
#pragma unroll
    for (int v = 0; v < NP; v++) {
      load_state(a.state + (size_t)(v + 1) * slot_sz, nc, i, sv);
      double acc[NP];
#pragma unroll
      for (int k = 0; k < NP; k++) acc[k] = 0.0;
#pragma unroll
      for (int m = 0; m < MAXDEG; m++) {
        if (m < deg) {
          const int fs = a.cf_face[e0 + m];
          const WbFaceGeom g = load_face(a.face, a.nface, fs >> 1);
          load_state(a.state, nc, a.cf_other[e0 + m], so);
          double term[NP];
          face_term<NP, NC, NPH>(g, fs & 1, sv, so, vol, term);
#pragma unroll
          for (int k = 0; k < NP; k++) acc[k] = acc[k] + term[k];
        }
      }
      source_terms<NP, NC, NPH>(a.src, i, sv, vol, acc);
      const double vscale = 1.0 / a.dx[(size_t)v * a.ninterior + i];
#pragma unroll
      for (int k = 0; k < NP; k++) {
        const double Lv = a.Lvar[((size_t)(v + 1) * NP + k) * a.nowned + i];
        const double Fp = wb_form_residual(a.form, Lv, acc[k], (size_t)i * NP + k);
        blk[v * NP + k] = (Fp + (-1.0) * F0[k]) * vscale;
      }
    }
    double *out = a.val + (size_t)a.diagpos[i] * NP * NP;
#pragma unroll
    for (int q = 0; q < NP * NP; q++) out[q] = blk[q];
  }

  // off-diagonal blocks: neighbour's variable v perturbed, only the shared face changes
#pragma unroll
  for (int m = 0; m < MAXDEG; m++) {
    if (m < deg) {
      const int bpos = a.cf_bpos[e0 + m];
      if (bpos >= 0) {
        const int o = a.cf_other[e0 + m];
        const int fs = a.cf_face[e0 + m];
        const WbFaceGeom g = load_face(a.face, a.nface, fs >> 1);
        double blk[NP * NP];
#pragma unroll
        for (int v = 0; v < NP; v++) {
          load_state(a.state + (size_t)(v + 1) * slot_sz, nc, o, so);
          double term[NP];
          face_term<NP, NC, NPH>(g, fs & 1, s0, so, vol, term);
          double acc[NP];
#pragma unroll
          for (int k = 0; k < NP; k++) acc[k] = 0.0;
#pragma unroll
          for (int q = 0; q < MAXDEG; q++) {
            if (q < deg) {
#pragma unroll
              for (int k = 0; k < NP; k++) acc[k] = acc[k] + (q == m ? term[k] : t[q][k]);
            }
          }
          source_terms<NP, NC, NPH>(a.src, i, s0, vol, acc);
          const double vscale = 1.0 / a.dx[(size_t)v * a.ninterior + o];
#pragma unroll
          for (int k = 0; k < NP; k++) {
            const double Fp = wb_form_residual(a.form, L0[k], acc[k], (size_t)i * NP + k);
            blk[v * NP + k] = (Fp + (-1.0) * F0[k]) * vscale;
          }
        }
        double *out = a.val + (size_t)bpos * NP * NP;
#pragma unroll
        for (int q = 0; q < NP * NP; q++) out[q] = blk[q];
      }
    }
  }
}

// The same assembly with EIGHT LANES PER BLOCK ROW: lane m of a row evaluates face m of the cell -- its base inflow
// term, the terms with the cell's own variables perturbed (for the diagonal block) and the terms with the neighbour's
// variables perturbed (for its off-diagonal block): 1 + 2 np flux evaluations per lane instead of (1 + 2 np) x faces
// per thread, seven times as many threads with less state each, so the long dependent chains of the flux function
// overlap across lanes instead of queueing in one thread.  The face sums are then formed in ascending face order from
// the lanes' terms with width-8 shuffles -- the order of the reference's sequential face loop, so F' and F round as
// in the colouring loop and the blocks are bit-identical to k_jacobian's.  The row's blocks are adjacent in the BAIJ
// value array: the lanes' 32-byte (bs = 2) block stores of a row form one contiguous span.
template <int EOS>
__global__ void __launch_bounds__(256) k_jacobian_lanes(const JacArgs a) {
  constexpr int NP = WbEosTraits<EOS>::NP, NC = WbEosTraits<EOS>::NC, NPH = WbEosTraits<EOS>::NPH;
  constexpr int NF = WbStateLayout<NC, NPH>::NF;
  const int m = threadIdx.x & 7;
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 3);
  const bool row_ok = row < a.nowned;
  const int i = row_ok ? row : a.nowned - 1;  // lanes of rows past the end compute on the last row and store nothing
  const size_t nc = a.ncell, slot_sz = (size_t)NF * nc;
  const int e0 = a.cf_ptr[i];
  const int deg = a.cf_ptr[i + 1] - e0;
  const bool has = m < deg;
  const double vol = a.vol[i];
  WbCellState<NC, NPH> s0, sv, so;
  load_state(a.state, nc, i, s0);
  const int e = e0 + (has ? m : 0);
  const int fs = a.cf_face[e], o = a.cf_other[e], bpos = has ? a.cf_bpos[e] : -1;
  const WbFaceGeom g = load_face(a.face, a.nface, fs >> 1);
  double t[NP], d[NP][NP], ot[NP][NP];
  load_state(a.state, nc, o, so);
  face_term<NP, NC, NPH>(g, fs & 1, s0, so, vol, t);
#pragma unroll
  for (int v = 0; v < NP; v++) {  // own variable v perturbed: this face's term of the diagonal block's column v
    load_state(a.state + (size_t)(v + 1) * slot_sz, nc, i, sv);
    face_term<NP, NC, NPH>(g, fs & 1, sv, so, vol, d[v]);
  }
#pragma unroll
  for (int v = 0; v < NP; v++) {  // the neighbour's variable v perturbed: only this face changes
    if (bpos >= 0) {
      load_state(a.state + (size_t)(v + 1) * slot_sz, nc, o, so);
      face_term<NP, NC, NPH>(g, fs & 1, s0, so, vol, ot[v]);
    } else {
#pragma unroll
      for (int k = 0; k < NP; k++) ot[v][k] = 0.0;
    }
  }
  // face sums in ascending face order
  double accF[NP], accD[NP][NP], accO[NP][NP];
#pragma unroll
  for (int k = 0; k < NP; k++) {
    accF[k] = 0.0;
#pragma unroll
    for (int v = 0; v < NP; v++) {
      accD[v][k] = 0.0;
      accO[v][k] = 0.0;
    }
  }
#pragma unroll
  for (int q = 0; q < 8; q++) {
    double tq[NP], dq[NP][NP];
#pragma unroll
    for (int k = 0; k < NP; k++) {
      tq[k] = __shfl_sync(0xffffffffu, t[k], q, 8);
#pragma unroll
      for (int v = 0; v < NP; v++) dq[v][k] = __shfl_sync(0xffffffffu, d[v][k], q, 8);
    }
    if (q < deg) {
#pragma unroll
      for (int k = 0; k < NP; k++) {
        accF[k] = accF[k] + tq[k];
#pragma unroll
        for (int v = 0; v < NP; v++) {
          accD[v][k] = accD[v][k] + dq[v][k];
          accO[v][k] = accO[v][k] + (q == m ? ot[v][k] : tq[k]);
        }
      }
    }
  }
  double L0[NP], F0[NP];
#pragma unroll
  for (int k = 0; k < NP; k++) L0[k] = a.Lvar[(size_t)k * a.nowned + i];
  source_terms<NP, NC, NPH>(a.src, i, s0, vol, accF);
#pragma unroll
  for (int k = 0; k < NP; k++) F0[k] = wb_form_residual(a.form, L0[k], accF[k], (size_t)i * NP + k);
  if (m == 0) {  // diagonal block
    double blk[NP * NP];
#pragma unroll
    for (int v = 0; v < NP; v++) {
      load_state(a.state + (size_t)(v + 1) * slot_sz, nc, i, sv);
      source_terms<NP, NC, NPH>(a.src, i, sv, vol, accD[v]);
      const double vscale = 1.0 / a.dx[(size_t)v * a.ninterior + i];
#pragma unroll
      for (int k = 0; k < NP; k++) {
        const double Lv = a.Lvar[((size_t)(v + 1) * NP + k) * a.nowned + i];
        const double Fp = wb_form_residual(a.form, Lv, accD[v][k], (size_t)i * NP + k);
        blk[v * NP + k] = (Fp + (-1.0) * F0[k]) * vscale;
      }
    }
    if (row_ok) {
      double *out = a.val + (size_t)a.diagpos[i] * NP * NP;
#pragma unroll
      for (int q = 0; q < NP * NP; q++) out[q] = blk[q];
    }
  }
  if (bpos >= 0) {  // this lane's off-diagonal block
    double blk[NP * NP];
#pragma unroll
    for (int v = 0; v < NP; v++) {
      source_terms<NP, NC, NPH>(a.src, i, s0, vol, accO[v]);
      const double vscale = 1.0 / a.dx[(size_t)v * a.ninterior + o];
#pragma unroll
      for (int k = 0; k < NP; k++) {
        const double Fp = wb_form_residual(a.form, L0[k], accO[v][k], (size_t)i * NP + k);
        blk[v * NP + k] = (Fp + (-1.0) * F0[k]) * vscale;
      }
    }
    if (row_ok) {
      double *out = a.val + (size_t)bpos * NP * NP;
#pragma unroll
      for (int q = 0; q < NP * NP; q++) out[q] = blk[q];
    }
  }
}

// ---------------------------------------------------------------- K9: transitions

struct TransArgs {
  const double *y_old;
  double *y, *search;
  int32_t *region, *old_region;
  const int32_t *region_iter;
  const double *T_iter;
  int *flags;  // [0] err, [1] changed_search, [2] last cell changed
  int nowned;
};
// src/flow_simulation.F90:2419-2576, thread per owned cell
template <int EOS>
__global__ void k_transitions(const __grid_constant__ WbEosParams e, const TransArgs a) {
  constexpr int NP = WbEosTraits<EOS>::NP;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.nowned) return;
  const int region = a.region[c];
  a.old_region[c] = region;  // :2502
  if (EOS == WB_EOS_W) {
    const double p = a.y[c] * e.scale[0][region];
    if (p < 0.0 || p > 100.e6) atomicMax(&a.flags[0], 1);
    return;
  } else {
    const int oreg = a.region_iter[c];
    double yn[NP], yo[NP], primary[NP], old_primary[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) {
      yn[k] = a.y[(size_t)c * NP + k];
      yo[k] = a.y_old[(size_t)c * NP + k];
    }
    wb_unscale<NP>(e, yn, region, primary);
    wb_unscale<NP>(e, yo, oreg, old_primary);
    int new_region = region;
    bool transition = false;
    int err;
    bool changed = false;
    if (EOS == WB_EOS_WCE) {
      err = wb_wge_transition(e.thermo, old_primary, primary, oreg, a.T_iter[c], new_region, transition);
      if (err == 0) err = wb_wge_check_primary(primary, new_region, changed);  // may clamp Pg (eos_wge.F90:594-606)
    } else {
      err = wb_we_transition(e.thermo, old_primary, primary, oreg, a.T_iter[c], new_region, transition);
      if (err == 0) err = wb_we_check_primary(primary, new_region);
    }
    if (err) {
      atomicMax(&a.flags[0], 1);
      return;
    }
    if (transition || changed) {  // changed_y: flow_simulation.F90:2521-2545
      a.region[c] = new_region;
      wb_scale<NP>(e, primary, new_region, yn);
#pragma unroll
      for (int k = 0; k < NP; k++) {
        a.y[(size_t)c * NP + k] = yn[k];
        a.search[(size_t)c * NP + k] = yo[k] - yn[k];  // :2538-2542
      }
      a.flags[1] = 1;
      if (c == a.nowned - 1) a.flags[2] = 1;
    }
  }
}

// ---------------------------------------------------------------- K8: scaled max norm

// max_i |v_i| / max(|s_i|, tol) with the first index attaining it (src/dm_utils.F90:644-685)
__global__ void k_max_scaled(const double *__restrict__ v, const double *__restrict__ s, double tol, int n,
                             double *__restrict__ part_val, int *__restrict__ part_idx) {
  double best = -1.0;
  int bi = 0x7fffffff;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double q = fabs(v[i]) / fmax(fabs(s[i]), tol);
    if (q > best) {
      best = q;
      bi = i;
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_down_sync(0xffffffffu, best, off);
    const int oi = __shfl_down_sync(0xffffffffu, bi, off);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  __shared__ double sv[32];
  __shared__ int si[32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    sv[w] = best;
    si[w] = bi;
  }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    best = l < nw ? sv[l] : -1.0;
    bi = l < nw ? si[l] : 0x7fffffff;
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, best, off);
      const int oi = __shfl_down_sync(0xffffffffu, bi, off);
      if (ov > best || (ov == best && oi < bi)) {
        best = ov;
        bi = oi;
      }
    }
    if (l == 0) {
      part_val[blockIdx.x] = best;
      part_idx[blockIdx.x] = bi;
    }
  }
}

// ---------------------------------------------------------------- small utility kernels

__global__ void k_face_perm(const int32_t *__restrict__ face_cells, const double *__restrict__ perm,
                            const int32_t *__restrict__ permdir, double *__restrict__ face, int nface, int ncell) {
  // harmonic average of the support cells' permeability along the face direction
  // (src/face.F90:381-398); permeability_factor is 1 for eos_we / eos_w
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nface) return;
  const int c1 = face_cells[2 * f], c2 = face_cells[2 * f + 1];
  const int d = permdir[f] - 1;
  const double k1 = perm[(size_t)d * ncell + c1] * 1.0, k2 = perm[(size_t)d * ncell + c2] * 1.0;
  const size_t nf = nface;
  face[5 * nf + f] = wb_harmonic(face[nf + f], face[2 * nf + f], face[3 * nf + f], k1, k2);
}

__global__ void k_copy_field(const double *__restrict__ src, double *__restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
__global__ void k_copy_i32(const int32_t *__restrict__ src, int32_t *__restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// y + region <-> halo message of (np+1) doubles per cell
__global__ void k_pack_yr(const double *__restrict__ y, const int32_t *__restrict__ region, int n, int np,
                          double *__restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  for (int k = 0; k < np; k++) out[(size_t)c * (np + 1) + k] = y[(size_t)c * np + k];
  out[(size_t)c * (np + 1) + np] = (double)region[c];
}
__global__ void k_unpack_yr(const double *__restrict__ in, int c0, int n, int np, double *__restrict__ y,
                            int32_t *__restrict__ region) {
  const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  for (int k = 0; k < np; k++) y[(size_t)c * np + k] = in[(size_t)c * (np + 1) + k];
  region[c] = (int)in[(size_t)c * (np + 1) + np];
}

// ================================================================ host side

#define DISPATCH_EOS(ctx, CALL)                                  \
  do {                                                           \
    if ((ctx)->eos.eos == WB_EOS_WE) { CALL(WB_EOS_WE); }        \
    else if ((ctx)->eos.eos == WB_EOS_WCE) { CALL(WB_EOS_WCE); } \
    else { CALL(WB_EOS_W); }                                     \
  } while (0)

// residual form of the context's time-stepping method for step size dt
static WbResForm wb_res_form(const wb_ctx *c, const double *d_lhs_last, double dt) {
  WbResForm f;
  f.method = c->method;
  f.l1 = d_lhs_last;
  f.l2 = c->d_lhs_last2;
  if (c->method == WB_METHOD_BDF2) {
    const double q = dt / c->dt_last, q1 = q + 1.0;
    f.a0 = 1.0 + 2.0 * q; f.a1 = -q1 * q1; f.a2 = q * q; f.cR = -dt * q1;
  } else {
    f.a0 = 1.0; f.a1 = -1.0; f.a2 = 0.0; f.cR = -dt;
  }
  return f;
}

WbSources wb_sources_args(const wb_ctx *c) {
  WbSources S = {c->nsrc > 0 ? c->d_src_head : nullptr, c->d_src_cell, c->d_src_comp, c->d_src_rate, c->d_src_enth, c->nsrc,
                 c->d_src_ctrl, c->d_src_pi, c->d_src_pref, c->d_src_limit,
                 c->d_src_sep_n, c->d_src_sep_h, c->d_src_limit_w, c->d_src_limit_s,
                 c->d_src_ptab_n, c->d_src_ptab};
  return S;
}

template <class T> static int dev_alloc(T **p, size_t n) {
  WB_CUDA(cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)));
  return 0;
}
template <class T> static int dev_upload(T **p, const std::vector<T> &v) {
  WB_TRY(dev_alloc(p, v.size()));
  if (!v.empty()) WB_CUDA(wb_memcpy_sync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

struct MeshDev {  // extra device arrays private to this file
  double *d_perm = nullptr;       // [3][ncell]
  int32_t *d_permdir = nullptr;   // [nface]
  double *d_bprimary = nullptr;   // [(ncell-ninterior)*np]
  int32_t *d_old_region = nullptr;
  double *d_yr = nullptr;         // halo message buffer [(ninterior)*(np+1)]
  double *d_w = nullptr, *d_F = nullptr;  // coloured-FD work vectors
};
static std::map<wb_ctx *, MeshDev> g_meshdev;
static MeshDev &meshdev(wb_ctx *c) {
  std::lock_guard<std::mutex> lk(wb_registry_mutex());
  return g_meshdev[c];  // node addresses of a std::map are stable
}

void wb_flow_release(wb_ctx *c) {
  std::lock_guard<std::mutex> lk(wb_registry_mutex());
  auto it = g_meshdev.find(c);
  if (it == g_meshdev.end()) return;
  MeshDev &md = it->second;
  cudaFree(md.d_perm); cudaFree(md.d_permdir); cudaFree(md.d_bprimary); cudaFree(md.d_old_region);
  cudaFree(md.d_yr); cudaFree(md.d_w); cudaFree(md.d_F);
  g_meshdev.erase(it);
}

static int recompute_face_perm(wb_ctx *c) {
  MeshDev &md = meshdev(c);
  if (c->nface == 0) return 0;
  k_face_perm<<<wb_grid(c->nface, 256), 256, 0, c->stream>>>(c->d_face_cells, md.d_perm, md.d_permdir, c->d_face,
                                                           c->nface, c->ncell);
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int wb_set_mesh(wb_ctx *c, int ncell, int ninterior, int nowned, int nface, const int32_t *face_cells,
                           const double *face_geom, const double *cell_geom, const double *rock) {
  WB_CHECK(c && face_geom && cell_geom && rock && (face_cells || nface == 0), "wb_set_mesh: null argument");
  WB_CHECK(nowned <= ninterior && ninterior <= ncell && nowned > 0, "wb_set_mesh: need 0 < nowned <= ninterior <= ncell");
  WB_CUDA(cudaSetDevice(c->device));
  wb_newton_invalidate_pc(c);
  wb_free_mesh(c);
  MeshDev &md = meshdev(c);
  cudaFree(md.d_perm); cudaFree(md.d_permdir); cudaFree(md.d_bprimary); cudaFree(md.d_old_region);
  cudaFree(md.d_yr); cudaFree(md.d_w); cudaFree(md.d_F);
  md = MeshDev();
  c->ncell = ncell; c->ninterior = ninterior; c->nowned = nowned; c->nface = nface;
  const int np = c->np;
  c->h_face_cells.assign(face_cells, face_cells + 2 * (size_t)nface);
  c->h_rock.assign(rock, rock + 8 * (size_t)ncell);

  // ---- face SoA + permeability direction
  std::vector<double> fsoa(6 * (size_t)std::max(nface, 1), 0.0);
  std::vector<int32_t> permdir(std::max(nface, 1), 1);
  for (int f = 0; f < nface; f++) {
    const double *g = face_geom + 12 * (size_t)f;
    WB_CHECK(face_cells[2 * f] >= 0 && face_cells[2 * f] < ncell && face_cells[2 * f + 1] >= 0 &&
                 face_cells[2 * f + 1] < ncell, "wb_set_mesh: face %d has a cell index out of range", f);
    fsoa[f] = g[0];
    fsoa[(size_t)nface + f] = g[1];
    fsoa[2 * (size_t)nface + f] = g[2];
    fsoa[3 * (size_t)nface + f] = g[3];
    fsoa[4 * (size_t)nface + f] = g[7];
    permdir[f] = (int32_t)(g[11] + 0.5);
    WB_CHECK(permdir[f] >= 1 && permdir[f] <= 3, "wb_set_mesh: face %d permeability direction %d", f, permdir[f]);
  }
  WB_TRY(dev_upload(&c->d_face, fsoa));
  WB_TRY(dev_upload(&md.d_permdir, permdir));
  WB_TRY(dev_upload(&c->d_face_cells, c->h_face_cells));

  // ---- cell arrays
  std::vector<double> vol(ncell), rockp(5 * (size_t)ncell), perm(3 * (size_t)ncell);
  for (int i = 0; i < ncell; i++) {
    vol[i] = cell_geom[4 * (size_t)i + 3];
    const double *r = rock + 8 * (size_t)i;
    rockp[i] = r[WB_R_POR];
    rockp[(size_t)ncell + i] = r[WB_R_RHO];
    rockp[2 * (size_t)ncell + i] = r[WB_R_CP];
    rockp[3 * (size_t)ncell + i] = r[WB_R_WET];
    rockp[4 * (size_t)ncell + i] = r[WB_R_DRY];
    for (int d = 0; d < 3; d++) perm[(size_t)d * ncell + i] = r[d];
  }
  WB_TRY(dev_upload(&c->d_vol, vol));
  WB_TRY(dev_upload(&c->d_rockp, rockp));
  WB_TRY(dev_upload(&md.d_perm, perm));

  // ---- cell -> faces CSR over owned cells, ascending face order
  c->h_cf_ptr.assign(nowned + 1, 0);
  for (int f = 0; f < nface; f++)
    for (int s = 0; s < 2; s++) {
      const int cc = face_cells[2 * f + s];
      if (cc < nowned) c->h_cf_ptr[cc + 1]++;
    }
  c->maxdeg = 0;
  for (int i = 0; i < nowned; i++) {
    c->maxdeg = std::max(c->maxdeg, c->h_cf_ptr[i + 1]);
    c->h_cf_ptr[i + 1] += c->h_cf_ptr[i];
  }
  c->ncf = c->h_cf_ptr[nowned];
  c->h_cf_face.assign(std::max(c->ncf, 1), 0);
  c->h_cf_other.assign(std::max(c->ncf, 1), 0);
  {
    std::vector<int32_t> fill(c->h_cf_ptr.begin(), c->h_cf_ptr.end() - 1);
    for (int f = 0; f < nface; f++)
      for (int s = 0; s < 2; s++) {
        const int cc = face_cells[2 * f + s];
        if (cc < nowned) {
          const int e = fill[cc]++;
          c->h_cf_face[e] = 2 * f + s;
          c->h_cf_other[e] = face_cells[2 * f + (1 - s)];
        }
      }
  }

  // ---- BAIJ pattern: row i = {i} U face neighbours with dofs (src/dm_utils.F90:1041-1051)
  wb_mat &J = c->J;
  J.ctx = c; J.nb = nowned; J.ncolb = ninterior; J.bs = np; J.owns = true;
  J.h_rowptr.assign(nowned + 1, 0);
  std::vector<int32_t> cols;
  cols.reserve((size_t)c->ncf + nowned);
  std::vector<int32_t> row;
  for (int i = 0; i < nowned; i++) {
    row.clear();
    row.push_back(i);
    for (int e = c->h_cf_ptr[i]; e < c->h_cf_ptr[i + 1]; e++)
      if (c->h_cf_other[e] < ninterior) row.push_back(c->h_cf_other[e]);
    std::sort(row.begin(), row.end());
    row.erase(std::unique(row.begin(), row.end()), row.end());
    cols.insert(cols.end(), row.begin(), row.end());
    J.h_rowptr[i + 1] = (int32_t)cols.size();
  }
  J.h_colidx = cols;
  J.nnzb = (int)cols.size();
  std::vector<int32_t> bpos(std::max(c->ncf, 1), -1), diagpos(nowned, -1);
  for (int i = 0; i < nowned; i++) {
    const int32_t *b = J.h_colidx.data() + J.h_rowptr[i], *e_ = J.h_colidx.data() + J.h_rowptr[i + 1];
    diagpos[i] = (int32_t)(std::lower_bound(b, e_, i) - J.h_colidx.data());
    for (int e = c->h_cf_ptr[i]; e < c->h_cf_ptr[i + 1]; e++) {
      const int o = c->h_cf_other[e];
      if (o < ninterior) bpos[e] = (int32_t)(std::lower_bound(b, e_, o) - J.h_colidx.data());
    }
  }
  WB_TRY(dev_upload(&c->d_cf_ptr, c->h_cf_ptr));
  WB_TRY(dev_upload(&c->d_cf_face, c->h_cf_face));
  WB_TRY(dev_upload(&c->d_cf_other, c->h_cf_other));
  WB_TRY(dev_upload(&c->d_cf_bpos, bpos));
  WB_TRY(dev_upload(&c->d_diagpos, diagpos));
  // (padded: the TMA-staged SpMV rounds its bulk copies to 16 bytes)
  WB_CUDA(cudaMalloc(&J.d_rowptr, sizeof(int32_t) * J.h_rowptr.size() + WB_PAD_BYTES));
  WB_CUDA(cudaMalloc(&J.d_colidx, sizeof(int32_t) * std::max<size_t>(J.h_colidx.size(), 1) + WB_PAD_BYTES));
  WB_CUDA(wb_memset_sync(J.d_colidx, 0, sizeof(int32_t) * std::max<size_t>(J.h_colidx.size(), 1) + WB_PAD_BYTES));
  WB_CUDA(wb_memcpy_sync(J.d_rowptr, J.h_rowptr.data(), sizeof(int32_t) * J.h_rowptr.size(), cudaMemcpyHostToDevice));
  WB_CUDA(wb_memcpy_sync(J.d_colidx, J.h_colidx.data(), sizeof(int32_t) * J.h_colidx.size(), cudaMemcpyHostToDevice));
  WB_CUDA(cudaMalloc(&J.d_val, (size_t)J.nnzb * np * np * sizeof(double) + WB_PAD_BYTES));
  WB_CUDA(wb_memset_sync(J.d_val, 0, (size_t)J.nnzb * np * np * sizeof(double) + WB_PAD_BYTES));
  WB_TRY(wb_mat_build_tiles(&J));
  WB_TRY(dev_alloc(&J.d_xloc, (size_t)(ninterior - nowned + 1) * np));  // ghost entries of x for the SpMV
  WB_CUDA(wb_memset_sync(J.d_xloc, 0, (size_t)(ninterior - nowned + 1) * np * sizeof(double)));

  // ---- state
  const size_t nslot = np + 1;
  WB_TRY(dev_alloc(&c->d_region, ncell));
  WB_TRY(dev_alloc(&c->d_region_iter, ncell));
  WB_TRY(dev_alloc(&c->d_region_step, ncell));
  WB_TRY(dev_alloc(&md.d_old_region, ncell));
  WB_TRY(dev_alloc(&c->d_T_iter, ncell));
  WB_TRY(dev_alloc(&c->d_T_step, ncell));
  WB_TRY(dev_alloc(&c->d_state, nslot * c->nf * ncell));
  WB_TRY(dev_alloc(&c->d_Lvar, nslot * np * nowned));
  WB_TRY(dev_alloc(&c->d_dx, (size_t)np * ninterior));
  WB_TRY(dev_alloc(&c->d_yloc, (size_t)ninterior * np));
  WB_TRY(dev_alloc(&c->d_balances, (size_t)nowned * np));
  WB_TRY(dev_alloc(&md.d_bprimary, (size_t)(ncell - ninterior) * np));
  WB_TRY(dev_alloc(&md.d_yr, (size_t)ninterior * (np + 1)));
  WB_CUDA(wb_memset_sync(c->d_region, 0, sizeof(int32_t) * ncell));
  WB_CUDA(wb_memset_sync(c->d_region_iter, 0, sizeof(int32_t) * ncell));
  WB_CUDA(wb_memset_sync(c->d_region_step, 0, sizeof(int32_t) * ncell));
  WB_CUDA(wb_memset_sync(md.d_old_region, 0, sizeof(int32_t) * ncell));
  WB_CUDA(wb_memset_sync(c->d_T_iter, 0, sizeof(double) * ncell));
  WB_CUDA(wb_memset_sync(c->d_state, 0, sizeof(double) * nslot * c->nf * ncell));
  WB_CUDA(wb_memset_sync(c->d_Lvar, 0, sizeof(double) * nslot * np * nowned));
  WB_CUDA(wb_memset_sync(c->d_yloc, 0, sizeof(double) * ninterior * np));
  WB_CUDA(wb_memset_sync(c->d_balances, 0, sizeof(double) * nowned * np));
  c->first_cell = 0;
  c->ncell_global = nowned;
  c->h_color.clear();
  c->ncolor = 0;
  WB_TRY(recompute_face_perm(c));
  WB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int wb_jacobian_pattern(wb_ctx *c, int *nb, int *bs, int *nnzb, const int32_t **rowptr,
                                   const int32_t **colidx, double **vals) {
  if (nb) *nb = c->J.nb;
  if (bs) *bs = c->J.bs;
  if (nnzb) *nnzb = c->J.nnzb;
  if (rowptr) *rowptr = c->J.d_rowptr;
  if (colidx) *colidx = c->J.d_colidx;
  if (vals) {
    *vals = c->J.d_val;
    c->J.external_vals = true;  // the caller may write values behind our back from now on
  }
  return 0;
}

extern "C" int wb_cell_faces_get(wb_ctx *c, int *ncf, int32_t *cf_ptr, int32_t *cf_face, int32_t *cf_other) {
  WB_CHECK(c->ncell > 0, "wb_cell_faces_get: no mesh");
  if (ncf) *ncf = c->ncf;
  if (cf_ptr) memcpy(cf_ptr, c->h_cf_ptr.data(), sizeof(int32_t) * (c->nowned + 1));
  if (cf_face) memcpy(cf_face, c->h_cf_face.data(), sizeof(int32_t) * c->ncf);
  if (cf_other) memcpy(cf_other, c->h_cf_other.data(), sizeof(int32_t) * c->ncf);
  return 0;
}

extern "C" int wb_jacobian_get(wb_ctx *c, int32_t *rowptr, int32_t *colidx, double *vals) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CUDA(cudaStreamSynchronize(c->stream));
  if (rowptr) memcpy(rowptr, c->J.h_rowptr.data(), sizeof(int32_t) * (c->J.nb + 1));
  if (colidx) memcpy(colidx, c->J.h_colidx.data(), sizeof(int32_t) * c->J.nnzb);
  if (vals) WB_CUDA(wb_memcpy_sync(vals, c->J.d_val, sizeof(double) * (size_t)c->J.nnzb * c->J.bs * c->J.bs,
                               cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int wb_jacobian_mat(wb_ctx *c, wb_mat **out) {
  *out = &c->J;
  return 0;
}

// ---- y (owned, device) -> d_yloc incl. ghosts (halo carries region too)
static int load_y(wb_ctx *c, const double *d_y) {
  const int np = c->np;
  if (d_y != c->d_yloc)
    WB_CUDA(cudaMemcpyAsync(c->d_yloc, d_y, sizeof(double) * (size_t)c->nowned * np, cudaMemcpyDeviceToDevice,
                            c->stream));
  if (c->nranks > 1 && c->halo.nneigh > 0) {
    MeshDev &md = meshdev(c);
    k_pack_yr<<<wb_grid(c->nowned, 256), 256, 0, c->stream>>>(c->d_yloc, c->d_region, c->nowned, np, md.d_yr);
    WB_LAUNCH(c);
    WB_TRY(wb_halo_exchange(c, md.d_yr, np + 1));
    const int ng = c->ninterior - c->nowned;
    if (ng > 0) {
      k_unpack_yr<<<wb_grid(ng, 256), 256, 0, c->stream>>>(md.d_yr, c->nowned, c->ninterior, np, c->d_yloc,
                                                         c->d_region);
      WB_LAUNCH(c);
    }
  }
  return 0;
}

// fluid properties + balances of variants [v0, v0+nv) into slots [slot0, ...)
static int launch_eos(wb_ctx *c, int slot0, int v0, int nv, double fd_err, double fd_umin) {
  EosArgs a;
  a.y = c->d_yloc; a.region = c->d_region; a.rockp = c->d_rockp; a.state = c->d_state; a.Lvar = c->d_Lvar;
  a.dx = c->d_dx; a.flags = c->d_flags; a.ncell = c->ncell; a.ninterior = c->ninterior; a.nowned = c->nowned;
  a.slot0 = slot0; a.variant0 = v0; a.fd_err = fd_err; a.fd_umin = fd_umin;
  dim3 grid(wb_grid(c->ninterior, 128), nv);
#define CALL(E) k_eos<E><<<grid, 128, 0, c->stream>>>(c->eos, a)
  DISPATCH_EOS(c, CALL);
#undef CALL
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  return 0;
}

static int launch_residual(wb_ctx *c, int slot, const double *d_lhs_last, double dt, double *d_lhs, double *d_rhs,
                           double *d_r) {
  ResidualArgs a;
  a.state = c->d_state + (size_t)slot * c->nf * c->ncell;
  a.Lvar = c->d_Lvar + (size_t)slot * c->np * c->nowned;
  a.face = c->d_face; a.vol = c->d_vol; a.cf_ptr = c->d_cf_ptr; a.cf_face = c->d_cf_face; a.cf_other = c->d_cf_other;
  a.form = wb_res_form(c, d_lhs_last, dt);
  a.lhs = d_lhs; a.rhs = d_rhs; a.r = (d_lhs_last || c->method == WB_METHOD_DIRECTSS) ? d_r : nullptr; a.dt = dt;
  a.ncell = c->ncell; a.nowned = c->nowned; a.nface = c->nface;
  a.src = wb_sources_args(c);
#define CALL(E) k_residual<E><<<wb_grid(c->nowned, 128), 128, 0, c->stream>>>(a)
  DISPATCH_EOS(c, CALL);
#undef CALL
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  return 0;
}

// ---- time-stepping method ----------------------------------------------------------------------
extern "C" int wb_set_method(wb_ctx *c, int method, double dt_last, const double *lhs_last2) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(method == WB_METHOD_BEULER || method == WB_METHOD_BDF2 || method == WB_METHOD_DIRECTSS,
           "wb_set_method: unknown method %d", method);
  if (method == WB_METHOD_BDF2) {
    WB_CHECK(c->ncell > 0 && lhs_last2 && dt_last > 0.0, "wb_set_method: BDF2 needs a mesh, dt_last > 0 and lhs_last2");
    const size_t n = (size_t)c->nowned * c->np;
    if (!c->d_lhs_last2) WB_CUDA(cudaMalloc(&c->d_lhs_last2, sizeof(double) * n));
    WB_CUDA(cudaMemcpyAsync(c->d_lhs_last2, lhs_last2, sizeof(double) * n, cudaMemcpyDefault, c->stream));
    WB_CUDA(cudaStreamSynchronize(c->stream));
    c->dt_last = dt_last;
  }
  c->method = method;
  return 0;
}

// ---- sources ---------------------------------------------------------------------------------
// (source_network%assemble_cell_inflows, src/flow_simulation.F90:1468-1473)
extern "C" int wb_set_sources(wb_ctx *c, int n, const int32_t *cell, const int32_t *component, const double *rate,
                              const double *enthalpy) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(c->ncell > 0, "wb_set_sources: no mesh");
  WB_CHECK(n <= 0 || (!wb_is_device_ptr(cell) && !wb_is_device_ptr(component) && !wb_is_device_ptr(rate) &&
                      !wb_is_device_ptr(enthalpy)),
           "wb_set_sources: source arrays are read on the host (set-up data): pass host arrays");
  WB_CUDA(cudaStreamSynchronize(c->stream));
  void *old[] = {c->d_src_head, c->d_src_cell, c->d_src_comp, c->d_src_rate, c->d_src_enth};
  for (void *p : old) cudaFree(p);
  c->d_src_head = c->d_src_cell = c->d_src_comp = nullptr;
  c->d_src_rate = c->d_src_enth = nullptr;
  c->nsrc = 0;
  c->h_src_order.clear();
  cudaFree(c->d_src_ctrl); cudaFree(c->d_src_pi); cudaFree(c->d_src_pref); cudaFree(c->d_src_limit);
  c->d_src_ctrl = nullptr;
  c->d_src_pi = c->d_src_pref = c->d_src_limit = nullptr;
  c->h_src_ctrl.clear(); c->h_src_pi.clear(); c->h_src_pref.clear(); c->h_src_limit.clear();
  cudaFree(c->d_src_sep_n); cudaFree(c->d_src_sep_h); cudaFree(c->d_src_limit_w); cudaFree(c->d_src_limit_s);
  c->d_src_sep_n = nullptr;
  c->d_src_sep_h = c->d_src_limit_w = c->d_src_limit_s = nullptr;
  cudaFree(c->d_src_ptab_n); cudaFree(c->d_src_ptab);
  c->d_src_ptab_n = nullptr;
  c->d_src_ptab = nullptr;
  cudaFree(c->d_trc_inj);  // tracer injection rates belong to the old source list
  c->d_trc_inj = nullptr;
  if (n <= 0) return 0;
  std::vector<int> order(n);
  for (int k = 0; k < n; k++) {
    WB_CHECK(cell[k] >= 0 && cell[k] < c->nowned, "wb_set_sources: source %d is not in an owned cell", k);
    WB_CHECK(component[k] >= 0 && component[k] <= c->np, "wb_set_sources: source %d: bad component %d", k, component[k]);
    order[k] = k;
  }
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cell[x] < cell[y]; });
  std::vector<int32_t> head(c->nowned, -1), sc(n), sk(n);
  std::vector<double> sr(n), se(n);
  for (int k = 0; k < n; k++) {
    const int o = order[k];
    sc[k] = cell[o]; sk[k] = component[o] | (component[o] << 8); sr[k] = rate[o]; se[k] = enthalpy[o];
    if (head[sc[k]] < 0) head[sc[k]] = k;
  }
  WB_TRY(dev_upload(&c->d_src_head, head));
  WB_TRY(dev_upload(&c->d_src_cell, sc));
  WB_TRY(dev_upload(&c->d_src_comp, sk));
  WB_TRY(dev_upload(&c->d_src_rate, sr));
  WB_TRY(dev_upload(&c->d_src_enth, se));
  c->nsrc = n;
  c->h_src_order = order;
  return 0;
}

// injection and production component of every source of the last wb_set_sources (get_components,
// src/source_setup.F90:2052-2083): which one applies is decided from the sign of the rate at every evaluation
extern "C" int wb_set_source_components(wb_ctx *c, int n, const int32_t *injection_component,
                                        const int32_t *production_component) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(n == c->nsrc && n > 0, "wb_set_source_components: %d components for %d sources", n, c->nsrc);
  WB_CUDA(cudaStreamSynchronize(c->stream));
  std::vector<int32_t> sk(n);
  for (int k = 0; k < n; k++) {
    const int o = c->h_src_order[k];
    WB_CHECK(injection_component[o] >= 1 && injection_component[o] <= c->np,
             "wb_set_source_components: source %d: bad injection component %d", o, injection_component[o]);
    WB_CHECK(production_component[o] >= 0 && production_component[o] <= c->np,
             "wb_set_source_components: source %d: bad production component %d", o, production_component[o]);
    sk[k] = injection_component[o] | (production_component[o] << 8);
  }
  WB_CUDA(wb_memcpy_sync(c->d_src_comp, sk.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice));
  return 0;
}

// ---- source controls -----------------------------------------------------------------------------
static int upload_source_controls(wb_ctx *c) {
  cudaFree(c->d_src_ctrl); cudaFree(c->d_src_pi); cudaFree(c->d_src_pref); cudaFree(c->d_src_limit);
  c->d_src_ctrl = nullptr;
  c->d_src_pi = c->d_src_pref = c->d_src_limit = nullptr;
  if (c->h_src_ctrl.empty()) return 0;
  WB_TRY(dev_upload(&c->d_src_ctrl, c->h_src_ctrl));
  WB_TRY(dev_upload(&c->d_src_pi, c->h_src_pi));
  WB_TRY(dev_upload(&c->d_src_pref, c->h_src_pref));
  WB_TRY(dev_upload(&c->d_src_limit, c->h_src_limit));
  return 0;
}
static void ensure_source_controls(wb_ctx *c) {
  if ((int)c->h_src_ctrl.size() == c->nsrc) return;
  c->h_src_ctrl.assign(c->nsrc, 0);
  c->h_src_pi.assign(c->nsrc, 0.0);
  c->h_src_pref.assign(c->nsrc, 0.0);
  c->h_src_limit.assign(c->nsrc, 0.0);
}

extern "C" int wb_set_source_controls(wb_ctx *c, int n, const int32_t *source, const double *productivity,
                                      const double *reference_pressure, const int32_t *direction, const double *limit) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CUDA(cudaStreamSynchronize(c->stream));
  c->h_src_ctrl.clear(); c->h_src_pi.clear(); c->h_src_pref.clear(); c->h_src_limit.clear();
  if (n <= 0) {
    if (c->d_src_sep_n) ensure_source_controls(c);  // separators stay in force: they are evaluated behind these arrays
    return upload_source_controls(c);
  }
  WB_CHECK(c->nsrc > 0, "wb_set_source_controls: no sources");
  WB_CHECK(!wb_is_device_ptr(source) && !wb_is_device_ptr(productivity) && !wb_is_device_ptr(reference_pressure) &&
               !wb_is_device_ptr(direction) && !wb_is_device_ptr(limit),
           "wb_set_source_controls: control arrays are read on the host (set-up data): pass host arrays");
  const int ns = c->nsrc;
  std::vector<int> pos(ns);  // input position -> sorted position
  for (int k = 0; k < ns; k++) pos[c->h_src_order[k]] = k;
  ensure_source_controls(c);
  for (int k = 0; k < n; k++) {
    WB_CHECK(source[k] >= 0 && source[k] < ns, "wb_set_source_controls: source index %d out of range", source[k]);
    const int d = direction ? direction[k] : 0;
    WB_CHECK(d >= 0 && d <= 2, "wb_set_source_controls: direction %d", d);
    const int q = pos[source[k]];
    c->h_src_ctrl[q] = (productivity[k] > 0.0 ? 1 : 0) | (d << 1);
    c->h_src_pi[q] = productivity[k];
    c->h_src_pref[q] = reference_pressure[k];
    c->h_src_limit[q] = limit ? limit[k] : 0.0;
  }
  return upload_source_controls(c);
}

// recharge / injectivity controls (recharge_source_control_iterator, src/source_control.F90:554-577; both input keys set
// up the same control, src/source_setup.F90:2984-3092): rate = -coefficient (P - reference pressure) of the source's
// cell, before the direction control and the limiters.  Edits the entries of these sources in the control arrays:
// call after wb_set_source_controls (which starts from scratch).
extern "C" int wb_set_source_recharge(wb_ctx *c, int n, const int32_t *source, const double *coefficient,
                                      const double *reference_pressure) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CUDA(cudaStreamSynchronize(c->stream));
  if (n <= 0) return 0;
  WB_CHECK(c->nsrc > 0, "wb_set_source_recharge: no sources");
  WB_CHECK(!wb_is_device_ptr(source) && !wb_is_device_ptr(coefficient) && !wb_is_device_ptr(reference_pressure),
           "wb_set_source_recharge: the arrays are read on the host (set-up data): pass host arrays");
  const int ns = c->nsrc;
  std::vector<int> pos(ns);
  for (int k = 0; k < ns; k++) pos[c->h_src_order[k]] = k;
  ensure_source_controls(c);
  for (int k = 0; k < n; k++) {
    WB_CHECK(source[k] >= 0 && source[k] < ns, "wb_set_source_recharge: source index %d out of range", source[k]);
    const int q = pos[source[k]];
    c->h_src_ctrl[q] = (c->h_src_ctrl[q] & 6) | 8;  // keeps the direction, replaces deliverability
    c->h_src_pi[q] = coefficient[k];
    c->h_src_pref[q] = reference_pressure[k];
  }
  return upload_source_controls(c);
}

// separator_stage_init (src/separator.F90:108-136) on the host: reference water and steam enthalpies u + P / rho on the
// saturation line at the stage's pressure, with the context's thermodynamic formulation
extern "C" int wb_separator_stage(wb_ctx *c, double pressure, double *ref_water_enthalpy, double *ref_steam_enthalpy) {
  double ts = 0.0, rho = 0.0, u = 0.0;
  int err = wb_saturation_temperature(c->eos.thermo, pressure, ts);
  if (err) return 1;
  if (wb_region_properties(c->eos.thermo, 1, pressure, ts, rho, u)) return 1;
  *ref_water_enthalpy = u + pressure / rho;
  if (wb_region_properties(c->eos.thermo, 2, pressure, ts, rho, u)) return 1;
  *ref_steam_enthalpy = u + pressure / rho;
  return 0;
}

extern "C" int wb_set_source_separators(wb_ctx *c, int n, const int32_t *source, const int32_t *nstage, const double *pressure,
                                        const double *limit_water, const double *limit_steam) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(c->d_src_sep_n); cudaFree(c->d_src_sep_h); cudaFree(c->d_src_limit_w); cudaFree(c->d_src_limit_s);
  c->d_src_sep_n = nullptr;
  c->d_src_sep_h = c->d_src_limit_w = c->d_src_limit_s = nullptr;
  if (n <= 0) return 0;
  WB_CHECK(c->nsrc > 0, "wb_set_source_separators: no sources");
  WB_CHECK(!wb_is_device_ptr(source) && !wb_is_device_ptr(nstage) && !wb_is_device_ptr(pressure) &&
               !wb_is_device_ptr(limit_water) && !wb_is_device_ptr(limit_steam),
           "wb_set_source_separators: the arrays are read on the host (set-up data): pass host arrays");
  const int ns = c->nsrc;
  if (!c->d_src_ctrl) {  // the rate evaluation looks at the separators only behind the control arrays: create empty ones
    ensure_source_controls(c);
    WB_TRY(upload_source_controls(c));
  }
  std::vector<int> pos(ns);  // input position -> sorted position
  for (int k = 0; k < ns; k++) pos[c->h_src_order[k]] = k;
  std::vector<int32_t> sn(ns, 0);
  std::vector<double> sh(4 * (size_t)ns, 0.0), lw(ns, 0.0), ls(ns, 0.0);
  for (int k = 0; k < n; k++) {
    WB_CHECK(source[k] >= 0 && source[k] < ns, "wb_set_source_separators: source index %d out of range", source[k]);
    WB_CHECK(nstage[k] >= 0 && nstage[k] <= 2, "wb_set_source_separators: %d separator stages (at most 2)", nstage[k]);
    const int q = pos[source[k]];
    sn[q] = nstage[k];
    for (int i = 0; i < nstage[k]; i++)
      WB_CHECK(wb_separator_stage(c, pressure[2 * k + i], &sh[4 * (size_t)q + 2 * i], &sh[4 * (size_t)q + 2 * i + 1]) == 0,
               "wb_set_source_separators: separator pressure %g outside the saturation line", pressure[2 * k + i]);
    lw[q] = limit_water ? limit_water[k] : 0.0;
    ls[q] = limit_steam ? limit_steam[k] : 0.0;
  }
  WB_TRY(dev_upload(&c->d_src_sep_n, sn));
  WB_TRY(dev_upload(&c->d_src_sep_h, sh));
  WB_TRY(dev_upload(&c->d_src_limit_w, lw));
  WB_TRY(dev_upload(&c->d_src_limit_s, ls));
  return 0;
}

// Reference pressure of sources on deliverability as a table against the flowing enthalpy (coordinate 0) or the pressure
// (1) of the source's cell: "deliverability": {"pressure": {"enthalpy": [[h, P], ...]}} of the input
// (deliverability_source_control_flow_rate, src/source_control.F90:376-388).  Like the separators the tables stay in force
// until the source list changes; n <= 0 removes them.
static_assert(WB_PTAB_MAX == WB_PRESSURE_TABLE_MAX, "table size of the header and of the device code");
extern "C" int wb_set_source_pressure_table(wb_ctx *c, int n, const int32_t *source, const int32_t *coordinate,
                                            const int32_t *step, const int32_t *npts, const double *table) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(c->d_src_ptab_n); cudaFree(c->d_src_ptab);
  c->d_src_ptab_n = nullptr;
  c->d_src_ptab = nullptr;
  if (n <= 0) return 0;
  WB_CHECK(c->nsrc > 0, "wb_set_source_pressure_table: no sources");
  WB_CHECK(source && npts && table, "wb_set_source_pressure_table: null array");
  WB_CHECK(!wb_is_device_ptr(source) && !wb_is_device_ptr(coordinate) && !wb_is_device_ptr(step) && !wb_is_device_ptr(npts) &&
               !wb_is_device_ptr(table),
           "wb_set_source_pressure_table: the arrays are read on the host (set-up data): pass host arrays");
  const int ns = c->nsrc;
  std::vector<int> pos(ns);  // input position -> sorted position
  for (int k = 0; k < ns; k++) pos[c->h_src_order[k]] = k;
  std::vector<int32_t> word(ns, 0);
  std::vector<double> tab(2 * WB_PTAB_MAX * (size_t)ns, 0.0);
  for (int k = 0; k < n; k++) {
    WB_CHECK(source[k] >= 0 && source[k] < ns, "wb_set_source_pressure_table: source index %d out of range", source[k]);
    WB_CHECK(npts[k] >= 1 && npts[k] <= WB_PTAB_MAX, "wb_set_source_pressure_table: %d points (1 .. %d)", npts[k], WB_PTAB_MAX);
    const int q = pos[source[k]];
    for (int i = 1; i < npts[k]; i++)
      WB_CHECK(table[2 * WB_PTAB_MAX * (size_t)k + 2 * i] > table[2 * WB_PTAB_MAX * (size_t)k + 2 * i - 2],
               "wb_set_source_pressure_table: the coordinates of source %d do not increase", source[k]);
    word[q] = npts[k] | ((coordinate && coordinate[k]) ? 256 : 0) | ((step && step[k]) ? 512 : 0);
    for (int i = 0; i < 2 * npts[k]; i++) tab[2 * WB_PTAB_MAX * (size_t)q + i] = table[2 * WB_PTAB_MAX * (size_t)k + i];
  }
  WB_TRY(dev_upload(&c->d_src_ptab_n, word));
  WB_TRY(dev_upload(&c->d_src_ptab, tab));
  return 0;
}

template <int EOS>
__global__ void k_source_separated(const WbSources S, const double *state, int ncell, const int32_t *order, double *out) {
  constexpr int NC = WbEosTraits<EOS>::NC, NPH = WbEosTraits<EOS>::NPH;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S.n) return;
  WbCellState<NC, NPH> s;
  load_state(state, (size_t)ncell, S.cell[k], s);
  double sep[5];
  wb_source_separated(S, k, s, wb_source_rate(S, k, s), sep);
  for (int i = 0; i < 5; i++) out[5 * (size_t)order[k] + i] = sep[i];
}

// the separated-flow output fields of the sources (water_rate, water_enthalpy, steam_rate, steam_enthalpy,
// steam_fraction: src/source_network_node.F90:95-112) at the state of the last unperturbed evaluation
extern "C" int wb_get_source_separated(wb_ctx *c, double *out5) {
  WB_CUDA(cudaSetDevice(c->device));
  if (c->nsrc == 0) return 0;
  int rc = 0;
  WbStage st(c);
  double *d_out = st.out(out5, 5 * (size_t)c->nsrc, &rc);
  if (rc) return rc;
  int32_t *d_order = nullptr;
  std::vector<int32_t> order(c->h_src_order.begin(), c->h_src_order.end());
  WB_TRY(dev_upload(&d_order, order));
  const WbSources S = wb_sources_args(c);
#define CALL(E) k_source_separated<E><<<wb_grid(c->nsrc, 128), 128, 0, c->stream>>>(S, c->d_state, c->ncell, d_order, d_out)
  DISPATCH_EOS(c, CALL);
#undef CALL
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  rc = st.finish();
  cudaFree(d_order);
  return rc;
}

template <int EOS>
__global__ void k_source_rates(const WbSources S, const double *state, int ncell, const int32_t *order, double *out) {
  constexpr int NC = WbEosTraits<EOS>::NC, NPH = WbEosTraits<EOS>::NPH;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S.n) return;
  WbCellState<NC, NPH> s;
  load_state(state, (size_t)ncell, S.cell[k], s);
  out[order[k]] = wb_source_rate(S, k, s);
}

extern "C" int wb_get_source_rates(wb_ctx *c, double *rate) {
  WB_CUDA(cudaSetDevice(c->device));
  if (c->nsrc == 0) return 0;
  int rc = 0;
  WbStage st(c);
  double *d_out = st.out(rate, (size_t)c->nsrc, &rc);
  if (rc) return rc;
  int32_t *d_order = nullptr;
  std::vector<int32_t> order(c->h_src_order.begin(), c->h_src_order.end());
  WB_TRY(dev_upload(&d_order, order));
  const WbSources S = wb_sources_args(c);
#define CALL(E) k_source_rates<E><<<wb_grid(c->nsrc, 128), 128, 0, c->stream>>>(S, c->d_state, c->ncell, d_order, d_out)
  DISPATCH_EOS(c, CALL);
#undef CALL
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  rc = st.finish();
  cudaFree(d_order);
  return rc;
}

// error flag of the property kernels, reduced over ranks (mpi_broadcast_error_flag)
static int check_err(wb_ctx *c) {
  WB_TRY(wb_reduce_flags(c, 4));
  return c->h_flags[0] ? 1 : 0;
}

extern "C" int wb_fluid_init(wb_ctx *c, const double *y, const int32_t *region) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(c->ncell > 0, "wb_fluid_init: no mesh");
  int rc = 0;
  {
    WbStage st(c);
    const double *dy = st.in(y, (size_t)c->nowned * c->np, &rc);
    const int32_t *dr = st.in(region, (size_t)c->nowned, &rc);
    if (rc) return rc;
    WB_CUDA(cudaMemcpyAsync(c->d_region, dr, sizeof(int32_t) * c->nowned, cudaMemcpyDeviceToDevice, c->stream));
    WB_TRY(load_y(c, dy));
    WB_TRY(launch_eos(c, 0, 0, 1, 0.0, 0.0));
    c->eval_variant = 0;
    WB_TRY(st.finish());
  }
  return check_err(c);
}

extern "C" int wb_set_boundaries(wb_ctx *c, int n, const int32_t *ghost_cells, const int32_t *interior_cells,
                                 const double *primary, const int32_t *region) {
  WB_CUDA(cudaSetDevice(c->device));
  if (n == 0) return 0;
  WB_CHECK(!wb_is_device_ptr(ghost_cells) && !wb_is_device_ptr(interior_cells),
           "wb_set_boundaries: ghost_cells / interior_cells are read on the host (set-up data): pass host arrays");
  MeshDev &md = meshdev(c);
  const int ncell = c->ncell;
  // rock copied from the interior cell (src/mesh.F90:1189-1193)
  for (int i = 0; i < n; i++) {
    const int g = ghost_cells[i], ic = interior_cells[i];
    WB_CHECK(g >= c->ninterior && g < ncell && ic >= 0 && ic < c->ninterior,
             "wb_set_boundary: ghost %d / interior %d out of range", g, ic);
    memcpy(&c->h_rock[8 * (size_t)g], &c->h_rock[8 * (size_t)ic], 8 * sizeof(double));
  }
  std::vector<double> rockp(5 * (size_t)ncell), perm(3 * (size_t)ncell);
  for (int i = 0; i < ncell; i++) {
    const double *r = &c->h_rock[8 * (size_t)i];
    rockp[i] = r[WB_R_POR];
    rockp[(size_t)ncell + i] = r[WB_R_RHO];
    rockp[2 * (size_t)ncell + i] = r[WB_R_CP];
    rockp[3 * (size_t)ncell + i] = r[WB_R_WET];
    rockp[4 * (size_t)ncell + i] = r[WB_R_DRY];
    for (int d = 0; d < 3; d++) perm[(size_t)d * ncell + i] = r[d];
  }
  WB_CUDA(cudaMemcpyAsync(c->d_rockp, rockp.data(), rockp.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  WB_CUDA(cudaMemcpyAsync(md.d_perm, perm.data(), perm.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  WB_CUDA(cudaStreamSynchronize(c->stream));
  WB_TRY(recompute_face_perm(c));
  int rc = 0;
  {
    WbStage st(c);
    BoundaryArgs a;
    a.cells = st.in(ghost_cells, (size_t)n, &rc);
    a.primary = st.in(primary, (size_t)n * c->np, &rc);
    a.region = st.in(region, (size_t)n, &rc);
    if (rc) return rc;
    a.rockp = c->d_rockp; a.state = c->d_state; a.region_out = c->d_region; a.bprimary = md.d_bprimary;
    a.flags = c->d_flags; a.n = n; a.ncell = ncell; a.ninterior = c->ninterior; a.nslots = c->np + 1;
#define CALL(E) k_boundary<E><<<wb_grid(n, 128), 128, 0, c->stream>>>(c->eos, a)
    DISPATCH_EOS(c, CALL);
#undef CALL
    WB_LAUNCH(c);
    WB_CUDA(cudaGetLastError());
    WB_TRY(st.finish());
  }
  return check_err(c);
}

// Rock records of the interior cells replaced between time steps (time-dependent permeability / porosity,
// flow_simulation_update_rock_properties src/flow_simulation.F90:2051-2089 with the table controls of
// src/rock_control.F90:49-116).  Boundary ghost cells keep the records they copied at set-up (src/mesh.F90:1189-1193),
// as they do in the reference, whose controls list interior cells only (src/rock_setup.F90:404-412).
extern "C" int wb_set_rock(wb_ctx *c, const double *rock) {
  WB_CHECK(c && rock, "wb_set_rock: null argument");
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(!wb_is_device_ptr(rock), "wb_set_rock: rock is read on the host (set-up data): pass a host array");
  WB_CHECK(c->ncell > 0 && c->h_rock.size() == 8 * (size_t)c->ncell, "wb_set_rock: no mesh (wb_set_mesh comes first)");
  MeshDev &md = meshdev(c);
  const int ncell = c->ncell, ni = c->ninterior;
  for (int i = 0; i < ni; i++) {
    const double *r = rock + 8 * (size_t)i;
    WB_CHECK(r[WB_R_POR] >= 0.0 && r[WB_R_POR] <= 1.0 && r[0] >= 0.0 && r[1] >= 0.0 && r[2] >= 0.0,
             "wb_set_rock: cell %d: porosity %g, permeability %g %g %g", i, r[WB_R_POR], r[0], r[1], r[2]);
  }
  memcpy(c->h_rock.data(), rock, 8 * (size_t)ni * sizeof(double));
  std::vector<double> rockp(5 * (size_t)ncell), perm(3 * (size_t)ncell);
  for (int i = 0; i < ncell; i++) {
    const double *r = &c->h_rock[8 * (size_t)i];
    rockp[i] = r[WB_R_POR];
    rockp[(size_t)ncell + i] = r[WB_R_RHO];
    rockp[2 * (size_t)ncell + i] = r[WB_R_CP];
    rockp[3 * (size_t)ncell + i] = r[WB_R_WET];
    rockp[4 * (size_t)ncell + i] = r[WB_R_DRY];
    for (int d = 0; d < 3; d++) perm[(size_t)d * ncell + i] = r[d];
  }
  WB_CUDA(cudaMemcpyAsync(c->d_rockp, rockp.data(), rockp.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  WB_CUDA(cudaMemcpyAsync(md.d_perm, perm.data(), perm.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  WB_CUDA(cudaStreamSynchronize(c->stream));
  WB_TRY(recompute_face_perm(c));
  return 0;
}

extern "C" int wb_set_boundary(wb_ctx *c, int ghost_cell, int interior_cell, const double *primary, int region) {
  int32_t g = ghost_cell, ic = interior_cell, r = region;
  return wb_set_boundaries(c, 1, &g, &ic, primary, &r);
}

extern "C" int wb_get_fluid(wb_ctx *c, double *fluid) {
  WB_CUDA(cudaSetDevice(c->device));
  MeshDev &md = meshdev(c);
  int rc = 0;
  WbStage st(c);
  RecordArgs a;
  a.y = c->d_yloc; a.bprimary = md.d_bprimary; a.region = c->d_region; a.old_region = md.d_old_region;
  a.fluid = st.out(fluid, (size_t)c->ncell * c->dof, &rc);
  if (rc) return rc;
  a.ncell = c->ncell; a.ninterior = c->ninterior; a.dof = c->dof;
#define CALL(E) k_fluid_record<E><<<wb_grid(c->ncell, 128), 128, 0, c->stream>>>(c->eos, a)
  DISPATCH_EOS(c, CALL);
#undef CALL
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  return st.finish();
}

extern "C" int wb_get_regions(wb_ctx *c, int32_t *region) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CUDA(cudaStreamSynchronize(c->stream));
  WB_CUDA(wb_memcpy_sync(region, c->d_region, sizeof(int32_t) * c->ncell,
                     wb_is_device_ptr(region) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
  return 0;
}

// last_iteration_fluid <- fluid (src/flow_simulation.F90:2108-2122): only region and temperature
// of the snapshot are ever read back (by the transitions)
extern "C" int wb_pre_iteration(wb_ctx *c) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CUDA(cudaMemcpyAsync(c->d_region_iter, c->d_region, sizeof(int32_t) * c->ncell, cudaMemcpyDeviceToDevice, c->stream));
  WB_CUDA(cudaMemcpyAsync(c->d_T_iter, c->d_state + (size_t)c->ncell, sizeof(double) * c->ncell,
                          cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}
extern "C" int wb_pre_timestep(wb_ctx *c) {  // :2022-2035
  WB_CUDA(cudaSetDevice(c->device));
  WB_CUDA(cudaMemcpyAsync(c->d_region_step, c->d_region, sizeof(int32_t) * c->ncell, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}
extern "C" int wb_pre_retry_timestep(wb_ctx *c) {  // :2093-2104
  WB_CUDA(cudaSetDevice(c->device));
  WB_CUDA(cudaMemcpyAsync(c->d_region, c->d_region_step, sizeof(int32_t) * c->ncell, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

// ---- function evaluation -----------------------------------------------------------------

// device-pointer core of pre_eval; perturbed evaluations go to the scratch slot 1 so the
// stored (unperturbed) state in slot 0 survives, as `fluid` does in the reference
int wb_pre_eval_dev(wb_ctx *c, const double *d_y, bool unperturbed) {
  WbScopedTimer tm(c, "fluid_props");
  WB_TRY(load_y(c, d_y));
  c->eval_variant = unperturbed ? 0 : 1;
  WB_TRY(launch_eos(c, c->eval_variant, 0, 1, 0.0, 0.0));
  return 0;
}

extern "C" int wb_pre_eval(wb_ctx *c, const double *y, const int32_t *perturbed, int nperturbed) {
  (void)perturbed;  // a masked evaluation equals a full one: properties depend on (y, region) only
  WB_CUDA(cudaSetDevice(c->device));
  int rc = 0;
  {
    WbStage st(c);
    const double *dy = st.in(y, (size_t)c->nowned * c->np, &rc);
    if (rc) return rc;
    WB_TRY(wb_pre_eval_dev(c, dy, nperturbed == 0));
    WB_TRY(st.finish());
  }
  return check_err(c);
}

extern "C" int wb_cell_balances(wb_ctx *c, double *lhs) {
  WB_CUDA(cudaSetDevice(c->device));
  int rc = 0;
  WbStage st(c);
  double *dl = st.out(lhs, (size_t)c->nowned * c->np, &rc);
  if (rc) return rc;
  {
    WbScopedTimer tm(c, "cell_balances");
    WB_TRY(launch_residual(c, c->eval_variant, nullptr, 0.0, dl, nullptr, nullptr));
  }
  return st.finish();
}

extern "C" int wb_cell_inflows(wb_ctx *c, double *rhs) {
  WB_CUDA(cudaSetDevice(c->device));
  int rc = 0;
  WbStage st(c);
  double *dr = st.out(rhs, (size_t)c->nowned * c->np, &rc);
  if (rc) return rc;
  {
    WbScopedTimer tm(c, "cell_inflows");
    WB_TRY(launch_residual(c, c->eval_variant, nullptr, 0.0, nullptr, dr, nullptr));
  }
  return st.finish();
}

// device-pointer core of the BE residual (no flag check, no sync)
int wb_residual_be_dev(wb_ctx *c, const double *d_y, const double *d_lhs_last, double dt, bool unperturbed,
                       double *d_lhs, double *d_rhs, double *d_r) {
  WB_TRY(wb_pre_eval_dev(c, d_y, unperturbed));
  WbScopedTimer tm(c, "cell_inflows");
  WB_TRY(launch_residual(c, c->eval_variant, d_lhs_last, dt, d_lhs, d_rhs, d_r));
  return 0;
}

extern "C" int wb_residual_be(wb_ctx *c, const double *y, const double *lhs_last, double dt,
                              const int32_t *perturbed, int nperturbed, double *lhs, double *rhs, double *r) {
  (void)perturbed;
  WB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)c->nowned * c->np;
  int rc = 0;
  {
    WbStage st(c);
    const double *dy = st.in(y, n, &rc);
    const double *dl = st.in(lhs_last, n, &rc);
    double *o_lhs = st.out(lhs, n, &rc), *o_rhs = st.out(rhs, n, &rc), *o_r = st.out(r, n, &rc);
    if (rc) return rc;
    WB_TRY(wb_residual_be_dev(c, dy, dl, dt, nperturbed == 0, o_lhs, o_rhs, o_r));
    WB_TRY(st.finish());
  }
  return check_err(c);
}

// second stage: reduce the per-block (value, index) pairs
__global__ void k_max_final(const double *__restrict__ pv, const int *__restrict__ pi, int n, double *out_v,
                            long long *out_i, long long offset) {
  double best = -1.0;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (pv[i] > best || (pv[i] == best && pi[i] < bi)) {
      best = pv[i];
      bi = pi[i];
    }
  }
  __shared__ double sv[1024];
  __shared__ int si[1024];
  sv[threadIdx.x] = best;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const double ov = sv[threadIdx.x + s];
      const int oi = si[threadIdx.x + s];
      if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) {
        sv[threadIdx.x] = ov;
        si[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out_v[0] = sv[0];
    out_i[0] = (long long)si[0] + offset;
  }
}

// result left in c->h_red[0] (value) / ((long long*)c->h_red)[1] (global index) after sync
int wb_max_scaled_core(wb_ctx *c, const double *d_v, const double *d_s, double tol, int n, double *maxval,
                       int64_t *maxloc) {
  const int nblk = std::min(wb_grid(n, 256), 4 * WB_NUM_SMS);
  double *pv = c->d_red;
  int *pi = (int *)(c->d_red + nblk);
  double *res = c->d_red + 2 * nblk + 2;  // [value, index] per rank slot
  k_max_scaled<<<nblk, 256, 0, c->stream>>>(d_v, d_s, tol, n, pv, pi);
  WB_LAUNCH(c);
  const int slot = c->nranks > 1 ? c->rank : 0;
  k_max_final<<<1, 1024, 0, c->stream>>>(pv, pi, nblk, res + 2 * slot, (long long *)(res + 2 * slot + 1),
                                         (long long)c->first_cell * c->np);
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  int nslots = 1;
  if (c->nranks > 1) {
    WB_NCCL(wb_nccl()->AllGather(res + 2 * slot, res, 2, ncclDouble, c->comm, c->stream));
    nslots = c->nranks;
  }
  WB_CUDA(cudaMemcpyAsync(c->h_red, res, sizeof(double) * 2 * nslots, cudaMemcpyDeviceToHost, c->stream));
  WB_CUDA(cudaStreamSynchronize(c->stream));
  double best = -1.0;
  int64_t bi = 0;
  for (int r = 0; r < nslots; r++) {
    const double v = c->h_red[2 * r];
    long long idx;
    memcpy(&idx, &c->h_red[2 * r + 1], sizeof(idx));
    if (v > best) {  // ranks ascend with global index: first maximum wins
      best = v;
      bi = idx;
    }
  }
  if (maxval) *maxval = best;
  if (maxloc) *maxloc = bi;
  return 0;
}

extern "C" int wb_max_scaled(wb_ctx *c, const double *v, const double *scale, double tol, double *maxval,
                             int64_t *maxloc) {
  WB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)c->nowned * c->np;
  int rc = 0;
  WbStage st(c);
  const double *dv = st.in(v, n, &rc), *ds = st.in(scale, n, &rc);
  if (rc) return rc;
  return wb_max_scaled_core(c, dv, ds, tol, (int)n, maxval, maxloc);
}

// ---- Jacobian ------------------------------------------------------------------------------

// device core: base state (slot 0) must already hold the properties of y when base_valid
int wb_jacobian_be_dev(wb_ctx *c, const double *d_y, const double *d_lhs_last, double dt, double fd_err,
                       double fd_umin, bool base_valid) {
  WbScopedTimer tm(c, "jacobian");
  WB_CHECK(c->maxdeg <= 8, "wb_jacobian_be: cell with %d faces (max 8 supported by the assembly kernel)", c->maxdeg);
  if (!base_valid) {
    WB_TRY(load_y(c, d_y));
    WB_TRY(launch_eos(c, 0, 0, c->np + 1, fd_err, fd_umin));
    c->eval_variant = 0;
  } else {
    WB_TRY(launch_eos(c, 1, 1, c->np, fd_err, fd_umin));
  }
  JacArgs a;
  a.state = c->d_state; a.Lvar = c->d_Lvar; a.dx = c->d_dx; a.face = c->d_face; a.vol = c->d_vol;
  a.form = wb_res_form(c, d_lhs_last, dt);
  a.cf_ptr = c->d_cf_ptr; a.cf_face = c->d_cf_face; a.cf_other = c->d_cf_other;
  a.cf_bpos = c->d_cf_bpos; a.diagpos = c->d_diagpos; a.val = c->J.d_val; a.dt = dt;
  c->J.version++;
  a.ncell = c->ncell; a.ninterior = c->ninterior; a.nowned = c->nowned; a.nface = c->nface;
  a.src = wb_sources_args(c);
  static int lanes = -1;  // WB_JAC_LANES = 1: eight lanes per row (measured slower: 2.04 vs 1.73 ms per assembly at 1 M cells); default: thread per row
  if (lanes < 0) {
    const char *e = getenv("WB_JAC_LANES");
    lanes = e ? atoi(e) : 0;
  }
  const int grid = wb_grid(c->nowned, 128);
  if (lanes) {
    const int gl = wb_grid((size_t)c->nowned * 8, 256);
#define CALL(E) k_jacobian_lanes<E><<<gl, 256, 0, c->stream>>>(a)
    DISPATCH_EOS(c, CALL);
#undef CALL
  } else if (c->maxdeg <= 6) {
#define CALL(E) k_jacobian<E, 6><<<grid, 128, 0, c->stream>>>(a)
    DISPATCH_EOS(c, CALL);
#undef CALL
  } else {
#define CALL(E) k_jacobian<E, 8><<<grid, 128, 0, c->stream>>>(a)
    DISPATCH_EOS(c, CALL);
#undef CALL
  }
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int wb_jacobian_be(wb_ctx *c, const double *y, const double *lhs_last, double dt, double fd_err,
                              double fd_umin, double *vals_out) {
  WB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)c->nowned * c->np;
  int rc = 0;
  {
    WbStage st(c);
    const double *dy = st.in(y, n, &rc), *dl = st.in(lhs_last, n, &rc);
    if (rc) return rc;
    WB_TRY(wb_jacobian_be_dev(c, dy, dl, dt, fd_err, fd_umin, false));
    if (vals_out) {
      const size_t nv = (size_t)c->J.nnzb * c->np * c->np;
      WB_CUDA(cudaMemcpyAsync(vals_out, c->J.d_val, nv * sizeof(double),
                              wb_is_device_ptr(vals_out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                              c->stream));
    }
    WB_TRY(st.finish());
  }
  return check_err(c);
}

// ---- the reference's colouring loop (MatFDColoringApply), for parity checks ---------------

__global__ void k_perturb_color(const double *__restrict__ y, const int32_t *__restrict__ color, int k, int var,
                                int nb, int np, double err, double umin, double *__restrict__ w,
                                double *__restrict__ vscale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nb) return;
  for (int q = 0; q < np; q++) {
    double v = y[(size_t)c * np + q];
    if (q == var && color[c] == k) {
      const double dx = fd_step(v, err, umin);
      vscale[c] = 1.0 / dx;
      v += dx;
    }
    w[(size_t)c * np + q] = v;
  }
}

__global__ void k_fd_scatter(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                             const int32_t *__restrict__ color, int k, int var, int nb, int np,
                             const double *__restrict__ Fp, const double *__restrict__ F0,
                             const double *__restrict__ vscale, double *__restrict__ val) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nb) return;
  for (int e = rowptr[r]; e < rowptr[r + 1]; e++) {
    const int col = colidx[e];
    if (col < nb && color[col] == k) {
      for (int ii = 0; ii < np; ii++) {
        const double w2 = Fp[(size_t)r * np + ii] + (-1.0) * F0[(size_t)r * np + ii];
        val[(size_t)e * np * np + var * np + ii] = w2 * vscale[col];
      }
    }
  }
}

// greedy distance-2 colouring of the block columns in natural order
static void color_pattern(wb_ctx *c) {
  const wb_mat &J = c->J;
  const int nb = J.nb;
  c->h_color.assign(nb, -1);
  std::vector<int32_t> mark;
  int ncolor = 0;
  for (int col = 0; col < nb; col++) {
    for (int k = J.h_rowptr[col]; k < J.h_rowptr[col + 1]; k++) {
      const int r = J.h_colidx[k];
      if (r >= nb) continue;
      for (int k2 = J.h_rowptr[r]; k2 < J.h_rowptr[r + 1]; k2++) {
        const int c2 = J.h_colidx[k2];
        if (c2 < nb && c->h_color[c2] >= 0) {
          if ((int)mark.size() <= c->h_color[c2]) mark.resize(c->h_color[c2] + 1, -1);
          mark[c->h_color[c2]] = col;
        }
      }
    }
    int q = 0;
    while (q < ncolor && q < (int)mark.size() && mark[q] == col) q++;
    if (q == ncolor) ncolor++;
    c->h_color[col] = q;
  }
  c->ncolor = ncolor;
}

extern "C" int wb_jacobian_be_colored(wb_ctx *c, const double *y, const double *lhs_last, double dt,
                                      double fd_err, double fd_umin, double *vals_out, int *ncolors) {
  WB_CUDA(cudaSetDevice(c->device));
  WB_CHECK(c->nranks == 1, "wb_jacobian_be_colored: single-GPU parity path only");
  MeshDev &md = meshdev(c);
  const int nb = c->nowned, np = c->np;
  const size_t n = (size_t)nb * np;
  if (c->h_color.empty()) color_pattern(c);
  if (ncolors) *ncolors = c->ncolor;
  if (!md.d_w) {
    WB_TRY(dev_alloc(&md.d_w, 3 * n + nb));
    WB_TRY(dev_alloc(&md.d_F, n));
  }
  double *d_w = md.d_w, *d_Fp = md.d_w + n, *d_vscale = md.d_w + 2 * n, *d_F0 = md.d_F;
  int32_t *d_color = nullptr;
  WB_TRY(dev_upload(&d_color, c->h_color));
  int rc = 0, err = 0;
  {
    WbStage st(c);
    const double *dy = st.in(y, n, &rc), *dl = st.in(lhs_last, n, &rc);
    if (rc) return rc;
    WB_TRY(wb_residual_be_dev(c, dy, dl, dt, true, nullptr, nullptr, d_F0));
    WB_CUDA(cudaMemsetAsync(c->J.d_val, 0, sizeof(double) * (size_t)c->J.nnzb * np * np, c->stream));
    c->J.version++;
    for (int k = 0; k < c->ncolor; k++)
      for (int var = 0; var < np; var++) {
        k_perturb_color<<<wb_grid(nb, 256), 256, 0, c->stream>>>(dy, d_color, k, var, nb, np, fd_err, fd_umin,
                                                               d_w, d_vscale);
        WB_LAUNCH(c);
        WB_TRY(wb_residual_be_dev(c, d_w, dl, dt, false, nullptr, nullptr, d_Fp));
        k_fd_scatter<<<wb_grid(nb, 256), 256, 0, c->stream>>>(c->J.d_rowptr, c->J.d_colidx, d_color, k, var, nb,
                                                            np, d_Fp, d_F0, d_vscale, c->J.d_val);
        WB_LAUNCH(c);
      }
    WB_CUDA(cudaGetLastError());
    // leave the stored state as the unperturbed one
    WB_TRY(wb_pre_eval_dev(c, dy, true));
    if (vals_out) {
      const size_t nv = (size_t)c->J.nnzb * np * np;
      WB_CUDA(cudaMemcpyAsync(vals_out, c->J.d_val, nv * sizeof(double),
                              wb_is_device_ptr(vals_out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                              c->stream));
    }
    WB_TRY(st.finish());
    err = check_err(c);
  }
  cudaFree(d_color);
  return err;
}

// ---- transitions -----------------------------------------------------------------------------

int wb_fluid_transitions_dev(wb_ctx *c, const double *d_y_old, double *d_search, double *d_y) {
  WbScopedTimer tm(c, "fluid_trans");
  MeshDev &md = meshdev(c);
  TransArgs a;
  a.y_old = d_y_old; a.y = d_y; a.search = d_search; a.region = c->d_region; a.old_region = md.d_old_region;
  a.region_iter = c->d_region_iter; a.T_iter = c->d_T_iter; a.flags = c->d_flags; a.nowned = c->nowned;
#define CALL(E) k_transitions<E><<<wb_grid(c->nowned, 128), 128, 0, c->stream>>>(c->eos, a)
  DISPATCH_EOS(c, CALL);
#undef CALL
  WB_LAUNCH(c);
  WB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int wb_fluid_transitions(wb_ctx *c, const double *y_old, double *search, double *y,
                                    int *changed_search, int *changed_y) {
  WB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)c->nowned * c->np;
  int rc = 0;
  {
    WbStage st(c);
    const double *dyo = st.in(y_old, n, &rc);
    double *ds = st.out(search, n, &rc, true), *dy = st.out(y, n, &rc, true);
    if (rc) return rc;
    WB_TRY(wb_fluid_transitions_dev(c, dyo, ds, dy));
    WB_TRY(st.finish());
  }
  WB_TRY(wb_reduce_flags(c, 4));
  if (changed_search) *changed_search = c->h_flags[1];
  if (changed_y) *changed_y = c->h_flags[2];
  return c->h_flags[0] ? 1 : 0;
}
