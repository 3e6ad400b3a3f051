// wb_tracer.cuh -- row assembly of the passive-tracer auxiliary linear system (see wb_tracer.cu).
// Host-compilable: tests/hostcheck runs the same source on the CPU against the oracle.
#pragma once
#include <math.h>

#include "../../include/waiwera_b200.h"
#include "wb_state.cuh"

#define WB_GAS_CONSTANT 8.3144598  // src/thermodynamics.F90: gas_constant

struct WbTracerDev {
  int phase[WB_MAX_TRACERS];  // 0-based here
  double diffusion[WB_MAX_TRACERS], decay[WB_MAX_TRACERS], activation[WB_MAX_TRACERS];
};

struct TracerArgs {
  const double *state;  // SoA state of the last unperturbed evaluation
  const double *face, *vol, *rockp;
  const int32_t *cf_ptr, *cf_face, *cf_other, *cf_bpos, *diagpos, *rowptr;
  WbSources src;
  const double *inj;  // [nsrc*nt] or null
  WbTracerDev trc;
  int method;               // WB_METHOD_*
  double sA, sD, s0, s2, sb;  // scales of Ar, Al (diagonal), Al_last x_last, Al_last2 x_last2, br
  const double *al_last, *x_last, *al_last2, *x_last2;
  const double *xb;  // [(ncell-ninterior)*nt] Dirichlet mass fractions or null
  double *val, *b, *al;  // val / b may be null (balances only)
  int ncell, ninterior, nowned, nface;
};

// v[p] for a run-time phase index without spilling the array to local memory
template <int N> WB_HD double pick(const double (&v)[N], int p) {
  double r = v[0];
#pragma unroll
  for (int k = 1; k < N; k++)
    if (p == k) r = v[k];
  return r;
}

// cell%tracer_balance_coefs (src/cell.F90:146-164): porosity * saturation * density
template <int NC, int NPH>
WB_HD double balance_coef(double por, const WbCellState<NC, NPH> &s, int p) {
  return por * pick(s.sat, p) * pick(s.rho, p);
}
// cell%diffusion_factor (src/cell.F90:168-201): porosity * density * tortuosity, tortuosity = 1 * saturation
template <int NC, int NPH>
WB_HD double diffusion_factor(double por, const WbCellState<NC, NPH> &s, int p) {
  const double rock_tortuosity = 1.0;
  const double tortuosity = rock_tortuosity * pick(s.sat, p);
  return por * pick(s.rho, p) * tortuosity;
}

// Thread per owned cell = block row of A_aux.  The reference's face loop adds to the row of each support cell in
// face order (advective entry at the upstream column, then the two diffusive entries), then the sources in source
// order, then the decay term; re-running that sequence per row reproduces its sums without atomics.
template <int EOS, int NT>
WB_HD void wb_tracer_row(const TracerArgs &a, int i) {
  constexpr int NP = WbEosTraits<EOS>::NP, NC = WbEosTraits<EOS>::NC, NPH = WbEosTraits<EOS>::NPH;
  const size_t nc = a.ncell;
  WbCellState<NC, NPH> si, so;
  load_state(a.state, nc, i, si);
  const double por = a.rockp[i], vol = a.vol[i];
  double coef[NT];
#pragma unroll
  for (int t = 0; t < NT; t++) {
    coef[t] = balance_coef(por, si, a.trc.phase[t]);
    if (a.al) a.al[(size_t)i * NT + t] = coef[t];
  }
  if (!a.val) return;

  const int k0 = a.rowptr[i], k1 = a.rowptr[i + 1];
  for (int k = k0; k < k1; k++) {
    double *blk = a.val + (size_t)k * NT * NT;
#pragma unroll
    for (int q = 0; q < NT * NT; q++) blk[q] = 0.0;
  }
  double diag[NT], br[NT], bdy[NT];
#pragma unroll
  for (int t = 0; t < NT; t++) diag[t] = br[t] = bdy[t] = 0.0;

  // faces (src/flow_simulation.F90:1632-1684)
  const int e0 = a.cf_ptr[i], e1 = a.cf_ptr[i + 1];
  for (int e = e0; e < e1; e++) {
    const int fs = a.cf_face[e], side = fs & 1, o = a.cf_other[e];
    const WbFaceGeom g = load_face(a.face, a.nface, fs >> 1);
    load_state(a.state, nc, o, so);
    const double por_o = a.rockp[o];
    double flux[NP], pf[NPH];
    if (side == 0) wb_face_flux<NP, NC, NPH>(g, si, so, flux, pf);
    else wb_face_flux<NP, NC, NPH>(g, so, si, flux, pf);
    const double sign_i = side == 0 ? -1.0 : 1.0;  // flux_sign of this cell
    const int bpos = a.cf_bpos[e];
#pragma unroll
    for (int t = 0; t < NT; t++) {
      const int p = a.trc.phase[t];
      const double tracer_phase_flux = pick(pf, p);
      const int up = tracer_phase_flux >= 0.0 ? 0 : 1;  // upstream support cell (0: cell 1)
      const double tracer_flow = tracer_phase_flux * g.area;
      const double fi = diffusion_factor(por, si, p), fo = diffusion_factor(por_o, so, p);
      const double dfac = side == 0 ? wb_harmonic(g.d1, g.d2, g.d12, fi, fo) : wb_harmonic(g.d1, g.d2, g.d12, fo, fi);
      double off = 0.0;
      // advective entry at (row i, upstream column)
      const double Fa = sign_i * tracer_flow / vol;
      if (up == side) diag[t] = diag[t] + Fa;
      else off = off + Fa;
      // diffusive entries, support cell 1 then 2: -sign_i sign_j A D_f D / (d12 V_i)
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const double sign_j = j == 0 ? -1.0 : 1.0;
        const double Fd = -sign_i * sign_j * g.area * dfac * a.trc.diffusion[t] / (g.d12 * vol);
        if (j == side) diag[t] = diag[t] + Fd;
        else off = off + Fd;
      }
      if (bpos >= 0) {
        a.val[(size_t)bpos * NT * NT + t * NT + t] += off;
      } else if (a.xb) {
        // Dirichlet ghost column: its identity row makes x = x_boundary, moved to the right-hand side
        bdy[t] = bdy[t] + (a.sA * off) * a.xb[(size_t)(o - a.ninterior) * NT + t];
      }
    }
  }

  // sources in source order (tracer_source_iterator :1717-1771)
  if (a.src.head) {
    int k = a.src.head[i];
    if (k >= 0) {
      for (; k < a.src.n && a.src.cell[k] == i; k++) {
        const double rate = wb_source_rate(a.src, k, si);
        if (wb_source_component(a.src.comp[k], rate) < NP) {
          if (rate < 0.0) {
            // fluid%phase_flow_fractions (src/fluid.F90:394-411)
            double frac[NPH], sum = 0.0;
#pragma unroll
            for (int p = 0; p < NPH; p++) {
              frac[p] = (si.phases & (1 << p)) ? si.mob[p] : 0.0;
              sum += frac[p];
            }
#pragma unroll
            for (int t = 0; t < NT; t++) {
              const double q = (pick(frac, a.trc.phase[t]) / sum) * rate / vol;
              diag[t] = diag[t] + q;
            }
          } else if (a.inj) {
#pragma unroll
            for (int t = 0; t < NT; t++) br[t] = br[t] + a.inj[(size_t)k * NT + t] / vol;
          }
        }
      }
    }
  }

  // decay (apply_tracer_decay :1775-1831; tracer_decay src/tracer.F90:48-61)
#pragma unroll
  for (int t = 0; t < NT; t++) {
    const double Tk = si.T + WB_TC_K;
    const double rate = a.trc.decay[t] * exp(-a.trc.activation[t] / (WB_GAS_CONSTANT * Tk));
    diag[t] = diag[t] + (-rate * coef[t]);
  }

  // setup_linear (src/timestepper.F90:458-581): MatScale, MatDiagonalSet(ADD), right-hand side
  double *dblk = a.val + (size_t)a.diagpos[i] * NT * NT;
  for (int k = k0; k < k1; k++) {
    if (k == a.diagpos[i]) continue;
    double *blk = a.val + (size_t)k * NT * NT;
#pragma unroll
    for (int t = 0; t < NT; t++) blk[t * NT + t] = a.sA * blk[t * NT + t];
  }
#pragma unroll
  for (int t = 0; t < NT; t++) {
    const size_t idx = (size_t)i * NT + t;
    double d = a.sA * diag[t];
    double rhs;
    if (a.method == WB_METHOD_DIRECTSS) {
      rhs = -1.0 * br[t];
    } else {
      d = d + coef[t] * a.sD;
      rhs = a.al_last[idx] * a.x_last[idx];
      if (a.method == WB_METHOD_BDF2) {
        rhs = rhs * a.s0;
        rhs = rhs + a.s2 * (a.al_last2[idx] * a.x_last2[idx]);
      }
      rhs = rhs + a.sb * br[t];
    }
    rhs = rhs - bdy[t];
    // aux_pre_solve (:1878-1899): phase of the tracer absent => row = identity, b = 0
    if (!(si.phases & (1 << a.trc.phase[t]))) {
      for (int k = k0; k < k1; k++) a.val[(size_t)k * NT * NT + t * NT + t] = 0.0;
      d = 1.0;
      rhs = 0.0;
    }
    dblk[t * NT + t] = d;
    if (a.b) a.b[idx] = rhs;
  }
}

#if defined(__CUDACC__)
template <int EOS, int NT>
__global__ void __launch_bounds__(128) k_tracer_assemble(const TracerArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.nowned) wb_tracer_row<EOS, NT>(a, i);
}
#endif
