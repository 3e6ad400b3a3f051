"""The tracer row assembly the CUDA kernel k_tracer_assemble runs (waiwera_b200/csrc/wb_tracer.cuh, compiled for
the host by tests/hostcheck) against the oracle's restatement of aux_lhs / aux_rhs / setup_linear / aux_pre_solve
(oracle/wo_tracer.c; src/flow_simulation.F90:1489-1959, src/timestepper.F90:458-581) on a 3-D two-phase mesh with
Dirichlet boundary cells, production and injection sources, diffusion and Arrhenius decay.  CPU-side check of
the source only: the product has no CPU path; the GPU parity tests are tests/test_gpu_tracer.py."""
import ctypes as C

import numpy as np
import pytest

from test_device_headers_host import hc  # noqa: F401  (fixture)
from util import SEED, boundary_values, make_problem, make_problem_wce

TRACERS = dict(phases=[1, 2, 1], diffusion=[1.0e-6, 2.0e-5, 0.0], decay=[0.0, 1.0e-7, 1.0e-6],
               activation=[0.0, 0.0, 2000.0])


def tracer_case(wo, eos="we", nt=3, seed=SEED):
    """oracle flow object at an unperturbed evaluation + everything the tracer system needs"""
    rng = np.random.default_rng(seed)
    if eos == "we":
        m, y, region, prm = make_problem(wo, dims=(6, 5, 4), two_phase_layers=2, top_boundary=True)
    else:
        m, y, region, prm = make_problem_wce(wo, dims=(6, 5, 4), two_phase_layers=2)
    f = wo.Flow(prm, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    nb = m.ncell - m.ninterior
    bprim, breg = (boundary_values(m) if nb else (np.zeros((0, f.np)), np.zeros(0, np.int32)))
    for k in range(nb):
        assert f.set_boundary(int(m.boundary["ghost_cells"][k]), int(m.boundary["interior_cells"][k]), bprim[k], int(breg[k])) == 0
    # production from a two-phase and a liquid cell, injection (with tracer) into two cells, a heat source (ignored)
    n = m.nowned
    cells = np.array([3, n - 2, 7, 7, n // 2, 11], np.int32)
    comps = np.array([1, 0, 1, 1, 1, f.np], np.int32)
    rates = np.array([-0.5, -0.2, 0.3, 0.1, 0.25, 1.0e3])
    enth = np.array([0.0, 0.0, 4.0e5, 8.0e5, 1.0e5, 0.0])
    f.set_sources(cells, comps, rates, enth)
    t = {k: v[:nt] for k, v in TRACERS.items()}
    f.set_tracers(t["phases"], t["diffusion"], t["decay"], t["activation"])
    inj = rng.uniform(0.0, 1.0e-4, (len(cells), nt))
    f.set_tracer_injection(inj)
    # source controls: the first producer on deliverability (production only), the second with a total limiter
    ctrl = dict(sources=np.array([0, 1], np.int32), pi=np.array([2.0e-12, 0.0]), pref=np.array([2.0e5, 0.0]),
                direction=np.array([1, 0], np.int32), limit=np.array([0.0, 0.15]))
    f.set_source_controls(ctrl["sources"], ctrl["pi"], ctrl["pref"], ctrl["direction"], ctrl["limit"])
    assert f.fluid_init(y, region) == 0
    err, L0 = f.lhs(y)
    assert err == 0
    err, _, _, _ = f.residual(y, L0, 1.0e5)          # unperturbed evaluation: fluid + phase fluxes at y
    assert err == 0
    nrows = f.ntrows
    x_last = rng.uniform(0.0, 0.01, nrows * nt)
    x_last2 = rng.uniform(0.0, 0.01, nrows * nt)
    al_last = f.tracer_balances() * rng.uniform(0.95, 1.05, nrows * nt)
    al_last2 = f.tracer_balances() * rng.uniform(0.95, 1.05, nrows * nt)
    # unscaled primaries exactly as the oracle's fluid_properties sees them: eos%unscale of the scaled y
    primary = np.zeros((m.nowned, f.np))
    for c in range(m.nowned):
        wo.lib().wo_eos_unscale(f.eos, wo.dp(y[c * f.np:(c + 1) * f.np].copy()), int(region[c]), wo.dp(primary[c]))
    prim_all = np.vstack([primary[:m.nowned], bprim]) if nb else primary[:m.nowned]
    reg_all = np.concatenate([region[:m.nowned], breg]).astype(np.int32)
    src = dict(cells=cells, comps=comps, rates=rates, inj=inj, ctrl=ctrl)
    tracer_case.last = dict(y=y, region=region)   # scaled state, for the GPU tests that rebuild the same problem
    return m, f, prm, t, src, x_last, x_last2, al_last, al_last2, np.ascontiguousarray(prim_all), reg_all


def host_assemble(wo, hc, m, f, prm, t, src, method, dt, dt_last, al_last, x_last, al_last2, x_last2, prim_all, reg_all):
    nt, n = f.nt, m.nowned
    J = f.bsr()
    rowptr, colidx, _ = wo.bsr_arrays(J)
    rowptr, colidx = rowptr.copy(), colidx.copy()
    wo.lib().wo_bsr_destroy(J)
    val = np.zeros(len(colidx) * nt * nt)
    b, al = np.zeros(n * nt), np.zeros(n * nt)
    nb = m.ncell - m.ninterior
    xb = np.ascontiguousarray(x_last.reshape(-1, nt)[n:].reshape(-1)) if nb else None
    arr = lambda a: np.ascontiguousarray(a, np.float64)
    a0, x0 = arr(al_last.reshape(-1, nt)[:n].reshape(-1)), arr(x_last.reshape(-1, nt)[:n].reshape(-1))
    a2, x2 = arr(al_last2.reshape(-1, nt)[:n].reshape(-1)), arr(x_last2.reshape(-1, nt)[:n].reshape(-1))
    ns, c = len(src["cells"]), src["ctrl"]
    cw, cpi, cpref, clim = np.zeros(ns, np.int32), np.zeros(ns), np.zeros(ns), np.zeros(ns)
    for k, sidx in enumerate(c["sources"]):      # control word as wb_set_source_controls packs it
        cw[sidx] = (1 if c["pi"][k] > 0 else 0) | (int(c["direction"][k]) << 1)
        cpi[sidx], cpref[sidx], clim[sidx] = c["pi"][k], c["pref"][k], c["limit"][k]
    rc = hc.hc_tracer_assemble(
        C.byref(prm), nt, m.ncell, n, m.nface, wo.ip(np.ascontiguousarray(m.face_cells.reshape(-1))),
        wo.dp(arr(m.face_geom.reshape(-1))), wo.dp(arr(m.cell_geom.reshape(-1))), wo.dp(arr(f.L and np.ctypeslib.as_array(
            C.cast(f.mesh.rock, C.POINTER(C.c_double)), shape=(m.ncell * 8,)).copy())),
        wo.dp(prim_all.reshape(-1)), wo.ip(reg_all), wo.ip(np.array(t["phases"], np.int32)), wo.dp(arr(t["diffusion"])),
        wo.dp(arr(t["decay"])), wo.dp(arr(t["activation"])), len(src["cells"]), wo.ip(src["cells"]), wo.ip(src["comps"]),
        wo.dp(src["rates"]), wo.ip(cw), wo.dp(cpi), wo.dp(cpref), wo.dp(clim), wo.dp(arr(src["inj"].reshape(-1))), method, dt, dt_last, wo.dp(a0), wo.dp(x0), wo.dp(a2),
        wo.dp(x2), wo.dp(xb), wo.ip(rowptr), wo.ip(colidx), wo.dp(val), wo.dp(b), wo.dp(al))
    assert rc == 0
    return rowptr, colidx, val.reshape(-1, nt * nt), b, al


def eliminate_boundary(wo, A, b, x_last, nowned, nt):
    """owned part of the oracle's extended system: blocks with owned columns in pattern order, and
    b_i - sum_j A_ij x_j over the Dirichlet columns j (identity rows: x_j = x_last_j)"""
    rowptr, colidx, val = wo.bsr_arrays(A)
    vals, be = [], b[:nowned * nt].copy()
    for i in range(nowned):
        for k in range(rowptr[i], rowptr[i + 1]):
            j = colidx[k]
            if j < nowned:
                vals.append(val[k].copy())
            else:
                for it in range(nt):
                    be[i * nt + it] = be[i * nt + it] - val[k][it * nt + it] * x_last[j * nt + it]
    return np.array(vals), be


@pytest.mark.parametrize("method", [0, 1, 2])
@pytest.mark.parametrize("nt", [1, 2, 3])
def test_tracer_rows_match_oracle(wo, hc, method, nt):
    m, f, prm, t, src, x_last, x_last2, al_last, al_last2, prim_all, reg_all = tracer_case(wo, "we", nt)
    dt, dt_last = 8.64e5, 5.0e5
    A = f.tracer_pattern()
    b_ref, al_ref = f.tracer_setup_linear(A, dt, al_last, x_last, method=method, dt_last=dt_last, al_last2=al_last2,
                                          x_last2=x_last2)
    vals_ref, b_el = eliminate_boundary(wo, A, b_ref, x_last, m.nowned, nt)
    rowptr, colidx, val, b, al = host_assemble(wo, hc, m, f, prm, t, src, method, dt, dt_last, al_last, x_last,
                                               al_last2, x_last2, prim_all, reg_all)
    assert val.shape == vals_ref.shape
    # entries: same operation order on both sides, FMA contraction off => identical unless the Arrhenius exp differs
    scale = np.abs(vals_ref).max()
    assert np.abs(val - vals_ref).max() <= 1e-15 * scale
    exact = (val == vals_ref).mean()
    assert exact > 0.99, exact
    assert np.abs(b - b_el).max() <= 1e-15 * np.abs(b_el).max()
    if method != 2:
        assert (al == al_ref[:m.nowned * nt]).all()
    # the case exercises what it claims to: absent phases, upstream on both sides, boundary elimination
    d = val[[np.searchsorted(colidx[rowptr[i]:rowptr[i + 1]], i) + rowptr[i] for i in range(m.nowned)]]
    if nt >= 2:
        assert (d[:, nt + 1] == 1.0).any() and (d[:, nt + 1] != 1.0).any()      # vapour tracer: absent in liquid cells
    assert (b_el != b_ref[:m.nowned * nt]).any()
    wo.lib().wo_bsr_destroy(A)


def test_tracer_rows_match_oracle_wce(wo, hc):
    """three-primary EOS (water + CO2 + energy): phase fluxes sum over two mass components"""
    m, f, prm, t, src, x_last, x_last2, al_last, al_last2, prim_all, reg_all = tracer_case(wo, "wce", 2)
    A = f.tracer_pattern()
    b_ref, al_ref = f.tracer_setup_linear(A, 8.64e5, al_last, x_last)
    vals_ref, b_el = eliminate_boundary(wo, A, b_ref, x_last, m.nowned, 2)
    rowptr, colidx, val, b, al = host_assemble(wo, hc, m, f, prm, t, src, 0, 8.64e5, 0.0, al_last, x_last, al_last2,
                                               x_last2, prim_all, reg_all)
    assert np.abs(val - vals_ref).max() <= 1e-15 * np.abs(vals_ref).max()
    assert np.abs(b - b_el).max() <= 1e-15 * np.abs(b_el).max()
    assert (al == al_ref[:m.nowned * 2]).all()
    wo.lib().wo_bsr_destroy(A)
