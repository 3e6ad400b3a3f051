"""GPU parity of the passive-tracer auxiliary linear problem through the C ABI:
wb_set_tracers / wb_set_tracer_injection / wb_tracer_cell_balances / wb_tracer_setup_linear / wb_tracer_solve
against the oracle (oracle/wo_tracer.c) and, end to end, against the AUTOUGH2 listings of the reference's
1-D tracer benchmark (tests/test_tracer_oned.py)."""
import ctypes as C

import numpy as np
import pytest

from test_tracer_host import eliminate_boundary, tracer_case
from util import gpu_flow

pytestmark = pytest.mark.gpu


def gpu_case(wo, flow, eos, nt):
    m, f, prm, t, src, x_last, x_last2, al_last, al_last2, prim_all, reg_all = tracer_case(wo, eos, nt)
    y, region = tracer_case.last["y"], tracer_case.last["region"]
    sim = gpu_flow(wo, flow, m, prm, y, region)
    assert sim.set_sources(src["cells"], src["comps"], src["rates"], np.zeros(len(src["cells"]))) == 0
    c = src["ctrl"]
    assert sim.set_source_controls(c["sources"], c["pi"], c["pref"], c["direction"], c["limit"]) == 0
    assert sim.set_tracers(t["phases"], t["diffusion"], t["decay"], t["activation"]) == 0
    assert sim.set_tracer_injection(src["inj"]) == 0
    err, L0 = sim.lhs(y)
    assert err == 0
    n = m.nowned
    own = lambda v: np.ascontiguousarray(v.reshape(-1, nt)[:n].reshape(-1))
    xb = np.ascontiguousarray(x_last.reshape(-1, nt)[n:].reshape(-1)) if m.ncell > m.ninterior else None
    return m, f, sim, own, xb, (x_last, x_last2, al_last, al_last2)


@pytest.mark.parametrize("eos,nt,method", [("we", 1, 0), ("we", 3, 0), ("we", 2, 1), ("we", 3, 2), ("wce", 2, 0)])
def test_tracer_system_matches_oracle(wo, eos, nt, method):
    from waiwera_b200 import flow
    m, f, sim, own, xb, (x_last, x_last2, al_last, al_last2) = gpu_case(wo, flow, eos, nt)
    dt, dt_last = 8.64e5, 5.0e5
    A = f.tracer_pattern()
    b_ref, al_ref = f.tracer_setup_linear(A, dt, al_last, x_last, method=method, dt_last=dt_last, al_last2=al_last2,
                                          x_last2=x_last2)
    vals_ref, b_el = eliminate_boundary(wo, A, b_ref, x_last, m.nowned, nt)
    if method == 1:
        assert sim.set_method(flow.METHOD_BDF2, dt_last, np.zeros(m.nowned * sim.np)) == 0
    elif method == 2:
        assert sim.set_method(flow.METHOD_DIRECTSS) == 0
    assert np.array_equal(sim.tracer_balances(), f.tracer_balances()[:m.nowned * nt]) or \
        np.abs(sim.tracer_balances() - f.tracer_balances()[:m.nowned * nt]).max() < 1e-13 * np.abs(al_ref).max()
    Ad, b, al = sim.tracer_setup_linear(dt, own(al_last), own(x_last), xb, own(al_last2), own(x_last2))
    val = Ad.get_values(len(vals_ref)).reshape(-1, nt * nt)
    assert np.abs(val - vals_ref).max() <= 1e-13 * np.abs(vals_ref).max()
    assert np.abs(b - b_el).max() <= 1e-13 * np.abs(b_el).max()
    if method != 2:
        assert np.abs(al - al_ref[:m.nowned * nt]).max() <= 1e-13 * np.abs(al_ref).max()
    # the Krylov solve of the assembled system against the oracle's solve of its (extended) system
    k = wo.KspOpts()
    k.type, k.restart, k.maxit, k.rtol, k.atol, k.dtol = wo.KSP_BCGS, 30, 10000, 1e-12, 1e-50, 1e5
    L = wo.lib()
    pc = L.wo_pc_create(A, wo.PC_BJACOBI_ILU0, None)
    x_ref = np.zeros(len(b_ref))
    its, rn = C.c_int(), C.c_double()
    assert L.wo_ksp_solve(A, pc, C.byref(k), wo.dp(b_ref), wo.dp(x_ref), C.byref(its), C.byref(rn)) > 0
    L.wo_pc_destroy(pc)
    x, al2, reason, kits = sim.tracer_solve(dt, own(al_last), own(x_last), xb, own(al_last2), own(x_last2),
                                            opts=flow.ksp_opts(type=flow.KSP_BCGS, rtol=1e-12))
    assert reason > 0
    assert np.abs(x - x_ref[:m.nowned * nt]).max() <= 1e-9 * np.abs(x_ref).max()
    # GMRES through the same entry point
    x2, _, reason2, _ = sim.tracer_solve(dt, own(al_last), own(x_last), xb, own(al_last2), own(x_last2),
                                         opts=flow.ksp_opts(type=flow.KSP_GMRES, rtol=1e-12))
    assert reason2 > 0 and np.abs(x2 - x_ref[:m.nowned * nt]).max() <= 1e-9 * np.abs(x_ref).max()
    L.wo_bsr_destroy(A)
    sim.destroy()


@pytest.mark.parametrize("case", ["single", "two"])
def test_oned_tracer_benchmark_on_gpu(wo, case):
    """the reference's 1-D tracer benchmark end to end on the CUDA path: flow Newton step + tracer solve per time
    step, against the oracle's run (tight) and the AUTOUGH2 listing (the reference's own tolerance)"""
    from waiwera_b200 import flow
    import test_tracer_oned as T
    from util import wb_params_from_oracle
    m_ref, y_ss, hist_ref, prod_ref = T.run_oracle(wo, case)
    c = T.CASES[case]
    m, y, region = T.problem(case)
    sim = flow.FlowSimulation(wb_params_from_oracle(wo, flow, T.params(wo)), m)
    assert sim.set_boundaries(m.boundary["ghost_cells"], m.boundary["interior_cells"], np.array([c["primary"]], float),
                              np.array([c["region"]], np.int32)) == 0
    assert sim.set_sources([T.NX - 1], [1], [c["rate"]], [0.0]) == 0
    assert sim.set_tracers([1]) == 0
    assert sim.fluid_init(y, region) == 0
    o = flow.newton_opts(max_iterations=8, rel_tol=1e-9, pc_type=flow.PC_BJACOBI_ILU0,
                         ksp=flow.ksp_opts(type=flow.KSP_BCGS))
    err, _ = sim.lhs(y)
    assert err == 0
    al = sim.tracer_balances()
    x = np.zeros(T.NX)
    xb = np.array([T.X_BOUNDARY])
    for step in range(c["nsteps"]):
        err, L0 = sim.lhs(y)
        assert err == 0
        sim.pre_timestep()
        res = sim.newton_solve(y, L0, T.DT, o)
        assert res.reason > 0
        err, _, _, _ = sim.residual(y, L0, T.DT)      # unperturbed evaluation at the converged state
        assert err == 0
        x, al, reason, its = sim.tracer_solve(T.DT, al, x, xb, opts=flow.ksp_opts(type=flow.KSP_BCGS, rtol=1e-10))
        assert reason > 0
        y_ref, x_ref = hist_ref[step]
        assert np.abs(y - y_ref).max() / np.abs(y_ref).max() < 1e-7
        assert np.abs(x - x_ref[:T.NX]).max() < 1e-9
        rows = np.array(T.GOLD[case]["tables"][step])[:T.NX]
        assert np.abs(x - rows[:, 4]).max() < 2e-6
    sim.destroy()


def test_tracers_can_be_removed_and_reset(wo):
    from waiwera_b200 import flow
    m, f, sim, own, xb, (x_last, x_last2, al_last, al_last2) = gpu_case(wo, flow, "we", 2)
    a1 = sim.tracer_balances()
    assert sim.set_tracers([2]) == 0                     # replaces the tracer set: vapour tracer only
    a2 = sim.tracer_balances()
    assert a2.shape == (m.nowned,) and np.array_equal(a2, a1.reshape(-1, 2)[:, 1])
    assert sim.set_tracers([]) == 0
    with pytest.raises(Exception):
        sim.tracer_balances()
    sim.destroy()


def test_full_size_tracer_properties(wo):
    """BASELINE config 2 size (100^3 cells, eos_we, top 20 layers two-phase): size-independent properties of the
    tracer step.  Closed box, no sources, no decay => the liquid tracer's mass sum_i V_i Al_i x_i is conserved by
    A x = b (advective and diffusive entries cancel in the volume-weighted column sums); the vapour tracer is
    pinned to zero where there is no vapour; x >= 0 (upstream weighting makes A an M-matrix); the solve is
    deterministic.  Prints the device times of the assembly kernel and of the solve."""
    import json
    from waiwera_b200 import flow
    from util import make_problem
    m, y, region, prm = make_problem(wo, dims=(100, 100, 100), two_phase_layers=20)
    sim = gpu_flow(wo, flow, m, prm, y, region)
    err, _ = sim.lhs(y)
    assert err == 0
    assert sim.set_tracers([1, 2], diffusion=[1.0e-6, 1.0e-5]) == 0
    n, nt = m.nowned, 2
    rng = np.random.default_rng(7)
    x0 = rng.uniform(0.0, 0.01, (n, nt))
    two_phase = region[:n] == 4
    x0[~two_phase, 1] = 0.0                      # no vapour: no vapour tracer
    x0 = np.ascontiguousarray(x0.reshape(-1))
    al0 = sim.tracer_balances()
    from waiwera_b200 import mesh as wmesh
    sim.set_pc_blocks(wmesh.cube_blocks(m, 10))  # block Jacobi over 10^3-cell cubes, ILU(0) inside (as bench.py)
    dt = 1.0e6
    x, al, reason, its = sim.tracer_solve(dt, al0, x0, opts=flow.ksp_opts(type=flow.KSP_BCGS, rtol=1e-12))
    assert reason > 0
    xa, x0a, ala, al0a = x.reshape(n, nt), x0.reshape(n, nt), al.reshape(n, nt), al0.reshape(n, nt)
    vol = m.cell_geom[:n, 3]
    mass0, mass1 = np.sum(vol * al0a[:, 0] * x0a[:, 0]), np.sum(vol * ala[:, 0] * xa[:, 0])
    assert abs(mass1 - mass0) <= 1e-9 * mass0, (mass0, mass1)
    assert np.array_equal(ala, al0a)             # same state: same balance coefficients
    assert (xa[~two_phase, 1] == 0.0).all()
    assert xa.min() >= -1e-12                    # M-matrix (upstream weighting), b >= 0 => x >= 0
    x2, _, reason2, its2 = sim.tracer_solve(dt, al0, x0, opts=flow.ksp_opts(type=flow.KSP_BCGS, rtol=1e-12))
    assert its2 == its and np.array_equal(x2, x)
    t_setup, t_solve = sim.timer("tracer_setup"), sim.timer("tracer_solve")
    print("TRACER_FULL_SIZE " + json.dumps(dict(cells=n, tracers=nt, ksp_iterations=its, setup_ms=t_setup[0] / max(t_setup[1], 1),
                                                solve_ms=t_solve[0] / max(t_solve[1], 1))))
    sim.destroy()
