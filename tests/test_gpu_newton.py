"""GPU parity: a full Newton solve of one backward-Euler step (SNES newtonls replay) vs the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

from util import SEED, make_problem, oracle_flow, gpu_flow, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def flow():
    from waiwera_b200 import flow as _flow
    return _flow


def oracle_newton(wo, ref, y, L0, dt, pc_type, ksp_type, nblocks, max_it=8):
    A = ref.bsr()
    nb = A.contents.nb
    color = np.zeros(nb, np.int32)
    nc = wo.lib().wo_bsr_coloring(A, wo.ip(color))
    o = wo.NewtonOpts()
    o.max_iterations, o.min_iterations = max_it, 0
    o.rel_tol, o.abs_tol, o.update_rel_tol, o.update_abs_tol = 1e-5, 1.0, 1e-10, 1.0
    o.fd_err, o.fd_umin, o.pc_type = 1e-8, 1e-2, pc_type
    o.ksp.type, o.ksp.restart, o.ksp.maxit = ksp_type, 30, 10000
    o.ksp.rtol, o.ksp.atol, o.ksp.dtol = 1e-5, 1e-50, 1e5
    bor = None if nblocks == 1 else ((np.arange(nb, dtype=np.int64) * nblocks) // nb).astype(np.int32)
    res = wo.NewtonResult()
    yy = y.copy()
    wo.lib().wo_newton_solve_be(ref.h, A, wo.ip(color), nc, wo.ip(bor), C.byref(o), dt, wo.dp(L0), wo.dp(yy), C.byref(res))
    wo.lib().wo_bsr_destroy(A)
    return yy, res


@pytest.mark.parametrize("cfg", [
    dict(thermo=0, two_phase_layers=0, top_boundary=True, pc=2, ksp=0, nblocks=1),
    dict(thermo=0, two_phase_layers=2, top_boundary=True, pc=2, ksp=1, nblocks=1),
    dict(thermo=1, two_phase_layers=2, top_boundary=False, pc=1, ksp=0, nblocks=1),
    dict(thermo=0, two_phase_layers=0, top_boundary=True, pc=2, ksp=0, nblocks=3),
])
def test_newton_step_matches_oracle(wo, flow, cfg):
    """one BE step from the hydrostatic state with a cold Dirichlet top: same convergence reason and
    Newton iteration count, per-iteration max scaled residual and final solution equal to solver tolerance"""
    m, y, region, prm = make_problem(wo, dims=(6, 5, 8), thermo=cfg["thermo"], two_phase_layers=cfg["two_phase_layers"],
                                     top_boundary=cfg["top_boundary"])
    ref = oracle_flow(wo, m, prm, y, region)
    sim = gpu_flow(wo, flow, m, prm, y, region)
    _, L0 = ref.lhs(y)
    ref.L.wo_flow_pre_timestep(ref.h)
    sim.lhs(y)
    sim.pre_timestep()
    dt = 1.0e6
    y0, res0 = oracle_newton(wo, ref, y, L0, dt, cfg["pc"], cfg["ksp"], cfg["nblocks"])
    o = flow.newton_opts(pc_type=cfg["pc"], pc_nblocks=cfg["nblocks"], ksp=flow.ksp_opts(type=cfg["ksp"]))
    y1 = y.copy()
    res1 = sim.newton_solve(y1, L0, dt, o)
    assert res0.reason == res1.reason and res1.reason > 0
    assert res0.iterations == res1.iterations
    for it in range(res0.iterations + 1):
        a, b = res0.max_residual[it], res1.max_residual[it]
        # early iterations agree to rounding; later ones inherit the linear-solve tolerance
        assert abs(a - b) <= (1e-9 if it == 0 else 2e-2) * max(abs(a), 1e-8), (it, a, b)
    assert abs(res0.max_residual[0] - res1.max_residual[0]) <= 1e-10 * res0.max_residual[0]
    assert relerr(y1, y0) < 1e-6
    assert np.array_equal(ref.regions()[:m.nowned], sim.regions()[:m.nowned])
    assert abs(res0.linear_iterations - res1.linear_iterations) <= res0.iterations + 1
    sim.destroy()
