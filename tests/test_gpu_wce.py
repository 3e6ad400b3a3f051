"""GPU parity for eos_wce (BASELINE config 4 physics: water + CO2 + energy, 3 primaries,
BAIJ block size 3): fluid records, L / R / BE residual, local-FD Jacobian vs the oracle's coloured FD Jacobian,
transitions with the partial-pressure clamp, and a Newton solve -- all through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from util import SEED, make_problem_wce, oracle_flow, gpu_flow, relerr

pytestmark = pytest.mark.gpu
RESIDUAL_TOL = 1e-10


@pytest.fixture(scope="module")
def flow():
    from waiwera_b200 import flow as _flow
    return _flow


def curves(wo, which):
    if which == 0:
        return None, None
    return (wo.make_relperm("corey", slr=0.3, ssr=0.05),
            wo.make_cappress("van_genuchten", P0=0.125e5, lambda_=0.45, slr=1e-3, sls=1.0, Pmax=1e6))


CASES = [dict(thermo=0, two_phase_layers=0), dict(thermo=0, two_phase_layers=2), dict(thermo=1, two_phase_layers=2)]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_wce_fluid_records_match_oracle(wo, flow, case):
    rp, cp = curves(wo, case % 2)
    m, y, region, prm = make_problem_wce(wo, relperm=rp, cappress=cp, **CASES[case])
    ref = oracle_flow(wo, m, prm, y, region)
    sim = gpu_flow(wo, flow, m, prm, y, region)
    assert sim.np == 3 and sim.dof == 26
    a, b = ref.fluid(), sim.fluid()
    for col in (2, 4):
        assert np.array_equal(a[:, col], b[:, col])
    scale = np.maximum(np.abs(a).max(axis=0), 1e-300)
    # CO2 properties go through pow / log10 (CUDA vs glibc differ by an ulp or two); IFC-67 adds its T_sat Newton tolerance
    tol = 1e-12 if CASES[case]["thermo"] == 0 else 5e-11
    assert (np.abs(a - b) / scale).max() < tol
    sim.destroy()


@pytest.mark.parametrize("case", range(len(CASES)))
def test_wce_residual_matches_oracle(wo, flow, case):
    rp, cp = curves(wo, case % 2)
    m, y, region, prm = make_problem_wce(wo, relperm=rp, cappress=cp, **CASES[case])
    ref = oracle_flow(wo, m, prm, y, region)
    sim = gpu_flow(wo, flow, m, prm, y, region)
    e0, L0 = ref.lhs(y)
    e1, L1 = sim.lhs(y)
    assert e0 == e1 == 0
    assert relerr(L1, L0) < 1e-12
    rng = np.random.default_rng(SEED + case)
    y2 = y * (1 + 1e-4 * rng.uniform(-1, 1, len(y)))
    dt = 1.0e6
    e0, lhs0, rhs0, r0 = ref.residual(y2, L0, dt)
    e1, lhs1, rhs1, r1 = sim.residual(y2, L0, dt)
    assert e0 == e1 == 0
    assert relerr(lhs1, lhs0) < 1e-12
    assert relerr(rhs1, rhs0) < RESIDUAL_TOL
    assert relerr(r1, r0) < RESIDUAL_TOL
    assert abs(np.linalg.norm(r1) - np.linalg.norm(r0)) <= RESIDUAL_TOL * np.linalg.norm(r0)
    mv0, ml0 = wo.max_scaled(r0, L0, 1.0)
    mv1, ml1 = sim.max_scaled(r1, L0, 1.0)
    assert ml0 == ml1 and abs(mv0 - mv1) <= RESIDUAL_TOL * abs(mv0)
    sim.destroy()


@pytest.mark.parametrize("case", [0, 1])
def test_wce_jacobian_matches_oracle(wo, flow, case):
    rp, cp = curves(wo, case % 2)
    m, y, region, prm = make_problem_wce(wo, relperm=rp, cappress=cp, **CASES[case])
    ref = oracle_flow(wo, m, prm, y, region)
    sim = gpu_flow(wo, flow, m, prm, y, region)
    _, L0 = ref.lhs(y)
    rng = np.random.default_rng(SEED + 10 + case)
    y2 = y * (1 + 1e-4 * rng.uniform(-1, 1, len(y)))
    dt = 1.0e6
    e, _, _, F0 = ref.residual(y2, L0, dt)
    assert e == 0
    A = ref.bsr()
    nb = A.contents.nb
    color = np.zeros(nb, np.int32)
    nc = wo.lib().wo_bsr_coloring(A, wo.ip(color))
    assert wo.lib().wo_fd_jacobian(ref.h, wo.dp(y2), wo.dp(L0), dt, wo.dp(F0), wo.ip(color), nc, 1e-8, 1e-2, A) == 0
    rowptr, colidx, val = [a.copy() for a in wo.bsr_arrays(A)]
    nb_g, bs, rp_g, ci_g = sim.jacobian_pattern()
    assert nb_g == m.nowned and bs == 3
    assert np.array_equal(rp_g, rowptr) and np.array_equal(ci_g, colidx)
    assert sim.jacobian(y2, L0, dt) == 0
    Jl = sim.jacobian_values()
    rows = np.repeat(np.arange(nb), np.diff(rowptr))
    v3 = np.abs(val).reshape(-1, 3, 3)  # [block][col][row]
    rowmax = np.zeros((nb, 3))
    for ii in range(3):
        np.maximum.at(rowmax[:, ii], rows, v3[:, :, ii].max(axis=1))
    scale = np.tile(rowmax[rows], (1, 3))  # entry q = col*3 + row -> row = q % 3
    assert (np.abs(Jl - val) / np.maximum(scale, 1e-300)).max() < 1e-5
    assert sim.jacobian(y2, L0, dt, colored=True) == 0
    Jc = sim.jacobian_values()
    assert (np.abs(Jc - Jl) / np.maximum(scale, 1e-300)).max() < 1e-12
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()


def test_wce_transitions_match_oracle(wo, flow):
    m, y, region, prm = make_problem_wce(wo, two_phase_layers=3)
    ref = oracle_flow(wo, m, prm, y, region)
    sim = gpu_flow(wo, flow, m, prm, y, region)
    ref.L.wo_flow_pre_eval(ref.h, wo.dp(y), None, 0)
    assert sim.pre_eval(y) == 0
    ref.L.wo_flow_pre_iteration(ref.h)
    sim.pre_iteration()
    rng = np.random.default_rng(SEED + 21)
    n = m.nowned
    search = np.zeros(3 * n)
    tp = region == 4
    search[1::3] = np.where(tp, rng.choice([-0.7, 0.0, 0.9], n), -rng.choice([0.0, 0.0, 1.4], n))
    search[0::3] = rng.uniform(-0.05, 0.05, n)
    search[2::3] = rng.choice([0.0, 0.0, 0.3, -1.1], n) * y[2::3]   # some cells driven to Pg < 0 or Pg > P
    ynew = y - search
    s0, y0 = search.copy(), ynew.copy()
    s1, y1 = search.copy(), ynew.copy()
    cs0, cy0 = C.c_int(), C.c_int()
    e0 = ref.L.wo_flow_fluid_transitions(ref.h, wo.dp(y), wo.dp(s0), wo.dp(y0), C.byref(cs0), C.byref(cy0))
    e1, cs1, cy1 = sim.fluid_transitions(y, s1, y1)
    assert e0 == e1 == 0
    assert cs0.value == cs1
    r0, r1 = ref.regions(), sim.regions()
    assert np.array_equal(r0[:n], r1[:n])
    assert (r0[:n] != region).sum() > 10
    assert np.abs(y0 - y1).max() <= 1e-12 * np.abs(y0).max()
    assert np.abs(s0 - s1).max() <= 1e-12 * max(np.abs(s0).max(), 1.0)
    assert (y1[2::3] >= 0).all()   # the clamp of check_primary_variables was applied
    sim.destroy()


@pytest.mark.parametrize("ksp", [0, 1])
def test_wce_newton_step_matches_oracle(wo, flow, ksp):
    from test_gpu_newton import oracle_newton
    m, y, region, prm = make_problem_wce(wo, dims=(6, 5, 8), two_phase_layers=2)
    ref = oracle_flow(wo, m, prm, y, region)
    sim = gpu_flow(wo, flow, m, prm, y, region)
    _, L0 = ref.lhs(y)
    sim.lhs(y)
    y0, res0 = oracle_newton(wo, ref, y, L0, 1.0e5, 2, ksp, 1)
    y1 = y.copy()
    res1 = sim.newton_solve(y1, L0, 1.0e5, flow.newton_opts(pc_type=flow.PC_BJACOBI_ILU0, ksp=flow.ksp_opts(type=ksp)))
    assert res0.reason == res1.reason and res0.iterations == res1.iterations, (res0.reason, res1.reason, res0.iterations, res1.iterations)
    assert relerr(y1, y0) < 1e-6
    sim.destroy()


def test_full_size_wce_properties(wo, flow):
    """BASELINE config 4 size (100x100x50 = 500 k cells, eos_wce, BAIJ bs = 3, 3.46 M blocks): size-independent
    properties -- closed box => the inflows conserve water, CO2 and energy (sum_i V_i R_i = 0 to rounding), the
    evaluation is deterministic, the Jacobian SpMV is linear, per-cell balances agree with the oracle on a sample,
    and one Newton step with BiCGStab + ILU(0) sub-domains reduces the residual."""
    m, y, region, prm = make_problem_wce(wo, dims=(100, 100, 50), two_phase_layers=10)
    sim = gpu_flow(wo, flow, m, prm, y, region)
    e, L = sim.lhs(y)
    assert e == 0
    dt = 1.0e5
    e, lhs, rhs, r = sim.residual(y, L, dt)
    assert e == 0
    vol = m.cell_geom[:m.nowned, 3]
    for k in range(3):
        tot = np.sum(vol * rhs[k::3])
        assert abs(tot) <= 1e-9 * np.sum(vol * np.abs(rhs[k::3]))
    assert np.array_equal(sim.residual(y, L, dt)[3], r)
    eos = wo.lib().wo_eos_create(C.byref(prm))
    for c in range(0, m.nowned, 50021):
        fl = np.zeros(26)
        fl[2] = region[c]
        pr = np.zeros(3)
        wo.lib().wo_eos_unscale(eos, wo.dp(y[3 * c:3 * c + 3].copy()), int(region[c]), wo.dp(pr))
        assert wo.lib().wo_eos_bulk_properties(eos, wo.dp(pr), wo.dp(fl)) == 0
        assert wo.lib().wo_eos_phase_properties(eos, wo.dp(pr), wo.dp(m.rock[c].copy()), wo.dp(fl)) == 0
        bal = np.zeros(3)
        wo.lib().wo_cell_balance(wo.dp(m.rock[c].copy()), wo.dp(fl), 2, 2, 3, wo.dp(bal))
        assert np.allclose(bal, L[3 * c:3 * c + 3], rtol=1e-12, atol=0)
    wo.lib().wo_eos_destroy(eos)
    assert sim.jacobian(y, L, dt) == 0
    nb, bs, rowptr, colidx = sim.jacobian_pattern()
    assert bs == 3 and len(colidx) == 3460000
    J = sim.jacobian_mat()
    rng = np.random.default_rng(SEED)
    x1, x2 = rng.uniform(-1, 1, nb * 3), rng.uniform(-1, 1, nb * 3)
    a1, a2, a12 = np.zeros(nb * 3), np.zeros(nb * 3), np.zeros(nb * 3)
    J.mult(x1, a1)
    J.mult(x2, a2)
    J.mult(1.5 * x1 - 0.25 * x2, a12)
    assert relerr(a12, 1.5 * a1 - 0.25 * a2) < 1e-13
    from waiwera_b200 import mesh as wmesh
    sim.set_pc_blocks(wmesh.cube_blocks(m, 10))
    y1 = y.copy()
    res = sim.newton_solve(y1, L, dt, flow.newton_opts(max_iterations=2, pc_type=flow.PC_BJACOBI_ILU0,
                                                       ksp=flow.ksp_opts(type=flow.KSP_BCGS)))
    assert res.iterations >= 1 and res.reason in (3, 4, -5)
    assert res.max_residual[1] < 0.5 * res.max_residual[0]
    sim.destroy()
