"""GPU parity of the persistent (sub-domain-resident) GMRES kernel: the whole KSPSolve as one cooperative kernel
(wb_fused.cu) against the CPU oracle's GMRES and against the launch-per-operation solver of wb_linalg.cu, which runs
the same arithmetic with one kernel per operation.  Sub-domain shapes are chosen to hit every CTA layout: fewer
sub-domains than SMs, several per CTA (1, 2 and 4 groups), ragged sub-domains, rows of 2 and 8 blocks (MINC), block
sizes 1 / 2 / 3, restarts shorter than the solve, the iteration limit, a zero right-hand side."""
import ctypes as C

import numpy as np
import pytest

from test_gpu_linalg import random_bsr
from util import SEED, make_problem, gpu_flow, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def flow():
    from waiwera_b200 import flow as _flow
    return _flow


def box_blocks(dims, box):
    nx, ny, nz = dims
    idx = np.arange(nx * ny * nz)
    i, j, k = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    bx, by = -(-nx // box[0]), -(-ny // box[1])
    key = (i // box[0]) + bx * ((j // box[1]) + by * (k // box[2]))
    _, inv = np.unique(key, return_inverse=True)
    return inv.astype(np.int32)


def solve_three_ways(wo, flow, sim, M, A, bor, b, restart, maxit, rtol, norm_mode=0, ksp=0):
    """oracle, launch-per-operation GPU solver, persistent kernel (norm_mode 0: VecNorm as a second reduction, the
    reference's arithmetic; 2: the norm from the dot-product pass; ksp 0: GMRES, 1: BiCGStab)"""
    L = flow._lib.lib()
    n = len(b)
    pc_ref = wo.lib().wo_pc_create(A, 2, wo.ip(bor))
    o = wo.KspOpts()
    o.type, o.restart, o.maxit, o.rtol, o.atol, o.dtol = ksp, restart, maxit, rtol, 1e-50, 1e5
    x0 = np.zeros(n)
    its0, rn0 = C.c_int(), C.c_double()
    reason0 = wo.lib().wo_ksp_solve(A, pc_ref, C.byref(o), wo.dp(b), wo.dp(x0), C.byref(its0), C.byref(rn0))
    wo.lib().wo_pc_destroy(pc_ref)
    res = [(reason0, its0.value, rn0.value, x0)]
    opts = flow.ksp_opts(type=ksp, restart=restart, maxit=maxit, rtol=rtol)
    L.wb_ksp_set_fused_norm(norm_mode)
    for fused in (0, 2):   # 2: the persistent kernel wherever it can run (1 = automatic choice)
        L.wb_ksp_set_fused(fused)
        pc = flow.PC(M, 2, 1, bor)
        x = np.full(n, 7.0)          # the solvers start from zero whatever the buffer holds
        launches = sim.launches()
        reason, its, rn = flow.ksp_solve(M, pc, b, x, opts)
        res.append((reason, its, rn, x, sim.launches() - launches))
        pc.destroy()
    L.wb_ksp_set_fused(1)
    L.wb_ksp_set_fused_norm(1)
    return res


CASES = [
    # dims, bs, sub-domain box, restart, maxit, rtol
    dict(dims=(8, 8, 8), bs=2, box=(4, 4, 4), restart=30, maxit=10000, rtol=1e-8),     # 8 sub-domains, one per CTA
    dict(dims=(12, 12, 10), bs=2, box=(2, 2, 2), restart=30, maxit=10000, rtol=1e-8),  # 180 sub-domains: 2 groups per CTA
    dict(dims=(24, 24, 16), bs=2, box=(3, 3, 2), restart=30, maxit=10000, rtol=1e-8),  # 512 sub-domains: 4 groups per CTA
    dict(dims=(10, 9, 7), bs=2, box=(4, 4, 3), restart=5, maxit=10000, rtol=1e-9),     # ragged boxes, many restart cycles
    dict(dims=(10, 9, 7), bs=2, box=(4, 4, 3), restart=5, maxit=13, rtol=1e-12),       # iteration limit inside a cycle
    dict(dims=(9, 7, 8), bs=3, box=(3, 4, 4), restart=30, maxit=10000, rtol=1e-8),     # bs = 3 (eos_wce)
    dict(dims=(12, 5, 4), bs=1, box=(4, 5, 2), restart=30, maxit=10000, rtol=1e-8),    # bs = 1 (eos_w)
    dict(dims=(20, 20, 20), bs=2, box=(10, 10, 10), restart=30, maxit=10000, rtol=1e-8),  # the bench's 10^3 cubes
]


@pytest.mark.parametrize("norm_mode", [0, 2])
@pytest.mark.parametrize("case", CASES)
def test_fused_gmres_matches_oracle_and_unfused(wo, flow, case, norm_mode):
    dims, bs = case["dims"], case["bs"]
    m, A, rowptr, colidx, val = random_bsr(wo, dims, bs, SEED + 21, diag_boost=3.0)
    _, y0, region, prm = make_problem(wo, dims=dims)
    sim = gpu_flow(wo, flow, m, prm, y0, region)
    nb = m.nowned
    M = flow.Mat.create(sim, nb, nb, bs, rowptr, colidx, val)
    bor = box_blocks(dims, case["box"])
    b = np.random.default_rng(SEED + 5).uniform(-1, 1, nb * bs)
    ref, unfused, fused = solve_three_ways(wo, flow, sim, M, A, bor, b, case["restart"], case["maxit"], case["rtol"],
                                           norm_mode)
    assert fused[4] <= 3, "the persistent path is one solver launch (+ the matrix refresh)"
    assert unfused[4] > 3 * max(unfused[1], 1)
    assert ref[0] == unfused[0] == fused[0]
    tol_its = 0 if case["maxit"] < 100 else 1
    assert abs(fused[1] - ref[1]) <= tol_its and abs(fused[1] - unfused[1]) <= tol_its, (ref[1], unfused[1], fused[1])
    # same operations, different association of the dot-product sums: agreement far inside the solver tolerance
    # (a solve cut off by the iteration limit compares iterates after the same number of iterations: rounding only)
    lim = 1e-6 if case["maxit"] > 100 else 1e-8
    assert relerr(fused[3], unfused[3]) < lim and relerr(fused[3], ref[3]) < lim
    if case["maxit"] > 100:
        ax = np.zeros(nb * bs)
        M.mult(fused[3], ax)
        assert relerr(ax, b) < 1e-5
        assert abs(fused[2] - ref[2]) <= 1e-6 * abs(ref[2]) + 1e-3 * case["rtol"] * np.linalg.norm(b)
    M.destroy()
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()


@pytest.mark.parametrize("case", [CASES[0], CASES[1], CASES[2], CASES[3], CASES[5], CASES[6], CASES[7],
                                  dict(dims=(10, 9, 7), bs=2, box=(4, 4, 3), restart=30, maxit=7, rtol=1e-12)])
def test_fused_bcgs_matches_oracle_and_unfused(wo, flow, case):
    """KSPSolve_BCGS in the persistent kernel against the oracle's BiCGStab and the launch-per-operation GPU BiCGStab"""
    dims, bs = case["dims"], case["bs"]
    m, A, rowptr, colidx, val = random_bsr(wo, dims, bs, SEED + 23, diag_boost=3.0)
    _, y0, region, prm = make_problem(wo, dims=dims)
    sim = gpu_flow(wo, flow, m, prm, y0, region)
    nb = m.nowned
    M = flow.Mat.create(sim, nb, nb, bs, rowptr, colidx, val)
    bor = box_blocks(dims, case["box"])
    b = np.random.default_rng(SEED + 6).uniform(-1, 1, nb * bs)
    ref, unfused, fused = solve_three_ways(wo, flow, sim, M, A, bor, b, 30, case["maxit"], case["rtol"], 0, ksp=1)
    assert fused[4] <= 3 and unfused[4] > 3 * max(unfused[1], 1)
    assert ref[0] == unfused[0] == fused[0], (ref[:3], unfused[:3], fused[:3])
    # BiCGStab's iteration count moves with the summation order of its dot products (three implementations, three
    # orders: e.g. 71 / 66 / 72): a band; a solve cut off by the iteration limit runs the same number of iterations
    tol_its = 0 if case["maxit"] < 100 else max(2, int(0.15 * ref[1]))
    assert abs(fused[1] - ref[1]) <= tol_its and abs(fused[1] - unfused[1]) <= tol_its, (ref[1], unfused[1], fused[1])
    lim = 1e-6 if case["maxit"] > 100 else 1e-8
    assert relerr(fused[3], unfused[3]) < lim and relerr(fused[3], ref[3]) < lim
    if case["maxit"] > 100:
        ax = np.zeros(nb * bs)
        M.mult(fused[3], ax)
        assert relerr(ax, b) < 1e-5
    M.destroy()
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()


def test_fused_bcgs_zero_rhs_and_repeat(wo, flow):
    dims, bs = (8, 8, 6), 2
    m, A, rowptr, colidx, val = random_bsr(wo, dims, bs, SEED + 24, diag_boost=3.0)
    _, y0, region, prm = make_problem(wo, dims=dims)
    sim = gpu_flow(wo, flow, m, prm, y0, region)
    nb = m.nowned
    M = flow.Mat.create(sim, nb, nb, bs, rowptr, colidx, val)
    bor = box_blocks(dims, (4, 4, 3))
    pc = flow.PC(M, 2, 1, bor)
    x = np.ones(nb * bs)
    launches = sim.launches()
    reason, its, rn = flow.ksp_solve(M, pc, np.zeros(nb * bs), x, flow.ksp_opts(type=1))
    assert sim.launches() - launches <= 3
    assert reason > 0 and its == 0 and not x.any()
    b = np.random.default_rng(3).uniform(-1, 1, nb * bs)
    xs = []
    for _ in range(3):
        x = np.zeros(nb * bs)
        r = flow.ksp_solve(M, pc, b, x, flow.ksp_opts(type=1, rtol=1e-10))
        xs.append((r, x))
    assert xs[0][0] == xs[1][0] == xs[2][0] and xs[0][0][0] > 0
    assert np.array_equal(xs[0][1], xs[1][1]) and np.array_equal(xs[0][1], xs[2][1])
    pc.destroy()
    M.destroy()
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()


def test_fused_gmres_zero_rhs_and_repeat(wo, flow):
    """zero right-hand side converges at once with x = 0; repeated solves with one PC reproduce themselves bit for
    bit (fixed reduction orders, no atomics on the data path)"""
    dims, bs = (8, 8, 6), 2
    m, A, rowptr, colidx, val = random_bsr(wo, dims, bs, SEED + 22, diag_boost=3.0)
    _, y0, region, prm = make_problem(wo, dims=dims)
    sim = gpu_flow(wo, flow, m, prm, y0, region)
    nb = m.nowned
    M = flow.Mat.create(sim, nb, nb, bs, rowptr, colidx, val)
    bor = box_blocks(dims, (4, 4, 3))
    pc = flow.PC(M, 2, 1, bor)
    x = np.ones(nb * bs)
    reason, its, rn = flow.ksp_solve(M, pc, np.zeros(nb * bs), x, flow.ksp_opts(type=0))
    assert reason > 0 and its == 0 and not x.any()
    b = np.random.default_rng(3).uniform(-1, 1, nb * bs)
    xs = []
    for _ in range(3):
        x = np.zeros(nb * bs)
        r = flow.ksp_solve(M, pc, b, x, flow.ksp_opts(type=0, rtol=1e-10))
        xs.append((r, x))
    assert xs[0][0] == xs[1][0] == xs[2][0]
    assert np.array_equal(xs[0][1], xs[1][1]) and np.array_equal(xs[0][1], xs[2][1])
    # new matrix values, same pattern: refactor, solve again (the kernel's copy of the matrix follows the values)
    val2 = val.copy()
    val2 *= 1.5
    M.set_values(val2)
    pc.refactor()
    x2 = np.zeros(nb * bs)
    r2 = flow.ksp_solve(M, pc, b, x2, flow.ksp_opts(type=0, rtol=1e-10))
    assert r2[0] > 0 and relerr(1.5 * x2, xs[0][1]) < 1e-8
    pc.destroy()
    M.destroy()
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()


def test_fused_gmres_minc_rows(wo, flow):
    """irregular rows (fracture cells 8 blocks, matrix cells 2): a Newton system of a MINC mesh, sub-domains = boxes of
    fracture cells with their matrix cells"""
    from waiwera_b200 import mesh as wmesh
    from util import oracle_flow
    base = wmesh.structured(8, 6, 6, dx=10.0, seed=SEED)
    m = wmesh.add_minc(base, volumes=(0.1, 0.9), spacing=(50., 50., 50.), matrix_permeability_factor=0.01)
    primary, region = wmesh.hydrostatic_state(base, seed=SEED)
    y = np.ascontiguousarray(wmesh.scale_primaries(np.concatenate([primary, primary]), np.concatenate([region, region]))).reshape(-1)
    region2 = np.concatenate([region, region])
    prm = wo.make_params()
    ref = oracle_flow(wo, m, prm, y, region2)
    sim = gpu_flow(wo, flow, m, prm, y, region2)
    e, L0 = sim.lhs(y)
    e, _, _, r = sim.residual(y * (1 + 1e-4), L0, 1.0e6)
    assert e == 0 and sim.jacobian(y * (1 + 1e-4), L0, 1.0e6) == 0
    J = sim.jacobian_mat()
    nb, bs, rowptr, colidx = sim.jacobian_pattern()
    vals = sim.jacobian_values()
    A = ref.bsr()
    wo.bsr_arrays(A)[2][:] = vals
    bor = wmesh.minc_cube_blocks(m, 4)
    out = solve_three_ways(wo, flow, sim, J, A, bor, r, 30, 10000, 1e-8)
    assert out[0][0] == out[1][0] == out[2][0] > 0
    assert abs(out[2][1] - out[0][1]) <= 1 and abs(out[2][1] - out[1][1]) <= 1
    assert relerr(out[2][3], out[0][3]) < 1e-6 and relerr(out[2][3], out[1][3]) < 1e-6
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()
