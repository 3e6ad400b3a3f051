"""GPU parity: BAIJ SpMV, preconditioners and Krylov solvers through the C ABI vs the CPU oracle.

The reference holds no test vectors for these PETSc-side operators (parity
unpinned at the operator level); the oracle is written to PETSc 3.22's documented semantics and
these tests hold the CUDA path to it.
"""
import ctypes as C

import numpy as np
import pytest

from util import SEED, make_problem, oracle_flow, gpu_flow, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def flow():
    from waiwera_b200 import flow as _flow
    return _flow


def random_bsr(wo, dims, bs, seed, diag_boost=8.0):
    """FV-adjacency BAIJ pattern of a box mesh with random, diagonally dominant blocks"""
    from waiwera_b200 import mesh as wmesh
    m = wmesh.structured(*dims, seed=seed)
    mm = wo.Mesh()
    fc = np.ascontiguousarray(m.face_cells.reshape(-1))
    mm.ncell, mm.ninterior, mm.nowned, mm.nface = m.ncell, m.ninterior, m.nowned, m.nface
    mm.face_cells, mm.face_geom, mm.cell_geom, mm.rock = wo.ip(fc), wo.dp(m.face_geom.reshape(-1)), wo.dp(m.cell_geom.reshape(-1)), wo.dp(m.rock.reshape(-1))
    A = wo.lib().wo_bsr_from_mesh(C.byref(mm), bs)
    rowptr, colidx, val = wo.bsr_arrays(A)
    rng = np.random.default_rng(seed)
    val[:] = rng.uniform(-1, 1, val.shape)
    rows = np.repeat(np.arange(m.nowned), np.diff(rowptr))
    diag = np.flatnonzero(colidx == rows)
    for k in range(bs):
        val[diag, k * bs + k] += diag_boost
    return m, A, rowptr, colidx, val


@pytest.mark.parametrize("bs,dims", [(2, (9, 8, 7)), (3, (6, 5, 7)), (1, (10, 3, 4)), (2, (1, 1, 50)), (2, (3, 1, 1))])
def test_spmv_matches_oracle(wo, flow, bs, dims):
    """K5: y = A x for bs = 1, 2, 3, ragged rows (boundary cells have fewer blocks)"""
    m, A, rowptr, colidx, val = random_bsr(wo, dims, bs, SEED + bs)
    _, y0, region, prm = make_problem(wo, dims=dims)
    sim = gpu_flow(wo, flow, m, prm, y0, region)
    nb = m.nowned
    M = flow.Mat.create(sim, nb, nb, bs, rowptr, colidx, val)
    rng = np.random.default_rng(SEED)
    x = rng.uniform(-1, 1, nb * bs)
    ref, got = np.zeros(nb * bs), np.zeros(nb * bs)
    wo.lib().wo_bsr_spmv(A, wo.dp(x), wo.dp(ref))
    M.mult(x, got)
    assert relerr(got, ref) < 1e-14
    assert np.abs(got - ref).max() <= 4e-15 * np.abs(val).max() * np.abs(x).max() * 8 * bs
    M.destroy()
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()


@pytest.mark.parametrize("bs", [2, 3])
@pytest.mark.parametrize("pctype,nblocks", [(1, 1), (2, 1), (2, 5)])
def test_pc_apply_matches_oracle(wo, flow, bs, pctype, nblocks):
    """K6: point-block Jacobi, global ILU(0), block-Jacobi ILU(0) with 5 sub-domains"""
    dims = (7, 6, 5)
    m, A, rowptr, colidx, val = random_bsr(wo, dims, bs, SEED + 3 * bs)
    _, y0, region, prm = make_problem(wo, dims=dims)
    sim = gpu_flow(wo, flow, m, prm, y0, region)
    nb = m.nowned
    M = flow.Mat.create(sim, nb, nb, bs, rowptr, colidx, val)
    bor = None
    if nblocks > 1:
        bor = ((np.arange(nb, dtype=np.int64) * nblocks) // nb).astype(np.int32)
    pc_ref = wo.lib().wo_pc_create(A, pctype, wo.ip(bor))
    assert pc_ref
    pc = flow.PC(M, pctype, nblocks, bor)
    rng = np.random.default_rng(SEED + 1)
    r = rng.uniform(-1, 1, nb * bs)
    z0, z1 = np.zeros(nb * bs), np.zeros(nb * bs)
    wo.lib().wo_pc_apply(pc_ref, wo.dp(r), wo.dp(z0))
    pc.apply(r, z1)
    assert relerr(z1, z0) < 1e-12
    # repeated applies (epoch-based ready flags) stay correct
    pc.apply(r, z1)
    assert relerr(z1, z0) < 1e-12
    # refactor after the values change
    val2 = val * 1.5
    M.set_values(val2)
    assert pc.refactor() == 0
    pc.apply(r, z1)
    assert relerr(z1, z0 / 1.5) < 1e-12
    wo.lib().wo_pc_destroy(pc_ref)
    pc.destroy()
    M.destroy()
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()


def test_ilu0_exact_on_block_tridiagonal(wo, flow):
    """ILU(0) has no dropped fill on a 1-D chain: the PC apply is an exact solve (size-independent property),
    checked at 200 000 rows where the level schedule is one row per level ... the worst case for the sweep."""
    dims = (1, 1, 200000)
    bs = 2
    m, A, rowptr, colidx, val = random_bsr(wo, dims, bs, SEED + 5)
    from waiwera_b200 import mesh as wmesh
    _, y0, region, prm = make_problem(wo, dims=(4, 4, 4))
    sim = gpu_flow(wo, flow, wmesh.structured(4, 4, 4), prm, y0, region)
    nb = m.nowned
    M = flow.Mat.create(sim, nb, nb, bs, rowptr, colidx, val)
    pc = flow.PC(M, 2, 1)
    rng = np.random.default_rng(SEED + 2)
    x = rng.uniform(-1, 1, nb * bs)
    b, z = np.zeros(nb * bs), np.zeros(nb * bs)
    M.mult(x, b)
    pc.apply(b, z)
    assert relerr(z, x) < 1e-10
    pc.destroy()
    M.destroy()
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()


@pytest.mark.parametrize("ksptype", [0, 1])
@pytest.mark.parametrize("pctype,nblocks", [(0, 1), (1, 1), (2, 1), (2, 4)])
def test_ksp_matches_oracle(wo, flow, ksptype, pctype, nblocks):
    """K7: GMRES(30) / BiCGStab, left PC, zero initial guess: same reason, iteration count within 1,
    solution equal to the oracle's well inside the solver tolerance"""
    dims, bs = (8, 7, 6), 2
    m, A, rowptr, colidx, val = random_bsr(wo, dims, bs, SEED + 7, diag_boost=4.0)
    _, y0, region, prm = make_problem(wo, dims=dims)
    sim = gpu_flow(wo, flow, m, prm, y0, region)
    nb = m.nowned
    M = flow.Mat.create(sim, nb, nb, bs, rowptr, colidx, val)
    bor = None if nblocks == 1 else ((np.arange(nb, dtype=np.int64) * nblocks) // nb).astype(np.int32)
    pc_ref = wo.lib().wo_pc_create(A, pctype, wo.ip(bor))
    pc = flow.PC(M, pctype, nblocks, bor)
    rng = np.random.default_rng(SEED + 3)
    b = rng.uniform(-1, 1, nb * bs)
    o = wo.KspOpts()
    o.type, o.restart, o.maxit, o.rtol, o.atol, o.dtol = ksptype, 30, 10000, 1e-8, 1e-50, 1e5
    x0 = np.zeros(nb * bs)
    its0, rn0 = C.c_int(), C.c_double()
    reason0 = wo.lib().wo_ksp_solve(A, pc_ref, C.byref(o), wo.dp(b), wo.dp(x0), C.byref(its0), C.byref(rn0))
    x1 = np.zeros(nb * bs)
    reason1, its1, rn1 = flow.ksp_solve(M, pc, b, x1, flow.ksp_opts(type=ksptype, rtol=1e-8))
    assert reason0 == reason1 and reason1 > 0
    assert abs(its0.value - its1) <= 1
    assert relerr(x1, x0) < 1e-6
    ax = np.zeros(nb * bs)
    M.mult(x1, ax)
    assert relerr(ax, b) < 1e-6
    # check_every must not change the result (kernels are predicated on the device-side flag)
    flow._lib.lib().wb_ksp_set_check_every(1)
    x2 = np.zeros(nb * bs)
    reason2, its2, _ = flow.ksp_solve(M, pc, b, x2, flow.ksp_opts(type=ksptype, rtol=1e-8))
    flow._lib.lib().wb_ksp_set_check_every(4)
    assert (reason2, its2) == (reason1, its1) and np.array_equal(x1, x2)
    wo.lib().wo_pc_destroy(pc_ref)
    pc.destroy()
    M.destroy()
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()


def test_gmres_restart_and_maxit(wo, flow):
    """restart cycles and the iteration limit: weak PC, tight tolerance, restart 5"""
    dims, bs = (6, 6, 6), 2
    m, A, rowptr, colidx, val = random_bsr(wo, dims, bs, SEED + 9, diag_boost=2.5)
    _, y0, region, prm = make_problem(wo, dims=dims)
    sim = gpu_flow(wo, flow, m, prm, y0, region)
    nb = m.nowned
    M = flow.Mat.create(sim, nb, nb, bs, rowptr, colidx, val)
    pc_ref = wo.lib().wo_pc_create(A, 1, None)
    pc = flow.PC(M, 1, 1)
    b = np.random.default_rng(SEED + 4).uniform(-1, 1, nb * bs)
    for maxit in (10000, 7):
        o = wo.KspOpts()
        o.type, o.restart, o.maxit, o.rtol, o.atol, o.dtol = 0, 5, maxit, 1e-9, 1e-50, 1e5
        x0 = np.zeros(nb * bs)
        its0, rn0 = C.c_int(), C.c_double()
        reason0 = wo.lib().wo_ksp_solve(A, pc_ref, C.byref(o), wo.dp(b), wo.dp(x0), C.byref(its0), C.byref(rn0))
        x1 = np.zeros(nb * bs)
        reason1, its1, rn1 = flow.ksp_solve(M, pc, b, x1, flow.ksp_opts(type=0, restart=5, maxit=maxit, rtol=1e-9))
        assert reason0 == reason1
        assert abs(its0.value - its1) <= (1 if maxit > 7 else 0)
        assert relerr(x1, x0) < 1e-5
    wo.lib().wo_pc_destroy(pc_ref)
    pc.destroy()
    M.destroy()
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()


def test_full_size_spmv_properties(wo, flow):
    """BASELINE config 2 size (1 M rows, bs 2, 6.94 M blocks): linearity A(ax+by) = aAx + bAy, a
    checksum (1^T A x = (A^T 1)^T x computed on the host from the blocks) and agreement with the oracle SpMV"""
    dims, bs = (100, 100, 100), 2
    m, A, rowptr, colidx, val = random_bsr(wo, dims, bs, SEED + 11)
    _, y0, region, prm = make_problem(wo, dims=(4, 4, 4))
    from waiwera_b200 import mesh as wmesh
    sim = gpu_flow(wo, flow, wmesh.structured(4, 4, 4), prm, y0, region)
    nb = m.nowned
    assert len(colidx) == 6940000
    M = flow.Mat.create(sim, nb, nb, bs, rowptr, colidx, val)
    rng = np.random.default_rng(SEED + 5)
    x, y = rng.uniform(-1, 1, nb * bs), rng.uniform(-1, 1, nb * bs)
    ax, ay, axy = np.zeros(nb * bs), np.zeros(nb * bs), np.zeros(nb * bs)
    M.mult(x, ax)
    M.mult(y, ay)
    M.mult(2.5 * x - 0.75 * y, axy)
    assert relerr(axy, 2.5 * ax - 0.75 * ay) < 1e-14
    colsum = np.zeros((nb, bs))
    v3 = val.reshape(-1, bs, bs)  # [block][col][row]
    np.add.at(colsum, colidx, v3.sum(axis=2))
    assert abs(ax.sum() - (colsum.reshape(-1) * x).sum()) < 1e-9 * np.abs(ax).sum()
    ref = np.zeros(nb * bs)
    wo.lib().wo_bsr_spmv(A, wo.dp(x), wo.dp(ref))
    assert relerr(ax, ref) < 1e-14
    M.destroy()
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()


@pytest.mark.parametrize("bs,dims,box", [(2, (9, 8, 7), (3, 4, 4)), (3, (6, 5, 7), (3, 3, 4)), (1, (10, 3, 4), (5, 3, 2)),
                                         (2, (24, 20, 20), (10, 10, 10))])
def test_asm_overlap1_matches_oracle(wo, flow, bs, dims, box):
    """PCASM (restricted, overlap 1) + ILU(0) on box sub-domains: PC apply, refactor, and a GMRES(30) solve against
    the oracle (whose ASM is pinned to the definition in test_oracle_linalg_independent.py); fewer iterations than
    block Jacobi on the same sub-domains"""
    import ctypes as C
    from test_gpu_fused import box_blocks
    m, A, rowptr, colidx, val = random_bsr(wo, dims, bs, SEED + 31 + bs, diag_boost=3.0)
    _, y0, region, prm = make_problem(wo, dims=dims)
    sim = gpu_flow(wo, flow, m, prm, y0, region)
    nb = m.nowned
    M = flow.Mat.create(sim, nb, nb, bs, rowptr, colidx, val)
    bor = box_blocks(dims, box)
    pc_ref = wo.lib().wo_pc_create(A, wo.PC_ASM_ILU0, wo.ip(bor))
    pc = flow.PC(M, flow.PC_ASM_ILU0, 1, bor)
    rng = np.random.default_rng(SEED)
    r = rng.uniform(-1, 1, nb * bs)
    z0, z1 = np.zeros(nb * bs), np.zeros(nb * bs)
    wo.lib().wo_pc_apply(pc_ref, wo.dp(r), wo.dp(z0))
    pc.apply(r, z1)
    assert relerr(z1, z0) < 1e-12
    o = wo.KspOpts()
    o.type, o.restart, o.maxit, o.rtol, o.atol, o.dtol = 0, 30, 10000, 1e-8, 1e-50, 1e5
    x0, x1 = np.zeros(nb * bs), np.zeros(nb * bs)
    its0, rn0 = C.c_int(), C.c_double()
    reason0 = wo.lib().wo_ksp_solve(A, pc_ref, C.byref(o), wo.dp(r), wo.dp(x0), C.byref(its0), C.byref(rn0))
    reason, its, rn = flow.ksp_solve(M, pc, r, x1, flow.ksp_opts(type=0, restart=30, maxit=10000, rtol=1e-8))
    assert reason == reason0 > 0 and abs(its - its0.value) <= 1 and relerr(x1, x0) < 1e-6
    pcb = flow.PC(M, flow.PC_BJACOBI_ILU0, 1, bor)
    xb = np.zeros(nb * bs)
    _, its_bj, _ = flow.ksp_solve(M, pcb, r, xb, flow.ksp_opts(type=0, restart=30, maxit=10000, rtol=1e-8))
    assert its <= its_bj
    pcb.destroy()
    M.set_values(val * 1.5)
    assert pc.refactor() == 0
    pc.apply(r, z1)
    assert relerr(z1, z0 / 1.5) < 1e-12
    wo.lib().wo_pc_destroy(pc_ref)
    pc.destroy()
    M.destroy()
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()
