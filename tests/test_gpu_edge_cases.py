"""Edge cases of the path through the C ABI: the smallest meshes the reference's own suites use (single cell -- the
ncg/co2_one_cell benchmark; 1-D columns), meshes without faces, sub-domains of one row, ragged block rows, zero
right-hand sides, sources in every cell, host and device pointers mixed."""
import ctypes as C

import numpy as np
import pytest

from waiwera_b200 import mesh as wmesh
from util import SEED, make_problem, make_problem_wce, oracle_flow, gpu_flow, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def flow():
    from waiwera_b200 import flow as _flow
    return _flow


@pytest.mark.parametrize("eos", ["we", "wce"])
def test_single_cell_mesh(wo, flow, eos):
    """one cell, no faces (test/benchmark/ncg/co2_one_cell): residual = L - L_last - dt * sources, 1x1 block Jacobian"""
    mk = make_problem if eos == "we" else make_problem_wce
    m, y, region, prm = mk(wo, dims=(1, 1, 1))
    assert m.nface == 0
    ref = oracle_flow(wo, m, prm, y, region)
    sim = gpu_flow(wo, flow, m, prm, y, region)
    npv = sim.np
    comp = [1, npv]                      # mass injection + a heat source in the same cell
    rate = [0.5, 2.0e4]
    enth = [8.0e5, 0.0]
    ref.set_sources([0, 0], comp, rate, enth)
    sim.set_sources([0, 0], comp, rate, enth)
    e0, L0 = ref.lhs(y)
    e1, L1 = sim.lhs(y)
    assert e0 == e1 == 0 and relerr(L1, L0) < 1e-12
    y2 = y * 1.0001
    e0, _, rhs0, r0 = ref.residual(y2, L0, 1.0e3)
    e1, _, rhs1, r1 = sim.residual(y2, L0, 1.0e3)
    assert e0 == e1 == 0 and relerr(rhs1, rhs0) < 1e-13 and relerr(r1, r0) < 1e-10
    vol = m.cell_geom[0, 3]
    assert abs(rhs1[0] - 0.5 / vol) < 1e-15 * abs(rhs1[0]) + 1e-300
    nb, bs, rowptr, colidx = sim.jacobian_pattern()
    assert nb == 1 and list(rowptr) == [0, 1] and list(colidx) == [0]
    assert sim.jacobian(y2, L0, 1.0e3) == 0
    yy = y.copy()
    res = sim.newton_solve(yy, L0, 1.0e3, flow.newton_opts(pc_type=flow.PC_BJACOBI_ILU0))
    assert res.reason > 0
    sim.destroy()


def test_one_dimensional_columns_and_production(wo, flow):
    """1-D column (rows of 2-3 blocks, one row per ILU level) with a production well: mobility-weighted flow
    fractions (source.F90:403-438) in the residual and in the diagonal Jacobian blocks"""
    m, y, region, prm = make_problem(wo, dims=(1, 1, 40), two_phase_layers=6)
    ref = oracle_flow(wo, m, prm, y, region)
    sim = gpu_flow(wo, flow, m, prm, y, region)
    cells, comp, rate, enth = [3, 20, 20], [0, 1, 2], [-0.02, 0.01, 500.0], [0.0, 4.0e5, 0.0]
    ref.set_sources(cells, comp, rate, enth)
    sim.set_sources(cells, comp, rate, enth)
    _, L0 = ref.lhs(y)
    sim.lhs(y)
    y2 = y * (1 + 1e-4 * np.random.default_rng(SEED).uniform(-1, 1, len(y)))
    e0, _, rhs0, r0 = ref.residual(y2, L0, 1.0e5)
    e1, _, rhs1, r1 = sim.residual(y2, L0, 1.0e5)
    assert e0 == e1 == 0 and relerr(rhs1, rhs0) < 1e-10 and relerr(r1, r0) < 1e-10
    A = ref.bsr()
    color = np.zeros(A.contents.nb, np.int32)
    nc = wo.lib().wo_bsr_coloring(A, wo.ip(color))
    assert wo.lib().wo_fd_jacobian(ref.h, wo.dp(y2), wo.dp(L0), 1.0e5, wo.dp(r0), wo.ip(color), nc, 1e-8, 1e-2, A) == 0
    rowptr, colidx, val = [a.copy() for a in wo.bsr_arrays(A)]
    assert sim.jacobian(y2, L0, 1.0e5) == 0
    Jl = sim.jacobian_values()
    scale = np.abs(val).max()
    assert np.abs(Jl - val).max() / scale < 1e-6
    # the production cell's diagonal block differs from the source-free one
    sim.set_sources([], [], [], [])
    assert sim.jacobian(y2, L0, 1.0e5) == 0
    J0 = sim.jacobian_values()
    d3 = int(np.flatnonzero(colidx[rowptr[3]:rowptr[4]] == 3)[0]) + rowptr[3]
    assert np.abs(J0[d3] - Jl[d3]).max() > 0
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()


def test_pc_subdomains_of_one_row_and_zero_rhs(wo, flow):
    """every row its own sub-domain (ILU(0) degenerates to point-block Jacobi); zero right-hand side converges at
    iteration 0 with x = 0 (KSPConvergedDefault on rnorm0 = 0)"""
    m, y, region, prm = make_problem(wo, dims=(5, 4, 3))
    sim = gpu_flow(wo, flow, m, prm, y, region)
    _, L0 = sim.lhs(y)
    assert sim.jacobian(y, L0, 1.0e6) == 0
    J = sim.jacobian_mat()
    nb = m.nowned
    pc1 = flow.PC(J, flow.PC_BJACOBI_ILU0, 1, np.arange(nb, dtype=np.int32))
    pc2 = flow.PC(J, flow.PC_PBJACOBI)
    r = np.random.default_rng(SEED).uniform(-1, 1, nb * 2)
    z1, z2 = np.zeros(nb * 2), np.zeros(nb * 2)
    pc1.apply(r, z1)
    pc2.apply(r, z2)
    assert relerr(z1, z2) < 1e-13
    for ksp in (flow.KSP_GMRES, flow.KSP_BCGS):
        x = np.ones(nb * 2)
        reason, its, rn = flow.ksp_solve(J, pc2, np.zeros(nb * 2), x, flow.ksp_opts(type=ksp))
        assert reason > 0 and its == 0 and not x.any()
    pc1.destroy()
    pc2.destroy()
    sim.destroy()


def test_mixed_host_and_device_pointers_and_repeated_meshes(wo, flow):
    """y on the device, lhs_last on the host; re-creating contexts of different EOS / sizes in one process"""
    import torch
    for dims, mk in (((4, 3, 2), make_problem), ((3, 3, 3), make_problem_wce), ((2, 2, 9), make_problem)):
        m, y, region, prm = mk(wo, dims=dims)
        ref = oracle_flow(wo, m, prm, y, region)
        sim = gpu_flow(wo, flow, m, prm, y, region)
        _, L0 = ref.lhs(y)
        sim.lhs(y)
        yd = torch.tensor(y * 1.0002, device="cuda")
        rd = torch.zeros_like(yd)
        torch.cuda.synchronize()
        err, _, _, _ = sim.residual(yd, L0, 1.0e5, r=rd, want_parts=False)
        _, _, _, r0 = ref.residual(y * 1.0002, L0, 1.0e5)
        assert err == 0 and relerr(rd.cpu().numpy(), r0) < 1e-10
        sim.destroy()
