"""MINC dual-porosity meshes (BASELINE config 5 / SURVEY.md section 8d-5): the generator's geometry against the
reference's known answers (test/unit/src/minc_test.F90:210-394), the array contract of the MINC cells and faces
(src/mesh.F90:3120-3160, 2286-2380), and GPU parity of the irregular-sparsity path (rows of 2 and 8 blocks)."""
import ctypes as C

import numpy as np
import pytest

from waiwera_b200 import mesh as wmesh
from util import SEED, psat_fn, oracle_flow, gpu_flow, relerr


def test_minc_geometry_known_answers():
    """minc_test.F90:273-394 test_geometry (root-finder tolerance 1e-8 in the reference)"""
    cases = [
        ([0.1, 0.9], [50.], [0.1, 0.9], [0.036], [0., 25. / 3.]),
        ([0.1, 0.3, 0.6], [100.], [0.1, 0.3, 0.6], [0.018, 0.018], [0., 25. / 3., 100. / 9.]),
        ([10, 20, 30, 40], [100.], [0.1, 0.2, 0.3, 0.4], [0.018, 0.018, 0.018], [0., 50. / 9., 25. / 3., 7400. / 999.]),
        ([5, 20, 30, 45], [100., 100.], [0.05, 0.2, 0.3, 0.45], [0.038, 0.033763886490617179, 0.026153393818234151],
         [0., 2.78691708403391, 5.006902875674109, 8.6030900201459914]),
        ([5, 20, 30, 45], [100., 80.], [0.05, 0.2, 0.3, 0.45], [0.04275, 0.038046846447656414, 0.029623680667175543],
         [0., 2.4753441451477878, 4.433244588928682, 7.595274770950577]),
        ([10, 30, 60], [100., 80., 90.], [0.10, 0.3, 0.6], [0.0605, 0.046229920797811137],
         [0., 2.8192309717077664, 7.7871646178561607]),
    ]
    for vols, sp, ev, ea, ed in cases:
        v, a, d = wmesh.minc_geometry(vols, sp)
        assert np.allclose(v, ev, rtol=1e-14)
        assert np.allclose(a, ea, rtol=1e-7)
        assert np.allclose(d, ed, rtol=1e-7, atol=1e-12)


def test_minc_inner_connection_distance_known_answers():
    """minc_test.F90:210-269: a single level's outer distance is inner_connection_distance(0)"""
    assert np.isclose(wmesh.minc_geometry([0.1, 0.9], [50.])[2][1], 25. / 3., rtol=1e-14)
    assert np.isclose(wmesh.minc_geometry([0.1, 0.9], [50., 80.])[2][1], 100. / 13., rtol=1e-14)
    assert np.isclose(wmesh.minc_geometry([0.1, 0.9], [50., 80., 60.])[2][1], 360. / 59., rtol=1e-14)


def test_minc_mesh_contract():
    """numbering, volumes and face records of the MINC cells (mesh.F90:2286-2380, 3120-3160)"""
    base = wmesh.structured(5, 4, 3, dx=10.0)
    n = base.ninterior
    m = wmesh.add_minc(base, volumes=(0.1, 0.3, 0.6), spacing=(50., 50., 50.))
    v, a, d = wmesh.minc_geometry((0.1, 0.3, 0.6), (50., 50., 50.))
    assert m.ncell == m.nowned == 3 * n and m.nface == base.nface + 2 * n
    # original faces first and untouched; one face per MINC cell, support (level m-1, level m)
    assert np.array_equal(m.face_cells[:base.nface], base.face_cells)
    assert np.array_equal(m.face_cells[base.nface:base.nface + n], np.stack([np.arange(n), np.arange(n) + n], 1))
    assert np.array_equal(m.face_cells[base.nface + n:], np.stack([np.arange(n) + n, np.arange(n) + 2 * n], 1))
    V = base.cell_geom[:, 3]
    for lvl in range(3):
        assert np.allclose(m.cell_geom[lvl * n:(lvl + 1) * n, 3], V * v[lvl], rtol=1e-15)
        assert np.array_equal(m.cell_geom[lvl * n:(lvl + 1) * n, :3], base.cell_geom[:, :3])
    f1 = m.face_geom[base.nface:base.nface + n]
    assert np.allclose(f1[:, 0], V * a[0]) and np.allclose(f1[:, 1], d[0]) and np.allclose(f1[:, 2], d[1])
    assert np.allclose(f1[:, 3], d[0] + d[1]) and (f1[:, 4:8] == 0).all() and (f1[:, 11] == 1).all()
    # total pore volume is conserved
    assert np.isclose(m.cell_geom[:, 3].sum(), V.sum(), rtol=1e-14)
    # owner of a matrix cell = owner of its fracture cell; sub-domain likewise
    own = wmesh.minc_owner(m, (1, 2, 2))
    assert np.array_equal(own[:n], own[n:2 * n]) and np.array_equal(own[:n], own[2 * n:])
    blk = wmesh.minc_cube_blocks(m, 2)
    assert np.array_equal(blk[:n], blk[n:2 * n])
    pm = wmesh.partition(m, own, 1, 4)
    nat = pm.natural[:pm.nowned]
    assert set(nat[nat >= n] % n) <= set(nat[nat < n])


def minc_problem(wo, eos, dims=(6, 5, 4), volumes=(0.1, 0.3, 0.6)):
    base = wmesh.structured(*dims, dx=10.0, seed=SEED)
    m = wmesh.add_minc(base, volumes=volumes, spacing=(50., 50., 50.), matrix_permeability_factor=0.01)
    n = base.ninterior
    nlev = len(volumes) - 1
    if eos == "we":
        primary, region = wmesh.hydrostatic_state(base, seed=SEED, two_phase_layers=1, thermo_psat=psat_fn(wo, 0))
        prm = wo.make_params(eos=wo.EOS_WE)
    else:
        primary, region = wmesh.wce_state(base, seed=SEED, two_phase_layers=1, thermo_psat=psat_fn(wo, 0))
        prm = wo.make_params(eos=wo.EOS_WCE)
    rng = np.random.default_rng(SEED + 7)
    prims, regs = [primary], [region]
    for lvl in range(nlev):  # matrix cells slightly out of equilibrium with their fracture cell
        p = primary.copy()
        p[:, 0] *= 1.0 + 1e-3 * rng.uniform(-1, 1, n)
        prims.append(p)
        regs.append(region)
    primary, region = np.concatenate(prims), np.concatenate(regs)
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1)
    return m, y, region, prm


@pytest.mark.gpu
@pytest.mark.parametrize("eos", ["we", "wce"])
def test_minc_residual_jacobian_pc_match_oracle(wo, eos):
    from waiwera_b200 import flow
    m, y, region, prm = minc_problem(wo, eos)
    ref = oracle_flow(wo, m, prm, y, region)
    sim = gpu_flow(wo, flow, m, prm, y, region)
    bs = sim.np
    e0, L0 = ref.lhs(y)
    e1, L1 = sim.lhs(y)
    assert e0 == e1 == 0 and relerr(L1, L0) < 1e-12
    rng = np.random.default_rng(SEED + 3)
    y2 = y * (1 + 1e-4 * rng.uniform(-1, 1, len(y)))
    dt = 1.0e5
    e0, _, rhs0, r0 = ref.residual(y2, L0, dt)
    e1, _, rhs1, r1 = sim.residual(y2, L0, dt)
    assert e0 == e1 == 0 and relerr(rhs1, rhs0) < 1e-10 and relerr(r1, r0) < 1e-10
    # irregular pattern: fracture rows have up to 8 blocks, innermost matrix rows 2
    A = ref.bsr()
    color = np.zeros(A.contents.nb, np.int32)
    nc = wo.lib().wo_bsr_coloring(A, wo.ip(color))
    assert wo.lib().wo_fd_jacobian(ref.h, wo.dp(y2), wo.dp(L0), dt, wo.dp(r0), wo.ip(color), nc, 1e-8, 1e-2, A) == 0
    rowptr, colidx, val = [a.copy() for a in wo.bsr_arrays(A)]
    nb, bsg, rp_g, ci_g = sim.jacobian_pattern()
    assert bsg == bs and np.array_equal(rp_g, rowptr) and np.array_equal(ci_g, colidx)
    nnz = np.diff(rowptr)
    assert nnz.max() == 8 and nnz.min() == 2
    assert sim.jacobian(y2, L0, dt) == 0
    Jl = sim.jacobian_values()
    rows = np.repeat(np.arange(nb), nnz)
    v3 = np.abs(val).reshape(-1, bs, bs)
    rowmax = np.zeros((nb, bs))
    for ii in range(bs):
        np.maximum.at(rowmax[:, ii], rows, v3[:, :, ii].max(axis=1))
    scale = np.tile(rowmax[rows], (1, bs))
    assert (np.abs(Jl - val) / np.maximum(scale, 1e-300)).max() < 1e-5
    # SpMV + ILU(0) over fracture-cube sub-domains that carry their matrix cells (8-block rows: generic level path)
    J = sim.jacobian_mat()
    x = rng.uniform(-1, 1, nb * bs)
    Aval = np.ctypeslib.as_array(A.contents.val, shape=(len(colidx), bs * bs))
    Aval[:] = Jl                     # same values on both sides for the SpMV / PC comparison
    ax, ref_ax = np.zeros(nb * bs), np.zeros(nb * bs)
    J.mult(x, ax)
    wo.lib().wo_bsr_spmv(A, wo.dp(x), wo.dp(ref_ax))
    assert relerr(ax, ref_ax) < 1e-14
    bor = wmesh.minc_cube_blocks(m, 3)
    pc_ref = wo.lib().wo_pc_create(A, 2, wo.ip(bor))
    pc = flow.PC(J, flow.PC_BJACOBI_ILU0, 1, bor)
    z0, z1 = np.zeros(nb * bs), np.zeros(nb * bs)
    wo.lib().wo_pc_apply(pc_ref, wo.dp(x), wo.dp(z0))
    pc.apply(x, z1)
    assert relerr(z1, z0) < 1e-11
    wo.lib().wo_pc_destroy(pc_ref)
    pc.destroy()
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()


@pytest.mark.gpu
def test_full_size_minc_properties(wo):
    """BASELINE config 5 size on one GPU: 100^3 fracture cells + one MINC level = 2 M cells (eos_we, 9.94 M blocks,
    rows of 2 and 8 blocks): conservation of the inflows including the fracture-matrix exchange, determinism,
    pattern counts, SpMV linearity and an ILU(0) sub-domain apply that is an exact inverse on its own product."""
    from waiwera_b200 import flow
    base = wmesh.structured(100, 100, 100, dx=10.0, seed=SEED)
    m = wmesh.add_minc(base, volumes=(0.1, 0.9), spacing=(50., 50., 50.), matrix_permeability_factor=0.01)
    n = base.ninterior
    primary, region = wmesh.hydrostatic_state(base, seed=SEED)
    rng = np.random.default_rng(SEED + 9)
    pm = primary.copy()
    pm[:, 0] *= 1.0 + 1e-3 * rng.uniform(-1, 1, n)
    y = np.ascontiguousarray(wmesh.scale_primaries(np.concatenate([primary, pm]), np.concatenate([region, region]))).reshape(-1)
    region2 = np.concatenate([region, region])
    sim = flow.FlowSimulation(flow.make_params(), m)
    assert sim.fluid_init(y, region2) == 0
    e, L = sim.lhs(y)
    assert e == 0
    e, lhs, rhs, r = sim.residual(y, L, 1.0e6)
    assert e == 0
    vol = m.cell_geom[:, 3]
    for k in range(2):
        tot = np.sum(vol * rhs[k::2])
        assert abs(tot) <= 1e-9 * np.sum(vol * np.abs(rhs[k::2]))
    assert np.array_equal(sim.residual(y, L, 1.0e6)[3], r)
    nb, bs, rowptr, colidx = sim.jacobian_pattern()
    assert nb == 2 * n and len(colidx) == 6940000 + 3 * n   # 6.94 M stencil blocks + matrix diag + 2 couplings per pair
    nnz = np.diff(rowptr)
    assert nnz.max() == 8 and nnz.min() == 2
    assert sim.jacobian(y, L, 1.0e6) == 0
    J = sim.jacobian_mat()
    x1, x2 = rng.uniform(-1, 1, nb * 2), rng.uniform(-1, 1, nb * 2)
    a1, a2, a12 = np.zeros(nb * 2), np.zeros(nb * 2), np.zeros(nb * 2)
    J.mult(x1, a1)
    J.mult(x2, a2)
    J.mult(2.0 * x1 + 0.5 * x2, a12)
    assert relerr(a12, 2.0 * a1 + 0.5 * a2) < 1e-13
    sim.destroy()
