"""BASELINE config 1 (plumbing): Model Intercomparison Study problem 1, the radial Avdonin problem, transcribed
from test/benchmark/model_intercomparison_study/problem1/run/problem1.json -- 40-cell 1-D radial mesh (Pappus
volumes, src/mesh.F90:340-432), eos we, IFC-67, P = 5 MPa, T = 170 degC, injection of 10 kg/s at h = 678 052.78 J/kg
in cell 0, Dirichlet ghost cell at r = 1000 m, backward Euler with the step sizes of the input file, nonlinear
relative tolerance 1e-6.  The reference's own benchmark test (test_problem1.py) compares with the analytical
(Avdonin) solution digitised in data/*.dat; the same curves are committed under tests/golden/ and the oracle's
whole Newton / time-stepping path is checked against them.  The CUDA path then has to reproduce the oracle's run."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from waiwera_b200 import mesh as wmesh

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "mis_problem1.json")))

STEP_SIZES = [100000.0, 150000.0, 225000.0, 337500.0, 506250.0, 759380.0, 1139100.0, 1708600.0, 2562900.0, 3844300.0,
              5766500.0, 8649800.0, 12975000.0, 16700000.0]
T_STOP = 1.0e9
NR, DR, THICK = 40, 25.0, 100.0


def steps():
    t, out, k = 0.0, [], 0
    while t < T_STOP * (1 - 1e-12):
        dt = STEP_SIZES[min(k, len(STEP_SIZES) - 1)]
        dt = min(dt, T_STOP - t)
        out.append(dt)
        t += dt
        k += 1
    return out


def problem():
    m = wmesh.radial_1d(NR, DR, THICK)
    m.rock[:, 0:3] = 1e-12
    m.rock[:, 3:5] = 20.0
    m.rock[:, 5], m.rock[:, 6], m.rock[:, 7] = 0.2, 2500.0, 1000.0
    primary = np.tile([5.0e6, 170.0], (NR, 1))
    region = np.ones(NR, np.int32)
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1)
    return m, y, region


def params(wo):
    # degenerate linear relative permeability limits [0, 0] as in the input (kr_l(1) = 1, kr_v(0) = 0 by clamping)
    return wo.make_params(eos=wo.EOS_WE, thermo=wo.THERMO_IFC67, gravity=(0.0, -9.8, 0.0),
                          relperm=wo.make_relperm("linear", liquid=(0.0, 0.0), vapour=(0.0, 0.0)))


def run_oracle(wo, record_cell=1):
    m, y, region = problem()
    prm = params(wo)
    f = wo.Flow(prm, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    assert f.set_boundary(int(m.boundary["ghost_cells"][0]), int(m.boundary["interior_cells"][0]), np.array([5.0e6, 170.0]), 1) == 0
    f.set_sources([0], [1], [10.0], [678052.7777224329])
    assert f.fluid_init(y, region) == 0
    L = wo.lib()
    A = f.bsr()
    color = np.zeros(A.contents.nb, np.int32)
    nc = L.wo_bsr_coloring(A, wo.ip(color))
    o = wo.NewtonOpts()
    o.max_iterations, o.min_iterations = 8, 0
    o.rel_tol, o.abs_tol, o.update_rel_tol, o.update_abs_tol = 1e-6, 1.0, 1e-10, 1.0
    o.fd_err, o.fd_umin, o.pc_type = 1e-8, 1e-2, wo.PC_BJACOBI_ILU0
    o.ksp.type, o.ksp.restart, o.ksp.maxit = wo.KSP_BCGS, 30, 10000
    o.ksp.rtol, o.ksp.atol, o.ksp.dtol = 1e-5, 1e-50, 1e5
    t, hist, its = 0.0, [], 0
    for dt in steps():
        err, L0 = f.lhs(y)
        assert err == 0
        L.wo_flow_pre_timestep(f.h)
        res = wo.NewtonResult()
        L.wo_newton_solve_be(f.h, A, wo.ip(color), nc, None, C.byref(o), dt, wo.dp(L0), wo.dp(y), C.byref(res))
        assert res.reason > 0, (t, dt, res.reason)
        its += res.iterations
        t += dt
        hist.append((t, y[2 * record_cell + 1] * 100.0))
    L.wo_bsr_destroy(A)
    return m, y.copy(), np.array(hist), its


@pytest.fixture(scope="module")
def oracle_run(wo):
    return run_oracle(wo)


def test_oracle_reproduces_avdonin_solution(oracle_run):
    """test_problem1.py:85-140: temperature profile at t = 1e9 s and history at r = 37.5 m against the analytical
    solution (the reference accepts 2e-2 relative on the digitised curves)"""
    m, y, hist, its = oracle_run
    T = y[1::2] * 100.0
    P = y[0::2] * 1.0e6
    rc = m.cell_geom[:NR, 0]
    r_a, T_a = np.array(GOLD["temperature_r_analytical"]).T
    sel = (r_a >= rc[0]) & (r_a <= rc[-1])
    Ti = np.interp(r_a[sel], rc, T)
    assert np.abs(Ti - T_a[sel]).max() / T_a[sel].max() < 2e-2   # 0.43 K on the 10 K front: 25 m cells, upwind + BE
    t_a, Th_a = np.array(GOLD["temperature_time_analytical"]).T
    Th = np.interp(t_a, hist[:, 0], hist[:, 1])
    assert np.abs(Th - Th_a).max() / Th_a.max() < 2e-2
    # physical sanity of the end state: cold front at the well, undisturbed far field, pressure drives flow outwards
    assert 159.9 < T[0] < 160.6 and abs(T[-1] - 170.0) < 0.2
    assert (np.diff(P) < 0).all() and abs(P[-1] - 5.0e6) < 2e4
    assert its < 6 * len(steps())


def test_oracle_matches_autough2_listing(oracle_run):
    """test_problem1.py:77-83: "AUTOUGH2 t = 1.e9 s" FieldWithinTolTC(Temperature, tolerance 1e-4) and the pressure at
    1e-3 -- the reference's own end-to-end golden output for this path (AUTOUGH2 run of the same input, 71 steps)"""
    m, y, hist, its = oracle_run
    assert len(steps()) == 71
    T = y[1::2] * 100.0
    P = y[0::2] * 1.0e6
    Tg, Pg = np.array(GOLD["autough2_final"]["temperature"]), np.array(GOLD["autough2_final"]["pressure"])
    assert np.abs(T - Tg).max() / np.abs(Tg).max() < 1e-4, np.abs(T - Tg).max()
    assert np.abs(P - Pg).max() / np.abs(Pg).max() < 1e-3, np.abs(P - Pg).max()


@pytest.mark.gpu
def test_cuda_path_reproduces_oracle_run(wo, oracle_run):
    """the same 70+ backward-Euler steps through wb_set_sources / wb_set_boundaries / wb_newton_solve_be"""
    from waiwera_b200 import flow
    from util import wb_params_from_oracle
    m, y_ref, hist_ref, its_ref = oracle_run
    _, y, region = problem()
    sim = flow.FlowSimulation(wb_params_from_oracle(wo, flow, params(wo)), m)
    assert sim.set_boundaries(m.boundary["ghost_cells"], m.boundary["interior_cells"], np.array([[5.0e6, 170.0]]), np.array([1], np.int32)) == 0
    assert sim.set_sources([0], [1], [10.0], [678052.7777224329]) == 0
    assert sim.fluid_init(y, region) == 0
    o = flow.newton_opts(max_iterations=8, rel_tol=1e-6, pc_type=flow.PC_BJACOBI_ILU0,
                         ksp=flow.ksp_opts(type=flow.KSP_BCGS))
    its = 0
    hist = []
    t = 0.0
    for dt in steps():
        err, L0 = sim.lhs(y)
        assert err == 0
        sim.pre_timestep()
        res = sim.newton_solve(y, L0, dt, o)
        assert res.reason > 0
        its += res.iterations
        t += dt
        hist.append(y[3] * 100.0)
    assert abs(its - its_ref) <= 3
    assert np.abs(y - y_ref).max() / np.abs(y_ref).max() < 1e-7
    assert np.abs(np.array(hist) - hist_ref[:, 1]).max() < 1e-5
    sim.destroy()
