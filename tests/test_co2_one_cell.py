"""The reference's one-cell CO2 benchmark (test/benchmark/ncg/co2_one_cell: co2_one_cell.json,
test_co2_one_cell.py; O'Sullivan et al. 1985, fig. 5) -- the end-to-end pin of the eos_wce path (BASELINE configs
4 and 5 run eos_wce): a 1 m^3 two-phase cell at 7.69 MPa / 260 degC with 30 bar of CO2 is produced at 5 kg/s with
Corey relative permeabilities (mobility-weighted production of both components) until it is almost dry, 38
backward-Euler steps of 0.5 s, IFC-67.  The reference compares pressure, temperature, vapour saturation and the
production enthalpy with the AUTOUGH2 listing at 1e-3; the listing's history is committed as
tests/golden/co2_one_cell.json (tools/make_golden.py).  Two identical cells instead of one (no gravity, equal
states: zero flux between them) so that the mesh has a face."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from waiwera_b200 import mesh as wmesh

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "co2_one_cell.json")))
DT, NSTEPS, RATE = 0.5, 38, -5.0
PRIMARY = [7694336.789042256, 0.2, 3000000.0]


def problem():
    m = wmesh.structured(2, 1, 1, dx=1.0, gravity=(0.0, 0.0, 0.0), heterogeneous=False)
    m.rock[:, 0:3] = 1e-15
    m.rock[:, 3:5] = 1.5
    m.rock[:, 5], m.rock[:, 6], m.rock[:, 7] = 0.15, 2500.0, 900.0
    primary = np.tile(PRIMARY, (2, 1))
    region = np.full(2, 4, np.int32)
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1)
    return m, y, region


def params(wo):
    return wo.make_params(eos=wo.EOS_WCE, thermo=wo.THERMO_IFC67, gravity=(0.0, 0.0, 0.0),
                          relperm=wo.make_relperm("corey", slr=0.3, ssr=0.05))


def production_enthalpy(fluid_record):
    """source%enthalpy of a producing source: mobility-weighted phase enthalpies (src/fluid.F90:417-436)"""
    fl = fluid_record
    phases = int(round(fl[4]))
    mob, h = [], []
    for p in range(2):
        ph = fl[8 + 9 * p: 8 + 9 * (p + 1)]
        mob.append(ph[3] * ph[0] / ph[1] if phases & (1 << p) else 0.0)
        h.append(ph[5])
    return (mob[0] * h[0] + mob[1] * h[1]) / (mob[0] + mob[1])


def check_history(hist):
    """hist: per step (P, T, Sv, Pco2, production enthalpy at the end of the step)"""
    hist = np.array(hist)
    gold = np.array(GOLD["element"])[1:]               # row 0 is the initial state
    assert len(hist) == len(gold) == NSTEPS
    for col, name in enumerate(["pressure", "temperature", "gas_saturation"]):
        err = np.abs(hist[:, col] - gold[:, col]).max() / np.abs(gold[:, col]).max()
        assert err < 1e-3, (name, err)
    # CO2 partial pressure spans 12 decades as the cell dries out: compare while it is above 1 Pa
    sel = gold[:, 3] > 1.0
    assert np.abs(hist[sel, 3] / gold[sel, 3] - 1.0).max() < 5e-3
    he = np.array(GOLD["source_enthalpy"])[1:]
    assert np.abs(hist[:, 4] - he).max() / np.abs(he).max() < 1e-3
    # measured (oracle): P 1.6e-4, T 7e-5, Sv 1.2e-5, production enthalpy 1.5e-4 of the AUTOUGH2 history


def newton_opts(wo):
    o = wo.NewtonOpts()
    o.max_iterations, o.min_iterations = 8, 0
    o.rel_tol, o.abs_tol, o.update_rel_tol, o.update_abs_tol = 1e-5, 1.0, 1e-10, 1.0
    o.fd_err, o.fd_umin, o.pc_type = 1e-8, 1e-2, wo.PC_BJACOBI_ILU0
    o.ksp.type, o.ksp.restart, o.ksp.maxit = wo.KSP_BCGS, 30, 10000
    o.ksp.rtol, o.ksp.atol, o.ksp.dtol = 1e-5, 1e-50, 1e5
    return o


def run_oracle(wo):
    m, y, region = problem()
    f = wo.Flow(params(wo), m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    f.set_sources([0, 1], [0, 0], [RATE, RATE], [0.0, 0.0])
    assert f.fluid_init(y, region) == 0
    L = wo.lib()
    J = f.bsr()
    color = np.zeros(J.contents.nb, np.int32)
    nc = L.wo_bsr_coloring(J, wo.ip(color))
    o = newton_opts(wo)
    hist, regions, ys = [], [], []
    for step in range(NSTEPS):
        err, L0 = f.lhs(y)
        assert err == 0
        L.wo_flow_pre_timestep(f.h)
        res = wo.NewtonResult()
        L.wo_newton_solve_be(f.h, J, wo.ip(color), nc, None, C.byref(o), DT, wo.dp(L0), wo.dp(y), C.byref(res))
        assert res.reason > 0, (step, res.reason)
        assert f.residual(y, L0, DT)[0] == 0
        fl = f.fluid()[0]
        hist.append((fl[0], fl[1], fl[8 + 9 + 2], fl[7], production_enthalpy(fl)))
        regions.append(int(f.regions()[0]))
        ys.append(y.copy())
    L.wo_bsr_destroy(J)
    return hist, regions, ys


@pytest.fixture(scope="module")
def oracle_run(wo):
    return run_oracle(wo)


def test_oracle_matches_autough2_co2_one_cell(oracle_run):
    hist, regions, ys = oracle_run
    check_history(hist)


@pytest.mark.gpu
def test_cuda_path_reproduces_co2_one_cell(wo, oracle_run):
    from waiwera_b200 import flow
    from util import wb_params_from_oracle
    hist_ref, regions_ref, ys_ref = oracle_run
    m, y, region = problem()
    sim = flow.FlowSimulation(wb_params_from_oracle(wo, flow, params(wo)), m)
    assert sim.set_sources([0, 1], [0, 0], [RATE, RATE], [0.0, 0.0]) == 0
    assert sim.fluid_init(y, region) == 0
    o = flow.newton_opts(max_iterations=8, rel_tol=1e-5, pc_type=flow.PC_BJACOBI_ILU0,
                         ksp=flow.ksp_opts(type=flow.KSP_BCGS))
    hist = []
    for step in range(NSTEPS):
        err, L0 = sim.lhs(y)
        assert err == 0
        sim.pre_timestep()
        res = sim.newton_solve(y, L0, DT, o)
        assert res.reason > 0
        assert sim.residual(y, L0, DT)[0] == 0
        fl = sim.fluid()[0]
        hist.append((fl[0], fl[1], fl[8 + 9 + 2], fl[7], production_enthalpy(fl)))
        assert int(sim.regions()[0]) == regions_ref[step]
        # both runs stop Newton at 1e-5 of the residual with an inexact (rtol 1e-5) Krylov solve: equal to that level
        assert np.abs(y - ys_ref[step]).max() / np.abs(ys_ref[step]).max() < 1e-4
    check_history(hist)
    sim.destroy()
