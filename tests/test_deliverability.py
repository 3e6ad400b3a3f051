"""The reference's deliverability benchmark (test/benchmark/source/deliverability: deliv_{delv,delt,delg_flow}.json,
test_deliverability.py): ten 100 m cubes in a row, eos we, IFC-67, hot liquid at 2.9 MPa / 230 degC, Dirichlet
boundary at the far end, a well on deliverability in cell 0 -- rate = -PI * sum_p mobility_p * (P - 0.5 MPa),
re-evaluated at every function evaluation (src/source_control.F90:322-507), so that it enters the finite-difference
Jacobian.  delv: PI = 1e-11 m3; delt: the same with a 20 kg/s total-flow limiter (src/source_network_node.F90:245-315);
delg_flow: PI calculated from an initial rate of 20 kg/s (:407-468); delg_pi_table: PI from a table in time, averaged over
each time step by the host (endpoint averaging) and handed to the device before the step.  80 prescribed steps to 3.26e8 s, the cell
boils.  Golden output: the AUTOUGH2 listings (tests/golden/deliverability.json); the reference accepts 5e-3 on P, T,
Sv of the last output and 1e-2 on the histories of the production cell, the generation rate and the enthalpy."""
import json
import os

import numpy as np
import pytest

from util import OracleSim, run_adaptive, we_fields, we_production_enthalpy
from waiwera_b200 import mesh as wmesh

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "deliverability.json")))
NX, DX = 10, 100.0
CASES = ["delv", "delt", "delg_flow", "delg_pi_table"]


def problem(case):
    g = GOLD[case]
    m = wmesh.structured(NX, 1, 1, dx=DX, gravity=(0.0, 0.0, 0.0), heterogeneous=False)
    m.rock[:, 0:3] = 1e-13
    m.rock[:, 3:5] = 1.5
    m.rock[:, 5], m.rock[:, 6], m.rock[:, 7] = 0.1, 2600.0, 900.0
    m = wmesh.add_boundary(m, [NX - 1], (1.0, 0.0, 0.0), 0.5 * DX, DX * DX, 1, gravity=(0.0, 0.0, 0.0))
    primary = np.tile(g["initial"], (NX, 1))
    region = np.ones(NX, np.int32)
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1)
    return m, y, region


def params(mod):
    return mod.make_params(eos=mod.EOS_WE, thermo=mod.THERMO_IFC67, gravity=(0.0, 0.0, 0.0),
                           relperm=mod.make_relperm("linear", liquid=(0.0, 1.0), vapour=(0.0, 1.0)),
                           cappress=mod.make_cappress("linear", saturation_limits=(0.0, 0.0), pressure=0.0))


def controls(case, fluid0):
    """(productivity index, reference pressure, direction, limit) of the well from the input's source value;
    delg_flow: calculate_PI_from_rate at the initial state (src/source_control.F90:407-468)"""
    s = GOLD[case]["source"][0]
    pref = s["deliverability"]["pressure"]
    if isinstance(s["deliverability"].get("productivity"), dict):
        pi = None                                      # table in time: pi_for_step
    elif "productivity" in s["deliverability"]:
        pi = s["deliverability"]["productivity"]
    else:
        phases = int(round(fluid0[4]))
        mob = sum((fluid0[7 + 8 * p + 3] * fluid0[7 + 8 * p] / fluid0[7 + 8 * p + 1]) for p in range(2) if phases & (1 << p))
        pi = abs(s["rate"]) / (mob * (fluid0[0] - pref) * fluid0[5])
    limit = s.get("limiter", {}).get("limit", 0.0)
    return pi, pref, 1, limit


def newton_opts_oracle(wo):
    o = wo.NewtonOpts()
    o.max_iterations, o.min_iterations = 8, 0
    o.rel_tol, o.abs_tol, o.update_rel_tol, o.update_abs_tol = 1e-7, 1.0, 1e-10, 1.0
    o.fd_err, o.fd_umin, o.pc_type = 1e-8, 1e-2, wo.PC_BJACOBI_ILU0
    o.ksp.type, o.ksp.restart, o.ksp.maxit = wo.KSP_BCGS, 30, 10000
    o.ksp.rtol, o.ksp.atol, o.ksp.dtol = 1e-5, 1e-50, 1e5
    return o


def pi_for_step(case, t0, t1):
    """productivity index for the step [t0, t1] from its table: interpolation_table%average with the input's
    interpolation / averaging (src/source_control.F90:486, src/interpolation.F90:565-680)"""
    from waiwera_b200 import ingest
    s = GOLD[case]["source"][0]
    tab = np.array(s["deliverability"]["productivity"]["time"], float)
    return ingest._table_average(tab, s.get("interpolation", "linear"), t0, t1, s.get("averaging", "integrate"))


def run(case, sim, y, rates, opts=None, set_controls=None):
    g = GOLD[case]
    hist = []
    t = 0.0
    for dt in g["step_sizes"]:
        dt = min(dt, g["stop"] - t)
        if set_controls is not None:
            set_controls(pi_for_step(case, t, t + dt))
        t1, _, _, _ = run_adaptive(sim, y, dt, dt, opts=opts, max_steps=1)
        assert t1 == dt                               # no step cuts: the prescribed step list is followed
        t += dt
        fl = np.asarray(sim.fluid())
        hist.append((t, we_fields(fl, NX), we_production_enthalpy(fl[0]), rates()[0]))
    return hist


def check(case, hist):
    g = GOLD[case]
    first = 1 if g["times"][0] == 0.0 else 0          # some listings have no table at t = 0
    gt = np.array(g["times"])[first:]
    t = np.array([h[0] for h in hist])
    assert np.allclose(t, gt, rtol=1e-9)
    rel = lambda a, b: np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)
    out, gold = hist[-1][1], np.array(g["tables"][-1])
    errs = {}
    for col, name in enumerate(GOLD["columns"]):
        errs["final " + name] = rel(out[:, col], gold[:, col])
        errs["history " + name] = rel([h[1][0, col] for h in hist], [tab[0][col] for tab in g["tables"][first:]])
    sfirst = 1 if g["source_times"][0] == 0.0 else 0
    errs["rate"] = rel([h[3] for h in hist], g["rate"][sfirst:])
    errs["enthalpy"] = rel([h[2] for h in hist], g["enthalpy"][sfirst:])
    for k, v in errs.items():
        assert v < 5e-5, (case, k, v)      # measured <= 9.2e-6: the listing's printed digits (reference: 5e-3 / 1e-2)
    return errs


def run_oracle(wo, case):
    m, y, region = problem(case)
    f = wo.Flow(params(wo), m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    assert f.set_boundary(int(m.boundary["ghost_cells"][0]), NX - 1, np.array(GOLD[case]["boundary"], float), 1) == 0
    f.set_sources([0], [0], [GOLD[case]["source"][0].get("rate", 0.0)], [0.0])
    assert f.fluid_init(y, region) == 0
    pi, pref, direction, limit = controls(case, f.fluid()[0])
    setc = lambda v: f.set_source_controls([0], [v], [pref], [direction], [limit])
    if pi is not None:
        setc(pi)
    sim = OracleSim(wo, f, newton_opts_oracle(wo))
    hist = run(case, sim, y, lambda: f.source_rates(1), set_controls=setc if pi is None else None)
    sim.destroy()
    return hist, y


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_autough2_deliverability(wo, case):
    hist, y = run_oracle(wo, case)
    check(case, hist)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_path_reproduces_deliverability(wo, case):
    from waiwera_b200 import flow
    hist_ref, y_ref = run_oracle(wo, case)
    m, y, region = problem(case)
    sim = flow.FlowSimulation(params(flow), m)
    assert sim.set_boundaries(m.boundary["ghost_cells"], m.boundary["interior_cells"], np.array([GOLD[case]["boundary"]], float),
                              np.array([1], np.int32)) == 0
    assert sim.set_sources([0], [0], [GOLD[case]["source"][0].get("rate", 0.0)], [0.0]) == 0
    assert sim.fluid_init(y, region) == 0
    pi, pref, direction, limit = controls(case, sim.fluid()[0])
    setc = lambda v: sim.set_source_controls([0], [v], [pref], [direction], [limit])
    if pi is not None:
        assert setc(pi) == 0
    o = flow.newton_opts(max_iterations=8, rel_tol=1e-7, pc_type=flow.PC_BJACOBI_ILU0, ksp=flow.ksp_opts(type=flow.KSP_BCGS))
    hist = run(case, sim, y, lambda: sim.source_rates(), opts=o, set_controls=setc if pi is None else None)
    check(case, hist)
    assert np.abs(y - y_ref).max() / np.abs(y_ref).max() < 1e-5
    assert abs(hist[-1][3] - hist_ref[-1][3]) < 1e-5 * abs(hist_ref[-1][3])
    sim.destroy()
