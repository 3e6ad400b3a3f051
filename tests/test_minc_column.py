"""The reference's MINC column benchmark (test/benchmark/minc/column: minc_column_{single,minc}.json,
test_minc_column.py): an 11-layer, 1 km column (100 m x 100 m), eos we, IFC-67, two-phase on top of hot liquid,
atmospheric Dirichlet boundary, 10 kg/s injected at the bottom and 25 kg/s produced from layer 6 for 90 days with
the adaptive backward-Euler stepper.  The fractures around the well dry out (regions 4 -> 2) while the matrix stays
two-phase.  Run as a single-porosity model and with MINC in layers 3-8 (fracture volume fraction 0.1, two matrix
levels 0.3 / 0.6, three fracture planes at 5 m spacing, matrix permeability 1e-18 m2: BASELINE config 5's mesh
type).  Golden output: the AUTOUGH2 listings shipped with the benchmark (tests/golden/minc_column.json); the
reference accepts 2.5e-2 on P, T, Sv of the last output, 2e-2 on their history in the production cell and 1e-2 on
the production enthalpy history."""
import json
import os

import numpy as np
import pytest

from util import OracleSim, run_adaptive
from waiwera_b200 import mesh as wmesh

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "minc_column.json")))
NL, AREA = 11, 1.0e4
BOTTOM = np.array([40.0, 90.0, 150.0, 220.0, 300.0, 390.0, 490.0, 600.0, 720.0, 850.0, 1000.0])
THICK = np.diff(np.concatenate([[0.0], BOTTOM]))
GRAVITY = (0.0, 0.0, -9.8)
BOUNDARY = [100000.022163, 20.0000000004]
ZONE = np.arange(2, 8)                 # layers with centres in z = [-600, -100]
PROD_CELL, INJ_CELL = 5, 10
T_STOP, DT0, DT_MAX = 7776000.0, 43200.0, 259200.0


def column_mesh(minc):
    z = -(np.cumsum(THICK) - 0.5 * THICK)
    cell_geom = np.zeros((NL, 4))
    cell_geom[:, 0:2] = 50.0
    cell_geom[:, 2] = z
    cell_geom[:, 3] = AREA * THICK
    k = np.arange(NL - 1)
    fg = np.zeros((NL - 1, 12))
    fg[:, 0] = AREA
    fg[:, 1], fg[:, 2] = 0.5 * THICK[:-1], 0.5 * THICK[1:]
    fg[:, 3] = fg[:, 1] + fg[:, 2]
    fg[:, 4:7] = (0.0, 0.0, -1.0)
    fg[:, 7] = float(np.dot(GRAVITY, (0.0, 0.0, -1.0)))
    fg[:, 8:10] = 50.0
    fg[:, 10] = -np.cumsum(THICK)[:-1]
    fg[:, 11] = 3
    rock = np.zeros((NL, 8))
    rock[:, 0:3] = 1e-13
    rock[:, 3:5] = 1.5
    rock[:, 5], rock[:, 6], rock[:, 7] = 0.1, 2600.0, 900.0
    m = wmesh.Mesh(ncell=NL, ninterior=NL, nowned=NL, face_cells=np.ascontiguousarray(np.stack([k, k + 1], 1).astype(np.int32)),
                   face_geom=np.ascontiguousarray(fg), cell_geom=np.ascontiguousarray(cell_geom),
                   rock=np.ascontiguousarray(rock), dims=(1, 1, NL), natural=np.arange(NL, dtype=np.int64), ncell_global=NL)
    if minc:
        m.rock[ZONE, 5] = 0.5          # rock type "fract"
        matrix = [1e-18, 1e-18, 1e-18, 1.5, 1.5, 0.0555555555556, 2600.0, 900.0]
        m = wmesh.add_minc(m, volumes=(0.1, 0.3, 0.6), spacing=(5.0, 5.0, 5.0), cells=ZONE, matrix_rock=matrix)
    return wmesh.add_boundary(m, [0], (0.0, 0.0, 1.0), 0.5 * THICK[0], AREA, 3, gravity=GRAVITY)


def problem(case):
    minc = case == "minc"
    m = column_mesh(minc)
    primary = np.array(GOLD[case]["initial_primary"])
    region = np.array(GOLD[case]["initial_region"], np.int32)
    if minc:            # matrix cells start from their fracture cell's state
        primary = np.concatenate([primary, primary[ZONE], primary[ZONE]])
        region = np.concatenate([region, region[ZONE], region[ZONE]]).astype(np.int32)
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1)
    return m, y, region


def params(wo):
    return wo.make_params(eos=wo.EOS_WE, thermo=wo.THERMO_IFC67, gravity=GRAVITY,
                          relperm=wo.make_relperm("linear", liquid=(0.0, 1.0), vapour=(0.0, 1.0)),
                          cappress=wo.make_cappress("linear", saturation_limits=(0.0, 0.0), pressure=0.0))


def newton_opts(wo):
    o = wo.NewtonOpts()
    o.max_iterations, o.min_iterations = 8, 0
    o.rel_tol, o.abs_tol, o.update_rel_tol, o.update_abs_tol = 1e-5, 1.0, 1e-10, 1.0
    o.fd_err, o.fd_umin, o.pc_type = 1e-8, 1e-2, wo.PC_BJACOBI_ILU0
    o.ksp.type, o.ksp.restart, o.ksp.maxit = wo.KSP_BCGS, 30, 10000
    o.ksp.rtol, o.ksp.atol, o.ksp.dtol = 1e-5, 1e-50, 1e5
    return o


SOURCES = dict(cells=[INJ_CELL, PROD_CELL], components=[1, 0], rates=[10.0, -25.0], enthalpies=[1037600.46717, 0.0])


def fields(fluid, n):
    fl = np.asarray(fluid)[:n]
    return np.stack([fl[:, 0], fl[:, 1], fl[:, 7 + 8 + 2]], 1)          # P, T, vapour saturation


def production_enthalpy(fl):
    phases = int(round(fl[4]))
    mob = [(fl[7 + 8 * p + 3] * fl[7 + 8 * p] / fl[7 + 8 * p + 1]) if phases & (1 << p) else 0.0 for p in range(2)]
    return (mob[0] * fl[7 + 5] + mob[1] * fl[7 + 8 + 5]) / (mob[0] + mob[1])


def check(case, hist):
    """hist: [(time, fields of all cells, production enthalpy)] per step"""
    g = GOLD[case]
    t = np.array([h[0] for h in hist])
    assert abs(t[-1] - T_STOP) < 1e-6 * T_STOP
    out, gold = hist[-1][1], np.array(g["tables"][-1])
    assert out.shape == gold.shape
    for col, name in enumerate(GOLD["columns"]):
        err = np.linalg.norm(out[:, col] - gold[:, col]) / np.linalg.norm(gold[:, col])
        assert err < 2.5e-2, (case, name, err)
    # history in the production cell and production enthalpy, interpolated to the AUTOUGH2 output times
    gt = np.array(g["times"])[1:]
    for col, name in enumerate(GOLD["columns"]):
        mine = np.interp(gt, t, [h[1][PROD_CELL, col] for h in hist])
        ref = np.array([tab[PROD_CELL][col] for tab in g["tables"]])[1:]
        assert np.linalg.norm(mine - ref) / np.linalg.norm(ref) < 2e-2, (case, name)
    he = np.interp(np.array(g["source_times"])[1:], t, [h[2] for h in hist])
    ref = np.array(g["production_enthalpy"])[1:]
    assert np.linalg.norm(he - ref) / np.linalg.norm(ref) < 1e-2


def run(sim, m, y, opts=None):
    hist = []
    n = m.ninterior
    t, dt = 0.0, DT0
    while t < T_STOP * (1 - 1e-12):
        t1, nsteps, nits, nretry = run_adaptive(sim, y, dt, min(dt, T_STOP - t), opts=opts, max_steps=1, reduction=0.25,
                                                amplification=2.0, its_min=6, its_max=8)
        its = nits
        t += t1
        fl = sim.fluid()
        hist.append((t, fields(fl, n), production_enthalpy(np.asarray(fl)[PROD_CELL])))
        dt = t1                                        # the step that was accepted (after any reductions)
        if its < 6:
            dt = min(2.0 * dt, DT_MAX)
    return hist


def run_oracle(wo, case):
    m, y, region = problem(case)
    f = wo.Flow(params(wo), m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    assert f.set_boundary(int(m.boundary["ghost_cells"][0]), 0, np.array(BOUNDARY), 1) == 0
    f.set_sources(SOURCES["cells"], SOURCES["components"], SOURCES["rates"], SOURCES["enthalpies"])
    assert f.fluid_init(y, region) == 0
    sim = OracleSim(wo, f, newton_opts(wo))
    hist = run(sim, m, y)
    regions = f.regions()[:m.ninterior].copy()
    sim.destroy()
    return hist, regions, y


@pytest.mark.parametrize("case", ["single", "minc"])
def test_oracle_matches_autough2_minc_column(wo, case):
    hist, regions, y = run_oracle(wo, case)
    check(case, hist)
    if case == "minc":
        assert set(regions.tolist()) >= {1, 2, 4}      # liquid, dry steam (fractures at the well) and two-phase cells


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["single", "minc"])
def test_cuda_path_reproduces_minc_column(wo, case):
    from waiwera_b200 import flow
    from util import wb_params_from_oracle
    hist_ref, regions_ref, y_ref = run_oracle(wo, case)
    m, y, region = problem(case)
    sim = flow.FlowSimulation(wb_params_from_oracle(wo, flow, params(wo)), m)
    assert sim.set_boundaries(m.boundary["ghost_cells"], m.boundary["interior_cells"], np.array([BOUNDARY]),
                              np.array([1], np.int32)) == 0
    assert sim.set_sources(SOURCES["cells"], SOURCES["components"], SOURCES["rates"], SOURCES["enthalpies"]) == 0
    assert sim.fluid_init(y, region) == 0
    o = flow.newton_opts(max_iterations=8, rel_tol=1e-5, pc_type=flow.PC_BJACOBI_ILU0,
                         ksp=flow.ksp_opts(type=flow.KSP_BCGS))
    hist = run(sim, m, y, opts=o)
    check(case, hist)
    assert len(hist) == len(hist_ref)                  # same step-size history
    assert np.array_equal(sim.regions()[:m.ninterior], regions_ref)
    assert np.abs(hist[-1][1] - hist_ref[-1][1]).max(axis=0).tolist() < [100.0, 1e-2, 1e-4]
    sim.destroy()
