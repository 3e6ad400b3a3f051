"""Geothermal Model Intercomparison Study problems 2 (radial flow to a well: a single-phase, b two-phase, c flashing
front), 4 (1-D vertical two-phase column with drainage, 40 years), 5 (2-D areal production, a without and b with
later re-injection through a rate table) and 6 (3-D 5 x 5 x 5 field with two rock types, atmosphere / side / bottom
boundaries and a stepped production rate, adaptive steps) -- test/benchmark/model_intercomparison_study/problem{2,4,5,6}
-- run FROM THE REFERENCE'S OWN INPUT FILES (JSON + gmsh; fixtures under tests/golden/inputs/ made by
tools/make_golden.py::convert_input, read by waiwera_b200.ingest; problem 6's mesh is rebuilt from its MULgraph geometry
file, the shipped ExodusII file being HDF5-based)
through the oracle's Newton / time-stepping path, against the AUTOUGH2 listings shipped with them
(tests/golden/mis_problems.json: P, T, Sv of every cell at 12 output times, full histories of the production cell
and three more, production enthalpy history).  The reference accepts 2e-3 (problem 4), 1e-3..1e-2 (problems 2, 5) on
these fields; the CUDA path then has to reproduce the oracle's runs."""
import json
import os

import numpy as np
import pytest

from util import OracleSim, run_input
from waiwera_b200 import ingest

HERE = os.path.dirname(os.path.abspath(__file__))
INP = os.path.join(HERE, "golden", "inputs")
GOLD = json.load(open(os.path.join(HERE, "golden", "mis_problems.json")))
CASES = ["problem2a", "problem2b", "problem2c", "problem4", "problem5a", "problem5b", "problem6"]
# relative L2 error accepted over tables / histories: pressure, temperature, vapour saturation, production enthalpy.
# Measured with the oracle: 2a, 2b, 5a, 5b agree with the listings to their printed digits (P, T 1e-6..9e-6, Sv
# <= 6e-5; the reference accepts 1e-4 / 1e-3); problem 4 to 3e-4 / 9e-5 / 1e-3 (2e-3); in 2c AUTOUGH2 left the prescribed
# step list after t = 5828 s (its own step cuts at the flashing front), until then the runs agree to the printed digits,
# afterwards to 2e-3 / 9e-4 / 1e-2 (the reference accepts 1e-2).  Problem 6: the reference compares only the production
# enthalpy history with AUTOUGH2 (2e-2; here 6e-3) and the well's P / Sv with digitised curves (1.5e-2 .. 7.5e-2); here
# the production cell's pressure follows the listing to 2e-5 until it boils, the top layer drifts to 1.3e-2 below it
# (from the first step on and uniformly over the layer: the atmosphere connection, where AUTOUGH2's interface density
# differs), and the adaptive step sequences part after the cell dries out (141 steps against 145).
TIGHT = (5e-5, 5e-5, 3e-4, 5e-5)
TOL = {"problem2a": TIGHT, "problem2b": TIGHT, "problem2c": (3e-3, 1e-3, 1.5e-2, 3e-3),
       "problem4": (2e-3, 2e-3, 2e-3, 2e-3), "problem5a": TIGHT, "problem5b": TIGHT,
       "problem6": (4e-3, 6e-3, 2e-2, 1e-2)}


def newton_opts(mod, p):
    nl = p.time["step"]["solver"]["nonlinear"]
    tol = nl["tolerance"]["function"]
    if hasattr(mod, "newton_opts"):
        return mod.newton_opts(max_iterations=nl["maximum"]["iterations"], rel_tol=tol["relative"] or 1e-5,
                               abs_tol=tol["absolute"] or 1.0, pc_type=mod.PC_BJACOBI_ILU0, ksp=mod.ksp_opts(type=mod.KSP_BCGS))
    o = mod.NewtonOpts()
    o.max_iterations, o.min_iterations = nl["maximum"]["iterations"], 0
    o.rel_tol, o.abs_tol, o.update_rel_tol, o.update_abs_tol = tol["relative"] or 1e-5, tol["absolute"] or 1.0, 1e-10, 1.0
    o.fd_err, o.fd_umin, o.pc_type = 1e-8, 1e-2, mod.PC_BJACOBI_ILU0
    o.ksp.type, o.ksp.restart, o.ksp.maxit = mod.KSP_BCGS, 30, 10000
    o.ksp.rtol, o.ksp.atol, o.ksp.dtol = 1e-5, 1e-50, 1e5
    return o


def run_oracle(wo, case):
    p = ingest.load(os.path.join(INP, case + ".input.json"), mod=wo)
    m = p.mesh
    f = wo.Flow(p.params, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    for k in range(len(p.boundary_region)):
        assert f.set_boundary(int(m.boundary["ghost_cells"][k]), int(m.boundary["interior_cells"][k]),
                              p.boundary_primary[k], int(p.boundary_region[k])) == 0
    f.set_sources(p.source_cells, p.source_components, p.source_rates, p.source_enthalpies)
    assert f.fluid_init(p.y, p.region) == 0
    sim = OracleSim(wo, f, newton_opts(wo, p))
    hist, y = run_input(p, sim)
    sim.destroy()
    return p, hist, y


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def compare(case, hist):
    """errors of the run against the golden listing: [P, T, Sv] over the tables, over the cell histories, enthalpy"""
    g = GOLD[case]
    t = np.array([h[0] for h in hist])
    f = np.array([h[1] for h in hist])                       # [step][cell][field]
    gt = np.array(g["times"])
    at = lambda tt, cell, col: np.interp(tt, t, f[:, cell, col])
    etab = []
    for col in range(3):
        mine, ref = [], []
        for ti, tab in zip(g["table_index"], g["tables"]):
            if gt[ti] <= 0 or gt[ti] > t[-1] * (1 + 1e-9):
                continue
            tab = np.array(tab)
            mine.append([at(gt[ti], c, col) for c in range(g["ncell"])])
            ref.append(tab[:, col])
        etab.append(rel(np.concatenate(mine), np.concatenate(ref)) if np.abs(np.concatenate(ref)).max() > 0 else
                    float(np.abs(np.concatenate(mine)).max()))
    sel = (gt > 0) & (gt <= t[-1] * (1 + 1e-9))
    ehist = []
    for col in range(3):
        mine = np.array([[at(tt, c, col) for c in g["history_cells"]] for tt in gt[sel]])
        ref = np.array(g["history"])[sel][:, :, col]
        ehist.append(rel(mine, ref) if np.abs(ref).max() > 0 else float(np.abs(mine).max()))
    st = np.array(g["source_times"])
    s2 = (st > 0) & (st <= t[-1] * (1 + 1e-9))
    eh = rel(np.interp(st[s2], t, [h[2] for h in hist]), np.array(g["production_enthalpy"])[s2])
    return etab, ehist, eh


@pytest.mark.parametrize("case", CASES)
def test_oracle_runs_reference_input_to_the_autough2_answer(wo, case):
    p, hist, y = run_oracle(wo, case)
    etab, ehist, eh = compare(case, hist)
    tol = TOL[case]
    assert all(e < tl for e, tl in zip(etab, tol[:3])), (case, "tables", etab)
    assert all(e < tl for e, tl in zip(ehist, tol[:3])), (case, "histories", ehist)
    assert eh < tol[3], (case, "enthalpy", eh)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["problem2c", "problem4", "problem5b"])
def test_cuda_path_runs_reference_input(wo, case):
    from waiwera_b200 import flow
    p_ref, hist_ref, y_ref = run_oracle(wo, case)
    p = ingest.load(os.path.join(INP, case + ".input.json"), mod=flow)
    m = p.mesh
    sim = flow.FlowSimulation(p.params, m)
    if len(p.boundary_region):
        assert sim.set_boundaries(m.boundary["ghost_cells"], m.boundary["interior_cells"], p.boundary_primary, p.boundary_region) == 0
    assert sim.set_sources(p.source_cells, p.source_components, p.source_rates, p.source_enthalpies) == 0
    assert sim.fluid_init(p.y, p.region) == 0
    hist, y = run_input(p, sim, opts=newton_opts(flow, p))
    etab, ehist, eh = compare(case, hist)
    tol = TOL[case]
    assert all(e < tl for e, tl in zip(etab, tol[:3])) and all(e < tl for e, tl in zip(ehist, tol[:3])) and eh < tol[3]
    assert len(hist) == len(hist_ref)
    assert np.abs(y - y_ref).max() / np.abs(y_ref).max() < 1e-4
    sim.destroy()
