"""The reference's 1-D MINC doublet benchmark (test/benchmark/minc/doublet_1d: minc_1d_{single,50,100,200}.json,
test_minc_1d.py): ten 50 m cubes, eos we, IFC-67, two-phase at 8.5 MPa, cold water (500 kJ/kg) injected at one end and
the same rate produced at the other for 50 years; single porosity, and MINC over the whole mesh with one matrix level
(volume fractions 0.1 / 0.9, three fracture planes) at fracture spacings of 50, 100 and 200 m -- the thermal sweep
depends on the fracture-matrix heat exchange, i.e. on the MINC geometry (src/minc.F90:393-544).  Golden: the last
ELEMENT table of the AUTOUGH2 listings (tests/golden/minc_doublet.json); the reference accepts 2e-3 on P, T, Sv."""
import json
import os

import numpy as np
import pytest

from util import OracleSim, run_adaptive, we_fields
from waiwera_b200 import mesh as wmesh

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "minc_doublet.json")))
NX, DX = 10, 50.0
CASES = ["single", "50", "100", "200"]


def problem(case):
    m = wmesh.structured(NX, 1, 1, dx=DX, gravity=(0.0, 0.0, 0.0), heterogeneous=False)
    m.rock[:, 0:3] = 6e-15
    m.rock[:, 3:5] = 2.1
    m.rock[:, 5], m.rock[:, 6], m.rock[:, 7] = 0.1, 2650.0, 1000.0
    if case != "single":
        m.rock[:, 5] = 0.5                   # rock type "fract"
        matrix = [1e-18, 1e-18, 1e-18, 2.1, 2.1, 0.0555555555556, 2650.0, 1000.0]
        sp = float(case)
        m = wmesh.add_minc(m, volumes=(0.1, 0.9), spacing=(sp, sp, sp), matrix_rock=matrix)
    n = m.ninterior
    primary = np.tile([8.5e6, 0.01], (n, 1))
    region = np.full(n, 4, np.int32)
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1)
    return m, y, region


def params(mod):
    return mod.make_params(eos=mod.EOS_WE, thermo=mod.THERMO_IFC67, gravity=(0.0, 0.0, 0.0),
                           relperm=mod.make_relperm("corey", slr=0.3, ssr=0.05),
                           cappress=mod.make_cappress("linear", saturation_limits=(0.0, 1.0), pressure=0.0))


def newton_opts(mod):
    if hasattr(mod, "newton_opts"):
        return mod.newton_opts(max_iterations=8, rel_tol=1e-5, pc_type=mod.PC_BJACOBI_ILU0, ksp=mod.ksp_opts(type=mod.KSP_BCGS))
    o = mod.NewtonOpts()
    o.max_iterations, o.min_iterations = 8, 0
    o.rel_tol, o.abs_tol, o.update_rel_tol, o.update_abs_tol = 1e-5, 1.0, 1e-10, 1.0
    o.fd_err, o.fd_umin, o.pc_type = 1e-8, 1e-2, mod.PC_BJACOBI_ILU0
    o.ksp.type, o.ksp.restart, o.ksp.maxit = mod.KSP_BCGS, 30, 10000
    o.ksp.rtol, o.ksp.atol, o.ksp.dtol = 1e-5, 1e-50, 1e5
    return o


SOURCES = ([0, NX - 1], [1, 0], [0.1, -0.1], [500000.0, 0.0])


def run(case, sim, y, opts=None):
    g = GOLD[case]
    st, ad = g["step"], g["step"]["adapt"]
    t, dt = 0.0, st["size"]
    dt_max = st["maximum"]["size"] or np.inf
    while t < g["stop"] * (1 - 1e-12):
        dt = min(dt, g["stop"] - t, dt_max)
        t1, _, its, _ = run_adaptive(sim, y, dt, dt, opts=opts, max_steps=1, reduction=ad["reduction"], amplification=1.0,
                                     its_min=0, its_max=10 ** 9)
        t += t1
        dt = t1 * (ad["amplification"] if its < ad["minimum"] else 1.0)
    return t


def check(case, out):
    gold = np.array(GOLD[case]["element"])
    assert out.shape == gold.shape
    errs = [np.linalg.norm(out[:, c] - gold[:, c]) / np.linalg.norm(gold[:, c]) for c in range(3)]
    assert all(e < 2e-3 for e in errs), (case, errs)
    return errs


def run_oracle(wo, case):
    m, y, region = problem(case)
    f = wo.Flow(params(wo), m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    f.set_sources(*SOURCES)
    assert f.fluid_init(y, region) == 0
    sim = OracleSim(wo, f, newton_opts(wo))
    run(case, sim, y)
    out = we_fields(f.fluid(), m.ninterior)
    sim.destroy()
    return out, y


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_autough2_minc_doublet(wo, case):
    out, y = run_oracle(wo, case)
    check(case, out)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["single", "100"])
def test_cuda_path_reproduces_minc_doublet(wo, case):
    from waiwera_b200 import flow
    out_ref, y_ref = run_oracle(wo, case)
    m, y, region = problem(case)
    sim = flow.FlowSimulation(params(flow), m)
    assert sim.set_sources(*SOURCES) == 0
    assert sim.fluid_init(y, region) == 0
    run(case, sim, y, opts=newton_opts(flow))
    out = we_fields(sim.fluid(), m.ninterior)
    check(case, out)
    assert np.abs(out - out_ref).max(axis=0).tolist() < [500.0, 0.05, 1e-3]
    sim.destroy()
