"""The reference's one-cell tracer decay benchmark (test/benchmark/tracer/decay: decay.json, test_tracer_decay.py):
eos we, 1 MPa / 60 degC, porosity 0.1, three tracers (no decay; constant decay 1e-6 1/s; Arrhenius decay with
activation energy 2 kJ/mol), initial mass fraction 1e-3, BDF2 with 20 steps of 86 400 s (first step backward Euler,
src/timestepper.F90:519-523).  The benchmark's expected result is the exact solution X exp(-k t) within 1e-2; the
time discretisation itself is pinned much tighter here by the closed-form BE / BDF2 recurrences.
Two cells instead of one (no flow between them: no gravity, uniform state) so that the mesh has a face."""
import ctypes as C
from math import exp

import numpy as np
import pytest

from waiwera_b200 import mesh as wmesh

DT, NSTEPS, X0 = 86400.0, 20, 1.0e-3
K0, EA, RGAS, TEMP = 1.0e-6, 2.0e3, 8.3144598, 60.0
RATES = [0.0, K0, K0 * exp(-EA / (RGAS * (TEMP + 273.15)))]


def problem():
    m = wmesh.structured(2, 1, 1, dx=10.0, gravity=(0.0, 0.0, 0.0), heterogeneous=False)
    m.rock[:, 5] = 0.1
    primary = np.tile([1.0e6, TEMP], (2, 1))
    region = np.ones(2, np.int32)
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1)
    return m, y, region


def params(wo):
    return wo.make_params(eos=wo.EOS_WE, thermo=wo.THERMO_IAPWS, gravity=(0.0, 0.0, 0.0))


def recurrence():
    """BE then BDF2 (constant step, r = 1) applied to dX/dt = -k X in closed form"""
    out = []
    for k in RATES:
        xs = [X0, X0 / (1.0 + DT * k)]
        for _ in range(NSTEPS - 1):
            xs.append((4.0 * xs[-1] - xs[-2]) / (3.0 + 2.0 * DT * k))
        out.append(xs[1:])
    return np.array(out).T          # [step][tracer]


def check_history(hist):
    hist = np.array(hist)
    t = DT * np.arange(1, NSTEPS + 1)
    for it, k in enumerate(RATES):
        exact = X0 * np.exp(-k * t)
        assert np.abs(hist[:, it] - exact).max() < 1e-2 * X0        # the benchmark's tolerance
    assert np.abs(hist - recurrence()).max() < 1e-12 * X0


def ksp(wo):
    k = wo.KspOpts()
    k.type, k.restart, k.maxit, k.rtol, k.atol, k.dtol = wo.KSP_BCGS, 30, 100, 1e-14, 1e-50, 1e5
    return k


def test_oracle_decay_benchmark(wo):
    m, y, region = problem()
    f = wo.Flow(params(wo), m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    f.set_tracers([1, 1, 1], decay=[0.0, K0, K0], activation=[0.0, 0.0, EA])
    assert f.fluid_init(y, region) == 0
    err, L0 = f.lhs(y)
    assert err == 0
    assert f.residual(y, L0, DT)[0] == 0
    L = wo.lib()
    A = f.tracer_pattern()
    al = f.tracer_balances()
    x, x2, al2 = np.full(6, X0), None, None
    hist = []
    for step in range(NSTEPS):
        b, al_new = f.tracer_setup_linear(A, DT, al, x, method=0 if step == 0 else 1, dt_last=DT, al_last2=al2, x_last2=x2)
        pc = L.wo_pc_create(A, wo.PC_BJACOBI_ILU0, None)
        xn = np.zeros(6)
        its, rn = C.c_int(), C.c_double()
        assert L.wo_ksp_solve(A, pc, C.byref(ksp(wo)), wo.dp(b), wo.dp(xn), C.byref(its), C.byref(rn)) > 0
        L.wo_pc_destroy(pc)
        x2, al2, x, al = x, al, xn, al_new
        hist.append(x[:3].copy())
        assert np.array_equal(x[:3], x[3:])          # both cells identical
    L.wo_bsr_destroy(A)
    check_history(hist)


@pytest.mark.gpu
def test_cuda_decay_benchmark(wo):
    from waiwera_b200 import flow
    from util import wb_params_from_oracle
    m, y, region = problem()
    sim = flow.FlowSimulation(wb_params_from_oracle(wo, flow, params(wo)), m)
    assert sim.set_tracers([1, 1, 1], decay=[0.0, K0, K0], activation=[0.0, 0.0, EA]) == 0
    assert sim.fluid_init(y, region) == 0
    err, L0 = sim.lhs(y)
    assert err == 0
    al = sim.tracer_balances()
    x, x2, al2 = np.full(6, X0), None, None
    hist = []
    for step in range(NSTEPS):
        if step == 1:
            assert sim.set_method(flow.METHOD_BDF2, DT, L0) == 0
        x_new, al_new, reason, its = sim.tracer_solve(DT, al, x, None, al2, x2,
                                                      opts=flow.ksp_opts(type=flow.KSP_BCGS, rtol=1e-14, maxit=100))
        assert reason > 0
        x2, al2, x, al = x, al, x_new, al_new
        hist.append(x[:3].copy())
    check_history(hist)
    sim.destroy()
