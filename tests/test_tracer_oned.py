"""the passive-tracer auxiliary linear problem (src/flow_simulation.F90:1489-1959,
src/timestepper.F90:458-494, 2347-2353).  The reference pins this path with its tracer benchmarks; the 1-D
problems of test/benchmark/tracer/oned (test_tracer_1d.py) are transcribed here: 10 cells of 10 m x 1 m x 1 m,
eos we, IFC-67, no gravity, Dirichlet boundary with tracer mass fraction 0.01 on the x = 0 face of cell 0,
production in cell 9, backward Euler steps of 864 000 s from the state in the Waiwera output file the benchmark
ships as its initial condition (oned_*_ss.h5: uniform 3 MPa / 20 degC in the single-phase case, the steady two-phase
flow of oned_two_phase_ss.json in the other).  The AUTOUGH2 listings
shipped with the benchmark (committed as tests/golden/tracer_oned.json by tools/make_golden.py) are the golden
output; the reference accepts 1e-3 relative on pressure and tracer mass fraction at the last output.
The CUDA path then has to reproduce the oracle's run."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from waiwera_b200 import mesh as wmesh

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "tracer_oned.json")))
NX, DX, DT = 10, 10.0, 864000.0
CASES = {
    # boundary primaries / region, production rate (kg/s), number of transient steps
    "single": dict(primary=[3.0e6, 20.0], region=1, rate=-0.00277777777778, nsteps=10),
    "two": dict(primary=[1.0e5, 0.5], region=4, rate=-2.77777777778e-05, nsteps=30),
}
X_BOUNDARY = 0.01


def problem(case):
    c = CASES[case]
    m = wmesh.structured(NX, 1, 1, dx=DX, dy=1.0, dz=1.0, gravity=(0.0, 0.0, 0.0), heterogeneous=False)
    m.rock[:, 0:3] = 1e-13
    m.rock[:, 3:5] = 1.0
    m.rock[:, 5], m.rock[:, 6], m.rock[:, 7] = 0.1, 2500.0, 1000.0
    m = wmesh.add_boundary(m, [0], (-1.0, 0.0, 0.0), 0.5 * DX, 1.0, 1, gravity=(0.0, 0.0, 0.0))
    init = GOLD[case]["initial"]
    second = init["vapour_saturation"] if case == "two" else init["temperature"]
    primary = np.stack([init["pressure"], second], 1)
    region = np.full(NX, c["region"], np.int32)
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1)
    return m, y, region


def params(wo):
    return wo.make_params(eos=wo.EOS_WE, thermo=wo.THERMO_IFC67, gravity=(0.0, 0.0, 0.0),
                          relperm=wo.make_relperm("linear", liquid=(0.0, 1.0), vapour=(0.0, 1.0)),
                          cappress=wo.make_cappress("linear", saturation_limits=(0.0, 1.0), pressure=0.0))


def newton_opts(wo):
    o = wo.NewtonOpts()
    o.max_iterations, o.min_iterations = 8, 0
    o.rel_tol, o.abs_tol, o.update_rel_tol, o.update_abs_tol = 1e-9, 1.0, 1e-10, 1.0
    o.fd_err, o.fd_umin, o.pc_type = 1e-8, 1e-2, wo.PC_BJACOBI_ILU0
    o.ksp.type, o.ksp.restart, o.ksp.maxit = wo.KSP_BCGS, 30, 10000
    o.ksp.rtol, o.ksp.atol, o.ksp.dtol = 1e-5, 1e-50, 1e5
    return o


def ksp_opts(wo, rtol=1e-10):
    k = wo.KspOpts()
    k.type, k.restart, k.maxit, k.rtol, k.atol, k.dtol = wo.KSP_BCGS, 30, 10000, rtol, 1e-50, 1e5
    return k


def steady_state_steps():
    """oned_two_phase_ss.json: adaptive backward Euler from 864 000 s, doubling, to t = 1e15 s"""
    return [DT * 2.0 ** k for k in range(31)]


def run_oracle(wo, case, steady_state=False):
    c = CASES[case]
    m, y, region = problem(case)
    if steady_state:
        y = np.ascontiguousarray(wmesh.scale_primaries(np.tile(c["primary"], (NX, 1)), region)).reshape(-1)
    f = wo.Flow(params(wo), m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    assert f.set_boundary(int(m.boundary["ghost_cells"][0]), 0, np.array(c["primary"], float), c["region"]) == 0
    f.set_sources([NX - 1], [1], [c["rate"]], [0.0])
    f.set_tracers([1])
    assert f.fluid_init(y, region) == 0
    L = wo.lib()
    J = f.bsr()
    color = np.zeros(J.contents.nb, np.int32)
    nc = L.wo_bsr_coloring(J, wo.ip(color))
    o = newton_opts(wo)

    def flow_step(dt):
        err, L0 = f.lhs(y)
        assert err == 0
        L.wo_flow_pre_timestep(f.h)
        res = wo.NewtonResult()
        L.wo_newton_solve_be(f.h, J, wo.ip(color), nc, None, C.byref(o), dt, wo.dp(L0), wo.dp(y), C.byref(res))
        assert res.reason > 0, (dt, res.reason)
        # SNES ends on an unperturbed evaluation at the converged y: fluid, fluxes at the new state
        err, _, _, _ = f.residual(y, L0, dt)
        assert err == 0

    if steady_state:
        for dt in steady_state_steps():
            flow_step(dt)
        L.wo_bsr_destroy(J)
        return y.copy()
    y_ss = y.copy()

    A = f.tracer_pattern()
    n = f.ntrows
    x = np.zeros(n)
    x[NX] = X_BOUNDARY            # boundary ghost row: the Dirichlet mass fraction
    err, _ = f.lhs(y)
    assert err == 0
    al = f.tracer_balances()
    k = ksp_opts(wo)
    hist, prod = [], []
    for step in range(c["nsteps"]):
        flow_step(DT)
        b, al_new = f.tracer_setup_linear(A, DT, al, x)
        pc = L.wo_pc_create(A, wo.PC_BJACOBI_ILU0, None)
        xn = np.zeros(n)
        its, rn = C.c_int(), C.c_double()
        reason = L.wo_ksp_solve(A, pc, C.byref(k), wo.dp(b), wo.dp(xn), C.byref(its), C.byref(rn))
        L.wo_pc_destroy(pc)
        assert reason > 0
        x, al = xn, al_new
        hist.append((y.copy(), x.copy()))
        prod.append(c["rate"] * x[NX - 1])   # liquid-only production: tracer flow = flow fraction * rate * X
    L.wo_bsr_destroy(A)
    L.wo_bsr_destroy(J)
    return m, y_ss, hist, prod


@pytest.fixture(scope="module", params=["single", "two"])
def oracle_run(request, wo):
    return request.param, run_oracle(wo, request.param)


def test_oracle_matches_autough2_tracer_listing(oracle_run):
    """test_tracer_1d.py:86-98: FieldWithinTolTC(Pressure, Tracer mass fraction; tolerance 1e-3, absolute 1e-4) at
    the last output against AUTOUGH2, and the tracer production history at 1e-3"""
    case, (m, y_ss, hist, prod) = oracle_run
    g = GOLD[case]
    assert len(hist) == len(g["times"])
    tab = np.array(g["tables"][-1])[:NX]            # last output; rows a..j (the boundary block is the last row)
    y, x = hist[-1]
    P = y[0::2] * 1.0e6
    assert np.abs(P - tab[:, 0]).max() / np.abs(tab[:, 0]).max() < 1e-3
    X = x[:NX]
    assert np.abs(X - tab[:, 4]).max() < 1e-3 * np.abs(tab[:, 4]).max() + 1e-6, (X, tab[:, 4])
    # every output time, not only the last: the tracer front
    for (yk, xk), rows in zip(hist, g["tables"]):
        rows = np.array(rows)[:NX]
        assert np.abs(xk[:NX] - rows[:, 4]).max() < 2e-6     # measured 5.4e-7: the listing prints 6 digits
    if case == "single":
        gp = np.array([s[3] for s in g["source"]])
        assert np.abs(np.array(prod) - gp).max() < 2e-3 * np.abs(gp).max()


def test_oracle_tracer_system_properties(oracle_run):
    """structure of A = Al - dt Ar after aux_pre_solve: identity boundary row, zero-diffusion upwind coupling only
    to the upstream neighbour, mass fractions bounded by the boundary value"""
    case, (m, y_ss, hist, prod) = oracle_run
    for _, x in hist:
        assert abs(x[NX] - X_BOUNDARY) < 1e-14        # identity row, solved by the Krylov method like any other
        assert (x[:NX] >= -1e-12).all() and (x[:NX] <= X_BOUNDARY * (1 + 1e-9)).all()
        assert (np.diff(x[:NX]) <= 1e-12).all()      # monotone front behind the inlet
    xs = np.array([x[:NX] for _, x in hist])
    assert (np.diff(xs, axis=0) >= -1e-12).all()     # the front only advances


def test_oracle_steady_state_matches_waiwera_output(wo):
    """oned_two_phase_ss.json run to steady state with the oracle's Newton path against the Waiwera result file
    the benchmark ships (oned_two_phase_ss.h5): pressure and vapour saturation of the steady two-phase flow"""
    y = run_oracle(wo, "two", steady_state=True)
    init = GOLD["two"]["initial"]
    P, S = y[0::2] * 1.0e6, y[1::2]
    # measured: 1.5e-11 Pa and 5.6e-16 -- the oracle's Newton path lands on the real Waiwera run's doubles
    assert np.abs(P - np.array(init["pressure"])).max() < 1e-6
    assert np.abs(S - np.array(init["vapour_saturation"])).max() < 1e-11
