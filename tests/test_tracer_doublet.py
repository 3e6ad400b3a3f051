"""The reference's tracer doublet benchmark (test/benchmark/tracer/doublet: doublet_ss.json, doublet.json,
test_doublet.py) from its own input files: a 100-cell row (200 m x 10 m x 10 m), eos we, IFC-67, 0.5 kg/s of
100 degC water injected at one end, a producer on deliverability with a total-flow limiter at the other
(wb_set_source_controls), a 3.6 h tracer pulse in the injector (rate table, step interpolation, endpoint averaging),
tracer diffusion 1e-4 m2/s, adaptive backward Euler to half a year.  Everything the auxiliary tracer solve has:
advection by the stored phase fluxes, diffusion, table injection, production through a controlled source.
Golden: the steady state in the shipped Waiwera output file doublet_ss.h5 (the transient's initial condition) and
the AUTOUGH2 listing (tests/golden/tracer_doublet.json); the reference accepts 1e-3 (absolute 1e-6) on the tracer mass
fraction at every output and 1e-3 on the tracer production rate history."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from util import OracleSim, run_adaptive
from waiwera_b200 import ingest
from waiwera_b200 import mesh as wmesh

HERE = os.path.dirname(os.path.abspath(__file__))
INP = os.path.join(HERE, "golden", "inputs")
GOLD = json.load(open(os.path.join(HERE, "golden", "tracer_doublet.json")))
N = 100


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


def setup_sources(p, sim):
    assert sim.set_sources(p.source_cells, p.source_components, p.source_rates, p.source_enthalpies) in (0, None)
    c = p.source_controls
    assert all(k["productivity"] is not None for k in c)
    r = sim.set_source_controls([k["source"] for k in c], [k["productivity"] if k["deliverability"] else 0.0 for k in c],
                                [k["reference_pressure"] for k in c], [k["direction"] for k in c], [k["limit"] for k in c])
    assert r in (0, None)


class OracleTracerSim(OracleSim):
    """OracleSim + the tracer calls of flow.FlowSimulation"""

    def set_source_controls(self, *a):
        self.f.set_source_controls(*a)

    def set_tracers(self, phases, diffusion=None, decay=None, activation=None):
        self.f.set_tracers(phases, diffusion, decay, activation)
        self.A = self.f.tracer_pattern()

    def set_tracer_injection(self, rates):
        self.f.set_tracer_injection(rates)

    def tracer_balances(self):
        return self.f.tracer_balances()

    def tracer_solve(self, dt, al, x, xb=None, opts=None):
        b, al_new = self.f.tracer_setup_linear(self.A, dt, al, x)
        k = self.wo.KspOpts()
        k.type, k.restart, k.maxit, k.rtol, k.atol, k.dtol = self.wo.KSP_BCGS, 30, 10000, 1e-10, 1e-50, 1e5
        pc = self.L.wo_pc_create(self.A, self.wo.PC_BJACOBI_ILU0, None)
        xn = np.zeros(len(b))
        its, rn = C.c_int(), C.c_double()
        reason = self.L.wo_ksp_solve(self.A, pc, C.byref(k), self.wo.dp(b), self.wo.dp(xn), C.byref(its), C.byref(rn))
        self.L.wo_pc_destroy(pc)
        return xn, al_new, reason, its.value

    def source_rates(self):
        return self.f.source_rates(self.f.L and 2)


def run_transient(p, sim, y, opts=None, ksp=None):
    """the transient of doublet.json: flow Newton step, then the tracer solve, per adaptive time step"""
    st, ad = p.time["step"], p.time["step"]["adapt"]
    stop, dt_max, nmax = p.time["stop"], st["maximum"]["size"], st["maximum"]["number"]
    sim.set_tracers([1], diffusion=[p.tracers[0].get("diffusion", 0.0)])
    err, _ = sim.lhs(y)
    assert err == 0
    al = sim.tracer_balances()
    x = np.zeros(len(al))
    t, dt, k = 0.0, st["size"], 0
    hist = []
    while t < stop * (1 - 1e-12) and k < nmax:
        dt = min(dt, stop - t, dt_max)
        t1, _, its, _ = run_adaptive(sim, y, dt, dt, opts=opts, max_steps=1, reduction=ad["reduction"], amplification=1.0,
                                     its_min=0, its_max=10 ** 9)
        sim.set_tracer_injection(ingest.tracer_rates_at(p, t, t + t1))
        x, al, reason, _ = (sim.tracer_solve(t1, al, x, None, opts=ksp) if ksp is not None else sim.tracer_solve(t1, al, x))
        assert reason > 0
        t += t1
        k += 1
        rate = sim.source_rates()[1]
        hist.append((t, np.asarray(sim.fluid())[:N, 0].copy(), x[:N].copy(), rate, rate * x[N - 1]))
        dt = t1 * (ad["amplification"] if its < ad["minimum"] else 1.0)
    return hist


def make_oracle(wo, name, y=None, region=None):
    p = ingest.load(os.path.join(INP, name + ".input.json"), mod=wo)
    m = p.mesh
    f = wo.Flow(p.params, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    sim = OracleTracerSim(wo, f, newton_opts(wo))
    y = p.y.copy() if y is None else y
    region = p.region if region is None else region
    setup_sources(p, sim)
    assert f.fluid_init(y, region) == 0
    return p, f, sim, y


def steady_state_primaries():
    primary = np.stack([GOLD["steady_pressure"], GOLD["steady_temperature"]], 1)
    region = np.ones(N, np.int32)
    return np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1), region


def check(hist):
    t = np.array([h[0] for h in hist])
    X = np.array([h[2] for h in hist])
    P = np.array([h[1] for h in hist])
    worst = 0.0
    for ti, tab, ptab in zip(GOLD["times"], GOLD["tracer"], GOLD["pressure"]):
        tab, ptab = np.array(tab), np.array(ptab)
        k = np.argmin(np.abs(t - ti))
        assert abs(t[k] - ti) <= 1e-3 * ti + 30.0          # same step history (AUTOUGH2 stops at 1.578e7 s)
        err = np.abs(X[k] - tab).max()
        worst = max(worst, err / max(tab.max(), 1e-300))
        assert err <= 1e-3 * tab.max() + 1e-9, (ti, err, tab.max())
        assert np.abs(P[k] - ptab).max() / ptab.max() < 1e-4
    st = np.array(GOLD["source_times"])
    gp = np.array(GOLD["tracer_production"])
    mine = np.interp(st[1:], t, [h[4] for h in hist])
    assert np.abs(mine - gp[1:]).max() <= 1e-3 * np.abs(gp).max() + 1e-12
    return worst


def test_oracle_steady_state_matches_waiwera_output(wo):
    """doublet_ss.json to t = 1e15 s against the doubles in the shipped Waiwera result file"""
    p, f, sim, y = make_oracle(wo, "doublet_ss")
    st = p.time["step"]
    run_adaptive(sim, y, st["size"], p.time["stop"], max_steps=st["maximum"]["number"], reduction=st["adapt"]["reduction"],
                 amplification=st["adapt"]["amplification"], its_min=st["adapt"]["minimum"], its_max=st["adapt"]["maximum"])
    P, T = y[0::2] * 1.0e6, y[1::2] * 1.0e2
    assert np.abs(P - np.array(GOLD["steady_pressure"])).max() < 1e-3          # Pa
    assert np.abs(T - np.array(GOLD["steady_temperature"])).max() < 1e-8
    assert abs(f.source_rates(2)[1] + 0.5) < 1e-9                                # the producer takes what is injected
    sim.destroy()


def test_oracle_matches_autough2_tracer_doublet(wo):
    y0, region = steady_state_primaries()
    p, f, sim, y = make_oracle(wo, "doublet", y0, region)
    hist = run_transient(p, sim, y)
    check(hist)
    sim.destroy()


@pytest.mark.gpu
def test_cuda_path_reproduces_tracer_doublet(wo):
    from waiwera_b200 import flow
    y0, region = steady_state_primaries()
    p_ref, f, osim, y_ref = make_oracle(wo, "doublet", y0, region)
    hist_ref = run_transient(p_ref, osim, y_ref)
    osim.destroy()
    p = ingest.load(os.path.join(INP, "doublet.input.json"), mod=flow)
    sim = flow.FlowSimulation(p.params, p.mesh)
    setup_sources(p, sim)
    assert sim.fluid_init(y0, region) == 0
    y = y0.copy()
    hist = run_transient(p, sim, y, opts=newton_opts(flow), ksp=flow.ksp_opts(type=flow.KSP_BCGS, rtol=1e-10))
    check(hist)
    assert len(hist) == len(hist_ref)
    xr, xg = np.array([h[2] for h in hist_ref]), np.array([h[2] for h in hist])
    assert np.abs(xg - xr).max() <= 1e-6 * xr.max()
    sim.destroy()
