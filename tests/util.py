"""Shared helpers of the parity tests: build the same problem for the oracle and for the CUDA path."""
import ctypes as C

import numpy as np

from waiwera_b200 import mesh as wmesh

SEED = 20240917


def wb_params_from_oracle(wo, flow, prm):
    """wb_params with the same contents as the oracle's wo_params (independent struct definitions,
    identical field layout)."""
    return flow.make_params(eos={wo.EOS_WE: flow.EOS_WE, wo.EOS_W: flow.EOS_W, wo.EOS_WCE: flow.EOS_WCE, wo.EOS_WAE: flow.EOS_WAE}[prm.eos],
                            thermo=prm.thermo, relperm=prm.relperm, cappress=prm.cappress,
                            gravity=tuple(prm.gravity), extrapolate=prm.extrapolate,
                            eos_w_temperature=prm.eos_w_temperature,
                            partial_pressure_scale=prm.partial_pressure_scale)


def psat_fn(wo, thermo):
    th = wo.lib().wo_thermo_create(thermo, 0)

    def f(t):
        if np.ndim(t) > 0:
            return np.array([f(v) for v in np.asarray(t).reshape(-1)])
        p = C.c_double()
        assert wo.lib().wo_saturation_pressure(th, float(t), C.byref(p)) == 0
        return p.value
    return f


def make_problem(wo, dims=(8, 7, 6), thermo=0, two_phase_layers=0, top_boundary=False, relperm=None, cappress=None,
                 seed=SEED, dx=10.0):
    """mesh + initial state (scaled y, region) + oracle params"""
    m = wmesh.structured(*dims, dx=dx, seed=seed, top_boundary=top_boundary)
    primary, region = wmesh.hydrostatic_state(m, seed=seed, two_phase_layers=two_phase_layers,
                                              thermo_psat=psat_fn(wo, thermo))
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1)
    prm = wo.make_params(eos=wo.EOS_WE, thermo=thermo, relperm=relperm, cappress=cappress)
    return m, y, region, prm


def make_problem_wce(wo, dims=(8, 7, 6), thermo=0, two_phase_layers=0, relperm=None, cappress=None, seed=SEED,
                     dx=10.0, partial_pressure_scale=0.0):
    """eos_wce (water + CO2 + energy, 3 primaries, BAIJ bs=3): mesh + scaled state + oracle params"""
    m = wmesh.structured(*dims, dx=dx, seed=seed)
    primary, region = wmesh.wce_state(m, seed=seed, two_phase_layers=two_phase_layers, thermo_psat=psat_fn(wo, thermo))
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region, partial_pressure_scale=partial_pressure_scale)).reshape(-1)
    prm = wo.make_params(eos=wo.EOS_WCE, thermo=thermo, relperm=relperm, cappress=cappress,
                         partial_pressure_scale=partial_pressure_scale)
    return m, y, region, prm


def boundary_values(m):
    """Dirichlet values of the top boundary ghost cells: 1 bar, 15 degC liquid"""
    nb = len(m.boundary["ghost_cells"])
    return np.tile(np.array([1.0e5, 15.0]), (nb, 1)), np.ones(nb, np.int32)


def oracle_flow(wo, m, prm, y, region):
    f = wo.Flow(prm, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    if m.boundary:
        bp, br = boundary_values(m)
        for g, ic, p, r in zip(m.boundary["ghost_cells"], m.boundary["interior_cells"], bp, br):
            assert f.set_boundary(int(g), int(ic), p, int(r)) == 0
    assert f.fluid_init(y, np.ascontiguousarray(region, np.int32)) == 0
    return f


def gpu_flow(wo, flow, m, prm, y, region, device=0):
    sim = flow.FlowSimulation(wb_params_from_oracle(wo, flow, prm), m, device=device)
    if m.boundary:
        bp, br = boundary_values(m)
        assert sim.set_boundaries(m.boundary["ghost_cells"], m.boundary["interior_cells"], bp, br) == 0
    assert sim.fluid_init(y, region) == 0
    return sim


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


class OracleSim:
    """The oracle behind the method names of flow.FlowSimulation that the time-stepping helpers use, so that one
    driver runs the reference benchmarks through the oracle and through the CUDA path."""

    def __init__(self, wo, f, opts):
        self.wo, self.f, self.L, self.opts = wo, f, wo.lib(), opts
        self.J = f.bsr()
        self.color = np.zeros(self.J.contents.nb, np.int32)
        self.ncolor = self.L.wo_bsr_coloring(self.J, wo.ip(self.color))

    def lhs(self, y):
        return self.f.lhs(y)

    def pre_timestep(self):
        self.L.wo_flow_pre_timestep(self.f.h)

    def pre_retry_timestep(self):
        self.L.wo_flow_pre_retry_timestep(self.f.h)

    def newton_solve(self, y, L0, dt, opts=None):
        res = self.wo.NewtonResult()
        self.L.wo_newton_solve_be(self.f.h, self.J, self.wo.ip(self.color), self.ncolor, None, C.byref(self.opts), dt,
                                  self.wo.dp(L0), self.wo.dp(y), C.byref(res))
        return res

    def residual(self, y, L0, dt):
        return self.f.residual(y, L0, dt)

    def set_sources(self, cells, components, rates, enthalpies):
        self.f.set_sources(cells, components, rates, enthalpies)
        return 0

    def set_source_components(self, injection, production):
        self.f.set_source_components(injection, production)
        return 0

    def set_source_controls(self, sources, pi, pref, direction, limit):
        self.f.set_source_controls(sources, pi, pref, direction, limit)
        return 0

    def set_source_recharge(self, *a):
        self.f.set_source_recharge(*a)
        return 0

    def set_source_separators(self, *a):
        return self.f.set_source_separators(*a)

    def set_source_pressure_table(self, *a):
        return self.f.set_source_pressure_table(*a)

    def set_rock(self, rock):
        return self.f.set_rock(rock)

    def source_rates(self, n):
        return self.f.source_rates(n)

    def fluid(self):
        return self.f.fluid()

    def regions(self):
        return self.f.regions()

    def destroy(self):
        self.L.wo_bsr_destroy(self.J)


def run_adaptive(sim, y, dt0, t_stop, opts=None, max_steps=500, reduction=0.2, amplification=2.0, its_min=5,
                 its_max=8, max_tries=20):
    """timestepper_step with the "iteration" step-size adaptor (src/timestepper.F90:863-1476, 2330-2375): a step that
    does not converge is retried with the step size times `reduction` after pre_retry_timestep; a step that converged
    in fewer than its_min Newton iterations is followed by one `amplification` times larger (more than its_max:
    reduced).  y is advanced in place; returns (time, steps, total Newton iterations, retries)."""
    t, dt, nsteps, nits, nretry = 0.0, dt0, 0, 0, 0
    while t < t_stop * (1.0 - 1e-14) and nsteps < max_steps:
        dt = min(dt, t_stop - t)
        err, L0 = sim.lhs(y)
        assert err == 0
        sim.pre_timestep()
        y0 = y.copy()
        for attempt in range(max_tries):
            res = sim.newton_solve(y, L0, dt, opts)
            if res.reason > 0:
                break
            nretry += 1
            dt *= reduction
            y[:] = y0
            sim.pre_retry_timestep()
        assert res.reason > 0, (t, dt, res.reason)
        t += dt
        nsteps += 1
        nits += res.iterations
        if res.iterations < its_min:
            dt *= amplification
        elif res.iterations > its_max:
            dt *= reduction
    err, L0 = sim.lhs(y)        # unperturbed evaluation: fluid records of the final state
    assert err == 0
    return t, nsteps, nits, nretry


def we_fields(fluid, n):
    """P, T, vapour saturation of the first n cells from eos_we fluid records (23 doubles)"""
    fl = np.asarray(fluid)[:n]
    return np.stack([fl[:, 0], fl[:, 1], fl[:, 7 + 8 + 2]], 1)


def we_production_enthalpy(fl):
    """enthalpy of the fluid a producing source takes from an eos_we cell (src/fluid.F90:417-436)"""
    phases = int(round(fl[4]))
    mob = [(fl[7 + 8 * p + 3] * fl[7 + 8 * p] / fl[7 + 8 * p + 1]) if phases & (1 << p) else 0.0 for p in range(2)]
    return (mob[0] * fl[7 + 5] + mob[1] * fl[7 + 8 + 5]) / (mob[0] + mob[1])


def wge_fields(fluid, n):
    """P, T, vapour saturation, gas mass fraction in the vapour and in the liquid, gas partial pressure of the first n
    cells from eos_wge (wce / wae) fluid records (26 doubles)"""
    fl = np.asarray(fluid)[:n]
    return np.stack([fl[:, 0], fl[:, 1], fl[:, 17 + 2], fl[:, 17 + 8], fl[:, 8 + 8], fl[:, 7]], 1)


def apply_input_controls(p, sim, t0, t1):
    """the source controls of an ingested input for the step [t0, t1] (ingest.controls_at: tables in time averaged over
    the step): deliverability / direction / total limiter, recharge, separators with water / steam limits.  First call:
    a productivity index to be calculated from the given rate is taken from the fluid state sim holds
    (calculate_PI_from_rate, src/source_control.F90:407-468; eos we), a recharge reference pressure "initial" from the
    cell's pressure."""
    from waiwera_b200 import ingest
    fl = None
    for c in p.source_controls:
        if c["deliverability"] and c["productivity"] is None:
            fl = np.asarray(sim.fluid()) if fl is None else fl
            f0 = fl[int(p.source_cells[c["source"]])]
            phases = int(round(f0[4]))
            mob = sum((f0[7 + 8 * q + 3] * f0[7 + 8 * q] / f0[7 + 8 * q + 1]) for q in range(2) if phases & (1 << q))
            c["productivity"] = abs(p.source_rates[c["source"]]) / (mob * (f0[0] - c["reference_pressure"]) * f0[5])
    for c in getattr(p, "source_recharge", []):
        if c["reference_pressure"] is None:
            fl = np.asarray(sim.fluid()) if fl is None else fl
            c["reference_pressure"] = float(fl[int(p.source_cells[c["source"]])][0])
    ctrl, seps = ingest.controls_at(p, t0, t1)
    if ctrl:
        r = sim.set_source_controls([c["source"] for c in ctrl], [c["productivity"] if c["deliverability"] else 0.0 for c in ctrl],
                                    [c["reference_pressure"] or 0.0 for c in ctrl], [c["direction"] for c in ctrl],
                                    [c["limit"] for c in ctrl])
        assert not r
    rc = getattr(p, "source_recharge", [])
    if rc:
        r = sim.set_source_recharge([c["source"] for c in rc], [c["coefficient"] for c in rc], [c["reference_pressure"] for c in rc])
        assert not r
    if seps:
        r = sim.set_source_separators([q["source"] for q in seps], [q["pressure"] for q in seps],
                                      [q["limit_water"] for q in seps], [q["limit_steam"] for q in seps])
        assert not r
    pt = getattr(p, "source_pressure_tables", [])
    if pt:
        r = sim.set_source_pressure_table([q["source"] for q in pt], [q["table"] for q in pt], [q["coordinate"] for q in pt],
                                          [q["step"] for q in pt])
        assert not r


def run_input(problem, sim, opts=None, fields=None, controls=False, well=0, on_step=None):
    """Runs an ingested input (waiwera_b200.ingest.Problem, eos_we) through `sim` (flow.FlowSimulation or OracleSim,
    mesh / boundaries / fluid_init already done) with the time stepping of its "time" value: a list of step sizes
    (the last one repeats), or one size with the "iteration" adaptor, maximum size / number, stop time; table sources
    are averaged over each step; with `controls` the source controls of the input (ingest.controls_at: tables in time
    averaged over the step) are set before every step.  Returns [(time, P/T/Sv of the interior cells, production
    enthalpy of source `well`)]; on_step(time, sim) is called after every step."""
    from waiwera_b200 import ingest
    p = problem
    st = p.time["step"]
    sizes = st["size"] if isinstance(st["size"], list) else [st["size"]]
    adapt = st.get("adapt", {}).get("on", False)
    dt_max = (st.get("maximum") or {}).get("size") or np.inf
    nmax = (st.get("maximum") or {}).get("number") or 10 ** 9
    stop = p.time.get("stop")
    stop = np.inf if stop is None else stop
    ad = st.get("adapt", {})
    n = p.mesh.ninterior
    y = p.y.copy()
    t, k, dt = 0.0, 0, sizes[0]
    hist = []
    prod = int(p.source_cells[well]) if len(p.source_cells) else 0
    while t < stop * (1 - 1e-12) and k < nmax:
        if not adapt:
            dt = sizes[min(k, len(sizes) - 1)]
        dt = min(dt, stop - t, dt_max)
        if p.source_tables:
            rates = ingest.rates_at(p, t, t + dt)
            assert sim.set_sources(p.source_cells, ingest.components_at(p, rates), rates, p.source_enthalpies) == 0
            sim.set_source_components(p.source_injection_components, p.source_production_components)
        if controls:
            apply_input_controls(p, sim, t, t + dt)
        t1, _, its, _ = run_adaptive(sim, y, dt, dt, opts=opts, max_steps=1, reduction=ad.get("reduction", 0.2),
                                     amplification=1.0, its_min=0, its_max=10 ** 9)
        t += t1
        k += 1
        fl = sim.fluid()
        if fields is None:
            hist.append((t, we_fields(fl, n), we_production_enthalpy(np.asarray(fl)[prod])))
        else:
            hist.append((t, fields(fl, n), 0.0))
        if on_step is not None:
            on_step(t, sim)
        dt = t1
        if adapt:
            if its < ad.get("minimum", 5):
                dt = dt * ad.get("amplification", 2.0)
            elif its > ad.get("maximum", 8):
                dt = dt * ad.get("reduction", 0.2)
    return hist, y
