"""Runs a Waiwera input deck through the CUDA path: `python -m waiwera_b200.run deck.json [-o output.h5]`.

The host side the reference keeps in timestepper.F90 / flow_simulation.F90 around the Newton step, restated as the few
lines a caller of this library needs -- ingest.load -> flow.FlowSimulation -> backward-Euler steps with the input's step
sizes, "iteration" step-size adaptor (src/timestepper.F90:863-1476, 2330-2375), step cuts on failed Newton solves,
source tables and controls averaged over each step (ingest.rates_at / controls_at), the tracer solve after every flow
step (timestepper.F90:2347-2353) -> the output file of
output.write_output (the layout CREDO's benchmark scripts and `initial.filename` restarts read).  It is not on the hot
path and not part of the parity claims; tests/test_run.py drives it with the checker standing in for the engine on CPU
and with the engine itself on the GPU.

run(problem, sim, ...) takes any object with the method names of flow.FlowSimulation."""
import argparse
import json
import os
import sys

import numpy as np

from . import ingest, output


def production_enthalpy(rec, ncomp, nphase):
    """flowing enthalpy of a producing source from the fluid record of its cell (src/fluid.F90:417-436): phase
    enthalpies weighted by the mobilities"""
    base, stride = 7 + ncomp - 1, 8 + ncomp - 1
    phases = int(round(rec[4]))
    mob = [rec[base + q * stride + 3] * rec[base + q * stride] / rec[base + q * stride + 1] if phases & (1 << q) else 0.0
           for q in range(nphase)]
    tot = sum(mob)
    return sum(m * rec[base + q * stride + 5] for q, m in enumerate(mob)) / tot if tot > 0 else 0.0


def _source_rates(sim, n):
    try:
        return np.asarray(sim.source_rates(), float)
    except TypeError:
        return np.asarray(sim.source_rates(n), float)


def apply_controls(p, sim, t0, t1):
    """source rate tables and controls of the input for the step [t0, t1].  On the first call a productivity index to be
    calculated from the given rate is taken from the fluid state sim holds (calculate_PI_from_rate,
    src/source_control.F90:407-468), a recharge reference pressure "initial" from the cell's pressure."""
    nsrc = len(p.source_cells)
    if nsrc == 0:
        return
    if p.source_tables or any(key == "factor" for _, key in getattr(p, "source_control_tables", {})):
        rates = ingest.rates_at(p, t0, t1)
        assert not sim.set_sources(p.source_cells, ingest.components_at(p, rates), rates, p.source_enthalpies)
        sim.set_source_components(p.source_injection_components, p.source_production_components)
    fl = None
    nc, nph = output._EOS[p.eos]
    for c in p.source_controls:
        if c["deliverability"] and c["productivity"] is None:
            fl = np.asarray(sim.fluid()) if fl is None else fl
            rec = fl[int(p.source_cells[c["source"]])]
            base, stride = 7 + len(nc) - 1, 8 + len(nc) - 1
            phases = int(round(rec[4]))
            mob = sum(rec[base + q * stride + 3] * rec[base + q * stride] / rec[base + q * stride + 1]
                      for q in range(len(nph)) if phases & (1 << q))
            c["productivity"] = abs(p.source_rates[c["source"]]) / (mob * (rec[0] - c["reference_pressure"]) * rec[5])
    for c in getattr(p, "source_recharge", []):
        if c["reference_pressure"] is None:
            fl = np.asarray(sim.fluid()) if fl is None else fl
            c["reference_pressure"] = float(fl[int(p.source_cells[c["source"]])][0])
    ctrl, seps = ingest.controls_at(p, t0, t1)
    if ctrl:
        assert not sim.set_source_controls([c["source"] for c in ctrl], [c["productivity"] if c["deliverability"] else 0.0 for c in ctrl],
                                           [c["reference_pressure"] or 0.0 for c in ctrl], [c["direction"] for c in ctrl],
                                           [c["limit"] for c in ctrl])
    rc = getattr(p, "source_recharge", [])
    if rc:
        assert not sim.set_source_recharge([c["source"] for c in rc], [c["coefficient"] for c in rc],
                                           [c["reference_pressure"] for c in rc])
    if seps:
        assert not sim.set_source_separators([q["source"] for q in seps], [q["pressure"] for q in seps],
                                             [q["limit_water"] for q in seps], [q["limit_steam"] for q in seps])
    pt = getattr(p, "source_pressure_tables", [])
    if pt:
        assert not sim.set_source_pressure_table([q["source"] for q in pt], [q["table"] for q in pt],
                                                 [q["coordinate"] for q in pt], [q["step"] for q in pt])


def setup_tracers(p, sim):
    """the "tracer" value of the input (setup_tracers, src/tracer.F90:64-150) -> (initial mass fractions [nowned * nt],
    mass fractions of the boundary ghost cells or None)"""
    nt = len(p.tracers)
    phase = {"liquid": 1, "vapour": 2}
    sim.set_tracers([phase[str(t.get("phase", "liquid")).lower()] for t in p.tracers],
                    diffusion=[float(t.get("diffusion", 0.0)) for t in p.tracers],
                    decay=[float(t.get("decay", 0.0)) for t in p.tracers],
                    activation=[float(t.get("activation", 0.0)) for t in p.tracers])
    m = p.mesh
    x = np.tile(np.asarray(p.initial_tracer, float)[:nt], m.nowned)
    xb = np.ascontiguousarray(p.boundary_tracer[:, :nt], float).reshape(-1) if len(p.boundary_region) else None
    return x, xb


def run(p, sim, opts=None, log=None, tracer_ksp=None, tracer_history=None):
    """Advances the ingested problem p on sim from time.start to time.stop.  Returns (times, fluids, source_history, y):
    the output times (the initial state first), the fluid records [ncell, dof] at those times, per output time the
    [nsources, 3] array of (component, rate, enthalpy), and the final scaled primaries.  With tracers in the input the
    auxiliary linear problem is solved after every flow step (timestepper.F90:2347-2353) and the mass fractions
    [nowned, nt] of every output time are appended to the list tracer_history."""
    tm = p.time or {}
    st = tm.get("step", {})
    size = st.get("size", 0.1)
    sizes = list(size) if isinstance(size, list) else [size]
    ad = st.get("adapt", {}) or {}
    adapt = bool(ad.get("on", False))
    method = str(st.get("method", "beuler")).lower()
    if method != "beuler":
        raise NotImplementedError("time.step.method %r: this driver steps with backward Euler only" % method)
    if adapt and str(ad.get("method", "iteration")).lower() != "iteration":
        raise NotImplementedError("time.step.adapt.method %r: only the iteration-count adaptor is built" % ad.get("method"))
    mx = st.get("maximum", {}) or {}
    dt_max = mx.get("size") or np.inf
    nmax = mx.get("number")
    nmax = 10 ** 9 if nmax is None else int(nmax)
    tries = int(mx.get("tries", 10))
    stop = tm.get("stop")
    stop = np.inf if stop is None else float(stop)
    assert np.isfinite(stop) or nmax < 10 ** 9, "the input gives neither a stop time nor a maximum number of steps"
    reduction, amplification = ad.get("reduction", 0.2), ad.get("amplification", 2.0)
    its_min, its_max = ad.get("minimum", 5), ad.get("maximum", 8)
    nc, nph = output._EOS[p.eos]
    nsrc = len(p.source_cells)
    y = p.y.copy()
    t = float(tm.get("start", 0.0) or 0.0)
    times, fluids, sources = [], [], []

    def record():
        fl = np.asarray(sim.fluid()).copy()
        times.append(t)
        fluids.append(fl)
        if nsrc:
            r = _source_rates(sim, nsrc)
            comp = np.where(r > 0, p.source_injection_components, p.source_production_components)
            h = [p.source_enthalpies[k] if r[k] > 0 else production_enthalpy(fl[int(p.source_cells[k])], len(nc), len(nph))
                 for k in range(nsrc)]
            sources.append(np.stack([comp, r, h], 1))

    err, L0 = sim.lhs(y)
    assert err == 0, "the initial state is outside the range of the thermodynamics"
    apply_controls(p, sim, t, t + min(sizes[0], dt_max))
    sim.residual(y, L0, min(sizes[0], dt_max))     # one function evaluation: the source rates of the initial state
    nt = len(p.tracers)
    if nt:
        x, xb = setup_tracers(p, sim)
        err, L0 = sim.lhs(y)
        al = sim.tracer_balances()
        if tracer_history is not None:
            tracer_history.append(x.reshape(-1, nt).copy())
    record()
    k, dt = 0, sizes[0]
    while t < stop * (1.0 - 1e-12) and k < nmax:
        if k < len(sizes):
            dt = sizes[k]
        elif not adapt:
            dt = sizes[-1]
        dt = min(dt, stop - t, dt_max)
        err, L0 = sim.lhs(y)
        assert err == 0
        sim.pre_timestep()
        y0 = y.copy()
        for attempt in range(tries):
            rock = ingest.rock_at(p, t + dt)            # pre_try_timestep: rock tables at the time the step ends at; L0
            if rock is not None:                        # stays the balance of the last step (src/timestepper.F90:2333)
                assert sim.set_rock(rock) == 0
            apply_controls(p, sim, t, t + dt)
            res = sim.newton_solve(y, L0, dt, opts)
            if res.reason > 0:
                break
            dt *= reduction
            y[:] = y0
            sim.pre_retry_timestep()
            if log:
                log("step %d at t = %.6g: Newton failed (reason %d), step size cut to %.6g" % (k + 1, t, res.reason, dt))
        if res.reason <= 0:
            raise RuntimeError("time step %d at t = %g did not converge after %d tries" % (k + 1, t, tries))
        if nt:
            # the flow step left the fluxes and fluid of the new state behind: the tracer system is built from them
            sim.set_tracer_injection(ingest.tracer_rates_at(p, t, t + dt))
            x, al, reason, _ = sim.tracer_solve(dt, al, x, xb, opts=tracer_ksp)
            if reason <= 0:
                raise RuntimeError("the tracer solve of time step %d did not converge (reason %d)" % (k + 1, reason))
            if tracer_history is not None:
                tracer_history.append(np.asarray(x).reshape(-1, nt).copy())
        t += dt
        k += 1
        err, L0 = sim.lhs(y)           # fluid records and source rates of the new state
        assert err == 0
        record()
        if log:
            log("step %d: t = %.6g, size %.6g, %d Newton iterations" % (k, t, dt, res.iterations))
        if adapt:
            if res.iterations < its_min:
                dt *= amplification
            elif res.iterations > its_max:
                dt *= reduction
    return np.array(times), fluids, sources, y


def run_file(path, output_path=None, sim=None, opts=None, log=None):
    """ingest the deck, run it on the CUDA path (or on `sim`), write the output file; returns its path"""
    from . import flow
    p = ingest.load(path, mod=flow)
    m = p.mesh
    if sim is None:
        sim = flow.FlowSimulation(p.params, m)
        if len(p.boundary_region):
            assert sim.set_boundaries(m.boundary["ghost_cells"], m.boundary["interior_cells"], p.boundary_primary, p.boundary_region) == 0
        if len(p.source_cells):
            assert sim.set_sources(p.source_cells, p.source_components, p.source_rates, p.source_enthalpies) == 0
            sim.set_source_components(p.source_injection_components, p.source_production_components)
        assert sim.fluid_init(p.y, p.region) == 0
    if opts is None:
        nl = (p.time.get("step", {}).get("solver", {}) or {}).get("nonlinear", {}) or {}
        tol = (nl.get("tolerance", {}) or {}).get("function", {}) or {}
        opts = flow.newton_opts(max_iterations=(nl.get("maximum", {}) or {}).get("iterations") or 8,
                                rel_tol=tol.get("relative") or 1e-5, abs_tol=tol.get("absolute") or 1.0,
                                pc_type=flow.PC_BJACOBI_ILU0, ksp=flow.ksp_opts(type=flow.KSP_BCGS))
    tracers = [] if p.tracers else None
    tk = flow.ksp_opts(type=flow.KSP_BCGS, rtol=1e-10) if p.tracers else None
    times, fluids, sources, y = run(p, sim, opts=opts, log=log, tracer_ksp=tk, tracer_history=tracers)
    out = p.doc.get("output") or {}
    if output_path is None:
        name = out.get("filename") if isinstance(out, dict) else None
        output_path = os.path.join(os.path.dirname(os.path.abspath(path)), name or os.path.splitext(os.path.basename(path))[0] + ".h5")
    write_results(p, output_path, times, fluids, sources, tracers)
    return output_path


def write_results(p, output_path, times, fluids, sources, tracers=None):
    """the output file of a run: the states the input's "output" value asks for ("initial", "frequency", "final";
    src/flow_simulation.F90 output setup) in the reference's layout"""
    out = p.doc.get("output") or {}
    m = p.mesh
    keep = np.ones(len(times), bool)
    if isinstance(out, dict):
        freq = out.get("frequency", 1)
        keep[1:] = False
        if freq:
            keep[freq::freq] = True
        keep[0] = bool(out.get("initial", True))
        if out.get("final", True):
            keep[-1] = True
    idx = np.nonzero(keep)[0]
    # fluid fields: the ones a restart needs plus those the input asks for ("output": {"fields": {"fluid": [...]}})
    fields = list(output.REQUIRED[p.eos])
    asked = (out.get("fields") or {}).get("fluid", []) if isinstance(out, dict) else []
    fields += [f for f in ([asked] if isinstance(asked, str) else asked) if f not in fields]
    output.write_output(output_path, m, p.eos, times[idx], [fluids[i] for i in idx],
                        source_cells=p.source_cells if len(p.source_cells) else None,
                        source_history=[sources[i] for i in idx] if len(p.source_cells) else None, fields=fields,
                        tracer_names=[t.get("name", "tracer_%d" % k) for k, t in enumerate(p.tracers)] if tracers else None,
                        tracer_history=[tracers[i] for i in idx] if tracers else None)


def describe(path):
    """what ingest makes of a deck, without touching the GPU: a dictionary for `--info`"""
    p = ingest.load(path)
    m = p.mesh
    n0 = getattr(m, "minc_cells", m.ninterior)
    tm = p.time or {}
    return {"input": os.path.abspath(path), "eos": p.eos, "dimension": int(m.dim), "cells": int(m.ninterior),
            "original_cells": int(n0), "minc_levels": int(m.minc_levels), "faces": int(m.nface),
            "boundary_faces": int(len(p.boundary_region)), "gravity": [float(g) for g in m.gravity],
            "initial": "given" if p.y is not None else "restart file not found",
            "sources": int(len(p.source_cells)), "source_controls": len(p.source_controls),
            "separators": len(p.source_separators), "recharge": len(p.source_recharge),
            "pressure_tables": len(p.source_pressure_tables), "rate_tables": len(p.source_tables),
            "tracers": [t.get("name") for t in p.tracers], "stop": tm.get("stop"),
            "step_method": (tm.get("step") or {}).get("method", "beuler")}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("input")
    ap.add_argument("-o", "--output")
    ap.add_argument("-q", "--quiet", action="store_true")
    ap.add_argument("--info", action="store_true", help="read the deck, print what was understood and stop (no GPU needed)")
    a = ap.parse_args(argv)
    if a.info:
        print(json.dumps(describe(a.input)))
        return
    path = run_file(a.input, a.output, log=None if a.quiet else lambda s: print(s, file=sys.stderr))
    print(json.dumps({"output": path}))


if __name__ == "__main__":
    main()
