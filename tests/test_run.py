"""waiwera_b200/run.py: a deck in, an output file out (`python -m waiwera_b200.run deck.json`) -- the few host lines
around the Newton step that the reference keeps in timestepper.F90 / flow_simulation.F90.  On CPU the checker stands in for
the engine (the driver only uses the method names of flow.FlowSimulation); on the GPU the command runs as a user would
type it.  Checked against the AUTOUGH2 listings of the decks (tests/golden/benchmarks_from_input.json) through the output
file it writes, read back with h5lite as a restart or a CREDO script would."""
import json
import os

import numpy as np
import pytest

from test_benchmarks_from_input import GOLD, INP
from test_mis_problems import newton_opts
from util import OracleSim
from waiwera_b200 import h5lite, ingest, output, run


def check_output_file(case, path, nsteps_expected=None):
    g = GOLD[case]
    h = h5lite.H5File(path)
    t = h["time"].reshape(-1)
    assert t[0] == 0.0 and abs(t[-1] - g["times"][-1]) <= 1e-4 * g["times"][-1]
    if nsteps_expected is not None:
        assert len(t) == nsteps_expected + 1
    final = np.array(g["final"])
    n = len(final)
    rel = lambda a, b: np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)
    assert rel(h["cell_fields/fluid_pressure"][-1][:n], final[:, 0]) < 2e-4
    assert rel(h["cell_fields/fluid_temperature"][-1][:n], final[:, 1]) < 2e-4
    assert np.abs(h["cell_fields/fluid_vapour_saturation"][-1][:n] - final[:, 2]).max() < 2e-4
    # source fields: rate and enthalpy histories of the listing (its first table is the state after the first step or t = 0)
    st = np.array(g["source_times"])
    s2 = st > 0
    rate = h["source_fields/source_rate"]
    assert rate.shape == (len(t), len(g["rate"][0]))
    for k in range(rate.shape[1]):
        assert rel(np.interp(st[s2], t[1:], rate[1:, k]), np.array(g["rate"])[s2, k]) < 1e-4
        assert rel(np.interp(st[s2], t[1:], h["source_fields/source_enthalpy"][1:, k]), np.array(g["enthalpy"])[s2, k]) < 1e-4
    assert np.array_equal(h["cell_index"].reshape(-1), np.arange(n))
    # the file restarts a run
    prim, region, time = output.read_restart(path, "we")
    assert time == t[-1] and np.allclose(prim[:, 0], final[:, 0], rtol=1e-3)
    return len(t) - 1


@pytest.mark.parametrize("case", ["deliv_delg_flow", "deliv_delw", "minc_1d_100"])
def test_driver_with_the_checker_as_engine(wo, tmp_path, case):
    p = ingest.load(os.path.join(INP, case + ".input.json"), mod=wo)
    p.doc["output"] = {"initial": True, "frequency": 1, "final": True, "fields": {"fluid": ["liquid_saturation", "vapour_density"]}}
    m = p.mesh
    f = wo.Flow(p.params, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    for k in range(len(p.boundary_region)):
        assert f.set_boundary(int(m.boundary["ghost_cells"][k]), int(m.boundary["interior_cells"][k]),
                              p.boundary_primary[k], int(p.boundary_region[k])) == 0
    f.set_sources(p.source_cells, p.source_components, p.source_rates, p.source_enthalpies)
    f.set_source_components(p.source_injection_components, p.source_production_components)
    assert f.fluid_init(p.y, p.region) == 0
    sim = OracleSim(wo, f, newton_opts(wo, p))
    log = []
    times, fluids, sources, y = run.run(p, sim, log=log.append)
    sim.destroy()
    assert len(times) == len(fluids) == len(sources) and len(log) >= len(times) - 1
    path = str(tmp_path / "out.h5")
    run.write_results(p, path, times, fluids, sources)
    check_output_file(case, path)
    h = h5lite.H5File(path)          # the extra fluid fields the "output" value asks for
    assert np.allclose(h["cell_fields/fluid_liquid_saturation"] + h["cell_fields/fluid_vapour_saturation"], 1.0, atol=1e-14)
    assert h.shape("cell_fields/fluid_vapour_density") == h.shape("cell_fields/fluid_pressure")
    if case.startswith("minc"):          # flow_simulation_output_minc_data: level and parent of every cell
        h = h5lite.H5File(path)
        n0 = m.minc_cells
        assert h["minc/level"].reshape(-1).tolist() == [0] * n0 + [1] * (m.ninterior - n0)
        assert h["minc/parent"].reshape(-1).tolist() == list(range(n0)) + list(m.minc_zone)
    else:
        assert "minc/level" not in h5lite.H5File(path)
    # "frequency": 0 keeps the initial and the final state only
    p.doc["output"] = {"frequency": 0}
    run.write_results(p, path, times, fluids, sources)
    assert h5lite.H5File(path)["time"].reshape(-1).tolist() == [times[0], times[-1]]



class OracleTracerEngine(OracleSim):
    """OracleSim + the tracer calls of flow.FlowSimulation (the checker keeps the boundary ghost rows inside its tracer
    system, the engine takes their mass fractions as a separate array)"""

    def set_tracers(self, phases, diffusion=None, decay=None, activation=None):
        self.f.set_tracers(phases, diffusion, decay, activation)
        self.A = self.f.tracer_pattern()
        self.nt = len(phases)

    def set_tracer_injection(self, rates):
        self.f.set_tracer_injection(rates)

    def tracer_balances(self):
        self.al_full = self.f.tracer_balances()
        return self.al_full[:self.f.nowned * self.nt].copy()

    def tracer_solve(self, dt, al, x, xb=None, opts=None):
        import ctypes as C
        n = self.f.nowned * self.nt
        x_full = np.concatenate([x, xb]) if xb is not None else np.asarray(x)
        b, al_new = self.f.tracer_setup_linear(self.A, dt, self.al_full, x_full)
        k = self.wo.KspOpts()
        k.type, k.restart, k.maxit, k.rtol, k.atol, k.dtol = self.wo.KSP_BCGS, 30, 10000, 1e-10, 1e-50, 1e5
        pc = self.L.wo_pc_create(self.A, self.wo.PC_BJACOBI_ILU0, None)
        xn = np.zeros(len(b))
        its, rn = C.c_int(), C.c_double()
        reason = self.L.wo_ksp_solve(self.A, pc, C.byref(k), self.wo.dp(b), self.wo.dp(xn), C.byref(its), C.byref(rn))
        self.L.wo_pc_destroy(pc)
        self.al_full = al_new
        return xn[:n], al_new[:n], reason, its.value


def test_driver_runs_a_tracer_deck(wo, tmp_path):
    """test/benchmark/tracer/oned (oned_single_phase.json): restart from the steady-state file the deck names, ten flow
    steps each followed by the tracer solve, Dirichlet tracer boundary -> cell_fields/tracer_tracer of the output file
    against the AUTOUGH2 listing (the reference accepts 1e-3; the listing prints 6 digits)"""
    gold = json.load(open(os.path.join(os.path.dirname(INP), "tracer_oned.json")))["single"]
    p = ingest.load(os.path.join(INP, "oned_single_phase.input.json"), mod=wo)
    m = p.mesh
    assert p.y is not None and len(p.tracers) == 1 and p.boundary_tracer[0, 0] == 0.01
    f = wo.Flow(p.params, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    assert f.set_boundary(int(m.boundary["ghost_cells"][0]), int(m.boundary["interior_cells"][0]), p.boundary_primary[0],
                          int(p.boundary_region[0])) == 0
    f.set_sources(p.source_cells, p.source_components, p.source_rates, p.source_enthalpies)
    assert f.fluid_init(p.y, p.region) == 0
    sim = OracleTracerEngine(wo, f, newton_opts(wo, p))
    tracers = []
    times, fluids, sources, y = run.run(p, sim, tracer_history=tracers)
    sim.destroy()
    assert len(times) == len(tracers) == len(gold["times"]) + 1 and np.allclose(times[1:], gold["times"])
    path = str(tmp_path / "tracer.h5")
    p.doc["output"] = {"initial": True, "frequency": 1, "final": True}
    run.write_results(p, path, times, fluids, sources, tracers)
    h = h5lite.H5File(path)
    X = h["cell_fields/tracer_tracer"]
    assert X.shape == (len(times), 10) and np.all(X[0] == 0.0)
    for k, rows in enumerate(gold["tables"]):
        rows = np.array(rows)[:10]
        assert np.abs(X[k + 1] - rows[:, 4]).max() < 2e-6
        assert np.abs(h["cell_fields/fluid_pressure"][k + 1] - rows[:, 0]).max() / rows[:, 0].max() < 1e-3


def test_info_describes_a_deck_without_a_gpu(capsys):
    run.main([os.path.join(INP, "minc_3d_base.input.json"), "--info"])
    d = json.loads(capsys.readouterr().out)
    assert (d["eos"], d["dimension"], d["cells"], d["original_cells"], d["minc_levels"]) == ("we", 3, 161, 125, 2)
    assert (d["boundary_faces"], d["sources"], d["source_controls"], d["gravity"]) == (25, 26, 1, [0.0, 0.0, -9.8])
    assert d["faces"] == 300 + 36 + 25 and d["initial"] == "given" and d["stop"] == 126100000
    run.main([os.path.join(INP, "deliv_delg_pwb_table.input.json"), "--info"])
    d = json.loads(capsys.readouterr().out)
    assert d["pressure_tables"] == 1 and d["cells"] == 10 and d["tracers"] == []


@pytest.mark.parametrize("case", [c for c in GOLD if not c.startswith("_") and c != "columns"])
def test_driver_reproduces_every_single_well_deck(wo, case):
    """all 14 decks of test_benchmarks_from_input.py through run.run (the driver a user gets) instead of the tests' own
    stepping helper: same agreement with the AUTOUGH2 listings"""
    from test_benchmarks_from_input import errors, tolerance
    p = ingest.load(os.path.join(INP, case + ".input.json"), mod=wo)
    m = p.mesh
    f = wo.Flow(p.params, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    for k in range(len(p.boundary_region)):
        assert f.set_boundary(int(m.boundary["ghost_cells"][k]), int(m.boundary["interior_cells"][k]),
                              p.boundary_primary[k], int(p.boundary_region[k])) == 0
    f.set_sources(p.source_cells, p.source_components, p.source_rates, p.source_enthalpies)
    f.set_source_components(p.source_injection_components, p.source_production_components)
    assert f.fluid_init(p.y, p.region) == 0
    sim = OracleSim(wo, f, newton_opts(wo, p))
    times, fluids, sources, y = run.run(p, sim)
    sim.destroy()
    n = m.ninterior
    hist = [(t, np.stack([fl[:n, 0], fl[:n, 1], fl[:n, output.fluid_field_column("we", "vapour_saturation")]], 1), 0.0)
            for t, fl in zip(times[1:], fluids[1:])]
    rates = np.array([s[:, 1] for s in sources[1:]])
    err, herr, er = errors(case, hist, rates)
    tl = tolerance(case)
    assert all(e < tl[0] for e in err) and all(e < tl[1] for e in herr) and er < tl[2], (case, err, herr, er)


@pytest.mark.parametrize("case", ["problem2c", "problem4", "problem5b", "problem6"])
def test_driver_reproduces_mis_problems(wo, case):
    """MIS problems with a flashing front and step cuts (2c), drainage under gravity (4), a rate table with re-injection
    (5b) and the 3-D field with adaptive steps (6) through run.run: the agreement of test_mis_problems.py"""
    import test_mis_problems as T
    p = ingest.load(os.path.join(INP, case + ".input.json"), mod=wo)
    m = p.mesh
    f = wo.Flow(p.params, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    for k in range(len(p.boundary_region)):
        assert f.set_boundary(int(m.boundary["ghost_cells"][k]), int(m.boundary["interior_cells"][k]),
                              p.boundary_primary[k], int(p.boundary_region[k])) == 0
    f.set_sources(p.source_cells, p.source_components, p.source_rates, p.source_enthalpies)
    f.set_source_components(p.source_injection_components, p.source_production_components)
    assert f.fluid_init(p.y, p.region) == 0
    sim = OracleSim(wo, f, newton_opts(wo, p))
    times, fluids, sources, y = run.run(p, sim)
    sim.destroy()
    n = m.ninterior
    sv = output.fluid_field_column("we", "vapour_saturation")
    cell = int(p.source_cells[0])
    hist = [(t, np.stack([fl[:n, 0], fl[:n, 1], fl[:n, sv]], 1), run.production_enthalpy(fl[cell], 1, 2))
            for t, fl in zip(times[1:], fluids[1:])]
    etab, ehist, eh = T.compare(case, hist)
    tol = T.TOL[case]
    assert all(e < tl for e, tl in zip(etab, tol[:3])) and all(e < tl for e, tl in zip(ehist, tol[:3])) and eh < tol[3], (case, etab, ehist, eh)


@pytest.mark.parametrize("case", ["infiltration", "heat_pipe"])
def test_driver_reproduces_the_air_water_benchmarks(wo, case):
    """eos wae decks (ncg/infiltration, ncg/heat_pipe) through run.run: the agreement of test_wae_benchmarks.py"""
    import test_wae_benchmarks as W
    from util import wge_fields
    p = ingest.load(os.path.join(INP, case + ".input.json"), mod=wo)
    m = p.mesh
    f = wo.Flow(p.params, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    for k in range(len(p.boundary_region)):
        assert f.set_boundary(int(m.boundary["ghost_cells"][k]), int(m.boundary["interior_cells"][k]),
                              p.boundary_primary[k], int(p.boundary_region[k])) == 0
    f.set_sources(p.source_cells, p.source_components, p.source_rates, p.source_enthalpies)
    assert f.fluid_init(p.y, p.region) == 0
    sim = OracleSim(wo, f, W.newton_opts(wo, p))
    times, fluids, sources, y = run.run(p, sim)
    regions = f.regions()[:m.ninterior].copy()
    sim.destroy()
    W.check(case, [(t, wge_fields(fl, m.ninterior), 0.0) for t, fl in zip(times[1:], fluids[1:])], regions)


def _oracle_engine(wo, p):
    m = p.mesh
    f = wo.Flow(p.params, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    for k in range(len(p.boundary_region)):
        assert f.set_boundary(int(m.boundary["ghost_cells"][k]), int(m.boundary["interior_cells"][k]),
                              p.boundary_primary[k], int(p.boundary_region[k])) == 0
    f.set_sources(p.source_cells, p.source_components, p.source_rates, p.source_enthalpies)
    f.set_source_components(p.source_injection_components, p.source_production_components)
    assert f.fluid_init(p.y, p.region) == 0
    return OracleSim(wo, f, newton_opts(wo, p))


def test_driver_updates_rock_tables_before_every_step(wo, tmp_path):
    """Rock permeabilities / porosities given as tables in time (src/rock_control.F90, src/rock_setup.F90:383-463):
    run.run sets the records of the time a step ends at before the step is tried (pre_try_timestep,
    src/timestepper.F90:2333, src/flow_simulation.F90:2040-2089) while the balance of the last step keeps the porosity
    it was computed with; checked against the same steps taken by hand, and against the run without tables"""
    from rock_control_deck import write_deck
    p = ingest.load(write_deck(tmp_path), mod=wo)
    assert len(p.rock_controls) == 3
    sim = _oracle_engine(wo, p)
    times, fluids, sources, y = run.run(p, sim)
    sim.destroy()
    assert len(times) == 13
    # by hand
    sim = _oracle_engine(wo, p)
    yh = p.y.copy()
    sizes = p.time["step"]["size"]
    t = 0.0
    for k in range(12):
        e, L0 = sim.lhs(yh)
        assert e == 0
        sim.pre_timestep()
        assert sim.set_rock(ingest.rock_at(p, t + sizes[k])) == 0
        run.apply_controls(p, sim, t, t + sizes[k])
        assert sim.newton_solve(yh, L0, sizes[k]).reason > 0
        t += sizes[k]
    sim.destroy()
    assert abs(times[-1] - t) <= 1e-9 * t and np.array_equal(y, yh)
    # the tables matter: the same deck with the rock of the start time throughout
    p0 = ingest.load(write_deck(tmp_path), mod=wo)
    p0.rock_controls = []
    sim = _oracle_engine(wo, p0)
    _, _, _, y0 = run.run(p0, sim)
    sim.destroy()
    assert np.abs(y - y0).max() / np.abs(y0).max() > 1e-4


def test_checker_set_rock_equals_a_fresh_mesh(wo):
    """the checker's own set_rock (what the GPU test of wb_set_rock compares with): residual after replacing the rock
    records == residual of an engine built on the new records, bit for bit"""
    from util import make_problem, oracle_flow
    m, y, region, prm = make_problem(wo, dims=(5, 4, 6), thermo=0, two_phase_layers=2, top_boundary=False)
    rng = np.random.default_rng(5)
    rock2 = m.rock.copy()
    rock2[:, 0:3] *= 10.0 ** rng.uniform(-0.5, 0.5, (m.ncell, 3))
    rock2[:, 5] = rng.uniform(0.05, 0.3, m.ncell)
    a = oracle_flow(wo, m, prm, y, region)
    _, L0 = a.lhs(y)
    assert a.set_rock(rock2[:m.ninterior]) == 0
    m.rock = rock2
    b = oracle_flow(wo, m, prm, y, region)
    ra, rb = a.residual(y * 1.00001, L0, 1.0e5), b.residual(y * 1.00001, L0, 1.0e5)
    assert ra[0] == rb[0] == 0 and np.array_equal(ra[3], rb[3]) and np.abs(rb[3]).max() > 0
    assert np.array_equal(a.lhs(y)[1], b.lhs(y)[1]) and not np.array_equal(a.lhs(y)[1], L0)
