"""wb_set_rock on the CUDA path: rock records of the interior cells replaced between time steps (time-dependent
permeability / porosity, flow_simulation_update_rock_properties src/flow_simulation.F90:2051-2089 with the table controls
of src/rock_control.F90:49-116).  Written after round 2's GPU budget was spent: **not yet run on a B200** (DESIGN.md
section 6b) -- the file sorts last so that an `-x` run reaches every other test first.  The entry is host code (the
re-upload sequence of wb_set_boundaries) in front of kernels the rest of the suite covers; the CPU halves of these tests
are test_run.py::test_driver_updates_rock_tables_before_every_step / test_checker_set_rock_equals_a_fresh_mesh and
test_ingest.py::test_rock_controls_known_answers (the reference's rock_control_test.F90 known answers)."""
import numpy as np
import pytest

from test_gpu_flow import oracle_jacobian
from test_mis_problems import newton_opts
from util import OracleSim, gpu_flow, make_problem, oracle_flow, relerr
from waiwera_b200 import ingest, run
from waiwera_b200._lib import WbError


@pytest.mark.gpu
def test_set_rock_matches_oracle(wo):
    """lhs, residual and FD Jacobian after wb_set_rock == the oracle after its set_rock, on a mesh with Dirichlet
    boundary ghost cells (which keep the records they copied at set-up in both); a bad record is refused"""
    from waiwera_b200 import flow
    m, y, region, prm = make_problem(wo, dims=(6, 5, 8), thermo=0, two_phase_layers=2, top_boundary=True)     # the smoke() problem
    ref = oracle_flow(wo, m, prm, y, region)
    sim = gpu_flow(wo, flow, m, prm, y, region)
    _, L0 = ref.lhs(y)
    e, L0g = sim.lhs(y)
    assert e == 0 and relerr(L0g, L0) < 1e-13
    rng = np.random.default_rng(11)
    rock2 = np.array(m.rock[:m.ninterior], float)
    rock2[:, 0:3] *= 10.0 ** rng.uniform(-0.5, 0.5, (m.ninterior, 3))
    rock2[:, 5] = rng.uniform(0.05, 0.3, m.ninterior)
    assert ref.set_rock(rock2) == 0 and sim.set_rock(rock2) == 0
    y2 = y * (1 + 1e-4 * rng.uniform(-1, 1, len(y)))
    dt = 1.0e6
    e0, _, _, r0 = ref.residual(y2, L0, dt)
    e1, _, _, r1 = sim.residual(y2, L0, dt)
    assert e0 == e1 == 0 and relerr(r1, r0) < 1e-10, relerr(r1, r0)
    A, rowptr, colidx, val, F0 = oracle_jacobian(wo, ref, y2, L0, dt)
    assert sim.jacobian(y2, L0, dt) == 0
    J = sim.jacobian_values()
    rows = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
    rowmax = np.zeros((len(rowptr) - 1, 2))
    for ii in range(2):
        np.maximum.at(rowmax[:, ii], rows, np.abs(val[:, [ii, 2 + ii]]).max(axis=1))
    scale = np.stack([rowmax[rows, 0], rowmax[rows, 1], rowmax[rows, 0], rowmax[rows, 1]], 1)
    assert (np.abs(J - val) / np.maximum(scale, 1e-300)).max() < 1e-4      # FD noise floor, as in test_gpu_flow.py
    wo.lib().wo_bsr_destroy(A)
    # the new porosity is in the balances, and differs from the old one
    e, L1g = sim.lhs(y)
    assert e == 0 and relerr(L1g, ref.lhs(y)[1]) < 1e-13 and relerr(L1g, L0) > 1e-3
    # putting the old records back restores the first residual
    assert ref.set_rock(m.rock[:m.ninterior]) == 0 and sim.set_rock(m.rock[:m.ninterior]) == 0
    assert relerr(sim.residual(y2, L0, dt)[3], ref.residual(y2, L0, dt)[3]) < 1e-10
    bad = rock2.copy()
    bad[3, 5] = 1.5
    with pytest.raises(WbError):
        sim.set_rock(bad)
    sim.destroy()


@pytest.mark.gpu
def test_cuda_path_runs_a_deck_with_rock_tables(wo, tmp_path):
    """run.run on the CUDA path == run.run on the checker for the deck of tests/rock_control_deck.py"""
    from rock_control_deck import write_deck
    from test_run import _oracle_engine
    from waiwera_b200 import flow
    path = write_deck(tmp_path)
    po = ingest.load(path, mod=wo)
    ref = _oracle_engine(wo, po)
    t0, f0, s0, y0 = run.run(po, ref)
    ref.destroy()
    p = ingest.load(path, mod=flow)
    m = p.mesh
    sim = flow.FlowSimulation(p.params, m)
    assert sim.set_boundaries(m.boundary["ghost_cells"], m.boundary["interior_cells"], p.boundary_primary, p.boundary_region) == 0
    assert sim.set_sources(p.source_cells, p.source_components, p.source_rates, p.source_enthalpies) == 0
    sim.set_source_components(p.source_injection_components, p.source_production_components)
    assert sim.fluid_init(p.y, p.region) == 0
    t1, f1, s1, y1 = run.run(p, sim, opts=newton_opts(flow, p))
    sim.destroy()
    assert len(t1) == len(t0) == 13 and np.allclose(t1, t0, rtol=1e-12)
    assert np.abs(y1 - y0).max() / np.abs(y0).max() < 1e-4
