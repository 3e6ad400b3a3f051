"""GPU cases written after round 2's measurements (passed on a B200: profiles/r2z_zz_new_cases.log):
new host-side input paths onto kernels that the rest of the suite already covers.  They live in a file that sorts last so that an `-x` run reaches every other test
first.  The CPU halves of these tests (the same decks through the oracle) are in test_minc_production3d.py,
test_benchmarks_from_input.py and test_run.py."""
import json
import os

import numpy as np
import pytest

import test_benchmarks_from_input as B
import test_minc_production3d as P3
import test_run as R
from test_mis_problems import newton_opts
from util import run_input
from waiwera_b200 import ingest


@pytest.mark.gpu
def test_cuda_path_runs_minc_production3d(wo):
    """the base case through the CUDA path, to the reference's own acceptance tolerance (1e-2) and against the oracle run"""
    from waiwera_b200 import flow
    p_ref, hist_ref, y_ref, rates_ref = P3.run_oracle(wo, "base")
    p = ingest.load(os.path.join(B.INP, "minc_3d_base.input.json"), mod=flow)
    m = p.mesh
    sim = flow.FlowSimulation(p.params, m)
    assert sim.set_boundaries(m.boundary["ghost_cells"], m.boundary["interior_cells"], p.boundary_primary, p.boundary_region) == 0
    assert sim.set_sources(p.source_cells, p.source_components, p.source_rates, p.source_enthalpies) == 0
    assert sim.fluid_init(p.y, p.region) == 0
    well = len(p.source_cells) - 1
    rates = []
    hist, y = run_input(p, sim, opts=newton_opts(flow, p), controls=True, well=well,
                        on_step=lambda t, s: rates.append(s.source_rates()[well]))
    err, herr, eh, er = P3.errors("base", hist, np.array(rates))
    assert all(e < 1e-2 for e in err + herr) and eh < 1e-2 and er < 1e-2, (err, herr, eh, er)
    assert len(hist) == len(hist_ref)
    assert np.abs(y - y_ref).max() / np.abs(y_ref).max() < 1e-3
    sim.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["deliv_delw", "deliv_delg_limit"])
def test_cuda_path_runs_separated_limiter_decks(wo, case):
    """the two deliverability decks with limiters on the separated water / steam flow (not among the hand-built cases of
    test_deliverability.py) through the CUDA path: the reference's acceptance tolerances against the listing, and the
    oracle's run"""
    from waiwera_b200 import flow
    p_ref, hist_ref, y_ref, rates_ref = B.run_oracle(wo, case)
    p = ingest.load(os.path.join(B.INP, case + ".input.json"), mod=flow)
    m = p.mesh
    sim = flow.FlowSimulation(p.params, m)
    assert sim.set_boundaries(m.boundary["ghost_cells"], m.boundary["interior_cells"], p.boundary_primary, p.boundary_region) == 0
    assert sim.set_sources(p.source_cells, p.source_components, p.source_rates, p.source_enthalpies) == 0
    assert sim.fluid_init(p.y, p.region) == 0
    rates = []
    hist, y = run_input(p, sim, opts=newton_opts(flow, p), controls=True, on_step=lambda t, s: rates.append(np.array(s.source_rates())))
    err, herr, er = B.errors(case, hist, np.array(rates))
    assert all(e < 5e-3 for e in err) and all(e < 1e-2 for e in herr) and er < 1e-2, (err, herr, er)
    assert len(hist) == len(hist_ref)
    assert np.abs(y - y_ref).max() / np.abs(y_ref).max() < 1e-4
    sim.destroy()


@pytest.mark.gpu
def test_command_line_on_the_cuda_path(tmp_path):
    """python -m waiwera_b200.run deck.json -o out.h5, as a user would type it"""
    import shutil
    import subprocess
    import sys
    case = "deliv_delg_flow"
    for fn in (case + ".input.json", "gdeliv.ascii.msh"):
        shutil.copy(os.path.join(B.INP, fn), str(tmp_path / fn))
    out = str(tmp_path / "out.h5")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "waiwera_b200.run", str(tmp_path / (case + ".input.json")), "-o", out, "-q"],
                       cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert json.loads(r.stdout.strip().splitlines()[-1])["output"] == out
    R.check_output_file(case, out, nsteps_expected=len(B.GOLD[case]["times"]) - 1)
