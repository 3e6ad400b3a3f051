"""Reference pressure of a well on deliverability tabulated against the flowing enthalpy or the pressure of its cell
(wb_set_source_pressure_table; "deliverability": {"pressure": {"enthalpy": [[h, P], ...]}}, SRC_PRESSURE_TABLE_COORD_ENTHALPY
in src/source_control.F90:359-403) on the CUDA path: source by source against the oracle (pinned on the reference's known
answer in tests/test_oracle_kat.py, and equal to the device header on the host in tests/test_device_headers_host.py), and
the reference's deliv_delg_pwb_table deck run from its input file against its AUTOUGH2 listing."""
import os

import numpy as np
import pytest

from test_benchmarks_from_input import INP, errors, run_oracle
from test_mis_problems import newton_opts
from util import run_input
from waiwera_b200 import ingest
from waiwera_b200 import mesh as wmesh

TABLES = [[[0.0, 22.0e5], [11.0e5, 20.0e5], [28.0e5, 0.5e5]],
          [[5.0e5, 1.0e5], [6.5e5, 2.0e5], [9.0e5, 4.0e5], [20.0e5, 8.0e5], [26.0e5, 12.0e5], [27.0e5, 13.0e5], [27.5e5, 14.0e5], [29.0e5, 25.0e5]],
          [[15.0e5, 3.0e5]]]


@pytest.mark.gpu
def test_source_rates_with_pressure_tables_match_oracle(wo):
    from waiwera_b200 import flow
    m = wmesh.structured(3, 1, 1, dx=10.0, heterogeneous=False)
    cells = [([30.0e5, 0.4], 4), ([30.0e5, 150.0], 1), ([1.0e5, 150.0], 2)]          # two-phase, liquid, vapour
    primary = np.array([c[0] for c in cells])
    region = np.array([c[1] for c in cells], np.int32)
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1)
    cases = [dict(cell=cell, table=t, coord=coord, step=step, direction=direction)
             for cell in range(3) for t in TABLES for coord in (0, 1) for step in (0, 1) for direction in (0, 1)]
    n = len(cases)
    with_table = list(range(0, n, 2)) + [1]          # the other sources keep the fixed reference pressure
    f = wo.Flow(wo.make_params(eos=wo.EOS_WE), m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    sim = flow.FlowSimulation(flow.make_params(eos=flow.EOS_WE), m)
    rates = {}
    for name, obj in (("oracle", f), ("cuda", sim)):
        assert obj.fluid_init(y, region) == 0
        assert not obj.set_sources([c["cell"] for c in cases], [0] * n, [-1.0] * n, [0.0] * n)
        assert not obj.set_source_controls(list(range(n)), [2e-12] * n, [7.0e5] * n, [c["direction"] for c in cases], [0.0] * n)
        assert not obj.set_source_pressure_table(with_table, [cases[k]["table"] for k in with_table],
                                                 [cases[k]["coord"] for k in with_table], [cases[k]["step"] for k in with_table])
        e, L0 = obj.lhs(y)
        assert e == 0 and obj.residual(y, L0, 1.0e3)[0] == 0
        rates[name] = np.asarray(f.source_rates(n) if name == "oracle" else sim.source_rates())
    assert np.abs(rates["cuda"] - rates["oracle"]).max() <= 1e-12 * np.abs(rates["oracle"]).max()
    assert np.allclose(rates["cuda"], rates["oracle"], rtol=1e-9, atol=0.0)
    assert (rates["oracle"] != 0.0).sum() > n // 2
    # the tables go when asked, the fixed reference pressure is back
    assert not sim.set_source_pressure_table([], [])
    e, L0 = sim.lhs(y)
    assert sim.residual(y, L0, 1.0e3)[0] == 0
    plain = sim.source_rates()
    assert not f.set_source_pressure_table([], [])
    e, L0 = f.lhs(y)
    f.residual(y, L0, 1.0e3)
    assert np.allclose(plain, f.source_rates(n), rtol=1e-9, atol=0.0) and np.abs(plain - rates["cuda"]).max() > 1e-3
    sim.destroy()


@pytest.mark.gpu
def test_cuda_path_runs_the_pwb_table_deck(wo):
    from waiwera_b200 import flow
    case = "deliv_delg_pwb_table"
    p_ref, hist_ref, y_ref, rates_ref = run_oracle(wo, case)
    p = ingest.load(os.path.join(INP, case + ".input.json"), mod=flow)
    m = p.mesh
    sim = flow.FlowSimulation(p.params, m)
    assert sim.set_boundaries(m.boundary["ghost_cells"], m.boundary["interior_cells"], p.boundary_primary, p.boundary_region) == 0
    assert sim.set_sources(p.source_cells, p.source_components, p.source_rates, p.source_enthalpies) == 0
    assert sim.fluid_init(p.y, p.region) == 0
    rates = []
    hist, y = run_input(p, sim, opts=newton_opts(flow, p), controls=True, on_step=lambda t, s: rates.append(np.array(s.source_rates())))
    err, herr, er = errors(case, hist, np.array(rates))
    assert all(e < 5e-3 for e in err) and all(e < 1e-2 for e in herr) and er < 1e-2, (err, herr, er)
    assert len(hist) == len(hist_ref)
    assert np.abs(y - y_ref).max() / np.abs(y_ref).max() < 1e-4
    sim.destroy()
