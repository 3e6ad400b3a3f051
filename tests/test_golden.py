"""The reference's own golden VECTORS for the path (tests/golden/reference_vectors.json, extracted from
test/unit/data/flow_simulation/*.h5 by tools/make_golden.py): cell_balances of test_flow_simulation_lhs, the
scaled primary vector and the rock records of test_flow_simulation_init.  The oracle is pinned against them on
the CPU; the CUDA path is compared with the same vectors through the C ABI on the GPU."""
import json
import os

import numpy as np
import pytest

from waiwera_b200 import mesh as wmesh

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))


def lhs_problem():
    """test_lhs.json: 12 cells (4x3, one layer), eos w at 20 degC, P = 2 bar everywhere, rock1 (porosity 0.1)"""
    m = wmesh.structured(4, 3, 1, dx=100.0, heterogeneous=False)
    m.rock[:, 0:3] = (1e-14, 2e-14, 3e-14)
    m.rock[:, 3:5] = 1.5
    m.rock[:, 5], m.rock[:, 6], m.rock[:, 7] = 0.1, 2600.0, 900.0
    y = np.full(12, G["lhs"]["pressure"] / 1.0e6)
    region = np.ones(12, np.int32)
    return m, y, region


def test_oracle_lhs_matches_reference_golden_file(wo):
    """flow_simulation_test.F90:126-158 (vec_diff_test against lhs.h5)"""
    m, y, region = lhs_problem()
    prm = wo.make_params(eos=wo.EOS_W, thermo=wo.THERMO_IAPWS, eos_w_temperature=G["lhs"]["temperature"])
    f = wo.Flow(prm, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    assert f.fluid_init(y, region) == 0
    err, lhs = f.lhs(y)
    assert err == 0
    gold = np.array(G["lhs"]["values"])
    assert np.abs(lhs - gold).max() / gold.max() < 1e-14, (lhs, gold)


def test_scaled_primary_matches_reference_golden_file():
    """flow_simulation_test.F90:116 (primary.h5): the solution vector holds primaries scaled by eos%scale"""
    prim = np.full((12, 1), G["primary_scaled"]["pressure"])
    y = wmesh.scale_primaries(prim, np.ones(12, np.int32))
    assert np.array_equal(np.asarray(y).reshape(-1), np.array(G["primary_scaled"]["values"]))


def test_rock_record_layout_matches_reference_golden_file():
    """flow_simulation_test.F90:118 (rock.h5): 8 doubles per cell in the order the mesh generator emits"""
    rec = np.array(G["rock"]["values"]).reshape(12, 8)
    # rock2 (cells 0-3) / rock1 (cells 4-11) of test_init.json: wet/dry conductivity, porosity, density, specific heat
    assert np.array_equal(rec[0, 3:], [2.4, 1.4, 0.08, 2500.0, 890.0])
    assert np.array_equal(rec[11, 3:], [2.5, 1.5, 0.1, 2600.0, 900.0])
    assert np.array_equal(rec[11, :3], [1e-14, 2e-14, 3e-14])
    d = wmesh.default_rock(1, None, heterogeneous=False)[0]
    # generator columns: permeability(3), wet, dry conductivity, porosity, density, specific heat
    assert d[3] == d[4] == 2.5 and d[5] == 0.1 and d[6] == 2200.0 and d[7] == 1000.0


@pytest.mark.gpu
def test_cuda_lhs_matches_reference_golden_file():
    """the same golden vector through wb_create / wb_fluid_init / wb_pre_eval / wb_cell_balances"""
    from waiwera_b200 import flow
    m, y, region = lhs_problem()
    sim = flow.FlowSimulation(flow.make_params(eos=flow.EOS_W, thermo=flow.THERMO_IAPWS,
                                               eos_w_temperature=G["lhs"]["temperature"]), m)
    assert sim.fluid_init(y, region) == 0
    err, lhs = sim.lhs(y)
    assert err == 0
    gold = np.array(G["lhs"]["values"])
    assert np.abs(lhs - gold).max() / gold.max() < 1e-14, (lhs, gold)
    sim.destroy()
