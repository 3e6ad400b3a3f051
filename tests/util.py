"""Shared helpers of the parity tests: build the same problem for the oracle and for the CUDA path."""
import ctypes as C

import numpy as np

from waiwera_b200 import mesh as wmesh

SEED = 20240917


def wb_params_from_oracle(wo, flow, prm):
    """wb_params with the same contents as the oracle's wo_params (independent struct definitions,
    identical field layout)."""
    return flow.make_params(eos={wo.EOS_WE: flow.EOS_WE, wo.EOS_W: flow.EOS_W, wo.EOS_WCE: flow.EOS_WCE}[prm.eos],
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
