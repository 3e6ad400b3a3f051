"""The device physics headers, compiled for the host, against the CPU oracle.

tests/hostcheck/hostcheck.cpp wraps waiwera_b200/csrc/wb_eos.cuh + wb_thermo.cuh +
wb_iapws_gen.cuh (the exact source the CUDA kernels inline) behind a few C entry
points and is compiled here with g++.  This is a CPU-side check of the source
only -- the product never loads it and has no CPU path; the GPU parity tests
proper are tests/test_gpu_*.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
LIB = os.path.join(HERE, "hostcheck", "_hostcheck.so")
SEED = 20240917


@pytest.fixture(scope="module")
def hc(wo):
    deps = [SRC] + [os.path.join(HERE, "..", "waiwera_b200", "csrc", f)
                    for f in ("wb_eos.cuh", "wb_thermo.cuh", "wb_iapws_gen.cuh", "wb_state.cuh", "wb_tracer.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-o", LIB, SRC])
    L = C.CDLL(LIB)
    d, i, dp, ip = C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.hc_region_properties.argtypes = [i, i, i, d, d, dp]
    L.hc_region_viscosity.restype = d
    L.hc_region_viscosity.argtypes = [i, i, d, d, d]
    L.hc_sat_pressure.argtypes = [i, d, dp]
    L.hc_sat_temperature.argtypes = [i, d, dp]
    L.hc_relperm.argtypes = [C.c_void_p, d, dp]
    L.hc_cappress.restype = d
    L.hc_cappress.argtypes = [C.c_void_p, d, d]
    L.hc_we_fluid.argtypes = [C.c_void_p, dp, i, dp]
    L.hc_we_flux.argtypes = [C.c_void_p, dp, dp, dp, dp, i, dp, i, dp, dp]
    L.hc_we_transition.argtypes = [C.c_void_p, dp, dp, i, d, ip, ip]
    L.hc_wce_fluid.argtypes = [C.c_void_p, dp, i, dp]
    L.hc_wce_flux.argtypes = [C.c_void_p, dp, dp, dp, dp, i, dp, i, dp, dp]
    L.hc_wce_transition.argtypes = [C.c_void_p, dp, dp, i, d, ip, ip, ip]
    L.hc_wce_scale.argtypes = [C.c_void_p, dp, i, dp, dp]
    L.hc_tracer_assemble.argtypes = [C.c_void_p, i, i, i, i, ip, dp, dp, dp, dp, ip, ip, dp, dp, dp, i, ip, ip, dp, ip, dp, dp, dp, dp, i,
                                     d, d, dp, dp, dp, dp, dp, ip, ip, dp, dp, dp]
    L.hc_source_rate.argtypes = [C.c_void_p, dp, i, dp, i, d, d, d, d, i, dp, d, d, dp]
    L.hc_separator_stage.argtypes = [i, d, dp, dp]
    L.hc_source_rate_ptab.argtypes = [C.c_void_p, dp, i, dp, i, d, d, i, dp, dp]
    return L


def close(a, b, tol=1e-13):
    return abs(a - b) <= tol * max(abs(a), abs(b), 1e-300)


@pytest.mark.parametrize("thermo", [0, 1])
def test_region_properties_match_oracle(wo, hc, thermo):
    rng = np.random.default_rng(SEED)
    th = wo.lib().wo_thermo_create(thermo, 0)
    try:
        cases = [(1, rng.uniform(1e5, 90e6, 200), rng.uniform(5, 340, 200)),
                 (2, rng.uniform(1e3, 10e6, 200), rng.uniform(200, 780, 200))]
        if thermo == 0:
            cases.append((3, rng.uniform(200, 600, 100), rng.uniform(360, 500, 100)))  # (rho, t)
        for region, ps, ts in cases:
            for p, t in zip(ps, ts):
                ref, got = np.zeros(2), np.zeros(2)
                e0 = wo.lib().wo_region_properties(th, region, wo.dp(np.array([p, t])), wo.dp(ref))
                e1 = hc.hc_region_properties(thermo, 0, region, p, t, wo.dp(got))
                assert e0 == e1
                if e0 == 0:
                    # same operation order, FMA contraction off on both sides: bit-identical
                    assert ref[0] == got[0] and ref[1] == got[1], (region, p, t, ref, got)
                    if region != 3:
                        v0 = wo.lib().wo_region_viscosity(th, region, t, p, ref[0])
                        v1 = hc.hc_region_viscosity(thermo, region, t, p, ref[0])
                        assert v0 == v1
        # out-of-range error returns
        for region, p, t in [(1, 20e6, 360.0), (1, 101e6, 60.0), (2, 1e5, 801.0)]:
            ref, got = np.zeros(2), np.zeros(2)
            assert wo.lib().wo_region_properties(th, region, wo.dp(np.array([p, t])), wo.dp(ref)) == \
                hc.hc_region_properties(thermo, 0, region, p, t, wo.dp(got)) == 1
    finally:
        wo.lib().wo_thermo_destroy(th)


@pytest.mark.parametrize("thermo", [0, 1])
def test_saturation_matches_oracle(wo, hc, thermo):
    rng = np.random.default_rng(SEED + 1)
    th = wo.lib().wo_thermo_create(thermo, 0)
    try:
        for t in list(rng.uniform(1.0, 373.0, 200)) + [0.5, 374.1, 380.0]:
            a, b = C.c_double(), C.c_double()
            assert wo.lib().wo_saturation_pressure(th, t, C.byref(a)) == hc.hc_sat_pressure(thermo, t, C.byref(b))
            assert a.value == b.value
        for p in list(rng.uniform(700.0, 22.0e6, 200)) + [500.0, 23e6]:
            a, b = C.c_double(), C.c_double()
            assert wo.lib().wo_saturation_temperature(th, p, C.byref(a)) == hc.hc_sat_temperature(thermo, p, C.byref(b))
            assert a.value == b.value
    finally:
        wo.lib().wo_thermo_destroy(th)


def curve_cases(wo):
    return [
        (wo.make_relperm("linear"), wo.make_cappress("zero")),
        (wo.make_relperm("linear", liquid=(0.1, 0.9), vapour=(0.2, 0.8)), wo.make_cappress("linear", saturation_limits=(0.1, 0.9), pressure=0.2e5)),
        (wo.make_relperm("linear", liquid=(0.0, 0.0), vapour=(0.0, 0.0)), wo.make_cappress("zero")),
        (wo.make_relperm("corey", slr=0.3, ssr=0.05), wo.make_cappress("van_genuchten", P0=0.125e5, lambda_=0.45, slr=1e-3, sls=1.0, Pmax=1e6)),
        (wo.make_relperm("grant", slr=0.3, ssr=0.1), wo.make_cappress("van_genuchten")),
        (wo.make_relperm("pickens", power=2.5), wo.make_cappress("zero")),
        (wo.make_relperm("fully_mobile"), wo.make_cappress("zero")),
        (wo.make_relperm("van_genuchten", lambda_=0.45, slr=0.1, sls=0.95), wo.make_cappress("zero")),
        (wo.make_relperm("van_genuchten", lambda_=0.5, slr=0.1, sls=1.0, ssr=0.1, sum_unity=False), wo.make_cappress("zero")),
        (wo.make_relperm("table", liquid=[(0, 0), (0.3, 0.1), (0.8, 0.7), (1, 1)], vapour=[(0, 0), (0.5, 0.6), (1, 1)]),
         wo.make_cappress("table", pressure=[(0, -1e5), (0.4, -2e4), (1.0, 0.0)])),
    ]


def test_curves_match_oracle(wo, hc):
    rng = np.random.default_rng(SEED + 2)
    sls = list(rng.uniform(-0.1, 1.1, 100)) + [0.0, 1.0, 0.5, 0.9995, 0.3, 0.05]
    for rp, cp in curve_cases(wo):
        for sl in sls:
            if rp.type == wo.RP_PICKENS and sl < 0:
                continue
            a, b = np.zeros(2), np.zeros(2)
            wo.lib().wo_relperm_values(C.byref(rp), sl, wo.dp(a))
            hc.hc_relperm(C.byref(rp), sl, wo.dp(b))
            assert np.array_equal(a, b, equal_nan=True), (rp.type, sl, a, b)
            assert wo.lib().wo_cappress_value(C.byref(cp), sl, 100.0) == hc.hc_cappress(C.byref(cp), sl, 100.0) or \
                np.isnan(hc.hc_cappress(C.byref(cp), sl, 100.0))


def we_cell(wo, thermo, rng, region):
    """random valid (primary, region) for eos_we"""
    th = wo.lib().wo_thermo_create(thermo, 0)
    try:
        if region == 1:
            t = rng.uniform(10, 300)
            ps = C.c_double()
            wo.lib().wo_saturation_pressure(th, t, C.byref(ps))
            return np.array([ps.value + rng.uniform(1e4, 2e7), t])
        if region == 2:
            t = rng.uniform(120, 340)
            ps = C.c_double()
            wo.lib().wo_saturation_pressure(th, t, C.byref(ps))
            return np.array([ps.value * rng.uniform(0.05, 0.95), t])
        return np.array([rng.uniform(1e5, 1.5e7), rng.uniform(0.01, 0.99)])
    finally:
        wo.lib().wo_thermo_destroy(th)


@pytest.mark.parametrize("thermo", [0, 1])
def test_we_fluid_record_matches_oracle(wo, hc, thermo):
    rng = np.random.default_rng(SEED + 3)
    rock = np.array([1e-13, 1e-13, 1e-14, 2.5, 1.5, 0.1, 2200.0, 1000.0])
    for rp, cp in curve_cases(wo)[:5]:
        prm = wo.make_params(eos=wo.EOS_WE, thermo=thermo, relperm=rp, cappress=cp)
        eos = wo.lib().wo_eos_create(C.byref(prm))
        try:
            for region in (1, 2, 4):
                for _ in range(40):
                    primary = we_cell(wo, thermo, rng, region)
                    ref = np.zeros(23)
                    ref[2] = region
                    e0 = wo.lib().wo_eos_bulk_properties(eos, wo.dp(primary), wo.dp(ref))
                    if e0 == 0:
                        e0 = wo.lib().wo_eos_phase_properties(eos, wo.dp(primary), wo.dp(rock), wo.dp(ref))
                    got = np.zeros(23)
                    e1 = hc.hc_we_fluid(C.byref(prm), wo.dp(primary), region, wo.dp(got))
                    assert e0 == e1
                    if e0 == 0:
                        assert np.array_equal(ref, got), (region, primary, ref - got)
        finally:
            wo.lib().wo_eos_destroy(eos)


@pytest.mark.parametrize("thermo", [0, 1])
def test_we_flux_and_balance_match_oracle(wo, hc, thermo):
    rng = np.random.default_rng(SEED + 4)
    rp, cp = curve_cases(wo)[3]
    prm = wo.make_params(eos=wo.EOS_WE, thermo=thermo, relperm=rp, cappress=cp)
    eos = wo.lib().wo_eos_create(C.byref(prm))
    try:
        for trial in range(300):
            r1, r2 = rng.choice([1, 2, 4], 2)
            p1, p2 = we_cell(wo, thermo, rng, r1), we_cell(wo, thermo, rng, r2)
            if trial % 3 == 0:  # nearly equal states: small gradients, either upstream direction
                r2 = r1
                p2 = p1 * (1 + rng.uniform(-1e-4, 1e-4, 2))
            rock1 = np.array([1e-13, 2e-13, 1e-14, 2.5, 1.5, 0.1, 2200.0, 1000.0]) * rng.uniform(0.5, 1.5, 8)
            rock2 = np.array([1e-13, 2e-13, 1e-14, 2.5, 1.5, 0.1, 2200.0, 1000.0]) * rng.uniform(0.5, 1.5, 8)
            d1, d2 = rng.uniform(1, 20, 2)
            g = np.zeros(12)
            g[0], g[1], g[2], g[3] = rng.uniform(1, 100), d1, d2, d1 + d2
            g[7] = rng.choice([0.0, -9.8, 9.8, 3.3])
            g[11] = float(rng.integers(1, 4))
            if trial % 10 == 1:  # Dirichlet boundary face: d2 = 0
                g[2], g[3] = 0.0, d1
            f1, f2 = np.zeros(23), np.zeros(23)
            f1[2], f2[2] = r1, r2
            ok = True
            for pr, fl, rk in ((p1, f1, rock1), (p2, f2, rock2)):
                e = wo.lib().wo_eos_bulk_properties(eos, wo.dp(pr), wo.dp(fl))
                if e == 0:
                    e = wo.lib().wo_eos_phase_properties(eos, wo.dp(pr), wo.dp(rk), wo.dp(fl))
                ok = ok and e == 0
            if not ok:
                continue
            ref = np.zeros(4)
            wo.lib().wo_face_flux(wo.dp(g), wo.dp(rock1), wo.dp(rock2), wo.dp(f1), wo.dp(f2), 1, 2, 2, 2, 0, wo.dp(ref))
            bref = np.zeros(2)
            wo.lib().wo_cell_balance(wo.dp(rock1), wo.dp(f1), 1, 2, 2, wo.dp(bref))
            got, bgot = np.zeros(4), np.zeros(2)
            assert hc.hc_we_flux(C.byref(prm), wo.dp(g), wo.dp(rock1), wo.dp(rock2), wo.dp(p1), int(r1), wo.dp(p2), int(r2),
                                 wo.dp(got), wo.dp(bgot)) == 0
            assert np.array_equal(ref, got), (trial, ref, got)
            assert np.array_equal(bref, bgot)
    finally:
        wo.lib().wo_eos_destroy(eos)


@pytest.mark.parametrize("thermo", [0, 1])
def test_we_transitions_match_oracle(wo, hc, thermo):
    rng = np.random.default_rng(SEED + 5)
    prm = wo.make_params(eos=wo.EOS_WE, thermo=thermo)
    eos = wo.lib().wo_eos_create(C.byref(prm))
    th = wo.lib().wo_thermo_create(thermo, 0)
    ntrans = 0
    try:
        for trial in range(400):
            old_region = int(rng.choice([1, 2, 4]))
            oldp = we_cell(wo, thermo, rng, old_region)
            if old_region == 4:
                newp = oldp + np.array([rng.uniform(-2e5, 2e5), rng.uniform(-1.2, 1.2)])
                if trial % 7 == 0:
                    newp[1] = oldp[1]  # degenerate inverse interpolation -> fallback branch
                    oldp[1] = newp[1] = rng.choice([-0.2, 1.3])
            else:
                ps = C.c_double()
                wo.lib().wo_saturation_pressure(th, oldp[1], C.byref(ps))
                newp = np.array([ps.value * rng.uniform(0.7, 1.3), oldp[1] + rng.uniform(-5, 5)])
            old_fluid = np.zeros(23)
            old_fluid[2] = old_region
            assert wo.lib().wo_eos_bulk_properties(eos, wo.dp(oldp), wo.dp(old_fluid)) == 0
            fluid = old_fluid.copy()
            p_ref = newp.copy()
            tr = C.c_int()
            e0 = wo.lib().wo_eos_transition(eos, wo.dp(oldp), wo.dp(p_ref), wo.dp(old_fluid), wo.dp(fluid), C.byref(tr))
            if e0 == 0:
                ch = C.c_int()
                e0 = wo.lib().wo_eos_check_primary_variables(eos, wo.dp(fluid), wo.dp(p_ref), C.byref(ch))
            p_got = newp.copy()
            reg, tr1 = C.c_int(old_region), C.c_int()
            e1 = hc.hc_we_transition(C.byref(prm), wo.dp(oldp), wo.dp(p_got), old_region, old_fluid[1], C.byref(reg), C.byref(tr1))
            assert (e0 != 0) == (e1 != 0), (trial, e0, e1)
            if e0 == 0:
                assert tr.value == tr1.value
                assert int(round(fluid[2])) == reg.value
                assert np.array_equal(p_ref, p_got), (trial, p_ref, p_got)
                ntrans += tr.value
        assert ntrans > 50
    finally:
        wo.lib().wo_eos_destroy(eos)
        wo.lib().wo_thermo_destroy(th)


# ---------------------------------------------------------------- eos_wce (water + CO2 + energy)

def wce_cell(wo, thermo, rng, region):
    """random valid (P, T | Sv, Pg) for eos_wce: the water partial pressure P - Pg plays the role of the eos_we pressure"""
    pw = we_cell(wo, thermo, rng, region)
    pg = rng.choice([0.0, rng.uniform(1e3, 5e5), rng.uniform(1e5, 3e6)])
    return np.array([pw[0] + pg, pw[1], pg])


@pytest.mark.parametrize("gas", ["co2", "air"])
@pytest.mark.parametrize("thermo", [0, 1])
def test_wce_fluid_record_matches_oracle(wo, hc, thermo, gas):
    """a5: all 26 fields of the eos_wge fluid record (CO2: eos_wce, air: eos_wae), device header vs oracle, bit for bit"""
    rng = np.random.default_rng(SEED + 13)
    rock = np.array([1e-13, 1e-13, 1e-14, 2.5, 1.5, 0.1, 2200.0, 1000.0])
    n_ok = 0
    for rp, cp in curve_cases(wo)[:4]:
        prm = wo.make_params(eos=wo.EOS_WCE if gas == "co2" else wo.EOS_WAE, thermo=thermo, relperm=rp, cappress=cp)
        eos = wo.lib().wo_eos_create(C.byref(prm))
        try:
            for region in (1, 2, 4):
                for _ in range(40):
                    primary = wce_cell(wo, thermo, rng, region)
                    ref = np.zeros(26)
                    ref[2] = region
                    e0 = wo.lib().wo_eos_bulk_properties(eos, wo.dp(primary), wo.dp(ref))
                    if e0 == 0:
                        e0 = wo.lib().wo_eos_phase_properties(eos, wo.dp(primary), wo.dp(rock), wo.dp(ref))
                    got = np.zeros(26)
                    e1 = hc.hc_wce_fluid(C.byref(prm), wo.dp(primary), region, wo.dp(got))
                    assert e0 == e1
                    if e0 == 0:
                        n_ok += 1
                        # pow / log10 come from the same libm on both sides here: bit-identical
                        assert np.array_equal(ref, got), (region, primary, ref - got)
        finally:
            wo.lib().wo_eos_destroy(eos)
    assert n_ok > 300


@pytest.mark.parametrize("fixed_scale", [0.0, 1.0e6])
def test_wce_scaling_matches_oracle(wo, hc, fixed_scale):
    """adaptive (reference default) and fixed partial-pressure scaling, eos_wge.F90:96-110, 639-674"""
    rng = np.random.default_rng(SEED + 14)
    prm = wo.make_params(eos=wo.EOS_WCE, partial_pressure_scale=fixed_scale)
    eos = wo.lib().wo_eos_create(C.byref(prm))
    try:
        for region in (1, 2, 4):
            for _ in range(20):
                primary = wce_cell(wo, 0, rng, region)
                y0, y1, back = np.zeros(3), np.zeros(3), np.zeros(3)
                wo.lib().wo_eos_scale(eos, wo.dp(primary), region, wo.dp(y0))
                hc.hc_wce_scale(C.byref(prm), wo.dp(primary), region, wo.dp(y1), wo.dp(back))
                assert np.array_equal(y0, y1)
                ref_back = np.zeros(3)
                wo.lib().wo_eos_unscale(eos, wo.dp(y0), region, wo.dp(ref_back))
                assert np.array_equal(back, ref_back)
    finally:
        wo.lib().wo_eos_destroy(eos)


@pytest.mark.parametrize("thermo", [0, 1])
def test_wce_flux_and_balance_match_oracle(wo, hc, thermo):
    """a8-a10 with two components: component + energy fluxes, phase fluxes and the 3 balances"""
    rng = np.random.default_rng(SEED + 15)
    rp, cp = curve_cases(wo)[3]
    prm = wo.make_params(eos=wo.EOS_WCE, thermo=thermo, relperm=rp, cappress=cp)
    eos = wo.lib().wo_eos_create(C.byref(prm))
    n_ok = 0
    try:
        for trial in range(300):
            r1, r2 = rng.choice([1, 2, 4], 2)
            p1, p2 = wce_cell(wo, thermo, rng, r1), wce_cell(wo, thermo, rng, r2)
            if trial % 3 == 0:
                r2 = r1
                p2 = p1 * (1 + rng.uniform(-1e-4, 1e-4, 3))
            rock1 = np.array([1e-13, 2e-13, 1e-14, 2.5, 1.5, 0.1, 2200.0, 1000.0]) * rng.uniform(0.5, 1.5, 8)
            rock2 = np.array([1e-13, 2e-13, 1e-14, 2.5, 1.5, 0.1, 2200.0, 1000.0]) * rng.uniform(0.5, 1.5, 8)
            d1, d2 = rng.uniform(1, 20, 2)
            g = np.zeros(12)
            g[0], g[1], g[2], g[3] = rng.uniform(1, 100), d1, d2, d1 + d2
            g[7] = rng.choice([0.0, -9.8, 9.8, 3.3])
            g[11] = float(rng.integers(1, 4))
            if trial % 10 == 1:
                g[2], g[3] = 0.0, d1
            f1, f2 = np.zeros(26), np.zeros(26)
            f1[2], f2[2] = r1, r2
            ok = True
            for pr, fl, rk in ((p1, f1, rock1), (p2, f2, rock2)):
                e = wo.lib().wo_eos_bulk_properties(eos, wo.dp(pr), wo.dp(fl))
                if e == 0:
                    e = wo.lib().wo_eos_phase_properties(eos, wo.dp(pr), wo.dp(rk), wo.dp(fl))
                ok = ok and e == 0
            if not ok:
                continue
            n_ok += 1
            ref = np.zeros(5)
            wo.lib().wo_face_flux(wo.dp(g), wo.dp(rock1), wo.dp(rock2), wo.dp(f1), wo.dp(f2), 2, 3, 2, 2, 0, wo.dp(ref))
            bref = np.zeros(3)
            wo.lib().wo_cell_balance(wo.dp(rock1), wo.dp(f1), 2, 2, 3, wo.dp(bref))
            got, bgot = np.zeros(5), np.zeros(3)
            assert hc.hc_wce_flux(C.byref(prm), wo.dp(g), wo.dp(rock1), wo.dp(rock2), wo.dp(p1), int(r1), wo.dp(p2), int(r2),
                                  wo.dp(got), wo.dp(bgot)) == 0
            assert np.array_equal(ref, got), (trial, ref, got)
            assert np.array_equal(bref, bgot)
        assert n_ok > 200
    finally:
        wo.lib().wo_eos_destroy(eos)


@pytest.mark.parametrize("thermo", [0, 1])
def test_wce_transitions_match_oracle(wo, hc, thermo):
    """a15 for eos_wge: transitions + the Pg-clamping check_primary_variables"""
    rng = np.random.default_rng(SEED + 16)
    prm = wo.make_params(eos=wo.EOS_WCE, thermo=thermo)
    eos = wo.lib().wo_eos_create(C.byref(prm))
    th = wo.lib().wo_thermo_create(thermo, 0)
    ntrans = nchanged = 0
    try:
        for trial in range(500):
            old_region = int(rng.choice([1, 2, 4]))
            oldp = wce_cell(wo, thermo, rng, old_region)
            dpg = rng.uniform(-1.2, 0.5) * oldp[2] + rng.choice([0.0, rng.uniform(-2e4, 2e4)])
            if old_region == 4:
                newp = oldp + np.array([rng.uniform(-2e5, 2e5), rng.uniform(-1.2, 1.2), dpg])
                if trial % 7 == 0:
                    newp[1] = oldp[1]
                    oldp[1] = newp[1] = rng.choice([-0.2, 1.3])
            else:
                ps = C.c_double()
                wo.lib().wo_saturation_pressure(th, oldp[1], C.byref(ps))
                pg = max(oldp[2] + dpg, -1e4)
                newp = np.array([ps.value * rng.uniform(0.7, 1.3) + max(pg, 0.0), oldp[1] + rng.uniform(-5, 5), pg])
            if trial % 11 == 0:
                newp[2] = newp[0] * 1.01  # partial pressure above the total pressure: clamped
            old_fluid = np.zeros(26)
            old_fluid[2] = old_region
            assert wo.lib().wo_eos_bulk_properties(eos, wo.dp(oldp), wo.dp(old_fluid)) == 0
            fluid = old_fluid.copy()
            p_ref = newp.copy()
            tr, ch = C.c_int(), C.c_int()
            e0 = wo.lib().wo_eos_transition(eos, wo.dp(oldp), wo.dp(p_ref), wo.dp(old_fluid), wo.dp(fluid), C.byref(tr))
            if e0 == 0:
                e0 = wo.lib().wo_eos_check_primary_variables(eos, wo.dp(fluid), wo.dp(p_ref), C.byref(ch))
            p_got = newp.copy()
            reg, tr1, ch1 = C.c_int(old_region), C.c_int(), C.c_int()
            e1 = hc.hc_wce_transition(C.byref(prm), wo.dp(oldp), wo.dp(p_got), old_region, old_fluid[1], C.byref(reg),
                                      C.byref(tr1), C.byref(ch1))
            assert (e0 != 0) == (e1 != 0), (trial, e0, e1)
            if e0 == 0:
                assert tr.value == tr1.value and ch.value == ch1.value
                assert int(round(fluid[2])) == reg.value
                assert np.array_equal(p_ref, p_got), (trial, p_ref, p_got)
                ntrans += tr.value
                nchanged += ch.value
        assert ntrans > 50 and nchanged > 10
    finally:
        wo.lib().wo_eos_destroy(eos)
        wo.lib().wo_thermo_destroy(th)


@pytest.mark.parametrize("thermo", [0, 1])
def test_source_controls_match_oracle(wo, hc, thermo):
    """wb_source_rate / wb_source_separated of wb_state.cuh (deliverability, recharge / injectivity, direction, total /
    water / steam limiters, one- and two-stage separators) against the oracle's source_network update, source by source
    on two-phase, liquid and vapour cells; the separator stage enthalpies bit for bit (the reference's known answers for
    them are in tests/test_separator.py)"""
    from waiwera_b200 import flow, mesh as wmesh
    L = wo.lib()
    for pr in (1.0e5, 5.5e5, 10.0e5, 14.5e5):
        a, b = C.c_double(), C.c_double()
        th = L.wo_thermo_create(thermo, 0)
        assert L.wo_separator_stage(th, pr, C.byref(a), C.byref(b)) == 0
        L.wo_thermo_destroy(th)
        ha, hb = C.c_double(), C.c_double()
        assert hc.hc_separator_stage(thermo, pr, C.byref(ha), C.byref(hb)) == 0
        assert (ha.value, hb.value) == (a.value, b.value)
    m = wmesh.structured(3, 1, 1, dx=10.0, heterogeneous=False)
    cells = [([30.0e5, 0.4], 4), ([30.0e5, 150.0], 1), ([1.0e5, 150.0], 2)]          # two-phase, liquid, vapour
    primary = np.array([c[0] for c in cells])
    region = np.array([c[1] for c in cells], np.int32)
    prm_o = wo.make_params(eos=wo.EOS_WE, thermo=thermo)
    prm_h = flow.make_params(eos=flow.EOS_WE, thermo=thermo)
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1)
    f = wo.Flow(prm_o, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    assert f.fluid_init(y, region) == 0
    rng = np.random.default_rng(SEED + 9)
    sep1, sep2 = [2.0e5], [8.0e5, 1.5e5]
    cases = []
    for cell in range(3):
        for kind in ("fixed", "deliv", "recharge"):
            for direction in (0, 1, 2):
                for seps in ([], sep1, sep2):
                    cases.append(dict(cell=cell, kind=kind, direction=direction, seps=seps,
                                      rate=float(rng.choice([-6.0, -0.5, 4.0])), pi=float(rng.uniform(1e-13, 1e-11)),
                                      coef=float(rng.uniform(1e-7, 1e-5)), pref=float(rng.choice([0.5e5, 20.0e5, 45.0e5])),
                                      limit=float(rng.choice([0.0, 0.3, 50.0])), lw=float(rng.choice([0.0, 0.2, 50.0])),
                                      ls=float(rng.choice([0.0, 0.05, 50.0]))))
    n = len(cases)
    f.set_sources([c["cell"] for c in cases], [0] * n, [c["rate"] for c in cases], [0.0] * n)
    f.set_source_controls(list(range(n)), [c["pi"] if c["kind"] == "deliv" else 0.0 for c in cases], [c["pref"] for c in cases],
                          [c["direction"] for c in cases], [c["limit"] for c in cases])
    rc = [k for k, c in enumerate(cases) if c["kind"] == "recharge"]
    f.set_source_recharge(rc, [cases[k]["coef"] for k in rc], [cases[k]["pref"] for k in rc])
    assert f.set_source_separators(list(range(n)), [c["seps"] for c in cases], [c["lw"] for c in cases], [c["ls"] for c in cases]) == 0
    e, L0 = f.lhs(y)
    assert e == 0 and f.residual(y, L0, 1.0e3)[0] == 0
    ref = f.source_rates(n)
    nonzero = limited = separated = 0
    for k, c in enumerate(cases):
        ctrl = (1 if c["kind"] == "deliv" else 0) | (c["direction"] << 1) | (8 if c["kind"] == "recharge" else 0)
        sh = np.zeros(4)
        for q, pr in enumerate(c["seps"]):
            a, b = C.c_double(), C.c_double()
            assert hc.hc_separator_stage(thermo, pr, C.byref(a), C.byref(b)) == 0
            sh[2 * q], sh[2 * q + 1] = a.value, b.value
        out = np.zeros(6)
        pi = c["coef"] if c["kind"] == "recharge" else (c["pi"] if c["kind"] == "deliv" else 0.0)
        r = hc.hc_source_rate(C.addressof(prm_h), wo.dp(np.ascontiguousarray(primary[c["cell"]])), int(region[c["cell"]]),
                              wo.dp(np.ascontiguousarray(m.rock[c["cell"]])), ctrl, pi, c["pref"], c["limit"], c["rate"],
                              len(c["seps"]), wo.dp(sh), c["lw"], c["ls"], wo.dp(out))
        assert r == 0
        assert close(out[0], ref[k], 1e-14) or (out[0] == 0.0 and ref[k] == 0.0), (c, out[0], ref[k])
        sep = f.source_separated(k, ref[k])
        assert np.allclose(out[1:], sep, rtol=1e-13, atol=0.0), (c, out[1:], sep)
        nonzero += ref[k] != 0.0
        separated += sep[2] != 0.0
        base = c["rate"] if c["kind"] == "fixed" else None
        limited += base is not None and ref[k] != 0.0 and abs(ref[k]) < abs(base) * (1 - 1e-12)
    assert nonzero > n // 3 and separated > 10 and limited > 5


@pytest.mark.parametrize("thermo", [0, 1])
def test_source_pressure_table_matches_oracle(wo, hc, thermo):
    """wb_source_rate with a reference-pressure table (wb_set_source_pressure_table: against the flowing enthalpy or the
    pressure, linear or step, inside and beyond the table) against the oracle, whose value for the reference's own case is
    pinned in tests/test_oracle_kat.py (source 11 of source_control_test.F90: -10.3366086953508 kg/s)"""
    from waiwera_b200 import flow, mesh as wmesh
    m = wmesh.structured(3, 1, 1, dx=10.0, heterogeneous=False)
    cells = [([30.0e5, 0.4], 4), ([30.0e5, 150.0], 1), ([1.0e5, 150.0], 2)]          # two-phase, liquid, vapour
    primary = np.array([c[0] for c in cells])
    region = np.array([c[1] for c in cells], np.int32)
    prm_o = wo.make_params(eos=wo.EOS_WE, thermo=thermo)
    prm_h = flow.make_params(eos=flow.EOS_WE, thermo=thermo)
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1)
    f = wo.Flow(prm_o, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    assert f.fluid_init(y, region) == 0
    tables = [[[0.0, 22.0e5], [11.0e5, 20.0e5], [28.0e5, 0.5e5]],                     # the reference's test table
              [[5.0e5, 1.0e5], [6.5e5, 2.0e5], [9.0e5, 4.0e5], [20.0e5, 8.0e5], [26.0e5, 12.0e5], [27.0e5, 13.0e5], [27.5e5, 14.0e5], [29.0e5, 25.0e5]],
              [[15.0e5, 3.0e5]]]
    cases = [dict(cell=cell, table=t, coord=coord, step=step, direction=direction)
             for cell in range(3) for t in tables for coord in (0, 1) for step in (0, 1) for direction in (0, 1)]
    n = len(cases)
    f.set_sources([c["cell"] for c in cases], [0] * n, [-1.0] * n, [0.0] * n)
    f.set_source_controls(list(range(n)), [2e-12] * n, [7.0e5] * n, [c["direction"] for c in cases], [0.0] * n)
    assert f.set_source_pressure_table(list(range(n)), [c["table"] for c in cases], [c["coord"] for c in cases],
                                       [c["step"] for c in cases]) == 0
    e, L0 = f.lhs(y)
    assert e == 0 and f.residual(y, L0, 1.0e3)[0] == 0
    ref = f.source_rates(n)
    plain = []
    for k, c in enumerate(cases):
        tab = np.zeros(16)
        tab[:2 * len(c["table"])] = np.array(c["table"]).reshape(-1)
        out = np.zeros(1)
        word = len(c["table"]) | (256 if c["coord"] else 0) | (512 if c["step"] else 0)
        args = (C.addressof(prm_h), wo.dp(np.ascontiguousarray(primary[c["cell"]])), int(region[c["cell"]]),
                wo.dp(np.ascontiguousarray(m.rock[c["cell"]])), 1 | (c["direction"] << 1), 2e-12, 7.0e5)
        assert hc.hc_source_rate_ptab(*args, word, wo.dp(tab), wo.dp(out)) == 0
        assert close(out[0], ref[k], 1e-14) or (out[0] == 0.0 and ref[k] == 0.0), (c, out[0], ref[k])
        base = np.zeros(1)
        assert hc.hc_source_rate_ptab(*args, 0, wo.dp(tab), wo.dp(base)) == 0          # no table: the fixed reference pressure
        plain.append(base[0])
    assert len(set(np.round(ref / 1e-3).tolist())) > 12 and (np.abs(ref - np.array(plain)) > 1e-6).sum() > n // 2
