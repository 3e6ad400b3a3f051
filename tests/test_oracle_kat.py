"""Pins the CPU oracle to the reference's own known-answer unit tests.

Every expected value below is transcribed from the reference test named in
the docstring (paths relative to the Waiwera tree, v1.5.1).  Tolerances are
the reference module tolerances (test%tolerance) or tighter.
"""
import ctypes as C

import numpy as np
import pytest

TC_K = 273.15


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def region_props(wo, th, region, p, t):
    param = np.array([p, t])
    props = np.zeros(2)
    err = wo.lib().wo_region_properties(th, region, wo.dp(param), wo.dp(props))
    return err, props


@pytest.fixture(scope="module")
def iapws(wo):
    th = wo.lib().wo_thermo_create(wo.THERMO_IAPWS, 0)
    yield th
    wo.lib().wo_thermo_destroy(th)


@pytest.fixture(scope="module")
def ifc67(wo):
    th = wo.lib().wo_thermo_create(wo.THERMO_IFC67, 0)
    yield th
    wo.lib().wo_thermo_destroy(th)


# ---- test/unit/src/IAPWS_test.F90:62-352 (tolerance 1e-7) ----

def test_iapws_region1(wo, iapws):
    """IAPWS_test.F90:62-100"""
    pts = [(3.e6, 300.), (80.e6, 300.), (3.e6, 500.)]
    nu = [0.100215168e-2, 0.971180894e-3, 0.120241800e-2]
    u = [0.112324818e6, 0.106448356e6, 0.971934985e6]
    for (p, tk), v, uu in zip(pts, nu, u):
        err, props = region_props(wo, iapws, 1, p, tk - TC_K)
        assert err == 0
        assert rel(props[0], 1.0 / v) < 1e-7
        assert rel(props[1], uu) < 1e-7
    for p, t in [(20.e6, 360.), (101.e6, 60.)]:
        assert region_props(wo, iapws, 1, p, t)[0] == 1


def test_iapws_region2(wo, iapws):
    """IAPWS_test.F90:104-142"""
    pts = [(0.0035e6, 300.), (0.0035e6, 700.), (30.e6, 700.)]
    nu = [0.394913866e2, 0.923015898e2, 0.542946619e-2]
    u = [0.241169160e7, 0.301262819e7, 0.246861076e7]
    for (p, tk), v, uu in zip(pts, nu, u):
        err, props = region_props(wo, iapws, 2, p, tk - TC_K)
        assert err == 0
        assert rel(props[0], 1.0 / v) < 1e-7
        assert rel(props[1], uu) < 1e-7
    for p, t in [(20.e6, 801.), (101.e6, 60.)]:
        assert region_props(wo, iapws, 2, p, t)[0] == 1


def test_iapws_region3(wo, iapws):
    """IAPWS_test.F90:146-182 (density, T) -> (pressure, energy)"""
    pts = [(500., 650.), (200., 650.), (500., 750.)]
    pr = [0.255837018e8, 0.222930643e8, 0.783095639e8]
    u = [0.181226279e7, 0.226365868e7, 0.210206932e7]
    for (d, tk), pp, uu in zip(pts, pr, u):
        err, props = region_props(wo, iapws, 3, d, tk - TC_K)
        assert err == 0
        assert rel(props[0], pp) < 1e-7
        assert rel(props[1], uu) < 1e-7
    assert region_props(wo, iapws, 3, 800., 400.)[0] == 1


def test_iapws_saturation(wo, iapws):
    """IAPWS_test.F90:186-218"""
    L = wo.lib()
    for tk, p in zip([300., 500., 600.], [0.353658941e4, 0.263889776e7, 0.123443146e8]):
        ps, ts = C.c_double(), C.c_double()
        assert L.wo_saturation_pressure(iapws, tk - TC_K, C.byref(ps)) == 0
        assert rel(ps.value, p) < 1e-7
        assert L.wo_saturation_temperature(iapws, ps.value, C.byref(ts)) == 0
        assert rel(ts.value, tk - TC_K) < 1e-7
    ps = C.c_double()
    assert L.wo_saturation_pressure(iapws, 380., C.byref(ps)) == 1
    assert L.wo_saturation_temperature(iapws, 30.e6, C.byref(ps)) == 1


def test_iapws_viscosity(wo, iapws):
    """IAPWS_test.F90:222-252"""
    t = np.array([298.15, 298.15, 373.15, 433.15, 433.15, 873.15, 873.15, 873.15, 1173.15, 1173.15, 1173.15]) - TC_K
    d = [998., 1200., 1000., 1., 1000., 1., 100., 600., 1., 100., 400.]
    visc = np.array([889.735100, 1437.649467, 307.883622, 14.538324, 217.685358, 32.619287, 35.802262,
                     77.430195, 44.217245, 47.640433, 64.154608]) * 1e-6
    reg = [1, 1, 1, 2, 1, 2, 2, 3, 2, 2, 2]
    for ti, di, vi, ri in zip(t, d, visc, reg):
        v = wo.lib().wo_region_viscosity(iapws, ri, float(ti), 1.e5, di)
        assert rel(v, vi) < 1e-7


def test_iapws_boundary23(wo):
    """IAPWS_test.F90:256-276"""
    t0, p0 = 0.62315e3 - TC_K, 0.165291643e8
    assert rel(wo.lib().wo_boundary23_pressure(t0), p0) < 1e-7
    assert rel(wo.lib().wo_boundary23_temperature(p0), t0) < 1e-7


def test_iapws_phase_composition(wo, iapws):
    """IAPWS_test.F90:280-348"""
    cases = [(1, 1.e5, 20., 0b001), (3, 200.e5, 360., 0b001), (2, 1.e5, 110., 0b010),
             (2, 175.e5, 360., 0b010), (2, 150.e5, 700., 0b010), (3, 180.e5, 360., 0b010),
             (3, 210.e5, 380., 0b010), (4, 33.466518715101621e5, 240., 0b011),
             (3, 500.e5, 390., 0b100), (3, 560.e5, 500., 0b100), (2, 300.e5, 700., 0b100)]
    for region, p, t, expected in cases:
        assert wo.lib().wo_phase_composition(iapws, region, p, t) == expected


# ---- test/unit/src/IFC67_test.F90:64-287 ----

def test_ifc67_region1(wo, ifc67):
    """IFC67_test.F90:64-103"""
    pts = [(3.e6, 300.), (80.e6, 300.), (3.e6, 500.)]
    rho = [997.95721560998174, 1029.7256888266911, 831.84196191567298]
    u = [112247.43313085975, 106310.47344628950, 971985.91117384087]
    for (p, tk), r, uu in zip(pts, rho, u):
        err, props = region_props(wo, ifc67, 1, p, tk - TC_K)
        assert err == 0
        assert rel(props[0], r) < 1e-12
        assert rel(props[1], uu) < 1e-11
    for p, t in [(20.e6, 360.), (101.e6, 60.)]:
        assert region_props(wo, ifc67, 1, p, t)[0] == 1


def test_ifc67_region2(wo, ifc67):
    """IFC67_test.F90:107-146"""
    pts = [(0.0035e6, 300.), (0.0035e6, 700.), (30.e6, 700.)]
    rho = [2.5316826343790743e-2, 1.0834441421293962e-2, 183.90041953968711]
    u = [2412405.0932077002, 3012229.4965919587, 2474981.3799304822]
    for (p, tk), r, uu in zip(pts, rho, u):
        err, props = region_props(wo, ifc67, 2, p, tk - TC_K)
        assert err == 0
        assert rel(props[0], r) < 1e-12
        assert rel(props[1], uu) < 1e-11
    for p, t in [(20.e6, 801.), (101.e6, 60.)]:
        assert region_props(wo, ifc67, 2, p, t)[0] == 1


def test_ifc67_saturation(wo, ifc67):
    """IFC67_test.F90:150-182"""
    L = wo.lib()
    for tk, p in zip([300., 500., 600.], [0.35323426e4, 0.263961572e7, 0.123493902e8]):
        ps, ts = C.c_double(), C.c_double()
        assert L.wo_saturation_pressure(ifc67, tk - TC_K, C.byref(ps)) == 0
        assert rel(ps.value, p) < 1e-7
        assert L.wo_saturation_temperature(ifc67, ps.value, C.byref(ts)) == 0
        assert rel(ts.value, tk - TC_K) < 1e-7
    ps = C.c_double()
    assert L.wo_saturation_pressure(ifc67, 380., C.byref(ps)) == 1
    assert L.wo_saturation_temperature(ifc67, 30.e6, C.byref(ps)) == 1


def test_ifc67_viscosity(wo, ifc67):
    """IFC67_test.F90:186-220 (7-digit literals)"""
    L = wo.lib()
    for tk, p, v in zip([298.15, 373.15], [1977563.58349, 99834578.2816], [8.903129e-04, 2.988268e-04]):
        assert rel(L.wo_region_viscosity(ifc67, 1, tk - TC_K, p, 0.0), v) < 1e-6
    for d, v in zip([1., 100.], [3.249537e-05, 3.667671e-05]):
        assert rel(L.wo_region_viscosity(ifc67, 2, 873.15 - TC_K, 0.0, d), v) < 1e-6


def test_ifc67_phase_composition_and_extrapolate(wo, ifc67):
    """IFC67_test.F90:224-283"""
    L = wo.lib()
    assert L.wo_phase_composition(ifc67, 1, 1.e5, 20.) == 0b01
    assert L.wo_phase_composition(ifc67, 2, 1.e5, 110.) == 0b10
    assert L.wo_phase_composition(ifc67, 4, 33.466518715101621e5, 240.) == 0b11
    assert region_props(wo, ifc67, 1, 30.e6, 355.)[0] == 1
    ex = L.wo_thermo_create(wo.THERMO_IFC67, 1)
    assert region_props(wo, ex, 1, 30.e6, 355.)[0] == 0
    L.wo_thermo_destroy(ex)


# ---- test/unit/src/powertable_test.F90 (spot check: multiplication chains give exact small powers) ----

def test_powertable(wo):
    powers = np.array([-3, -1, 2, 5, 7, 17], np.int32)
    q = np.array([-3, -1, 0, 1, 2, 5, 7, 17], np.int32)
    out = np.zeros(len(q))
    wo.lib().wo_powertable_eval(wo.ip(powers), len(powers), 2.0, wo.ip(q), len(q), wo.dp(out))
    assert np.array_equal(out, 2.0 ** q.astype(float))
    wo.lib().wo_powertable_eval(wo.ip(powers), len(powers), 1.3, wo.ip(q), len(q), wo.dp(out))
    assert np.allclose(out, 1.3 ** q.astype(float), rtol=1e-14, atol=0)


# ---- test/unit/src/eos_we_test.F90:65-474 (tolerance 1e-9 / 1e-6 for transitions) ----

def _eos_we(wo, **kw):
    prm = wo.make_params(eos=wo.EOS_WE, thermo=wo.THERMO_IAPWS, **kw)
    e = wo.lib().wo_eos_create(C.byref(prm))
    return prm, e


def test_eos_we_fluid_properties(wo):
    """eos_we_test.F90:65-187"""
    prm, e = _eos_we(wo, relperm=wo.make_relperm("linear", liquid=(0.2, 0.8), vapour=(0.2, 0.8)))
    L = wo.lib()
    fluid = np.zeros(23)
    rock = np.zeros(8)
    pressure, sv = 27.967924557686445e5, 0.25
    primary = np.array([pressure, sv])
    fluid[2] = fluid[3] = 4.0
    fluid[5] = 1.0
    assert L.wo_eos_bulk_properties(e, wo.dp(primary), wo.dp(fluid)) == 0
    assert L.wo_eos_phase_properties(e, wo.dp(primary), wo.dp(rock), wo.dp(fluid)) == 0
    tol = 1e-9
    assert fluid[0] == pressure
    assert rel(fluid[1], 230.0) < tol
    assert int(round(fluid[4])) == 0b011
    liq, vap = fluid[7:15], fluid[15:23]
    exp_l = dict(rho=827.12247049977032, u=986828.18916209263, h=990209.54144729744, mu=1.1619412513757267e-4,
                 kr=11. / 12., pc=0.0, s=0.75)
    exp_v = dict(rho=13.984012253728331, u=2603010.010356456, h=2803009.2956133024, mu=1.6704837258831552e-5,
                 kr=1. / 12., pc=0.0, s=0.25)
    for ph, ex in ((liq, exp_l), (vap, exp_v)):
        assert rel(ph[0], ex["rho"]) < tol
        assert rel(ph[1], ex["mu"]) < tol
        assert rel(ph[2], ex["s"]) < tol
        assert rel(ph[3], ex["kr"]) < tol
        assert ph[4] == ex["pc"]
        assert rel(ph[5], ex["h"]) < tol
        assert rel(ph[6], ex["u"]) < tol
        assert ph[7] == 1.0
    L.wo_eos_destroy(e)


def test_eos_we_transitions(wo):
    """eos_we_test.F90:191-326 (transition_compare tolerance 1e-6, unit_test_utils.F90:35)"""
    prm, e = _eos_we(wo)
    L = wo.lib()
    small = 1e-6
    cases = [
        # old_region, old_T, old_primary, primary, expected_primary, expected_region, transition
        (1, 0.0, [1.e5, 20.], [1.e5, 20.], [1.e5, 20.], 1, False),
        (1, 0.0, [20.e5, 210.], [15.e5, 200.], [16.647121334271149e5, small], 4, True),
        (2, 0.0, [1.e5, 120.], [1.e5, 120.], [1.e5, 120.], 2, False),
        (2, 0.0, [84.0e5, 302.], [86.e5, 299.27215502281706], [85.621455812056474e5, 1. - small], 4, True),
        (4, 0.0, [1.e5, 0.5], [1.e5, 0.5], [1.e5, 0.5], 4, False),
        (4, 299.27215502281706, [85.e5, 0.1], [86.e5, -0.01], [85.90917681818182e5, 300.02645326107097], 1, True),
        (4, 212.38453531849041, [20.e5, 0.9], [20.1e5, 1.02], [20.08331325e5, 212.59487472987195], 2, True),
    ]
    for old_region, old_t, old_p, prim, exp_p, exp_region, exp_tr in cases:
        old_fluid, fluid = np.zeros(23), np.zeros(23)
        old_fluid[2] = fluid[2] = float(old_region)
        old_fluid[1] = old_t
        op, p = np.array(old_p), np.array(prim)
        tr = C.c_int()
        err = L.wo_eos_transition(e, wo.dp(op), wo.dp(p), wo.dp(old_fluid), wo.dp(fluid), C.byref(tr))
        assert err == 0
        assert bool(tr.value) == exp_tr
        assert int(round(fluid[2])) == exp_region
        for a, b in zip(p, exp_p):
            assert rel(a, b) < 1e-6
    L.wo_eos_destroy(e)


def test_eos_we_errors(wo):
    """eos_we_test.F90:330-396"""
    prm, e = _eos_we(wo, relperm=wo.make_relperm("linear", liquid=(0.2, 0.8), vapour=(0.2, 0.8)))
    L = wo.lib()
    for (p, t), region in zip([(20.e6, 360.), (101.e6, 20.)], [1, 2]):
        fluid, rock = np.zeros(23), np.zeros(8)
        fluid[2] = float(region)
        primary = np.array([p, t])
        err = L.wo_eos_bulk_properties(e, wo.dp(primary), wo.dp(fluid))
        if err == 0:
            err = L.wo_eos_phase_properties(e, wo.dp(primary), wo.dp(rock), wo.dp(fluid))
        assert err == 1
    L.wo_eos_destroy(e)


def test_eos_we_conductivity(wo):
    """eos_we_test.F90:400-474 (tol 1e-7)"""
    rock = np.zeros(8)
    rock[3:5] = [1.5, 1.0]
    fluid = np.zeros(23)
    for sl, ex in [(0.0, 1.0), (0.25, 1.25), (0.5, 1.3535534), (0.75, 1.4330127), (1.0, 1.5)]:
        fluid[9] = sl
        assert abs(wo.lib().wo_eos_conductivity(wo.dp(rock), wo.dp(fluid), 1) - ex) < 1e-7


def test_eos_w_properties(wo):
    """eos_w_test.F90:70-82: rho, u, h, mu at 1 bar / 20 degC (IAPWS)"""
    prm = wo.make_params(eos=wo.EOS_W, thermo=wo.THERMO_IAPWS, eos_w_temperature=20.0)
    L = wo.lib()
    e = L.wo_eos_create(C.byref(prm))
    fluid, rock = np.zeros(15), np.zeros(8)
    fluid[2] = 1.0
    primary = np.array([1.e5])
    assert L.wo_eos_bulk_properties(e, wo.dp(primary), wo.dp(fluid)) == 0
    assert L.wo_eos_phase_properties(e, wo.dp(primary), wo.dp(rock), wo.dp(fluid)) == 0
    assert rel(fluid[7 + 0], 998.20548637769673) < 1e-9
    assert rel(fluid[7 + 6], 83911.631393167205) < 1e-9
    assert rel(fluid[7 + 5], 84011.811167136271) < 1e-9
    assert rel(fluid[7 + 1], 1.0015972622270245e-3) < 1e-9
    L.wo_eos_destroy(e)


# ---- test/unit/src/cell_test.F90:98-141 ----

def test_cell_balance(wo):
    rock = np.array([0., 0., 0., 0., 0., 0.1, 2200., 950.])
    fluid = np.array([2.7e5, 130., 4., 4., 3., 1., 0., 0.,
                      935., 0., 0.8, 0., 0., 0., 5.461e5, 0.7, 0.3,
                      1.5, 0., 0.2, 0., 0., 0., 2.540e6, 0.4, 0.6])
    bal = np.zeros(3)
    wo.lib().wo_cell_balance(wo.dp(rock), wo.dp(fluid), 2, 2, 3, wo.dp(bal))
    for a, b in zip(bal, [52.372, 22.458, 2.8545448e8]):
        assert rel(a, b) < 1e-8


# ---- test/unit/src/rock_test.F90:55-135 ----

def test_rock_record_layout_and_energy(wo):
    """rock%assign (rock_test.F90:55-98): permeability(3), wet / dry conductivity, porosity, density, specific heat in
    that order -- the 8-double record of the C ABI; rock%energy (:102-135, src/rock.F90:142): density * specific heat * T
    = 2.717e8 J/m3 at 130 degC, seen through the energy balance of a cell with no fluid in it"""
    from waiwera_b200 import mesh as wmesh
    rock = wmesh.default_rock(1, None, heterogeneous=False)[0]
    assert rock.shape == (8,) and list(rock[3:]) == [2.5, 2.5, 0.1, 2200.0, 1000.0]      # defaults src/rock.F90:69-76
    rock = np.array([1.e-12, 1.e-13, 1.e-14, 2.5, 1.5, 0.1, 2200., 950.])
    fluid = np.zeros(26)
    fluid[1] = 130.0
    bal = np.zeros(3)
    wo.lib().wo_cell_balance(wo.dp(rock), wo.dp(fluid), 2, 2, 3, wo.dp(bal))
    assert bal[0] == 0.0 and bal[1] == 0.0 and rel(bal[2] / (1.0 - rock[5]), 2.717e8) < 1e-15


# ---- test/unit/src/face_test.F90:102-728 ----

def test_face_distances(wo):
    """face_test.F90:102-150"""
    c1, c2 = np.array([-80., 200., 50.]), np.array([100., 200., 50.])
    fc = np.array([0., 200., 50.])
    for normal, ex in [([1., 0., 0.], (80., 100.)), ([-1., 0., 0.], (-80., -100.))]:
        n = np.array(normal)
        dist, d12 = np.zeros(2), C.c_double()
        wo.lib().wo_face_calculate_distances(wo.dp(c1), wo.dp(c2), wo.dp(fc), wo.dp(n), wo.dp(dist),
                                             C.cast(C.byref(d12), wo.c_dp))
        assert np.allclose(dist, ex, rtol=1e-14)
        assert abs(d12.value - sum(ex)) < 1e-12


def test_face_harmonic_average(wo):
    """face_test.F90:273-343"""
    xs = [(240., 170.), (240., 170.), (240., 170.), (0., 170.), (240., 0.), (0., 0.)]
    ds = [(25., 32.), (0., 10.), (22., 0.), (25., 32.), (25., 32.), (25., 32.)]
    ex = [194.937133277, 170., 240., 0., 0., 0.]
    for x, dd, e in zip(xs, ds, ex):
        g = np.zeros(12)
        g[1:3] = dd
        g[3] = sum(dd)
        xv = np.array(x)
        got = wo.lib().wo_face_harmonic_average(wo.dp(g), wo.dp(xv))
        assert abs(got - e) <= 1e-9 * max(1.0, abs(e))


ROCK1 = [1.e-14, 2.e-14, 3.e-15, 2.5, 2.5, 0.1, 2200., 1000.]
FLUID_1PH = [1.e5, 20., 1., 1., 1., 1., 0., 998.2, 1.e-3, 1., 1., 0., 84011.8, 83911.6, 1.,
             0., 0., 0., 0., 0., 0., 0., 0.]


def _flux(wo, g, r1, r2, f1, f2):
    flux = np.zeros(4)
    a = [np.array(v, dtype=np.float64) for v in (g, r1, r2, f1, f2)]
    wo.lib().wo_face_flux(*[wo.dp(v) for v in a], 1, 2, 2, 2, 0, wo.dp(flux))
    return flux


def test_face_flux_zero_horizontal(wo):
    """face_test.F90:347-418"""
    g = [0., 25., 35., 60., 1., 0., 0., 0., 0., 0., 0., 1.]
    assert np.array_equal(_flux(wo, g, ROCK1, ROCK1, FLUID_1PH, FLUID_1PH), np.zeros(4))


def test_face_flux_vertical_gravity(wo):
    """face_test.F90:422-495"""
    g = [0., 25., 35., 60., 0., 0., -1., 9.8, 0., 0., 0., 3.]
    flux = _flux(wo, g, ROCK1, ROCK1, FLUID_1PH, FLUID_1PH)
    assert rel(flux[0], 2.9294255256e-5) < 1e-9
    assert rel(flux[1], 2.4610631137) < 1e-9
    assert flux[2] == flux[0]
    assert flux[3] == 0.0


def test_face_flux_hydrostatic(wo):
    """face_test.F90:499-580: gravity balances the pressure gradient"""
    g = [0., 25., 35., 60., 0., 0., -1., 9.8, 0., 0., 0., 3.]
    f1 = [2.e5, 20., 1., 1., 1., 1., 0., 998.2512244888, 0.00100156652270771, 1., 1., 0.,
          84105.9189422008, 83905.5685743839, 1., 0., 0., 0., 0., 0., 0., 0., 0.]
    f2 = [7.87050606076185e5, 20., 1., 1., 1., 1., 0., 998.5195444779, 0.00100138700807062, 1., 1., 0.,
          84658.2021844106, 83869.9846573438, 1., 0., 0., 0., 0., 0., 0., 0., 0.]
    flux = _flux(wo, g, ROCK1, ROCK1, f1, f2)
    # mass flux scale for this face is ~3e-5 (previous test); "zero" to the reference's tolerance
    assert abs(flux[0]) < 1e-6 * 2.9294255256e-5 * 1e3
    assert abs(flux[1]) < 1e-6 * 2.4610631137 * 1e3
    assert abs(flux[2]) < 1e-6 * 2.9294255256e-5 * 1e3
    assert flux[3] == 0.0


def test_face_flux_two_phase_vertical(wo):
    """face_test.F90:584-675: counter-flow"""
    g = [0., 25., 35., 60., 0., 0., -1., 9.8, 0., 0., 0., 3.]
    r2 = [2.e-14, 3.e-14, 6.e-15, 2.7, 2.7, 0.05, 2300., 995.]
    f1 = [6.2e5, 160., 4., 4., 3., 1., 0., 907.45, 1.7e-4, 0.25, 0.75, 0., 675574.7, 674893.5, 1.,
          3.26, 1.43e-5, 0.75, 0.25, 0., 2757430.53, 2567774.0, 1.]
    f2 = [8.2e5, 171.44, 4., 4., 3., 1., 0., 895.98, 1.58e-4, 0.4, 0.6, 0., 725517.1, 724601.9, 1.,
          4.26, 1.47e-5, 0.6, 0.4, 0., 2769308.8, 2576807.25, 1.]
    flux = _flux(wo, g, ROCK1, r2, f1, f2)
    assert rel(flux[0], 9.14772841429594e-5) < 1e-10
    assert rel(flux[1], 57.9124776818) < 1e-10
    assert rel(flux[2], 9.30959555690338e-5) < 1e-10
    assert rel(flux[3], -1.61867142607443e-6) < 1e-10


# ---- relative_permeability_test.F90:74-344, capillary_pressure_test.F90:50-201 ----

def _rp(wo, r, sl):
    out = np.zeros(2)
    wo.lib().wo_relperm_values(C.byref(r), sl, wo.dp(out))
    return out


def test_relperm_curves(wo):
    lin = wo.make_relperm("linear", liquid=(0.1, 0.8), vapour=(0.3, 0.75))
    for sl, ex in [(0.01, (0., 1.)), (0.2, (1. / 7., 1.)), (0.5, (4. / 7., 4. / 9.)), (0.9, (1., 0.))]:
        assert np.allclose(_rp(wo, lin, sl), ex, rtol=1e-13, atol=1e-15)
    pick = wo.make_relperm("pickens", power=2.0)
    for sl, ex in [(0.01, (1.e-4, 1.)), (0.5, (0.25, 1.)), (0.9, (0.81, 1.))]:
        assert np.allclose(_rp(wo, pick, sl), ex, rtol=1e-13)
    corey = wo.make_relperm("corey", slr=0.3, ssr=0.1)
    for sl, ex in [(0.01, (0., 1.)), (0.5, (1. / 81., 32. / 81.)), (0.95, (1., 0.))]:
        assert np.allclose(_rp(wo, corey, sl), ex, rtol=1e-12, atol=1e-15)
    grant = wo.make_relperm("grant", slr=0.3, ssr=0.1)
    for sl, ex in [(0.01, (0., 1.)), (0.5, (1. / 81., 80. / 81.)), (0.95, (1., 0.))]:
        assert np.allclose(_rp(wo, grant, sl), ex, rtol=1e-12, atol=1e-15)
    vg = wo.make_relperm("van_genuchten", slr=0.1, sls=0.8, lambda_=0.5)
    for sl, ex in [(0.01, (0., 1.)), (0.25, (0.00024977947758877213, 0.9997502205224112)),
                   (0.5, (0.024315039984298164, 0.9756849600157018)),
                   (0.75, (0.38106285486468433, 0.6189371451353156)), (0.95, (1., 0.))]:
        assert np.allclose(_rp(wo, vg, sl), ex, rtol=1e-10, atol=1e-15)
    tab = wo.make_relperm("table", liquid=[(0, 0), (0.7, 0.01), (0.95, 0.99), (1, 1)],
                          vapour=[(0, 0), (0.05, 0.01), (0.3, 0.99), (1, 1)])
    for sl, ex in [(0., (0., 1.)), (0.3, (0.01 * 3. / 7., (4. + 3. * 0.99) / 7.)), (0.7, (0.01, 0.99)),
                   (0.9, (0.2 * 0.01 + 0.8 * 0.99, 0.8 * 0.01 + 0.2 * 0.99)), (1., (1., 0.))]:
        assert np.allclose(_rp(wo, tab, sl), ex, rtol=1e-12, atol=1e-15)
    mob = wo.make_relperm("fully_mobile")
    assert np.array_equal(_rp(wo, mob, 0.2), [1., 1.])


def test_cappress_curves(wo):
    L = wo.lib()
    t = 20.0
    z = wo.make_cappress("zero")
    assert L.wo_cappress_value(C.byref(z), 0.6, t) == 0.0
    lin = wo.make_cappress("linear", saturation_limits=(0.1, 0.8), pressure=0.2e5)
    for sl, ex in [(0., -0.2e5), (0.6, -0.0571428571428e5), (0.9, 0.)]:
        assert abs(L.wo_cappress_value(C.byref(lin), sl, t) - ex) <= 1e-9 * max(abs(ex), 1)
    vg = wo.make_cappress("van_genuchten", P0=0.2e5, lambda_=0.5, slr=0.1, sls=0.8)
    for sl, ex in [(0., 0.), (0.12, -6.99714227381e5), (0.15, -2.79284800875e5), (0.25, -0.911652955412e5),
                   (0.6, -0.195959179423e5), (0.9, 0.)]:
        assert abs(L.wo_cappress_value(C.byref(vg), sl, t) - ex) <= 1e-9 * max(abs(ex), 1)
    vgm = wo.make_cappress("van_genuchten", P0=0.2e5, lambda_=0.5, slr=0.1, sls=0.8, Pmax=6.e5)
    for sl, ex in [(0., -6.e5), (0.12, -6.e5), (0.6, -0.195959179423e5)]:
        assert abs(L.wo_cappress_value(C.byref(vgm), sl, t) - ex) <= 1e-9 * max(abs(ex), 1)
    tab = wo.make_cappress("table", pressure=[(0, -5.e5), (0.4, -1.e5), (0.7, 0)])
    for sl, ex in [(0., -5.e5), (0.3, -2.e5), (0.7, 0.), (0.9, 0.)]:
        assert abs(L.wo_cappress_value(C.byref(tab), sl, t) - ex) <= 1e-9 * max(abs(ex), 1)


# ---- test/unit/src/ncg_co2_thermodynamics_test.F90 (AUTOUGH2 values; tolerances 1e-9 / 1e-8) ----

def test_co2_henrys_constant(wo):
    """ncg_co2_thermodynamics_test.F90:47-135 (zero-salt cases)"""
    for t, ex in [(20., 1.44811504032e+08), (100., 5.50571700000e+08), (240., 5.21847810624e+08),
                  (300., 3.71913900000e+08), (350., 2.23454746875e+08)]:
        assert rel(wo.lib().wo_co2_henrys_constant(t), ex) < 1e-12


def test_co2_energy_solution(wo):
    """ncg_co2_thermodynamics_test.F90:139-257 (zero-salt cases)"""
    L = wo.lib()
    for t, ex in [(20., -495750.87299689), (100., -180685.98723494), (240., 242741.64505202),
                  (300., 407409.27618764)]:
        assert rel(L.wo_co2_energy_solution(t, L.wo_co2_henrys_constant(t)), ex) < 1e-12


def test_co2_viscosity(wo):
    """ncg_co2_thermodynamics_test.F90:261-309 (tol 1e-8)"""
    pc = [0.1e6, 1.e6, 5.e6, 10.e6, 20.e6, 30.e6]
    ts = [20., 100., 200., 300., 350.]
    expected = np.array([
        1.47350850e-5, 1.63927474e-5, 2.37601356e-5, 3.29693708e-5, 9.99600434e-5, 1.19066342e-4,
        1.82742115e-5, 1.86681905e-5, 2.04192081e-5, 2.26079800e-5, 3.76607100e-5, 5.40893300e-5,
        2.24530737e-5, 2.26583470e-5, 2.35706728e-5, 2.47110800e-5, 2.94125600e-5, 3.50956800e-5,
        2.62857731e-5, 2.64270283e-5, 2.70548291e-5, 2.78395800e-5, 3.04311100e-5, 3.39737300e-5,
        2.80772517e-5, 2.81462344e-5, 2.84528241e-5, 2.88360612e-5, 2.93062944e-5, 3.36788644e-5]).reshape(5, 6)
    v = C.c_double()
    for ip_, p in enumerate(pc):
        for it, t in enumerate(ts):
            assert wo.lib().wo_co2_viscosity(p, t, C.byref(v)) == 0
            assert abs(v.value - expected[it, ip_]) < 1e-8 * max(1.0, abs(expected[it, ip_]))
            assert rel(v.value, expected[it, ip_]) < 1e-7
    assert wo.lib().wo_co2_viscosity(301.e5, 100., C.byref(v)) == 1


def test_co2_properties(wo):
    """ncg_co2_thermodynamics_test.F90:313-362 (tol 1e-9): (Pc, T) -> enthalpy, density"""
    data = np.array([
        0.0, 20.0, 17140.18077231938, 0.0,
        100000.0, 20.0, 16142.247883091828, 1.8142044368713437,
        0.0, 100.0, 87450.99131436742, 0.0,
        100000.0, 100.0, 87004.524163092, 1.4213754811567743,
        4000000.0, 100.0, 64355.3813832885, 62.608990505735434,
        9000000.0, 100.0, 20379.357776952613, 184.7959892299282,
        0.0, 240.0, 223594.37705727902, 0.0,
        100000.0, 240.0, 223439.99865083068, 1.0324489144812645,
        4000000.0, 240.0, 215608.4290498441, 42.27375154306431,
        9000000.0, 240.0, 200402.49860929986, 100.70459422220841,
        0.0, 300.0, 286380.4950504236, 0.0,
        100000.0, 300.0, 286273.71092985675, 0.9242369906584087,
        4000000.0, 300.0, 280856.58497462136, 37.5055455044134,
        9000000.0, 300.0, 270338.58607276645, 87.3627658128452]).reshape(14, 4)
    props = np.zeros(2)
    for pc, t, h, d in data:
        wo.lib().wo_co2_properties(pc, t, wo.dp(props))
        assert rel(props[1], h) < 1e-12
        assert abs(props[0] - d) <= 1e-12 * max(d, 1.0)


# ---- test/unit/src/eos_wge_test.F90:53-279 (eos wce is the concrete CO2 child, src/eos_wce.F90) ----

def test_eos_wge_transitions(wo):
    """all 14 cases of test_eos_wge_transition (transition_compare tolerance 1e-6)"""
    prm = wo.make_params(eos=wo.EOS_WCE, thermo=wo.THERMO_IAPWS)
    L = wo.lib()
    e = L.wo_eos_create(C.byref(prm))
    assert L.wo_eos_num_primary(e) == 3 and L.wo_eos_fluid_dof(e) == 26
    small = 1e-6
    t41, t42 = 299.27215502281706, 212.38453531849041
    cases = [
        (1, 0., [1.e5, 20., 0.], [1.e5, 20., 0.], [1.e5, 20., 0.], 1, False),
        (1, 0., [1.e5, 20., 0.2e5], [1.e5, 20., 0.2e5], [1.e5, 20., 0.2e5], 1, False),
        (1, 0., [20.e5, 210., 0.], [15.e5, 200., 0.], [16.647121334271149e5, small, 0.], 4, True),
        (1, 0., [21.e5, 210., 1.e5], [17.e5, 200., 2.e5], [18.31769706741692e5, small, 1.6705757331457702e5], 4, True),
        (2, 0., [1.e5, 120., 0.], [1.e5, 120., 0.], [1.e5, 120., 0.], 2, False),
        (2, 0., [1.e5, 120., 0.2e5], [1.e5, 120., 0.2e5], [1.e5, 120., 0.2e5], 2, False),
        (2, 0., [84.0e5, 302., 0.], [86.e5, t41, 0.], [85.621455812056474e5, 1. - small, 0.], 4, True),
        (2, 0., [86.0e5, 302., 2.e5], [87.e5, t41, 1.e5], [86.810727906028237e5, 1. - small, 1.1892720939717567e5], 4, True),
        (4, 0., [1.e5, 0.5, 0.], [1.e5, 0.5, 0.], [1.e5, 0.5, 0.], 4, False),
        (4, 0., [1.e5, 0.5, 0.2e5], [1.e5, 0.5, 0.2e5], [1.e5, 0.5, 0.2e5], 4, False),
        (4, t41, [85.e5, 0.1, 0.], [86.e5, -0.01, 0.], [85.909176818181816e5, 300.02645326107097, 0.], 1, True),
        (4, t41, [88.e5, 0.1, 3.e5], [87.5e5, -0.01, 1.5e5], [87.545540454545449e5, 300.02645326107097, 1.6363636363636365e5], 1, True),
        (4, t42, [20.e5, 0.9, 0.], [20.1e5, 1.02, 0.], [20.08331325e5, 212.59487472987195, 0.], 2, True),
        (4, t42, [22.e5, 0.9, 2.e5], [24.1e5, 1.02, 4.e5], [23.749979916666667e5, 212.59487472987195, 3.6666666666666663e5], 2, True),
    ]
    for old_region, old_t, old_p, prim, exp_p, exp_region, exp_tr in cases:
        old_fluid, fluid = np.zeros(26), np.zeros(26)
        old_fluid[2] = fluid[2] = float(old_region)
        old_fluid[1] = old_t
        op, p = np.array(old_p), np.array(prim)
        tr = C.c_int()
        err = L.wo_eos_transition(e, wo.dp(op), wo.dp(p), wo.dp(old_fluid), wo.dp(fluid), C.byref(tr))
        assert err == 0
        assert bool(tr.value) == exp_tr, (old_p, prim)
        assert int(round(fluid[2])) == exp_region
        for a, b in zip(p, exp_p):
            assert abs(a - b) <= 1e-6 * max(abs(b), 1.0), (old_p, prim, p, exp_p)
    L.wo_eos_destroy(e)


def test_eos_wce_scaling_and_consistency(wo):
    """adaptive partial-pressure scaling round trip (eos_wge.F90:639-674), check_primary_variables clamping
    (:573-635) and internal consistency of a two-phase wce record: pure-water limit equals eos_we, mass
    fractions sum to 1, u = h - P/rho, liquid density carries no free gas (effective_properties)"""
    L = wo.lib()
    prm = wo.make_params(eos=wo.EOS_WCE, thermo=wo.THERMO_IAPWS)
    e = L.wo_eos_create(C.byref(prm))
    prim = np.array([30.e5, 0.3, 4.e5])
    y, back = np.zeros(3), np.zeros(3)
    L.wo_eos_scale(e, wo.dp(prim), 4, wo.dp(y))
    assert np.allclose(y, [3.0, 0.3, 4.e5 / 30.e5], rtol=1e-15)
    L.wo_eos_unscale(e, wo.dp(y), 4, wo.dp(back))
    assert np.allclose(back, prim, rtol=1e-15)
    fluid = np.zeros(26)
    fluid[2] = 4.0
    ch = C.c_int()
    p2 = np.array([30.e5, 0.3, 31.e5])
    assert L.wo_eos_check_primary_variables(e, wo.dp(fluid), wo.dp(p2), C.byref(ch)) == 0
    assert ch.value == 1 and p2[2] == (1. - 1e-6) * 30.e5
    p3 = np.array([30.e5, 0.3, -5.])
    assert L.wo_eos_check_primary_variables(e, wo.dp(fluid), wo.dp(p3), C.byref(ch)) == 0 and p3[2] == 0.0 and ch.value == 1
    assert L.wo_eos_check_primary_variables(e, wo.dp(fluid), wo.dp(np.array([30.e5, 2.5, 1.e5])), C.byref(ch)) == 1
    rock = np.zeros(8)
    assert L.wo_eos_bulk_properties(e, wo.dp(prim), wo.dp(fluid)) == 0
    assert L.wo_eos_phase_properties(e, wo.dp(prim), wo.dp(rock), wo.dp(fluid)) == 0
    assert fluid[6] == 26.e5 and fluid[7] == 4.e5
    liq, vap = fluid[8:17], fluid[17:26]
    for ph in (liq, vap):
        assert abs(ph[7] + ph[8] - 1.0) < 1e-15
        assert rel(ph[6], ph[5] - fluid[0] / ph[0]) < 1e-14
    props = np.zeros(2)
    L.wo_co2_properties(4.e5, fluid[1], wo.dp(props))
    th = L.wo_thermo_create(0, 0)
    wp = np.zeros(2)
    L.wo_region_properties(th, 1, wo.dp(np.array([30.e5, fluid[1]])), wo.dp(wp))
    assert liq[0] == wp[0]                      # liquid: water density only
    L.wo_region_properties(th, 2, wo.dp(np.array([26.e5, fluid[1]])), wo.dp(wp))
    assert rel(vap[0], wp[0] + props[0]) < 1e-15 and rel(vap[8], props[0] / (wp[0] + props[0])) < 1e-14
    # pure-water limit: Pg = 0 reproduces eos_we
    prm_we = wo.make_params(eos=wo.EOS_WE, thermo=wo.THERMO_IAPWS)
    ewe = L.wo_eos_create(C.byref(prm_we))
    f3, f2 = np.zeros(26), np.zeros(23)
    f3[2] = f2[2] = 4.0
    assert L.wo_eos_bulk_properties(e, wo.dp(np.array([30.e5, 0.3, 0.0])), wo.dp(f3)) == 0
    assert L.wo_eos_phase_properties(e, wo.dp(np.array([30.e5, 0.3, 0.0])), wo.dp(rock), wo.dp(f3)) == 0
    assert L.wo_eos_bulk_properties(ewe, wo.dp(np.array([30.e5, 0.3])), wo.dp(f2)) == 0
    assert L.wo_eos_phase_properties(ewe, wo.dp(np.array([30.e5, 0.3])), wo.dp(rock), wo.dp(f2)) == 0
    assert f3[1] == f2[1]
    for k in range(7):  # density .. internal energy of both phases
        assert rel(f3[8 + k], f2[7 + k]) < 1e-13 and rel(f3[17 + k], f2[15 + k]) < 1e-13
    L.wo_eos_destroy(e)
    L.wo_eos_destroy(ewe)
    L.wo_thermo_destroy(th)


def test_fluid_sums(wo):
    """test/unit/src/fluid_test.F90:115-243: fluid%component_density / energy (through cell%balance with porosity 1),
    phase_mobilities, phase_flow_fractions, component_flow_fractions and specific_enthalpy (through a producing source
    of all mass components at -1 kg/s in a unit-volume cell: inflow = -(component flow fractions, enthalpy))"""
    L = wo.lib()
    rec = np.array([2.7e5, 130., 4., 4., 3., 1., 0., 0.,
                    935., 0., 0.8, 0., 0., 0., 5.461e5, 0.7, 0.3,
                    1.5, 0., 0.2, 0., 0., 0., 2.540e6, 0.4, 0.6])
    rock = np.array([1e-13, 1e-13, 1e-13, 2.5, 2.5, 1.0, 0.0, 0.0])      # porosity 1: balance = fluid sums
    bal = np.zeros(3)
    L.wo_cell_balance(wo.dp(rock), wo.dp(rec), 2, 2, 3, wo.dp(bal))
    assert np.allclose(bal[:2], [523.72, 224.58], rtol=1e-12)            # fluid_test.F90:125
    assert np.isclose(bal[2], 4.092448e8, rtol=1e-12)                    # :160
    # :197-230 (same record with viscosities, relative permeabilities and enthalpies)
    rec2 = np.array([2.7e5, 130., 4., 4., 3., 1., 0., 0.,
                     935., 1.e-6, 0.8, 0.7, 0., 83.9e3, 5.461e5, 0.7, 0.3,
                     1.5, 2.e-7, 0.2, 0.3, 0., 800.e3, 2.540e6, 0.4, 0.6])
    from waiwera_b200 import mesh as wmesh
    m = wmesh.structured(2, 1, 1, dx=1.0, gravity=(0.0, 0.0, 0.0), heterogeneous=False)
    f = wo.Flow(wo.make_params(eos=wo.EOS_WCE, gravity=(0.0, 0.0, 0.0)), m.ncell, m.ninterior, m.nowned,
                m.face_cells.reshape(-1), m.face_geom.reshape(-1), m.cell_geom.reshape(-1), m.rock.reshape(-1))
    f.set_sources([0], [0], [-1.0], [0.0])
    f.current_fluid()[:] = rec2                      # both cells: no flux between them, stored fluxes are zero
    rhs = np.zeros(6)
    assert L.wo_flow_cell_inflows(f.h, wo.dp(rhs)) == 0
    assert np.allclose(-rhs[:2], [0.6989722116, 0.3010277884], rtol=1e-9)    # component flow fractions
    assert np.isclose(-rhs[2], 86353.3307955843, rtol=1e-12)                 # specific enthalpy
    mob = np.array([935. * 0.7 / 1.e-6, 1.5 * 0.3 / 2.e-7])
    assert np.allclose(mob, [654500000., 2250000.]) and np.allclose(mob / mob.sum(), [0.9965740388, 0.0034259612])
    assert (rhs[3:] == 0.0).all()


# ---- test/unit/src/root_finder_test.F90:48-294 (Brent; the saturation-line search of the phase transitions) ----
class _RootFinder(C.Structure):
    _fields_ = [("interval", C.c_double * 2), ("root_tolerance", C.c_double), ("function_tolerance", C.c_double),
                ("root", C.c_double), ("max_iterations", C.c_int), ("iterations", C.c_int), ("err", C.c_int)]


_ROOT_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


def _find_root(wo, fn, interval=None):
    L = wo.lib()
    r = _RootFinder()
    L.wo_root_finder_init.argtypes = [C.POINTER(_RootFinder)]
    L.wo_root_finder_find.argtypes = [C.POINTER(_RootFinder), _ROOT_FN, C.c_void_p]
    L.wo_root_finder_init.restype = L.wo_root_finder_find.restype = None
    L.wo_root_finder_init(C.byref(r))
    if interval is not None:
        r.interval[0], r.interval[1] = interval
    cb = _ROOT_FN(lambda x, ctx: fn(x))
    L.wo_root_finder_find(C.byref(r), cb, None)
    return r


def test_root_finder(wo):
    import math
    r = _find_root(wo, lambda x: 0.5 - x)                                        # :48-91 linear
    assert r.err == 0 and abs(r.root - 0.5) <= r.root_tolerance and r.iterations <= 2
    assert _find_root(wo, lambda x: 0.5 - x, (0.75, 1.0)).err != 0                # interval not bracketed
    r = _find_root(wo, lambda x: (x - 0.75) ** 2 - 0.5)                          # :95-131 quadratic
    assert r.err == 0 and abs(r.root - (0.75 - math.sqrt(0.5))) <= r.root_tolerance and r.iterations <= 7
    r = _find_root(wo, lambda x: math.cos(x) - x ** 3, (0.0, 4.0))               # :135-174 Zhang
    assert r.err == 0 and abs(r.root - 0.8654740331015734) <= r.root_tolerance and r.iterations <= 12

    def invquad(x):                                                              # :178-222 inverse quadratic
        xs = x - 2.0 / 3.0
        return -math.sqrt(abs(xs)) if xs > 0 else math.sqrt(abs(xs))
    r = _find_root(wo, invquad, (-10.0, 10.0))
    assert r.err == 0 and abs(r.root - 2.0 / 3.0) <= r.root_tolerance and r.iterations <= 18
    # :226-294 saturation line intersection (IAPWS): from (20 bar, 210 degC) to (23 bar, 220 degC)
    th = wo.lib().wo_thermo_create(wo.THERMO_IAPWS, 0)

    def satdiff(x):
        P, T = 20.e5 + x * 3.e5, 210.0 + x * 10.0
        ps = C.c_double()
        wo.lib().wo_saturation_pressure(th, T, C.byref(ps))
        return ps.value - P
    r = _find_root(wo, satdiff)
    assert r.err == 0 and r.iterations <= 6
    assert abs((210.0 + r.root * 10.0) - 218.61315743282924) <= 10.0 * r.root_tolerance
    wo.lib().wo_thermo_destroy(th)


# ---- test/unit/src/interpolation_test.F90:73-296 (tables of the curves, rate tables of the sources) ----
DATA5_X, DATA5_Y = [0., 2.1, 3.7, 6.3, 8.9], [1., 2.0, 0.5, -1.1, -0.1]


class _Table(C.Structure):
    _fields_ = [("n", C.c_int), ("dim", C.c_int), ("index", C.c_int), ("x", C.POINTER(C.c_double)),
                ("val", C.POINTER(C.c_double))]


def test_interpolation_table(wo):
    L = wo.lib()
    t = _Table()
    x, v = np.array(DATA5_X), np.array(DATA5_Y)
    L.wo_table_init.argtypes = [C.POINTER(_Table), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int]
    L.wo_table_interpolate.argtypes = [C.POINTER(_Table), C.c_double, C.POINTER(C.c_double)]
    L.wo_table_destroy.argtypes = [C.POINTER(_Table)]
    L.wo_table_init.restype = L.wo_table_interpolate.restype = L.wo_table_destroy.restype = None
    L.wo_table_init(C.byref(t), wo.dp(x), wo.dp(v), 5, 1)
    for xq, expect, index in [(-0.5, 1.0, 0), (0.0, 1.0, 0), (1.0, 1.4761904761904763, 1), (4.5, 0.007692307692307665, 3),
                              (3.6, 0.59375, 2), (6.3, -1.1, 4), (10.0, -0.1, 5)]:      # :88-109
        y = C.c_double()
        L.wo_table_interpolate(C.byref(t), xq, C.byref(y))
        assert abs(y.value - expect) <= 1e-9 * max(abs(expect), 1.0), (xq, y.value)
        assert t.index == index, (xq, t.index)
    L.wo_table_destroy(C.byref(t))


def test_rate_table_averaging():
    """waiwera_b200.ingest.rates_at: endpoint and integral averaging of linear and step tables
    (interpolation_test.F90:223-373)"""
    from waiwera_b200 import ingest

    class P:
        pass
    tab = np.stack([DATA5_X, DATA5_Y], 1)
    cases = {
        ("linear", "endpoint"): [((-0.5, -0.1), 1.0), ((-0.5, 0.1), 1.0238095238095237), ((0.1, 2.0), 1.5),
                                 ((0.1, 3.0), 1.1019345238095237), ((3.1, 7.0), 0.11586538461538454),
                                 ((8.0, 12.0), -0.27307692307692316), ((1.0, 1.0), 1.4761904761904763)],
        ("step", "endpoint"): [((-0.5, -0.1), 1.0), ((-0.5, 0.1), 1.0), ((0.1, 2.0), 1.0), ((0.1, 3.0), 1.5),
                               ((3.1, 7.0), 0.45), ((8.0, 12.0), -0.6), ((1.0, 1.0), 1.0)],
        ("linear", "integrate"): [((-0.5, -0.1), 1.0), ((-0.5, 0.1), 1.003968253968254), ((0.1, 2.0), 1.5),
                                  ((0.1, 3.0), 1.5406660509031198), ((3.1, 7.0), -0.2530818540433925),
                                  ((8.0, 12.0), -0.1389423076923077), ((9.0, 12.0), -0.1), ((1.0, 1.0), 1.4761904761904763)],
        ("step", "integrate"): [((-0.5, -0.1), 1.0), ((-0.5, 0.1), 1.0), ((0.1, 2.0), 1.0), ((0.1, 3.0), 3.8 / 2.9),
                                ((3.1, 7.0), 1.73 / 3.9), ((8.0, 12.0), -0.325), ((1.0, 1.0), 1.0)]}
    for (interp, averaging), rows in cases.items():
        p = P()
        p.source_rates = np.zeros(1)
        p.source_tables = {0: (tab, interp, averaging)}
        for (t0, t1), expect in rows:
            assert abs(ingest.rates_at(p, t0, t1)[0] - expect) <= 1e-9, (interp, averaging, t0, t1)


def test_source_update_flow(wo):
    """test/unit/src/source_test.F90:53-194: source%update_flow for injection of either mass component or heat and
    for production of all components / one component / heat, from the two-phase two-component fluid record of the
    test; checked through cell_inflows on a unit-volume cell (inflow = flow / volume)"""
    from waiwera_b200 import mesh as wmesh
    rec = np.array([2.7e5, 130., 4., 4., 3., 1., 0., 0.,
                    935., 1.e-6, 0.8, 0.7, 0., 83.9e3, 5.461e5, 0.7, 0.3,
                    1.5, 2.e-7, 0.2, 0.3, 0., 800.e3, 2.540e6, 0.4, 0.6])
    m = wmesh.structured(2, 1, 1, dx=1.0, gravity=(0.0, 0.0, 0.0), heterogeneous=False)
    f = wo.Flow(wo.make_params(eos=wo.EOS_WCE, gravity=(0.0, 0.0, 0.0)), m.ncell, m.ninterior, m.nowned,
                m.face_cells.reshape(-1), m.face_geom.reshape(-1), m.cell_geom.reshape(-1), m.rock.reshape(-1))
    cases = [("inject 1", 10., 200.e3, 1, [10., 0., 2.e6]),
             ("inject 2", 5., 200.e3, 2, [0., 5., 1.e6]),
             ("inject heat", 1000., 0., 3, [0., 0., 1000.]),
             ("produce all", -5., 0., 0, [-3.4948610582, -1.5051389418, -431766.653977922]),
             ("produce 1", -5., 0., 1, [-5., 0., -431766.653977922]),
             ("produce heat", -5000., 0., 3, [0., 0., -5000.]),
             ("no flow 1", 0., 100.e3, 1, [0., 0., 0.])]
    for tag, rate, enthalpy, component, flow in cases:
        f.set_sources([0], [component], [rate], [enthalpy])
        f.current_fluid()[:] = rec
        rhs = np.zeros(6)
        assert wo.lib().wo_flow_cell_inflows(f.h, wo.dp(rhs)) == 0
        assert np.allclose(rhs[:3], flow, rtol=1e-9, atol=1e-12), (tag, rhs[:3])


def test_source_control_deliverability(wo):
    """test/unit/src/source_control_test.F90:226-610 (test_source_controls_pressure_reference.json), the controls this
    build has: source 1 (deliverability, production only) -12.8728519749 kg/s, source 2 (the same behind a 10 kg/s
    total limiter) -10, source 4 (productivity index calculated from an initial rate of -11) -11, source 6
    (deliverability, injection only: no flow) 0 -- in a two-phase cell at 50 bar, Sv = 0.8, kr = saturations"""
    from waiwera_b200 import mesh as wmesh
    L = wo.lib()
    th = L.wo_thermo_create(wo.THERMO_IAPWS, 0)
    P, sv = 50.e5, 0.8
    T = C.c_double()
    assert L.wo_saturation_temperature(th, P, C.byref(T)) == 0
    rec = np.zeros(26)
    rec[0], rec[1], rec[2], rec[3], rec[5] = P, T.value, 4.0, 4.0, 1.0
    rec[4] = L.wo_phase_composition(th, 4, P, T.value)
    for p, (sat, X) in enumerate([(1.0 - sv, [0.75, 0.25]), (sv, [0.9, 0.1])]):
        props = np.zeros(2)
        assert L.wo_region_properties(th, p + 1, wo.dp(np.array([P, T.value])), wo.dp(props)) == 0
        ph = rec[8 + 9 * p: 8 + 9 * (p + 1)]
        ph[0], ph[6] = props
        ph[1] = L.wo_region_viscosity(th, p + 1, T.value, P, props[0])
        ph[2] = ph[3] = sat
        ph[5] = props[1] + P / props[0]
        ph[7:9] = X
    L.wo_thermo_destroy(th)
    m = wmesh.structured(2, 1, 1, dx=1.0, gravity=(0.0, 0.0, 0.0), heterogeneous=False)
    f = wo.Flow(wo.make_params(eos=wo.EOS_WCE, gravity=(0.0, 0.0, 0.0)), m.ncell, m.ninterior, m.nowned,
                m.face_cells.reshape(-1), m.face_geom.reshape(-1), m.cell_geom.reshape(-1), m.rock.reshape(-1))
    f.set_sources([0, 0, 0, 0], [0, 0, 0, 0], [-1.0, -1.0, -11.0, -1.0], [0.0] * 4)
    f.current_fluid()[:] = rec
    mob = sum(rec[8 + 9 * p + 3] * rec[8 + 9 * p] / rec[8 + 9 * p + 1] for p in range(2))
    pi4 = 11.0 / (mob * (P - 2.0e5) * 1.0)                       # calculate_PI_from_rate :407-468
    f.set_source_controls([0, 1, 2, 3], [1e-12, 1e-12, pi4, 1e-12], [2.0e5] * 4, [1, 0, 0, 2], [0.0, 10.0, 0.0, 0.0])
    rhs = np.zeros(6)
    assert L.wo_flow_cell_inflows(f.h, wo.dp(rhs)) == 0
    rates = f.source_rates(4)
    assert np.allclose(rates, [-12.8728519749, -10.0, -11.0, 0.0], rtol=1e-9, atol=1e-12), rates
    # source 11: reference pressure tabulated against the flowing enthalpy (SRC_PRESSURE_TABLE_COORD_ENTHALPY)
    # -10.3366086953508 kg/s; the same table looked up at the cell's pressure (50 bar, beyond its last point: 0.5 bar)
    # gives 1e-12 * mobility * (50e5 - 0.5e5); with step interpolation the value at the point before the enthalpy
    table = [[0.0, 2200000.0], [1100000.0, 2000000.0], [2800000.0, 50000.0]]
    f.set_source_controls([0, 1, 2, 3], [1e-12] * 4, [2.0e5] * 4, [0] * 4, [0.0] * 4)
    assert f.set_source_pressure_table([0, 1, 2], [table] * 3, coordinate=[0, 1, 0], step=[0, 0, 1]) == 0
    assert L.wo_flow_cell_inflows(f.h, wo.dp(rhs)) == 0
    rates = f.source_rates(4)
    h = sum(rec[8 + 9 * p + 3] * rec[8 + 9 * p] / rec[8 + 9 * p + 1] * rec[8 + 9 * p + 5] for p in range(2)) / mob
    assert 1100000.0 < h < 2800000.0
    assert np.allclose(rates, [-10.3366086953508, -1e-12 * mob * (P - 0.5e5), -1e-12 * mob * (P - 2.0e6), -12.8728519749],
                       rtol=1e-9, atol=1e-12), rates


def test_source_separator_limiter_known_answers(wo):
    """test/unit/src/source_control_test.F90:226-610 again, the sources with a separator: source 3 (deliverability behind a
    10 bar separator and a steam limiter at 5 kg/s) -9.3081349399 kg/s with 5 kg/s of steam, source 16 (fixed -10 kg/s
    through a 15 / 5 bar two-stage separator) 5.60996474758954 kg/s of steam, source 17 (deliverability, 10 bar
    separator, limiter {"total": 10, "steam": 5}) -9.3081349399 / 5 -- same two-phase cell at 50 bar, Sv = 0.8"""
    from waiwera_b200 import mesh as wmesh
    L = wo.lib()
    th = L.wo_thermo_create(wo.THERMO_IAPWS, 0)
    P, sv = 50.e5, 0.8
    T = C.c_double()
    assert L.wo_saturation_temperature(th, P, C.byref(T)) == 0
    rec = np.zeros(26)
    rec[0], rec[1], rec[2], rec[3], rec[5] = P, T.value, 4.0, 4.0, 1.0
    rec[4] = L.wo_phase_composition(th, 4, P, T.value)
    for p, (sat, X) in enumerate([(1.0 - sv, [0.75, 0.25]), (sv, [0.9, 0.1])]):
        props = np.zeros(2)
        assert L.wo_region_properties(th, p + 1, wo.dp(np.array([P, T.value])), wo.dp(props)) == 0
        ph = rec[8 + 9 * p: 8 + 9 * (p + 1)]
        ph[0], ph[6] = props
        ph[1] = L.wo_region_viscosity(th, p + 1, T.value, P, props[0])
        ph[2] = ph[3] = sat
        ph[5] = props[1] + P / props[0]
        ph[7:9] = X
    L.wo_thermo_destroy(th)
    m = wmesh.structured(2, 1, 1, dx=1.0, gravity=(0.0, 0.0, 0.0), heterogeneous=False)
    f = wo.Flow(wo.make_params(eos=wo.EOS_WCE, gravity=(0.0, 0.0, 0.0)), m.ncell, m.ninterior, m.nowned,
                m.face_cells.reshape(-1), m.face_geom.reshape(-1), m.cell_geom.reshape(-1), m.rock.reshape(-1))
    f.set_sources([0, 0, 0], [0, 0, 0], [-1.0, -10.0, -1.0], [0.0] * 3)
    f.current_fluid()[:] = rec
    f.set_source_controls([0, 1, 2], [1e-12, 0.0, 1e-12], [2.0e5, 0.0, 2.0e5], [0, 0, 0], [0.0, 0.0, 10.0])
    assert f.set_source_separators([0, 1, 2], [[10.e5], [15.e5, 5.e5], [10.e5]], [0.0, 0.0, 0.0], [5.0, 0.0, 5.0]) == 0
    rhs = np.zeros(6)
    assert L.wo_flow_cell_inflows(f.h, wo.dp(rhs)) == 0
    rates = f.source_rates(3)
    assert np.allclose(rates, [-9.3081349399, -10.0, -9.3081349399], rtol=1e-9, atol=1e-12), rates
    steam = [f.source_separated(s, rates[s])[2] for s in range(3)]
    assert np.allclose(steam, [-5.0, -5.60996474758954, -5.0], rtol=1e-9, atol=1e-12), steam
    # sources 7, 8, 9, 12 of the same file: recharge 0.013 (P - 50.1 bar) both ways = +130; recharge 0.013 against
    # 49.99 bar, production only = -13; injectivity 0.011 against 49.99 bar, injection only = 0; recharge against the
    # cell's initial pressure = 0
    f.set_sources([0, 0, 0, 0], [1, 1, 1, 1], [0.0] * 4, [0.0] * 4)
    f.current_fluid()[:] = rec
    f.set_source_controls([0, 1, 2, 3], [0.0] * 4, [0.0] * 4, [0, 1, 2, 0], [0.0] * 4)
    f.set_source_recharge([0, 1, 2, 3], [0.013, 0.013, 0.011, 0.012], [50.1e5, 49.99e5, 49.99e5, P])
    assert L.wo_flow_cell_inflows(f.h, wo.dp(rhs)) == 0
    rates = f.source_rates(4)
    assert np.allclose(rates, [130.0, -13.0, 0.0, 0.0], rtol=1e-9, atol=1e-9), rates


def test_source_control_tables_known_answers(wo):
    """the sources of the same reference file whose control parameters are tables in time, evaluated over the test's
    interval [30, 120] s by ingest.controls_at (table averaging pinned by interpolation_test.F90): source 5 (productivity
    index table, endpoint averaging) -10.1910078135, source 10 (reference pressure table, step, endpoint: 1.8 bar)
    -12.9264888581701, sources 14 / 15 (rate factor tables, step / linear) 0.75 and 0.375 of -12.8728519749"""
    import json
    import os
    import shutil
    import tempfile
    from waiwera_b200 import ingest, mesh as wmesh
    inp = os.path.join(os.path.dirname(__file__), "golden", "inputs")
    doc = json.load(open(os.path.join(inp, "problem2a.input.json")))
    cell = doc["source"][0]["cell"]
    doc["source"] = [
        {"cell": cell, "direction": "out", "deliverability": {"productivity": {"time": [[0.0, 1e-12], [60.0, 9e-13], [90.0, 7e-13], [180.0, 5e-13]]}, "pressure": 200000.0}, "averaging": "endpoint"},
        {"cell": cell, "deliverability": {"productivity": 1e-12, "pressure": {"time": [[0.0, 200000.0], [60.0, 190000.0], [90.0, 160000.0], [180.0, 150000.0]]}}, "interpolation": "step", "averaging": "endpoint"},
        {"cell": cell, "direction": "production", "deliverability": {"productivity": 1e-12, "pressure": 200000.0}, "factor": {"time": [[0.0, 1.0], [30.0, 0.75], [120.0, 0.0]], "interpolation": "step"}},
        {"cell": cell, "direction": "production", "deliverability": {"productivity": 1e-12, "pressure": 200000.0}, "factor": [[0.0, 1.0], [30.0, 0.75], [120.0, 0.0]]},
    ]
    with tempfile.TemporaryDirectory() as d:
        for fn in os.listdir(inp):
            if fn.endswith(".msh"):
                shutil.copy(os.path.join(inp, fn), d)
        path = os.path.join(d, "in.json")
        json.dump(doc, open(path, "w"))
        p = ingest.load(path)
    ctrl, _ = ingest.controls_at(p, 30.0, 120.0)
    assert abs(ctrl[1]["reference_pressure"] - 1.8e5) < 1e-6          # asserted by the reference test itself
    L = wo.lib()
    th = L.wo_thermo_create(wo.THERMO_IAPWS, 0)
    P, sv = 50.e5, 0.8
    T = C.c_double()
    assert L.wo_saturation_temperature(th, P, C.byref(T)) == 0
    rec = np.zeros(26)
    rec[0], rec[1], rec[2], rec[3], rec[5] = P, T.value, 4.0, 4.0, 1.0
    rec[4] = L.wo_phase_composition(th, 4, P, T.value)
    for q, (sat, X) in enumerate([(1.0 - sv, [0.75, 0.25]), (sv, [0.9, 0.1])]):
        props = np.zeros(2)
        assert L.wo_region_properties(th, q + 1, wo.dp(np.array([P, T.value])), wo.dp(props)) == 0
        ph = rec[8 + 9 * q: 8 + 9 * (q + 1)]
        ph[0], ph[6] = props
        ph[1] = L.wo_region_viscosity(th, q + 1, T.value, P, props[0])
        ph[2] = ph[3] = sat
        ph[5] = props[1] + P / props[0]
        ph[7:9] = X
    L.wo_thermo_destroy(th)
    m = wmesh.structured(2, 1, 1, dx=1.0, gravity=(0.0, 0.0, 0.0), heterogeneous=False)
    f = wo.Flow(wo.make_params(eos=wo.EOS_WCE, gravity=(0.0, 0.0, 0.0)), m.ncell, m.ninterior, m.nowned,
                m.face_cells.reshape(-1), m.face_geom.reshape(-1), m.cell_geom.reshape(-1), m.rock.reshape(-1))
    f.set_sources([0] * 4, [0] * 4, [-1.0] * 4, [0.0] * 4)
    f.current_fluid()[:] = rec
    f.set_source_controls([c["source"] for c in ctrl], [c["productivity"] for c in ctrl], [c["reference_pressure"] for c in ctrl],
                          [c["direction"] for c in ctrl], [c["limit"] for c in ctrl])
    rhs = np.zeros(6)
    assert L.wo_flow_cell_inflows(f.h, wo.dp(rhs)) == 0
    rates = f.source_rates(4)
    expect = [-10.1910078135, -12.9264888581701, -12.8728519749 * 0.75, -12.8728519749 * 0.375]
    assert np.allclose(rates, expect, rtol=1e-9, atol=1e-12), rates


def test_eos_scaling(wo):
    """test/unit/src/eos_test.F90:94-200: eos%scale / eos%unscale of every EOS of this build with default, user and
    adaptive (partial pressure / pressure) scales"""
    L = wo.lib()
    cases = [
        (wo.EOS_W, {}, [([3.e5], 1, [0.3]), ([4.e5], 2, [0.4])]),
        (wo.EOS_W, dict(pressure_scale=1.e5), [([3.e5], 1, [3.0])]),
        (wo.EOS_WE, {}, [([1.e5, 20.], 1, [0.1, 0.2]), ([0.9e5, 100.], 2, [0.09, 1.]), ([13.e5, 0.4], 4, [1.3, 0.4])]),
        (wo.EOS_WE, dict(pressure_scale=1.e7, temperature_scale=200.),
         [([1.e5, 20.], 1, [0.01, 0.1]), ([0.9e5, 100.], 2, [0.009, 0.5]), ([13.e5, 0.4], 4, [0.13, 0.4])]),
        (wo.EOS_WCE, {}, [([15.e5, 40., 3.e5], 1, [1.5, 0.4, 0.2]), ([0.8e5, 110., 0.6e5], 2, [0.08, 1.1, 0.75]),
                          ([20.e5, 0.4, 10.e5], 4, [2.0, 0.4, 0.5]), ([100.e5, 0.5, 50.e5], 4, [10., 0.5, 0.5])]),
        (wo.EOS_WCE, dict(pressure_scale=1.e7, temperature_scale=200., partial_pressure_scale=1.e5),
         [([15.e5, 40., 2.e5], 1, [0.15, 0.2, 2.]), ([0.7e5, 110., 0.6e5], 2, [0.007, 0.55, 0.6]),
          ([13.e5, 0.4, 10.e5], 4, [0.13, 0.4, 10.0])]),
        (wo.EOS_WAE, dict(pressure_scale=1.e7, temperature_scale=200., partial_pressure_scale=1.e5),
         [([15.e5, 40., 2.e5], 1, [0.15, 0.2, 2.]), ([0.7e5, 110., 0.6e5], 2, [0.007, 0.55, 0.6]),
          ([13.e5, 0.4, 10.e5], 4, [0.13, 0.4, 10.0])])]
    for eos_id, scales, rows in cases:
        prm = wo.make_params(eos=eos_id)
        for k, v in scales.items():
            setattr(prm, k, v)
        eos = L.wo_eos_create(C.byref(prm))
        try:
            for primary, region, expect in rows:
                pr = np.array(primary)
                y, back = np.zeros(len(pr)), np.zeros(len(pr))
                L.wo_eos_scale(eos, wo.dp(pr), region, wo.dp(y))
                assert np.allclose(y, expect, rtol=1e-12), (eos_id, scales, primary, y)
                L.wo_eos_unscale(eos, wo.dp(y), region, wo.dp(back))
                assert np.allclose(back, pr, rtol=1e-12)
        finally:
            L.wo_eos_destroy(eos)
