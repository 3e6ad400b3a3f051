"""Separators and limiters on the separated water / steam flows of a source (the part of the
source network that acts on a single source).  Reference: src/separator.F90 (stage flash :108-166, multi-stage :212-260),
source_network_node_limit_rate (src/source_network_node.F90:245-315), input syntax src/source_setup.F90:2255-2330,
3117-3276.  The oracle is pinned by the reference's own known answers (test/unit/src/separator_test.F90:55-150); the
CUDA path is compared with the oracle through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from util import make_problem, oracle_flow, gpu_flow, relerr

P1 = [10.0e5]
P2 = [1.45e6, 0.55e6]
# separator pressures for the flow tests: below the pressure of the producing cells (4.7 bar, 150 degC), so that the
# produced fluid really flashes
Q1 = [1.0e5]
Q2 = [2.0e5, 0.5e5]


def stage_h(wo, pressures, thermo=0):
    th = wo.lib().wo_thermo_create(thermo, 0)
    h = np.zeros(2 * len(pressures))
    for i, p in enumerate(pressures):
        a, b = C.c_double(), C.c_double()
        assert wo.lib().wo_separator_stage(th, p, C.byref(a), C.byref(b)) == 0
        h[2 * i], h[2 * i + 1] = a.value, b.value
    return h


@pytest.mark.parametrize("pressures,two_phase", [
    (P1, (0.21709153586628488, -7.829084641337152, 762682.8443354106, -2.1709153586628487, 2777119.5376846623)),
    (P2, (0.256210105124, -7.437898948764703, 655876.6515067405, -2.5621010512352966, 2779615.4799612807))])
def test_separator_known_answers(wo, pressures, two_phase):
    """separator_test.F90:55-150: -10 kg/s at 500, 3000 and 1200 kJ/kg through a 10 bar separator and through a two-stage
    14.5 / 5.5 bar separator: steam fraction, water rate and enthalpy, steam rate and enthalpy"""
    h = stage_h(wo, pressures)
    out = np.zeros(5)

    def sep(rate, enthalpy):
        wo.lib().wo_separate(len(pressures), wo.dp(h), rate, enthalpy, wo.dp(out))
        return out[4], out[0], out[1], out[2], out[3]
    assert sep(-10.0, 500.0e3) == (0.0, -10.0, 500.0e3, 0.0, 0.0)                 # all water
    assert sep(-10.0, 3000.0e3) == (1.0, 0.0, 0.0, -10.0, 3000.0e3)               # all steam
    got = sep(-10.0, 1200.0e3)
    for g, e in zip(got, two_phase):
        assert abs(g - e) <= 1e-9 * abs(e)


def setup(wo, flow=None):
    """3 x 3 x 4 column mesh with two two-phase layers; three sources in a two-phase cell, a liquid cell, and one injector"""
    m, y, region, prm = make_problem(wo, dims=(3, 3, 4), two_phase_layers=2)
    tp = int(np.flatnonzero(region == 4)[0])
    lq = int(np.flatnonzero(region == 1)[0])
    cells, comps, rates, enth = [tp, lq, tp, tp], [0, 0, 1, 0], [-5.0, -4.0, 3.0, -6.0], [0.0, 0.0, 4.0e5, 0.0]
    f = oracle_flow(wo, m, prm, y, region)
    f.set_sources(cells, comps, rates, enth)
    sim = None
    if flow is not None:
        sim = gpu_flow(wo, flow, m, prm, y, region)
        assert sim.set_sources(cells, comps, rates, enth) == 0
    return m, y, region, f, sim, rates


def evaluate(f, y, n):
    e, L0 = f.lhs(y)
    assert e == 0
    out = f.residual(y, L0, 1.0e4)
    assert out[0] == 0
    return f.source_rates(n), out[-1]


CASES = [
    # pressures per source, water limits, steam limits, total limits
    dict(p=[Q1, Q1, Q1, Q2], lw=[0, 0, 0, 0], ls=[0.5, 0.5, 0.5, 0.5], lt=[0, 0, 0, 0]),      # steam limiters
    dict(p=[Q1, Q1, Q1, Q2], lw=[1.0, 1.0, 1.0, 2.0], ls=[0, 0, 0, 0], lt=[0, 0, 0, 0]),      # water limiters
    dict(p=[Q2, Q1, [], Q1], lw=[3.0, 0, 1.0, 1.5], ls=[0.4, 0.1, 1.0, 5.0], lt=[4.0, 3.0, 1.0, 0]),  # several types at once
    dict(p=[Q1, [], [], []], lw=[0, 0, 0, 0], ls=[0, 0, 0, 0], lt=[0, 2.0, 0, 0]),            # separator without limiter
    # recharge controls (rate = -c (P - Pref)): producing through a separator with a steam limiter, injecting, blocked
    # by its direction, and limited in total
    dict(p=[Q1, [], [], []], lw=[0, 0, 0, 0], ls=[0.3, 0, 0, 0], lt=[0, 0, 0, 2.5], dir=[0, 0, 1, 0],
         recharge=([0, 1, 2, 3], [2.0e-5, 1.0e-5, 1.0e-5, 3.0e-5], [1.0e5, 9.0e5, 9.0e5, 1.0e5])),
]


def apply_case(obj, case, n):
    obj.set_source_controls(list(range(n)), [0.0] * n, [0.0] * n, case.get("dir", [0] * n), case["lt"])
    if "recharge" in case:
        obj.set_source_recharge(*case["recharge"])
    r = obj.set_source_separators(list(range(n)), case["p"], case["lw"], case["ls"])
    assert r in (0, None)


@pytest.mark.parametrize("case", CASES)
def test_oracle_separated_limiters(wo, case):
    """the limited rate is the fixed rate times the smallest limit / |flow| over the limited flow types; injection and
    sources without a separator have no separated flows"""
    m, y, region, f, _, rates = setup(wo)
    n = len(rates)
    base, _ = evaluate(f, y, n)
    assert np.array_equal(base, rates)
    apply_case(f, case, n)
    got, _ = evaluate(f, y, n)
    if "recharge" in case:           # the rate before the limiters: -c (P - Pref), then the direction control
        fl = f.fluid()
        cells = [int(np.flatnonzero(region == 4)[0]), int(np.flatnonzero(region == 1)[0])]
        cells = [cells[0], cells[1], cells[0], cells[0]]
        rates = list(rates)
        for s, c, pr in zip(*case["recharge"]):
            rates[s] = -c * (fl[cells[s]][0] - pr)
            d = case["dir"][s]
            if (d == 1 and not rates[s] < 0) or (d == 2 and not rates[s] > 0):
                rates[s] = 0.0
        assert rates[0] < 0 and rates[1] > 0 and rates[2] == 0.0
    for s in range(n):
        sep0 = f.source_separated(s, rates[s])
        has_sep = len(case["p"][s]) > 0 and rates[s] < 0
        assert (sep0[0] != 0.0 or sep0[2] != 0.0) == has_sep
        if has_sep:
            assert abs(sep0[0] + sep0[2] - rates[s]) <= 1e-12 * abs(rates[s])      # water + steam = total
        scale = 1.0
        for lim, flowrate in ((case["lt"][s], rates[s]), (case["lw"][s], sep0[0]), (case["ls"][s], sep0[2])):
            if lim > 0 and abs(flowrate) > lim:
                scale = min(scale, lim / abs(flowrate))
        assert abs(got[s] - rates[s] * scale) <= 1e-14 * abs(rates[s])
        sep1 = f.source_separated(s, got[s])
        for lim, flowrate in ((case["lw"][s], sep1[0]), (case["ls"][s], sep1[2])):
            if lim > 0 and has_sep:
                assert abs(flowrate) <= lim * (1 + 1e-12)
    # the two-phase producer really exercises the flash: its steam fraction is strictly between 0 and 1
    sf = f.source_separated(0, rates[0])[4] if len(case["p"][0]) else 0.5
    assert 0.0 < sf < 1.0


def test_ingest_separator_and_limiter_syntax():
    from waiwera_b200 import ingest
    import json, os, shutil, tempfile
    inp = os.path.join(os.path.dirname(__file__), "golden", "inputs")
    doc = json.load(open(os.path.join(inp, "problem2a.input.json")))
    cell = doc["source"][0]["cell"]
    doc["source"] = [
        {"cell": cell, "rate": -5.0, "separator": {"pressure": [1.45e6, 0.55e6]}, "limiter": {"steam": 2.0, "total": 10.0}},
        {"cell": cell, "rate": -4.0, "limiter": {"type": "water", "limit": 1.5, "separator_pressure": 8.0e5}},
        {"cell": cell, "rate": -3.0, "separator": True},
        {"cell": cell, "rate": -2.0, "limiter": {"limit": 1.0}},
    ]
    with tempfile.TemporaryDirectory() as d:
        for fn in os.listdir(inp):
            if fn.endswith(".msh"):
                shutil.copy(os.path.join(inp, fn), d)
        path = os.path.join(d, "in.json")
        json.dump(doc, open(path, "w"))
        p = ingest.load(path)
    seps = {s["source"]: s for s in p.source_separators}
    assert seps[0] == dict(source=0, pressure=[1.45e6, 0.55e6], limit_water=0.0, limit_steam=2.0)
    assert seps[1] == dict(source=1, pressure=[8.0e5], limit_water=1.5, limit_steam=0.0)
    assert seps[2] == dict(source=2, pressure=[0.55e6], limit_water=0.0, limit_steam=0.0)
    assert 3 not in seps
    lim = {c["source"]: c["limit"] for c in p.source_controls}
    assert lim[0] == 10.0 and lim[1] == 0.0 and lim[3] == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_separated_limiters_match_oracle(wo, case):
    """rates, separated-flow outputs, the residual and the finite-difference Jacobian (which sees the limiter through the
    perturbed evaluations) against the oracle"""
    from waiwera_b200 import flow
    m, y, region, f, sim, rates = setup(wo, flow)
    n = len(rates)
    apply_case(f, case, n)
    apply_case(sim, case, n)
    ref_rates, ref_res = evaluate(f, y, n)
    e, L0 = sim.lhs(y)
    e, _, _, res = sim.residual(y, L0, 1.0e4)
    assert e == 0
    got = sim.source_rates()
    assert np.abs(got - ref_rates).max() <= 1e-13 * np.abs(ref_rates).max()
    assert relerr(res, ref_res) < 1e-12
    sep = sim.source_separated()
    for s in range(n):
        assert np.allclose(sep[s], f.source_separated(s, ref_rates[s]), rtol=1e-12, atol=1e-300)
    for pr in (P1[0], Q1[0], Q2[1]):
        hw, hs = sim.separator_stage(pr)
        assert np.allclose([hw, hs], stage_h(wo, [pr]), rtol=1e-14, atol=0)
    # Jacobian: the limited source's cell block differs from the unlimited one and matches the oracle's
    assert sim.jacobian(y, L0, 1.0e4) == 0
    J = sim.jacobian_values()
    A = f.bsr()
    color = np.zeros(A.contents.nb, np.int32)
    ncolor = wo.lib().wo_bsr_coloring(A, wo.ip(color))
    F0 = f.residual(y, f.lhs(y)[1], 1.0e4)[-1]
    assert wo.lib().wo_fd_jacobian(f.h, wo.dp(y), wo.dp(f.lhs(y)[1]), 1.0e4, wo.dp(F0), wo.ip(color), ncolor, 1e-8, 1e-2, A) == 0
    assert relerr(J, wo.bsr_arrays(A)[2]) < 1e-7
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()
