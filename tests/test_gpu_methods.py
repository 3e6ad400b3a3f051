"""a11: the three residual forms timestepper.F90 selects -- backward Euler (:345-374), variable-step BDF2 (:378-427)
and direct steady state (:431-452) -- through wb_set_method: residual and local-FD Jacobian vs the oracle's
residual and colour-by-colour FD Jacobian of the same form."""
import numpy as np
import pytest

from util import SEED, make_problem, oracle_flow, gpu_flow, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("method", [1, 2])
def test_method_residual_and_jacobian_match_oracle(wo, method):
    from waiwera_b200 import flow
    m, y, region, prm = make_problem(wo, two_phase_layers=2, top_boundary=True)
    ref = oracle_flow(wo, m, prm, y, region)
    sim = gpu_flow(wo, flow, m, prm, y, region)
    rng = np.random.default_rng(SEED + 31)
    _, L2 = ref.lhs(y * (1 + 2e-4 * rng.uniform(-1, 1, len(y))))   # "two steps back"
    _, L1 = ref.lhs(y * (1 + 1e-4 * rng.uniform(-1, 1, len(y))))   # "last step"
    _, L1 = ref.lhs(y)
    sim.lhs(y)
    dt, dt_last = 1.0e6, 0.6e6
    ref.set_method(method, dt_last, L2)
    assert sim.set_method(method, dt_last, L2) == 0
    y2 = y * (1 + 1e-4 * rng.uniform(-1, 1, len(y)))
    e0, lhs0, rhs0, r0 = ref.residual(y2, L1, dt)
    e1, lhs1, rhs1, r1 = sim.residual(y2, L1, dt)
    assert e0 == e1 == 0
    assert relerr(r1, r0) < 1e-10
    if method == 2:
        assert np.array_equal(r1, rhs1)                       # r = R
    else:
        q = dt / dt_last
        expect = (1 + 2 * q) * lhs1 - (q + 1) ** 2 * L1 + q * q * L2 - dt * (q + 1) * rhs1
        assert relerr(r1, expect) < 1e-12
    # Jacobian of the same form
    A = ref.bsr()
    nb = A.contents.nb
    color = np.zeros(nb, np.int32)
    nc = wo.lib().wo_bsr_coloring(A, wo.ip(color))
    assert wo.lib().wo_fd_jacobian(ref.h, wo.dp(y2), wo.dp(L1), dt, wo.dp(r0), wo.ip(color), nc, 1e-8, 1e-2, A) == 0
    rowptr, colidx, val = [a.copy() for a in wo.bsr_arrays(A)]
    assert sim.jacobian(y2, L1, dt) == 0
    Jl = sim.jacobian_values()
    rows = np.repeat(np.arange(nb), np.diff(rowptr))
    rowmax = np.zeros((nb, 2))
    for ii in range(2):
        np.maximum.at(rowmax[:, ii], rows, np.abs(val[:, [ii, 2 + ii]]).max(axis=1))
    scale = np.stack([rowmax[rows, 0], rowmax[rows, 1], rowmax[rows, 0], rowmax[rows, 1]], 1)
    assert (np.abs(Jl - val) / np.maximum(scale, 1e-300)).max() < 5e-6
    assert sim.jacobian(y2, L1, dt, colored=True) == 0
    Jc = sim.jacobian_values()
    assert (np.abs(Jc - Jl) / np.maximum(scale, 1e-300)).max() < 1e-12
    # back to backward Euler: the default form is restored
    sim.set_method(0)
    ref.set_method(0)
    assert relerr(sim.residual(y2, L1, dt)[3], ref.residual(y2, L1, dt)[3]) < 1e-10
    wo.lib().wo_bsr_destroy(A)
    sim.destroy()
