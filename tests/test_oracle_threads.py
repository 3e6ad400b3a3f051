"""The oracle's OpenMP loops (owner-computes cell chunks, per-thread EOS instances) must give bit-identical results
for every thread count: the CPU arm of bench.py runs them on all host cores, the parity tests compare against them."""
import ctypes as C

import numpy as np
import pytest

from util import make_problem, make_problem_wce, oracle_flow


@pytest.mark.parametrize("eos", ["we", "wce"])
def test_residual_and_jacobian_do_not_depend_on_thread_count(wo, eos):
    dims = (30, 30, 25)  # 22 500 cells: above the oracle's serial cut-off (WO_PAR_MIN)
    if eos == "we":
        m, y, region, prm = make_problem(wo, dims=dims, two_phase_layers=3)
    else:
        m, y, region, prm = make_problem_wce(wo, dims=dims, two_phase_layers=3)
    L = wo.lib()
    out = {}
    for nt in (1, 3, 8):
        assert L.wo_set_num_threads(nt) == nt
        f = oracle_flow(wo, m, prm, y, region)
        e, L0 = f.lhs(y)
        assert e == 0
        e, lhs, rhs, r = f.residual(y * (1.0 + 1e-5), L0, 1.0e5)
        assert e == 0
        A = f.bsr()
        color = np.zeros(A.contents.nb, np.int32)
        nc = L.wo_bsr_coloring(A, wo.ip(color))
        y1 = np.ascontiguousarray(y * (1.0 + 1e-5))
        assert L.wo_fd_jacobian(f.h, wo.dp(y1), wo.dp(L0), 1.0e5, wo.dp(r), wo.ip(color), nc, 1e-8, 1e-2, A) == 0
        vals = wo.bsr_arrays(A)[2].copy()
        # a transition sweep on a perturbed state
        ynew = np.ascontiguousarray(y * (1.0 + 2e-3))
        search = np.ascontiguousarray(y - ynew)
        cs, cy = C.c_int(), C.c_int()
        L.wo_flow_pre_iteration(f.h)
        et = L.wo_flow_fluid_transitions(f.h, wo.dp(y), wo.dp(search), wo.dp(ynew), C.byref(cs), C.byref(cy))
        out[nt] = (L0, lhs, rhs, r, vals, ynew.copy(), search.copy(), cs.value, cy.value, et, f.regions().copy())
        L.wo_bsr_destroy(A)
    import os
    L.wo_set_num_threads(os.cpu_count() or 1)
    for nt in (3, 8):
        for a, b in zip(out[1], out[nt]):
            assert np.array_equal(a, b)
