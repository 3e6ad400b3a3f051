#!/usr/bin/env python3
"""Kernel-level timing of the Krylov pieces on the bench problem (run under gpurun):
   python tools/microbench.py [--dims 100 100 100] [--cube 10]
Times with CUDA events (wb timers) the SpMV, the PC apply and whole GMRES solves capped at a fixed
iteration count with each preconditioner, so the per-iteration cost can be split into SpMV / PC /
Gram-Schmidt shares without a profiler."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs=3, default=[100, 100, 100])
    ap.add_argument("--cube", type=int, default=10)
    ap.add_argument("--its", type=int, default=300)
    a = ap.parse_args()
    import torch
    import bench
    from waiwera_b200 import flow, mesh as wmesh, _lib
    L = _lib.lib()
    m, y, region = bench.build_problem(tuple(a.dims))
    sim = flow.FlowSimulation(flow.make_params(), m)
    assert sim.fluid_init(y, region) == 0
    err, L0 = sim.lhs(y)
    assert sim.jacobian(y, L0, bench.DT) == 0
    J = sim.jacobian_mat()
    n = sim.n
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    z = torch.empty_like(x)
    out = {}

    def timed(name, fn, reps=30):
        for _ in range(3):
            fn()
        L.wb_timer_reset(sim.h)
        for _ in range(reps):
            fn()
        return sim.timer(name)

    t, c = timed("mat_mult", lambda: J.mult(x, z))
    out["spmv_us"] = 1e3 * t / c
    for label, pct, bor in (("ilu0_cube%d" % a.cube, flow.PC_BJACOBI_ILU0, wmesh.cube_blocks(m, a.cube)),
                            ("pbjacobi", flow.PC_PBJACOBI, None), ("none", flow.PC_NONE, None)):
        pc = flow.PC(J, pct, 1, bor)
        t, c = timed("pc_apply", lambda: pc.apply(x, z))
        out["pc_apply_%s_us" % label] = 1e3 * t / c
        b = torch.randn(n, dtype=torch.float64, device="cuda")
        sol = torch.empty_like(b)
        for ksp in (flow.KSP_GMRES, flow.KSP_BCGS):
            o = flow.ksp_opts(type=ksp, maxit=a.its, rtol=1e-30)
            flow.ksp_solve(J, pc, b, sol, o)
            L.wb_timer_reset(sim.h)
            reason, its, rn = flow.ksp_solve(J, pc, b, sol, o)
            t, c = sim.timer("ksp_solve")
            out["%s_%s_us_per_it" % ("gmres" if ksp == flow.KSP_GMRES else "bcgs", label)] = 1e3 * t / max(its, 1)
        pc.destroy()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
