#!/usr/bin/env python3
"""One GMRES solve of a bench configuration's Newton system with the persistent kernel (or WB_FUSED=0: the
launch-per-operation solver), capped at --maxit iterations: per-iteration time and the kernel's phase breakdown.
Small enough to run under ncu:  ncu --set full --import-source on -k regex:k_gmres_fused -c 1 python tools/fused_probe.py --dims 50 50 50"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--dims", type=int, nargs=3, default=None)
    ap.add_argument("--cube", type=int, default=10)
    ap.add_argument("--maxit", type=int, default=600)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    import torch
    import bench
    from waiwera_b200 import flow, _lib
    L = _lib.lib()
    prob = bench.Problem(a.config, 1, a.dims)
    m, y, region = prob.mesh, prob.y, prob.region
    sim = flow.FlowSimulation(flow.make_params(eos=prob.eos), m)
    assert sim.fluid_init(y, region) == 0
    err, L0 = sim.lhs(y)
    e, _, _, F0 = sim.residual(y, L0, prob.dt)
    assert sim.jacobian(y, L0, prob.dt) == 0
    J = sim.jacobian_mat()
    b = torch.from_numpy(F0).cuda()
    sol = torch.empty_like(b)
    pc = flow.PC(J, flow.PC_BJACOBI_ILU0, 1, prob.blocks(m, a.cube))
    o = flow.ksp_opts(type=flow.KSP_GMRES, maxit=a.maxit, rtol=1e-30)
    L.wb_timers_enable(1)
    out = {}
    for rep in range(a.reps):
        sim.ksp_breakdown()
        L.wb_timer_reset(sim.h)
        reason, its, rn = flow.ksp_solve(J, pc, b, sol, o)
        t, c = sim.timer("ksp_solve")
        out = {"dims": prob.dims, "its": its, "us_per_it": round(1e3 * t / max(its, 1), 2), "ctas": sim.ksp_breakdown_ctas(),
               "cta0": sim.ksp_breakdown()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
