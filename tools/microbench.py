#!/usr/bin/env python3
"""Kernel-level timing of the Krylov pieces on the bench problem (run under gpurun):
   python tools/microbench.py [--config 2] [--cube 10] [--global-ilu] [--restarts 30 60 100 200]
Times with CUDA events (wb timers) the SpMV, the PC apply and whole GMRES solves capped at a fixed
iteration count with each preconditioner, so the per-iteration cost can be split into SpMV / PC /
Gram-Schmidt shares without a profiler; optionally the global (one sub-domain) ILU(0) and a restart sweep of the
full solve of the Newton system (iterations and ms to rtol 1e-5)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--dims", type=int, nargs=3, default=None)
    ap.add_argument("--cube", type=int, default=10)
    ap.add_argument("--its", type=int, default=300)
    ap.add_argument("--global-ilu", action="store_true")
    ap.add_argument("--restarts", type=int, nargs="*", default=[])
    ap.add_argument("--skip-pcs", action="store_true")
    ap.add_argument("--fused-only", action="store_true", help="only the full GMRES(30) solve with the persistent kernel")
    ap.add_argument("--asm", action="store_true", help="PCASM (restricted, overlap 1) + ILU(0) on the same sub-domains")
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
    n = sim.n
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    z = torch.empty_like(x)
    out = {"workload": prob.name(argparse.Namespace(ksp="gmres", restart=30))}
    L.wb_timers_enable(1)

    def timed(name, fn, reps=30):
        for _ in range(3):
            fn()
        L.wb_timer_reset(sim.h)
        for _ in range(reps):
            fn()
        return sim.timer(name)

    t, c = timed("mat_mult", lambda: J.mult(x, z))
    out["spmv_us"] = 1e3 * t / c
    pcs = [("ilu0_cube%d" % a.cube, flow.PC_BJACOBI_ILU0, prob.blocks(m, a.cube), 30)]
    if not a.skip_pcs:
        pcs += [("pbjacobi", flow.PC_PBJACOBI, None, 30), ("none", flow.PC_NONE, None, 30)]
    if a.global_ilu:
        pcs.append(("ilu0_global", flow.PC_BJACOBI_ILU0, None, 5))
    if a.asm:
        pcs.append(("asm1_ilu0_cube%d" % a.cube, flow.PC_ASM_ILU0, prob.blocks(m, a.cube), 30))
    b = torch.from_numpy(F0).cuda()
    sol = torch.empty_like(b)
    for label, pct, bor, reps in pcs:
        if label.startswith("ilu0_cube"):
            # the persistent kernel (one launch per solve) against the launch-per-operation solver
            for fused in ((2,) if a.fused_only else (2, 0)):
                L.wb_ksp_set_fused(fused)
                pcf = flow.PC(J, pct, 1, bor)
                o = flow.ksp_opts(type=flow.KSP_GMRES, maxit=20000, rtol=1e-5)
                flow.ksp_solve(J, pcf, b, sol, o)
                sim.ksp_breakdown()
                L.wb_timer_reset(sim.h)
                reason, its, rn = flow.ksp_solve(J, pcf, b, sol, o)
                t, c = sim.timer("ksp_solve")
                out["gmres30_full_%s" % ("fused" if fused else "unfused")] = {
                    "its": its, "reason": reason, "rnorm": rn, "ms": t, "us_per_it": 1e3 * t / max(its, 1),
                    "breakdown_ctas_min_mean_max": sim.ksp_breakdown_ctas(), "breakdown_us": sim.ksp_breakdown()}
                pcf.destroy()
            L.wb_ksp_set_fused(1)
            if a.fused_only:
                continue
        pc = flow.PC(J, pct, 1, bor)
        t, c = timed("pc_apply", lambda: pc.apply(x, z), reps)
        out["pc_apply_%s_us" % label] = 1e3 * t / c
        L.wb_timer_reset(sim.h)
        pc.refactor()
        t, c = sim.timer("pc_setup")
        out["pc_setup_%s_ms" % label] = t / max(c, 1)
        for ksp in (flow.KSP_GMRES, flow.KSP_BCGS):
            if label == "ilu0_global" and ksp == flow.KSP_BCGS:
                continue
            o = flow.ksp_opts(type=ksp, maxit=a.its if label != "ilu0_global" else 60, rtol=1e-30)
            flow.ksp_solve(J, pc, b, sol, o)
            L.wb_timer_reset(sim.h)
            reason, its, rn = flow.ksp_solve(J, pc, b, sol, o)
            t, c = sim.timer("ksp_solve")
            out["%s_%s_us_per_it" % ("gmres" if ksp == flow.KSP_GMRES else "bcgs", label)] = 1e3 * t / max(its, 1)
        if label.startswith("asm1"):
            # full solves to rtol 1e-5 (launch-per-operation solver; the persistent kernel has no overlap support)
            for ksp, nm in ((flow.KSP_GMRES, "gmres30"), (flow.KSP_BCGS, "bcgs")):
                o = flow.ksp_opts(type=ksp, maxit=20000, rtol=1e-5)
                flow.ksp_solve(J, pc, b, sol, o)
                L.wb_timer_reset(sim.h)
                reason, its, rn = flow.ksp_solve(J, pc, b, sol, o)
                t, c = sim.timer("ksp_solve")
                out["%s_full_%s" % (nm, label)] = {"its": its, "reason": reason, "rnorm": rn, "ms": t, "us_per_it": 1e3 * t / max(its, 1)}
        if label.startswith("ilu0_cube"):
            for rs in a.restarts:
                o = flow.ksp_opts(type=flow.KSP_GMRES, restart=rs, maxit=20000, rtol=1e-5)
                flow.ksp_solve(J, pc, b, sol, o)
                L.wb_timer_reset(sim.h)
                reason, its, rn = flow.ksp_solve(J, pc, b, sol, o)
                t, c = sim.timer("ksp_solve")
                out["gmres_restart_%d" % rs] = {"its": its, "reason": reason, "ms": t, "us_per_it": 1e3 * t / max(its, 1)}
            if a.restarts:
                o = flow.ksp_opts(type=flow.KSP_BCGS, maxit=20000, rtol=1e-5)
                L.wb_timer_reset(sim.h)
                reason, its, rn = flow.ksp_solve(J, pc, b, sol, o)
                t, c = sim.timer("ksp_solve")
                out["bcgs_full"] = {"its": its, "reason": reason, "ms": t, "us_per_it": 1e3 * t / max(its, 1)}
        pc.destroy()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
