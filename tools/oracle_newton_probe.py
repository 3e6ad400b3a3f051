#!/usr/bin/env python3
"""Probe of the CPU oracle's Newton / Krylov behaviour on a bench configuration (tuning aid for the synthetic
configurations' time step; test infrastructure, never the product path).
  python tools/oracle_newton_probe.py --config 4 --dims 50 50 25 --dt 1e3 1e4 1e5 --ksp bcgs gmres"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--dims", type=int, nargs=3, default=None)
    ap.add_argument("--dt", type=float, nargs="*", default=[1e5])
    ap.add_argument("--ksp", nargs="*", default=["gmres"])
    ap.add_argument("--maxit", type=int, default=4)
    ap.add_argument("--ksp-maxit", type=int, default=3000)
    ap.add_argument("--cube", type=int, default=10)
    a = ap.parse_args()
    import bench
    from oracle import wo
    L = wo.lib()

    class Args:
        pc_blocks, pc_cube, ksp, restart = 1, a.cube, "gmres", 30
    prob = bench.Problem(a.config, 1, a.dims)
    arm = bench.CpuArm(prob, Args)
    for dt in a.dt:
        for ksp in a.ksp:
            arm.f.fluid_init(prob.y, prob.region)
            arm.f.lhs(prob.y)
            L.wo_flow_pre_timestep(arm.f.h)
            o = wo.NewtonOpts()
            o.max_iterations, o.min_iterations = a.maxit, 0
            o.rel_tol, o.abs_tol, o.update_rel_tol, o.update_abs_tol = 1e-5, 1.0, 1e-10, 1.0
            o.fd_err, o.fd_umin, o.pc_type = 1e-8, 1e-2, wo.PC_BJACOBI_ILU0
            o.ksp.type, o.ksp.restart, o.ksp.maxit = (wo.KSP_GMRES if ksp == "gmres" else wo.KSP_BCGS), 30, a.ksp_maxit
            o.ksp.rtol, o.ksp.atol, o.ksp.dtol = 1e-5, 1e-50, 1e5
            res = wo.NewtonResult()
            yy = prob.y.copy()
            t0 = time.perf_counter()
            L.wo_newton_solve_be(arm.f.h, arm.A, wo.ip(arm.color), arm.ncolor, wo.ip(arm.bor), C.byref(o), dt, wo.dp(arm.L0),
                                 wo.dp(yy), C.byref(res))
            reg = arm.f.regions()[:prob.mesh.nowned]
            print(json.dumps({"config": a.config, "dt": dt, "ksp": ksp, "reason": res.reason, "its": res.iterations,
                              "lin_its": list(res.lin_its[:res.iterations + 1]), "lin_reason": list(res.lin_reason[:res.iterations + 1]),
                              "max_res": [float("%.3g" % v) for v in res.max_residual[:res.iterations + 1]],
                              "region_changes": int((reg != prob.region).sum()), "s": round(time.perf_counter() - t0, 1)}), flush=True)


if __name__ == "__main__":
    main()
