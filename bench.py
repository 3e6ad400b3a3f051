#!/usr/bin/env python3
"""Newton-step benchmark of the B200-native Waiwera hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

A step is ONE Newton iteration of a backward-Euler time step on the 100x100x100 (1 M cell) eos_we /
IAPWS mesh of BASELINE.json configs[1]: residual evaluation, FD Jacobian assembly (BAIJ bs=2,
6.94 M blocks), PC set-up (block-Jacobi / ILU(0)), GMRES(30) solve to rtol 1e-5, update, phase
transitions and the convergence norm -- exactly what SNES newtonls does per iteration with the
callbacks timestepper.F90 registers.  Every step restarts from the same initial state, so all steps do
identical work.  Prints one JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT = 1.0e6
DIMS = (100, 100, 100)
METRIC = "newton_steps_per_sec"
UNIT = "Newton steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dims", type=int, nargs=3, default=list(DIMS))
    ap.add_argument("--pc", default="ilu0", choices=["ilu0", "pbjacobi", "none"])
    ap.add_argument("--pc-blocks", type=int, default=1, help="block-Jacobi sub-domains per GPU (contiguous row ranges)")
    ap.add_argument("--pc-cube", type=int, default=10,
                    help="block-Jacobi sub-domains = cubes of this many cells per side (0: use --pc-blocks)")
    ap.add_argument("--ksp", default="gmres", choices=["gmres", "bcgs"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-its", type=int, default=30)
    ap.add_argument("--spmv-launches", type=int, default=50)
    ap.add_argument("--no-p2p", action="store_true", help="N > 1: keep NCCL for the per-iteration exchanges (halo, dots, norm)")
    ap.add_argument("--ksp-maxit", type=int, default=10000,
                    help="cap on Krylov iterations (profiling runs only: a capped solve is not a bench value)")
    return ap.parse_args()


def workload_name(dims):
    return "eos_we IAPWS %dx%dx%d structured (%d cells), BAIJ bs=2, BE dt=1e6 s, Krylov+bjacobi/ILU(0) rtol 1e-5" % (
        dims[0], dims[1], dims[2], dims[0] * dims[1] * dims[2])


def build_problem(dims):
    """synthetic config 2: hydrostatic single-phase liquid + noise, heterogeneous rock, closed box"""
    from waiwera_b200 import mesh as wmesh
    m = wmesh.structured(*dims, dx=10.0, seed=wmesh.SEED)
    primary, region = wmesh.hydrostatic_state(m, seed=wmesh.SEED)
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1)
    return m, y, region


# ------------------------------------------------------------------ clocks sampler

class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) > 8:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU arm (oracle = port of the reference algorithm)

def cpu_newton_sample(dims, pc_blocks, sample_its, total_its_hint=None, pc_cube=0, ksp="gmres"):
    """Times the reference algorithm (oracle port: FD-coloured Jacobian, ILU(0), GMRES(30)) on the host cores.
    cpu_baseline leg (sample_its small): a bounded sample of the same workload -- the full mesh, one residual, one
    FD Jacobian, one PC set-up and `sample_its` Krylov iterations; the Newton-step time is that with the Krylov
    part extrapolated linearly to the iteration count `total_its_hint` of the full solve.
    --impl reference leg (sample_its = the solver's own limit, no hint): the whole Newton step, Krylov solve run to
    its rtol on the CPU, nothing extrapolated."""
    from oracle import wo
    wo.build()
    L = wo.lib()
    ncores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(ncores))
    m, y, region = build_problem(dims)
    prm = wo.make_params(eos=wo.EOS_WE, thermo=wo.THERMO_IAPWS)
    f = wo.Flow(prm, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    assert f.fluid_init(y, region) == 0
    t0 = time.perf_counter()
    e, L0 = f.lhs(y)
    e, lhs, rhs, F0 = f.residual(y, L0, DT)
    t_res = time.perf_counter() - t0
    A = f.bsr()
    nb = A.contents.nb
    color = np.zeros(nb, np.int32)
    nc = L.wo_bsr_coloring(A, wo.ip(color))
    t0 = time.perf_counter()
    assert L.wo_fd_jacobian(f.h, wo.dp(y), wo.dp(L0), DT, wo.dp(F0), wo.ip(color), nc, 1e-8, 1e-2, A) == 0
    t_jac = time.perf_counter() - t0
    nblk = max(pc_blocks, 1)
    bor = None if nblk == 1 else ((np.arange(nb, dtype=np.int64) * nblk) // nb).astype(np.int32)
    if pc_cube > 0:
        from waiwera_b200 import mesh as wmesh
        bor = wmesh.cube_blocks(m, pc_cube)
    t0 = time.perf_counter()
    pc = L.wo_pc_create(A, wo.PC_BJACOBI_ILU0, wo.ip(bor))
    t_pc = time.perf_counter() - t0
    o = wo.KspOpts()
    o.type, o.restart, o.maxit = (wo.KSP_GMRES if ksp == "gmres" else wo.KSP_BCGS), 30, sample_its
    o.rtol, o.atol, o.dtol = 1e-5, 1e-50, 1e5
    x = np.zeros(nb * 2)
    its, rn = C.c_int(), C.c_double()
    t0 = time.perf_counter()
    L.wo_ksp_solve(A, pc, C.byref(o), wo.dp(F0), wo.dp(x), C.byref(its), C.byref(rn))
    t_ksp = time.perf_counter() - t0
    # SpMV rate of the CPU path
    xx, yy = np.ones(nb * 2), np.zeros(nb * 2)
    t0 = time.perf_counter()
    for _ in range(5):
        L.wo_bsr_spmv(A, wo.dp(xx), wo.dp(yy))
    t_spmv = (time.perf_counter() - t0) / 5
    nnzb = A.contents.nnzb
    spmv_bytes = nnzb * (4 * 8 + 4) + (nb + 1) * 4 + 2 * nb * 2 * 8
    L.wo_pc_destroy(pc)
    L.wo_bsr_destroy(A)
    per_it = t_ksp / max(its.value, 1)
    total_its = total_its_hint if total_its_hint else its.value
    t_step = t_res + t_jac + t_pc + per_it * total_its
    how = ("Krylov part extrapolated to %d iterations" % total_its) if total_its != its.value else \
        "whole Krylov solve run on the CPU (nothing extrapolated)"
    return {"value": 1.0 / t_step, "unit": UNIT, "cores": ncores, "kind": "port",
            "sample": "full %dx%dx%d mesh: 1 residual (%.2fs) + 1 FD-coloured Jacobian, %d colours (%.2fs) + ILU(0) factor "
                      "(%.2fs) + %d %s iterations (%.3fs each); %s; "
                      "oracle port of the reference algorithm (not the PETSc binary), OpenMP over sub-domains / SpMV / dots"
                      % (dims[0], dims[1], dims[2], t_res, nc, t_jac, t_pc, its.value, ksp.upper(), per_it, how),
            "s_per_step": t_step, "spmv_gbs": spmv_bytes / t_spmv / 1e9, "ksp_iterations": total_its}


def run_reference(args):
    """The reference's CPU algorithm for the same Newton step on all host cores (oracle port: the reference itself is
    Fortran + PETSc and cannot be built in this image).  Every step is the WHOLE step -- residual, FD-coloured
    Jacobian, ILU(0) factor and the Krylov solve run to rtol 1e-5 -- about 45 s at 1 M cells on 16 cores, so at most
    two steps are timed whatever --steps says; the best one is reported."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dims = tuple(args.dims)
    t0 = time.perf_counter()
    samples = []
    nrep = max(1, min(args.steps, 2))
    for _ in range(nrep):
        samples.append(cpu_newton_sample(dims, args.pc_blocks, args.ksp_maxit, None, args.pc_cube, args.ksp))
    best = max(samples, key=lambda s: s["value"])
    out = {"metric": METRIC, "value": best["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": nrep,
           "warmup": 0, "ms_per_step": 1e3 * best["s_per_step"], "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
           "config": {"workload": workload_name(dims), "timing": "host wall clock around whole Newton steps (see cpu_baseline.sample)",
                      "ksp": args.ksp, "pc": args.pc, "ksp_iterations_per_step": best["ksp_iterations"]},
           "cpu_baseline": best,
           "e2e": {"value": best["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": time.perf_counter() - t0}
    print(json.dumps(out))


# ------------------------------------------------------------------ B200 arm

def run_b200(args):
    import torch
    import torch.distributed as dist
    from waiwera_b200 import flow, mesh as wmesh, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the waiwera_b200 path is CUDA only (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    dims = tuple(args.dims)
    gm, gy, gregion = build_problem(dims)
    if world > 1:
        owner = wmesh.box_owner(gm, wmesh.default_parts(world))
        m = wmesh.partition(gm, owner, rank, world)
        nat = m.natural[:m.nowned]
        y = np.ascontiguousarray(gy.reshape(-1, 2)[nat].reshape(-1))
        region = np.ascontiguousarray(gregion[nat])
    else:
        m, y, region = gm, gy, gregion
    sim = flow.FlowSimulation(flow.make_params(eos=flow.EOS_WE, thermo=flow.THERMO_IAPWS), m, device=local)
    p2p = False
    if world > 1:
        uid = torch.from_numpy(flow.FlowSimulation.unique_id()).cuda() if rank == 0 else torch.zeros(128, dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        sim.comm_init(rank, world, uid.cpu().numpy())
        if not args.no_p2p:
            try:
                p2p = sim.p2p_setup(dist, torch.device("cuda", local))
            except Exception as e:  # CUDA IPC not available on this box: NCCL carries the exchanges
                if rank == 0:
                    sys.stderr.write("bench.py: NVLink P2P path unavailable (%s); using NCCL\n" % e)
                p2p = False
    assert sim.fluid_init(y, region) == 0
    err, L0 = sim.lhs(y)
    assert err == 0
    if args.pc_cube > 0 and args.pc == "ilu0":
        sim.set_pc_blocks(wmesh.cube_blocks(m, args.pc_cube))
    pc_type = {"ilu0": flow.PC_BJACOBI_ILU0, "pbjacobi": flow.PC_PBJACOBI, "none": flow.PC_NONE}[args.pc]
    ksp_type = {"gmres": flow.KSP_GMRES, "bcgs": flow.KSP_BCGS}[args.ksp]
    opts = flow.newton_opts(max_iterations=1, pc_type=pc_type, pc_nblocks=args.pc_blocks, ksp=flow.ksp_opts(type=ksp_type, maxit=args.ksp_maxit))
    n = sim.n
    stream = torch.cuda.ExternalStream(sim.stream())
    y0_d = torch.from_numpy(y).cuda()
    L0_d = torch.from_numpy(L0).cuda()
    y_d = torch.empty_like(y0_d)
    y0_h = torch.from_numpy(y).pin_memory()
    L0_h = torch.from_numpy(L0).pin_memory()
    y_h = torch.empty(n, dtype=torch.float64).pin_memory()
    torch.cuda.synchronize()

    def step_device():
        with torch.cuda.stream(stream):
            y_d.copy_(y0_d, non_blocking=True)
        return sim.newton_solve(y_d, L0_d, DT, opts)

    def step_host():
        y_h.copy_(y0_h)
        return sim.newton_solve(y_h.numpy(), L0_h.numpy(), DT, opts)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        res = None
        for _ in range(k):
            res = fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res

    L.wb_timers_enable(0)
    for _ in range(args.warmup):
        res = step_device()
    launches0 = sim.launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, res = timed(step_device, args.steps)
    launches = sim.launches() - launches0
    for _ in range(1):
        step_host()
    ms_e2e, res_h = timed(step_host, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ksp_its = int(res.linear_iterations)

    # ---- phase breakdown + SpMV roofline (timers on: device events around each phase / kernel)
    L.wb_timers_enable(1)
    L.wb_timer_reset(sim.h)
    step_device()
    phases = {}
    for nm in ("fluid_props", "cell_inflows", "jacobian", "pc_setup", "ksp_solve", "fluid_trans"):
        t, cnt = sim.timer(nm)
        phases[nm] = {"ms": round(t, 4), "calls": cnt}
    roofline = None
    if world == 1:
        J = sim.jacobian_mat()
        x_d = torch.randn(n, dtype=torch.float64, device="cuda")
        z_d = torch.empty_like(x_d)
        torch.cuda.synchronize()
        for _ in range(5):
            J.mult(x_d, z_d)
        L.wb_timer_reset(sim.h)
        for _ in range(args.spmv_launches):
            J.mult(x_d, z_d)
        t, cnt = sim.timer("mat_mult")
        nb, bs, rowptr, colidx = sim.jacobian_pattern()
        nnzb = len(colidx)
        abytes = nnzb * (bs * bs * 8 + 4) + (nb + 1) * 4 + 2 * nb * bs * 8
        peaks = {}
        pk_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        which = "fallback 6650 GB/s (B200_PROFILING.md)"
        peak = 6650.0
        if os.path.exists(pk_file):
            peaks = json.load(open(pk_file))
            peak = float(peaks.get("hbm_gbs", peak))
            which = "MEASURED_PEAKS.json hbm_gbs"
        achieved = abytes / (t / cnt * 1e-3) / 1e9
        traffic = None
        tf = os.path.join(ROOT, "profiles", "spmv_traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {"kernel": "k_bsr_spmv<2> (GMRES block SpMV, K5)", "bound": "hbm", "achieved": round(achieved, 1),
                    "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                    "algorithmic_bytes": abytes, "us_per_launch": round(1e3 * t / cnt, 2), "launches_timed": cnt,
                    "peak_source": which, "frac_of_nominal_8TBs": round(achieved / 8000.0, 4)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = {"metric": METRIC, "value": args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(dims), "parallelism": "domain decomposition %s, halo of x inside SpMV%s" % (wmesh.default_parts(world), "" if world == 1 else (", per-iteration exchanges over NVLink P2P (CUDA IPC)" if p2p else ", NCCL exchanges")),
                      "pc": args.pc, "pc_subdomains": ("cubes of %d^3 cells" % args.pc_cube) if args.pc_cube > 0 else ("%d contiguous ranges per GPU" % args.pc_blocks),
                      "ksp": args.ksp,
                      "l2": "working set (Jacobian 222 MB + Krylov basis 496 MB) exceeds the 126 MB L2; no explicit flush",
                      "ksp_iterations_per_step": ksp_its, "newton_reason": int(res.reason),
                      "max_scaled_residual": [res.max_residual[0], res.max_residual[1]]},
           "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 2 * n * 8,
                   "d2h_bytes_per_step": n * 8 + C.sizeof(flow.NewtonResult), "ms_per_step": ms_e2e / args.steps},
           "gpu_launches": int(launches), "clocks": clocks, "phases_ms": phases}
    if roofline:
        out["roofline"] = roofline
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_newton_sample(dims, args.pc_blocks, args.cpu_sample_its, ksp_its, args.pc_cube, args.ksp)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
