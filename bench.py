#!/usr/bin/env python3
"""Newton-step benchmark of the B200-native Waiwera hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores
  python bench.py --config 4 | --config 5 ...              # the other measured configurations of BASELINE.json

A step is ONE Newton iteration of a backward-Euler time step: residual evaluation, FD Jacobian assembly (BAIJ),
PC set-up (block-Jacobi / ILU(0)), Krylov solve to rtol 1e-5, update, phase transitions, the residual at the new
iterate and the convergence norm -- what SNES newtonls does per iteration with the callbacks timestepper.F90
registers.  Every step restarts from the same initial state, so all steps do identical work.

  --config 2 (default, the headline): 100x100x100 (1 M cells) eos_we / IAPWS, BAIJ bs = 2, 6.94 M blocks; N > 1 splits
             the same mesh over N GPUs (strong scaling, BASELINE configs[1] and [2])
  --config 4: 100x100x50 (500 k cells) eos_wce with a band of cells straddling the saturation line, bs = 3,
             3.46 M blocks, 1 GPU (configs[3])
  --config 5: MINC dual porosity, eos_wce, 250 k cells (125 k fracture + 125 k matrix) PER GPU: 2 M cells on 8 GPUs
             (weak scaling, configs[4]); rows of 2 and 8 blocks

Prints one JSON line (rank 0).
"""
import os
import sys

# torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm (oracle, OpenMP) must use all host cores, and the
# OpenMP runtime reads the variable when it is first loaded -- so set it before anything imports numpy / torch.
_NCORES = os.cpu_count() or 1
if "--impl" in sys.argv and "reference" in sys.argv or int(os.environ.get("WORLD_SIZE", "1")) == 1:
    os.environ["OMP_NUM_THREADS"] = str(_NCORES)

import argparse  # noqa: E402
import ctypes as C  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "newton_steps_per_sec"
UNIT = "Newton steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 4, 5])
    ap.add_argument("--dims", type=int, nargs=3, default=None, help="override the mesh dimensions of the configuration")
    ap.add_argument("--dt", type=float, default=None)
    ap.add_argument("--pc", default="ilu0", choices=["ilu0", "asm", "pbjacobi", "none"],
                    help="ilu0: block Jacobi + ILU(0) on the --pc-cube sub-domains; asm: PCASM (restricted, overlap 1) + ILU(0) on the same sub-domains")
    ap.add_argument("--pc-blocks", type=int, default=1, help="block-Jacobi sub-domains per GPU (contiguous row ranges)")
    ap.add_argument("--pc-cube", type=int, default=10,
                    help="block-Jacobi sub-domains = cubes of this many cells per side (0: use --pc-blocks)")
    ap.add_argument("--ksp", default="gmres", choices=["gmres", "bcgs"])
    ap.add_argument("--restart", type=int, default=30, help="GMRES restart (timestepper.F90:1776-1778)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--cpu-sample-its", type=int, default=30)
    ap.add_argument("--cpu-budget-s", type=float, default=240.0, help="--impl reference: wall-clock budget for the timed steps")
    ap.add_argument("--spmv-launches", type=int, default=50)
    ap.add_argument("--no-p2p", action="store_true", help="N > 1: keep NCCL for the per-iteration exchanges (halo, dots, norm)")
    ap.add_argument("--ksp-maxit", type=int, default=10000,
                    help="cap on Krylov iterations (profiling runs only: a capped solve is not a bench value)")
    return ap.parse_args()


# ------------------------------------------------------------------ problems

class Problem:
    """global mesh + scaled initial state of one configuration (setup-time numpy; nothing here touches oracle/)"""

    def __init__(self, cfg, world, dims=None, dt=None):
        from waiwera_b200 import flow, mesh as wmesh
        self.cfg, self.world = cfg, world
        if cfg == 2:
            self.dims = tuple(dims or (100, 100, 100))
            self.eos, self.eos_name, self.npv = flow.EOS_WE, "eos_we", 2
            self.dt = dt or 1.0e6
            self.mesh = wmesh.structured(*self.dims, dx=10.0, seed=wmesh.SEED)
            primary, self.region = wmesh.hydrostatic_state(self.mesh, seed=wmesh.SEED)
            self.scaling = "strong"
            self.parts = wmesh.default_parts(world)
            self.owner = wmesh.box_owner(self.mesh, self.parts) if world > 1 else None
            self.minc = False
        elif cfg == 4:
            self.dims = tuple(dims or (100, 100, 50))
            self.eos, self.eos_name, self.npv = flow.EOS_WCE, "eos_wce", 3
            self.dt = dt or 1.0e5
            self.mesh = wmesh.structured(*self.dims, dx=10.0, seed=wmesh.SEED)
            nz = self.dims[2]
            primary, self.region = wmesh.wce_band_state(self.mesh, seed=wmesh.SEED, band=(2 * nz // 5, 3 * nz // 5))
            self.scaling = "strong"
            self.parts = wmesh.default_parts(world)
            self.owner = wmesh.box_owner(self.mesh, self.parts) if world > 1 else None
            self.minc = False
        else:
            # weak scaling: one 50^3 box of fracture cells (+ one MINC level) per GPU, boxes side by side in x and y so
            # that every N has the same depth (the same hydrostatic column, the same kind of Newton system)
            self.parts = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (4, 2, 1)}.get(world) or wmesh.default_parts(world)
            per = dims or (50, 50, 50)
            self.dims = tuple(per[k] * self.parts[k] for k in range(3))
            self.eos, self.eos_name, self.npv = flow.EOS_WCE, "eos_wce", 3
            self.dt = dt or 1.0e5
            base = wmesh.structured(*self.dims, dx=10.0, seed=wmesh.SEED)
            n = base.ninterior
            self.mesh = wmesh.add_minc(base, volumes=(0.1, 0.9), spacing=(50.0, 50.0, 50.0), matrix_permeability_factor=0.01)
            pf, rf = wmesh.wce_state(base, seed=wmesh.SEED)
            rng = np.random.default_rng(wmesh.SEED + 9)
            pm = pf.copy()
            pm[:, 0] *= 1.0 + 1e-3 * rng.uniform(-1, 1, n)   # matrix slightly out of equilibrium with the fractures
            primary, self.region = np.concatenate([pf, pm]), np.concatenate([rf, rf])
            self.scaling = "weak"
            self.owner = wmesh.minc_owner(self.mesh, self.parts) if world > 1 else None
            self.minc = True
        self.y = np.ascontiguousarray(wmesh.scale_primaries(primary, self.region)).reshape(-1)
        self.region = np.ascontiguousarray(self.region, np.int32)

    def blocks(self, m, cube):
        from waiwera_b200 import mesh as wmesh
        return wmesh.minc_cube_blocks(m, cube) if self.minc else wmesh.cube_blocks(m, cube)

    def name(self, args):
        d = self.dims
        ncell = self.mesh.ninterior
        what = {2: "%s IAPWS %dx%dx%d structured (%d cells)" % (self.eos_name, d[0], d[1], d[2], ncell),
                4: "%s IAPWS %dx%dx%d structured (%d cells), band of cells on the saturation line" % (self.eos_name, d[0], d[1], d[2], ncell),
                5: "%s IAPWS MINC dual porosity, %dx%dx%d fracture cells + 1 matrix level (%d cells)" % (self.eos_name, d[0], d[1], d[2], ncell)}[self.cfg]
        pc = {"ilu0": "bjacobi/ILU(0)", "asm": "ASM(overlap 1)/ILU(0)"}.get(getattr(args, "pc", "ilu0"), getattr(args, "pc", "ilu0"))
        return "config %d: %s, BAIJ bs=%d, BE dt=%g s, %s%s + %s rtol 1e-5" % (
            self.cfg, what, self.npv, self.dt, args.ksp.upper(), "(%d)" % args.restart if args.ksp == "gmres" else "", pc)


def solver_bytes_per_iteration(nb, bs, rowptr, colidx, block_of_row, restart, ksp):
    """Algorithmic bytes one Krylov iteration of the persistent solver kernel moves: the matrix
    once per product (blocks + 4-byte column indices), the ILU(0) factors of the sub-domains once per preconditioner
    apply (the blocks of the matrix pattern with both ends in one sub-domain), and the vector passes.  GMRES(m), column
    k of a cycle (k = 1..m, mean (m + 1) / 2): gathered operand + stored basis vector + stored product, dots against k
    vectors + the product, multi-AXPY over k vectors + product read and written = (2 k + 6) vectors.  BiCGStab: two
    products and applies, and 22 vector reads / writes (P, S, x, R updates and five dot products)."""
    rowptr = np.asarray(rowptr, np.int64)
    colidx = np.asarray(colidx, np.int64)
    nnzb = len(colidx)
    rows = np.repeat(np.arange(nb, dtype=np.int64), np.diff(rowptr))
    own = colidx < nb
    if block_of_row is None:
        nnzf = int(own.sum())
    else:
        b = np.asarray(block_of_row, np.int64)
        nnzf = int((own & (b[rows] == b[np.minimum(colidx, nb - 1)])).sum())
    blk = bs * bs * 8 + 4
    vec = nb * bs * 8
    if ksp == "bcgs":
        return 2 * (nnzb + nnzf) * blk + 22 * vec
    return (nnzb + nnzf) * blk + (restart + 1 + 6) * vec


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

class CpuArm:
    """The reference algorithm (oracle port: OpenMP owner-computes residual, FD-coloured Jacobian, block-Jacobi
    ILU(0), GMRES / BiCGStab) on all host cores, set up once; `step` times one whole Newton step."""

    def __init__(self, prob, args):
        from oracle import wo
        wo.build()
        self.wo, self.L = wo, wo.lib()
        self.cores = self.L.wo_set_num_threads(_NCORES)
        self.prob, self.args = prob, args
        m = prob.mesh
        eos = {2: wo.EOS_WE, 4: wo.EOS_WCE, 5: wo.EOS_WCE}[prob.cfg]
        prm = wo.make_params(eos=eos, thermo=wo.THERMO_IAPWS)
        self.f = wo.Flow(prm, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                         m.cell_geom.reshape(-1), m.rock.reshape(-1))
        assert self.f.fluid_init(prob.y, prob.region) == 0
        e, self.L0 = self.f.lhs(prob.y)
        assert e == 0
        self.L.wo_flow_pre_timestep(self.f.h)   # every step starts from this state (regions restored by pre_retry_timestep)
        self.A = self.f.bsr()
        self.nb, self.bs = self.A.contents.nb, self.A.contents.bs
        self.color = np.zeros(self.nb, np.int32)
        self.ncolor = self.L.wo_bsr_coloring(self.A, wo.ip(self.color))
        nblk = max(args.pc_blocks, 1)
        self.bor = None if nblk == 1 else ((np.arange(self.nb, dtype=np.int64) * nblk) // self.nb).astype(np.int32)
        if args.pc_cube > 0:
            self.bor = prob.blocks(m, args.pc_cube)

    def residual(self, y):
        e, lhs, rhs, r = self.f.residual(y, self.L0, self.prob.dt)
        assert e == 0
        return r

    def step(self, maxit):
        """one Newton iteration from the initial state; returns the phase times"""
        wo, L, p = self.wo, self.L, self.prob
        y = p.y
        t = {}
        L.wo_flow_pre_retry_timestep(self.f.h)   # regions of the initial state (a step's transitions change them)
        t0 = time.perf_counter()
        F0 = self.residual(y)
        t["residual"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        L.wo_flow_pre_iteration(self.f.h)
        assert L.wo_fd_jacobian(self.f.h, wo.dp(y), wo.dp(self.L0), p.dt, wo.dp(F0), wo.ip(self.color), self.ncolor,
                                1e-8, 1e-2, self.A) == 0
        t["jacobian"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        pc = L.wo_pc_create(self.A, wo.PC_ASM_ILU0 if self.args.pc == "asm" else wo.PC_BJACOBI_ILU0, wo.ip(self.bor))
        t["pc_setup"] = time.perf_counter() - t0
        o = wo.KspOpts()
        o.type, o.restart, o.maxit = (wo.KSP_GMRES if self.args.ksp == "gmres" else wo.KSP_BCGS), self.args.restart, maxit
        o.rtol, o.atol, o.dtol = 1e-5, 1e-50, 1e5
        x = np.zeros(self.nb * self.bs)
        its, rn = C.c_int(), C.c_double()
        t0 = time.perf_counter()
        reason = L.wo_ksp_solve(self.A, pc, C.byref(o), wo.dp(F0), wo.dp(x), C.byref(its), C.byref(rn))
        t["ksp"] = time.perf_counter() - t0
        L.wo_pc_destroy(pc)
        t0 = time.perf_counter()
        ynew = y - x                  # shell line search, lambda = 1, with the post-check (fluid_transitions)
        cs, cy = C.c_int(), C.c_int()
        et = L.wo_flow_fluid_transitions(self.f.h, wo.dp(y), wo.dp(x), wo.dp(ynew), C.byref(cs), C.byref(cy))
        if et == 0:
            self.f.residual(ynew, self.L0, p.dt)   # the line search's function evaluation at the new iterate
        t["residual_new"] = time.perf_counter() - t0
        t["transitions_err"] = et
        t["its"], t["ksp_reason"] = its.value, reason
        return t

    def spmv_gbs(self):
        wo, L = self.wo, self.L
        xx, yy = np.ones(self.nb * self.bs), np.zeros(self.nb * self.bs)
        L.wo_bsr_spmv(self.A, wo.dp(xx), wo.dp(yy))
        t0 = time.perf_counter()
        for _ in range(5):
            L.wo_bsr_spmv(self.A, wo.dp(xx), wo.dp(yy))
        t = (time.perf_counter() - t0) / 5
        nnzb, bs = self.A.contents.nnzb, self.bs
        return (nnzb * (bs * bs * 8 + 4) + (self.nb + 1) * 4 + 2 * self.nb * bs * 8) / t / 1e9


def cpu_baseline_sample(prob, args, total_its):
    """cpu_baseline leg of the B200 arm: a bounded sample of the same workload -- the full mesh, one residual, one
    FD-coloured Jacobian, one PC set-up, `cpu_sample_its` Krylov iterations and the residual at the new iterate; the
    Newton-step time is that with the Krylov part extrapolated linearly to the iteration count of the full solve."""
    arm = CpuArm(prob, args)
    t = arm.step(args.cpu_sample_its)
    per_it = t["ksp"] / max(t["its"], 1)
    t_step = t["residual"] + t["jacobian"] + t["pc_setup"] + per_it * total_its + t["residual_new"]
    d = prob.dims
    return {"value": 1.0 / t_step, "unit": UNIT, "cores": arm.cores, "kind": "port",
            "sample": "full %dx%dx%d mesh: 1 residual (%.2fs) + 1 FD-coloured Jacobian, %d colours (%.2fs) + ILU(0) factor "
                      "(%.2fs) + %d %s iterations (%.3fs each), Krylov part extrapolated to %d iterations; oracle port of "
                      "the reference algorithm (not the PETSc binary), gcc -O3 -march=native, OpenMP on %d threads "
                      "(owner-computes cell chunks, sub-domains, SpMV rows, dots)"
                      % (d[0], d[1], d[2], t["residual"], arm.ncolor, t["jacobian"], t["pc_setup"], t["its"], args.ksp.upper(),
                         per_it, total_its, arm.cores),
            "s_per_step": t_step, "spmv_gbs": arm.spmv_gbs(), "ksp_iterations": total_its}


def run_reference(args):
    """The reference's CPU algorithm for the same Newton step on all host cores (oracle port: the reference itself is
    Fortran + PETSc and cannot be built in this image).  Every timed step is the WHOLE step -- residual, FD-coloured
    Jacobian, ILU(0) factor, the Krylov solve run to rtol 1e-5 and the residual at the new iterate, nothing
    extrapolated.  A step takes tens of seconds, so after one warm-up step as many steps as fit into --cpu-budget-s
    (at least 2, at most --steps) are timed and their MEAN is reported; `steps` says how many."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    prob = Problem(args.config, 1 if args.config != 5 else args.gpus, args.dims, args.dt)
    arm = CpuArm(prob, args)
    samples = []
    if args.warmup > 0:
        arm.step(min(args.ksp_maxit, 30))       # warm-up: page in the arrays, spin up the thread team
    t_budget = time.perf_counter()
    while len(samples) < max(2, args.steps):
        t0 = time.perf_counter()
        t = arm.step(args.ksp_maxit)
        t["step"] = t["residual"] + t["jacobian"] + t["pc_setup"] + t["ksp"] + t["residual_new"]
        samples.append(t)
        if len(samples) >= 2 and time.perf_counter() - t_budget + (time.perf_counter() - t0) > args.cpu_budget_s:
            break
    s_step = float(np.mean([s["step"] for s in samples]))
    d = prob.dims
    mean = {k: float(np.mean([s[k] for s in samples])) for k in ("residual", "jacobian", "pc_setup", "ksp", "residual_new")}
    its = int(round(np.mean([s["its"] for s in samples])))
    base = {"value": 1.0 / s_step, "unit": UNIT, "cores": arm.cores, "kind": "port",
            "sample": "%d whole Newton steps on the full %dx%dx%d mesh (mean): residual %.2fs + FD-coloured Jacobian, %d colours, "
                      "%.2fs + ILU(0) factor %.2fs + %d %s iterations run to rtol (%.2fs) + residual at the new iterate %.2fs; "
                      "oracle port of the reference algorithm (not the PETSc binary), gcc -O3 -march=native, OpenMP on %d threads"
                      % (len(samples), d[0], d[1], d[2], mean["residual"], arm.ncolor, mean["jacobian"], mean["pc_setup"], its,
                         args.ksp.upper(), mean["ksp"], mean["residual_new"], arm.cores),
            "s_per_step": s_step, "spmv_gbs": arm.spmv_gbs(), "ksp_iterations": its,
            "step_s": [round(s["step"], 3) for s in samples]}
    out = {"metric": METRIC, "value": 1.0 / s_step, "unit": UNIT, "n_gpus": args.gpus, "steps": len(samples),
           "warmup": 1 if args.warmup > 0 else 0, "ms_per_step": 1e3 * s_step, "higher_is_better": True,
           "scaling": prob.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
           "config": {"workload": prob.name(args), "timing": "host wall clock around whole Newton steps, mean of `steps` (see cpu_baseline.sample)",
                      "ksp": args.ksp, "pc": args.pc, "ksp_iterations_per_step": its,
                      "ksp_reason": int(samples[-1]["ksp_reason"]), "omp_threads": arm.cores},
           "cpu_baseline": base,
           "e2e": {"value": 1.0 / s_step, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": time.perf_counter() - t_all}
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
    prob = Problem(args.config, world, args.dims, args.dt)
    gm, gy, gregion, npv, dt = prob.mesh, prob.y, prob.region, prob.npv, prob.dt
    if world > 1:
        m = wmesh.partition(gm, prob.owner, rank, world)
        nat = m.natural[:m.nowned]
        y = np.ascontiguousarray(gy.reshape(-1, npv)[nat].reshape(-1))
        region = np.ascontiguousarray(gregion[nat])
    else:
        m, y, region = gm, gy, gregion
        nat = None
    sim = flow.FlowSimulation(flow.make_params(eos=prob.eos, thermo=flow.THERMO_IAPWS), m, device=local)
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
    sim.pre_timestep()   # every step starts from this state: pre_retry_timestep restores the regions a step's transitions changed
    if args.pc_cube > 0 and args.pc in ("ilu0", "asm"):
        sim.set_pc_blocks(prob.blocks(m, args.pc_cube))
    pc_type = {"ilu0": flow.PC_BJACOBI_ILU0, "asm": flow.PC_ASM_ILU0, "pbjacobi": flow.PC_PBJACOBI, "none": flow.PC_NONE}[args.pc]
    ksp_type = {"gmres": flow.KSP_GMRES, "bcgs": flow.KSP_BCGS}[args.ksp]
    opts = flow.newton_opts(max_iterations=1, pc_type=pc_type, pc_nblocks=args.pc_blocks,
                            ksp=flow.ksp_opts(type=ksp_type, maxit=args.ksp_maxit, restart=args.restart))
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
        sim.pre_retry_timestep()
        with torch.cuda.stream(stream):
            y_d.copy_(y0_d, non_blocking=True)
        return sim.newton_solve(y_d, L0_d, dt, opts)

    def step_host():
        sim.pre_retry_timestep()
        y_h.copy_(y0_h)
        return sim.newton_solve(y_h.numpy(), L0_h.numpy(), dt, opts)

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

    # ---- phase breakdown (timers on: device events around each phase; every rank runs it, rank 0 reports its own)
    L.wb_timers_enable(1)
    L.wb_timer_reset(sim.h)
    step_device()
    phases = {}
    for nm in ("fluid_props", "cell_inflows", "jacobian", "pc_setup", "ksp_solve", "fluid_trans"):
        t, cnt = sim.timer(nm)
        phases[nm] = {"ms": round(t, 4), "calls": cnt}
    ksp_breakdown_ctas = sim.ksp_breakdown_ctas()
    ksp_breakdown = sim.ksp_breakdown()
    L.wb_timers_enable(0)

    # ---- parity against the CPU oracle at this state (outside every timed region): residual vector, max-scaled norm
    # and its argmax (timestepper.F90:1898-1951, dm_utils.F90:644-685), and the TRUE residual after the step
    y_fin = y_d.cpu().numpy()
    e, _, _, r1 = sim.residual(y_fin, L0, dt)          # regions as the step left them
    mv1, ml1 = sim.max_scaled(r1, L0, 1.0) if e == 0 else (float("nan"), -1)
    regions_changed = int((sim.regions()[:m.nowned] != region).sum())
    sim.pre_retry_timestep()
    e, _, _, r0 = sim.residual(y, L0, dt)
    assert e == 0
    mv0, ml0 = sim.max_scaled(r0, L0, 1.0)
    parity = None
    if not args.no_parity:
        if world > 1:
            parts = [None] * world
            dist.gather_object(dict(nat=nat, r=r0), parts if rank == 0 else None, dst=0)
            if rank == 0:
                r_glob = np.zeros(gm.ninterior * npv)
                for p in parts:
                    r_glob.reshape(-1, npv)[p["nat"]] = p["r"].reshape(-1, npv)
        else:
            r_glob = r0
        if rank == 0:
            arm = CpuArm(prob, args)
            r_cpu = arm.residual(gy)
            from oracle import wo
            mv_cpu, ml_cpu = wo.max_scaled(r_cpu, arm.L0, 1.0)
            rel = float(np.linalg.norm(r_glob - r_cpu) / np.linalg.norm(r_cpu))
            # the GPU argmax is a global index in rank-contiguous numbering; map the oracle's (natural) one the same way
            if world > 1:
                order = np.concatenate([p["nat"] for p in parts])
                inv = np.empty_like(order)
                inv[order] = np.arange(len(order))
                ml_cpu_g = int(inv[ml_cpu // npv] * npv + ml_cpu % npv)
            else:
                ml_cpu_g = int(ml_cpu)
            parity = {"residual_relerr": rel, "residual_norm2": float(np.linalg.norm(r_glob)),
                      "residual_norm2_relerr": float(abs(np.linalg.norm(r_glob) - np.linalg.norm(r_cpu)) / np.linalg.norm(r_cpu)),
                      "max_scaled": mv0, "max_scaled_rel": float(abs(mv0 - mv_cpu) / abs(mv_cpu)),
                      "argmax": int(ml0), "argmax_equal": bool(int(ml0) == ml_cpu_g),
                      "oracle": "CPU port evaluated on the same state outside the timed region", "tolerance": 1e-10}
            del arm

    # ---- SpMV roofline (the north-star kernel, K5): CUDA events on the context's stream around each launch
    roofline = None
    if world == 1:
        J = sim.jacobian_mat()
        x_d = torch.randn(n, dtype=torch.float64, device="cuda")
        z_d = torch.empty_like(x_d)
        torch.cuda.synchronize()
        L.wb_timers_enable(1)
        for _ in range(5):
            J.mult(x_d, z_d)
        L.wb_timer_reset(sim.h)
        for _ in range(args.spmv_launches):
            J.mult(x_d, z_d)
        t, cnt = sim.timer("mat_mult")
        L.wb_timers_enable(0)
        nb, bs, rowptr, colidx = sim.jacobian_pattern()
        nnzb = len(colidx)
        abytes = nnzb * (bs * bs * 8 + 4) + (nb + 1) * 4 + 2 * nb * bs * 8
        pk_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        which = "fallback 6650 GB/s (B200_PROFILING.md)"
        peak = 6650.0
        if os.path.exists(pk_file):
            peak = float(json.load(open(pk_file)).get("hbm_gbs", peak))
            which = "MEASURED_PEAKS.json hbm_gbs"
        achieved = abytes / (t / cnt * 1e-3) / 1e9
        traffic = None
        tf = os.path.join(ROOT, "profiles", "spmv_traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get("config%d" % args.config, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {"kernel": "BAIJ block SpMV bs=%d (K5, wb_mat_mult on the Jacobian)" % bs, "bound": "hbm",
                    "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                    "traffic": traffic, "algorithmic_bytes": abytes, "us_per_launch": round(1e3 * t / cnt, 2),
                    "launches_timed": cnt, "peak_source": which, "frac_of_nominal_8TBs": round(achieved / 8000.0, 4)}

    # ---- the dominant kernel of the step: the persistent solver kernel (one launch = one KSPSolve, 99 % of the step).
    # achieved = algorithmic bytes per iteration x iterations / the device time of the solve inside the timed steps
    roofline_solver = None
    if world == 1 and roofline and ksp_breakdown and args.pc == "ilu0":
        try:
            bor_ = prob.blocks(m, args.pc_cube) if args.pc_cube > 0 else None
            per_it = solver_bytes_per_iteration(nb, bs, rowptr, colidx, bor_, args.restart, args.ksp)
            t_solve = phases["ksp_solve"]["ms"] * 1e-3
            ach = per_it * max(ksp_its, 1) / t_solve / 1e9
            roofline_solver = {"kernel": "k_gmres_fused (persistent %s: the whole KSPSolve, the dominant kernel of the step)" % args.ksp.upper(),
                               "bound": "hbm", "achieved": round(ach, 1), "peak": roofline["peak"], "unit": "GB/s",
                               "frac": round(ach / roofline["peak"], 4), "algorithmic_bytes_per_iteration": int(per_it),
                               "iterations_per_launch": int(ksp_its), "ms_per_launch": round(phases["ksp_solve"]["ms"], 3),
                               "traffic_per_iteration": 1.111e9 if (args.config == 2 and args.ksp == "gmres") else None,
                               "traffic_source": "profiles/r2f_fused_ncu_metrics.csv (ncu --set full: 95.85 GB read + 4.16 GB written over 90 iterations)"}
        except Exception as e:          # never let the extra line break the bench
            roofline_solver = {"error": str(e)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ncell_global = gm.ninterior
    out = {"metric": METRIC, "value": args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": prob.scaling,
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": prob.name(args), "cells": int(ncell_global), "cells_per_gpu": int(m.nowned),
                      "parallelism": "domain decomposition %s, halo of x inside SpMV%s" % (prob.parts, "" if world == 1 else (", per-iteration exchanges over NVLink P2P (CUDA IPC)" if p2p else ", NCCL exchanges")),
                      "pc": args.pc, "pc_subdomains": ("cubes of %d^3 cells" % args.pc_cube) if args.pc_cube > 0 else ("%d contiguous ranges per GPU" % args.pc_blocks),
                      "ksp": args.ksp, "restart": args.restart,
                      "l2": "working set (Jacobian + factors + Krylov basis) exceeds the 126 MB L2 at 1 GPU; no explicit flush",
                      "ksp_iterations_per_step": ksp_its, "ksp_reason": int(res.lin_reason[0]), "ksp_rnorm": res.lin_rnorm[0],
                      "us_per_ksp_iteration": round(1e3 * phases["ksp_solve"]["ms"] / max(ksp_its, 1), 2),
                      "newton_reason": int(res.reason),
                      "max_scaled_residual": [res.max_residual[0], res.max_residual[1]],
                      "post_step_max_scaled_residual": mv1, "cells_changing_region_in_the_step": regions_changed,
                      "cell_updates_per_s": ncell_global * args.steps / (ms * 1e-3)},
           "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 2 * n * 8,
                   "d2h_bytes_per_step": n * 8 + C.sizeof(flow.NewtonResult), "ms_per_step": ms_e2e / args.steps},
           "gpu_launches": int(launches), "clocks": clocks, "phases_ms": phases, "parity": parity}
    if ksp_breakdown:
        out["ksp_breakdown_us_per_iteration"] = ksp_breakdown
        out["ksp_breakdown_min_mean_max_over_ctas"] = ksp_breakdown_ctas
    if roofline:
        out["roofline"] = roofline
    if roofline_solver:
        out["roofline_solver"] = roofline_solver
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_sample(prob, args, ksp_its)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
