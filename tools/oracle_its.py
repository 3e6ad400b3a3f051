#!/usr/bin/env python3
"""Krylov iteration counts of the CPU oracle on the config-2 Newton system (100^3 eos_we, dt = 1e6 s) for
different block-Jacobi sub-domain shapes, restarts and Krylov methods -- tuning aid for bench.py's defaults
(test infrastructure: runs the oracle, never the product path).

  python tools/oracle_its.py --cubes 5 10 --restarts 30 --ksp gmres
"""
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


def box_blocks(m, sx, sy, sz):
    nx, ny, nz = m.dims
    idx = m.natural[:m.nowned]
    i, j, k = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    bx, by = -(-nx // sx), -(-ny // sy)
    key = (i // sx) + bx * ((j // sy) + by * (k // sz))
    _, inv = np.unique(key, return_inverse=True)
    return inv.astype(np.int32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs=3, default=[100, 100, 100])
    ap.add_argument("--boxes", type=str, nargs="*", default=["10x10x10"], help="sub-domain shapes sx x sy x sz; 0 = global ILU(0)")
    ap.add_argument("--restarts", type=int, nargs="*", default=[30])
    ap.add_argument("--ksp", nargs="*", default=["gmres"])
    ap.add_argument("--dt", type=float, default=1.0e6)
    ap.add_argument("--maxit", type=int, default=10000)
    a = ap.parse_args()
    from oracle import wo
    from waiwera_b200 import mesh as wmesh
    wo.build()
    L = wo.lib()
    m = wmesh.structured(*a.dims, dx=10.0, seed=wmesh.SEED)
    primary, region = wmesh.hydrostatic_state(m, seed=wmesh.SEED)
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1)
    prm = wo.make_params(eos=wo.EOS_WE, thermo=wo.THERMO_IAPWS)
    f = wo.Flow(prm, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                m.cell_geom.reshape(-1), m.rock.reshape(-1))
    assert f.fluid_init(y, region) == 0
    e, L0 = f.lhs(y)
    e, lhs, rhs, F0 = f.residual(y, L0, a.dt)
    A = f.bsr()
    nb = A.contents.nb
    color = np.zeros(nb, np.int32)
    nc = L.wo_bsr_coloring(A, wo.ip(color))
    assert L.wo_fd_jacobian(f.h, wo.dp(y), wo.dp(L0), a.dt, wo.dp(F0), wo.ip(color), nc, 1e-8, 1e-2, A) == 0
    for box in a.boxes:
        if box == "0":
            bor = None
        else:
            sx, sy, sz = [int(v) for v in box.split("x")]
            bor = box_blocks(m, sx, sy, sz)
        pc = L.wo_pc_create(A, wo.PC_BJACOBI_ILU0, wo.ip(bor))
        for ksp in a.ksp:
            for rs in (a.restarts if ksp == "gmres" else [0]):
                o = wo.KspOpts()
                o.type, o.restart, o.maxit = (wo.KSP_GMRES if ksp == "gmres" else wo.KSP_BCGS), rs, a.maxit
                o.rtol, o.atol, o.dtol = 1e-5, 1e-50, 1e5
                x = np.zeros(nb * 2)
                its, rn = C.c_int(), C.c_double()
                t0 = time.perf_counter()
                reason = L.wo_ksp_solve(A, pc, C.byref(o), wo.dp(F0), wo.dp(x), C.byref(its), C.byref(rn))
                print(json.dumps({"box": box, "ksp": ksp, "restart": rs, "its": its.value, "reason": reason,
                                  "rnorm": rn.value, "s": round(time.perf_counter() - t0, 1)}), flush=True)
        L.wo_pc_destroy(pc)


if __name__ == "__main__":
    main()
