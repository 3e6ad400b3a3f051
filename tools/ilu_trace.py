#!/usr/bin/env python3
"""Per-CTA timeline of the sub-domain ILU(0) solve on the bench problem (tuning aid, run under gpurun)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from waiwera_b200 import flow, mesh as wmesh, _lib

L = _lib.lib()
cube = int(sys.argv[1]) if len(sys.argv) > 1 else 10
m, y, region = bench.build_problem((100, 100, 100))
sim = flow.FlowSimulation(flow.make_params(), m)
assert sim.fluid_init(y, region) == 0
err, L0 = sim.lhs(y)
assert sim.jacobian(y, L0, bench.DT) == 0
J = sim.jacobian_mat()
pc = flow.PC(J, flow.PC_BJACOBI_ILU0, 1, wmesh.cube_blocks(m, cube))
x = torch.randn(sim.n, dtype=torch.float64, device="cuda")
z = torch.empty_like(x)
nblk = int(wmesh.cube_blocks(m, cube).max()) + 1
L.wb_debug_pc_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
for ns in (-1, 2, 3, 4):
    buf = np.zeros(4 * nblk + 128, np.int64)
    out = buf[:4 * nblk].reshape(nblk, 4)
    for rep in range(3):
        rc = L.wb_debug_pc_trace(pc.h, x.data_ptr(), z.data_ptr(), buf.ctypes.data, ns)
    assert rc == 0
    t0 = out[:, 0].min()
    st, en = (out[:, 0] - t0) / 1e3, (out[:, 1] - t0) / 1e3
    dur = en - st
    print("nstage %d: kernel span %.1f us; CTA duration min/med/max %.1f/%.1f/%.1f us; start times: %d CTAs at <5us, last start %.1f us; SMs used %d"
          % (ns, en.max(), dur.min(), np.median(dur), dur.max(), (st < 5).sum(), st.max(), len(np.unique(out[:, 2]))))
    # concurrency profile
    for t in (10, 30, 50, 70, 90):
        print("   t=%d us: %d CTAs running" % (t, ((st <= t) & (en > t)).sum()))
