import sys, os, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import wo
from waiwera_b200 import flow
from test_gpu_linalg import random_bsr
from test_gpu_fused import box_blocks, CASES
from util import make_problem, gpu_flow, relerr, SEED
L = flow._lib.lib()
for rep in range(3):
  for ci, case in enumerate(CASES[:3]):
    dims, bs, box = case["dims"], case["bs"], case["box"]
    m, A, rowptr, colidx, val = random_bsr(wo, dims, bs, SEED + 21, diag_boost=3.0)
    _, y0, region, prm = make_problem(wo, dims=dims)
    sim = gpu_flow(wo, flow, m, prm, y0, region)
    nb = m.nowned
    M = flow.Mat.create(sim, nb, nb, bs, rowptr, colidx, val)
    bor = box_blocks(dims, box)
    b = np.random.default_rng(SEED + 5).uniform(-1, 1, nb * bs)
    opts = flow.ksp_opts(type=0, restart=case["restart"], maxit=case["maxit"], rtol=case["rtol"])
    for fused in (0, 2):
        L.wb_ksp_set_fused(fused)
        pc = flow.PC(M, 2, 1, bor)
        x = np.full(nb * bs, 7.0)
        reason, its, rn = flow.ksp_solve(M, pc, b, x, opts)
        ax = np.zeros(nb * bs); M.mult(x, ax)
        print(rep, ci, "fused" if fused else "unfused", reason, its, "%.3e" % rn, "true relres %.2e" % relerr(ax, b), flush=True)
        pc.destroy()
    M.destroy(); wo.lib().wo_bsr_destroy(A); sim.destroy()
