import sys, os, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import wo
from waiwera_b200 import flow
from test_gpu_linalg import random_bsr
from util import make_problem, gpu_flow, relerr
for dims, bs in (((12, 12, 10), 2), ((24, 24, 16), 2), ((24, 24, 16), 3), ((33, 20, 7), 2), ((40, 40, 40), 1)):
    m, A, rowptr, colidx, val = random_bsr(wo, dims, bs, 5)
    _, y0, region, prm = make_problem(wo, dims=dims)
    sim = gpu_flow(wo, flow, m, prm, y0, region)
    nb = m.nowned
    M = flow.Mat.create(sim, nb, nb, bs, rowptr, colidx, val)
    x = np.random.default_rng(1).uniform(-1, 1, nb * bs)
    ref, got = np.zeros(nb * bs), np.zeros(nb * bs)
    wo.lib().wo_bsr_spmv(A, wo.dp(x), wo.dp(ref))
    M.mult(x, got)
    bad = np.flatnonzero(np.abs(got - ref) > 1e-12 * np.abs(ref).max())
    print(dims, bs, "relerr", relerr(got, ref), "bad rows", len(bad), bad[:8] // bs)
    val2 = val * 2.0
    M.set_values(val2)
    M.mult(x, got)
    print("   after set_values", relerr(got, 2 * ref))
    sim.destroy()
