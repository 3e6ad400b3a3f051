import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from oracle import wo
from waiwera_b200 import flow
from util import *
m, y, region, prm = make_problem(wo, thermo=0, two_phase_layers=2)
ref = oracle_flow(wo, m, prm, y, region)
sim = gpu_flow(wo, flow, m, prm, y, region)
_, L0 = ref.lhs(y)
rng = np.random.default_rng(SEED + 11)
y2 = y * (1 + 1e-4 * rng.uniform(-1, 1, len(y)))
dt = 1e6
nb, bs, rowptr, colidx = sim.jacobian_pattern()
assert sim.jacobian(y2, L0, dt) == 0
Jl = sim.jacobian_values()
assert sim.jacobian(y2, L0, dt, colored=True) == 0
Jc = sim.jacobian_values()
print("ncolors", sim.ncolors)
rows = np.repeat(np.arange(nb), np.diff(rowptr))
d = np.abs(Jl - Jc)
idx = np.argsort(-d.max(axis=1))[:12]
for e in idx:
    print("row", rows[e], "col", colidx[e], "reg", region[rows[e]], region[colidx[e]], "Jl", Jl[e], "Jc", Jc[e], "diff", d[e])
print("nonzero diffs", (d.max(axis=1) > 0).sum(), "of", len(d))
which = d.max(axis=1) > 0
print("diag?", (rows[which] == colidx[which]).sum(), "offdiag", (rows[which] != colidx[which]).sum())
print("regions of rows with diffs", np.bincount(region[rows[which]]), "cols", np.bincount(region[colidx[which]]))
print("which entries (col-major idx) differ", (d > 0).sum(axis=0))
