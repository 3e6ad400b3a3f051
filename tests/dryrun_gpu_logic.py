"""Test infrastructure, not a test: runs the Python side of `gpu`-marked tests on a machine WITHOUT a GPU by putting the
oracle behind the method names of waiwera_b200.flow (FlowSimulation, newton_opts, ksp_opts).  It says nothing about the
CUDA path -- it only catches mistakes in a test's own set-up, arguments and tolerances before GPU minutes are spent on it.

    python tests/dryrun_gpu_logic.py test_zz_new_gpu_cases::test_cuda_path_runs_minc_production3d \
                                     "test_zz_new_gpu_cases::test_cuda_path_runs_separated_limiter_decks[deliv_delw]"

Only tests that go through ingest.load(..., mod=flow) / flow.FlowSimulation and the source-control setters can run."""
import importlib
import inspect
import os
import pathlib
import re
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import wo            # noqa: E402
from util import OracleSim       # noqa: E402


class ShimSimulation:
    def __init__(self, params, m, device=0):
        self.f = wo.Flow(params, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                         m.cell_geom.reshape(-1), m.rock.reshape(-1))
        self.sim, self.ncell, self._n = None, m.ncell, 0

    def set_boundaries(self, ghosts, interior, primary, region):
        for k in range(len(region)):
            assert self.f.set_boundary(int(ghosts[k]), int(interior[k]), np.asarray(primary[k], float), int(region[k])) == 0
        return 0

    def set_sources(self, cells, *a):
        self._n = len(cells)
        self.f.set_sources(cells, *a)
        return 0

    def set_source_components(self, *a):
        self.f.set_source_components(*a)
        return 0

    def set_source_controls(self, *a):
        self.f.set_source_controls(*a)
        return 0

    def set_source_recharge(self, *a):
        self.f.set_source_recharge(*a)
        return 0

    def set_source_separators(self, *a):
        return self.f.set_source_separators(*a)

    def set_source_pressure_table(self, *a):
        return self.f.set_source_pressure_table(*a)

    def set_rock(self, rock):
        return self.f.set_rock(rock)

    def source_rates(self):
        return self.f.source_rates(self._n)

    def fluid_init(self, y, region):
        return self.f.fluid_init(y, region)

    def lhs(self, y):
        return self.f.lhs(y)

    def residual(self, y, L0, dt):
        return self.f.residual(y, L0, dt)

    def pre_timestep(self):
        wo.lib().wo_flow_pre_timestep(self.f.h)

    def pre_retry_timestep(self):
        wo.lib().wo_flow_pre_retry_timestep(self.f.h)

    def newton_solve(self, y, L0, dt, opts=None):
        if self.sim is None:
            self.sim = OracleSim(wo, self.f, opts)
        self.sim.opts = opts
        return self.sim.newton_solve(y, L0, dt)

    def fluid(self):
        return self.f.fluid()

    def regions(self):
        return self.f.regions()

    def destroy(self):
        pass


def shim_module():
    fake = types.ModuleType("waiwera_b200.flow")
    for k in dir(wo):
        setattr(fake, k, getattr(wo, k))
    fake.FlowSimulation = ShimSimulation

    def newton_opts(max_iterations=8, rel_tol=1e-5, abs_tol=1.0, pc_type=None, ksp=None, **_):
        o = wo.NewtonOpts()
        o.max_iterations, o.min_iterations = max_iterations, 0
        o.rel_tol, o.abs_tol, o.update_rel_tol, o.update_abs_tol = rel_tol, abs_tol, 1e-10, 1.0
        o.fd_err, o.fd_umin, o.pc_type = 1e-8, 1e-2, wo.PC_BJACOBI_ILU0
        o.ksp.type, o.ksp.restart, o.ksp.maxit = wo.KSP_BCGS, 30, 10000
        o.ksp.rtol, o.ksp.atol, o.ksp.dtol = 1e-5, 1e-50, 1e5
        return o
    fake.newton_opts = newton_opts
    fake.ksp_opts = lambda **kw: None
    return fake


def main(argv):
    import waiwera_b200
    fake = shim_module()
    sys.modules["waiwera_b200.flow"] = fake
    waiwera_b200.flow = fake
    for spec in argv:
        mod, name = spec.split("::")
        m = re.match(r"(\w+)\[(.*)\]$", name)
        args = [wo] + ([m.group(2)] if m else [])
        fn = getattr(importlib.import_module(mod), m.group(1) if m else name)
        if "tmp_path" in inspect.signature(fn).parameters:
            args.append(pathlib.Path(tempfile.mkdtemp()))
        fn(*args)
        print("ok", spec)


if __name__ == "__main__":
    main(sys.argv[1:])
