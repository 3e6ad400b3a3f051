"""Host-side mirror of the reference's interface for the Newton-step path.

`FlowSimulation` exposes the `ode_type` hooks `flow_simulation_type` binds
(src/flow_simulation.F90:103-128: lhs, rhs, pre_eval, pre_iteration, pre_timestep,
pre_retry_timestep, post_linesearch, setup_jacobian) plus the Mat / PC / KSP / SNES calls
`timestepper.F90` makes, each forwarding to the C ABI in include/waiwera_b200.h.  Arrays may be
numpy arrays (host) or torch CUDA tensors (device, zero copy).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Params, Relperm, Cappress, KspOpts, NewtonOpts, NewtonResult, check

THERMO_IAPWS, THERMO_IFC67 = 0, 1
EOS_WE, EOS_W, EOS_WCE, EOS_WAE = 0, 1, 2, 3
RP_FULLY_MOBILE, RP_LINEAR, RP_PICKENS, RP_COREY, RP_GRANT, RP_VAN_GENUCHTEN, RP_TABLE = range(7)
CP_ZERO, CP_LINEAR, CP_VAN_GENUCHTEN, CP_TABLE = range(4)
PC_NONE, PC_PBJACOBI, PC_BJACOBI_ILU0, PC_ASM_ILU0 = 0, 1, 2, 3
PRESSURE_TABLE_MAX = 8          # WB_PRESSURE_TABLE_MAX
KSP_GMRES, KSP_BCGS = 0, 1
METHOD_BEULER, METHOD_BDF2, METHOD_DIRECTSS = 0, 1, 2


def ptr(a, dtype=None):
    """void* of a numpy array (host) or torch tensor (device); None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if dtype is not None:
            assert a.dtype == dtype, (a.dtype, dtype)
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    # torch tensor
    assert a.is_contiguous()
    return a.data_ptr()


def make_relperm(kind="linear", **kw):
    """wb_relperm from the "relative_permeability" value of the input (src/relative_permeability.F90:197-558;
    parameter layout: include/waiwera_b200.h).  kind: "fully_mobile", "linear", "pickens", "corey", "grant",
    "van_genuchten", "table"; keyword names are the input's ("liquid", "vapour", "power", "slr", "ssr", "lambda",
    "sls", "sum_unity")."""
    kind = kind.lower().replace(" ", "_")
    r = Relperm()
    if kind in ("fully_mobile", "fully mobile"):
        r.type = RP_FULLY_MOBILE
    elif kind == "linear":
        r.type = RP_LINEAR
        liq, vap = kw.get("liquid", (0.0, 1.0)), kw.get("vapour", (0.0, 1.0))
        r.p[0], r.p[1], r.p[2], r.p[3] = liq[0], liq[1], vap[0], vap[1]
    elif kind == "pickens":
        r.type = RP_PICKENS
        r.p[0] = kw.get("power", 1.0)
    elif kind in ("corey", "grant"):
        r.type = RP_COREY if kind == "corey" else RP_GRANT
        r.p[0], r.p[1] = kw.get("slr", 0.3), kw.get("ssr", 0.05 if kind == "corey" else 0.6)
    elif kind == "van_genuchten":
        r.type = RP_VAN_GENUCHTEN
        r.p[0] = kw.get("lambda", kw.get("lambda_", 0.45))
        r.p[1], r.p[2] = kw.get("slr", 1e-3), kw.get("sls", 1.0)
        # "sum_unity" (default true): vapour curve 1 - krl; otherwise the ssr curve, ssr default 0.6
        # (src/relative_permeability.F90:436-455)
        r.p[3] = 1.0 if kw.get("sum_unity", True) else 0.0
        r.p[4] = kw.get("ssr", 0.6)
    elif kind == "table":
        r.type = RP_TABLE
        liq, vap = kw["liquid"], kw["vapour"]
        assert len(liq) <= _lib.WB_MAX_TABLE and len(vap) <= _lib.WB_MAX_TABLE
        r.nl, r.nv = len(liq), len(vap)
        for k, (x, y) in enumerate(liq):
            r.lx[k], r.ly[k] = x, y
        for k, (x, y) in enumerate(vap):
            r.vx[k], r.vy[k] = x, y
    else:
        raise ValueError("relative permeability type %r" % kind)
    return r


def make_cappress(kind="zero", **kw):
    """wb_cappress from the "capillary_pressure" value of the input (src/capillary_pressure.F90:159-358)"""
    kind = kind.lower().replace(" ", "_")
    c = Cappress()
    if kind == "zero":
        c.type = CP_ZERO
    elif kind == "linear":
        c.type = CP_LINEAR
        lim = kw.get("saturation_limits", (0.0, 1.0))
        c.p[0], c.p[1], c.p[2] = lim[0], lim[1], kw.get("pressure", 0.125e5)
    elif kind == "van_genuchten":
        c.type = CP_VAN_GENUCHTEN
        c.p[0], c.p[1] = kw.get("P0", 0.125e5), kw.get("lambda", kw.get("lambda_", 0.45))
        c.p[2], c.p[3] = kw.get("slr", 1e-3), kw.get("sls", 1.0)
        c.p[4] = kw.get("Pmax", 0.0)
        c.p[5] = 1.0 if "Pmax" in kw else 0.0
    elif kind == "table":
        c.type = CP_TABLE
        pts = kw["pressure"]
        assert len(pts) <= _lib.WB_MAX_TABLE
        c.n = len(pts)
        for k, (x, y) in enumerate(pts):
            c.x[k], c.y[k] = x, y
    else:
        raise ValueError("capillary pressure type %r" % kind)
    return c


def make_params(eos=EOS_WE, thermo=THERMO_IAPWS, relperm=None, cappress=None, gravity=(0.0, 0.0, -9.8),
                extrapolate=0, eos_w_temperature=20.0, partial_pressure_scale=0.0):
    """wb_params from any object with the same fields (e.g. the oracle's ctypes structs in the tests)."""
    p = Params()
    p.eos, p.thermo, p.extrapolate = eos, thermo, extrapolate
    p.pressure_scale, p.temperature_scale = 1.0e6, 1.0e2
    p.partial_pressure_scale = partial_pressure_scale  # <= 0: adaptive (reference default)
    p.eos_w_temperature = eos_w_temperature
    if relperm is None:
        p.relperm.type = RP_LINEAR
        p.relperm.p[0], p.relperm.p[1], p.relperm.p[2], p.relperm.p[3] = 0.0, 1.0, 0.0, 1.0
    else:
        C.memmove(C.byref(p.relperm), C.byref(relperm), C.sizeof(Relperm))
    if cappress is None:
        p.cappress.type = CP_ZERO
    else:
        C.memmove(C.byref(p.cappress), C.byref(cappress), C.sizeof(Cappress))
    for k in range(3):
        p.gravity[k] = gravity[k]
    return p


def ksp_opts(type=KSP_GMRES, restart=30, maxit=10000, rtol=1e-5, atol=1e-50, dtol=1e5):
    """PETSc KSP defaults."""
    o = KspOpts()
    o.type, o.restart, o.maxit, o.rtol, o.atol, o.dtol = type, restart, maxit, rtol, atol, dtol
    return o


def newton_opts(max_iterations=8, min_iterations=0, rel_tol=1e-5, abs_tol=1.0, update_rel_tol=1e-10,
                update_abs_tol=1.0, fd_err=1e-8, fd_umin=1e-2, pc_type=PC_BJACOBI_ILU0, pc_nblocks=1, ksp=None):
    """defaults of src/timestepper.F90:1992-2020, 1572-1573"""
    o = NewtonOpts()
    o.max_iterations, o.min_iterations = max_iterations, min_iterations
    o.rel_tol, o.abs_tol, o.update_rel_tol, o.update_abs_tol = rel_tol, abs_tol, update_rel_tol, update_abs_tol
    o.fd_err, o.fd_umin, o.pc_type, o.pc_nblocks = fd_err, fd_umin, pc_type, pc_nblocks
    o.ksp = ksp if ksp is not None else ksp_opts()
    return o


class Mat:
    """BAIJ matrix on the GPU (MatCreateBAIJ / MatSetValuesBlocked / MatMult)."""

    def __init__(self, sim, handle, nb, bs, owned=True):
        self.sim, self.h, self.nb, self.bs, self.owned = sim, handle, nb, bs, owned
        self.L = _lib.lib()

    @classmethod
    def create(cls, sim, nb, ncolb, bs, rowptr, colidx, vals=None):
        L = _lib.lib()
        h = C.c_void_p()
        rowptr = np.ascontiguousarray(rowptr, np.int32)
        colidx = np.ascontiguousarray(colidx, np.int32)
        check(L.wb_mat_create(sim.h, nb, ncolb, bs, len(colidx), ptr(rowptr), ptr(colidx), ptr(vals), C.byref(h)),
              "wb_mat_create")
        return cls(sim, h, nb, bs)

    def set_values(self, vals):
        check(self.L.wb_mat_set_values(self.h, ptr(vals)), "wb_mat_set_values")

    def get_values(self, nnzb):
        vals = np.zeros(nnzb * self.bs * self.bs)
        check(self.L.wb_mat_get_values(self.h, ptr(vals)), "wb_mat_get_values")
        return vals

    def mult(self, x, y):
        check(self.L.wb_mat_mult(self.h, ptr(x), ptr(y)), "wb_mat_mult")
        return y

    def destroy(self):
        if self.owned and self.h:
            self.L.wb_mat_destroy(self.h)
            self.h = None


class PC:
    """PCSetUp / PCApply (pbjacobi, bjacobi + ILU(0))."""

    def __init__(self, mat, type=PC_BJACOBI_ILU0, nblocks=1, block_of_row=None):
        self.L = _lib.lib()
        self.mat = mat
        self.h = C.c_void_p()
        bor = None if block_of_row is None else np.ascontiguousarray(block_of_row, np.int32)
        self.rc = check(self.L.wb_pc_setup(mat.h, type, nblocks, ptr(bor), C.byref(self.h)), "wb_pc_setup")

    def refactor(self):
        return check(self.L.wb_pc_refactor(self.h), "wb_pc_refactor")

    def apply(self, r, z):
        check(self.L.wb_pc_apply(self.h, ptr(r), ptr(z)), "wb_pc_apply")
        return z

    def destroy(self):
        if self.h:
            self.L.wb_pc_destroy(self.h)
            self.h = None


def ksp_solve(mat, pc, b, x, opts=None):
    """KSPSolve; returns (reason, iterations, residual norm)."""
    L = _lib.lib()
    o = opts if opts is not None else ksp_opts()
    its, reason, rn = C.c_int(), C.c_int(), C.c_double()
    check(L.wb_ksp_solve(mat.h, pc.h, C.byref(o), ptr(b), ptr(x), C.byref(its), C.byref(reason), C.byref(rn)),
          "wb_ksp_solve")
    return reason.value, its.value, rn.value


class FlowSimulation:
    """The flow-simulation ODE on one GPU (one instance per rank)."""

    def __init__(self, params, mesh, device=0):
        L = _lib.lib()
        self.L = L
        self.params = params
        self.mesh = mesh
        self.h = C.c_void_p()
        check(L.wb_create(C.byref(params), device, C.byref(self.h)), "wb_create")
        self.np = L.wb_num_primary(self.h)
        self.dof = L.wb_fluid_dof(self.h)
        self._keep = (np.ascontiguousarray(mesh.face_cells, np.int32), np.ascontiguousarray(mesh.face_geom, np.float64),
                      np.ascontiguousarray(mesh.cell_geom, np.float64), np.ascontiguousarray(mesh.rock, np.float64))
        check(L.wb_set_mesh(self.h, mesh.ncell, mesh.ninterior, mesh.nowned, len(self._keep[0]),
                            ptr(self._keep[0]), ptr(self._keep[1]), ptr(self._keep[2]), ptr(self._keep[3])), "wb_set_mesh")
        self.ncell, self.ninterior, self.nowned = mesh.ncell, mesh.ninterior, mesh.nowned
        self.n = self.nowned * self.np
        check(L.wb_set_global_offset(self.h, mesh.first_cell, mesh.ncell_global or mesh.nowned), "wb_set_global_offset")

    # ---- multi-GPU plumbing: NCCL id is broadcast by the caller (torch.distributed)
    def comm_init(self, rank, nranks, unique_id):
        check(self.L.wb_comm_init(self.h, rank, nranks, ptr(unique_id)), "wb_comm_init")
        m = self.mesh
        if m.neigh_rank is not None and len(m.neigh_rank):
            check(self.L.wb_set_halo(self.h, len(m.neigh_rank), ptr(m.neigh_rank), ptr(m.send_ptr), ptr(m.send_idx),
                                     ptr(m.recv_ptr), ptr(m.recv_idx)), "wb_set_halo")

    # NVLink peer-to-peer path: export -> all-gather by the caller -> open
    def p2p_export(self):
        blob = np.zeros(self.L.wb_comm_p2p_blob_size(), np.uint8)
        check(self.L.wb_comm_p2p_export(self.h, ptr(blob)), "wb_comm_p2p_export")
        return blob

    def p2p_open(self, blobs):
        b = np.ascontiguousarray(blobs, np.uint8)
        check(self.L.wb_comm_p2p_open(self.h, ptr(b)), "wb_comm_p2p_open")
        return bool(self.L.wb_comm_p2p_enabled(self.h))

    def p2p_setup(self, dist, device):
        """convenience for torch.distributed hosts: all-gather the blobs and open them"""
        import torch
        mine = torch.from_numpy(self.p2p_export()).to(device)
        parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, mine)
        return self.p2p_open(torch.cat(parts).cpu().numpy())

    @staticmethod
    def unique_id():
        uid = np.zeros(128, np.uint8)
        check(_lib.lib().wb_comm_unique_id(ptr(uid)), "wb_comm_unique_id")
        return uid

    def destroy(self):
        if self.h:
            self.L.wb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # ---- state
    def fluid_init(self, y, region):
        return check(self.L.wb_fluid_init(self.h, ptr(y), ptr(np.ascontiguousarray(region, np.int32))), "wb_fluid_init")

    def set_boundaries(self, ghost_cells, interior_cells, primary, region):
        g = np.ascontiguousarray(ghost_cells, np.int32)
        ic = np.ascontiguousarray(interior_cells, np.int32)
        pr = np.ascontiguousarray(primary, np.float64)
        rg = np.ascontiguousarray(region, np.int32)
        return check(self.L.wb_set_boundaries(self.h, len(g), ptr(g), ptr(ic), ptr(pr), ptr(rg)), "wb_set_boundaries")

    def set_rock(self, rock):
        """rock records [ninterior, 8] of the interior cells replaced between time steps (time-dependent permeability /
        porosity: flow_simulation_update_rock_properties, src/flow_simulation.F90:2051-2089); boundary ghost cells keep
        the records they copied at set-up"""
        r = np.ascontiguousarray(rock, np.float64).reshape(-1)
        assert r.size == 8 * self.mesh.ninterior, "set_rock: %d values for %d interior cells" % (r.size, self.mesh.ninterior)
        return check(self.L.wb_set_rock(self.h, ptr(r)), "wb_set_rock")

    def set_method(self, method, dt_last=0.0, lhs_last2=None):
        """time-stepping residual form (timestepper.F90:345-452): METHOD_BEULER / METHOD_BDF2 / METHOD_DIRECTSS"""
        return check(self.L.wb_set_method(self.h, method, dt_last, ptr(lhs_last2)), "wb_set_method")

    def set_sources(self, cells, components, rates, enthalpies):
        """fixed-rate sources / sinks (source.F90:375-480); see wb_set_sources"""
        c = np.ascontiguousarray(cells, np.int32)
        k = np.ascontiguousarray(components, np.int32)
        r = np.ascontiguousarray(rates, np.float64)
        h = np.ascontiguousarray(enthalpies, np.float64)
        self.nsrc = len(c)
        return check(self.L.wb_set_sources(self.h, len(c), ptr(c), ptr(k), ptr(r), ptr(h)), "wb_set_sources")

    def set_source_components(self, injection_components, production_components):
        """injection / production component of every source, chosen by the sign of the rate at every evaluation"""
        a = np.ascontiguousarray(injection_components, np.int32)
        b = np.ascontiguousarray(production_components, np.int32)
        return check(self.L.wb_set_source_components(self.h, len(a), ptr(a), ptr(b)), "wb_set_source_components")

    def set_source_controls(self, sources, productivity, reference_pressure, direction=None, limit=None):
        """deliverability / direction / total-limiter controls of some of the sources; see wb_set_source_controls"""
        s = np.ascontiguousarray(sources, np.int32)
        pi = np.ascontiguousarray(productivity, np.float64)
        pr = np.ascontiguousarray(reference_pressure, np.float64)
        dr = None if direction is None else np.ascontiguousarray(direction, np.int32)
        lm = None if limit is None else np.ascontiguousarray(limit, np.float64)
        return check(self.L.wb_set_source_controls(self.h, len(s), ptr(s), ptr(pi), ptr(pr), ptr(dr), ptr(lm)),
                     "wb_set_source_controls")

    def set_source_recharge(self, sources, coefficient, reference_pressure):
        """recharge / injectivity controls: rate = -coefficient (P - reference pressure); see wb_set_source_recharge"""
        s = np.ascontiguousarray(sources, np.int32)
        cf = np.ascontiguousarray(coefficient, np.float64)
        pr = np.ascontiguousarray(reference_pressure, np.float64)
        return check(self.L.wb_set_source_recharge(self.h, len(s), ptr(s), ptr(cf), ptr(pr)), "wb_set_source_recharge")

    def set_source_separators(self, sources, pressures, limit_water=None, limit_steam=None):
        """separators (per source a list of 0, 1 or 2 stage pressures) and limits on the separated water / steam rates;
        see wb_set_source_separators"""
        s = np.ascontiguousarray(sources, np.int32)
        ns = np.array([len(p) for p in pressures], np.int32)
        pr = np.zeros(2 * max(len(s), 1))
        for k, p in enumerate(pressures):
            pr[2 * k:2 * k + len(p)] = p
        lw = None if limit_water is None else np.ascontiguousarray(limit_water, np.float64)
        ls = None if limit_steam is None else np.ascontiguousarray(limit_steam, np.float64)
        return check(self.L.wb_set_source_separators(self.h, len(s), ptr(s), ptr(ns), ptr(pr), ptr(lw), ptr(ls)),
                     "wb_set_source_separators")

    def set_source_pressure_table(self, sources, tables, coordinate=None, step=None):
        """reference pressure of sources on deliverability tabulated against the flowing enthalpy (coordinate 0, the
        default) or the pressure (1) of their cells: tables[k] = [(x, y), ...]; see wb_set_source_pressure_table"""
        s = np.ascontiguousarray(sources, np.int32)
        n = len(s)
        npts = np.array([len(t) for t in tables], np.int32)
        tab = np.zeros((max(n, 1), 2 * PRESSURE_TABLE_MAX))
        for k, t in enumerate(tables):
            tab[k, :2 * len(t)] = np.asarray(t, np.float64).reshape(-1)
        co = np.zeros(max(n, 1), np.int32) if coordinate is None else np.ascontiguousarray(coordinate, np.int32)
        st = np.zeros(max(n, 1), np.int32) if step is None else np.ascontiguousarray(step, np.int32)
        return check(self.L.wb_set_source_pressure_table(self.h, n, ptr(s), ptr(co), ptr(st), ptr(npts), ptr(tab)),
                     "wb_set_source_pressure_table")

    def separator_stage(self, pressure):
        hw, hs = C.c_double(), C.c_double()
        check(self.L.wb_separator_stage(self.h, float(pressure), C.byref(hw), C.byref(hs)), "wb_separator_stage")
        return hw.value, hs.value

    def source_separated(self):
        """[nsources][5]: water rate, water enthalpy, steam rate, steam enthalpy, steam fraction"""
        out = np.zeros((self.nsrc, 5))
        check(self.L.wb_get_source_separated(self.h, ptr(out)), "wb_get_source_separated")
        return out

    def source_rates(self):
        r = np.zeros(self.nsrc)
        check(self.L.wb_get_source_rates(self.h, ptr(r)), "wb_get_source_rates")
        return r

    def fluid(self):
        out = np.zeros((self.ncell, self.dof))
        check(self.L.wb_get_fluid(self.h, ptr(out)), "wb_get_fluid")
        return out

    def regions(self):
        r = np.zeros(self.ncell, np.int32)
        check(self.L.wb_get_regions(self.h, ptr(r)), "wb_get_regions")
        return r

    def pre_iteration(self):
        check(self.L.wb_pre_iteration(self.h))

    def pre_timestep(self):
        check(self.L.wb_pre_timestep(self.h))

    def pre_retry_timestep(self):
        check(self.L.wb_pre_retry_timestep(self.h))

    # ---- ode hooks
    def pre_eval(self, y, perturbed_columns=None):
        pc = None if perturbed_columns is None else np.ascontiguousarray(perturbed_columns, np.int32)
        return check(self.L.wb_pre_eval(self.h, ptr(y), ptr(pc), 0 if pc is None else len(pc)), "wb_pre_eval")

    def lhs(self, y, out=None):
        """pre_eval + cell_balances, as timestepper's residual routines call them"""
        out = np.zeros(self.n) if out is None else out
        err = self.pre_eval(y)
        if err == 0:
            err = check(self.L.wb_cell_balances(self.h, ptr(out)), "wb_cell_balances")
        return err, out

    def rhs(self, out=None):
        out = np.zeros(self.n) if out is None else out
        err = check(self.L.wb_cell_inflows(self.h, ptr(out)), "wb_cell_inflows")
        return err, out

    def residual(self, y, lhs_last, dt, perturbed=None, lhs=None, rhs=None, r=None, want_parts=True):
        if want_parts:
            lhs = np.zeros(self.n) if lhs is None else lhs
            rhs = np.zeros(self.n) if rhs is None else rhs
        r = np.zeros(self.n) if r is None else r
        pc = None if perturbed is None else np.ascontiguousarray(perturbed, np.int32)
        err = check(self.L.wb_residual_be(self.h, ptr(y), ptr(lhs_last), dt, ptr(pc), 0 if pc is None else len(pc),
                                          ptr(lhs), ptr(rhs), ptr(r)), "wb_residual_be")
        return err, lhs, rhs, r

    def max_scaled(self, v, scale, tol):
        mv, ml = C.c_double(), C.c_int64()
        check(self.L.wb_max_scaled(self.h, ptr(v), ptr(scale), tol, C.byref(mv), C.byref(ml)), "wb_max_scaled")
        return mv.value, ml.value

    # ---- Jacobian
    def jacobian_pattern(self):
        nb, bs, nnzb = C.c_int(), C.c_int(), C.c_int()
        check(self.L.wb_jacobian_pattern(self.h, C.byref(nb), C.byref(bs), C.byref(nnzb), None, None, None))
        rowptr, colidx = np.zeros(nb.value + 1, np.int32), np.zeros(nnzb.value, np.int32)
        check(self.L.wb_jacobian_get(self.h, ptr(rowptr), ptr(colidx), None), "wb_jacobian_get")
        return nb.value, bs.value, rowptr, colidx

    def cell_faces(self):
        """cell -> face gather lists (wb_cell_faces_get): cf_ptr, cf_face (2*face + side), cf_other"""
        ncf = C.c_int()
        check(self.L.wb_cell_faces_get(self.h, C.byref(ncf), None, None, None), "wb_cell_faces_get")
        p, f, o = np.zeros(self.nowned + 1, np.int32), np.zeros(ncf.value, np.int32), np.zeros(ncf.value, np.int32)
        check(self.L.wb_cell_faces_get(self.h, C.byref(ncf), ptr(p), ptr(f), ptr(o)), "wb_cell_faces_get")
        return p, f, o

    def jacobian_values(self):
        nb, bs, rowptr, colidx = self.jacobian_pattern()
        vals = np.zeros((len(colidx), bs * bs))
        check(self.L.wb_jacobian_get(self.h, None, None, ptr(vals)), "wb_jacobian_get")
        return vals

    def jacobian(self, y, lhs_last, dt, fd_err=1e-8, fd_umin=1e-2, colored=False, vals_out=None):
        if colored:
            nc = C.c_int()
            err = check(self.L.wb_jacobian_be_colored(self.h, ptr(y), ptr(lhs_last), dt, fd_err, fd_umin, ptr(vals_out),
                                                      C.byref(nc)), "wb_jacobian_be_colored")
            self.ncolors = nc.value
            return err
        return check(self.L.wb_jacobian_be(self.h, ptr(y), ptr(lhs_last), dt, fd_err, fd_umin, ptr(vals_out)),
                     "wb_jacobian_be")

    def jacobian_mat(self):
        h = C.c_void_p()
        check(self.L.wb_jacobian_mat(self.h, C.byref(h)))
        return Mat(self, h, self.nowned, self.np, owned=False)

    # ---- passive tracers: the auxiliary linear problem (flow_simulation.F90:1489-1959, timestepper.F90:458-581)
    def set_tracers(self, phases, diffusion=None, decay=None, activation=None):
        """setup_tracers (tracer.F90:64-150); phases are 1-based phase indices"""
        ph = np.ascontiguousarray(phases, np.int32)
        arr = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (diffusion, decay, activation)]
        self.nt = len(ph)
        return check(self.L.wb_set_tracers(self.h, len(ph), ptr(ph), ptr(arr[0]), ptr(arr[1]), ptr(arr[2])),
                     "wb_set_tracers")

    def set_tracer_injection(self, rates):
        r = None if rates is None else np.ascontiguousarray(rates, np.float64)
        return check(self.L.wb_set_tracer_injection(self.h, ptr(r)), "wb_set_tracer_injection")

    def tracer_balances(self):
        """aux_lhs: porosity * saturation * density of each tracer's phase, [nowned*nt]"""
        al = np.zeros(self.nowned * self.nt)
        check(self.L.wb_tracer_cell_balances(self.h, ptr(al)), "wb_tracer_cell_balances")
        return al

    def tracer_setup_linear(self, dt, al_last, x_last, x_boundary=None, al_last2=None, x_last2=None):
        """setup_linear + aux_pre_solve; returns (A, b, Al) with A a borrowed Mat (bs = nt)"""
        n = self.nowned * self.nt
        al, b = np.zeros(n), np.zeros(n)
        h = C.c_void_p()
        xb = None if x_boundary is None else np.ascontiguousarray(x_boundary, np.float64)
        check(self.L.wb_tracer_setup_linear(self.h, dt, ptr(al_last), ptr(x_last), ptr(al_last2), ptr(x_last2), ptr(xb),
                                            ptr(al), ptr(b), C.byref(h)), "wb_tracer_setup_linear")
        return Mat(self, h, self.nowned, self.nt, owned=False), b, al

    def tracer_solve(self, dt, al_last, x_last, x_boundary=None, al_last2=None, x_last2=None, opts=None,
                     pc_type=PC_BJACOBI_ILU0, pc_nblocks=1):
        """the auxiliary step of timestepper_step (timestepper.F90:2347-2353); returns (x, Al, reason, iterations)"""
        n = self.nowned * self.nt
        al, x = np.zeros(n), np.zeros(n)
        o = opts if opts is not None else ksp_opts()
        its, reason = C.c_int(), C.c_int()
        xb = None if x_boundary is None else np.ascontiguousarray(x_boundary, np.float64)
        check(self.L.wb_tracer_solve(self.h, C.byref(o), pc_type, pc_nblocks, dt, ptr(al_last), ptr(x_last),
                                     ptr(al_last2), ptr(x_last2), ptr(xb), ptr(al), ptr(x), C.byref(its),
                                     C.byref(reason)), "wb_tracer_solve")
        return x, al, reason.value, its.value

    # ---- transitions
    def fluid_transitions(self, y_old, search, y):
        cs, cy = C.c_int(), C.c_int()
        err = check(self.L.wb_fluid_transitions(self.h, ptr(y_old), ptr(search), ptr(y), C.byref(cs), C.byref(cy)),
                    "wb_fluid_transitions")
        return err, cs.value, cy.value

    # ---- Newton
    def set_pc_blocks(self, block_of_row):
        bor = None if block_of_row is None else np.ascontiguousarray(block_of_row, np.int32)
        check(self.L.wb_set_pc_blocks(self.h, ptr(bor)), "wb_set_pc_blocks")

    def newton_solve(self, y, lhs_last, dt, opts=None):
        o = opts if opts is not None else newton_opts()
        res = NewtonResult()
        check(self.L.wb_newton_solve_be(self.h, C.byref(o), dt, ptr(lhs_last), ptr(y), C.byref(res)), "wb_newton_solve_be")
        return res

    # ---- instrumentation
    def timer(self, name):
        ms, cnt = C.c_double(), C.c_int64()
        self.L.wb_timer_get(self.h, name.encode(), C.byref(ms), C.byref(cnt))
        return ms.value, cnt.value

    def ksp_breakdown(self, reset=True):
        """device time per Krylov iteration of the persistent GMRES kernel's phases (microseconds, CTA 0), or None
        when no fused solve ran since the last reset (wb_ksp_fused_profile)"""
        ns = (C.c_double * 7)()
        check(self.L.wb_ksp_fused_profile(self.h, ns, 1 if reset else 0), "wb_ksp_fused_profile")
        its = ns[6]
        if its <= 0:
            return None
        names = ("spmv_pc", "dots", "dots_reduce", "maxpy_norm_push", "norm_reduce_update", "other")
        out = {k: round(ns[i] * 1e-3 / its, 3) for i, k in enumerate(names)}
        out["iterations"] = int(its)
        return out

    def ksp_breakdown_ctas(self):
        """tuning aid: min / mean / max over the CTAs of the persistent kernel's phase times per iteration (us);
        call before ksp_breakdown(reset=True)"""
        L = self.L
        if not hasattr(L, "wb_debug_fused_profile_all"):
            return None
        L.wb_debug_fused_profile_all.restype = C.c_int
        L.wb_debug_fused_profile_all.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        buf = (C.c_double * (16 * 256))()
        n = L.wb_debug_fused_profile_all(self.h, buf, 256)
        a = np.array(buf[:16 * n]).reshape(n, 16)
        a = a[a[:, 6] > 0]
        if len(a) == 0:
            return None
        per = a[:, :6] * 1e-3 / a[:, 6:7]
        names = ("spmv_pc", "dots", "dots_reduce", "maxpy_norm_push", "norm_reduce_update", "other")
        out = {k: [round(float(per[:, i].min()), 2), round(float(per[:, i].mean()), 2), round(float(per[:, i].max()), 2)]
               for i, k in enumerate(names)}
        sp = a[:, 7] * 1e-3 / a[:, 6]
        out["spmv_share_of_spmv_pc"] = [round(float(sp.min()), 2), round(float(sp.mean()), 2), round(float(sp.max()), 2)]
        for k, nm in ((9, "level_first_wait_cycles"), (10, "level_own_cycles"), (11, "level_barrier_cycles")):
            v = a[:, k] / a[:, 6]
            out[nm] = [round(float(v.min())), round(float(v.mean())), round(float(v.max()))]
        return out

    def launches(self):
        return self.L.wb_launch_count(self.h)

    def stream(self):
        return self.L.wb_stream(self.h)
