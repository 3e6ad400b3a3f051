"""ctypes binding of the CPU oracle (oracle/_build/libwaiwera_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(waiwera_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libwaiwera_oracle.so")

WO_MAX_TABLE = 16
THERMO_IAPWS, THERMO_IFC67 = 0, 1
EOS_WE, EOS_W, EOS_WCE, EOS_WAE = 0, 1, 2, 3
RP_FULLY_MOBILE, RP_LINEAR, RP_PICKENS, RP_COREY, RP_GRANT, RP_VAN_GENUCHTEN, RP_TABLE = range(7)
CP_ZERO, CP_LINEAR, CP_VAN_GENUCHTEN, CP_TABLE = range(4)
PC_NONE, PC_PBJACOBI, PC_BJACOBI_ILU0, PC_ASM_ILU0 = 0, 1, 2, 3
KSP_GMRES, KSP_BCGS = 0, 1

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class Relperm(C.Structure):
    _fields_ = [("type", C.c_int), ("p", C.c_double * 8), ("nl", C.c_int), ("nv", C.c_int),
                ("lx", C.c_double * WO_MAX_TABLE), ("ly", C.c_double * WO_MAX_TABLE),
                ("vx", C.c_double * WO_MAX_TABLE), ("vy", C.c_double * WO_MAX_TABLE)]


class Cappress(C.Structure):
    _fields_ = [("type", C.c_int), ("p", C.c_double * 8), ("n", C.c_int),
                ("x", C.c_double * WO_MAX_TABLE), ("y", C.c_double * WO_MAX_TABLE)]


class Params(C.Structure):
    _fields_ = [("eos", C.c_int), ("thermo", C.c_int), ("extrapolate", C.c_int),
                ("pressure_scale", C.c_double), ("temperature_scale", C.c_double),
                ("partial_pressure_scale", C.c_double), ("eos_w_temperature", C.c_double),
                ("relperm", Relperm), ("cappress", Cappress), ("gravity", C.c_double * 3)]


class Mesh(C.Structure):
    _fields_ = [("ncell", C.c_int), ("ninterior", C.c_int), ("nowned", C.c_int), ("nface", C.c_int),
                ("face_cells", c_ip), ("face_geom", c_dp), ("cell_geom", c_dp), ("rock", c_dp)]


class Bsr(C.Structure):
    _fields_ = [("nb", C.c_int), ("bs", C.c_int), ("nnzb", C.c_int),
                ("rowptr", c_ip), ("colidx", c_ip), ("val", c_dp)]


class Tracer(C.Structure):
    _fields_ = [("phase", C.c_int), ("diffusion", C.c_double), ("decay", C.c_double), ("activation", C.c_double)]


class KspOpts(C.Structure):
    _fields_ = [("type", C.c_int), ("restart", C.c_int), ("maxit", C.c_int),
                ("rtol", C.c_double), ("atol", C.c_double), ("dtol", C.c_double)]


class NewtonOpts(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("min_iterations", C.c_int),
                ("rel_tol", C.c_double), ("abs_tol", C.c_double),
                ("update_rel_tol", C.c_double), ("update_abs_tol", C.c_double),
                ("fd_err", C.c_double), ("fd_umin", C.c_double), ("pc_type", C.c_int), ("ksp", KspOpts)]


class NewtonResult(C.Structure):
    _fields_ = [("reason", C.c_int), ("iterations", C.c_int), ("linear_iterations", C.c_int),
                ("max_residual", C.c_double * 32), ("lin_its", C.c_int * 32),
                ("lin_reason", C.c_int * 32), ("lin_rnorm", C.c_double * 32)]


def _cpu_key():
    """identifies the host CPU (model + ISA flags): the library is built -march=native, and the build directory
    travels from the build container to the GPU box, whose CPU may differ"""
    import hashlib
    model, flags = "", ""
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name") and not model:
                    model = line.split(":", 1)[1].strip()
                elif line.startswith("flags") and not flags:
                    flags = " ".join(sorted(line.split(":", 1)[1].split()))
                if model and flags:
                    break
    except OSError:
        pass
    return hashlib.sha1((model + "|" + flags).encode()).hexdigest()


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile); rebuilt from scratch when the host CPU is not the one the
    existing library was built for."""
    stamp = os.path.join(_HERE, "_build", "cpu.stamp")
    key = _cpu_key()
    try:
        same = open(stamp).read().strip() == key
    except OSError:
        same = False
    if force or not same:
        subprocess.check_call(["make", "-C", _HERE, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    with open(stamp, "w") as fh:
        fh.write(key + "\n")
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()  # no-op when up to date; rebuilds on a different host CPU
    L = C.CDLL(_LIB)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    sig = {
        "wo_set_num_threads": (i, [i]),
        "wo_powertable_eval": (None, [c_ip, i, d, c_ip, i, c_dp]),
        "wo_thermo_create": (vp, [i, i]),
        "wo_thermo_destroy": (None, [vp]),
        "wo_region_properties": (i, [vp, i, c_dp, c_dp]),
        "wo_region_viscosity": (d, [vp, i, d, d, d]),
        "wo_saturation_pressure": (i, [vp, d, c_dp]),
        "wo_saturation_temperature": (i, [vp, d, c_dp]),
        "wo_phase_composition": (i, [vp, i, d, d]),
        "wo_boundary23_pressure": (d, [d]),
        "wo_boundary23_temperature": (d, [d]),
        "wo_relperm_values": (None, [C.POINTER(Relperm), d, c_dp]),
        "wo_cappress_value": (d, [C.POINTER(Cappress), d, d]),
        "wo_eos_create": (vp, [C.POINTER(Params)]),
        "wo_eos_destroy": (None, [vp]),
        "wo_eos_num_primary": (i, [vp]),
        "wo_eos_fluid_dof": (i, [vp]),
        "wo_eos_scale": (None, [vp, c_dp, i, c_dp]),
        "wo_eos_unscale": (None, [vp, c_dp, i, c_dp]),
        "wo_eos_bulk_properties": (i, [vp, c_dp, c_dp]),
        "wo_eos_phase_properties": (i, [vp, c_dp, c_dp, c_dp]),
        "wo_eos_transition": (i, [vp, c_dp, c_dp, c_dp, c_dp, C.POINTER(i)]),
        "wo_eos_check_primary_variables": (i, [vp, c_dp, c_dp, C.POINTER(i)]),
        "wo_eos_conductivity": (d, [c_dp, c_dp, i]),
        "wo_co2_properties": (None, [d, d, c_dp]),
        "wo_co2_henrys_constant": (d, [d]),
        "wo_co2_energy_solution": (d, [d, d]),
        "wo_co2_viscosity": (i, [d, d, c_dp]),
        "wo_air_properties": (None, [d, d, c_dp]),
        "wo_air_henrys_constant": (d, [d, c_dp]),
        "wo_air_energy_solution": (d, [d, c_dp]),
        "wo_air_mixture_viscosity": (d, [d, d, d, i]),
        "wo_cell_balance": (None, [c_dp, c_dp, i, i, i, c_dp]),
        "wo_face_flux": (None, [c_dp, c_dp, c_dp, c_dp, c_dp, i, i, i, i, i, c_dp]),
        "wo_face_calculate_distances": (None, [c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
        "wo_face_harmonic_average": (d, [c_dp, c_dp]),
        "wo_flow_create": (vp, [C.POINTER(Params), C.POINTER(Mesh)]),
        "wo_flow_destroy": (None, [vp]),
        "wo_flow_eos": (vp, [vp]),
        "wo_flow_fluid": (c_dp, [vp]),
        "wo_flow_current_fluid": (c_dp, [vp]),
        "wo_flow_flux": (c_dp, [vp]),
        "wo_flow_fluid_init": (i, [vp, c_dp, c_ip]),
        "wo_flow_set_boundary": (i, [vp, i, i, c_dp, i]),
        "wo_flow_set_rock": (i, [vp, c_dp]),
        "wo_flow_set_sources": (None, [vp, i, c_ip, c_ip, c_dp, c_dp]),
        "wo_flow_set_source_components": (None, [vp, i, c_ip, c_ip]),
        "wo_flow_set_method": (None, [vp, i, d, c_dp]),
        "wo_flow_set_source_controls": (None, [vp, i, c_ip, c_dp, c_dp, c_ip, c_dp]),
        "wo_flow_set_source_recharge": (None, [vp, i, c_ip, c_dp, c_dp]),
        "wo_flow_get_source_rates": (None, [vp, c_dp]),
        "wo_separator_stage": (i, [vp, d, c_dp, c_dp]),
        "wo_separate": (None, [i, c_dp, d, d, c_dp]),
        "wo_flow_set_source_separators": (i, [vp, i, c_ip, c_ip, c_dp, c_dp, c_dp]),
        "wo_flow_set_source_pressure_table": (i, [vp, i, c_ip, c_ip, c_ip, c_ip, c_dp]),
        "wo_flow_source_separated": (None, [vp, i, d, c_dp]),
        "wo_flow_pre_eval": (i, [vp, c_dp, c_ip, i]),
        "wo_flow_cell_balances": (i, [vp, c_dp]),
        "wo_flow_cell_inflows": (i, [vp, c_dp]),
        "wo_flow_pre_iteration": (None, [vp]),
        "wo_flow_pre_timestep": (None, [vp]),
        "wo_flow_pre_retry_timestep": (None, [vp]),
        "wo_flow_fluid_transitions": (i, [vp, c_dp, c_dp, c_dp, C.POINTER(i), C.POINTER(i)]),
        "wo_flow_get_regions": (None, [vp, c_ip]),
        "wo_residual_be": (i, [vp, c_dp, c_dp, d, c_ip, i, c_dp, c_dp, c_dp]),
        "wo_flow_set_tracers": (None, [vp, i, C.POINTER(Tracer)]),
        "wo_flow_set_tracer_injection": (None, [vp, c_dp]),
        "wo_tracer_decay": (d, [C.POINTER(Tracer), d]),
        "wo_tracer_pattern": (C.POINTER(Bsr), [C.POINTER(Mesh), i]),
        "wo_tracer_cell_balances": (None, [vp, c_dp]),
        "wo_tracer_cell_inflows": (None, [vp, C.POINTER(Bsr), c_dp]),
        "wo_tracer_pre_solve": (None, [vp, C.POINTER(Bsr), c_dp, c_dp]),
        "wo_tracer_setup_linear": (None, [vp, i, d, d, c_dp, c_dp, c_dp, c_dp, C.POINTER(Bsr), c_dp, c_dp]),
        "wo_vec_max_pointwise_abs_scale": (None, [c_dp, c_dp, d, i, c_dp, C.POINTER(i)]),
        "wo_bsr_from_mesh": (C.POINTER(Bsr), [C.POINTER(Mesh), i]),
        "wo_bsr_destroy": (None, [C.POINTER(Bsr)]),
        "wo_bsr_spmv": (None, [C.POINTER(Bsr), c_dp, c_dp]),
        "wo_bsr_coloring": (i, [C.POINTER(Bsr), c_ip]),
        "wo_fd_jacobian": (i, [vp, c_dp, c_dp, d, c_dp, c_ip, i, d, d, C.POINTER(Bsr)]),
        "wo_pc_create": (vp, [C.POINTER(Bsr), i, c_ip]),
        "wo_pc_apply": (None, [vp, c_dp, c_dp]),
        "wo_pc_destroy": (None, [vp]),
        "wo_ksp_solve": (i, [C.POINTER(Bsr), vp, C.POINTER(KspOpts), c_dp, c_dp, C.POINTER(i), c_dp]),
        "wo_newton_solve_be": (i, [vp, C.POINTER(Bsr), c_ip, i, c_ip, C.POINTER(NewtonOpts), d, c_dp, c_dp,
                                   C.POINTER(NewtonResult)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def dp(a):
    """double* view of a contiguous float64 numpy array (or None)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_dp)


def ip(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_ip)


def make_relperm(kind="linear", **kw):
    r = Relperm()
    if kind == "fully_mobile":
        r.type = RP_FULLY_MOBILE
    elif kind == "linear":
        r.type = RP_LINEAR
        liq = kw.get("liquid", (0.0, 1.0))
        vap = kw.get("vapour", (0.0, 1.0))
        r.p[0], r.p[1], r.p[2], r.p[3] = liq[0], liq[1], vap[0], vap[1]
    elif kind == "pickens":
        r.type = RP_PICKENS
        r.p[0] = kw.get("power", 1.0)
    elif kind in ("corey", "grant"):
        r.type = RP_COREY if kind == "corey" else RP_GRANT
        r.p[0], r.p[1] = kw.get("slr", 0.3), kw.get("ssr", 0.05 if kind == "corey" else 0.6)
    elif kind == "van_genuchten":
        r.type = RP_VAN_GENUCHTEN
        r.p[0], r.p[1], r.p[2] = kw.get("lambda", kw.get("lambda_", 0.45)), kw.get("slr", 1e-3), kw.get("sls", 1.0)
        r.p[3] = 1.0 if kw.get("sum_unity", True) else 0.0
        r.p[4] = kw.get("ssr", 0.6)
    elif kind == "table":
        r.type = RP_TABLE
        liq, vap = kw["liquid"], kw["vapour"]
        r.nl, r.nv = len(liq), len(vap)
        for k, (x, y) in enumerate(liq):
            r.lx[k], r.ly[k] = x, y
        for k, (x, y) in enumerate(vap):
            r.vx[k], r.vy[k] = x, y
    else:
        raise ValueError(kind)
    return r


def make_cappress(kind="zero", **kw):
    c = Cappress()
    if kind == "zero":
        c.type = CP_ZERO
    elif kind == "linear":
        c.type = CP_LINEAR
        lim = kw.get("saturation_limits", (0.0, 1.0))
        c.p[0], c.p[1], c.p[2] = lim[0], lim[1], kw.get("pressure", 0.125e5)
    elif kind == "van_genuchten":
        c.type = CP_VAN_GENUCHTEN
        c.p[0], c.p[1], c.p[2], c.p[3] = kw.get("P0", 0.125e5), kw.get("lambda", kw.get("lambda_", 0.45)), kw.get("slr", 1e-3), kw.get("sls", 1.0)
        c.p[4] = kw.get("Pmax", 0.0)
        c.p[5] = 1.0 if "Pmax" in kw else 0.0
    elif kind == "table":
        c.type = CP_TABLE
        pts = kw["pressure"]
        c.n = len(pts)
        for k, (x, y) in enumerate(pts):
            c.x[k], c.y[k] = x, y
    else:
        raise ValueError(kind)
    return c


def make_params(eos=EOS_WE, thermo=THERMO_IAPWS, relperm=None, cappress=None, gravity=(0.0, 0.0, -9.8),
                extrapolate=0, eos_w_temperature=20.0, partial_pressure_scale=0.0):
    p = Params()
    p.eos, p.thermo, p.extrapolate = eos, thermo, extrapolate
    p.pressure_scale, p.temperature_scale = 1.0e6, 1.0e2
    p.partial_pressure_scale = partial_pressure_scale  # <= 0: adaptive Pg/P (the reference default)
    p.eos_w_temperature = eos_w_temperature
    p.relperm = relperm if relperm is not None else make_relperm("linear")
    p.cappress = cappress if cappress is not None else make_cappress("zero")
    for k in range(3):
        p.gravity[k] = gravity[k]
    return p


class Flow:
    """Thin OO wrapper over wo_flow for the tests and the CPU baseline."""

    def __init__(self, params, ncell, ninterior, nowned, face_cells, face_geom, cell_geom, rock):
        L = lib()
        self.L = L
        self.params = params
        self._keep = (np.ascontiguousarray(face_cells, np.int32), np.ascontiguousarray(face_geom, np.float64),
                      np.ascontiguousarray(cell_geom, np.float64), np.ascontiguousarray(rock, np.float64))
        m = Mesh()
        m.ncell, m.ninterior, m.nowned, m.nface = ncell, ninterior, nowned, len(self._keep[0]) // 2
        m.face_cells, m.face_geom, m.cell_geom, m.rock = ip(self._keep[0]), dp(self._keep[1]), dp(self._keep[2]), dp(self._keep[3])
        self.mesh = m
        self.h = L.wo_flow_create(C.byref(params), C.byref(m))
        assert self.h
        self.eos = L.wo_flow_eos(self.h)
        self.np = L.wo_eos_num_primary(self.eos)
        self.dof = L.wo_eos_fluid_dof(self.eos)
        self.ncell, self.nowned, self.ninterior = ncell, nowned, ninterior
        self.n = nowned * self.np

    def __del__(self):
        try:
            self.L.wo_flow_destroy(self.h)
        except Exception:
            pass

    def fluid(self):
        return np.ctypeslib.as_array(self.L.wo_flow_fluid(self.h), shape=(self.ncell, self.dof))

    def current_fluid(self):
        return np.ctypeslib.as_array(self.L.wo_flow_current_fluid(self.h), shape=(self.ncell, self.dof))

    def fluid_init(self, y, region):
        return self.L.wo_flow_fluid_init(self.h, dp(y), ip(region))

    def set_boundary(self, ghost, interior, primary, region):
        pr = np.ascontiguousarray(primary, np.float64)
        return self.L.wo_flow_set_boundary(self.h, ghost, interior, dp(pr), region)

    def set_rock(self, rock):
        r = np.ascontiguousarray(rock, np.float64).reshape(-1)
        return self.L.wo_flow_set_rock(self.h, dp(r))

    def regions(self):
        r = np.zeros(self.ncell, np.int32)
        self.L.wo_flow_get_regions(self.h, ip(r))
        return r

    def set_method(self, method, dt_last=0.0, lhs_last2=None):
        self.L.wo_flow_set_method(self.h, method, dt_last, dp(lhs_last2))

    def set_sources(self, cells, components, rates, enthalpies):
        c = np.ascontiguousarray(cells, np.int32)
        k = np.ascontiguousarray(components, np.int32)
        r = np.ascontiguousarray(rates, np.float64)
        h = np.ascontiguousarray(enthalpies, np.float64)
        self.L.wo_flow_set_sources(self.h, len(c), ip(c), ip(k), dp(r), dp(h))

    def set_source_components(self, injection_components, production_components):
        a = np.ascontiguousarray(injection_components, np.int32)
        b = np.ascontiguousarray(production_components, np.int32)
        self.L.wo_flow_set_source_components(self.h, len(a), ip(a), ip(b))

    def set_source_controls(self, sources, productivity, reference_pressure, direction=None, limit=None):
        s = np.ascontiguousarray(sources, np.int32)
        pi = np.ascontiguousarray(productivity, np.float64)
        pr = np.ascontiguousarray(reference_pressure, np.float64)
        dr = None if direction is None else np.ascontiguousarray(direction, np.int32)
        lm = None if limit is None else np.ascontiguousarray(limit, np.float64)
        self.L.wo_flow_set_source_controls(self.h, len(s), ip(s), dp(pi), dp(pr), ip(dr), dp(lm))

    def set_source_recharge(self, sources, coefficient, reference_pressure):
        s = np.ascontiguousarray(sources, np.int32)
        self.L.wo_flow_set_source_recharge(self.h, len(s), ip(s), dp(np.ascontiguousarray(coefficient, np.float64)),
                                           dp(np.ascontiguousarray(reference_pressure, np.float64)))

    def set_source_separators(self, sources, pressures, limit_water=None, limit_steam=None):
        """pressures: per source a list of 0, 1 or 2 separator stage pressures"""
        s = np.ascontiguousarray(sources, np.int32)
        ns = np.array([len(p) for p in pressures], np.int32)
        pr = np.zeros(2 * len(s))
        for k, p in enumerate(pressures):
            pr[2 * k:2 * k + len(p)] = p
        lw = None if limit_water is None else np.ascontiguousarray(limit_water, np.float64)
        ls = None if limit_steam is None else np.ascontiguousarray(limit_steam, np.float64)
        return self.L.wo_flow_set_source_separators(self.h, len(s), ip(s), ip(ns), dp(pr), dp(lw), dp(ls))

    def set_source_pressure_table(self, sources, tables, coordinate=None, step=None):
        """tables: per source a list of (x, y) points; coordinate 0 = flowing enthalpy (default), 1 = pressure"""
        s = np.ascontiguousarray(sources, np.int32)
        n = len(s)
        npts = np.array([len(t) for t in tables], np.int32)
        tab = np.zeros((max(n, 1), 16))
        for k, t in enumerate(tables):
            tab[k, :2 * len(t)] = np.asarray(t, float).reshape(-1)
        co = np.zeros(n, np.int32) if coordinate is None else np.ascontiguousarray(coordinate, np.int32)
        st = np.zeros(n, np.int32) if step is None else np.ascontiguousarray(step, np.int32)
        return self.L.wo_flow_set_source_pressure_table(self.h, n, ip(s), ip(co), ip(st), ip(npts), dp(tab))

    def source_separated(self, s, rate):
        """water rate, water enthalpy, steam rate, steam enthalpy, steam fraction of source s at the given rate"""
        out = np.zeros(5)
        self.L.wo_flow_source_separated(self.h, int(s), float(rate), dp(out))
        return out

    def source_rates(self, n):
        r = np.zeros(n)
        self.L.wo_flow_get_source_rates(self.h, dp(r))
        return r

    def residual(self, y, lhs_last, dt, perturbed=None):
        lhs, rhs, r = np.zeros(self.n), np.zeros(self.n), np.zeros(self.n)
        npert = 0 if perturbed is None else len(perturbed)
        err = self.L.wo_residual_be(self.h, dp(y), dp(lhs_last), dt, ip(perturbed), npert, dp(lhs), dp(rhs), dp(r))
        return err, lhs, rhs, r

    def lhs(self, y):
        """pre_eval (unperturbed) + cell_balances."""
        out = np.zeros(self.n)
        err = self.L.wo_flow_pre_eval(self.h, dp(y), None, 0)
        if err == 0:
            err = self.L.wo_flow_cell_balances(self.h, dp(out))
        return err, out

    def bsr(self):
        return self.L.wo_bsr_from_mesh(C.byref(self.mesh), self.np)

    # ---- passive tracers (auxiliary linear problem) ----
    def set_tracers(self, phases, diffusion=None, decay=None, activation=None):
        nt = len(phases)
        arr = (Tracer * nt)()
        for k in range(nt):
            arr[k].phase = int(phases[k])
            arr[k].diffusion = 0.0 if diffusion is None else float(diffusion[k])
            arr[k].decay = 0.0 if decay is None else float(decay[k])
            arr[k].activation = 0.0 if activation is None else float(activation[k])
        self.nt = nt
        self.L.wo_flow_set_tracers(self.h, nt, arr)
        self.ntrows = self.nowned + (self.ncell - self.ninterior)

    def set_tracer_injection(self, rates):
        r = np.ascontiguousarray(rates, np.float64)
        self.L.wo_flow_set_tracer_injection(self.h, dp(r))

    def tracer_pattern(self):
        return self.L.wo_tracer_pattern(C.byref(self.mesh), self.nt)

    def tracer_balances(self):
        al = np.zeros(self.ntrows * self.nt)
        self.L.wo_tracer_cell_balances(self.h, dp(al))
        return al

    def tracer_setup_linear(self, A, dt, al_last, x_last, method=0, dt_last=0.0, al_last2=None, x_last2=None):
        """setup_linear + aux_pre_solve for the state of the last unperturbed evaluation; returns b, Al."""
        n = self.ntrows * self.nt
        b, al = np.zeros(n), np.zeros(n)
        self.L.wo_tracer_setup_linear(self.h, method, dt, dt_last, dp(al_last), dp(x_last), dp(al_last2), dp(x_last2),
                                      A, dp(b), dp(al))
        return b, al


def bsr_arrays(A):
    a = A.contents
    rowptr = np.ctypeslib.as_array(a.rowptr, shape=(a.nb + 1,))
    colidx = np.ctypeslib.as_array(a.colidx, shape=(a.nnzb,))
    val = np.ctypeslib.as_array(a.val, shape=(a.nnzb, a.bs * a.bs))
    return rowptr, colidx, val


def max_scaled(v, scale, tol):
    mv, ml = C.c_double(), C.c_int()
    lib().wo_vec_max_pointwise_abs_scale(dp(v), dp(scale), tol, len(v), C.byref(mv), C.byref(ml))
    return mv.value, ml.value
