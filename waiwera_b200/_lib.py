"""ctypes binding of libwaiwera_b200.so (the C ABI declared in include/waiwera_b200.h).

The library is hand-written CUDA for sm_100a; there is no other implementation behind this
module.  Loading fails loudly if the shared library has not been built (python -m
waiwera_b200.build / __graft_entry__.build()).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libwaiwera_b200.so")

WB_MAX_TABLE = 16
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
vp = C.c_void_p


class Relperm(C.Structure):
    _fields_ = [("type", C.c_int), ("p", C.c_double * 8), ("nl", C.c_int), ("nv", C.c_int),
                ("lx", C.c_double * WB_MAX_TABLE), ("ly", C.c_double * WB_MAX_TABLE),
                ("vx", C.c_double * WB_MAX_TABLE), ("vy", C.c_double * WB_MAX_TABLE)]


class Cappress(C.Structure):
    _fields_ = [("type", C.c_int), ("p", C.c_double * 8), ("n", C.c_int),
                ("x", C.c_double * WB_MAX_TABLE), ("y", C.c_double * WB_MAX_TABLE)]


class Params(C.Structure):
    _fields_ = [("eos", C.c_int), ("thermo", C.c_int), ("extrapolate", C.c_int),
                ("pressure_scale", C.c_double), ("temperature_scale", C.c_double),
                ("partial_pressure_scale", C.c_double), ("eos_w_temperature", C.c_double),
                ("relperm", Relperm), ("cappress", Cappress), ("gravity", C.c_double * 3)]


class KspOpts(C.Structure):
    _fields_ = [("type", C.c_int), ("restart", C.c_int), ("maxit", C.c_int),
                ("rtol", C.c_double), ("atol", C.c_double), ("dtol", C.c_double)]


class NewtonOpts(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("min_iterations", C.c_int),
                ("rel_tol", C.c_double), ("abs_tol", C.c_double),
                ("update_rel_tol", C.c_double), ("update_abs_tol", C.c_double),
                ("fd_err", C.c_double), ("fd_umin", C.c_double),
                ("pc_type", C.c_int), ("pc_nblocks", C.c_int), ("ksp", KspOpts)]


class NewtonResult(C.Structure):
    _fields_ = [("reason", C.c_int), ("iterations", C.c_int), ("linear_iterations", C.c_int),
                ("max_residual", C.c_double * 32), ("lin_its", C.c_int * 32),
                ("lin_reason", C.c_int * 32), ("lin_rnorm", C.c_double * 32)]


# every symbol include/waiwera_b200.h declares: name -> (restype, argtypes)
i, d, i64 = C.c_int, C.c_double, C.c_int64
SIGNATURES = {
    "wb_last_error": (C.c_char_p, []),
    "wb_version": (i, []),
    "wb_create": (i, [C.POINTER(Params), i, C.POINTER(vp)]),
    "wb_destroy": (i, [vp]),
    "wb_num_primary": (i, [vp]),
    "wb_fluid_dof": (i, [vp]),
    "wb_set_mesh": (i, [vp, i, i, i, i, vp, vp, vp, vp]),
    "wb_jacobian_pattern": (i, [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "wb_jacobian_get": (i, [vp, vp, vp, vp]),
    "wb_comm_unique_id": (i, [vp]),
    "wb_comm_init": (i, [vp, i, i, vp]),
    "wb_set_halo": (i, [vp, i, vp, vp, vp, vp, vp]),
    "wb_comm_p2p_blob_size": (i, []),
    "wb_comm_p2p_export": (i, [vp, vp]),
    "wb_comm_p2p_open": (i, [vp, vp]),
    "wb_comm_p2p_enabled": (i, [vp]),
    "wb_comm_p2p_disable": (i, [vp]),
    "wb_set_global_offset": (i, [vp, i64, i64]),
    "wb_fluid_init": (i, [vp, vp, vp]),
    "wb_set_boundary": (i, [vp, i, i, vp, i]),
    "wb_set_boundaries": (i, [vp, i, vp, vp, vp, vp]),
    "wb_set_rock": (i, [vp, vp]),
    "wb_set_sources": (i, [vp, i, vp, vp, vp, vp]),
    "wb_set_method": (i, [vp, i, d, vp]),
    "wb_set_source_components": (i, [vp, i, vp, vp]),
    "wb_set_source_controls": (i, [vp, i, vp, vp, vp, vp, vp]),
    "wb_set_source_recharge": (i, [vp, i, vp, vp, vp]),
    "wb_get_source_rates": (i, [vp, vp]),
    "wb_set_source_separators": (i, [vp, i, vp, vp, vp, vp, vp]),
    "wb_set_source_pressure_table": (i, [vp, i, vp, vp, vp, vp, vp]),
    "wb_separator_stage": (i, [vp, d, vp, vp]),
    "wb_get_source_separated": (i, [vp, vp]),
    "wb_get_fluid": (i, [vp, vp]),
    "wb_get_regions": (i, [vp, vp]),
    "wb_pre_iteration": (i, [vp]),
    "wb_pre_timestep": (i, [vp]),
    "wb_pre_retry_timestep": (i, [vp]),
    "wb_pre_eval": (i, [vp, vp, vp, i]),
    "wb_cell_balances": (i, [vp, vp]),
    "wb_cell_inflows": (i, [vp, vp]),
    "wb_residual_be": (i, [vp, vp, vp, d, vp, i, vp, vp, vp]),
    "wb_max_scaled": (i, [vp, vp, vp, d, C.POINTER(d), C.POINTER(i64)]),
    "wb_jacobian_be": (i, [vp, vp, vp, d, d, d, vp]),
    "wb_jacobian_be_colored": (i, [vp, vp, vp, d, d, d, vp, C.POINTER(i)]),
    "wb_fluid_transitions": (i, [vp, vp, vp, vp, C.POINTER(i), C.POINTER(i)]),
    "wb_mat_create": (i, [vp, i, i, i, i, vp, vp, vp, C.POINTER(vp)]),
    "wb_mat_set_values": (i, [vp, vp]),
    "wb_mat_get_values": (i, [vp, vp]),
    "wb_mat_destroy": (i, [vp]),
    "wb_jacobian_mat": (i, [vp, C.POINTER(vp)]),
    "wb_mat_mult": (i, [vp, vp, vp]),
    "wb_pc_setup": (i, [vp, i, i, vp, C.POINTER(vp)]),
    "wb_pc_refactor": (i, [vp]),
    "wb_pc_apply": (i, [vp, vp, vp]),
    "wb_pc_destroy": (i, [vp]),
    "wb_ksp_solve": (i, [vp, vp, C.POINTER(KspOpts), vp, vp, C.POINTER(i), C.POINTER(i), C.POINTER(d)]),
    "wb_ksp_set_check_every": (i, [i]),
    "wb_set_pc_blocks": (i, [vp, vp]),
    "wb_cell_faces_get": (i, [vp, C.POINTER(i), vp, vp, vp]),
    "wb_newton_solve_be": (i, [vp, C.POINTER(NewtonOpts), d, vp, vp, C.POINTER(NewtonResult)]),
    "wb_set_tracers": (i, [vp, i, vp, vp, vp, vp]),
    "wb_set_tracer_injection": (i, [vp, vp]),
    "wb_tracer_cell_balances": (i, [vp, vp]),
    "wb_tracer_setup_linear": (i, [vp, d, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]),
    "wb_tracer_solve": (i, [vp, C.POINTER(KspOpts), i, i, d, vp, vp, vp, vp, vp, vp, vp, C.POINTER(i), C.POINTER(i)]),
    "wb_ksp_set_fused": (i, [i]),
    "wb_ksp_set_fused_norm": (i, [i]),
    "wb_ksp_fused_profile": (i, [vp, C.POINTER(d), i]),
    "wb_timer_get": (i, [vp, C.c_char_p, C.POINTER(d), C.POINTER(i64)]),
    "wb_timer_reset": (i, [vp]),
    "wb_timers_enable": (i, [i]),
    "wb_launch_count": (i64, [vp]),
    "wb_stream": (vp, [vp]),
}

_lib = None


def lib():
    """Load the CUDA library; raises if it is missing (there is no fallback implementation)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "waiwera_b200: %s not found. Build it with `python -m waiwera_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class WbError(RuntimeError):
    pass


def check(rc, what=""):
    """rc < 0: fatal (raise); rc >= 0 returned to the caller (0 ok, >0 physics/domain error)."""
    if rc < 0:
        raise WbError("%s failed (%d): %s" % (what, rc, lib().wb_last_error().decode()))
    return rc
