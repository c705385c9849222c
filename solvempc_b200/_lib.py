"""ctypes loader of libsolvempc_b200.so (CUDA kernels + C ABI, include/solvempc_b200.h).

There is no CPU fallback: if the library has not been built (``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C solvempc_b200/csrc``) importing the product raises, and without a CUDA device every compute call
raises ``SolveMpcError`` (SMPC_ERR_CUDA).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SOLVEMPC_B200_LIB") or os.path.join(_HERE, "libsolvempc_b200.so")   # (the override is for A/B builds in development)

HOST, DEVICE = 0, 1
OK, ERR_ARG, ERR_DATA, ERR_CUDA, ERR_STATE, ERR_IO = 0, 1, 2, 3, 4, 5
SOLVED, SOLVED_INACCURATE, MAX_ITER_REACHED = 1, 2, -2
PRIMAL_INFEASIBLE, DUAL_INFEASIBLE, UNSOLVED = -3, -4, -10
PRIMAL_INFEASIBLE_INACCURATE, DUAL_INFEASIBLE_INACCURATE = 3, 4


class SolveMpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"solvempc_b200 error {code}: {msg}")
        self.code = code


class Settings(C.Structure):
    """smpc_settings (OSQP settings; reference sets only verbosity + warm start, cpp:51-52)."""
    _fields_ = [
        ("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
        ("eps_abs", C.c_double), ("eps_rel", C.c_double),
        ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
        ("adaptive_rho_tolerance", C.c_double),
        ("max_iter", C.c_int), ("check_termination", C.c_int), ("scaling", C.c_int),
        ("adaptive_rho", C.c_int), ("adaptive_rho_interval", C.c_int),
        ("warm_start", C.c_int), ("scaled_termination", C.c_int), ("kernel", C.c_int),
    ]


class MpcConfig(C.Structure):
    _fields_ = [
        ("horizon", C.c_int), ("nx", C.c_int), ("n_state_rows", C.c_int),
        ("Q", C.c_double), ("R", C.c_double), ("RD", C.c_double), ("u_limit", C.c_double), ("xref", C.c_double),
        ("Ad", C.c_void_p), ("Bd", C.c_void_p), ("Cd", C.c_void_p), ("K", C.c_void_p),
        ("per_instance", C.c_int),
    ]


class MimoConfig(C.Structure):
    _fields_ = [
        ("horizon", C.c_int), ("nx", C.c_int), ("nu", C.c_int),
        ("Ad", C.c_void_p), ("Bd", C.c_void_p), ("Q", C.c_void_p), ("R", C.c_void_p), ("umin", C.c_void_p), ("umax", C.c_void_p),
    ]


# every symbol include/solvempc_b200.h declares (tests/test_cabi.py checks the header against this list)
_vp, _i, _dp = C.c_void_p, C.c_int, C.c_void_p
SIGNATURES = {
    "smpc_default_settings": (None, [C.POINTER(Settings)]),
    "smpc_last_error": (C.c_char_p, []),
    "smpc_version": (C.c_char_p, []),
    "smpc_device_count": (_i, []),
    "smpc_solver_create_shared": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _dp, _dp, _dp, _dp, _dp, C.POINTER(Settings)]),
    "smpc_solver_create_shared_csc": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _vp, _vp, _dp, _vp, _vp, _dp, _dp, _dp, _dp, C.POINTER(Settings)]),
    "smpc_solver_create_batched": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _dp, _dp, _i, _dp, _dp, C.POINTER(Settings)]),
    "smpc_solver_destroy": (_i, [_vp]),
    "smpc_solver_set_stream": (_i, [_vp, _vp]),
    "smpc_solver_dims": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "smpc_solver_update_lin_cost": (_i, [_vp, _dp, _i]),
    "smpc_solver_update_upper_bound": (_i, [_vp, _dp, _i]),
    "smpc_solver_update_lower_bound": (_i, [_vp, _dp, _i]),
    "smpc_solver_update_bounds": (_i, [_vp, _dp, _dp, _i]),
    "smpc_solver_warm_start": (_i, [_vp, _dp, _dp, _i]),
    "smpc_solver_cold_start": (_i, [_vp]),
    "smpc_solver_reset": (_i, [_vp]),
    "smpc_solver_set_cold_solves": (_i, [_vp, _i]),
    "smpc_solver_set_scheduling": (_i, [_vp, _i]),
    "smpc_solver_enable_timing": (_i, [_vp, _i]),
    "smpc_solver_kernel_ms": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(_i), _i]),
    "smpc_solver_set_polish": (_i, [_vp, _i, C.c_double, _i]),
    "smpc_solver_get_polish_status": (_i, [_vp, _vp, _i]),
    "smpc_solver_solve": (_i, [_vp]),
    "smpc_solver_get_solution": (_i, [_vp, _dp, _dp, _i]),
    "smpc_solver_get_info": (_i, [_vp, _vp, _vp, _dp, _dp, _dp, _dp, _vp, _i]),
    "smpc_solver_count_solved": (_i, [_vp, C.POINTER(_i)]),
    "smpc_solver_sync": (_i, [_vp]),
    "smpc_solver_get_scaling": (_i, [_vp, _dp, _dp, C.POINTER(C.c_double)]),
    "smpc_solver_launch_count": (C.c_longlong, [_vp]),
    "smpc_solver_kernel_name": (C.c_char_p, [_vp]),
    "smpc_solver_row_pairs": (_i, [_vp]),
    "smpc_shared_plan_inspect": (_i, [_i, _i, _dp, _dp, _dp, _dp, _dp, C.POINTER(Settings), _dp, _dp, C.POINTER(C.c_double),
                                      _dp, _dp, _dp, _dp, _dp, _dp, _vp]),
    "smpc_mpc_create": (_i, [C.POINTER(_vp), _i, C.POINTER(MpcConfig), _i, C.POINTER(Settings)]),
    "smpc_mpc_create_from_json": (_i, [C.POINTER(_vp), _i, C.c_char_p, _i, C.POINTER(Settings)]),
    "smpc_mpc_destroy": (_i, [_vp]),
    "smpc_mpc_set_stream": (_i, [_vp, _vp]),
    "smpc_mpc_dims": (_i, [_vp] + [C.POINTER(_i)] * 5),
    "smpc_mpc_solver": (_vp, [_vp]),
    "smpc_mpc_get_matrix": (_i, [_vp, C.c_char_p, _i, _dp, _i]),
    "smpc_mpc_set_state": (_i, [_vp, _dp, _dp, _dp, _i]),
    "smpc_mpc_controller_step": (_i, [_vp]),
    "smpc_mpc_controller_step_from": (_i, [_vp, _dp, _dp, _dp, _i]),
    "smpc_mpc_bind_results": (_i, [_vp, _vp, _vp, _i]),
    "smpc_mpc_sync": (_i, [_vp]),
    "smpc_mpc_plant_step": (_i, [_vp]),
    "smpc_mpc_closed_loop": (_i, [_vp, _i, C.c_double, _i, _vp, _i, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "smpc_mpc_get_state": (_i, [_vp, _dp, _dp, _i]),
    "smpc_mpc_get_control_status": (_i, [_vp, _dp, _vp, _i]),
    "smpc_mpc_get_step_vectors": (_i, [_vp, _dp, _dp, _i]),
    "smpc_mpc_launch_count": (C.c_longlong, [_vp]),
    "smpc_mimo_create": (_i, [C.POINTER(_vp), _i, C.POINTER(MimoConfig), _i, C.POINTER(Settings)]),
    "smpc_mimo_create_from_json": (_i, [C.POINTER(_vp), _i, C.c_char_p, _i, C.POINTER(Settings)]),
    "smpc_mimo_destroy": (_i, [_vp]),
    "smpc_mimo_set_stream": (_i, [_vp, _vp]),
    "smpc_mimo_dims": (_i, [_vp] + [C.POINTER(_i)] * 6),
    "smpc_mimo_solver": (_vp, [_vp]),
    "smpc_mimo_get_matrix": (_i, [_vp, C.c_char_p, _dp, _i]),
    "smpc_mimo_set_state": (_i, [_vp, _dp, _dp, _i]),
    "smpc_mimo_controller_step": (_i, [_vp]),
    "smpc_mimo_get_control": (_i, [_vp, _dp, _i]),
    "smpc_mimo_launch_count": (C.c_longlong, [_vp]),
    "smpc_wire_parse_frame": (_i, [C.c_char_p, _i, C.POINTER(C.c_double), _dp]),
    "smpc_wire_format_control": (_i, [C.c_double, C.c_char_p, _i, _i]),
    "smpc_feed_open": (_i, [C.POINTER(_vp), _i]),
    "smpc_feed_latest": (_i, [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_double), _dp]),
    "smpc_feed_stats": (_i, [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "smpc_feed_close": (_i, [_vp]),
}

_lib = None


def lib():
    """The loaded library; raises if it was not built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first (__graft_entry__.build() or "
                "`make -C solvempc_b200/csrc`). solvempc_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise SolveMpcError(rc, lib().smpc_last_error().decode("utf-8", "replace"))


def default_settings(**kw):
    s = Settings()
    lib().smpc_default_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise KeyError(k)
        setattr(s, k, v)
    return s
