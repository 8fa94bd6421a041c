"""ctypes view of `include/upright_b200.h` and loader of the CUDA library.

The library is the product: if `libupright_b200.so` is missing or no CUDA
device is present every entry point raises — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

UB_MAX_JOINTS = 9
UB_MAX_BODIES = 8
UB_MAX_CONTACTS = 32
UB_MAX_SPHERES = 16
UB_MAX_PAIRS = 32
UB_MAX_DYNAMIC_OBSTACLES = 4
UB_MAX_PROJECTILE_LINKS = 8
UB_MAX_NX = 27
UB_BODY_PARAMS = 10
UB_STATS = 8
UB_MAX_GATHER = 8
UB_IPC_HANDLE_BYTES = 64

UB_PTRS_DEVICE = 0x1
UB_WARM_START = 0x2
UB_COMPUTE_F64 = 0x4
UB_RESCUE_F64 = 0x8
UB_SHAPE_SPHERE, UB_SHAPE_HALFSPACE = 0, 1

STATUS_NAMES = {0: "converged", 1: "qp_maxiter", 2: "linesearch_failed", 3: "nan"}


class Joint(C.Structure):
    _fields_ = [("type", C.c_int32), ("reserved", C.c_int32), ("R", C.c_double * 9),
                ("p", C.c_double * 3), ("axis", C.c_double * 3)]


class Contact(C.Structure):
    _fields_ = [("body1", C.c_int32), ("body2", C.c_int32), ("mu", C.c_double),
                ("r_co_o1", C.c_double * 3), ("r_co_o2", C.c_double * 3),
                ("normal", C.c_double * 3), ("span", C.c_double * 6)]


class Sphere(C.Structure):
    _fields_ = [("link", C.c_int32), ("shape", C.c_int32), ("radius", C.c_double),
                ("offset", C.c_double * 3)]


class Pair(C.Structure):
    _fields_ = [("a", C.c_int32), ("b", C.c_int32)]


class SlackSettings(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("input_box", C.c_int32), ("state_box", C.c_int32),
                ("poly_ineq", C.c_int32), ("upper_L2_penalty", C.c_double),
                ("lower_L2_penalty", C.c_double)]


class ClosedLoopParams(C.Structure):
    """ub_closed_loop_params_t"""
    _fields_ = [("sim_dt", C.c_double), ("replan_period", C.c_double), ("n_steps", C.c_int32),
                ("log_stride", C.c_int32), ("use_feedback", C.c_int32), ("cold_start", C.c_int32),
                ("init_sqp_iteration", C.c_int32), ("sqp_iteration", C.c_int32),
                ("kp", C.c_double), ("kv", C.c_double), ("ka", C.c_double)]


UB_MAX_OBSTACLE_MODES = 8


class ObstacleMode(C.Structure):
    """ub_obstacle_mode_t"""
    _fields_ = [("time", C.c_double), ("position", C.c_double * 3), ("velocity", C.c_double * 3),
                ("acceleration", C.c_double * 3)]


class ProblemDesc(C.Structure):
    _fields_ = [
        ("nq", C.c_int32), ("nb", C.c_int32), ("nc", C.c_int32), ("nf", C.c_int32),
        ("N", C.c_int32), ("n_spheres", C.c_int32), ("n_pairs", C.c_int32),
        ("sqp_iteration", C.c_int32), ("qp_iter_max", C.c_int32),
        ("balancing_enabled", C.c_int32), ("obstacles_enabled", C.c_int32),
        ("qp_method", C.c_int32),
        ("dt", C.c_double),
        ("joints", Joint * UB_MAX_JOINTS),
        ("tool_R", C.c_double * 9), ("tool_p", C.c_double * 3),
        ("gravity", C.c_double * 3),
        ("state_weight", C.c_double * UB_MAX_NX),
        ("input_weight", C.c_double * UB_MAX_JOINTS),
        ("ee_weight", C.c_double * 6),
        ("force_weight", C.c_double),
        ("xd", C.c_double * UB_MAX_NX),
        ("state_lb", C.c_double * UB_MAX_NX), ("state_ub", C.c_double * UB_MAX_NX),
        ("input_lb", C.c_double * UB_MAX_JOINTS), ("input_ub", C.c_double * UB_MAX_JOINTS),
        ("force_lb", C.c_double), ("force_ub", C.c_double),
        ("body_params", (C.c_double * UB_BODY_PARAMS) * UB_MAX_BODIES),
        ("contacts", Contact * UB_MAX_CONTACTS),
        ("spheres", Sphere * UB_MAX_SPHERES),
        ("pairs", Pair * UB_MAX_PAIRS),
        ("minimum_distance", C.c_double),
        ("slacks", SlackSettings),
        ("rho_hard", C.c_double), ("qp_mu0", C.c_double), ("qp_thr0", C.c_double),
        ("qp_mu_target", C.c_double), ("qp_tol", C.c_double), ("reg_input", C.c_double),
        ("alpha_decay", C.c_double), ("alpha_min", C.c_double), ("g_max", C.c_double),
        ("g_min", C.c_double), ("gamma_c", C.c_double), ("armijo_factor", C.c_double),
        ("delta_tol", C.c_double), ("cost_tol", C.c_double),
        ("ee_box_enabled", C.c_int32), ("reserved0", C.c_int32),
        ("ee_box_lower", C.c_double * 3), ("ee_box_upper", C.c_double * 3),
        ("ia_cost_enabled", C.c_int32), ("reserved1", C.c_int32), ("ia_cost_weight", C.c_double),
        ("ia_span", C.c_double * 6),
        ("ia_constraint_enabled", C.c_int32), ("ia_use_angular_acceleration", C.c_int32),
        ("ia_align_with_fixed_vector", C.c_int32), ("reserved2", C.c_int32), ("ia_alpha", C.c_double),
        ("ia_normal", C.c_double * 3), ("ia_com", C.c_double * 3),
        ("n_dynamic_obstacles", C.c_int32), ("reserved3", C.c_int32),
        ("projectile_enabled", C.c_int32), ("n_projectile_links", C.c_int32),
        ("projectile_spheres", C.c_int32 * UB_MAX_PROJECTILE_LINKS),
        ("projectile_distances", C.c_double * UB_MAX_PROJECTILE_LINKS),
        ("projectile_scale", C.c_double), ("projectile_active", C.c_double),
    ]

    # convenience
    @property
    def nx(self):
        return 3 * self.nq

    @property
    def nu(self):
        return self.nq + self.nf * self.nc

    @property
    def n_eq(self):
        return 6 * self.nb if self.balancing_enabled else 0

    @property
    def n_fric(self):
        return 5 * self.nc if (self.balancing_enabled and self.nf == 3) else 0

    @property
    def n_obs(self):
        return self.n_pairs if self.obstacles_enabled else 0


LIB_NAME = "libupright_b200.so"
_lib = None


def library_path() -> Path:
    """The in-tree library; UB_LIBRARY names another build of it (diagnostic builds such as -DUB_DEBUG_NAN)."""
    override = os.environ.get("UB_LIBRARY")
    return Path(override).resolve() if override else Path(__file__).resolve().parent / LIB_NAME


def load_library():
    """Load the CUDA library; raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise RuntimeError(
            f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(upright_b200 has no CPU fallback)"
        )
    lib = C.CDLL(str(path), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else 0)
    vp = C.c_void_p
    lib.ub_last_error.restype = C.c_char_p
    lib.ub_version.restype = C.c_int
    lib.ub_problem_create.argtypes = [C.POINTER(ProblemDesc), C.POINTER(vp)]
    lib.ub_problem_create.restype = C.c_int
    lib.ub_problem_destroy.argtypes = [vp]
    lib.ub_problem_destroy.restype = None
    lib.ub_problem_dims.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.ub_problem_dims.restype = C.c_int
    lib.ub_workspace_bytes.argtypes = [vp, C.c_int32, C.c_uint32]
    lib.ub_workspace_bytes.restype = C.c_int64
    lib.ub_solve_batch.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                   C.c_int64, C.c_uint32, vp]
    lib.ub_solve_batch.restype = C.c_int
    lib.ub_eval.argtypes = [vp, C.c_char_p, C.c_int32, vp, vp, vp, vp, vp, C.c_int32,
                            C.POINTER(C.c_int32)]
    lib.ub_eval.restype = C.c_int
    lib.ub_closed_loop.argtypes = [vp, C.c_int32, vp, vp, vp, C.c_int32, vp, C.POINTER(ClosedLoopParams), vp, vp, vp,
                                   C.POINTER(C.c_int32), vp, C.c_uint32, vp]
    lib.ub_closed_loop.restype = C.c_int
    lib.ub_last_solve_ms.argtypes = [vp]
    lib.ub_last_solve_ms.restype = C.c_float
    lib.ub_launch_count.restype = C.c_int64
    lib.ub_closed_loop_set_obstacles.argtypes = [vp, C.c_int32, C.POINTER(C.c_int32), C.POINTER(ObstacleMode), C.c_int32, vp]
    lib.ub_closed_loop_set_obstacles.restype = C.c_int
    lib.ub_set_gather_targets.argtypes = [vp, C.c_int32, C.POINTER(vp), C.POINTER(vp), C.c_int64]
    lib.ub_set_gather_targets.restype = C.c_int
    lib.ub_gather_alloc.argtypes = [C.c_int64, C.POINTER(vp), C.c_char_p]
    lib.ub_gather_alloc.restype = C.c_int
    lib.ub_gather_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.ub_gather_open.restype = C.c_int
    lib.ub_gather_close.argtypes = [vp]
    lib.ub_gather_close.restype = C.c_int
    lib.ub_gather_free.argtypes = [vp]
    lib.ub_gather_free.restype = C.c_int
    lib.ub_measure_fma_peak.argtypes = [C.POINTER(C.c_double)]
    lib.ub_measure_fma_peak.restype = C.c_int
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "ub_last_error", "ub_version", "ub_problem_create", "ub_problem_destroy",
    "ub_problem_dims", "ub_workspace_bytes", "ub_solve_batch", "ub_eval",
    "ub_last_solve_ms", "ub_launch_count", "ub_set_option", "ub_workspace_layout", "ub_closed_loop",
    "ub_measure_fma_peak", "ub_set_gather_targets", "ub_closed_loop_set_obstacles",
    "ub_gather_alloc", "ub_gather_open", "ub_gather_close", "ub_gather_free",
]


def check(code: int):
    if code != 0:
        msg = load_library().ub_last_error()
        raise RuntimeError(f"upright_b200 error {code}: {msg.decode() if msg else '?'}")
