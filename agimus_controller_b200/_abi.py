"""ctypes mirror of ``include/agx.h`` (structs and constants only — no library is loaded here).

Shared by the product loader (``_lib.py``) and, in tests, by the oracle loader (``oracle/orc.py``).
"""
import ctypes as C

import numpy as np

AGX_MAX_NV = 16
AGX_MAX_CAPSULES = 4
AGX_MAX_COLLISION_PAIRS = 2
AGX_N_COST_TERMS = 13
AGX_N_COSTS = 5
AGX_STATUS_LINESEARCH = 4
AGX_JOINT_REVOLUTE = 0
AGX_JOINT_PRISMATIC = 1
AGX_POSE_PLACEMENT = 0
AGX_POSE_TRANSLATION_WORLD = 1

AGX_OK = 0
AGX_EINVAL = -1
AGX_EUNSUPPORTED = -2
AGX_ECUDA = -3
AGX_ENOMEM = -4

AGX_STATUS_CONVERGED = 0
AGX_STATUS_MAXITER = 1
AGX_STATUS_REGMAX = 2
AGX_STATUS_NAN = 3
AGX_STATUS_TIMEOUT = 5

_D = C.c_double
_I = C.c_int32


class AgxModel(C.Structure):
    """``struct agx_model`` — kinematic-tree table."""

    _fields_ = [
        ("nv", _I),
        ("frame_parent", _I),
        ("parent", _I * AGX_MAX_NV),
        ("jtype", _I * AGX_MAX_NV),
        ("axis", (_D * 3) * AGX_MAX_NV),
        ("placement_R", (_D * 9) * AGX_MAX_NV),
        ("placement_p", (_D * 3) * AGX_MAX_NV),
        ("mass", _D * AGX_MAX_NV),
        ("com", (_D * 3) * AGX_MAX_NV),
        ("inertia", (_D * 6) * AGX_MAX_NV),
        ("armature", _D * AGX_MAX_NV),
        ("gravity", _D * 3),
        ("frame_R", _D * 9),
        ("frame_p", _D * 3),
        ("cap_a0", (_D * 3) * AGX_MAX_CAPSULES),
        ("cap_a1", (_D * 3) * AGX_MAX_CAPSULES),
        ("cap_radius", _D * AGX_MAX_CAPSULES),
        ("col_alpha", _D),
        ("n_capsules", _I),
        ("cap_parent", _I * AGX_MAX_CAPSULES),
        ("n_pairs", _I),
        ("pair_a", _I * AGX_MAX_COLLISION_PAIRS),
        ("pair_b", _I * AGX_MAX_COLLISION_PAIRS),
        ("pose_mode", _I),
        ("reserved_", _I),
    ]


class AgxFddpOpts(C.Structure):
    """``struct agx_fddp_opts``."""

    _fields_ = [
        ("reg_min", _D),
        ("reg_max", _D),
        ("reg_incfactor", _D),
        ("reg_decfactor", _D),
        ("th_grad", _D),
        ("th_stepdec", _D),
        ("th_stepinc", _D),
        ("th_acceptstep", _D),
        ("th_acceptnegstep", _D),
        ("th_stop", _D),
        ("reg_init", _D),
        ("fixed_iters", _I),
        ("n_alphas", _I),
        ("eager_exit", _I),
        ("accept_rule", _I),
        ("max_solve_time", _D),
    ]


class AgxSqpOpts(C.Structure):
    """``struct agx_sqp_opts``."""

    _fields_ = [
        ("sigma", _D),
        ("reg", _D),
        ("mu", _D),
        ("termination_tolerance", _D),
        ("n_alphas", _I),
        ("eager_exit", _I),
        ("max_solve_time", _D),
    ]


def default_sqp_opts(termination_tolerance: float = 1e-3) -> AgxSqpOpts:
    """``mim_solvers.SolverCSQP`` as the reference configures it (ocp_base_croco.py:64-75, ocp_param_base.py:53-61):
    proximal sigma 1e-6, regularisation floor 1e-9, merit weight 10, KKT tolerance 1e-3, step lengths 2^-n, n < 10."""
    return AgxSqpOpts(sigma=1e-6, reg=1e-9, mu=10.0, termination_tolerance=termination_tolerance, n_alphas=10,
                      eager_exit=0, max_solve_time=0.0)


def ref_size(nv: int) -> int:
    """Doubles per node reference record: [xref nx][wx nx][uref nu][wu nu][Rref 9][pref 3][wpose 6][wcol 2]."""
    return 6 * nv + 20


def default_fddp_opts(fixed_iters: bool = False) -> AgxFddpOpts:
    """Crocoddyl ``SolverFDDP`` defaults."""
    return AgxFddpOpts(
        reg_min=1e-9,
        reg_max=1e9,
        reg_incfactor=10.0,
        reg_decfactor=10.0,
        th_grad=1e-12,
        th_stepdec=0.5,
        th_stepinc=0.01,
        th_acceptstep=0.1,
        th_acceptnegstep=2.0,
        th_stop=1e-9,
        reg_init=float("nan"),
        fixed_iters=1 if fixed_iters else 0,
        n_alphas=10,
        eager_exit=0,
        accept_rule=0,
        max_solve_time=0.0,
    )


def fill_array(dst, src) -> None:
    """Copy a (nested) numpy array into a (nested) ctypes array."""
    a = np.ascontiguousarray(src)
    C.memmove(dst, a.ctypes.data, min(C.sizeof(dst), a.nbytes))


_P = C.c_void_p


def bind(lib: C.CDLL) -> C.CDLL:
    """Declare the prototypes of every entry point of ``include/agx.h`` on a loaded library."""
    H = C.c_void_p
    lib.agx_ref_size.argtypes = [C.c_int]
    lib.agx_ref_size.restype = C.c_int
    lib.agx_fddp_opts_default.argtypes = [C.POINTER(AgxFddpOpts)]
    lib.agx_fddp_opts_default.restype = None
    lib.agx_create.argtypes = [C.POINTER(AgxModel), C.c_int, _P, C.c_int, C.c_int, C.c_int, C.POINTER(H)]
    lib.agx_create.restype = C.c_int
    lib.agx_destroy.argtypes = [H]
    lib.agx_destroy.restype = C.c_int
    lib.agx_last_error.argtypes = [H]
    lib.agx_last_error.restype = C.c_char_p
    lib.agx_set_refs.argtypes = [H, _P, _P]
    lib.agx_set_refs.restype = C.c_int
    lib.agx_set_refs_window.argtypes = [H, _P, C.c_int, C.c_int, _P, C.c_int, _P]
    lib.agx_set_refs_window.restype = C.c_int
    lib.agx_calc.argtypes = [H, _P, _P, _P, _P, _P]
    lib.agx_calc.restype = C.c_int
    lib.agx_calc_diff.argtypes = [H] + [_P] * 12
    lib.agx_calc_diff.restype = C.c_int
    lib.agx_rollout.argtypes = [H, _P, _P, _P, _P]
    lib.agx_rollout.restype = C.c_int
    lib.agx_integrate.argtypes = [H, _P, _P, C.c_double, C.c_int, _P, _P]
    lib.agx_integrate.restype = C.c_int
    lib.agx_rnea.argtypes = [H, _P, _P, _P, C.c_int, _P, _P]
    lib.agx_rnea.restype = C.c_int
    lib.agx_solve.argtypes = [H, _P, _P, _P, C.c_int, C.POINTER(AgxFddpOpts)] + [_P] * 9
    lib.agx_solve.restype = C.c_int
    lib.agx_set_capsule.argtypes = [H, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, _P]
    lib.agx_set_capsule.restype = C.c_int
    lib.agx_solve_sqp.argtypes = [H, _P, _P, _P, C.c_int, C.POINTER(AgxSqpOpts)] + [_P] * 9
    lib.agx_solve_sqp.restype = C.c_int
    lib.agx_sqp_opts_default.argtypes = [C.POINTER(AgxSqpOpts)]
    lib.agx_sqp_opts_default.restype = None
    lib.agx_cost_terms.argtypes = [H, _P, _P, _P, _P]
    lib.agx_cost_terms.restype = C.c_int
    lib.agx_cost_derivatives.argtypes = [H, _P, _P, _P, _P, _P]
    lib.agx_cost_derivatives.restype = C.c_int
    lib.agx_shift_warmstart.argtypes = [H, _P, _P, _P, _P, _P]
    lib.agx_shift_warmstart.restype = C.c_int
    lib.agx_riccati.argtypes = [H, _P, _P, _P, C.c_double, _P, _P, _P, _P]
    lib.agx_riccati.restype = C.c_int
    lib.agx_set_timing.argtypes = [H, C.c_int]
    lib.agx_set_timing.restype = C.c_int
    lib.agx_get_timing.argtypes = [H, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    lib.agx_get_timing.restype = C.c_int
    lib.agx_probe_fp64.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.agx_probe_fp64.restype = C.c_int
    lib.agx_launch_count.argtypes = [H]
    lib.agx_launch_count.restype = C.c_longlong
    return lib


EXPORTED_SYMBOLS = (
    "agx_ref_size", "agx_fddp_opts_default", "agx_create", "agx_destroy", "agx_last_error", "agx_set_refs",
    "agx_calc", "agx_calc_diff", "agx_rollout", "agx_integrate", "agx_rnea", "agx_solve", "agx_launch_count",
    "agx_set_timing", "agx_get_timing", "agx_probe_fp64", "agx_riccati", "agx_cost_terms", "agx_shift_warmstart", "agx_set_refs_window",
    "agx_solve_sqp", "agx_sqp_opts_default", "agx_set_capsule", "agx_cost_derivatives",
)
