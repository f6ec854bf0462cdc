"""ctypes binding of libdpilqr_b200.so (the C ABI declared in include/dpilqr_b200.h).

There is deliberately no fallback: if the CUDA library is missing or no CUDA device is
visible, every compute entry point raises.  The library is built in-tree by
``python -m dpilqr_b200.build`` (or ``__graft_entry__.build()``).
"""

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DPILQR_B200_LIB selects an experimental build (dpilqr_b200.build.build_variant); the default is the in-tree library
LIB_PATH = os.environ.get("DPILQR_B200_LIB") or os.path.join(_HERE, "lib", "libdpilqr_b200.so")

c_double_p = ctypes.c_void_p  # raw device/host addresses are passed as integers
c_int32_p = ctypes.c_void_p


class NativeError(RuntimeError):
    """An entry point of libdpilqr_b200.so returned an error code."""


class BatchStruct(ctypes.Structure):
    """Mirror of ``dpilqr_batch`` (include/dpilqr_b200.h)."""

    _fields_ = [
        ("n_problems", ctypes.c_int32),
        ("n_agents", ctypes.c_int32),
        ("s", ctypes.c_int32),
        ("c", ctypes.c_int32),
        ("horizon", ctypes.c_int32),
        ("n_cost", ctypes.c_int32),
        ("dt", ctypes.c_double),
        ("model", ctypes.c_void_p),
        ("n_dims", ctypes.c_void_p),
        ("cost_idx", ctypes.c_void_p),
        ("Q", ctypes.c_void_p),
        ("R", ctypes.c_void_p),
        ("Qf", ctypes.c_void_p),
        ("xf", ctypes.c_void_p),
        ("radius", ctypes.c_void_p),
        ("weights", ctypes.c_void_p),
        ("has_prox", ctypes.c_void_p),
        ("model_hint", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class SolveOpts(ctypes.Structure):
    """Mirror of ``dpilqr_solve_opts``."""

    _fields_ = [
        ("n_lqr_iter", ctypes.c_int32),
        ("n_alpha", ctypes.c_int32),
        ("tol", ctypes.c_double),
        ("t_kill", ctypes.c_double),
        ("record_trace", ctypes.c_int32),
        ("profile", ctypes.c_int32),
        ("bounded_search", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


KERNEL_KINDS = ["rollout", "linquad", "backward", "linesearch", "select", "backward_full"]


class Profile(ctypes.Structure):
    """Mirror of ``dpilqr_profile``."""

    _fields_ = [("ms", ctypes.c_double * 6), ("launches", ctypes.c_int64 * 6), ("units", ctypes.c_int64 * 6)]


# status bits (include/dpilqr_b200.h)
ST_NONFINITE = 1
ST_SINGULAR = 2
ST_POINT_NDIM = 4
ST_CONVERGED = 16
ST_LS_FAILED = 32
ST_ITER_LIMIT = 64
ST_TIME_LIMIT = 128
J_ABORTED = 1.7976931348623157e308  # DPILQR_J_ABORTED: candidate stopped by the bounded line search

EXPORTS = [
    "dpilqr_last_error", "dpilqr_version", "dpilqr_device_count", "dpilqr_model_nx", "dpilqr_model_nu",
    "dpilqr_stage_stride", "dpilqr_workspace_bytes", "dpilqr_f", "dpilqr_integrate", "dpilqr_linearize",
    "dpilqr_rollout_linesearch", "dpilqr_linearize_quadraticize", "dpilqr_game_cost", "dpilqr_stage_to_dense",
    "dpilqr_backward", "dpilqr_inter_graph", "dpilqr_solve_batch", "dpilqr_solve_batch_host", "dpilqr_get_profile", "dpilqr_debug_backward_timing",
    "dpilqr_release_cache", "dpilqr_random_setup",
]

_lib = None


def lib():
    """Load the shared library (once).  Raises NativeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} not found: build it with `python -m dpilqr_b200.build` "
            "(dpilqr_b200 has no CPU fallback)"
        )
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
    bp, op = ctypes.POINTER(BatchStruct), ctypes.POINTER(SolveOpts)
    L.dpilqr_last_error.restype = ctypes.c_char_p
    L.dpilqr_last_error.argtypes = []
    L.dpilqr_version.restype = i32
    L.dpilqr_device_count.restype = i32
    L.dpilqr_model_nx.argtypes = [i32]
    L.dpilqr_model_nu.argtypes = [i32]
    L.dpilqr_stage_stride.restype = i64
    L.dpilqr_stage_stride.argtypes = [i32, i32, i32]
    L.dpilqr_workspace_bytes.restype = i64
    L.dpilqr_workspace_bytes.argtypes = [i32, i32, i32, i32, i32, i32]
    L.dpilqr_f.argtypes = [i32, i64, vp, vp, vp, vp]
    L.dpilqr_integrate.argtypes = [i32, dbl, i64, vp, vp, vp, vp]
    L.dpilqr_linearize.argtypes = [i32, dbl, i64, vp, vp, vp, vp, vp]
    L.dpilqr_rollout_linesearch.argtypes = [bp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp]
    L.dpilqr_linearize_quadraticize.argtypes = [bp, vp, vp, vp, vp, vp]
    L.dpilqr_game_cost.argtypes = [bp, i64, vp, vp, i32, vp, vp]
    L.dpilqr_stage_to_dense.argtypes = [bp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.dpilqr_backward.argtypes = [bp, vp, vp, vp, vp, vp, vp]
    L.dpilqr_inter_graph.argtypes = [vp, i64, i32, i32, i32, vp, vp, vp]
    L.dpilqr_solve_batch.restype = i64
    L.dpilqr_solve_batch.argtypes = [bp, op, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]
    L.dpilqr_solve_batch_host.restype = i64
    L.dpilqr_solve_batch_host.argtypes = [bp, op, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32]
    L.dpilqr_get_profile.argtypes = [ctypes.POINTER(Profile), i32]
    L.dpilqr_debug_backward_timing.argtypes = [vp]
    L.dpilqr_release_cache.restype = i32
    L.dpilqr_random_setup.argtypes = [i64, i64, i32, i32, i32, dbl, dbl, vp, vp, vp]
    _lib = L
    return L


def last_error():
    return lib().dpilqr_last_error().decode("utf-8", "replace")


def check(rc):
    """Raise NativeError for a negative return code; pass non-negative values through."""
    if rc < 0:
        raise NativeError(f"libdpilqr_b200 error {rc}: {last_error()}")
    return rc


def get_profile(reset=False):
    """Per-kernel device time accumulated by solves run with profile=True: {kind: (ms, launches, units)}."""
    prof = Profile()
    lib().dpilqr_get_profile(ctypes.byref(prof), int(bool(reset)))
    return {k: (prof.ms[i], int(prof.launches[i]), int(prof.units[i])) for i, k in enumerate(KERNEL_KINDS)}


def require_device():
    """Fail loudly when no CUDA device is visible (no CPU path exists)."""
    return check(lib().dpilqr_device_count())
