"""ctypes binding of libpsb200.so (the C ABI in include/psb200.h).

There is deliberately no fallback: if the CUDA library is missing or no GPU is present the
product path raises (a CPU fallback would void every parity claim)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libpsb200.so")
_LIB = None

i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")

# every symbol include/psb200.h declares
SYMBOLS = [
    "psb200_create", "psb200_destroy", "psb200_set_parameters", "psb200_set_tolerance", "psb200_set_block_size",
    "psb200_analyze_pattern_csc", "psb200_factorize_csc", "psb200_solve", "psb200_solve_device", "psb200_get_info",
    "psb200_name", "psb200_last_error", "psb200_release_cached_memory", "psb200_factorize_csc_device", "psb200_residual_norm_device",
    "psb200_dirichlet_solve", "psb200_dirichlet_prefactorize", "psb200_dirichlet_solve_prefactorized", "psb200_dist_prepare", "psb200_dist_connect", "psb200_dist_reset", "psb200_dist_allgather", "psb200_residual_norm", "psb200_dist_local_range",
    "psb200_dist_plan_host", "psb200_dist_plan_host_aligned", "psb200_debug_get_csr",
    "psb200_spmv", "psb200_bench_spmv", "psb200_get_stream", "psb200_debug_set_aggregates", "psb200_debug_get_level",
    "psb200_precond_apply", "psb200_debug_get_aggregates",
    # include/psb200_io.h
    "psb200_market_load", "psb200_market_get_csc", "psb200_market_free", "psb200_market_save", "psb200_market_load_vector",
    "psb200_market_save_vector", "psb200_market_last_error",
    # include/psb200_nl.h
    "psb200_nl_create", "psb200_nl_destroy", "psb200_nl_minimize", "psb200_nl_get_info", "psb200_nl_last_error",
    "psb200_lbfgs_create", "psb200_lbfgs_destroy", "psb200_lbfgs_reset", "psb200_lbfgs_direction", "psb200_lbfgs_direction_device",
    "psb200_lbfgs_last_error", "psb200_nl_set_linear_solver_hook", "psb200_nl_set_iteration_callback", "psb200_nl_set_direction_filter",
    # include/psb200_problems.h
    "psb200_nh_create", "psb200_nh_destroy", "psb200_nh_pattern", "psb200_nh_value", "psb200_nh_gradient", "psb200_nh_hessian_device",
    "psb200_nh_hessian_host", "psb200_nh_last_error",
]


def build(verbose=False):
    """Compile libpsb200.so in-tree with nvcc for sm_100a (works without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j8", "libpsb200.so"]
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"psb200: {LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C polysolve_b200/csrc`). The CUDA backend has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    L.psb200_create.argtypes = [C.POINTER(H), C.c_char_p]
    L.psb200_destroy.argtypes = [H]
    L.psb200_set_parameters.argtypes = [H, C.c_char_p]
    L.psb200_set_tolerance.argtypes = [H, C.c_double]
    L.psb200_set_block_size.argtypes = [H, C.c_int]
    L.psb200_analyze_pattern_csc.argtypes = [H, C.c_int64, C.c_int64, i32p, i32p, C.c_int]
    L.psb200_factorize_csc.argtypes = [H, C.c_int64, C.c_int64, i32p, i32p, f64p]
    L.psb200_solve.argtypes = [H, f64p, f64p, C.c_int64]
    L.psb200_release_cached_memory.argtypes = [H]
    L.psb200_lbfgs_create.argtypes = [C.POINTER(C.c_void_p), C.c_int64, C.c_int, C.c_int]
    L.psb200_lbfgs_destroy.argtypes = [C.c_void_p]
    L.psb200_lbfgs_reset.argtypes = [C.c_void_p]
    L.psb200_lbfgs_direction.argtypes = [C.c_void_p, f64p, f64p, f64p, C.c_int64]
    L.psb200_lbfgs_direction_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.psb200_lbfgs_last_error.argtypes = [C.c_void_p]
    L.psb200_lbfgs_last_error.restype = C.c_char_p
    L.psb200_market_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.psb200_market_get_csc.argtypes = [C.c_void_p, i32p, i32p, f64p]
    L.psb200_market_free.argtypes = [C.c_void_p]
    L.psb200_market_save.argtypes = [C.c_char_p, C.c_int64, C.c_int64, i32p, i32p, f64p, C.c_int]
    L.psb200_market_load_vector.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    L.psb200_market_save_vector.argtypes = [C.c_char_p, f64p, C.c_int64]
    L.psb200_market_last_error.restype = C.c_char_p
    L.psb200_dirichlet_solve.argtypes = [H, C.c_int64, C.c_int64, i32p, i32p, f64p, f64p, i32p, C.c_int64, f64p, C.c_int]
    L.psb200_dirichlet_prefactorize.argtypes = [H, C.c_int64, C.c_int64, i32p, i32p, f64p, i32p, C.c_int64, C.c_int]
    L.psb200_dirichlet_solve_prefactorized.argtypes = [H, C.c_void_p, f64p, f64p, C.c_int64]
    L.psb200_factorize_csc_device.argtypes = [H, C.c_int64, C.c_int64, C.c_void_p, C.c_double]
    L.psb200_residual_norm_device.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_double)]
    L.psb200_solve_device.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int64]
    L.psb200_get_info.argtypes = [H, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.psb200_name.argtypes = [H]
    L.psb200_name.restype = C.c_char_p
    L.psb200_last_error.argtypes = [H]
    L.psb200_last_error.restype = C.c_char_p
    L.psb200_dist_prepare.argtypes = [H, C.c_int, C.c_int, C.c_int64, C.c_char_p]
    L.psb200_dist_connect.argtypes = [H, C.c_char_p]
    L.psb200_dist_reset.argtypes = [H]
    L.psb200_dist_allgather.argtypes = [H, f64p, C.c_int64]
    L.psb200_residual_norm.argtypes = [H, f64p, f64p, C.c_int64, C.POINTER(C.c_double)]
    L.psb200_dist_local_range.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
    L.psb200_dist_plan_host.argtypes = [C.c_int64, C.c_int64, i32p, i32p, C.c_int, C.c_int, C.c_int64, i64p, i64p,
                                        i32p, i32p, i32p, i32p, i32p, i32p, i32p]
    L.psb200_dist_plan_host_aligned.argtypes = [C.c_int64, C.c_int64, i32p, i32p, C.c_int, C.c_int, C.c_int64, C.c_int, i64p, i64p,
                                                i32p, i32p, i32p, i32p, i32p, i32p, i32p]
    L.psb200_debug_get_csr.argtypes = [H, i32p, i32p, i32p]
    L.psb200_spmv.argtypes = [H, f64p, f64p, C.c_int64]
    L.psb200_bench_spmv.argtypes = [H, C.c_char_p, C.c_int, C.POINTER(C.c_double)]
    L.psb200_get_stream.argtypes = [H]
    L.psb200_get_stream.restype = C.c_void_p
    L.psb200_debug_set_aggregates.argtypes = [H, C.c_int, C.c_void_p, C.c_int64]
    L.psb200_debug_get_level.argtypes = [H, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                         C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.c_void_p]
    L.psb200_precond_apply.argtypes = [H, f64p, f64p, C.c_int64]
    L.psb200_debug_get_aggregates.argtypes = [H, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    _LIB = L
    return L
