"""Host-side mirror of polysolve::nonlinear::Solver (Newton + line search) for the "CUDA" linear backend.

Same names and argument meaning as the reference (reference src/polysolve/nonlinear/Solver.hpp:37-66,
Problem.hpp:22-143); every call forwards to the C ABI of include/psb200_nl.h, whose driver
(polysolve_b200/csrc/newton.cpp) issues analyze_pattern -> factorize -> solve -> get_info on the
GPU linear solver exactly as Newton.cpp:189-211 does."""
import ctypes as C
import json

import numpy as np

from . import _lib


class Problem:
    """polysolve::nonlinear::Problem (reference Problem.hpp:22-143). Subclasses implement value, gradient and
    hessian (scipy.sparse CSC, the layout of StiffnessMatrix); the other hooks default as in the reference."""

    def value(self, x):
        raise NotImplementedError

    def gradient(self, x):
        raise NotImplementedError

    def hessian(self, x, project_to_psd=False):
        raise NotImplementedError

    def solution_changed(self, x):
        pass

    def is_step_valid(self, x0, x1):
        return True

    def max_step_size(self, x0, x1):
        return 1.0

    def line_search_begin(self, x0, x1):
        pass

    def line_search_end(self):
        pass

    def post_step(self, iteration, x, grad):
        pass

    def stop(self, x):
        return False

    # Optional, forwarded only when the problem object defines them (reference defaults otherwise, Problem.hpp:35,103-121):
    #   is_residual() -> bool
    #   after_line_search_custom_operation(x0, x1) -> bool      (True => solution_changed(x1))
    #   callback(state: dict of the solver's Criteria, x) -> bool   (False ends the loop)
    #   grad_norm(grad, norm_type) / step_norm(dx, norm_type) -> float     norm_type in "Euclidean" | "L2" | "Linf"
    #   grad_norm_rescaling(norm_type) / step_norm_rescaling(norm_type) / energy_norm_rescaling(norm_type) -> float


NORM_TYPES = ("Euclidean", "L2", "Linf")

_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_VALUE = C.CFUNCTYPE(C.c_double, C.c_void_p, _f64p, C.c_int64)
_GRAD = C.CFUNCTYPE(None, C.c_void_p, _f64p, C.c_int64, _f64p)
_HESS = C.CFUNCTYPE(C.c_int, C.c_void_p, _f64p, C.c_int64, C.c_int, C.POINTER(C.c_int64), C.POINTER(_i32p), C.POINTER(_i32p),
                    C.POINTER(_f64p))
_VOIDX = C.CFUNCTYPE(None, C.c_void_p, _f64p, C.c_int64)
_STEPV = C.CFUNCTYPE(C.c_int, C.c_void_p, _f64p, _f64p, C.c_int64)
_MAXST = C.CFUNCTYPE(C.c_double, C.c_void_p, _f64p, _f64p, C.c_int64)
_LSBEG = C.CFUNCTYPE(None, C.c_void_p, _f64p, _f64p, C.c_int64)
_LSEND = C.CFUNCTYPE(None, C.c_void_p)
_POST = C.CFUNCTYPE(None, C.c_void_p, C.c_int, _f64p, _f64p, C.c_int64)
_STOP = C.CFUNCTYPE(C.c_int, C.c_void_p, _f64p, C.c_int64)
_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)


class _CCriteria(C.Structure):
    _fields_ = [("iterations", C.c_int64), ("xDelta", C.c_double), ("fDelta", C.c_double), ("gradNorm", C.c_double),
                ("firstGradNorm", C.c_double), ("xDeltaDotGrad", C.c_double), ("relGradNorm", C.c_double),
                ("relXDelta", C.c_double), ("newtonDecrement", C.c_double), ("fDeltaCount", C.c_int64),
                ("energy", C.c_double), ("alpha", C.c_double), ("step", C.c_double)]


_ISRES = C.CFUNCTYPE(C.c_int, C.c_void_p)
_CALLB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(_CCriteria), _f64p, C.c_int64)
_NORM = C.CFUNCTYPE(C.c_double, C.c_void_p, _f64p, C.c_int64, C.c_int)
_RESC = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_int, C.c_int)
_ITCB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(_CCriteria))
_FILT = C.CFUNCTYPE(None, C.c_void_p, _f64p, _f64p, C.c_int64)


class _CProblem(C.Structure):
    _fields_ = [("user", C.c_void_p), ("value", _VALUE), ("gradient", _GRAD), ("hessian", _HESS),
                ("solution_changed", _VOIDX), ("is_step_valid", _STEPV), ("max_step_size", _MAXST),
                ("line_search_begin", _LSBEG), ("line_search_end", _LSEND), ("post_step", _POST), ("stop", _STOP),
                ("hessian_device", _HESS), ("is_residual", _ISRES), ("after_line_search_custom_operation", _STEPV),
                ("callback", _CALLB), ("grad_norm", _NORM), ("step_norm", _NORM), ("norm_rescaling", _RESC)]


def _vec(p, n):
    return np.ctypeslib.as_array(p, shape=(n,))


class NonlinearSolver:
    """polysolve::nonlinear::Solver. create(solver_params, linear_solver_params) mirrors Solver.cpp:124-186."""

    @staticmethod
    def create(solver_params=None, linear_solver_params=None):
        return NonlinearSolver(solver_params or {}, linear_solver_params or {"solver": "CUDA"})

    def __init__(self, solver_params, linear_solver_params):
        self._L = _lib.lib()
        L = self._L
        L.psb200_nl_create.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_char_p]
        L.psb200_nl_destroy.argtypes = [C.c_void_p]
        L.psb200_nl_minimize.argtypes = [C.c_void_p, C.POINTER(_CProblem), np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS"), C.c_int64]
        L.psb200_nl_get_info.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.psb200_nl_last_error.argtypes = [C.c_void_p]
        L.psb200_nl_last_error.restype = C.c_char_p
        self._h = C.c_void_p()
        rc = L.psb200_nl_create(C.byref(self._h), json.dumps(solver_params).encode(), json.dumps(linear_solver_params).encode())
        if rc:
            raise RuntimeError(L.psb200_nl_last_error(None).decode())

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.psb200_nl_destroy(h)
            self._h = None

    def set_linear_solver_hook(self, fn):
        """fn(solver) is called once for every linear solver the driver owns (a borrowed polysolve_b200.Solver), in creation
        order: a multi-GPU application connects them there (solver.dist_setup_torch())."""
        from .solver import Solver
        self._L.psb200_nl_set_linear_solver_hook.argtypes = [C.c_void_p, _HOOK, C.c_void_p]
        errors = []

        def hook(_, lin):
            try:
                fn(Solver(_borrowed=lin))
            except Exception as e:  # noqa: BLE001
                errors.append(e)
        cb = _HOOK(hook)
        rc = self._L.psb200_nl_set_linear_solver_hook(self._h, cb, None)
        if errors:
            raise errors[0]
        if rc:
            raise RuntimeError("psb200_nl_set_linear_solver_hook failed")

    def set_iteration_callback(self, fn):
        """Solver::set_iteration_callback: fn(state: dict of the solver's Criteria) -> bool, True stops the solve with status
        "Objective function specified to stop" (no error). None removes it."""
        self._L.psb200_nl_set_iteration_callback.argtypes = [C.c_void_p, _ITCB, C.c_void_p]
        self._it_errors = []

        def cb(_, st):
            try:
                return 1 if fn({k: getattr(st.contents, k) for k, _t in _CCriteria._fields_}) else 0
            except Exception as e:  # noqa: BLE001 -- exceptions must not cross the C frame
                self._it_errors.append(e)
                return 1
        self._it_cb = _ITCB(cb) if fn is not None else C.cast(None, _ITCB)   # kept alive with the solver
        if self._L.psb200_nl_set_iteration_callback(self._h, self._it_cb, None):
            raise RuntimeError("psb200_nl_set_iteration_callback failed")

    def set_direction_filter(self, fn):
        """Solver::set_direction_filter: fn(x, dx) edits dx in place (numpy views). None removes it."""
        self._L.psb200_nl_set_direction_filter.argtypes = [C.c_void_p, _FILT, C.c_void_p]
        self._it_errors = getattr(self, "_it_errors", [])

        def cb(_, xp, dp, nn):
            try:
                fn(_vec(xp, nn), _vec(dp, nn))
            except Exception as e:  # noqa: BLE001
                self._it_errors.append(e)
        self._filt_cb = _FILT(cb) if fn is not None else C.cast(None, _FILT)
        if self._L.psb200_nl_set_direction_filter(self._h, self._filt_cb, None):
            raise RuntimeError("psb200_nl_set_direction_filter failed")

    def minimize(self, problem, x):
        """x is in/out (float64, contiguous). Raises RuntimeError where the reference throws."""
        if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous):
            raise TypeError("x must be a contiguous float64 numpy array (it is updated in place)")
        n = x.shape[0]
        keep = {}
        errors = []

        def guard(default):
            def deco(fn):
                def wrapped(*a):
                    try:
                        return fn(*a)
                    except Exception as e:  # noqa: BLE001 -- exceptions must not cross the C frame
                        errors.append(e)
                        return default
                return wrapped
            return deco

        @guard(float("nan"))
        def value(_, xp, nn):
            return float(problem.value(_vec(xp, nn)))

        @guard(None)
        def gradient(_, xp, nn, gp):
            _vec(gp, nn)[:] = problem.gradient(_vec(xp, nn))

        @guard(1)
        def hessian(_, xp, nn, psd, nnz, outer, inner, vals):
            import scipy.sparse as sp
            H = problem.hessian(_vec(xp, nn), bool(psd))
            if not sp.isspmatrix_csc(H):
                H = sp.csc_matrix(H)
            H.sort_indices()
            o = np.ascontiguousarray(H.indptr, np.int32)
            i = np.ascontiguousarray(H.indices, np.int32)
            v = np.ascontiguousarray(H.data, np.float64)
            keep["h"] = (o, i, v)  # valid until the next hessian() call
            nnz[0] = int(o[-1])
            outer[0] = o.ctypes.data_as(_i32p)
            inner[0] = i.ctypes.data_as(_i32p)
            vals[0] = v.ctypes.data_as(_f64p)
            return 0

        @guard(1)
        def hessian_device(_, xp, nn, psd, nnz, outer, inner, vals):
            # the Problem assembles on the GPU: (outer, inner) host int32 arrays of the fixed pattern, device pointer of the values
            o, i, dptr = problem.hessian_device(_vec(xp, nn), bool(psd))
            keep["hd"] = (o, i)
            nnz[0] = int(o[-1])
            outer[0] = o.ctypes.data_as(_i32p)
            inner[0] = i.ctypes.data_as(_i32p)
            vals[0] = C.cast(dptr, _f64p)
            return 0

        @guard(None)
        def solution_changed(_, xp, nn):
            problem.solution_changed(_vec(xp, nn))

        @guard(0)
        def is_step_valid(_, x0, x1, nn):
            return 1 if problem.is_step_valid(_vec(x0, nn), _vec(x1, nn)) else 0

        @guard(0.0)
        def max_step_size(_, x0, x1, nn):
            return float(problem.max_step_size(_vec(x0, nn), _vec(x1, nn)))

        @guard(None)
        def ls_begin(_, x0, x1, nn):
            problem.line_search_begin(_vec(x0, nn), _vec(x1, nn))

        @guard(None)
        def ls_end(_):
            problem.line_search_end()

        @guard(None)
        def post_step(_, it, xp, gp, nn):
            problem.post_step(it, _vec(xp, nn), _vec(gp, nn))

        @guard(1)
        def stop(_, xp, nn):
            return 1 if problem.stop(_vec(xp, nn)) else 0

        @guard(0)
        def is_residual(_):
            return 1 if problem.is_residual() else 0

        @guard(0)
        def after_ls(_, x0, x1, nn):
            return 1 if problem.after_line_search_custom_operation(_vec(x0, nn), _vec(x1, nn)) else 0

        @guard(0)
        def callback(_, st, xp, nn):
            state = {k: getattr(st.contents, k) for k, _t in _CCriteria._fields_}
            return 1 if problem.callback(state, _vec(xp, nn)) else 0

        @guard(float("nan"))
        def grad_norm(_, gp, nn, nt):
            return float(problem.grad_norm(_vec(gp, nn), NORM_TYPES[nt]))

        @guard(float("nan"))
        def step_norm(_, dp, nn, nt):
            return float(problem.step_norm(_vec(dp, nn), NORM_TYPES[nt]))

        @guard(1.0)
        def rescaling(_, which, nt):
            fn = getattr(problem, ("grad_norm_rescaling", "step_norm_rescaling", "energy_norm_rescaling")[which], None)
            return float(fn(NORM_TYPES[nt])) if fn else 1.0

        def opt(name, ctype, fn):
            return ctype(fn) if hasattr(problem, name) else C.cast(None, ctype)

        has_resc = any(hasattr(problem, k) for k in ("grad_norm_rescaling", "step_norm_rescaling", "energy_norm_rescaling"))
        cp = _CProblem(None, _VALUE(value), _GRAD(gradient), _HESS(hessian), _VOIDX(solution_changed), _STEPV(is_step_valid),
                       _MAXST(max_step_size), _LSBEG(ls_begin), _LSEND(ls_end), _POST(post_step), _STOP(stop),
                       opt("hessian_device", _HESS, hessian_device), opt("is_residual", _ISRES, is_residual),
                       opt("after_line_search_custom_operation", _STEPV, after_ls), opt("callback", _CALLB, callback),
                       opt("grad_norm", _NORM, grad_norm), opt("step_norm", _NORM, step_norm),
                       _RESC(rescaling) if has_resc else C.cast(None, _RESC))
        rc = self._L.psb200_nl_minimize(self._h, C.byref(cp), x, n)
        errors.extend(getattr(self, "_it_errors", []))
        if getattr(self, "_it_errors", None):
            self._it_errors.clear()
        if errors:
            raise errors[0]
        if rc:
            raise RuntimeError(self._L.psb200_nl_last_error(self._h).decode())
        return x

    def get_info(self):
        need = C.c_size_t()
        buf = C.create_string_buffer(1 << 16)
        rc = self._L.psb200_nl_get_info(self._h, buf, len(buf), C.byref(need))
        if rc and need.value > len(buf):
            buf = C.create_string_buffer(need.value)
            rc = self._L.psb200_nl_get_info(self._h, buf, len(buf), C.byref(need))
        if rc:
            raise RuntimeError("psb200_nl_get_info failed")
        return json.loads(buf.value.decode())


class Lbfgs:
    """The L-BFGS memory of the reference's LBFGS strategy (LBFGS.cpp:22-61, LBFGSpp::BFGSMat) on device vectors
    (include/psb200_nl.h, polysolve_b200/csrc/lbfgs.cu)."""

    def __init__(self, n, history_size=6, device=-1):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        self.n = int(n)
        if self._L.psb200_lbfgs_create(C.byref(self._h), self.n, int(history_size), int(device)):
            raise RuntimeError(self._L.psb200_lbfgs_last_error(None).decode())

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.psb200_lbfgs_destroy(h)
            self._h = None

    def _check(self, rc):
        if rc:
            raise RuntimeError(self._L.psb200_lbfgs_last_error(self._h).decode())

    def reset(self):
        self._check(self._L.psb200_lbfgs_reset(self._h))

    def compute_update_direction(self, x, grad):
        x = np.ascontiguousarray(x, np.float64)
        grad = np.ascontiguousarray(grad, np.float64)
        if x.shape != grad.shape:
            raise RuntimeError("psb200_lbfgs_direction: x and grad differ in size")
        d = np.empty(max(x.shape[0], 1), np.float64)
        self._check(self._L.psb200_lbfgs_direction(self._h, x if x.size else d, grad if grad.size else d, d, x.shape[0]))
        return d[:x.shape[0]]

    def compute_update_direction_device(self, x_ptr, grad_ptr, dir_ptr):
        self._check(self._L.psb200_lbfgs_direction_device(self._h, x_ptr, grad_ptr, dir_ptr, self.n))
