"""Device-resident Neo-Hookean Problem (include/psb200_problems.h, polysolve_b200/csrc/neohookean.cu): the `Problem
subclass` of BASELINE config 5 with the interface of polysolve::nonlinear::Problem (reference
src/polysolve/nonlinear/Problem.hpp:49-67). value / gradient take and return host vectors; hessian_device leaves the
values in GPU memory for psb200_factorize_csc_device (the Newton driver uses it when present)."""
import ctypes as C

import numpy as np

from . import _lib
from .nonlinear import Problem

_i32pp = C.POINTER(C.POINTER(C.c_int32))
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def grid_tets(m):
    """Kuhn 6-tet split of an m^3-node unit-spacing grid (node id = i + m j + m^2 k), positively oriented."""
    import itertools
    idx = np.arange(m ** 3).reshape(m, m, m)  # [k, j, i]
    kk, jj, ii = np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij")
    X = np.stack([ii.reshape(-1), jj.reshape(-1), kk.reshape(-1)], 1).astype(np.float64)
    tets = []
    for perm in itertools.permutations(range(3)):
        v = [np.zeros(3, int)]
        for ax in perm:
            w = v[-1].copy()
            w[ax] += 1
            v.append(w)
        cols = [idx[tuple(slice(w[ax], w[ax] + m - 1) for ax in (2, 1, 0))].reshape(-1) for w in v]
        tets.append(np.stack(cols, 1))
    T = np.concatenate(tets, 0).astype(np.int32)
    d = X[T[:, 1:]] - X[T[:, :1]]
    neg = np.linalg.det(d) < 0
    T[neg, 2], T[neg, 3] = T[neg, 3].copy(), T[neg, 2].copy()
    return X, np.ascontiguousarray(T)


class NeoHookeanDevice(Problem):
    def __init__(self, X, T, mu=1.0, lam=1.5, fixed=None, device=-1):
        L = _lib.lib()
        self._L = L
        L.psb200_nh_create.argtypes = [C.POINTER(C.c_void_p), C.c_int64, _f64p, C.c_int64, np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS"),
                                       C.c_double, C.c_double, C.c_void_p, C.c_int]
        L.psb200_nh_destroy.argtypes = [C.c_void_p]
        L.psb200_nh_pattern.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _i32pp, _i32pp]
        L.psb200_nh_value.argtypes = [C.c_void_p, _f64p, C.POINTER(C.c_double)]
        L.psb200_nh_gradient.argtypes = [C.c_void_p, _f64p, _f64p]
        L.psb200_nh_hessian_device.argtypes = [C.c_void_p, _f64p, C.POINTER(C.c_void_p)]
        L.psb200_nh_hessian_host.argtypes = [C.c_void_p, _f64p, _f64p]
        L.psb200_nh_last_error.argtypes = [C.c_void_p]
        L.psb200_nh_last_error.restype = C.c_char_p
        X = np.ascontiguousarray(X, np.float64)
        T = np.ascontiguousarray(T, np.int32)
        self.nn = X.shape[0]
        self.n = 3 * self.nn
        fx = np.zeros(self.n, np.uint8) if fixed is None else np.ascontiguousarray(np.asarray(fixed).reshape(-1), np.uint8)
        self._fx = fx
        self._h = C.c_void_p()
        rc = L.psb200_nh_create(C.byref(self._h), self.nn, X.reshape(-1), T.shape[0], T.reshape(-1), mu, lam, fx.ctypes.data_as(C.c_void_p), device)
        if rc:
            raise RuntimeError(L.psb200_nh_last_error(None).decode())
        n, nnz = C.c_int64(), C.c_int64()
        po, pi = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
        self._check(L.psb200_nh_pattern(self._h, C.byref(n), C.byref(nnz), C.byref(po), C.byref(pi)))
        self.nnz = nnz.value
        # copies: the arrays the solver hashes every Newton step must outlive any call
        self.outer = np.ctypeslib.as_array(po, shape=(self.n + 1,)).copy()
        self.inner = np.ctypeslib.as_array(pi, shape=(self.nnz,)).copy()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.psb200_nh_destroy(h)
            self._h = None

    def _check(self, rc):
        if rc:
            raise RuntimeError(self._L.psb200_nh_last_error(self._h).decode())

    def value(self, x):
        v = C.c_double()
        self._check(self._L.psb200_nh_value(self._h, np.ascontiguousarray(x, np.float64), C.byref(v)))
        return v.value

    def gradient(self, x):
        g = np.empty(self.n)
        self._check(self._L.psb200_nh_gradient(self._h, np.ascontiguousarray(x, np.float64), g))
        return g

    def hessian_device(self, x, project_to_psd=False):
        d = C.c_void_p()
        self._check(self._L.psb200_nh_hessian_device(self._h, np.ascontiguousarray(x, np.float64), C.byref(d)))
        return self.outer, self.inner, d.value

    def hessian(self, x, project_to_psd=False):
        import scipy.sparse as sp
        vals = np.empty(self.nnz)
        self._check(self._L.psb200_nh_hessian_host(self._h, np.ascontiguousarray(x, np.float64), vals))
        return sp.csc_matrix((vals, self.inner, self.outer), shape=(self.n, self.n))


def stretch_problem(m, stretch=0.1, mu=1.0, lam=1.5, device=-1):
    """Config 5: m^3-node block, face x = 0 clamped, face x = m - 1 pulled by stretch * (m - 1); start = affine stretch."""
    X, T = grid_tets(m)
    fixed = np.zeros((m ** 3, 3), np.uint8)
    fixed[X[:, 0] == 0] = 1
    fixed[X[:, 0] == m - 1] = 1
    x0 = np.zeros((m ** 3, 3))
    x0[:, 0] = stretch * X[:, 0]
    return NeoHookeanDevice(X, T, mu, lam, fixed.reshape(-1), device), x0.reshape(-1)
