"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle (liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product (polysolve_b200/) never does.  PARITY UNPINNED: see
oracle_core.hpp -- Eigen 5.0.1 / AMGCL 1.4.3 are not available, the oracle restates them.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


class AmgParams(C.Structure):
    _fields_ = [(k, C.c_int) for k in (
        "max_levels", "coarse_enough", "direct_coarse", "ncycle", "npre", "npost", "pre_cycles",
        "degree", "power_iters", "scale", "estimate_spectral_radius", "relax_type")] + [
        (k, C.c_double) for k in ("higher", "lower", "sa_relax", "eps_strong")]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.orc_splitmix64_fill.argtypes = [C.c_uint64, C.c_int64, f64p]
    L.orc_poisson2d_nnz.restype = C.c_int64
    L.orc_poisson3d_nnz.restype = C.c_int64
    L.orc_poisson2d.argtypes = [C.c_int, i32p, i32p, f64p]
    L.orc_poisson3d.argtypes = [C.c_int, i32p, i32p, f64p]
    L.orc_convdiff2d.argtypes = [C.c_int, C.c_double, i32p, i32p, f64p]
    L.orc_prefactor_values.argtypes = [C.c_int64, i32p, i32p, C.c_int, f64p]
    L.orc_csc_to_csr.argtypes = [C.c_int64, C.c_int64, i32p, i32p, i32p, i32p, i32p]
    L.orc_partition_rows.argtypes = [C.c_int64, i32p, C.c_int, C.c_int, i64p]
    L.orc_halo_for_rank.argtypes = [C.c_int64, i32p, i32p, C.c_int64, C.c_int64, i32p, i32p, C.c_int64]
    L.orc_halo_for_rank.restype = C.c_int64
    L.orc_spmv_csc.argtypes = [C.c_int64, i32p, i32p, f64p, f64p, f64p]
    L.orc_spmv_csr.argtypes = [C.c_int64, i32p, i32p, f64p, f64p, f64p]
    L.orc_eigen_cg.argtypes = [C.c_int64, i32p, i32p, f64p, f64p, f64p, C.c_double, C.c_int64, C.c_int,
                               C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.orc_eigen_bicgstab.argtypes = [C.c_int64, i32p, i32p, f64p, f64p, f64p, C.c_double, C.c_int64, C.c_int,
                                     C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.orc_amg_default_params.argtypes = [C.POINTER(AmgParams)]
    L.orc_amg_create.argtypes = [C.c_int64, i32p, i32p, f64p, C.POINTER(AmgParams)]
    L.orc_amg_create.restype = C.c_void_p
    L.orc_amg_create_imposed.argtypes = [C.c_int64, i32p, i32p, f64p, C.POINTER(AmgParams), C.c_int,
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.orc_amg_create_imposed.restype = C.c_void_p
    L.orc_amg_destroy.argtypes = [C.c_void_p]
    L.orc_amg_num_levels.argtypes = [C.c_void_p]
    L.orc_amg_level_info.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int64)] * 3 + [C.POINTER(C.c_double)] * 3
    L.orc_amg_get_aggregates.argtypes = [C.c_void_p, C.c_int, i32p]
    L.orc_amg_get_matrix.argtypes = [C.c_void_p, C.c_int, C.c_int, i32p, i32p, f64p]
    L.orc_amg_matrix_rows.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_amg_matrix_rows.restype = C.c_int64
    L.orc_amg_apply.argtypes = [C.c_void_p, f64p, f64p]
    L.orc_amg_cg.argtypes = [C.c_void_p, f64p, f64p, C.c_double, C.c_int64, C.POINTER(C.c_double), C.c_void_p, C.c_int64]
    L.orc_amg_cg.restype = C.c_int64
    _LIB = L
    return L


# ---------------------------------------------------------------- helpers
def splitmix64(seed, n):
    out = np.empty(n, np.float64)
    lib().orc_splitmix64_fill(seed, n, out)
    return out


def _gen(nnz, n_rows, fn, *args):
    outer = np.empty(n_rows + 1, np.int32)
    inner = np.empty(nnz, np.int32)
    val = np.empty(nnz, np.float64)
    fn(*args, outer, inner, val)
    return outer, inner, val


def poisson2d(n):
    L = lib()
    return _gen(L.orc_poisson2d_nnz(n), n * n, L.orc_poisson2d, n)


def poisson3d(n):
    L = lib()
    return _gen(L.orc_poisson3d_nnz(n), n ** 3, L.orc_poisson3d, n)


def convdiff2d(n, c):
    L = lib()
    return _gen(L.orc_poisson2d_nnz(n), n * n, L.orc_convdiff2d, n, c)


def prefactor_values(outer, inner, rounds=10):
    n = len(outer) - 1
    out = np.empty(rounds * int(outer[-1]), np.float64)
    lib().orc_prefactor_values(n, outer, inner, rounds, out)
    return out.reshape(rounds, -1)


def csc_to_csr(nrows, outer, inner):
    ncols = len(outer) - 1
    nnz = int(outer[-1])
    rp = np.empty(nrows + 1, np.int32)
    ci = np.empty(nnz, np.int32)
    perm = np.empty(nnz, np.int32)
    lib().orc_csc_to_csr(nrows, ncols, outer, inner, rp, ci, perm)
    return rp, ci, perm


def partition_rows(row_ptr, world, align=1):
    off = np.empty(world + 1, np.int64)
    lib().orc_partition_rows(len(row_ptr) - 1, row_ptr, world, align, off)
    return off


def halo_for_rank(row_ptr, col_idx, r0, r1):
    n = len(row_ptr) - 1
    nl = int(row_ptr[r1] - row_ptr[r0])
    local_col = np.empty(max(nl, 1), np.int32)
    cap = max(nl, 1)
    halo = np.empty(cap, np.int32)
    nh = lib().orc_halo_for_rank(n, row_ptr, col_idx, r0, r1, local_col, halo, cap)
    assert nh >= 0
    return local_col[:nl], halo[:nh].copy()


def spmv_csc(outer, inner, val, x):
    y = np.empty(len(outer) - 1, np.float64)
    lib().orc_spmv_csc(len(outer) - 1, outer, inner, val, np.ascontiguousarray(x), y)
    return y


def spmv_csr(ptr, col, val, x):
    y = np.empty(len(ptr) - 1, np.float64)
    lib().orc_spmv_csr(len(ptr) - 1, ptr, col, val, np.ascontiguousarray(x), y)
    return y


def eigen_cg(ptr, idx, val, b, x0=None, tol=1e-10, max_iters=1000, mode=0, stop_after=0):
    """Eigen::ConjugateGradient<..., Lower|Upper, DiagonalPreconditioner>::solveWithGuess restatement.
    Returns (x, iterations(), error(), spmv_count)."""
    n = len(ptr) - 1
    x = np.zeros(n) if x0 is None else np.array(x0, np.float64, copy=True)
    it, err, sp = C.c_int64(), C.c_double(), C.c_int64()
    lib().orc_eigen_cg(n, ptr, idx, val, np.ascontiguousarray(b), x, tol, max_iters, mode, stop_after,
                       C.byref(it), C.byref(err), C.byref(sp))
    return x, it.value, err.value, sp.value


def eigen_bicgstab(ptr, idx, val, b, x0=None, tol=1e-10, max_iters=1000, mode=0):
    n = len(ptr) - 1
    x = np.zeros(n) if x0 is None else np.array(x0, np.float64, copy=True)
    it, err, sp = C.c_int64(), C.c_double(), C.c_int64()
    lib().orc_eigen_bicgstab(n, ptr, idx, val, np.ascontiguousarray(b), x, tol, max_iters, mode,
                             C.byref(it), C.byref(err), C.byref(sp))
    return x, it.value, err.value, sp.value


class Amg:
    """amgcl::make_solver<amg<builtin, smoothed_aggregation, chebyshev>, cg> restatement with
    polysolve's defaults (reference src/polysolve/linear/AMGCL.cpp:32-65)."""

    def __init__(self, ptr, col, val, imposed=None, **kw):
        L = lib()
        self.prm = AmgParams()
        L.orc_amg_default_params(C.byref(self.prm))
        for k, v in kw.items():
            assert hasattr(self.prm, k), k
            setattr(self.prm, k, v)
        self.n = len(ptr) - 1
        if imposed:
            arrs = [np.ascontiguousarray(a, np.int32) if a is not None else np.zeros(0, np.int32) for a in imposed]
            ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
            lens = (C.c_int64 * len(arrs))(*[len(a) for a in arrs])
            self.h = L.orc_amg_create_imposed(self.n, ptr, col, val, C.byref(self.prm), len(arrs), ptrs, lens)
        else:
            self.h = L.orc_amg_create(self.n, ptr, col, val, C.byref(self.prm))
        if not self.h:
            raise RuntimeError("oracle AMG setup failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_amg_destroy(self.h)
            self.h = None

    @property
    def num_levels(self):
        return lib().orc_amg_num_levels(self.h)

    def level_info(self, l):
        r, z, pz = C.c_int64(), C.c_int64(), C.c_int64()
        rho, d, c = C.c_double(), C.c_double(), C.c_double()
        lib().orc_amg_level_info(self.h, l, C.byref(r), C.byref(z), C.byref(pz), C.byref(rho), C.byref(d), C.byref(c))
        return dict(rows=r.value, nnz=z.value, p_nnz=pz.value, rho=rho.value, d=d.value, c=c.value)

    def aggregates(self, l):
        out = np.empty(self.level_info(l)["rows"], np.int32)
        lib().orc_amg_get_aggregates(self.h, l, out)
        return out

    def matrix(self, l, which="A"):
        w = {"A": 0, "P": 1, "R": 2}[which]
        info = self.level_info(l)
        rows = lib().orc_amg_matrix_rows(self.h, l, w)
        nnz = info["nnz"] if w == 0 else info["p_nnz"]
        ptr = np.empty(rows + 1, np.int32)
        col = np.empty(nnz, np.int32)
        val = np.empty(nnz, np.float64)
        lib().orc_amg_get_matrix(self.h, l, w, ptr, col, val)
        return ptr, col, val

    def apply(self, rhs):
        x = np.empty(self.n)
        lib().orc_amg_apply(self.h, np.ascontiguousarray(rhs), x)
        return x

    def cg(self, b, x0=None, tol=1e-10, maxiter=1000, hist=False):
        x = np.zeros(self.n) if x0 is None else np.array(x0, np.float64, copy=True)
        rel = C.c_double()
        h = np.zeros(maxiter) if hist else None
        it = lib().orc_amg_cg(self.h, np.ascontiguousarray(b), x, tol, maxiter, C.byref(rel),
                              h.ctypes.data if hist else None, maxiter if hist else 0)
        if hist:
            return x, it, rel.value, h[:it]
        return x, it, rel.value
