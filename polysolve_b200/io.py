"""Matrix Market replay of the reference's fixtures (include/psb200_io.h): the host-side mirror of
Eigen::loadMarket / saveMarket / loadMarketVector / saveMarketVector and of the tests' loadSymmetric
(reference tests/test_linear_solver.cpp:25-50). Matrices come back as scipy CSC, the layout of StiffnessMatrix."""
import ctypes as C

import numpy as np

from . import _lib


def _err(L):
    return RuntimeError(L.psb200_market_last_error().decode())


def load_market(path, symmetric=0):
    """symmetric: 0 = entries as stored (Eigen::loadMarket), 1 = mirror the stored triangle (loadSymmetric),
    -1 = follow the header."""
    import scipy.sparse as sp
    L = _lib.lib()
    h = C.c_void_p()
    rows, cols, nnz = C.c_int64(), C.c_int64(), C.c_int64()
    if L.psb200_market_load(str(path).encode(), int(symmetric), C.byref(h), C.byref(rows), C.byref(cols), C.byref(nnz)):
        raise _err(L)
    try:
        outer = np.empty(cols.value + 1, np.int32)
        inner = np.empty(max(nnz.value, 1), np.int32)
        vals = np.empty(max(nnz.value, 1), np.float64)
        if L.psb200_market_get_csc(h, outer, inner, vals):
            raise _err(L)
    finally:
        L.psb200_market_free(h)
    return sp.csc_matrix((vals[:nnz.value], inner[:nnz.value], outer), shape=(rows.value, cols.value))


def load_symmetric(path):
    """loadSymmetric of the reference's tests (test_linear_solver.cpp:25-50)."""
    return load_market(path, symmetric=1)


def save_market(A, path, symmetric=False):
    import scipy.sparse as sp
    A = sp.csc_matrix(A)
    A.sort_indices()
    L = _lib.lib()
    outer = np.ascontiguousarray(A.indptr, np.int32)
    inner = np.ascontiguousarray(A.indices if A.nnz else np.zeros(1), np.int32)
    vals = np.ascontiguousarray(A.data if A.nnz else np.zeros(1), np.float64)
    if L.psb200_market_save(str(path).encode(), A.shape[0], A.shape[1], outer, inner, vals, int(bool(symmetric))):
        raise _err(L)


def load_market_vector(path):
    L = _lib.lib()
    n = C.c_int64()
    if L.psb200_market_load_vector(str(path).encode(), None, 0, C.byref(n)):
        raise _err(L)
    out = np.empty(max(n.value, 1), np.float64)
    if L.psb200_market_load_vector(str(path).encode(), out.ctypes.data, out.shape[0], C.byref(n)):
        raise _err(L)
    return out[:n.value]


def save_market_vector(v, path):
    L = _lib.lib()
    v = np.ascontiguousarray(v, np.float64)
    if L.psb200_market_save_vector(str(path).encode(), v if v.size else np.zeros(1), v.shape[0]):
        raise _err(L)
