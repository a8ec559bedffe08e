"""Generates the committed golden fixtures. The reference's tests hold no golden vectors for this
path and its fixtures (polyfem-data) are not vendored, so the fixtures are produced
by INDEPENDENT solvers (scipy direct solve / scipy cg) on the deterministic C1-style inputs, plus
the iteration counts of the oracle restatement that SURVEY A.5 recorded. Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # C1: 32x32 Poisson, b = splitmix64(42), Jacobi-PCG tol 1e-10
    o, i, v = O.poisson2d(32)
    b = O.splitmix64(42, 1024)
    A = sp.csc_matrix((v, i, o), shape=(1024, 1024))
    x, it, err, _ = O.eigen_cg(o, i, v, b, tol=1e-10)
    xs, info = spla.cg(A, b, rtol=1e-10, M=sp.diags(1 / A.diagonal()), maxiter=1000)
    assert info == 0 and np.abs(xs - x).max() < 1e-12
    x_direct = spla.spsolve(A, b)
    np.savez_compressed(os.path.join(HERE, "c1_poisson2d_32.npz"), b=b, x=x, x_direct=x_direct, iters=it, err=err)
    # non-symmetric convection-diffusion, BiCGSTAB
    o, i, v = O.convdiff2d(32, 0.5)
    A = sp.csc_matrix((v, i, o), shape=(1024, 1024))
    x, it, err, _ = O.eigen_bicgstab(o, i, v, b, tol=1e-10)
    x_direct = spla.spsolve(A, b)
    assert np.abs(x - x_direct).max() < 1e-9
    np.savez_compressed(os.path.join(HERE, "convdiff2d_32.npz"), b=b, x=x, x_direct=x_direct, iters=it, err=err)
    # 3-D Poisson 24^3 manufactured solution (direct solve as the independent answer)
    n = 24
    o, i, v = O.poisson3d(n)
    xstar = O.splitmix64(42, n ** 3)
    b3 = O.spmv_csc(o, i, v, xstar)
    x, it, err, _ = O.eigen_cg(o, i, v, b3, tol=1e-8, max_iters=10000)
    np.savez_compressed(os.path.join(HERE, "poisson3d_24.npz"), b=b3, xstar=xstar, x=x, iters=it, err=err)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
