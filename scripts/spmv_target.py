"""ncu target: plain SpMV launches on a long-row matrix.
    python scripts/spmv_target.py poisson2:128 [kernel ...]     squared 3-D Poisson operator (25 nnz/row, stands for AMG level 1)
    python scripts/spmv_target.py elasticity:72 [kernel ...]    P1 linear elasticity (43 nnz/row)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import scipy.sparse as sp  # noqa: E402

import polysolve_b200 as psb  # noqa: E402

P = psb.problems
what, _, size = (sys.argv[1] if len(sys.argv) > 1 else "poisson2:128").partition(":")
size = int(size or 128)
kernels = sys.argv[2:] or [""]
if what == "elasticity":
    o, i, v, _ = P.elasticity3d(size)
else:
    o, i, v = P.poisson3d(size)
    A = sp.csr_matrix((v, i, o), shape=(size ** 3, size ** 3))
    A2 = (A @ A).tocsr()
    A2.sort_indices()
    o, i, v = A2.indptr.astype(np.int32), A2.indices.astype(np.int32), A2.data.astype(np.float64)
n = len(o) - 1
s = psb.Solver.create("CUDA", "")
s.factorize_raw(n, o, i, v)
for k in kernels:
    print(k or s.get_info()["spmv_kernel"], s.bench_spmv(reps=2, kernel=k))
