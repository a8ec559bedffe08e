"""ncu target: plain SpMV launches on the squared Poisson operator (25 nnz/row, stands for AMG level 1).
    python scripts/spmv_target.py [n] [kernel ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import scipy.sparse as sp  # noqa: E402

import polysolve_b200 as psb  # noqa: E402

P = psb.problems
pn = int(sys.argv[1]) if len(sys.argv) > 1 else 128
kernels = sys.argv[2:] or ["stream4"]
o, i, v = P.poisson3d(pn)
A = sp.csr_matrix((v, i, o), shape=(pn ** 3, pn ** 3))
A2 = (A @ A).tocsr()
A2.sort_indices()
s = psb.Solver.create("CUDA", "")
s.factorize_raw(pn ** 3, A2.indptr.astype(np.int32), A2.indices.astype(np.int32), A2.data.astype(np.float64))
for k in kernels:
    print(k, s.bench_spmv(reps=2, kernel=k))
