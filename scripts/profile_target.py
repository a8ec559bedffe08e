"""Short C2 workload for ncu captures: analyze + factorize + a 12-iteration Jacobi-PCG solve
(graphs off so every launch is a plain kernel launch). Usage: python scripts/profile_target.py [n] [precond]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import polysolve_b200 as psb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 216
precond = sys.argv[2] if len(sys.argv) > 2 else "jacobi"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 12
P = psb.problems
o, i, v = P.poisson3d(n)
N = n ** 3
b = P.spmv_csr(o, i, v, P.splitmix64(42, N))
s = psb.Solver.create("CUDA", "")
s.set_parameters({"CUDA": {"tolerance": 1e-8, "max_iter": iters, "use_graph": False, "check_every": 4, "precond": precond}})
s.factorize_raw(N, o, i, v)
x = np.zeros(N)
s.solve(b, x)
print(s.get_info())
