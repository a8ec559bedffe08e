"""ncu target: one C2 Jacobi-PCG solve with the persistent kernel (32 iterations, graphs off)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import polysolve_b200 as psb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 216
mode = sys.argv[2] if len(sys.argv) > 2 else "persistent"
P = psb.problems
o, i, v = P.poisson3d(n)
N = n ** 3
b = P.spmv_csr(o, i, v, P.splitmix64(42, N))
s = psb.Solver.create("CUDA", "")
s.set_parameters({"CUDA": {"tolerance": 1e-8, "max_iter": 32, "use_graph": False, "check_every": 8, "cg_kernel": mode}})
s.factorize_raw(N, o, i, v)
x = np.zeros(N)
s.solve(b, x)
print(s.get_info()["solver_iter"], s.get_info().get("persist_cycles"))
