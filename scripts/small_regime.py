"""Per-iteration cost of Jacobi-PCG on the per-GPU shares of the 8 / 4 / 2-GPU runs (108^3, 136^3, 171^3 rows on ONE GPU,
no communication): what is left above the HBM time is launch / drain / reduction overhead of the kernel chain.
    python scripts/small_regime.py [cg_kernel] [n ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import polysolve_b200 as psb  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "split"
sizes = [int(a) for a in sys.argv[2:]] or [108, 136, 171, 216]
P = psb.problems
for n in sizes:
    o, i, v = P.poisson3d(n)
    N = n ** 3
    b = P.spmv_csr(o, i, v, P.splitmix64(42, N))
    db = torch.from_numpy(b).cuda()
    s = psb.Solver.create("CUDA", "")
    s.set_parameters({"CUDA": {"tolerance": 1e-8, "max_iter": 10000, "cg_kernel": mode.split("-")[0], "check_every": 16,
                               "pdl": not mode.endswith("-nopdl")}})
    s.factorize_raw(N, o, i, v)
    dx = torch.zeros(N, dtype=torch.float64, device="cuda")
    best = 1e9
    for rep in range(5):
        dx.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.solve_device(db.data_ptr(), dx.data_ptr(), N)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    it = s.get_info()["solver_iter"]
    ideal_us = P.pcg_iter_bytes(N, int(o[-1])) / 6538.6e9 * 1e6
    print(f"{mode} n={n}^3 rows={N} iters={it} us/iter={1e6 * best / it:.1f} (HBM time {ideal_us:.1f}) it/s={it / best:.0f}", flush=True)
    del s
