"""C2 (216^3 Poisson) Jacobi-PCG: persistent cooperative kernel vs kernel-per-phase, device-resident timing."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import polysolve_b200 as psb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 216
P = psb.problems
o, i, v = P.poisson3d(n)
N = n ** 3
b = P.spmv_csr(o, i, v, P.splitmix64(42, N))
db = torch.from_numpy(b).cuda()
for mode in ("split", "persistent", "split", "persistent"):
    for ce in (16, 32):
        s = psb.Solver.create("CUDA", "")
        s.set_parameters({"CUDA": {"tolerance": 1e-8, "max_iter": 10000, "cg_kernel": mode, "check_every": ce}})
        s.factorize_raw(N, o, i, v)
        dx = torch.zeros(N, dtype=torch.float64, device="cuda")
        ts = []
        for rep in range(4):
            dx.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s.solve_device(db.data_ptr(), dx.data_ptr(), N)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        info = s.get_info()
        x = dx.cpu().numpy()
        rel = np.linalg.norm(P.spmv_csr(o, i, v, x) - b) / np.linalg.norm(b)
        print(mode, "check_every", ce, "iters", info["solver_iter"], info["solver_status"], "best ms", 1e3 * min(ts),
              "it/s", info["solver_iter"] / min(ts), "rel", rel, "launches", info["gpu_launches"],
              "phase_us_per_iter", [round(c / 1.965e3 / max(1, info["solver_iter"]), 2) for c in info.get("persist_cycles", [])], flush=True)
        del s
