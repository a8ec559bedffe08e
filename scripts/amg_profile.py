"""C3 workload (216^3 Poisson, SA-AMG-PCG with polysolve's AMGCL defaults) for timing and ncu launch lists.

    python scripts/amg_profile.py timers [n]       -> 4 x factorize with the per-level setup timers, 3 x solve
    python scripts/amg_profile.py launches [n]     -> factorize (outside the profiler range) + ONE solve with graphs off
                                                      between cudaProfilerStart/Stop (run under ncu --profile-from-start off)
    python scripts/amg_profile.py setup [n]        -> one warm factorize, then ONE factorize inside the profiler range
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import polysolve_b200 as psb  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "timers"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 216
P = psb.problems
o, i, v = P.poisson3d(n)
N = n ** 3
b = P.spmv_csr(o, i, v, P.splitmix64(42, N))
s = psb.Solver.create("CUDA", "")
prm = {"precond": "amg", "tolerance": 1e-8, "max_iter": 1000}
if mode == "launches":
    prm["use_graph"] = False
s.set_parameters({"CUDA": prm})
s.analyze_pattern_raw(N, o, i, N)
if mode == "timers":
    for k in range(4):
        t0 = time.perf_counter()
        s.factorize_raw(N, o, i, v)
        dt = time.perf_counter() - t0
        info = s.get_info()
        print(json.dumps({"factorize": k, "wall_s": dt, "setup_ms": [lv.get("setup_ms") for lv in info["amg"]["levels"]]}), flush=True)
    for k in range(3):
        x = np.zeros(N)
        t0 = time.perf_counter()
        s.solve(b, x)
        dt = time.perf_counter() - t0
        info = s.get_info()
        print(json.dumps({"solve": k, "wall_s": dt, "iters": info["num_iterations"], "launches": info["gpu_launches"],
                          "solve_ms": info.get("solve_ms"), "amg": {kk: vv for kk, vv in info["amg"].items() if kk != "levels"}}), flush=True)
elif mode == "setup":
    s.factorize_raw(N, o, i, v)
    import torch
    torch.cuda.cudart().cudaProfilerStart()
    s.factorize_raw(N, o, i, v)
    torch.cuda.cudart().cudaProfilerStop()
    print(json.dumps({"levels": [lv["rows"] for lv in s.get_info()["amg"]["levels"]]}))
else:
    s.factorize_raw(N, o, i, v)
    import torch
    x = np.zeros(N)
    torch.cuda.cudart().cudaProfilerStart()
    s.solve(b, x)
    torch.cuda.cudart().cudaProfilerStop()
    print(json.dumps({"iters": s.get_info()["num_iterations"]}))
