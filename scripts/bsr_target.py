"""ncu target: BSR-3 SpMV (72^3-node P1 elasticity, block size 3) -- a few launches of spmv_bsr3_kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polysolve_b200 as psb  # noqa: E402

P = psb.problems
o, i, v, b = P.elasticity3d(int(sys.argv[1]) if len(sys.argv) > 1 else 72)
n = len(b)
s = psb.Solver.create("CUDA", "")
s.set_parameters({"CUDA": {"block_size": 3}})
s.factorize_raw(n, o, i, v)
print(s.get_info()["spmv_kernel"], s.bench_spmv(reps=5))
for k in ("bsr:256:544:2", "bsr:256:544:1", "bsr:128:272:2", "bsr:128:272:3", "bsr:128:288:4", "bsr:512:1088:1"):
    if "--sweep" in sys.argv:
        print(k, round(s.bench_spmv(reps=30, kernel=k), 5), flush=True)
