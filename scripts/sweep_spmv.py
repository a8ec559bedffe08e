"""SpMV schedule / tile-shape sweep on the C2 matrix (216^3). Prints GB/s (algorithmic bytes) per variant."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polysolve_b200 as psb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 216
P = psb.problems
o, i, v = P.poisson3d(n)
N = n ** 3
s = psb.Solver.create("CUDA", "")
s.factorize_raw(N, o, i, v)
B = P.spmv_bytes(N, len(i))
variants = ["", "scalar", "vector2", "vector4", "vector8"]
for t, c, st in [(256, 2560, 3), (256, 2048, 2), (256, 2048, 3), (256, 2048, 4), (128, 1024, 2), (128, 1024, 3), (128, 1024, 4),
                 (512, 4096, 2), (512, 4096, 3), (64, 512, 4)]:
    variants.append(f"stream:{t}:{c}:{st}")
    for ctas in (1, 2, 3, 4, 6, 8):
        variants.append(f"stream:{t}:{c}:{st}:{ctas}")
for k in variants:
    try:
        ms = min(s.bench_spmv(reps=30, kernel=k) for _ in range(3))
        print(f"{k or 'configured':28s} {ms * 1e3:8.1f} us  {B / ms / 1e6:8.1f} GB/s  {B / ms / 1e6 / 6553.6:6.3f} of measured peak", flush=True)
    except Exception as e:
        print(k, "FAILED", e, flush=True)
