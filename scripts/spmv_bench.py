"""SpMV GB/s (algorithmic bytes 12 nnz + 20 n + 4) of every schedule on three matrix families:
poisson3d (7 nnz/row), its Galerkin-like square (25 nnz/row, stands for AMG level 1) and P1 elasticity (45 nnz/row).
    python scripts/spmv_bench.py [poisson_n] [elasticity_m]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import scipy.sparse as sp  # noqa: E402

import polysolve_b200 as psb  # noqa: E402

P = psb.problems
pn = int(sys.argv[1]) if len(sys.argv) > 1 else 128
em = int(sys.argv[2]) if len(sys.argv) > 2 else 72
peak = 6538.6
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def run(tag, o, i, v, kernels):
    n = len(o) - 1
    nnz = int(o[-1])
    s = psb.Solver.create("CUDA", "")
    s.factorize_raw(n, o, i, v)
    auto = s.get_info()["spmv_kernel"]
    x = P.splitmix64(3, n)
    y_ref = sp.csr_matrix((v, i, o), shape=(n, n)) @ x
    for k in ["auto"] + kernels:
        ms = s.bench_spmv(reps=30, kernel="" if k == "auto" else k)
        y = s.spmv(x) if k == "auto" else None
        err = float(np.max(np.abs(y - y_ref)) / np.max(np.abs(y_ref))) if y is not None else None
        gbs = P.spmv_bytes(n, nnz) / (ms * 1e-3) / 1e9
        print(json.dumps({"matrix": tag, "n": n, "nnz": nnz, "nnz_per_row": round(nnz / n, 1), "kernel": k if k != "auto" else "auto=" + auto,
                          "ms": round(ms, 4), "GB/s": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 3), "max_rel_err": err}), flush=True)


o, i, v = P.poisson3d(pn)
run(f"poisson3d_{pn}", o, i, v, ["vector2", "vector4"])
A = sp.csr_matrix((v, i, o), shape=(pn ** 3, pn ** 3))
A2 = (A @ A).tocsr()
A2.sort_indices()
run(f"poisson3d_{pn}_squared", A2.indptr.astype(np.int32), A2.indices.astype(np.int32), A2.data.astype(np.float64),
    ["stream4", "stream4n", "stream4m", "stream8n", "vector8"])
# 33 nnz/row with scattered columns: the level-1 Galerkin matrix of the 216^3 Laplacian has this density; here A^2 of the
# 7-point stencil plus the 8 body-diagonal neighbours
import itertools  # noqa: E402
n3 = 96
idx = np.arange(n3 ** 3).reshape(n3, n3, n3)
rows, cols = [], []
offs = [d for d in itertools.product((-2, -1, 0, 1, 2), repeat=3) if sum(abs(t) for t in d) <= 2] + list(itertools.product((-1, 1), repeat=3))
for d in offs:
    src = idx[max(0, -d[0]):n3 - max(0, d[0]), max(0, -d[1]):n3 - max(0, d[1]), max(0, -d[2]):n3 - max(0, d[2])]
    dst = idx[max(0, d[0]):n3 - max(0, -d[0]), max(0, d[1]):n3 - max(0, -d[1]), max(0, d[2]):n3 - max(0, -d[2])]
    rows.append(src.reshape(-1))
    cols.append(dst.reshape(-1))
M = sp.csr_matrix((np.ones(sum(len(r) for r in rows)), (np.concatenate(rows), np.concatenate(cols))), shape=(n3 ** 3, n3 ** 3))
M.sort_indices()
run(f"stencil33_{n3}", M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64), ["stream4", "stream4m", "stream8n", "vector16"])
del M
del A, A2
o, i, v, _ = P.elasticity3d(em)
run(f"elasticity3d_{em}", o, i, v, ["stream4", "stream8n", "vector16"])
# the same matrix as a block-3 matrix: BSR-3 schedule (76 B per block; GB/s still counted on the scalar-CSR bytes, so a
# value above the HBM peak means fewer bytes moved, and on its own algorithmic bytes 76 nnzb + 4 nb + 16 n)
s = psb.Solver.create("CUDA", "")
s.set_parameters({"CUDA": {"block_size": 3}})
n = len(o) - 1
s.factorize_raw(n, o, i, v)
x = P.splitmix64(3, n)
y_ref = sp.csr_matrix((v, i, o), shape=(n, n)) @ x
info = s.get_info()
ms = s.bench_spmv(reps=30)
err = float(np.max(np.abs(s.spmv(x) - y_ref)) / np.max(np.abs(y_ref)))
nnzb = int(o[-1]) // 9   # the pattern has full blocks
bsr_bytes = 76 * nnzb + 4 * (n // 3) + 16 * n
print(json.dumps({"matrix": f"elasticity3d_{em} block_size 3", "kernel": "auto=" + info["spmv_kernel"], "ms": round(ms, 4),
                  "GB/s_on_csr_bytes": round(P.spmv_bytes(n, int(o[-1])) / (ms * 1e-3) / 1e9, 1),
                  "GB/s_on_bsr_bytes": round(bsr_bytes / (ms * 1e-3) / 1e9, 1), "frac_of_measured_peak_bsr_bytes": round(bsr_bytes / (ms * 1e-3) / 1e9 / peak, 3),
                  "max_rel_err": err}), flush=True)
ms = s.bench_spmv(reps=30, kernel="stream8n")
print(json.dumps({"matrix": f"elasticity3d_{em} block_size 3", "kernel": "stream8n (scalar CSR of the same matrix)", "ms": round(ms, 4)}), flush=True)
