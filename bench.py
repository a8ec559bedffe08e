#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric : CG iterations / second (Jacobi-PCG, Eigen ordering) with the SpMV kernel's achieved HBM
         bandwidth against the measured roofline (MEASURED_PEAKS.json).
config : configs[1] of BASELINE.json -- 10M-DoF (216^3) 3-D Poisson 7-point, Jacobi-PCG, tol 1e-8.
step   : one complete solve(b, x) from x0 = 0 to the relative residual 1e-8.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n GRID]

`value` : device-resident (b, x in HBM), CUDA events on the solver's stream, max over ranks.
`e2e`   : the same metric through the public C-ABI call path with HOST buffers: every step does
          factorize_csc(host values) + solve(host b, host x), so the H2D of the matrix values and
          of b/x and the D2H of x are inside the timed region.
N > 1   : the same ONE system, row-partitioned; from 4 ranks on the Krylov method is the single-reduction CG (krylov = cg1r:
          the same iterates as the Eigen ordering in exact arithmetic, one all-reduce per iteration), stated in
          config.krylov; the other ordering is timed in the same run (other_krylov), and a `parity` block checks the
          N-rank path against the oracle on a small system (partition / halo lists bit-exact, x, iteration counts).
`--config c4`: BASELINE configs[3] (119^3-node P1 elasticity, block-3 SA-AMG-PCG) as a separate line.
`--impl reference`: the reference's CPU path (Eigen::ConjugateGradient restatement from oracle/,
          see DESIGN.md: Eigen/AMGCL are not installable offline) on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "cg_iters_per_sec"
UNIT = "iter/s"
TOL = 1e-8
MAX_ITER = 10000
# row partitions: how long a kernel may wait for a peer before the solve fails (library default 3 s). The bench starts on
# a cold box with up to 8 ranks sharing the host cores for set-up, so it allows more skew; a lost rank still surfaces.
COMM_TIMEOUT_S = 20.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi (B200_PROFILING.md 'clocks line'). Started before the warm-up steps -- nvidia-smi needs a few
    hundred ms to come up, longer than the timed region of a multi-GPU run -- and marked at the edges of the timed
    region: the reported median is over the samples inside it, or, when none fell inside, over the samples taken
    under the same load during warm-up (the count of both is reported)."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = self.tl = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_load(self):
        self.tl = time.perf_counter()

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def digest(rows):
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[0]))
                    mx.append(float(r[1]))
                    for nm, val in zip(names, r[3:7]):
                        if val.lower().startswith("active"):
                            reasons.add(nm)
                except Exception:
                    continue
            return sm, mx, reasons
        inside = [x for x in self.rows if self.t0 is not None and self.t1 is not None and self.t0 <= x[0] <= self.t1]
        loaded = [x for x in self.rows if self.tl is not None and self.t1 is not None and self.tl <= x[0] <= self.t1]
        sm, mx, reasons = digest(inside if inside else (loaded if loaded else self.rows))
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": len(digest(inside)[0]), "samples_under_load": len(digest(loaded)[0]),
                "samples_total": len(self.rows)}


def build_problem(n):
    from polysolve_b200 import problems as P
    N = n ** 3
    outer, inner, vals = P.poisson3d(n)
    xstar = P.splitmix64(42, N)
    b = P.spmv_csr(outer, inner, vals, xstar)
    return N, outer, inner, vals, b, xstar


def pinned_copy(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the reference arm is the CPU path on ALL host cores (libgomp reads
    # the variable when the oracle library is loaded, which happens below)
    if world > 1 and os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle import oracle as O
    n = args.n
    N, outer, inner, vals, b, _ = build_problem(n)
    threads = O.lib().orc_num_threads()
    sample_iters = args.ref_iters
    # warm-up + timed steps; one step = `sample_iters` loop trips of the CG on the full-size system
    # (a bounded sample: the complete solve needs ~650 trips, i.e. minutes on a CPU)
    mode = 1  # all host cores: row-parallel OpenMP CSR SpMV/dots (symmetric matrix: CSC arrays == CSR)
    for _ in range(max(1, min(args.warmup, 1))):
        O.eigen_cg(outer, inner, vals, b, tol=TOL, max_iters=MAX_ITER, mode=mode, stop_after=2)
    t0 = time.perf_counter()
    its = 0
    for _ in range(args.steps):
        _, it, _, _ = O.eigen_cg(outer, inner, vals, b, tol=TOL, max_iters=MAX_ITER, mode=mode, stop_after=sample_iters)
        its += it
    dt = time.perf_counter() - t0
    value = its / dt
    # disclosure: the literally faithful path (CSC column scatter, 1 thread = what
    # Solver::create("Eigen::ConjugateGradient") executes for a column-major matrix)
    t1 = time.perf_counter()
    _, it1, _, _ = O.eigen_cg(outer, inner, vals, b, tol=TOL, max_iters=MAX_ITER, mode=0, stop_after=max(4, sample_iters // 4))
    faithful = it1 / (time.perf_counter() - t1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"poisson3d_{n}^3 ({N} DoF, 7-pt), Jacobi-PCG tol {TOL}", "n": N, "nnz": int(outer[-1]),
                   "note": "reference deps Eigen 5.0.1 / AMGCL 1.4.3 not installable offline; CPU numbers are from an "
                           "in-repo restatement (oracle/) following SURVEY Appendix A"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} x {sample_iters} CG iterations on the full {N}-DoF system, OpenMP row-parallel CSR",
                         "eigen_faithful_1thread": faithful},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def amg_roofline(info, t_solve, hbm_peak, world=1):
    """Algorithmic HBM bytes of one AMG-PCG solve from the actual hierarchy: per preconditioner application
    level l is entered ncycle^l times; every entry of a non-last level runs ncycle x (pre-smooth + residual + post-smooth)
    = ncycle x (2 degree + 1) SpMVs (the first pre-smoothing step from a zero iterate is a vector kernel), one restriction
    and one prolongation per inner cycle; the last level runs 2 degree - 1 SpMVs. Every SpMV counts 12 nnz + 20 n bytes plus
    32 n for the fused Chebyshev epilogue; CG itself adds one fine SpMV and 11 vector passes per iteration."""
    amg = info["amg"]
    lv = amg["levels"]
    nc, deg, its = amg["ncycle"], amg["degree"], max(1, info["num_iterations"])
    per_apply = 0.0
    for l, L in enumerate(lv):
        entries = nc ** l
        n, nnz = L["rows"], L["nnz"]
        spmv = 12.0 * nnz + 20.0 * n
        if l + 1 < len(lv):
            k = entries * (nc * (2 * deg + 1) - 1)
            per_apply += k * (spmv + 32.0 * n) + entries * nc * 2 * (12.0 * L["p_nnz"] + 20.0 * n)
        else:
            per_apply += entries * (2 * deg - 1) * (spmv + 32.0 * n)
    n0, nnz0 = lv[0]["rows"], lv[0]["nnz"]
    total = (its + 1) * per_apply + its * (12.0 * nnz0 + 20.0 * n0 + 88.0 * n0)
    gbs = total / t_solve / 1e9
    return {"bound": "hbm", "achieved": gbs, "peak": hbm_peak * world, "unit": "GB/s", "frac": gbs / (hbm_peak * world),
            "algorithmic_bytes_per_solve": total, "what": "sum over levels of SpMV applications x (12 nnz + 20 n + 32 n epilogue) + transfer operators + CG"}


def amg_leg(args, psb, P, local, hbm_peak):
    """Config 3 of BASELINE.json (SA-AMG-PCG, polysolve's AMGCL defaults) next to the CPU restatement
    (oracle/, OpenMP, all host cores) on the SAME full-size system: setup and solve timed separately.
    --amg-cpu-n < n moves the CPU leg (and a like-for-like GPU run) to a smaller grid."""
    out = {}
    cpu_n = args.amg_cpu_n if args.amg_cpu_n > 0 else args.n
    legs = (("full", args.n),) if cpu_n == args.n else (("full", args.n), ("sample", cpu_n))
    for tag, n in legs:
        N, outer, inner, vals, b, _ = build_problem(n)
        s = psb.Solver.create("CUDA", "")
        s.set_parameters({"CUDA": {"precond": "amg", "tolerance": TOL, "max_iter": 1000, "device": local}})
        s.analyze_pattern_raw(N, outer, inner, N)
        s.factorize_raw(N, outer, inner, vals)  # warm-up (pool growth, module load)
        setups = []
        for _ in range(2):
            t0 = time.perf_counter()
            s.factorize_raw(N, outer, inner, vals)
            setups.append(time.perf_counter() - t0)
        t_setup = min(setups)
        hb, hx = pinned_copy(b), pinned_copy(np.zeros(N))
        x = hx.numpy()
        s.solve(hb.numpy(), x)
        solves = []
        for _ in range(2):
            x[:] = 0
            t0 = time.perf_counter()
            s.solve(hb.numpy(), x)
            solves.append(time.perf_counter() - t0)
        t_solve = min(solves)
        info = s.get_info()
        rel = float(np.linalg.norm(P.spmv_csr(outer, inner, vals, x) - b) / np.linalg.norm(b))
        rf = amg_roofline(info, t_solve, hbm_peak)
        out[tag] = {"n": N, "gpu_setup_s": t_setup, "gpu_solve_s": t_solve, "gpu_setup_s_all": setups, "gpu_solve_s_all": solves, "gpu_iters": info["num_iterations"], "rel_residual": rel,
                    "roofline": rf,
                    "levels": [lv["rows"] for lv in info["amg"]["levels"]], "operator_complexity": info["amg"]["operator_complexity"],
                    "level_spmv_kernels": [lv.get("spmv_kernel") for lv in info["amg"]["levels"]],
                    "gpu_setup_ms_by_level": [lv.get("setup_ms") for lv in info["amg"]["levels"]]}
        del s
        if n == cpu_n and not args.no_cpu:
            from oracle import oracle as O
            t0 = time.perf_counter()
            H = O.Amg(outer, inner, vals)
            c_setup = time.perf_counter() - t0
            t0 = time.perf_counter()
            _, itc, relc = H.cg(b, tol=TOL)
            c_solve = time.perf_counter() - t0
            out[tag].update({"cpu_setup_s": c_setup, "cpu_solve_s": c_solve, "cpu_iters": itc, "cpu_rel_residual": relc,
                             "cpu_cores": O.lib().orc_num_threads(), "cpu_kind": "port (oracle/ AMGCL restatement, OpenMP)",
                             "speedup_setup_plus_solve": (c_setup + c_solve) / (t_setup + t_solve),
                             "speedup_solve": c_solve / t_solve})
    return out


def amg_dist_leg(args, psb, P, local, world, N, outer, inner, vals, b, barrier, hbm_peak):
    """Config 3 on the row partition: SA-AMG-PCG with the PARTITIONED hierarchy (decoupled aggregation, rank-local P / R,
    distributed Galerkin product, per-level halo plans, small levels replicated). Wall clock between barriers, max over
    ranks (every rank times the same collective call)."""
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    s = psb.Solver.create("CUDA", "")
    s.set_parameters({"CUDA": {"precond": "amg", "tolerance": TOL, "max_iter": 1000, "device": local, "comm_timeout_s": COMM_TIMEOUT_S,
                               "amg": {"dist_mode": args.amg_dist_mode, "fused_push": bool(args.fused_push)}}})
    s.dist_setup_torch(halo_cap=1 << 20)
    s.analyze_pattern_raw(N, outer, inner, N)
    s.factorize_raw(N, outer, inner, vals)
    setups, solves = [], []
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        s.factorize_raw(N, outer, inner, vals)
        barrier()
        setups.append(time.perf_counter() - t0)
    hb, hx = pinned_copy(b), pinned_copy(np.zeros(N))
    x = hx.numpy()
    s.solve(hb.numpy(), x)
    for _ in range(2):
        x[:] = 0
        barrier()
        t0 = time.perf_counter()
        s.solve(hb.numpy(), x)
        barrier()
        solves.append(time.perf_counter() - t0)
    info = s.get_info()
    s.release_cached_memory()  # setup temporaries cached by the stream-ordered pool are not resident data
    torch.cuda.synchronize()
    used = free0 - torch.cuda.mem_get_info()[0]
    r0, r1 = s.dist_local_range()
    xt = torch.zeros(N, dtype=torch.float64, device="cuda")
    xt[r0:r1] = torch.from_numpy(x[r0:r1]).cuda()
    dist.all_reduce(xt, op=dist.ReduceOp.SUM)
    t = torch.tensor([min(setups), min(solves), float(used)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    xf = xt.cpu().numpy()
    rel = float(np.linalg.norm(P.spmv_csr(outer, inner, vals, xf) - b) / np.linalg.norm(b))
    amg = info["amg"]
    return {"n": N, "n_gpus": world, "gpu_setup_s": float(t[0]), "gpu_solve_s": float(t[1]), "gpu_iters": info["num_iterations"],
            "rel_residual": rel, "levels": [lv["rows"] for lv in amg["levels"]], "operator_complexity": amg["operator_complexity"],
            "partitioned_levels": amg.get("partitioned_levels"), "replicated_levels": amg.get("replicated_levels"),
            "dist_mode": info.get("amg_dist_mode"), "device_bytes_per_rank_max": int(t[2]),
            "fine_matrix_bytes_total": 12 * int(outer[-1]) + 4 * N,
            "roofline": amg_roofline(info, float(t[1]), hbm_peak, world),
            "what": "SA-AMG-PCG on the row partition: every level above amg.replicate_below non-zeros partitioned (decoupled aggregation, "
                    "rank-local P/R, distributed Galerkin product, per-level halo pushes over NVLink), small levels replicated; "
                    "device_bytes_per_rank_max = device memory this leg allocated on the fullest rank (comm buffer 336 MB included)"}


def parity_block(psb, P, local, world, rank, barrier):
    """N-rank path against the oracle (the checker) on a 40^3 Poisson system, printed with the bench line because the
    driver's pytest box has one GPU: partition offsets / halo lists / value map bit-exact against the oracle's index
    functions, x and the iteration count of the Eigen-ordering PCG, the single-reduction CG, and the AMG-PCG iteration
    count of the partitioned hierarchy next to the 1-GPU hierarchy's."""
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    n, tol = 40, 1e-8
    N = n ** 3
    o, i, v = P.poisson3d(n)
    b = P.spmv_csr(o, i, v, P.splitmix64(42, N))
    out = {"n": N, "tol": tol}
    rp, ci, perm = O.csc_to_csr(N, o, i)
    off0 = O.partition_rows(rp, world)
    plan = psb.Solver.dist_plan_host(N, o, i, rank, world, 1 << 16)
    a, e = int(off0[rank]), int(off0[rank + 1])
    lc0, halo0 = O.halo_for_rank(rp, ci, a, e)
    ok = (np.array_equal(plan["offsets"], off0) and np.array_equal(plan["halo_cols"], halo0)
          and np.array_equal(plan["perm"], perm[rp[a]:rp[e]]) and np.array_equal(plan["rp"], rp[a:e + 1] - rp[a]))
    x0, it0, _, _ = O.eigen_cg(o, i, v, b, tol=tol, max_iters=10000)

    def solve(prm):
        s = psb.Solver.create("CUDA", "")
        s.set_parameters({"CUDA": dict({"tolerance": tol, "max_iter": 10000, "device": local, "comm_timeout_s": COMM_TIMEOUT_S}, **prm)})
        s.dist_setup_torch(halo_cap=1 << 16)
        s.analyze_pattern_raw(N, o, i, N)
        s.factorize_raw(N, o, i, v)
        x = np.zeros(N)
        s.solve(b, x)
        info = s.get_info()
        r0, r1 = s.dist_local_range()
        xt = torch.zeros(N, dtype=torch.float64, device="cuda")
        xt[r0:r1] = torch.from_numpy(x[r0:r1]).cuda()
        dist.all_reduce(xt, op=dist.ReduceOp.SUM)
        return xt.cpu().numpy(), info

    x, info = solve({"krylov": "cg"})
    out["pcg_iters"], out["oracle_iters"] = info["solver_iter"], it0
    out["pcg_x_rel_diff_vs_oracle"] = float(np.linalg.norm(x - x0) / np.linalg.norm(x0))
    x, info = solve({"krylov": "cg1r"})
    out["cg1r_iters"] = info["solver_iter"]
    out["cg1r_x_rel_diff_vs_oracle"] = float(np.linalg.norm(x - x0) / np.linalg.norm(x0))
    x, info = solve({"precond": "amg", "max_iter": 200, "amg": {"replicate_below": 50000}})
    out["amg_partitioned_iters"] = info["solver_iter"]
    out["amg_partitioned_levels"] = [lv["rows"] for lv in info["amg"]["levels"]]
    out["amg_rel_residual"] = float(np.linalg.norm(P.spmv_csr(o, i, v, x) - b) / np.linalg.norm(b))
    s1 = psb.Solver.create("CUDA", "")
    s1.set_parameters({"CUDA": {"tolerance": tol, "max_iter": 200, "device": local, "precond": "amg"}})
    s1.factorize_raw(N, o, i, v)
    x1 = np.zeros(N)
    s1.solve(b, x1)
    out["amg_one_gpu_iters"] = s1.get_info()["solver_iter"]
    del s1
    flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["plan_bit_exact_all_ranks"] = bool(flag[0] == 1.0)
    out["ok"] = bool(out["plan_bit_exact_all_ranks"] and abs(out["pcg_iters"] - it0) <= max(1, 0.02 * it0)
                     and abs(out["cg1r_iters"] - it0) <= max(1, 0.02 * it0) and out["pcg_x_rel_diff_vs_oracle"] < 10 * tol
                     and out["cg1r_x_rel_diff_vs_oracle"] < 10 * tol and out["amg_partitioned_iters"] <= out["amg_one_gpu_iters"] + 1
                     and out["amg_rel_residual"] < 2 * tol)
    return out


def cusparse_leg(args, N, outer, inner, vals):
    """the reference's own GPU building blocks on the same box -- cusparseSpMV (CSR, fp64, ALG_DEFAULT,
    MASSolver.cu:245-290) and one unfused MAS-style PCG iteration (MASSolver.cu:469-595). Measurement only: the library
    libpsb200_cmp.so is separate from the product."""
    import ctypes as C
    from polysolve_b200 import problems as P
    path = os.path.join(ROOT, "polysolve_b200", "csrc", "libpsb200_cmp.so")
    L = C.CDLL(path)
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    L.psb200_cmp_run.argtypes = [C.c_int64, C.c_int64, i32p, i32p, f64p, C.c_int, C.c_int, f64p]
    L.psb200_cmp_last_error.restype = C.c_char_p
    out = np.zeros(4)
    nnz = int(outer[-1])
    rc = L.psb200_cmp_run(N, nnz, outer, inner, vals, 50, 50, out)  # symmetric matrix: the CSC arrays are the CSR arrays
    if rc:
        return {"error": L.psb200_cmp_last_error().decode()}
    b_spmv = P.spmv_bytes(N, nnz)
    return {"cusparse_spmv_ms": out[0], "cusparse_spmv_gbs": b_spmv / (out[0] * 1e-3) / 1e9,
            "unfused_iter_ms": out[1], "unfused_iter_per_s": 1e3 / out[1], "cusparse_workspace_bytes": int(out[2]),
            "checksum_sum_A1": out[3],
            "what": "cusparseSpMV CSR fp64 ALG_DEFAULT, 50 reps; MAS-style unfused PCG iteration (cusparseSpMV + 2 atomic dots + "
                    "3 axpby + diagonal apply + 2 scalar kernels), 50 iterations, same matrix, same GPU"}


def run_ours(args):
    import torch
    import polysolve_b200 as psb
    from polysolve_b200 import problems as P

    rank, world, local = dist_env()
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (one process per GPU)")
    torch.cuda.set_device(local)
    # nvidia-smi needs a few hundred ms to deliver its first sample: start it now, filter by time stamps later
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = args.n
    # strong scaling: ONE C2 system, row-range partitioned over the ranks (every rank builds the same
    # synthetic matrix on its host and keeps its own rows)
    N, outer, inner, vals, b, xstar = build_problem(n)
    nnz = int(outer[-1])
    hbm_peak, peak_src = peaks()

    # measured on an 8 x B200 box (profiles/r02_bench_n*.json): the single-reduction form wins from 4 ranks on (+2 % at 4,
    # +8 % at 8) and loses at 2 (-5 %: 12 % more bytes, and synchronisation is not yet the bottleneck there)
    krylov = args.krylov if args.krylov != "auto" else ("cg1r" if world >= 4 else "cg")
    s = psb.Solver.create("CUDA", "")
    s.set_parameters({"CUDA": {"krylov": krylov, "precond": "jacobi", "tolerance": TOL, "max_iter": MAX_ITER,
                               "check_every": args.check_every, "device": local, "interior_first": args.interior_first,
                               "comm_timeout_s": COMM_TIMEOUT_S}})
    if world > 1:
        s.dist_setup_torch(halo_cap=1 << 20)
    t0 = time.perf_counter()
    s.analyze_pattern_raw(N, outer, inner, N)
    t_analyze = time.perf_counter() - t0
    t0 = time.perf_counter()
    s.factorize_raw(N, outer, inner, vals)
    t_factorize = time.perf_counter() - t0
    r0, r1 = s.dist_local_range() if world > 1 else (0, N)
    nl = r1 - r0

    stream = torch.cuda.ExternalStream(s.stream())
    db = torch.from_numpy(b[r0:r1].copy()).cuda()
    dx = torch.zeros(nl, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        with torch.cuda.stream(stream):
            dx.zero_()
        s.solve_device(db.data_ptr(), dx.data_ptr(), nl)
        return s.get_info()["solver_iter"]

    # ---- device-resident timed region
    sampler.mark_load()
    for _ in range(args.warmup):
        device_step()
    launches0 = s.get_info()["gpu_launches"]
    barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    iters = 0
    for _ in range(args.steps):
        iters += device_step()
    e1.record(stream)
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = s.get_info()["gpu_launches"] - launches0
    info = s.get_info()
    xl = dx.cpu().numpy()

    # ---- e2e: host buffers through the C ABI (factorize values + solve), pinned host memory
    hv, hb = pinned_copy(vals), pinned_copy(b)
    hx = pinned_copy(np.zeros(N))
    # x0 = 0: on a row partition a rank reads and writes only its own rows of the full-length x (solve_host), so that is
    # the part it resets
    for _ in range(min(args.warmup, 2)):
        hx[r0:r1].zero_()
        s.factorize_raw(N, outer, inner, hv.numpy())
        s.solve(hb.numpy(), hx.numpy())
    barrier()
    t0 = time.perf_counter()
    e2e_iters = 0
    for _ in range(args.steps):
        hx[r0:r1].zero_()
        s.factorize_raw(N, outer, inner, hv.numpy())
        s.solve(hb.numpy(), hx.numpy())
        e2e_iters += s.get_info()["solver_iter"]
    barrier()
    e2e_dt = time.perf_counter() - t0

    # ---- the other Krylov ordering in the same run (N > 1: Eigen ordering next to the single-reduction headline)
    other = None
    if world > 1:
        other_k = "cg" if krylov == "cg1r" else "cg1r"
        s.set_parameters({"CUDA": {"krylov": other_k}})
        device_step()
        barrier()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record(stream)
        oit = 0
        for _ in range(2):
            oit += device_step()
        o1.record(stream)
        barrier()
        other = (other_k, oit, o0.elapsed_time(o1))
        s.set_parameters({"CUDA": {"krylov": krylov}})

    # ---- roofline of the dominant kernel (fused SpMV + p.Ap), measured live: one extra solve with CUDA
    #      events around every launch on the solver's stream (graphs off), after the timed region.
    s.set_parameters({"CUDA": {"profile": True}})
    device_step()
    prof = s.get_info().get("profile", {})
    s.set_parameters({"CUDA": {"profile": False}})
    plain_ms = s.bench_spmv(reps=50) if world == 1 else 0.0

    # ---- max over ranks; the iteration count is a property of the one shared system (not summed)
    x = np.zeros(N)
    x[r0:r1] = xl
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, e2e_dt, other[2]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_dt = float(t[0]), float(t[1])
        other = (other[0], other[1], float(t[2]))
        c = torch.tensor([launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        launches = int(c[0])
        xt = torch.from_numpy(x).cuda()
        dist.all_reduce(xt, op=dist.ReduceOp.SUM)
        x = xt.cpu().numpy()
    amg_dist = None
    parity = None
    if world > 1:
        # secondary legs: a failure must not cost the headline line; every rank catches on its own and keeps walking
        # through the same barriers
        del s
        try:
            parity = parity_block(psb, P, local, world, rank, barrier)
        except Exception as e:  # noqa: BLE001
            parity = {"ok": False, "error": str(e)[:300]}
        if not args.no_amg:
            try:
                amg_dist = amg_dist_leg(args, psb, P, local, world, N, outer, inner, vals, b, barrier, hbm_peak)
            except Exception as e:  # noqa: BLE001
                amg_dist = {"error": str(e)[:300]}
    if rank != 0:
        return
    rel_res = float(np.linalg.norm(P.spmv_csr(outer, inner, vals, x) - b) / np.linalg.norm(b))
    value = iters / (ms * 1e-3)
    e2e_value = e2e_iters / e2e_dt

    total_prof_ms = sum(v["ms"] for v in prof.values()) or 1.0
    k = prof.get("spmv_cg1r" if krylov == "cg1r" else "spmv_dot", {"ms": 0.0, "launches": 1})
    spmv_ms = k["ms"] / max(1, k["launches"])
    # per-launch algorithmic bytes of THIS rank's kernel (rank 0's row range)
    lnnz = info.get("dist", {}).get("local_nnz", nnz)
    b_spmv = P.spmv_bytes(nl, lnnz)
    achieved = b_spmv / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp) and world == 1:
        try:
            traffic = json.load(open(tp)).get("spmv_dot_dram_bytes_per_launch")
        except Exception:
            traffic = None
    iter_bytes = P.pcg_iter_bytes(N, nnz)
    ms_per_iter = ms / max(1, iters)

    # ---- CPU baseline on a bounded sample (rank 0, N == 1 only)
    cpu = None
    amg = None
    if world == 1 and not args.no_cpu:
        from oracle import oracle as O
        threads = O.lib().orc_num_threads()
        O.eigen_cg(outer, inner, vals, b, tol=TOL, max_iters=MAX_ITER, mode=1, stop_after=2)
        t0 = time.perf_counter()
        _, itc, _, _ = O.eigen_cg(outer, inner, vals, b, tol=TOL, max_iters=MAX_ITER, mode=1, stop_after=args.ref_iters)
        cpu_val = itc / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        _, it1, _, _ = O.eigen_cg(outer, inner, vals, b, tol=TOL, max_iters=MAX_ITER, mode=0, stop_after=max(4, args.ref_iters // 4))
        faithful = it1 / (time.perf_counter() - t0)
        cpu = {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.ref_iters} CG iterations on the full {N}-DoF system, OpenMP row-parallel CSR (oracle/)",
               "eigen_faithful_1thread": faithful}
    cusp = None
    if world == 1 and not args.no_cusparse:
        try:
            cusp = cusparse_leg(args, N, outer, inner, vals)
        except Exception as e:  # noqa: BLE001
            cusp = {"error": str(e)[:300]}
    if world == 1 and not args.no_amg:
        del s
        amg = amg_leg(args, psb, P, local, hbm_peak)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"poisson3d_{n}^3 ({N} DoF, 7-pt), Jacobi-PCG tol {TOL}", "n": N, "nnz": nnz,
                   "partition": f"one system, contiguous row ranges over {world} GPU(s), NVLink peer-memory halo push + fused all-reduce"
                   if world > 1 else "single GPU",
                   "l2_policy": "inputs_exceed_l2 (0.88 GB matrix + 0.4 GB vectors per iteration vs 126 MB L2)"
                   if world <= 2 else "per-GPU working set approaches L2 size at this rank count; no flush (strong scaling of the fixed system)",
                   "iters_per_solve": iters / args.steps, "rel_residual": rel_res,
                   "krylov": krylov + (" (single-reduction CG: one all-reduce + two kernels per iteration; same iterates as the Eigen "
                                       "ordering in exact arithmetic)" if krylov == "cg1r" else " (Eigen::ConjugateGradient ordering)"),
                   "spmv_kernel": info["spmv_kernel"], "check_every": args.check_every,
                   "analyze_s": t_analyze, "factorize_s": t_factorize},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (8 * nnz + 16 * N) // world, "d2h_bytes_per_step": 8 * N // world,
                "what": "x0 = 0 + factorize_csc(host values) + solve(host b, x) per step, pinned host buffers; bytes are per rank"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "spmv_stream_kernel<EpiCg1r> (fused SpMV + 3 dots + all-reduce)" if krylov == "cg1r"
                     else "spmv_stream_kernel<EpiDot> (fused SpMV + p.Ap)", "achieved": achieved,
                     "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": b_spmv, "avg_launch_ms": spmv_ms,
                     "kernel_share_of_step": k["ms"] / total_prof_ms,
                     "plain_spmv_gbs": (b_spmv / (plain_ms * 1e-3) / 1e9) if plain_ms > 0 else None,
                     "pcg_iter_gbs_all_gpus": iter_bytes / (ms_per_iter * 1e-3) / 1e9,
                     "pcg_iter_frac_of_n_gpu_peak": iter_bytes / (ms_per_iter * 1e-3) / 1e9 / (hbm_peak * world),
                     "profile_ms": {kk: vv["ms"] for kk, vv in prof.items()}},
    }
    if cpu:
        line["cpu_baseline"] = cpu
    if amg:
        line["amg_pcg"] = amg
    if amg_dist:
        line["amg_pcg_dist"] = amg_dist
    if parity:
        line["parity"] = parity
    if other:
        line["other_krylov"] = {"krylov": other[0], "value": other[1] / (other[2] * 1e-3), "unit": UNIT, "iters_per_solve": other[1] / 2}
    if cusp:
        line["cusparse"] = cusp
        if "cusparse_spmv_gbs" in cusp:
            line["cusparse"]["ours_vs_cusparse_spmv"] = achieved / cusp["cusparse_spmv_gbs"] if cusp["cusparse_spmv_gbs"] else None
            line["cusparse"]["ours_vs_unfused_iteration"] = value / cusp["unfused_iter_per_s"]
    print(json.dumps(line), flush=True)


def run_c4(args):
    """BASELINE configs[3]: P1-tet linear elasticity on m^3 nodes (m = 119: 5,055,477 DoF, 224 M scalar non-zeros),
    block-3 SA-AMG-PCG (AMGCL_Block<3> semantics, AMGCL.cpp:246-298), rows split over the ranks, tol 1e-8. One JSON line:
    value = setup + solve wall clock (max over ranks), with the hierarchy, the iteration count and the residual
    recomputed on the host."""
    import torch
    import polysolve_b200 as psb
    from polysolve_b200 import problems as P
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hbm_peak, _ = peaks()
    m = args.c4_nodes
    t0 = time.perf_counter()
    outer, inner, vals, b = P.elasticity3d(m)
    t_build = time.perf_counter() - t0
    N = 3 * m ** 3

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    s = psb.Solver.create("CUDA", "")
    s.set_parameters({"CUDA": {"precond": "amg", "block_size": 3, "tolerance": TOL, "max_iter": 1000, "device": local,
                               "comm_timeout_s": COMM_TIMEOUT_S,
                               "amg": {"dist_mode": args.amg_dist_mode, "fused_push": bool(args.fused_push)}}})
    if world > 1:
        s.dist_setup_torch(halo_cap=1 << 21)
    t0 = time.perf_counter()
    s.analyze_pattern_raw(N, outer, inner, N)
    t_analyze = time.perf_counter() - t0
    s.factorize_raw(N, outer, inner, vals)
    setups, solves = [], []
    for _ in range(max(1, args.steps // 3)):
        barrier()
        t0 = time.perf_counter()
        s.factorize_raw(N, outer, inner, vals)
        barrier()
        setups.append(time.perf_counter() - t0)
    hb, hx = pinned_copy(b), pinned_copy(np.zeros(N))
    x = hx.numpy()
    s.solve(hb.numpy(), x)
    for _ in range(max(2, args.steps // 2)):
        x[:] = 0
        barrier()
        t0 = time.perf_counter()
        s.solve(hb.numpy(), x)
        barrier()
        solves.append(time.perf_counter() - t0)
    info = s.get_info()
    s.release_cached_memory()  # setup temporaries cached by the stream-ordered pool are not resident data
    torch.cuda.synchronize()
    used = free0 - torch.cuda.mem_get_info()[0]
    r0, r1 = s.dist_local_range() if world > 1 else (0, N)
    t = torch.tensor([min(setups), min(solves), float(used)], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        xt = torch.zeros(N, dtype=torch.float64, device="cuda")
        xt[r0:r1] = torch.from_numpy(x[r0:r1]).cuda()
        dist.all_reduce(xt, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        x = xt.cpu().numpy()
    if rank != 0:
        return
    rel = float(np.linalg.norm(P.spmv_csr(outer, inner, vals, x) - b) / np.linalg.norm(b))
    amg = info["amg"]
    nnz = int(outer[-1])
    line = {"metric": "amg_pcg_setup_plus_solve_s", "value": float(t[0] + t[1]), "unit": "s", "n_gpus": world, "steps": len(solves), "warmup": 1,
            "ms_per_step": 1e3 * float(t[1]), "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"elasticity3d_{m}^3 nodes ({N} DoF, P1 tets, E=1 nu=0.3, x=0 clamped), block-3 SA-AMG-PCG (AMGCL defaults) tol {TOL}",
                       "n": N, "nnz": nnz, "nnz_blocks": nnz // 9, "dist_mode": info.get("amg_dist_mode"), "host_build_s": t_build, "analyze_s": t_analyze},
            "setup_s": float(t[0]), "solve_s": float(t[1]), "setup_s_all": setups, "solve_s_all": solves,
            "iters": info["num_iterations"], "rel_residual": rel, "levels": [lv["rows"] for lv in amg["levels"]],
            "level_nnz": [lv["nnz"] for lv in amg["levels"]], "operator_complexity": amg["operator_complexity"],
            "partitioned_levels": amg.get("partitioned_levels"), "replicated_levels": amg.get("replicated_levels"),
            "device_bytes_per_rank_max": int(t[2]), "fine_matrix_csr_bytes_total": 12 * nnz + 4 * N,
            "spmv_kernel": info["spmv_kernel"], "roofline": amg_roofline(info, float(t[1]), hbm_peak, world),
            "gpu_launches": info["gpu_launches"]}
    print(json.dumps(line), flush=True)


def run_amg_dist_only(args):
    """N > 1: only the partitioned SA-AMG-PCG leg of config 3 (A/B runs of amg.* switches without the Jacobi legs)."""
    import torch
    import torch.distributed as dist
    import polysolve_b200 as psb
    from polysolve_b200 import problems as P
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, outer, inner, vals, b, _ = build_problem(args.n)
    hbm_peak, _ = peaks()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()
    out = amg_dist_leg(args, psb, P, local, world, N, outer, inner, vals, b, barrier, hbm_peak)
    if rank == 0:
        out["fused_push"] = bool(args.fused_push)
        print(json.dumps({"amg_pcg_dist": out, "n_gpus": world}), flush=True)


def run_c5(args):
    """BASELINE configs[4]: Newton + backtracking on a nonlinear elasticity Problem (compressible Neo-Hookean P1 tets on an
    m^3-node block, m = 70: 1,029,000 DoF, mu = 1, lambda = 1.5, 10 % stretch Dirichlet), inner solve = GPU block-3
    SA-AMG-PCG. The Problem lives on the GPU (include/psb200_problems.h): energy / gradient by CUDA kernels, the Hessian
    assembled on the device straight into the analysed CSC pattern and handed to psb200_factorize_csc_device. On N GPUs the
    linear solvers are row-partitioned (connected through the driver's hook), every rank evaluates the Problem redundantly
    and the Newton step is all-gathered over NVLink."""
    import torch
    import polysolve_b200 as psb
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = args.c5_nodes
    t0 = time.perf_counter()
    prob, x0 = psb.neohookean.stretch_problem(m, stretch=0.1, mu=1.0, lam=1.5, device=local)
    t_build = time.perf_counter() - t0
    nl = {"solver": "Newton", "line_search": {"method": "Backtracking"}, "grad_norm_tol": 1e-8, "rel_grad_norm_tol": 0,
          "max_iterations": 100, "Newton": {"residual_tolerance": 1e-5}}
    lin = {"solver": "CUDA", "CUDA": {"precond": "amg", "block_size": 3, "tolerance": 1e-8, "max_iter": 1000, "device": local,
                                      "comm_timeout_s": COMM_TIMEOUT_S,
                                      "amg": {"dist_mode": args.amg_dist_mode}}}
    runs = []
    for rep in range(1 + max(1, args.steps // 2)):  # the first run warms up (module load, pool growth, graph capture)
        s = psb.NonlinearSolver.create(nl, lin)
        if world > 1:
            s.set_linear_solver_hook(lambda solver: solver.dist_setup_torch(halo_cap=1 << 20))
            import torch.distributed as dist
            dist.barrier()
        x = x0.copy()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.minimize(prob, x)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        info = s.get_info()
        runs.append((dt, info, x))
        del s
    runs = runs[1:]
    dt, info, x = min(runs, key=lambda r: r[0])
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    chk = torch.tensor([float(np.abs(x).sum())], dtype=torch.float64, device="cuda")
    same = True
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(lo[0] == hi[0])  # every rank holds the bit-identical iterate
    if rank != 0:
        return
    inner = [i["solver_iter"] for i in info["internal_solver"]]
    g = prob.gradient(x)
    line = {"metric": "newton_wall_s", "value": float(tt[0]), "unit": "s", "n_gpus": world, "steps": len(runs), "warmup": 0,
            "ms_per_step": 1e3 * float(tt[0]), "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Neo-Hookean P1 tets on {m}^3 nodes ({3 * m ** 3} DoF, mu=1, lambda=1.5, 10% stretch), Newton + Backtracking, "
                                   f"inner = block-3 SA-AMG-PCG tol 1e-8", "n": 3 * m ** 3, "nnz": int(prob.nnz), "problem_build_s": t_build,
                       "hessian": "assembled on the device into the analysed CSC pattern (psb200_factorize_csc_device)",
                       "dist": "row-partitioned linear solvers, Problem evaluated redundantly per rank, Newton step all-gathered over NVLink" if world > 1 else "single GPU"},
            "newton_iterations": info["iterations"], "status": info["status"], "succeeded": info["succeeded"],
            "inner_iterations": inner, "inner_iterations_total": int(sum(inner)), "grad_norm": float(np.linalg.norm(g)), "energy": info["energy"],
            "time_assembly_s": info.get("time_assembly"), "time_linear_s": info.get("time_inverting"), "time_line_search_s": info.get("time_line_search"),
            "time_grad_s": info.get("time_grad"), "amg_levels": [lv["rows"] for lv in info["internal_solver"][-1]["amg"]["levels"]],
            "amg_dist_mode": info["internal_solver"][-1].get("amg_dist_mode"), "iterate_identical_on_all_ranks": same,
            "all_runs_s": [r[0] for r in runs]}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=216, help="grid points per side (216 -> 10,077,696 DoF)")
    ap.add_argument("--check-every", type=int, default=16)
    ap.add_argument("--ref-iters", type=int, default=40, help="CG iterations per CPU sample step")
    ap.add_argument("--krylov", default="auto", choices=["auto", "cg", "cg1r"], help="auto: cg (Eigen ordering) on 1-2 GPUs, cg1r (single reduction) from 4 ranks")
    ap.add_argument("--no-cusparse", action="store_true", help="skip the cuSPARSE / MAS-style comparator (N = 1)")
    ap.add_argument("--amg-dist-mode", default="partitioned", choices=["partitioned", "global", "local"])
    ap.add_argument("--config", default="c2", choices=["c2", "c4", "c5"], help="c2: the headline (10M-DoF Poisson Jacobi-PCG + the C3 AMG legs); c4: 119^3-node elasticity, block-3 AMG-PCG; c5: Newton on 70^3-node Neo-Hookean")
    ap.add_argument("--c4-nodes", type=int, default=119)
    ap.add_argument("--c5-nodes", type=int, default=70)
    ap.add_argument("--fused-push", action="store_true", help="multi-GPU AMG legs: amg.fused_push = true (halo pushed from the SpMV epilogue)")
    ap.add_argument("--amg-dist-only", action="store_true", help="N > 1: run only the partitioned SA-AMG-PCG leg (config 3) and print its JSON")
    ap.add_argument("--interior-first", action="store_true", help="row partitions: SpMV tiles without halo columns first, late halo wait")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-amg", action="store_true", help="skip the AMG-PCG (config 3) leg")
    ap.add_argument("--amg-cpu-n", type=int, default=0, help="grid side of the CPU AMG leg (0 = the full --n system)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.amg_dist_only:
        run_amg_dist_only(args)
    elif args.config == "c4":
        run_c4(args)
    elif args.config == "c5":
        run_c5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
