"""GPU tests of the row-partitioned Jacobi-PCG (NVLink peer-memory halo push + fused all-reduce).
The world_size-1 case runs on any GPU box; the multi-rank cases need >= 2 GPUs (one process per GPU,
rendezvous over gloo at 127.0.0.1) and are skipped otherwise.  N-rank result == 1-rank result == oracle."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(P, n, block):
    if block > 1:
        o, i, v, b = P.elasticity3d(n)
        return o, i, v, b, 3 * n ** 3
    o, i, v = P.poisson3d(n)
    N = n ** 3
    return o, i, v, P.spmv_csr(o, i, v, P.splitmix64(42, N)), N


def _worker(rank, world, port, n, tol, q, precond="jacobi", block=1, amg_mode="global", extra=None):
    import torch
    import torch.distributed as dist

    import polysolve_b200 as psb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = psb.problems
        o, i, v, b, N = _problem(P, n, block)
        s = psb.Solver.create("CUDA", "")
        prm = {"tolerance": tol, "max_iter": 10000, "device": rank, "check_every": 8, "precond": precond,
               "block_size": block, "amg": {"dist_mode": amg_mode}}
        prm.update(extra or {})
        s.set_parameters({"CUDA": prm})
        s.dist_setup_torch(halo_cap=1 << 16)
        s.analyze_pattern_raw(N, o, i, N)
        s.factorize_raw(N, o, i, v)
        a, e = s.dist_local_range()
        x = np.zeros(N)
        s.solve(b, x)
        info = s.get_info()
        # second solve from the converged iterate: 0 iterations on every rank
        x2 = x.copy()
        s.solve(b, x2)
        it2 = s.get_info()["solver_iter"]
        q.put((rank, a, e, x[a:e].copy(), info["solver_iter"], info["solver_error"], info["solver_status"], it2, info["dist"]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run(world, n, tol, precond="jacobi", block=1, amg_mode="global", extra=None):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    port = _free_port()
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, tol, q, precond, block, amg_mode, extra)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return sorted(res, key=lambda t: t[0])


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_dist_pcg_matches_oracle(orc, world):
    if world > max(1, _ngpu()):
        pytest.skip(f"needs {world} GPUs")
    n, tol = 40, 1e-8
    N = n ** 3
    o, i, v = orc.poisson3d(n)
    b = orc.spmv_csc(o, i, v, orc.splitmix64(42, N))
    x0, it0, err0, _ = orc.eigen_cg(o, i, v, b, tol=tol, max_iters=10000)
    res = _run(world, n, tol)
    rp, ci, _ = orc.csc_to_csr(N, o, i)
    off0 = orc.partition_rows(rp, world)
    x = np.zeros(N)
    for rank, a, e, xs, it, err, status, it2, dinfo in res:
        assert (a, e) == (off0[rank], off0[rank + 1])  # partition offsets bit-exact vs the oracle
        x[a:e] = xs
        assert status == "Converged"
        assert it == res[0][4]                           # every rank reports the same iteration count
        assert abs(it - it0) <= max(1, 0.02 * it0)       # same algorithm, different summation order
        assert err < tol
        assert it2 == 0
        assert dinfo["world"] == world
    assert np.linalg.norm(x - x0) / np.linalg.norm(x0) < 10 * tol
    assert np.linalg.norm(orc.spmv_csc(o, i, v, x) - b) / np.linalg.norm(b) < 2 * tol


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("extra", [{"cg_kernel": "persistent"}, {"cg_kernel": "persistent", "interior_first": True},
                                   {"interior_first": True}])
def test_dist_pcg_kernel_variants(orc, world, extra):
    """The persistent cooperative kernel and the interior-first tile order on the row partition: same result as the
    default path (oracle iteration count within the summation-order band, solution to the solver tolerance)."""
    if world > max(1, _ngpu()):
        pytest.skip(f"needs {world} GPUs")
    n, tol = 40, 1e-8
    N = n ** 3
    o, i, v = orc.poisson3d(n)
    b = orc.spmv_csc(o, i, v, orc.splitmix64(42, N))
    x0, it0, _, _ = orc.eigen_cg(o, i, v, b, tol=tol, max_iters=10000)
    res = _run(world, n, tol, extra=extra)
    x = np.zeros(N)
    for rank, a, e, xs, it, err, status, it2, dinfo in res:
        x[a:e] = xs
        assert status == "Converged" and it == res[0][4]
        assert abs(it - it0) <= max(1, 0.02 * it0)
        assert err < tol and it2 == 0
    assert np.linalg.norm(x - x0) / np.linalg.norm(x0) < 10 * tol


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_dist_amg_pcg_global_hierarchy(psb, orc, world):
    """AMG-PCG on the row partition, amg.dist_mode = global (default): the hierarchy of the WHOLE matrix, level 0 of the
    cycle partitioned (halo push per smoothing step, restriction summed across ranks), coarse levels replicated. Same
    hierarchy and arithmetic as the single-GPU solver up to summation order: same iteration count, same solution."""
    if world > max(1, _ngpu()):
        pytest.skip(f"needs {world} GPUs")
    n, tol = 40, 1e-8
    N = n ** 3
    o, i, v = orc.poisson3d(n)
    b = orc.spmv_csc(o, i, v, orc.splitmix64(42, N))
    s1 = psb.Solver.create("CUDA", "")
    s1.set_parameters({"CUDA": {"tolerance": tol, "max_iter": 1000, "precond": "amg"}})
    s1.factorize_raw(N, o, i, v)
    x1 = np.zeros(N)
    s1.solve(b, x1)
    it1 = s1.get_info()["solver_iter"]
    levels1 = [lv["rows"] for lv in s1.get_info()["amg"]["levels"]]
    del s1
    res = _run(world, n, tol, "amg")
    x = np.zeros(N)
    for rank, a, e, xs, it, err, status, it2, dinfo in res:
        x[a:e] = xs
        assert status == "Converged"
        assert it == res[0][4] == it1, (it, it1)     # the 1-GPU iteration count, on every rank
        assert err < tol
        assert it2 == 0
    assert np.linalg.norm(x - x1) / np.linalg.norm(x1) < 1e-9   # same iterates up to summation order
    assert np.linalg.norm(orc.spmv_csc(o, i, v, x) - b) / np.linalg.norm(b) < 2 * tol
    assert len(levels1) >= 2


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_dist_amg_pcg_rank_local(orc, world):
    """amg.dist_mode = local: rank-local SA-AMG of the diagonal block (block-Jacobi across ranks) inside the global CG.
    The solution equals the oracle's to the solver tolerance; the iteration count grows with the rank count."""
    if world > max(1, _ngpu()):
        pytest.skip(f"needs {world} GPUs")
    n, tol = 40, 1e-8
    N = n ** 3
    o, i, v = orc.poisson3d(n)
    b = orc.spmv_csc(o, i, v, orc.splitmix64(42, N))
    x0, _, _, _ = orc.eigen_cg(o, i, v, b, tol=1e-12, max_iters=10000)
    res = _run(world, n, tol, "amg", 1, "local")
    x = np.zeros(N)
    for rank, a, e, xs, it, err, status, it2, dinfo in res:
        x[a:e] = xs
        assert status == "Converged"
        assert it == res[0][4] and 1 <= it <= 80
        assert err < tol
        assert it2 == 0
    # a different preconditioner stops at a different point inside the tolerance ball: error <= cond(A) * residual
    assert np.linalg.norm(x - x0) / np.linalg.norm(x0) < 1e-6
    assert np.linalg.norm(orc.spmv_csc(o, i, v, x) - b) / np.linalg.norm(b) < 2 * tol


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_dist_block3_amg_pcg_elasticity(psb, orc, world):
    """C4-shaped (BASELINE config 4): P1 linear elasticity, block-3 SA-AMG-PCG on the row partition. Offsets are
    multiples of 3 and equal the oracle's aligned partition; the solution equals a direct solve."""
    if world > max(1, _ngpu()):
        pytest.skip(f"needs {world} GPUs")
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    m, tol = 16, 1e-8
    o, i, v, b = psb.problems.elasticity3d(m)
    N = 3 * m ** 3
    A = sp.csc_matrix((v, i, o), shape=(N, N))
    x0 = spla.spsolve(A, b)
    res = _run(world, m, tol, "amg", 3)
    rp, ci, _ = orc.csc_to_csr(N, o, i)
    off0 = orc.partition_rows(rp, world, align=3)
    x = np.zeros(N)
    for rank, a, e, xs, it, err, status, it2, dinfo in res:
        assert (a, e) == (off0[rank], off0[rank + 1]) and a % 3 == 0
        x[a:e] = xs
        assert status == "Converged"
        assert it == res[0][4] and 1 <= it <= 200   # rank-local hierarchy: grows with the rank count (83 on 4 ranks)
        assert err < tol
        assert it2 == 0
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) < 2 * tol
    assert np.linalg.norm(x - x0) / np.linalg.norm(x0) < 1e-5


def test_two_devices_in_one_process(psb, orc):
    """Solver instances on different GPUs driven from one host thread (every C-ABI call makes its device current)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    o, i, v = orc.poisson3d(20)
    N = 20 ** 3
    b = orc.splitmix64(3, N)
    ss = []
    for dev in (0, 1):
        s = psb.Solver.create("CUDA", "")
        s.set_parameters({"CUDA": {"tolerance": 1e-10, "device": dev}})
        ss.append(s)
    for s in ss:              # interleaved: analyze both, then factorize both, then solve both
        s.analyze_pattern_raw(N, o, i, N)
    for s in ss:
        s.factorize_raw(N, o, i, v)
    xs = []
    for s in reversed(ss):
        x = np.zeros(N)
        s.solve(b, x)
        xs.append(x)
    assert np.array_equal(xs[0], xs[1])
    assert np.linalg.norm(orc.spmv_csc(o, i, v, xs[0]) - b) < 1e-8
