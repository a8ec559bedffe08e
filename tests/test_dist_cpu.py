"""CPU tests of the multi-GPU host logic (row partition, halo lists, column remap, send lists):
bit-exact against the oracle's index functions, plus a world_size-2 gloo run that emulates the halo
exchange with torch.distributed and checks the distributed SpMV against the full one."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

HALO_CAP = 4096


def matrices(orc):
    rng = np.random.default_rng(5)
    out = {"poisson3d_12": orc.poisson3d(12), "poisson2d_40": orc.poisson2d(40), "convdiff_24": orc.convdiff2d(24, 0.3)}
    A = sp.random(500, 500, density=0.02, random_state=rng, format="csc") + sp.eye(500, format="csc")
    A = sp.csc_matrix(A)
    A.sort_indices()
    out["random_unsym"] = (A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64))
    return out


def compact_cols(plan, world):
    """product layout (nl + q*cap + pos) -> the oracle's compact layout (nl + position in the halo list)"""
    nl = plan["n_local"]
    seg = np.concatenate([[0], np.cumsum(plan["recv_count"])])
    ci = plan["ci"].astype(np.int64)
    out = ci.copy()
    h = ci >= nl
    q = (ci[h] - nl) // HALO_CAP
    pos = (ci[h] - nl) % HALO_CAP
    out[h] = nl + seg[q] + pos
    return out


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_plan_matches_oracle_bit_exact(psb, orc, world):
    for name, (o, i, v) in matrices(orc).items():
        n = len(o) - 1
        rp, ci, perm = orc.csc_to_csr(n, o, i)
        off0 = orc.partition_rows(rp, world)
        plans = [psb.Solver.dist_plan_host(n, o, i, r, world, HALO_CAP) for r in range(world)]
        for r, P in enumerate(plans):
            assert np.array_equal(P["offsets"], off0), name
            a, b = int(off0[r]), int(off0[r + 1])
            lc0, halo0 = orc.halo_for_rank(rp, ci, a, b)
            assert P["n_local"] == b - a
            assert np.array_equal(P["rp"], rp[a:b + 1] - rp[a]), name
            assert np.array_equal(P["halo_cols"], halo0), name
            assert np.array_equal(compact_cols(P, world), lc0), name
            assert np.array_equal(P["perm"], perm[rp[a]:rp[b]]), name  # values map: vals_local = vals_csc[perm]
        # what g sends to q is exactly what q expects from g, in the same (ascending) order
        for g, Pg in enumerate(plans):
            for q, Pq in enumerate(plans):
                sent = Pg["send_rows"][Pg["send_begin"][q]:Pg["send_begin"][q + 1]].astype(np.int64) + off0[g]
                want = Pq["halo_cols"][(Pq["halo_cols"] >= off0[g]) & (Pq["halo_cols"] < off0[g + 1])]
                assert np.array_equal(sent, want), (name, g, q)
                assert len(sent) == Pq["recv_count"][g]
                if g == q:
                    assert len(sent) == 0


@pytest.mark.parametrize("world", [2, 3, 8])
def test_plan_block_aligned_offsets(psb, orc, world):
    """Block problems (set_block_size(3)): offsets are multiples of the block size, equal to the oracle's
    aligned partition, and every other index array still matches the oracle for those offsets."""
    o, i, v, _ = psb.problems.elasticity3d(6)
    n = len(o) - 1
    rp, ci, perm = orc.csc_to_csr(n, o, i)
    off0 = orc.partition_rows(rp, world, align=3)
    assert np.all(off0 % 3 == 0)
    for r in range(world):
        P = psb.Solver.dist_plan_host(n, o, i, r, world, HALO_CAP, align=3)
        assert np.array_equal(P["offsets"], off0)
        a, b = int(off0[r]), int(off0[r + 1])
        lc0, halo0 = orc.halo_for_rank(rp, ci, a, b)
        assert np.array_equal(P["rp"], rp[a:b + 1] - rp[a])
        assert np.array_equal(P["halo_cols"], halo0)
        assert np.array_equal(compact_cols(P, world), lc0)
        assert np.array_equal(P["perm"], perm[rp[a]:rp[b]])
    with pytest.raises(RuntimeError):
        psb.Solver.dist_plan_host(n, o, i, 0, world, HALO_CAP, align=7)  # 648 is not a multiple of 7


@pytest.mark.parametrize("world", [2, 3, 4])
def test_plan_block_halos_are_whole_nodes(psb, orc, world):
    """Block problems exchange whole nodes: with a pattern whose 3 x 3 blocks are incomplete (entries dropped at random, so
    a rank may read only one dof of a remote node) the halo list of every rank is the node-completion of the oracle's halo
    list, the send lists are whole nodes too, and what g sends to q is exactly what q expects -- the invariant the
    partitioned block AMG needs to expand its rows to full blocks including the halo columns."""
    rng = np.random.default_rng(7)
    o, i, v, _ = psb.problems.elasticity3d(5)
    n = len(o) - 1
    A = sp.csc_matrix((v, i, o), shape=(n, n)).tocoo()
    keep = (rng.random(A.nnz) < 0.6) | (A.row == A.col)
    A = sp.csc_matrix((A.data[keep], (A.row[keep], A.col[keep])), shape=(n, n))
    A.sort_indices()
    o, i = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    rp, ci, perm = orc.csc_to_csr(n, o, i)
    off0 = orc.partition_rows(rp, world, align=3)
    plans = [psb.Solver.dist_plan_host(n, o, i, r, world, HALO_CAP, align=3) for r in range(world)]
    for r, P in enumerate(plans):
        a, b = int(off0[r]), int(off0[r + 1])
        _, halo0 = orc.halo_for_rank(rp, ci, a, b)
        want = np.unique((halo0[:, None] // 3 * 3 + np.arange(3)[None, :]).reshape(-1)) if len(halo0) else halo0
        assert np.array_equal(P["halo_cols"], want)
        assert len(P["halo_cols"]) % 3 == 0 and np.all(P["halo_cols"].reshape(-1, 3) % 3 == np.arange(3))
        assert np.array_equal(P["perm"], perm[rp[a]:rp[b]])           # the matrix entries themselves are untouched
    for g, Pg in enumerate(plans):
        for q, Pq in enumerate(plans):
            sent = Pg["send_rows"][Pg["send_begin"][q]:Pg["send_begin"][q + 1]].astype(np.int64) + off0[g]
            want = Pq["halo_cols"][(Pq["halo_cols"] >= off0[g]) & (Pq["halo_cols"] < off0[g + 1])]
            # the sender cannot know which of its columns the reader's ROWS touch beyond the pattern it sees: it sends the
            # whole nodes of every column with an entry in a row of q -- exactly the reader's completed list
            assert np.array_equal(sent, want), (g, q)


def test_plan_edge_cases(psb, orc):
    """More ranks than rows (empty ranks), no coupling at all (no halo anywhere), one dense row and column (every rank
    needs row 0, rank 0 needs everybody): offsets, halo lists and send lists still equal the oracle's."""
    B = sp.lil_matrix((50, 50))
    B.setdiag(1.0)
    B[0, :] = 1.0
    B[:, 0] = 1.0
    cases = [("diag5", sp.identity(5, format="csc"), 8), ("one", sp.identity(1, format="csc"), 2),
             ("tri3", sp.diags([[-1.0] * 2, [2.0] * 3, [-1.0] * 2], [-1, 0, 1], format="csc"), 4), ("arrow", B.tocsc(), 4)]
    for name, A, world in cases:
        A = sp.csc_matrix(A)
        A.sort_indices()
        n = A.shape[0]
        o, i = A.indptr.astype(np.int32), A.indices.astype(np.int32)
        rp, ci, perm = orc.csc_to_csr(n, o, i)
        off0 = orc.partition_rows(rp, world)
        plans = [psb.Solver.dist_plan_host(n, o, i, r, world, HALO_CAP) for r in range(world)]
        for r, P in enumerate(plans):
            a, b = int(off0[r]), int(off0[r + 1])
            lc0, halo0 = orc.halo_for_rank(rp, ci, a, b)
            assert np.array_equal(P["offsets"], off0) and P["n_local"] == b - a, name
            assert np.array_equal(P["halo_cols"], halo0) and np.array_equal(compact_cols(P, world), lc0), name
        for g, Pg in enumerate(plans):
            for q, Pq in enumerate(plans):
                sent = Pg["send_rows"][Pg["send_begin"][q]:Pg["send_begin"][q + 1]].astype(np.int64) + off0[g]
                want = Pq["halo_cols"][(Pq["halo_cols"] >= off0[g]) & (Pq["halo_cols"] < off0[g + 1])]
                assert np.array_equal(sent, want), (name, g, q)
    assert sum(P["n_local"] == 0 for P in [psb.Solver.dist_plan_host(5, *[sp.identity(5, format="csc").indptr.astype(np.int32),
               sp.identity(5, format="csc").indices.astype(np.int32)], r, 8, HALO_CAP) for r in range(8)]) == 3


def test_plan_randomised_matrices_and_rank_counts(psb, orc):
    """120 random square patterns (empty to half dense, with and without a diagonal) on 1-8 ranks: every index array equals
    the oracle's, send lists equal the readers' halo lists. Dense couplings send a row to several ranks, so the send list
    of a rank can be longer than n (this once overflowed the wrapper's n-entry buffer)."""
    import random
    rnd = random.Random(5)
    rng = np.random.default_rng(5)
    longest = 0
    for t in range(120):
        n = rnd.randint(1, 60)
        world = rnd.randint(1, 8)
        A = sp.random(n, n, density=rnd.choice([0, 0.02, 0.1, 0.5]), random_state=rng, format="csc")
        if rnd.random() < 0.7:
            A = A + sp.identity(n, format="csc")
        A = sp.csc_matrix(A)
        A.sort_indices()
        o, i = A.indptr.astype(np.int32), A.indices.astype(np.int32)
        rp, ci, perm = orc.csc_to_csr(n, o, i)
        off0 = orc.partition_rows(rp, world)
        plans = [psb.Solver.dist_plan_host(n, o, i, r, world, HALO_CAP) for r in range(world)]
        for r, P in enumerate(plans):
            a, b = int(off0[r]), int(off0[r + 1])
            lc0, halo0 = orc.halo_for_rank(rp, ci, a, b)
            assert np.array_equal(P["offsets"], off0) and np.array_equal(P["rp"], rp[a:b + 1] - rp[a]), t
            assert np.array_equal(P["halo_cols"], halo0) and np.array_equal(compact_cols(P, world), lc0), t
            assert np.array_equal(P["perm"], perm[rp[a]:rp[b]]), t
            longest = max(longest, len(P["send_rows"]) / n)
        for g, Pg in enumerate(plans):
            for q, Pq in enumerate(plans):
                sent = Pg["send_rows"][Pg["send_begin"][q]:Pg["send_begin"][q + 1]].astype(np.int64) + off0[g]
                want = Pq["halo_cols"][(Pq["halo_cols"] >= off0[g]) & (Pq["halo_cols"] < off0[g + 1])]
                assert np.array_equal(sent, want), (t, g, q)
    assert longest > 1.0   # the case the n-entry buffer could not hold is exercised


def test_plan_rejects_corrupted_index_arrays(psb):
    """The host-only plan entry point validates the pattern itself (outer monotone from 0 to nnz, inner in range): 600
    corrupted inputs come back as errors -- a non-monotone outer array once corrupted the heap."""
    import random
    rnd = random.Random(9)
    rng = np.random.default_rng(9)
    ok = err = 0
    for _ in range(600):
        n = rnd.randint(1, 40)
        world = rnd.randint(1, 8)
        A = sp.csc_matrix(sp.random(n, n, density=rnd.choice([0.05, 0.3]), random_state=rng, format="csc") + sp.identity(n, format="csc"))
        A.sort_indices()
        o, i = A.indptr.astype(np.int32).copy(), A.indices.astype(np.int32).copy()
        m = rnd.random()
        corrupted = False
        if m < 0.35:
            i[rnd.randrange(len(i))] = rnd.choice([-1, n, n + 5, 2 ** 31 - 1, -2 ** 31])
            corrupted = True
        elif m < 0.7:
            k = rnd.randrange(len(o))
            new = rnd.choice([-1, len(i) + 3, 2 ** 31 - 1])
            lo = o[k - 1] if k > 0 else 0
            hi = o[k + 1] if k + 1 < len(o) else new - 1
            corrupted = not (lo <= new <= hi) or k == 0 or k == n
            o[k] = new
        try:
            for r in range(world):
                psb.Solver.dist_plan_host(n, o, i, r, world, HALO_CAP)
            ok += 1
            assert not corrupted
        except RuntimeError:
            err += 1
            assert corrupted
    assert ok > 100 and err > 100


def test_plan_threaded_path_matches_oracle(psb, orc):
    """Above 2^20 entries the plan builder splits its passes over host threads (column ranges dealt in order): a
    2.6 M-entry matrix on 3 and 8 ranks still equals the oracle bit for bit, and repeated builds are identical."""
    o, i, v = orc.poisson3d(72)
    n = 72 ** 3
    assert int(o[-1]) > 2 * (1 << 20)
    rp, ci, perm = orc.csc_to_csr(n, o, i)
    for world in (3, 8):
        off0 = orc.partition_rows(rp, world)
        for r in (0, world // 2, world - 1):
            P = psb.Solver.dist_plan_host(n, o, i, r, world, 1 << 16)
            a, b = int(off0[r]), int(off0[r + 1])
            lc0, halo0 = orc.halo_for_rank(rp, ci, a, b)
            assert np.array_equal(P["offsets"], off0) and np.array_equal(P["rp"], rp[a:b + 1] - rp[a])
            assert np.array_equal(P["halo_cols"], halo0) and np.array_equal(P["perm"], perm[rp[a]:rp[b]])
            Q = psb.Solver.dist_plan_host(n, o, i, r, world, 1 << 16)
            assert all(np.array_equal(P[k], Q[k]) for k in ("ci", "perm", "send_rows", "send_begin", "recv_count", "halo_cols"))


def test_plan_halo_capacity_error(psb, orc):
    o, i, v = orc.poisson3d(12)
    with pytest.raises(RuntimeError):
        psb.Solver.dist_plan_host(12 ** 3, o, i, 0, 2, halo_cap=16)  # a 12x12 plane does not fit in 16 slots


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    import polysolve_b200 as psb
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o, i, v = orc.convdiff2d(30, 0.4)  # unsymmetric values, symmetric pattern
        n = 900
        x = orc.splitmix64(9, n)
        P = psb.Solver.dist_plan_host(n, o, i, rank, world, HALO_CAP)
        a, b = int(P["offsets"][rank]), int(P["offsets"][rank + 1])
        nl = b - a
        # "halo push" emulated with an all-gather of (destination, values) pairs
        outbox = {dst: x[a:b][P["send_rows"][P["send_begin"][dst]:P["send_begin"][dst + 1]]] for dst in range(world)}
        boxes = [None] * world
        dist.all_gather_object(boxes, outbox)
        xext = np.zeros(nl + world * HALO_CAP)
        xext[:nl] = x[a:b]
        for src in range(world):
            vals = boxes[src][rank]
            assert len(vals) == P["recv_count"][src]
            xext[nl + src * HALO_CAP: nl + src * HALO_CAP + len(vals)] = vals
        vloc = v[P["perm"]]
        y_loc = np.add.reduceat(vloc * xext[P["ci"]], P["rp"][:-1].astype(np.int64))
        y_full = sp.csc_matrix((v, i, o), shape=(n, n)) @ x
        err = float(np.abs(y_loc - y_full[a:b]).max())
        # dot-product all-reduce in rank order (what comm_allreduce does on the device)
        parts = [None] * world
        dist.all_gather_object(parts, float(y_loc @ x[a:b]))
        total = sum(parts)
        q.put((rank, err, abs(total - float(y_full @ x))))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_distributed_spmv(psb, orc):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    world = 2
    port = _free_port()
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, derr in res:
        assert err < 1e-13, (rank, err)
        assert derr < 1e-10


def _gloo_cg1r_worker(rank, world, port, q):
    """Row-partitioned single-reduction PCG as `krylov = cg1r` runs it (dist.cu: cg1r_update_kernel pushes the halo of the
    new u, the SpMV epilogue reduces gamma, delta, |r|^2 in ONE all-reduce), restated over gloo on the product's host
    plan: per trip ONE halo exchange and ONE all-reduce of three doubles."""
    import torch
    import torch.distributed as dist

    import polysolve_b200 as psb
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o, i, v = orc.poisson3d(10)
        n = 1000
        v = v * (1.0 + 0.05 * orc.splitmix64(13, len(v)))
        v = 0.5 * (v + v[orc.csc_to_csr(n, o, i)[2]])
        b = orc.splitmix64(42, n)
        tol, max_iters = 1e-10, 1000
        P = psb.Solver.dist_plan_host(n, o, i, rank, world, HALO_CAP)
        a, e = int(P["offsets"][rank]), int(P["offsets"][rank + 1])
        nl = e - a
        vloc = v[P["perm"]]
        rp = P["rp"][:-1].astype(np.int64)
        diag = sp.csc_matrix((v, i, o), shape=(n, n)).diagonal()[a:e]
        dinv = 1.0 / diag
        n_exchanges = n_reductions = 0

        def spmv(u_loc):
            nonlocal n_exchanges
            outbox = {dst: u_loc[P["send_rows"][P["send_begin"][dst]:P["send_begin"][dst + 1]]] for dst in range(world)}
            boxes = [None] * world
            dist.all_gather_object(boxes, outbox)
            n_exchanges += 1
            ext = np.zeros(nl + world * HALO_CAP)
            ext[:nl] = u_loc
            for src in range(world):
                vals = boxes[src][rank]
                ext[nl + src * HALO_CAP: nl + src * HALO_CAP + len(vals)] = vals
            return np.add.reduceat(vloc * ext[P["ci"]], rp)

        def allreduce(*vals):
            nonlocal n_reductions
            t = torch.tensor(vals, dtype=torch.float64)
            dist.all_reduce(t)
            n_reductions += 1
            return [float(z) for z in t]

        bl = b[a:e]
        x = np.zeros(nl)
        r = bl - spmv(x)
        rn2, bn2 = allreduce(float(r @ r), float(bl @ bl))
        thr = tol * tol * bn2
        p = np.zeros(nl)
        s = np.zeros(nl)
        u = dinv * r
        it = 0
        gamma_old = alpha_old = None
        ex0, red0 = n_exchanges, n_reductions
        trips = 0
        while True:
            w = spmv(u)
            gamma, delta, rn2 = allreduce(float(r @ u), float(w @ u), float(r @ r))
            trips += 1
            if gamma_old is not None and rn2 < thr:
                break
            if gamma_old is None:
                beta, alpha = 0.0, gamma / delta
            else:
                beta = gamma / gamma_old
                alpha = gamma / (delta - beta * gamma / alpha_old)
                it += 1
                if it >= max_iters:
                    break
            gamma_old, alpha_old = gamma, alpha
            p = u + beta * p
            s = w + beta * s
            x = x + alpha * p
            r = r - alpha * s
            u = dinv * r
        q.put((rank, a, e, x, it, float(np.sqrt(rn2 / bn2)), n_exchanges - ex0, n_reductions - red0, trips))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_single_reduction_cg(psb, orc):
    """The partitioned cg1r equals the single-process restatement (oracle/cg1r_oracle.py) and the Eigen-ordering oracle:
    same iteration count (+-1 for the different summation order), same x; one halo exchange and one all-reduce per trip."""
    import torch.multiprocessing as mp
    from oracle import cg1r_oracle
    ctx = mp.get_context("spawn")
    world = 2
    port = _free_port()
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_cg1r_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o, i, v = orc.poisson3d(10)
    n = 1000
    v = v * (1.0 + 0.05 * orc.splitmix64(13, len(v)))
    v = 0.5 * (v + v[orc.csc_to_csr(n, o, i)[2]])
    b = orc.splitmix64(42, n)
    A = sp.csc_matrix((v, i, o), shape=(n, n)).tocsr()
    x1, it1, err1, _ = cg1r_oracle.cg1r(A, b, dinv=1.0 / A.diagonal(), tol=1e-10, max_iters=1000)
    x0, it0, _, _ = orc.eigen_cg(o, i, v, b, tol=1e-10, max_iters=1000)
    x = np.zeros(n)
    for rank, a, e, xl, it, err, nex, nred, trips in res:
        x[a:e] = xl
        assert abs(it - it1) <= 1 and abs(it - it0) <= 1 and err < 1e-10
        assert nex == trips and nred == trips            # ONE exchange and ONE reduction per trip
    assert len({r[4] for r in res}) == 1                 # every rank took the same decisions
    assert np.linalg.norm(x - x1) / np.linalg.norm(x1) < 1e-9
    assert np.linalg.norm(x - x0) / np.linalg.norm(x0) < 1e-9
