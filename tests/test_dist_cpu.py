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
