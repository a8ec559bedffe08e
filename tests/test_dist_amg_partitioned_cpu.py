"""CPU blueprint of the PARTITIONED multi-GPU AMG setup (polysolve_b200/csrc/amg_dist.cu), run by two
gloo ranks exactly as the GPUs do it and checked against the single-process algebra:

  * decoupled aggregation: aggregates of a rank use only its own rows (here: greedy aggregation of the diagonal block);
  * rank-local P: smoothed with the filtered matrix (connections leaving the rank lumped into the diagonal), so P is block
    diagonal over the ranks, constants stay in its range, R = P^T is rank-local;
  * distributed Galerkin product: ONE exchange of the P rows of the halo columns, then A_c[my aggregates, :] = R (A P_ext)
    with global coarse column ids;
  * next-level halo plan from the owners' point of view: the consumers' requests give the send lists.

Checked: the row blocks the two ranks compute are exactly the rows of P^T A P of the single process with P = blockdiag(P_r);
P 1 = 1 on rows whose filtered diagonal is regular; the requested / sent lists of the next level match."""
import os
import socket

import numpy as np
import scipy.sparse as sp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def greedy_aggregates(A):
    """plain aggregation on the pattern of A (any deterministic aggregation does for the algebra checked here)"""
    n = A.shape[0]
    agg = -np.ones(n, int)
    k = 0
    for i in range(n):
        if agg[i] >= 0:
            continue
        nb = A.indices[A.indptr[i]:A.indptr[i + 1]]
        free = [j for j in nb if agg[j] < 0]
        if len(free) <= 1 and len(nb) > 1 and any(agg[j] >= 0 for j in nb):
            agg[i] = agg[[j for j in nb if agg[j] >= 0][0]]
            continue
        agg[free] = k
        agg[i] = k
        k += 1
    return agg, k


def local_prolongation(A_loc, a, b, omega=2.0 / 3.0):
    """P_r = (I - omega D_f^-1 A_f) P_tent on the rows [a, b) of one rank: A_f = diagonal block with the entries that leave
    the rank lumped into the diagonal (amg_dist.cu lump_halo_kernel + build_prolongation)."""
    nl = b - a
    A_loc = A_loc.tocsr()
    D = A_loc[:, a:b].tolil()
    lump = np.asarray(A_loc.sum(axis=1)).ravel() - np.asarray(A_loc[:, a:b].sum(axis=1)).ravel()
    D.setdiag(D.diagonal() + lump)
    Af = D.tocsr()
    agg, na = greedy_aggregates(Af)
    Pt = sp.csr_matrix((np.ones(nl), (np.arange(nl), agg)), shape=(nl, na))
    S = sp.identity(nl, format="csr") - omega * sp.diags(1.0 / Af.diagonal()) @ Af
    return (S @ Pt).tocsr(), na


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 10
        o, i, v = orc.poisson3d(n)
        N = n ** 3
        A = sp.csr_matrix((v, i, o), shape=(N, N))  # symmetric: CSC arrays == CSR arrays
        rp = A.indptr.astype(np.int32)
        off = orc.partition_rows(rp, world)
        a, b = int(off[rank]), int(off[rank + 1])
        A_loc = A[a:b, :]
        P_loc, na = local_prolongation(A_loc, a, b)
        # coarse offsets = prefix sums of the aggregate counts (dist_gather_ll)
        cnt = [None] * world
        dist.all_gather_object(cnt, na)
        coff = np.concatenate([[0], np.cumsum(cnt)])
        # halo columns of my rows and the exchange of their P rows (lengths + (global column, value) pairs)
        cols = np.unique(A_loc.indices)
        halo = cols[(cols < a) | (cols >= b)]
        want = [halo[(halo >= off[r]) & (halo < off[r + 1])] for r in range(world)]
        reqs = [None] * world
        dist.all_gather_object(reqs, want)                       # reqs[s][r] = rows of rank r that rank s reads
        send = {s: reqs[s][rank] for s in range(world) if s != rank}
        payload = {s: [(P_loc.indices[P_loc.indptr[g - a]:P_loc.indptr[g - a + 1]] + coff[rank],
                        P_loc.data[P_loc.indptr[g - a]:P_loc.indptr[g - a + 1]]) for g in rows] for s, rows in send.items()}
        allp = [None] * world
        dist.all_gather_object(allp, payload)
        # P_ext: my rows (global coarse columns) + the received halo rows, indexed by GLOBAL fine row
        ncg = int(coff[-1])
        rows_, cols_, vals_ = [], [], []
        Pl = P_loc.tocoo()
        rows_.append(Pl.row + a)
        cols_.append(Pl.col + coff[rank])
        vals_.append(Pl.data)
        for r in range(world):
            if r == rank:
                continue
            for g, (c, w) in zip(want[r], allp[r][rank]):
                rows_.append(np.full(len(c), g))
                cols_.append(c)
                vals_.append(w)
        Pext = sp.csr_matrix((np.concatenate(vals_), (np.concatenate(rows_), np.concatenate(cols_))), shape=(N, ncg))
        Ac_rows = (P_loc.T @ (A_loc @ Pext)).tocsr()             # na x ncg, global coarse columns
        # next-level plan: requests of the off-range coarse columns
        ccols = np.unique(Ac_rows.indices)
        chalo = ccols[(ccols < coff[rank]) | (ccols >= coff[rank + 1])]
        cwant = [chalo[(chalo >= coff[r]) & (chalo < coff[r + 1])] for r in range(world)]
        creq = [None] * world
        dist.all_gather_object(creq, cwant)
        csend = {s: creq[s][rank] for s in range(world) if s != rank}
        q.put((rank, a, b, P_loc, int(coff[rank]), Ac_rows, cwant, csend))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_partitioned_galerkin_equals_global_triple_product(orc):
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    port = _free_port()
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 10
    o, i, v = orc.poisson3d(n)
    N = n ** 3
    A = sp.csr_matrix((v, i, o), shape=(N, N))
    P = sp.block_diag([r[3] for r in res], format="csr")          # rank-local P: block diagonal over the ranks
    Ac = (P.T @ A @ P).tocsr()
    for rank, a, b, P_loc, c0, Ac_rows, cwant, csend in res:
        na = P_loc.shape[1]
        ref = Ac[c0:c0 + na, :]
        assert abs(Ac_rows - ref).max() < 1e-13                   # the distributed product == the rows of P^T A P
        assert (Ac_rows != 0).nnz == (ref != 0).nnz
        # constants stay in the range of P on interior rows of the global matrix (row sum zero): P 1 = 1 there
        rowsum = np.asarray(A[a:b, :].sum(axis=1)).ravel()
        ones = P_loc @ np.ones(na)
        assert np.allclose(ones[np.abs(rowsum) < 1e-14], 1.0)
    # what rank g sends at the next level is what rank q asked for (symmetric neighbour relation for a symmetric matrix)
    for g in range(world):
        for s in range(world):
            if g != s:
                assert np.array_equal(res[g][7][s], res[s][6][g])
                assert (len(res[g][7][s]) > 0) == (len(res[s][7][g]) > 0)
    assert abs(Ac - Ac.T).max() < 1e-13
