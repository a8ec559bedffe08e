"""CPU tests: the oracle against known answers, analytic eigenpairs, scipy and the committed golden
fixtures (tests/golden/make_golden.py). The reference holds no golden vectors for this path
(SURVEY 8c), so these are what pins the restatement."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def csc(o, i, v):
    n = len(o) - 1
    return sp.csc_matrix((v, i, o), shape=(n, n))


def test_splitmix_known_values(orc):
    b = orc.splitmix64(42, 3)
    # SURVEY 8d fixes these three values
    assert b.tolist() == [0.4831297575436466, -0.6801792142461598, -0.4427977394897227]


def test_c1_known_answer(orc):
    """C1 => 115 reported iterations, error() 7.0695e-11, ||Ax-b|| 1.3329e-9."""
    o, i, v = orc.poisson2d(32)
    b = orc.splitmix64(42, 1024)
    x, it, err, spmvs = orc.eigen_cg(o, i, v, b, tol=1e-10, max_iters=1000)
    assert it == 115 and spmvs == 116
    assert abs(err - 7.0695e-11) < 1e-14
    res = np.linalg.norm(csc(o, i, v) @ x - b)
    assert abs(res - 1.3329e-9) < 1e-12 and res < 1e-8  # reference acceptance: test_linear_solver.cpp:160-162
    # warm start from the converged x => 0 iterations (property pinned at test_linear_solver.cpp:449)
    _, it2, _, _ = orc.eigen_cg(o, i, v, b, x0=x, tol=1e-10)
    assert it2 == 0


def test_cg_matches_scipy(orc):
    o, i, v = orc.poisson2d(32)
    A = csc(o, i, v)
    b = orc.splitmix64(42, 1024)
    x, it, err, _ = orc.eigen_cg(o, i, v, b, tol=1e-10)
    xs, info = spla.cg(A, b, rtol=1e-10, M=sp.diags(1 / A.diagonal()), maxiter=1000)
    assert info == 0
    assert np.abs(xs - x).max() < 1e-12


def test_golden_c1(orc):
    g = np.load(os.path.join(GOLD, "c1_poisson2d_32.npz"))
    o, i, v = orc.poisson2d(32)
    x, it, err, _ = orc.eigen_cg(o, i, v, g["b"], tol=1e-10)
    assert it == int(g["iters"])
    np.testing.assert_allclose(x, g["x"], rtol=0, atol=1e-13)
    # and the golden solution really solves the system (independent of the oracle)
    assert np.linalg.norm(csc(o, i, v) @ g["x"] - g["b"]) < 1e-8


def test_golden_bicgstab(orc):
    g = np.load(os.path.join(GOLD, "convdiff2d_32.npz"))
    o, i, v = orc.convdiff2d(32, 0.5)
    x, it, err, _ = orc.eigen_bicgstab(o, i, v, g["b"], tol=1e-10)
    assert it == int(g["iters"])
    np.testing.assert_allclose(x, g["x"], rtol=0, atol=1e-12)
    xs = spla.spsolve(csc(o, i, v).tocsc(), g["b"])
    assert np.abs(xs - x).max() < 1e-9


def test_eigenpairs_poisson3d(orc):
    """Analytic eigenpairs of the Dirichlet Laplacian: A v = lambda v (SURVEY 8c known-answer (i))."""
    n = 12
    o, i, v = orc.poisson3d(n)
    k = np.arange(1, n + 1)
    for (a, b_, c) in [(1, 1, 1), (2, 5, 3), (n, n, n)]:
        s = lambda m: np.sin(np.pi * m * k / (n + 1))
        vec = np.einsum("i,j,k->kji", s(a), s(b_), s(c)).ravel()  # x fastest
        lam = 4 * sum(np.sin(np.pi * m / (2 * (n + 1))) ** 2 for m in (a, b_, c))
        y = orc.spmv_csc(o, i, v, vec)
        np.testing.assert_allclose(y, lam * vec, atol=1e-12)
        y2 = orc.spmv_csr(o, i, v, vec)
        np.testing.assert_allclose(y2, lam * vec, atol=1e-12)


def test_transpose_index_oracle(orc):
    rng = np.random.default_rng(7)
    A = sp.random(300, 300, density=0.03, random_state=rng, format="csc") + sp.eye(300, format="csc")
    A = sp.csc_matrix(A)
    A.sort_indices()
    o, i, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data
    rp, ci, perm = orc.csc_to_csr(300, o, i)
    R = A.tocsr()
    R.sort_indices()
    assert np.array_equal(rp, R.indptr) and np.array_equal(ci, R.indices)
    np.testing.assert_array_equal(v[perm], R.data)


def test_partition_and_halo_oracle(orc):
    o, i, v = orc.poisson3d(10)
    n = 1000
    off = orc.partition_rows(o, 4)
    assert off[0] == 0 and off[-1] == n and np.all(np.diff(off) > 0)
    nnz_share = np.diff(o[off])
    assert nnz_share.max() - nnz_share.min() <= 2 * 7 * 100  # balanced to within a plane
    for g in range(4):
        lc, halo = orc.halo_for_rank(o, i, int(off[g]), int(off[g + 1]))
        nl = off[g + 1] - off[g]
        glob = np.where(lc < nl, lc + off[g], 0)
        glob[lc >= nl] = halo[lc[lc >= nl] - nl]
        assert np.array_equal(glob, i[o[off[g]]:o[off[g + 1]]])
        assert np.all(np.diff(halo) > 0)


def test_prefactor_values_symmetric_spd(orc):
    """Value generator of the reference's pattern-reuse test (test_linear_solver.cpp:262-283)."""
    o, i, _ = orc.poisson2d(8)
    vals = orc.prefactor_values(o, i, rounds=3)
    for r in range(3):
        A = csc(o, i, vals[r])
        assert abs(A - A.T).max() == 0
        d = A.diagonal()
        assert d.min() >= 10 and d.max() <= 500
        assert np.all(np.linalg.eigvalsh(A.toarray()) > 0)
    assert not np.array_equal(vals[0], vals[1])


def test_amg_known_answer(orc):
    """32^3 Poisson, polysolve's AMGCL defaults => 32768 -> 4192 -> 117 rows,
    nnz 223232 / 114356 / 4349, 4 CG iterations to 1e-10 and 3 to 1e-8 (7 / 5 with ncycle 1)."""
    o, i, v = orc.poisson3d(32)
    b = orc.spmv_csc(o, i, v, orc.splitmix64(42, 32 ** 3))
    H = orc.Amg(o, i, v)
    assert H.num_levels == 3
    infos = [H.level_info(l) for l in range(3)]
    assert [q["rows"] for q in infos] == [32768, 4192, 117]
    assert [q["nnz"] for q in infos] == [223232, 114356, 4349]
    x, it, rel = H.cg(b, tol=1e-10)
    assert it == 4 and rel < 1e-10
    assert np.linalg.norm(csc(o, i, v) @ x - b) / np.linalg.norm(b) < 1e-9
    _, it8, _ = H.cg(b, tol=1e-8)
    assert it8 == 3
    H1 = orc.Amg(o, i, v, ncycle=1)
    assert H1.cg(b, tol=1e-10)[1] == 7 and H1.cg(b, tol=1e-8)[1] == 5
    # warm start => 0 iterations (test_linear_solver.cpp:432-450)
    assert H.cg(b, x0=x, tol=1e-10)[1] == 0


def test_amg_galerkin_and_transfer_properties(orc):
    o, i, v = orc.poisson3d(16)
    H = orc.Amg(o, i, v, coarse_enough=200)
    assert H.num_levels >= 2
    A0 = sp.csr_matrix(H.matrix(0, "A")[::-1][0:1] + H.matrix(0, "A")[1::-1], shape=(4096, 4096)) if False else None
    pa, ca, va = H.matrix(0, "A")
    pp, cp, vp = H.matrix(0, "P")
    pr, cr, vr = H.matrix(0, "R")
    p1, c1, v1 = H.matrix(1, "A")
    nc = len(p1) - 1
    A = sp.csr_matrix((va, ca, pa), shape=(4096, 4096))
    P = sp.csr_matrix((vp, cp, pp), shape=(4096, nc))
    R = sp.csr_matrix((vr, cr, pr), shape=(nc, 4096))
    Ac = sp.csr_matrix((v1, c1, p1), shape=(nc, nc))
    assert abs(R - P.T).max() == 0
    assert abs(Ac - R @ A @ P).max() < 1e-12
    # smoothed-aggregation prolongator preserves constants in the interior: rows of P sum to 1 where
    # the row of A sums to 0 (P = (I - w D^-1 A) P_tent and P_tent 1 = 1)
    rows_interior = np.asarray(abs(A @ np.ones(4096)) < 1e-14).ravel()
    np.testing.assert_allclose(np.asarray(P.sum(axis=1)).ravel()[rows_interior], 1.0, atol=1e-13)
    agg = H.aggregates(0)
    assert agg.min() >= 0 and agg.max() == nc - 1 and len(np.unique(agg)) == nc


def test_block_oracle_reduces_to_scalar_oracle(orc):
    """The block-arithmetic AMG oracle with B = 1 is the scalar AMGCL restatement (same hierarchy, same iterates)."""
    from oracle import amg_block_oracle as BO
    o, i, v = orc.poisson3d(20)
    b = orc.spmv_csc(o, i, v, orc.splitmix64(42, 8000))
    H1, Hc = BO.BlockAmg(o, i, v, 1), orc.Amg(o, i, v)
    assert [r for r, _ in H1.level_sizes()] == [Hc.level_info(l)["rows"] for l in range(Hc.num_levels)]
    assert [z for _, z in H1.level_sizes()] == [Hc.level_info(l)["nnz"] for l in range(Hc.num_levels)]
    x1, it1, _ = H1.cg(b, tol=1e-8)
    xc, itc, _ = Hc.cg(b, tol=1e-8)
    assert it1 == itc and np.abs(x1 - xc).max() < 1e-13


def test_elasticity_generator_matches_element_assembly():
    """C4 generator (stencil-wise assembly) against a brute-force element loop; SPD; block oracle converges on it."""
    import scipy.sparse as sp
    from polysolve_b200 import problems as P
    from oracle import amg_block_oracle as BO
    m = 5
    o, i, v, b = P.elasticity3d(m)
    N = len(b)
    A = sp.csr_matrix((v, i, o), shape=(N, N))
    nid = lambda p: int(p[0] + m * p[1] + m * m * p[2])  # noqa: E731
    ref = np.zeros((N, N))
    for cz in range(m - 1):
        for cy in range(m - 1):
            for cx in range(m - 1):
                for T in P._kuhn_tets():
                    K = P._p1_elastic_element(T, 1.0, 0.3)
                    ids = [nid(T[a] + np.array([cx, cy, cz])) for a in range(4)]
                    dof = np.array([3 * q + c for q in ids for c in range(3)])
                    ref[np.ix_(dof, dof)] += K
    cl = np.array([3 * nid((0, j, k)) + c for j in range(m) for k in range(m) for c in range(3)])
    ref[cl, :] = 0
    ref[:, cl] = 0
    ref[cl, cl] = 1
    assert np.abs(ref - A.toarray()).max() < 1e-14
    assert np.linalg.eigvalsh(ref).min() > 0
    o, i, v, b = P.elasticity3d(12)
    H = BO.BlockAmg(o, i, v, 3)
    x, it, rel = H.cg(b, tol=1e-8)
    assert len(H.levels) == 2 and 0 < it < 30 and rel < 1e-8


def test_single_reduction_cg_restatement_equals_eigen_ordering(orc):
    """oracle/cg1r_oracle.py (what `krylov = cg1r` runs on the GPU) against the Eigen-ordering oracle: same iterates in
    exact arithmetic => iteration counts within +-2 %, same x, same counting rule (C1: 115), same start-up rules."""
    from oracle import cg1r_oracle

    o, i, v = orc.poisson2d(32)
    b = orc.splitmix64(42, 1024)
    A = csc(o, i, v).tocsr()
    dinv = 1.0 / A.diagonal()
    x0, it0, err0, _ = orc.eigen_cg(o, i, v, b, tol=1e-10, max_iters=1000)
    x, it, err, status = cg1r_oracle.cg1r(A, b, dinv=dinv, tol=1e-10, max_iters=1000)
    assert status == "Converged" and abs(it - it0) <= 2 and it0 == 115
    assert err < 1e-10 and np.linalg.norm(x - x0) / np.linalg.norm(x0) < 1e-9
    # warm start => 0 iterations, x untouched; zero rhs => x = 0
    x2, it2, _, _ = cg1r_oracle.cg1r(A, b, x0=x, dinv=dinv, tol=1e-10)
    assert it2 == 0 and np.array_equal(x2, x)
    x3, it3, err3, _ = cg1r_oracle.cg1r(A, np.zeros(1024), x0=b, dinv=dinv)
    assert it3 == 0 and err3 == 0.0 and not x3.any()
    # max_iter: the counter equals the number of trips (Eigen: `while (i < maxIters)`)
    for mi in (1, 2, 7):
        _, itm, _, st = cg1r_oracle.cg1r(A, b, dinv=dinv, tol=1e-14, max_iters=mi)
        _, ite, _, _ = orc.eigen_cg(o, i, v, b, tol=1e-14, max_iters=mi)
        assert itm == mi == ite and st == "Reach max iterations"
    # a perturbed 3-D operator (non-constant diagonal)
    o, i, v = orc.poisson3d(14)
    N = 14 ** 3
    v = v * (1.0 + 0.05 * orc.splitmix64(13, len(v)))
    v = 0.5 * (v + v[orc.csc_to_csr(N, o, i)[2]])
    b = orc.splitmix64(7, N)
    A = csc(o, i, v).tocsr()
    x0, it0, _, _ = orc.eigen_cg(o, i, v, b, tol=1e-9, max_iters=5000)
    x, it, err, _ = cg1r_oracle.cg1r(A, b, dinv=1.0 / A.diagonal(), tol=1e-9, max_iters=5000)
    assert abs(it - it0) <= max(1, 0.02 * it0) and np.linalg.norm(x - x0) / np.linalg.norm(x0) < 1e-7
