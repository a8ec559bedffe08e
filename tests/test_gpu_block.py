"""Block mode (set_block_size(B) / AMGCL_Block<B>, reference src/polysolve/linear/AMGCL.cpp:111-123,246-298).

The GPU realises AMGCL's B x B value type by its scalar expansion (amg.cu "block mode"); the oracle
(oracle/amg_block_oracle.py) restates the algorithm in genuine block arithmetic. Parity is checked both ways, as for the
scalar path: oracle (greedy) aggregates imposed on the GPU, and the GPU's MIS-2 aggregates imposed on the oracle.
The matrix is C4's P1-tet elasticity operator at test size."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def make(psb, B, **kw):
    s = psb.Solver.create("CUDA", "")
    p = dict(precond="amg", tolerance=1e-8, max_iter=500)
    p.update(kw)
    s.set_parameters({"CUDA": p})
    s.set_block_size(B)
    return s


@pytest.mark.parametrize("B", [2, 3])
def test_block_expansion_bit_exact(psb, B):
    """analyze_pattern with block_size B: the stored CSR is the block_matrix adapter's view -- full B x B blocks on the union
    pattern of each block row, bit-exact against scipy's tobsr(); perm maps fill-in to -1."""
    rng = np.random.default_rng(5)
    n = 60 * B
    M = sp.random(n, n, density=0.03, random_state=rng, format="csr") + sp.identity(n)
    M = (M + M.T).tocsc()
    M.sort_indices()
    o, i, v = M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.copy()
    s = psb.Solver.create("CUDA", "")
    s.set_block_size(B)
    s.analyze_pattern_raw(n, o, i, n)
    s.factorize_raw(n, o, i, v)
    full = M.tocsr().tobsr((B, B))
    full.sort_indices()
    nnz_full = full.data.size
    rp, ci, perm = s.debug_get_csr(n, nnz_full)
    # expected scalar CSR of the full-block pattern
    nb = n // B
    exp_rp = np.zeros(n + 1, np.int64)
    cnt = np.diff(full.indptr)
    exp_rp[1:] = np.cumsum(np.repeat(cnt * B, B))
    assert np.array_equal(rp, exp_rp)
    exp_ci = np.concatenate([np.tile((B * full.indices[full.indptr[I]:full.indptr[I + 1]][:, None] + np.arange(B)).ravel(), B)
                             for I in range(nb)])
    assert np.array_equal(ci, exp_ci)
    vals = np.where(perm >= 0, v[np.maximum(perm, 0)], 0.0)
    A = sp.csr_matrix((vals, ci, rp), shape=(n, n))
    assert abs(A - M).max() == 0
    x = rng.standard_normal(n)
    np.testing.assert_allclose(s.spmv(x), M @ x, rtol=1e-13, atol=1e-13)


def test_block3_jacobi_pcg_matches_oracle(psb, orc):
    P = psb.problems
    o, i, v, b = P.elasticity3d(10)
    n = len(b)
    s = psb.Solver.create("CUDA", "")
    s.set_parameters({"CUDA": {"tolerance": 1e-9, "max_iter": 5000}})
    s.set_block_size(3)
    s.factorize_raw(n, o, i, v)
    x = np.zeros(n)
    s.solve(b, x)
    x0, it0, _, _ = orc.eigen_cg(o, i, v, b, tol=1e-9, max_iters=5000)
    info = s.get_info()
    assert info["solver_status"] == "Converged" and abs(info["solver_iter"] - it0) <= 2
    assert np.linalg.norm(x - x0) / np.linalg.norm(x0) < 1e-7


def _node_csr(A_bsr):
    return A_bsr.indptr, A_bsr.indices


def test_block3_amg_with_oracle_aggregates(psb):
    """Greedy (AMGCL) node aggregates imposed on the GPU: same hierarchy sizes, same spectral-radius estimates, same CG
    iteration count and the same solution as the block-arithmetic oracle."""
    from oracle import amg_block_oracle as BO
    P = psb.problems
    o, i, v, b = P.elasticity3d(14)
    n = len(b)
    H = BO.BlockAmg(o, i, v, 3)
    assert len(H.levels) >= 2
    s = make(psb, 3)
    for l in range(len(H.levels) - 1):
        s.debug_set_aggregates(l, H.levels[l].agg.astype(np.int32))
    s.factorize_raw(n, o, i, v)
    amg = s.get_info()["amg"]
    assert amg["block_size"] == 3
    assert [q["rows"] for q in amg["levels"]] == [L.A.shape[0] for L in H.levels]
    # the GPU keeps explicit zero blocks the block arithmetic drops only if a whole block cancels: compare nnz as blocks
    assert [q["nnz"] for q in amg["levels"]] == [L.A.data.size for L in H.levels]
    for q, L in zip(amg["levels"], H.levels):
        assert abs(q["rho"] - L.rho) < 1e-9 * L.rho
    for l in range(len(H.levels) - 1):
        assert abs(amg["levels"][l]["omega"] - H.levels[l].omega) < 1e-12
        rows, cols, rp, ci, va = s.debug_get_level(l, "P")
        Pg = sp.csr_matrix((va, ci, rp), shape=(rows, cols))
        assert abs(Pg - H.levels[l].P).max() < 1e-12
        rows, cols, rp, ci, va = s.debug_get_level(l + 1, "A")
        Ag = sp.csr_matrix((va, ci, rp), shape=(rows, cols))
        assert abs(Ag - H.levels[l + 1].Acsr).max() < 1e-11
    # one preconditioner application, then the full solve
    r = P.splitmix64(3, n)
    z = s.precond_apply(r)
    zo = H.apply(r)
    assert np.linalg.norm(z - zo) / np.linalg.norm(zo) < 1e-9
    x = np.zeros(n)
    s.solve(b, x)
    xo, ito, relo = H.cg(b, tol=1e-8)
    info = s.get_info()
    assert info["num_iterations"] == ito and info["solver_status"] == "Converged"
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-8


@pytest.mark.parametrize("B,m", [(3, 12), (3, 16)])
def test_block_amg_mis2_matches_oracle_with_same_aggregates(psb, B, m):
    from oracle import amg_block_oracle as BO
    P = psb.problems
    o, i, v, b = P.elasticity3d(m)
    n = len(b)
    s = make(psb, B)
    s.factorize_raw(n, o, i, v)
    amg = s.get_info()["amg"]
    nlev = len(amg["levels"])
    imposed = []
    for l in range(nlev - 1):
        rows = amg["levels"][l]["rows"]
        agg, na = s.debug_get_aggregates(l, rows)
        assert na == amg["levels"][l + 1]["rows"]
        node = agg[::B]
        # scalar ids are B * node id + component
        assert np.array_equal(agg[node.repeat(B) >= 0], (node.repeat(B) + np.tile(np.arange(B), rows // B))[node.repeat(B) >= 0])
        imposed.append(np.where(node >= 0, node // B, node))
    H = BO.BlockAmg(o, i, v, B, imposed=imposed)
    assert [L.A.shape[0] for L in H.levels] == [q["rows"] for q in amg["levels"]]
    x = np.zeros(n)
    s.solve(b, x)
    xo, ito, _ = H.cg(b, tol=1e-8)
    info = s.get_info()
    assert info["num_iterations"] == ito
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-8
    A = sp.csr_matrix((v, i, o), shape=(n, n))
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) < 1e-7  # reference bar, tests/test_linear_solver.cpp:600-601


def test_block3_amg_c4_shape(psb):
    """C4's operator at 48^3 nodes (331,776 DoF): block-3 SA-AMG-PCG to 1e-8, residual recomputed on the host; the scalar
    (block_size 1) AMG on the same matrix needs more iterations -- the reason AMGCL_Block exists."""
    P = psb.problems
    o, i, v, b = P.elasticity3d(48)
    n = len(b)
    A = sp.csr_matrix((v, i, o), shape=(n, n))
    its = {}
    for B in (3, 1):
        s = make(psb, B)
        s.factorize_raw(n, o, i, v)
        x = np.zeros(n)
        s.solve(b, x)
        info = s.get_info()
        assert info["solver_status"] == "Converged", info
        assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) < 1e-7
        its[B] = info["num_iterations"]
    assert its[3] <= its[1]


@pytest.mark.parametrize("m", [11, 24])
def test_bsr3_schedule_matches_scalar_csr_and_oracle(psb, orc, m):
    """BSR-3 storage (76 B per 3 x 3 block; the reference's GPU path stores BSR too, mas_utils/BSRMatrix.cu:195-474): the
    block-3 solver picks it automatically, its SpMV equals the oracle's CSC product to 1e-13 and the scalar-CSR schedules of
    the same matrix, with unsymmetric values and an uneven block count per row; PCG on it takes the oracle's iterations."""
    P = psb.problems
    o, i, v, b = P.elasticity3d(m)
    n = len(b)
    v = v * (1.0 + 0.1 * orc.splitmix64(17, len(v)))   # unsymmetric values: exercises the row-major block layout
    x = orc.splitmix64(3, n)
    y0 = orc.spmv_csc(o, i, v, x)                       # CSC product = A x
    # the oracle multiplies the CSC arrays as A; the solver reads the same arrays as CSC and transposes them to CSR
    s = psb.Solver.create("CUDA", "")
    s.set_parameters({"CUDA": {"tolerance": 1e-9, "max_iter": 5000}})
    s.set_block_size(3)
    s.factorize_raw(n, o, i, v)
    info = s.get_info()
    assert info["spmv_kernel"] == ("bsr3" if n >= 3072 else info["spmv_kernel"])
    y = s.spmv(x)
    assert np.abs(y - y0).max() <= 1e-13 * np.abs(y0).max()
    for k in ("stream8n", "vector16", "bsr"):
        s.set_parameters({"CUDA": {"spmv_kernel": k}})
        assert np.abs(s.spmv(x) - y0).max() <= 1e-13 * np.abs(y0).max(), k
    assert s.get_info()["spmv_kernel"] == "bsr3"
    # symmetric values again: Jacobi-PCG on the BSR schedule == the oracle
    o, i, v, b = P.elasticity3d(m)
    s.factorize_raw(n, o, i, v)
    xs = np.zeros(n)
    s.solve(b, xs)
    x0, it0, _, _ = orc.eigen_cg(o, i, v, b, tol=1e-9, max_iters=5000)
    assert abs(s.get_info()["solver_iter"] - it0) <= max(2, 0.02 * it0)
    assert np.linalg.norm(xs - x0) / np.linalg.norm(x0) < 1e-7
