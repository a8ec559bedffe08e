"""Dirichlet pre-processing (reference src/polysolve/linear/FEMSolver.cpp:97-372): the oracle's restatement
is checked on the CPU against the properties the reference comments state (":99-106"); the GPU path is compared with
the oracle and with an independent direct solve of the oracle's system."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _problem(orc, n=14, seed=3):
    o, i, v = orc.poisson3d(n)
    N = n ** 3
    v = v * (1.0 + 0.1 * orc.splitmix64(seed, len(v)))
    v = 0.5 * (v + v[orc.csc_to_csr(N, o, i)[2]])
    A = sp.csc_matrix((v, i, o), shape=(N, N))
    f = orc.splitmix64(seed + 1, N)
    grid = np.arange(N).reshape(n, n, n)
    nodes = np.unique(np.concatenate([grid[0].ravel(), grid[:, :, -1].ravel(), [N // 2]])).astype(np.int32)
    return A, f, nodes


def test_oracle_dirichlet_system_properties(orc):
    from oracle import fem_oracle as F
    A, f, nodes = _problem(orc)
    At, g = F.dirichlet_system(A, f, nodes)
    N = A.shape[0]
    free = np.setdiff1d(np.arange(N), nodes)
    # rows and columns of the Dirichlet dofs are identity; the free block is untouched
    assert abs(At[nodes][:, nodes] - sp.identity(len(nodes))).max() == 0
    assert abs(At[nodes][:, free]).max() == 0 and abs(At[free][:, nodes]).max() == 0
    assert abs(At[free][:, free] - A[free][:, free]).max() == 0
    # g[i] = f[i] on the boundary, f[i] - sum_{j in Gamma} a_ij f[j] elsewhere (FEMSolver.cpp:102-104)
    assert np.array_equal(g[nodes], f[nodes])
    np.testing.assert_allclose(g[free], f[free] - A[free][:, nodes] @ f[nodes], rtol=0, atol=1e-14)
    # the solution takes the prescribed values and solves the free block
    u = spla.spsolve(At, g)
    np.testing.assert_allclose(u[nodes], f[nodes], rtol=0, atol=1e-14)
    np.testing.assert_allclose(A[free] @ u, f[free], rtol=0, atol=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("precond", ["jacobi", "amg"])
def test_gpu_dirichlet_solve_matches_oracle(psb, orc, precond):
    from oracle import fem_oracle as F
    A, f, nodes = _problem(orc)
    N = A.shape[0]
    At, g = F.dirichlet_system(A, f, nodes)
    u0 = spla.spsolve(At, g)
    s = psb.Solver.create("CUDA", "")
    s.set_parameters({"CUDA": {"tolerance": 1e-11, "max_iter": 5000, "precond": precond}})
    fin, u = f.copy(), np.zeros(N)
    s.dirichlet_solve(A, fin, nodes, u, N)
    info = s.get_info()
    assert info["solver_status"] == "Converged"
    np.testing.assert_allclose(fin, g, rtol=0, atol=1e-13)            # f = g on return (FEMSolver.cpp:282)
    assert np.array_equal(fin[nodes], f[nodes])
    np.testing.assert_allclose(u, u0, rtol=0, atol=1e-8)
    assert np.linalg.norm(At @ u - g) / np.linalg.norm(g) < 1e-10
    if precond == "jacobi":
        # the Krylov loop sees the same numbers as the reference's rebuilt matrix: same iteration count as the oracle CG on A~
        At.sort_indices()
        _, it0, _, _ = orc.eigen_cg(At.indptr.astype(np.int32), At.indices.astype(np.int32), At.data, g, tol=1e-11, max_iters=5000)
        assert abs(info["solver_iter"] - it0) <= 1


@pytest.mark.gpu
def test_gpu_dirichlet_prefactorized(psb, orc):
    from oracle import fem_oracle as F
    A, f, nodes = _problem(orc, seed=9)
    N = A.shape[0]
    At, g = F.dirichlet_system(A, f, nodes)
    u0 = spla.spsolve(At, g)
    s = psb.Solver.create("CUDA", "")
    s.set_parameters({"CUDA": {"tolerance": 1e-11, "max_iter": 5000}})
    s.prefactorize(A, nodes, N)
    # several right-hand sides on one factorization, lifted with the original matrix
    for k in range(3):
        fk = f * (k + 1.0)
        fin, u = fk.copy(), np.zeros(N)
        s.dirichlet_solve_prefactorized(A, fin, u)
        np.testing.assert_allclose(fin, g * (k + 1.0), rtol=0, atol=1e-12)
        np.testing.assert_allclose(u, u0 * (k + 1.0), rtol=0, atol=1e-7)
    # lifted with the matrix prefactorize rewrote (the reference's in-place usage): the lifting term vanishes
    fin, u = f.copy(), np.zeros(N)
    s.dirichlet_solve_prefactorized(None, fin, u)
    assert np.array_equal(fin, f)
    np.testing.assert_allclose(At @ u, f, rtol=0, atol=1e-9)
    # errors: a node id out of range, a solve before prefactorize
    with pytest.raises(RuntimeError):
        s.prefactorize(A, np.array([N], np.int32), N)
    s2 = psb.Solver.create("CUDA", "")
    with pytest.raises(RuntimeError):
        s2.dirichlet_solve_prefactorized(None, f.copy(), np.zeros(N))
