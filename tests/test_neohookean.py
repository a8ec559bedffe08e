"""BASELINE config 5: the device-resident Neo-Hookean Problem (include/psb200_problems.h) and the Newton driver with a
device-assembled Hessian (Newton::solve_sparse_linear_system, reference Newton.cpp:173-214, without a PCIe crossing of
the matrix). Checker: oracle/neohookean_oracle.py (numpy restatement, itself pinned by finite differences here)."""
import numpy as np
import pytest

PARAMS = {"solver": "Newton", "line_search": {"method": "Backtracking"}, "grad_norm_tol": 1e-8, "rel_grad_norm_tol": 0,
          "max_iterations": 50, "Newton": {"residual_tolerance": 1e-5}}


def test_oracle_gradient_and_hessian_match_finite_differences():
    from oracle import neohookean_oracle as NH
    prob, x0 = NH.stretch_problem(5)
    rng = np.random.default_rng(0)
    free = ~prob.fixed
    x = x0 + 0.02 * rng.standard_normal(prob.n) * free
    d = rng.standard_normal(prob.n) * free
    e = 1e-6
    g, H = prob.gradient(x), prob.hessian(x)
    fd = (prob.value(x + e * d) - prob.value(x - e * d)) / (2 * e)
    assert abs(fd - g @ d) <= 1e-7 * abs(fd)
    gd = (prob.gradient(x + e * d) - prob.gradient(x - e * d)) / (2 * e)
    Hd = H @ d
    Hd[prob.fixed] = 0
    assert np.linalg.norm(gd - Hd) <= 1e-7 * np.linalg.norm(gd)
    assert abs(H - H.T).max() < 1e-12
    # fixed dofs: identity rows and columns
    D = H.toarray()
    fx = np.flatnonzero(prob.fixed)
    assert np.array_equal(D[fx][:, fx], np.eye(len(fx))) and not D[fx][:, np.flatnonzero(free)].any()


def test_oracle_newton_converges_on_the_stretch_problem():
    import scipy.sparse.linalg as spla
    from oracle import neohookean_oracle as NH
    from oracle import newton_oracle as NO
    prob, x0 = NH.stretch_problem(6)
    x, info = NO.minimize(prob, x0.copy(), PARAMS, lambda H, rhs, g0: (spla.spsolve(H.tocsc(), rhs), 1))
    assert info["status"] == "GradNormTolerance" and info["iterations"] <= 6
    u = x.reshape(-1, 3)
    assert abs(u[:, 1]).max() > 1e-3          # lateral contraction: the affine start is not the solution
    assert np.allclose(x[prob.fixed], x0[prob.fixed])


def test_device_problem_fails_loudly_without_gpu(psb):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("checks the no-GPU failure mode")
    except ImportError:
        pass
    with pytest.raises(RuntimeError, match="no CUDA device"):
        psb.neohookean.stretch_problem(4)


@pytest.mark.gpu
@pytest.mark.parametrize("m", [4, 7])
def test_device_problem_matches_oracle(psb, m):
    """Energy, gradient and the Hessian assembled by the CUDA kernels straight into the CSC pattern == the numpy
    restatement (1e-12 relative); the pattern has full 3 x 3 blocks and sorted rows; two evaluations are bit-identical."""
    from oracle import neohookean_oracle as NH
    po, x0 = NH.stretch_problem(m)
    pd, x0d = psb.neohookean.stretch_problem(m)
    assert np.array_equal(x0, x0d)
    rng = np.random.default_rng(1)
    x = x0 + 0.03 * rng.standard_normal(po.n) * (~po.fixed)
    assert abs(pd.value(x) - po.value(x)) <= 1e-12 * abs(po.value(x))
    g0, g1 = po.gradient(x), pd.gradient(x)
    assert np.abs(g1 - g0).max() <= 1e-12 * np.abs(g0).max()
    H0, H1 = po.hessian(x), pd.hessian(x)
    H1.sort_indices()
    assert np.all(np.diff(pd.outer) % 3 == 0)
    assert all(np.all(np.diff(pd.inner[pd.outer[c]:pd.outer[c + 1]]) > 0) for c in range(0, pd.n, 37))   # rows ascending
    assert abs(H1 - H0).max() <= 1e-12 * abs(H0).max()
    assert (H1 != 0).nnz <= H1.nnz and H1.nnz == pd.nnz
    assert np.array_equal(pd.gradient(x), g1) and pd.value(x) == pd.value(x)
    assert np.array_equal(pd.hessian(x).data, pd.hessian(x).data)
    # an inverted element gives +inf (the line search treats it as an invalid step)
    xb = x0.copy()
    xb[3 * (m * m + m + 1)] += 5.0
    assert pd.value(xb) == float("inf") and po.value(xb) == float("inf")


@pytest.mark.gpu
@pytest.mark.parametrize("precond", ["jacobi", "amg"])
def test_newton_with_device_hessian_matches_oracle(psb, precond):
    """Newton + Backtracking on the Neo-Hookean stretch problem: the driver assembles H on the GPU (hessian_device ->
    psb200_factorize_csc_device), solves with the GPU PCG (block size 3 for AMG) and reaches the oracle's minimiser in the
    same number of Newton iterations."""
    import scipy.sparse.linalg as spla
    from oracle import neohookean_oracle as NH
    from oracle import newton_oracle as NO
    m = 8
    po, x0 = NH.stretch_problem(m)
    xo, io = NO.minimize(po, x0.copy(), PARAMS, lambda H, rhs, g0: (spla.spsolve(H.tocsc(), rhs), 1))
    pd, _ = psb.neohookean.stretch_problem(m)
    lin = {"solver": "CUDA", "CUDA": {"tolerance": 1e-10, "max_iter": 2000, "precond": precond, "block_size": 3 if precond == "amg" else 1,
                                      "amg": {"coarse_enough": 300}}}
    s = psb.NonlinearSolver.create(PARAMS, lin)
    x = x0.copy()
    s.minimize(pd, x)
    info = s.get_info()
    assert info["succeeded"] and info["status"] == "Gradient vector norm too small"
    assert info["iterations"] == io["iterations"]
    assert np.abs(x - xo).max() < 1e-7
    assert all(i["solver_status"] == "Converged" for i in info["internal_solver"])
    assert [i["analyze_skipped"] for i in info["internal_solver"]][1:] == [True] * (len(info["internal_solver"]) - 1)
