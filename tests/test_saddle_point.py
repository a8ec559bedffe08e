"""The SaddlePointSolver call site (reference src/polysolve/linear/SaddlePointSolver.cpp:168-215): its inner
solves are created by name with Solver::create, so registering "CUDA" makes them run on the GPU -- asymmetric solves with
Jacobi-PCG / BiCGSTAB on As, the Schur-like solve with BiCGSTAB on Ss (negative definite). The outer loop is the oracle's
restatement; only the inner solvers differ between the CPU and the GPU test."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _system(orc, n=18, m=60, seed=1):
    o, i, v = orc.poisson2d(n)
    A = sp.csc_matrix((v, i, o), shape=(n * n, n * n)) + 0.5 * sp.identity(n * n)
    rng = np.random.default_rng(seed)
    B = sp.random(n * n, m, density=0.03, random_state=rng, format="csc")
    C = -0.05 * sp.identity(m, format="csc")
    K = sp.bmat([[A, B], [B.T, C]], format="csc")
    return K, n * n, rng.standard_normal(n * n + m)


class _Direct:
    def analyze_pattern(self, A, k): pass
    def factorize(self, A): self.lu = spla.splu(sp.csc_matrix(A))
    def solve(self, b, x): x[:] = self.lu.solve(b)


def test_oracle_saddle_point_with_direct_inner_solves(orc):
    from oracle import saddle_oracle as SO
    K, p, rhs = _system(orc)
    # Ss = Cs - Bs^T Bs only approximates the Schur complement, so the outer loop is a genuinely iterative method:
    # the residual falls by ~3x per outer iteration whatever the accuracy of the inner solves
    s = SO.SaddlePointSolver(lambda role: _Direct(), conv_tol=1e-5)
    s.analyze_pattern(K, p)
    s.factorize(K)
    x = np.zeros(K.shape[0])
    s.solve(rhs, x)
    assert s.final_res_norm < 1e-5 and 5 <= s.num_iterations <= 15
    np.testing.assert_allclose(x, spla.spsolve(K, rhs), rtol=0, atol=1e-4)


@pytest.mark.gpu
def test_saddle_point_inner_solves_on_gpu(psb, orc):
    from oracle import saddle_oracle as SO
    K, p, rhs = _system(orc)

    def make(role):
        s = psb.Solver.create("CUDA", "")
        # symmetric_solver on Ss = Cs - Bs^T Bs (negative definite): BiCGSTAB; asymmetric_solver on the SPD As: Jacobi-PCG
        s.set_parameters({"CUDA": {"krylov": "bicgstab" if role == "symmetric" else "cg", "tolerance": 1e-12, "max_iter": 5000}})
        return s
    s = SO.SaddlePointSolver(make, conv_tol=1e-5)
    s.analyze_pattern(K, p)
    s.factorize(K)
    x = np.zeros(K.shape[0])
    s.solve(rhs, x)
    ref = SO.SaddlePointSolver(lambda role: _Direct(), conv_tol=1e-5)
    ref.analyze_pattern(K, p)
    ref.factorize(K)
    xr = np.zeros(K.shape[0])
    ref.solve(rhs, xr)
    assert s.final_res_norm < 1e-5, s.final_res_norm
    assert abs(s.num_iterations - ref.num_iterations) <= 1      # GPU inner solves == exact inner solves, outer loop unchanged
    np.testing.assert_allclose(x, spla.spsolve(K, rhs), rtol=0, atol=1e-4)
