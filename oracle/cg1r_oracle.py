"""TEST INFRASTRUCTURE ONLY. PARITY UNPINNED (see oracle_core.hpp header).

numpy restatement of the single-reduction (Chronopoulos-Gear) form of the Jacobi-PCG that `krylov = "cg1r"` runs
(polysolve_b200/csrc/dist.cu: EpiCg1r / FinCg1r / cg1r_update_kernel). There is no reference counterpart for the
re-ordering itself: what it must reproduce is Eigen's `conjugate_gradient()` as polysolve configures it
(reference src/polysolve/linear/Solver.cpp:433-435, EigenSolver.tpp:75-114) -- same start-up (zero rhs,
converged initial guess), same threshold tol^2 * |b|^2 on |r|^2, same counting rule (the counter is incremented
after a trip that did not converge), same reported error sqrt(|r|^2 / |b|^2).

    u = M r, w = A u, gamma = r.u, delta = w.u          (one reduction: gamma, delta, |r|^2)
    beta = gamma / gamma_old, alpha = gamma / (delta - beta * gamma / alpha_old)
    p = u + beta p, s = w + beta s, x += alpha p, r -= alpha s
"""
import numpy as np


def cg1r(A, b, x0=None, dinv=None, tol=1e-10, max_iters=1000):
    """A: scipy sparse (symmetric). Returns (x, iterations, error, status)."""
    n = A.shape[0]
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
    dinv = np.ones(n) if dinv is None else dinv
    bn2 = float(b @ b)
    if bn2 == 0.0:  # Eigen: x = 0, 0 iterations, error 0
        return np.zeros(n), 0, 0.0, "Converged"
    r = b - A @ x
    rn2 = float(r @ r)
    thr = tol * tol * bn2
    if rn2 < thr:  # Eigen: the initial guess already satisfies the tolerance
        return x, 0, float(np.sqrt(rn2 / bn2)), "Converged"
    p = np.zeros(n)
    s = np.zeros(n)
    u = dinv * r
    it = 0
    gamma_old = alpha_old = None
    while True:
        w = A @ u
        gamma, delta, rn2 = float(r @ u), float(w @ u), float(r @ r)
        if gamma_old is not None and rn2 < thr:
            return x, it, float(np.sqrt(rn2 / bn2)), "Converged"
        if not np.isfinite(rn2) or delta != delta:
            return x, it, float("nan"), "Breakdown"
        if gamma_old is None:
            beta, alpha = 0.0, gamma / delta
        else:
            beta = gamma / gamma_old
            alpha = gamma / (delta - beta * gamma / alpha_old)
            it += 1
            if it >= max_iters:
                return x, it, float(np.sqrt(rn2 / bn2)), "Reach max iterations"
        gamma_old, alpha_old = gamma, alpha
        p = u + beta * p
        s = w + beta * s
        x = x + alpha * p
        r = r - alpha * s
        u = dinv * r
