"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the compressible Neo-Hookean P1-tetrahedron energy used as the
`Problem` of BASELINE config 5 ("Newton + backtracking on a nonlinear elasticity Problem subclass"; the reference's
Problem interface is src/polysolve/nonlinear/Problem.hpp:22-143, its users supply value / gradient / hessian).

    W(F) = mu/2 (|F|^2 - 3) - mu ln J + lambda/2 (ln J)^2,   F = I + grad u,   J = det F
    P(F) = mu (F - F^-T) + lambda ln J F^-T
    dP[dF] = mu dF + (mu - lambda ln J) F^-T dF^T F^-T + lambda (F^-T : dF) F^-T

Unknowns are the nodal displacements (3 per node, node-major); fixed dofs keep their value: zero gradient, identity row and
column in the Hessian. Only tests/ and bench.py's CPU leg import this file; the product (polysolve_b200/csrc/neohookean.cu)
is checked against it and against finite differences."""
import numpy as np
import scipy.sparse as sp


def grid_tets(m):
    """Kuhn 6-tet split of the (m-1)^3 cells of an m^3-node unit-spacing grid; node id = i + m j + m^2 k."""
    import itertools
    idx = np.arange(m ** 3).reshape(m, m, m)  # [k, j, i]
    X = np.stack(np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij"), -1).reshape(-1, 3)[:, ::-1].astype(float)
    tets = []
    c = (slice(0, m - 1),) * 3
    for perm in itertools.permutations(range(3)):
        v = [np.zeros(3, int)]
        for ax in perm:
            w = v[-1].copy()
            w[ax] += 1
            v.append(w)
        cols = []
        for w in v:
            sl = tuple(slice(w[ax], w[ax] + m - 1) for ax in (2, 1, 0))
            cols.append(idx[sl].reshape(-1))
        tets.append(np.stack(cols, 1))
    T = np.concatenate(tets, 0).astype(np.int32)
    # positive orientation
    d = X[T[:, 1:]] - X[T[:, :1]]
    neg = np.linalg.det(d) < 0
    T[neg, 2], T[neg, 3] = T[neg, 3].copy(), T[neg, 2].copy()
    return X, T


class NeoHookean:
    def __init__(self, X, T, mu=1.0, lam=1.5, fixed=None):
        self.X, self.T, self.mu, self.lam = np.asarray(X, float), np.asarray(T, np.int64), mu, lam
        self.nn = self.X.shape[0]
        self.n = 3 * self.nn
        self.fixed = np.zeros(self.n, bool) if fixed is None else np.asarray(fixed, bool)
        Dm = np.transpose(self.X[self.T[:, 1:]] - self.X[self.T[:, :1]], (0, 2, 1))  # columns = edge vectors
        self.Dminv = np.linalg.inv(Dm)
        self.vol = np.abs(np.linalg.det(Dm)) / 6.0
        G = np.zeros((len(self.T), 4, 3))  # grad N_a
        G[:, 1:, :] = self.Dminv           # row a-1 of Dm^-1
        G[:, 0, :] = -self.Dminv.sum(1)
        self.G = G

    def _F(self, x):
        u = x.reshape(-1, 3)
        pos = self.X + u
        Ds = np.transpose(pos[self.T[:, 1:]] - pos[self.T[:, :1]], (0, 2, 1))
        return Ds @ self.Dminv

    def value(self, x):
        F = self._F(x)
        J = np.linalg.det(F)
        if np.any(J <= 0):
            return float("inf")
        lj = np.log(J)
        W = 0.5 * self.mu * ((F ** 2).sum((1, 2)) - 3) - self.mu * lj + 0.5 * self.lam * lj ** 2
        return float((W * self.vol).sum())

    def gradient(self, x):
        F = self._F(x)
        J = np.linalg.det(F)
        FinvT = np.transpose(np.linalg.inv(F), (0, 2, 1))
        lj = np.log(J)[:, None, None]
        P = self.mu * (F - FinvT) + self.lam * lj * FinvT
        g = np.zeros((self.nn, 3))
        for a in range(4):
            np.add.at(g, self.T[:, a], self.vol[:, None] * np.einsum("trk,tk->tr", P, self.G[:, a, :]))
        g = g.reshape(-1)
        g[self.fixed] = 0.0
        return g

    def hessian(self, x, psd=False):
        F = self._F(x)
        J = np.linalg.det(F)
        Finv = np.linalg.inv(F)
        lj = np.log(J)
        p = np.einsum("tlr,tal->tar", Finv, self.G)   # p_a = F^-T grad N_a
        rows, cols, vals = [], [], []
        for a in range(4):
            for b in range(4):
                gg = np.einsum("tk,tk->t", self.G[:, a], self.G[:, b])
                K = (self.mu * gg)[:, None, None] * np.eye(3)[None] \
                    + (self.mu - self.lam * lj)[:, None, None] * np.einsum("tr,tc->trc", p[:, b], p[:, a]) \
                    + self.lam * np.einsum("tr,tc->trc", p[:, a], p[:, b])
                K = K * self.vol[:, None, None]
                for r in range(3):
                    for c in range(3):
                        rows.append(3 * self.T[:, a] + r)
                        cols.append(3 * self.T[:, b] + c)
                        vals.append(K[:, r, c])
        H = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(self.n, self.n)).tocsc()
        H.sum_duplicates()
        # fixed dofs: identity rows / columns (explicit zeros keep the pattern)
        fx = self.fixed
        H = H.tocoo()
        kill = fx[H.row] | fx[H.col]
        data = H.data.copy()
        data[kill] = 0.0
        data[kill & (H.row == H.col)] = 1.0
        H = sp.coo_matrix((data, (H.row, H.col)), shape=(self.n, self.n)).tocsc()
        H.sort_indices()
        return H

    # Problem.hpp defaults
    def solution_changed(self, x): pass
    def is_step_valid(self, x0, x1): return bool(np.all(np.linalg.det(self._F(x1)) > 0))
    def max_step_size(self, x0, x1): return 1.0
    def line_search_begin(self, x0, x1): pass
    def line_search_end(self): pass
    def post_step(self, it, x, g): pass
    def stop(self, x): return False


def stretch_problem(m, stretch=0.1, mu=1.0, lam=1.5):
    """Config 5 geometry: m^3-node block, face x = 0 clamped, face x = m - 1 displaced by stretch * (m - 1) in x;
    the start vector is the affine stretch u_x = stretch * x (it satisfies both Dirichlet conditions)."""
    X, T = grid_tets(m)
    fixed = np.zeros((m ** 3, 3), bool)
    fixed[X[:, 0] == 0] = True
    fixed[X[:, 0] == m - 1] = True
    x0 = np.zeros((m ** 3, 3))
    x0[:, 0] = stretch * X[:, 0]
    return NeoHookean(X, T, mu, lam, fixed.reshape(-1)), x0.reshape(-1)
