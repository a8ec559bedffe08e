"""TEST INFRASTRUCTURE ONLY. PARITY UNPINNED (see oracle_core.hpp header).

Block-valued SA-AMG-PCG: what polysolve::linear::AMGCL_Block<B> obtains from AMGCL 1.4.3 with
value_type = static_matrix<double,B,B> (reference src/polysolve/linear/AMGCL.cpp:246-298: block_matrix adapter at
:270-272, reinterpret_as_rhs at :288-292; parameters AMGCL.cpp:32-65). Restated in *block arithmetic* with numpy /
scipy.sparse.bsr (last bullet): Frobenius norms for strength and Gershgorin, the B x B inverse
of the diagonal block wherever the scalar algorithm divides by a_ii, aggregation on block rows. It is deliberately an
independent formulation from the GPU's scalar-expansion implementation (polysolve_b200/csrc/amg.cu "block mode").
Sized for tests (sequential greedy aggregation is a Python loop)."""
import numpy as np
import scipy.sparse as sp


def _splitmix_unit(seed, n):
    m = np.uint64(0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * np.arange(1, n + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return 2.0 * (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 1.0


class Level:
    pass


class BlockAmg:
    def __init__(self, ptr, col, val, B, imposed=None, max_levels=6, coarse_enough=3000, ncycle=2, npre=1, npost=1, degree=16,
                 power_iters=100, higher=2.0, lower=0.008333333333, sa_relax=1.0, eps_strong=0.0):
        n = len(ptr) - 1
        assert n % B == 0
        self.B = B
        self.prm = dict(max_levels=max_levels, coarse_enough=coarse_enough // B, ncycle=ncycle, npre=npre, npost=npost,
                        degree=degree, power_iters=power_iters, higher=higher, lower=lower, sa_relax=sa_relax)
        A = sp.csr_matrix((val, col, ptr), shape=(n, n)).tobsr((B, B))  # amgcl::adapter::block_matrix
        A.sort_indices()
        self.levels = []
        eps = eps_strong
        imposed = imposed or []
        while A is not None and A.shape[0] // B > self.prm["coarse_enough"]:
            L = self._relax_setup(A, 1000 + len(self.levels) + 1)
            self.levels.append(L)
            if len(self.levels) >= max_levels:
                A = None
                break
            li = len(self.levels) - 1
            strong = self._strong(A, eps)
            if li < len(imposed) and imposed[li] is not None:
                agg = np.asarray(imposed[li], np.int64)
                nc = int(agg.max()) + 1
            else:
                agg, nc = self._plain_aggregates(A, strong)
            eps *= 0.5
            L.agg = agg
            if nc == 0:
                A = None
                break
            omega = sa_relax * (4.0 / 3.0) / self._gershgorin(A, L.Dinv)
            L.omega = omega
            L.P = self._prolongation(A, L.Dinv, strong, agg, nc, omega)
            L.R = L.P.T.tocsr()
            A = (L.R @ (A.tocsr() @ L.P)).tobsr((B, B))
            A.sort_indices()
        if A is not None:
            self.levels.append(self._relax_setup(A, 1000 + len(self.levels) + 1))

    # ---- pieces
    def _diag_blocks(self, A):
        B, nb = self.B, A.shape[0] // self.B
        D = np.zeros((nb, B, B))
        rows = np.repeat(np.arange(nb), np.diff(A.indptr))
        m = A.indices == rows
        D[rows[m]] = A.data[m]
        return D

    def _relax_setup(self, A, seed):
        L = Level()
        L.A = A
        L.Acsr = A.tocsr()
        L.Dinv = np.linalg.inv(self._diag_blocks(A))
        n = A.shape[0]
        B = self.B
        # amgcl spectral_radius<true>(A, power_iters): power iteration on D^-1 A, block vectors
        b0 = _splitmix_unit(seed, n)
        b0 /= np.sqrt(b0 @ b0)
        radius = 1.0
        it = 0
        while it < self.prm["power_iters"]:
            b1 = self._apply_dinv(L.Dinv, L.Acsr @ b0)
            radius = float(np.sum(np.abs(np.sum((b1 * b0).reshape(-1, B), axis=1))))  # sum_i |<b1_i, b0_i>|
            it += 1
            if it < self.prm["power_iters"]:
                b0 = b1 / np.sqrt(b1 @ b1)
        L.rho = radius
        hi, lo = radius * self.prm["higher"], radius * self.prm["lower"]
        L.d, L.c = 0.5 * (hi + lo), 0.5 * (hi - lo)
        return L

    def _apply_dinv(self, Dinv, v):
        B = self.B
        return np.einsum("nij,nj->ni", Dinv, v.reshape(-1, B)).reshape(-1)

    def _strong(self, A, eps):
        """plain_aggregates: strong(i,j) <=> j != i and eps^2 ||D_i|| ||D_j|| < ||A_ij||^2 (Frobenius)."""
        nb = A.shape[0] // self.B
        rows = np.repeat(np.arange(nb), np.diff(A.indptr))
        nrm = np.sqrt(np.sum(A.data ** 2, axis=(1, 2)))
        dn = np.sqrt(np.sum(self._diag_blocks(A) ** 2, axis=(1, 2)))
        return (A.indices != rows) & (eps * eps * dn[rows] * dn[A.indices] < nrm * nrm)

    def _plain_aggregates(self, A, strong):
        nb = A.shape[0] // self.B
        ptr, col = A.indptr, A.indices
        idv = np.full(nb, -2, np.int64)
        has = np.add.reduceat(strong.astype(np.int64), ptr[:-1]) if len(strong) else np.zeros(nb, np.int64)
        has[np.diff(ptr) == 0] = 0
        idv[has > 0] = -1
        count = 0
        for i in range(nb):
            if idv[i] != -1:
                continue
            cur = count
            count += 1
            idv[i] = cur
            neib = []
            for k in range(ptr[i], ptr[i + 1]):
                c = col[k]
                if strong[k] and idv[c] != -2:
                    idv[c] = cur
                    neib.append(c)
            for c in neib:
                for k in range(ptr[c], ptr[c + 1]):
                    cc = col[k]
                    if strong[k] and idv[cc] == -1:
                        idv[cc] = cur
        if count == 0:
            return idv, 0
        used = np.zeros(count, np.int64)
        used[idv[idv >= 0]] = 1
        ren = np.cumsum(used) - 1
        idv[idv >= 0] = ren[idv[idv >= 0]]
        return idv, int(used.sum())

    def _gershgorin(self, A, Dinv):
        nb = A.shape[0] // self.B
        nrm = np.sqrt(np.sum(A.data ** 2, axis=(1, 2)))
        s = np.add.reduceat(nrm, A.indptr[:-1])
        s[np.diff(A.indptr) == 0] = 0
        return float(np.max(s * np.sqrt(np.sum(Dinv ** 2, axis=(1, 2)))))

    def _prolongation(self, A, Dinv, strong, agg, nc, omega):
        """P = (I - omega D_f^-1 A_f) P_tent, block arithmetic; diagonal block -> (1 - omega) I, strong off-diagonal
        block A_ij -> -omega D_f^-1 A_ij into column agg(j); weak blocks are lumped into D_f."""
        B = self.B
        nb = A.shape[0] // B
        rows = np.repeat(np.arange(nb), np.diff(A.indptr))
        isdiag = A.indices == rows
        Df = np.zeros((nb, B, B))
        lump = isdiag | ~strong
        np.add.at(Df, rows[lump], A.data[lump])
        Dfinv = np.linalg.inv(Df)
        keep = (isdiag | strong) & (agg[A.indices] >= 0)
        r, j = rows[keep], A.indices[keep]
        blk = np.where(isdiag[keep][:, None, None], (1.0 - omega) * np.eye(B)[None], -omega * np.einsum("nij,njk->nik", Dfinv[r], A.data[keep]))
        # accumulate blocks with equal (row, aggregate)
        key = r * nc + agg[j]
        uk, inv = np.unique(key, return_inverse=True)
        data = np.zeros((len(uk), B, B))
        np.add.at(data, inv, blk)
        prow, pcol = uk // nc, uk % nc
        indptr = np.zeros(nb + 1, np.int64)
        np.add.at(indptr, prow + 1, 1)
        indptr = np.cumsum(indptr)
        P = sp.bsr_matrix((data, pcol, indptr), shape=(nb * B, nc * B))
        return P.tocsr()

    # ---- cycle (amgcl amg::cycle) and CG (amgcl solver::cg), identical to the scalar restatement
    def _cheb(self, L, rhs, x):
        d, c = L.d, L.c
        p = np.zeros_like(x)
        alpha = beta = 0.0
        for k in range(self.prm["degree"]):
            r = self._apply_dinv(L.Dinv, rhs - L.Acsr @ x)
            if k == 0:
                alpha, beta = 1.0 / d, 0.0
            elif k == 1:
                alpha = 2 * d * (1.0 / (2 * d * d - c * c))
                beta = alpha * d - 1
            else:
                alpha = 1.0 / (d - 0.25 * alpha * c * c)
                beta = alpha * d - 1
            p = alpha * r + beta * p
            x = x + p
        return x

    def _cycle(self, l, rhs, x):
        L = self.levels[l]
        if l + 1 == len(self.levels):
            for _ in range(self.prm["npre"] + self.prm["npost"]):
                x = self._cheb(L, rhs, x)
            return x
        for _ in range(self.prm["ncycle"]):
            for _ in range(self.prm["npre"]):
                x = self._cheb(L, rhs, x)
            t = rhs - L.Acsr @ x
            u = self._cycle(l + 1, L.R @ t, np.zeros(L.R.shape[0]))
            x = x + L.P @ u
            for _ in range(self.prm["npost"]):
                x = self._cheb(L, rhs, x)
        return x

    def apply(self, rhs):
        return self._cycle(0, np.asarray(rhs, float), np.zeros(len(rhs)))

    def cg(self, b, x0=None, tol=1e-10, maxiter=1000):
        A = self.levels[0].Acsr
        x = np.zeros(len(b)) if x0 is None else np.array(x0, float)
        nb = float(np.sqrt(b @ b))
        if nb < np.finfo(float).eps:
            return np.zeros(len(b)), 0, nb
        eps = max(tol * nb, np.finfo(float).tiny)
        r = b - A @ x
        rho1, rho2 = 2.0, 1.0
        rn = float(np.sqrt(r @ r))
        it = 0
        p = None
        while it < maxiter and rn > eps:
            s = self.apply(r)
            rho2, rho1 = rho1, float(r @ s)
            p = s + (rho1 / rho2) * p if it else s.copy()
            q = A @ p
            alpha = rho1 / float(q @ p)
            x += alpha * p
            r -= alpha * q
            rn = float(np.sqrt(r @ r))
            it += 1
        return x, it, rn / nb

    def level_sizes(self):
        return [(L.A.shape[0], int(L.Acsr.nnz if False else L.A.data.size)) for L in self.levels]
