"""Deterministic synthetic systems of BASELINE.json's configs, built with
numpy only (product-side generators: bench.py and smoke() must not depend on oracle/).

All matrices come back as raw Eigen-style CSC triples (outer int32[n+1], inner int32[nnz],
vals f64[nnz]) -- the layout that crosses the C ABI."""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed, n):
    """U(-1,1) stream from splitmix64 (stands in for Eigen's setRandom,
    reference tests/test_linear_solver.cpp:137-140). Bit-portable between C++ and numpy."""
    with np.errstate(over="ignore"):
        s = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * np.arange(1, n + 1, dtype=np.uint64)
        z = s
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return 2.0 * (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 1.0


def _stencil(dims, diag):
    """Dirichlet Laplacian on a lexicographic grid (x fastest). Symmetric, so CSC == CSR."""
    dims = list(dims)
    N = int(np.prod(dims))
    idx = np.arange(N, dtype=np.int64)
    strides = [1]
    for d in dims[:-1]:
        strides.append(strides[-1] * d)
    coords = [(idx // s) % d for s, d in zip(strides, dims)]
    # neighbour offsets in ascending column order: -s_k ... -s_0, 0, +s_0 ... +s_k
    offs = [(-s, c > 0) for s, c in reversed(list(zip(strides, coords)))]
    offs.append((0, np.ones(N, bool)))
    offs += [(s, c < d - 1) for s, c, d in zip(strides, coords, dims)]
    counts = np.zeros(N, np.int64)
    for _, m in offs:
        counts += m
    outer = np.zeros(N + 1, np.int64)
    np.cumsum(counts, out=outer[1:])
    nnz = int(outer[-1])
    inner = np.empty(nnz, np.int32)
    vals = np.empty(nnz, np.float64)
    pos = outer[:-1].copy()
    for off, m in offs:
        p = pos[m]
        inner[p] = (idx[m] + off).astype(np.int32)
        vals[p] = diag if off == 0 else -1.0
        pos[m] += 1
    return outer.astype(np.int32), inner, vals


def poisson2d(n):
    """C1: 2-D 5-point Laplacian n x n, diag 4 / off -1 (N = n^2)."""
    return _stencil((n, n), 4.0)


def poisson3d(n):
    """C2/C3: 3-D 7-point Laplacian n^3, diag 6 / off -1 (N = n^3; n = 216 -> 10,077,696 DoF)."""
    return _stencil((n, n, n), 6.0)


def spmv_csr(ptr, col, val, x):
    """Reference-free numpy SpMV for building b = A x* (symmetric matrices: CSC arrays work as CSR)."""
    prod = val * x[col]
    out = np.add.reduceat(prod, ptr[:-1].astype(np.int64))
    out[ptr[1:] == ptr[:-1]] = 0.0
    return out


def spmv_bytes(n, nnz):
    """Algorithmic bytes of one fp64/int32 CSR SpMV: 12 nnz + 20 N + 4."""
    return 12 * nnz + 20 * n + 4


def pcg_iter_bytes(n, nnz):
    """Compulsory traffic of one fused Jacobi-PCG iteration: B_spmv + 88 N."""
    return spmv_bytes(n, nnz) + 88 * n


class QuarticSpringGrid3D:
    """A polysolve::nonlinear::Problem on the C2 grid: nodes of an n^3 Dirichlet grid joined by nonlinear springs,
        E(u) = sum_edges phi(u_i - u_j) - f.u,   phi(d) = d^2/2 + kappa d^4/4   (edges to the boundary use u = 0).
    Convex, so the Hessian is SPD with the 7-point pattern of poisson3d(n) and values that change every Newton step --
    the analyze-once / factorize-many protocol of Newton.cpp:189-204 at any size (n = 100 -> 1,000,000 DoF).
    Duck-types polysolve_b200.nonlinear.Problem."""

    def __init__(self, n, kappa=10.0, seed=3):
        self.n, self.kappa = n, float(kappa)
        self.N = n ** 3
        self.outer, self.inner, _ = poisson3d(n)
        idx = np.arange(self.N, dtype=np.int64)
        strides = [1, n, n * n]
        coords = [(idx // s) % n for s in strides]
        # same column order as _stencil: -s2, -s1, -s0, 0, +s0, +s1, +s2
        self.dirs = [(-s, c > 0) for s, c in reversed(list(zip(strides, coords)))] + [(0, None)] + \
                    [(s, c < n - 1) for s, c in zip(strides, coords)]
        counts = np.zeros(self.N, np.int64)
        for off, m in self.dirs:
            counts += 1 if m is None else m
        pos = np.zeros(self.N + 1, np.int64)
        np.cumsum(counts, out=pos[1:])
        assert int(pos[-1]) == int(self.outer[-1])
        cur = pos[:-1].copy()
        self.slots = []
        for off, m in self.dirs:
            mm = np.ones(self.N, bool) if m is None else m
            self.slots.append(cur[mm].copy())
            cur[mm] += 1
        self.f = 0.5 + 0.5 * splitmix64(seed, self.N)

    def _diffs(self, x):
        """d[k][i] = u_i - u_neighbour for the 6 directions (neighbour value 0 outside the grid)."""
        out = []
        for off, m in self.dirs:
            if off == 0:
                continue
            nb = np.zeros(self.N)
            nb[m] = x[np.nonzero(m)[0] + off]
            out.append(x - nb)
        return out

    def value(self, x):
        e = 0.0
        for (off, m), d in zip([d for d in self.dirs if d[0] != 0], self._diffs(x)):
            w = np.where(m, 0.5, 1.0)  # interior edges are seen from both ends
            e += float(np.sum(w * (0.5 * d * d + 0.25 * self.kappa * d ** 4)))
        return e - float(self.f @ x)

    def gradient(self, x):
        g = -self.f.copy()
        for d in self._diffs(x):
            g += d + self.kappa * d ** 3
        return g

    def hessian(self, x, project_to_psd=False):
        import scipy.sparse as sp
        vals = np.zeros(int(self.outer[-1]))
        diag = np.zeros(self.N)
        k = 0
        for (off, m), slot in zip(self.dirs, self.slots):
            if off == 0:
                diag_slot = slot
                continue
            d = self._diffs_one(x, off, m)
            h = 1.0 + 3.0 * self.kappa * d * d
            diag += h
            vals[slot] = -h[m]
            k += 1
        vals[diag_slot] = diag
        return sp.csc_matrix((vals, self.inner, self.outer), shape=(self.N, self.N))

    def _diffs_one(self, x, off, m):
        nb = np.zeros(self.N)
        nb[m] = x[np.nonzero(m)[0] + off]
        return x - nb

    # hooks of Problem.hpp with their default behaviour
    def solution_changed(self, x):
        pass

    def is_step_valid(self, x0, x1):
        return True

    def max_step_size(self, x0, x1):
        return 1.0

    def line_search_begin(self, x0, x1):
        pass

    def line_search_end(self):
        pass

    def post_step(self, iteration, x, grad):
        pass

    def stop(self, x):
        return False


# ----------------------------------------------------------------------------------------- elasticity (C4 / C5 family)
def _kuhn_tets():
    """The 6 tetrahedra of the Kuhn split of the unit cube: paths 000 -> 111 along a permutation of the axes."""
    import itertools
    tets = []
    for perm in itertools.permutations(range(3)):
        v = [np.zeros(3, int)]
        for ax in perm:
            w = v[-1].copy()
            w[ax] += 1
            v.append(w)
        tets.append(np.array(v))
    return tets


def _p1_elastic_element(X, E, nu):
    """12 x 12 stiffness of a P1 tetrahedron with vertices X (4 x 3), isotropic linear elasticity (Voigt)."""
    M = np.hstack([np.ones((4, 1)), X.astype(float)])
    G = np.linalg.inv(M)[1:, :]          # G[:, a] = grad of the a-th barycentric function
    vol = abs(np.linalg.det(M)) / 6.0
    Bm = np.zeros((6, 12))
    for a in range(4):
        gx, gy, gz = G[:, a]
        Bm[:, 3 * a:3 * a + 3] = [[gx, 0, 0], [0, gy, 0], [0, 0, gz], [gy, gx, 0], [0, gz, gy], [gz, 0, gx]]
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    mu = E / (2 * (1 + nu))
    D = np.zeros((6, 6))
    D[:3, :3] = lam
    D[:3, :3] += 2 * mu * np.eye(3)
    D[3:, 3:] = mu * np.eye(3)
    return vol * Bm.T @ D @ Bm


def elasticity3d(m, E=1.0, nu=0.3):
    """C4: P1-tet (Kuhn 6-tet split) linear elasticity on an m^3-node unit-spacing grid, node-major 3-dof ordering,
    face x = 0 clamped as identity rows/columns (cf. reference FEMSolver.cpp:136-161). Returns the scalar CSC triple
    (symmetric, so CSC == CSR) and the right-hand side b = unit body force in -z on the free nodes.
    Assembled stencil-wise: for every (tet type, local vertex pair) one 3 x 3 block is added to a slab of nodes."""
    N = m ** 3
    tets = _kuhn_tets()
    blocks = {}   # offset (dx,dy,dz) -> [values (m,m,m,3,3) indexed [k,j,i], present (m,m,m) bool]

    def slab(o):  # nodes cell_origin + o over all cells, as slices of the (k, j, i) grid
        return tuple(slice(o[ax], o[ax] + m - 1) for ax in (2, 1, 0))

    for T in tets:
        K = _p1_elastic_element(T, E, nu)
        for a in range(4):
            for b in range(4):
                d = tuple(int(t) for t in (T[b] - T[a]))
                if d not in blocks:
                    blocks[d] = [np.zeros((m, m, m, 3, 3)), np.zeros((m, m, m), bool)]
                sl = slab(T[a])
                blocks[d][0][sl] += K[3 * a:3 * a + 3, 3 * b:3 * b + 3]
                blocks[d][1][sl] = True
    # clamp x = 0: rows and columns of those nodes become identity
    offs = sorted(blocks, key=lambda d: d[0] + m * d[1] + m * m * d[2])
    ii = np.arange(m)
    for d in offs:
        V, Pm = blocks[d]
        row_clamped = np.zeros((m, m, m), bool)
        row_clamped[:, :, 0] = True
        col_clamped = np.zeros((m, m, m), bool)
        col_i = ii + d[0]
        col_clamped[:, :, (col_i == 0)] = True
        kill = row_clamped | col_clamped
        if d == (0, 0, 0):
            V[row_clamped] = np.eye(3)
        else:
            Pm &= ~kill
            V[kill] = 0.0
    # block CSR in ascending column order
    present = np.stack([blocks[d][1].reshape(N) for d in offs], axis=1)            # N x noff
    cnt = present.sum(axis=1)
    brp = np.zeros(N + 1, np.int64)
    np.cumsum(cnt, out=brp[1:])
    node = np.arange(N, dtype=np.int64)
    coloff = np.array([d[0] + m * d[1] + m * m * d[2] for d in offs], np.int64)
    rows_idx, off_idx = np.nonzero(present)                                        # row-major: ascending offsets per row
    bci = node[rows_idx] + coloff[off_idx]
    allv = np.stack([blocks[d][0].reshape(N, 3, 3) for d in offs], axis=1)          # N x noff x 3 x 3
    bva = allv[rows_idx, off_idx]                                                   # nnzb x 3 x 3
    del allv
    nnzb = bci.shape[0]
    # scalar CSR: row 3 I + r = blocks of I in order, each contributing columns 3 J .. 3 J + 2
    len3 = 3 * cnt
    ptr = np.zeros(3 * N + 1, np.int64)
    np.cumsum(np.repeat(len3, 3), out=ptr[1:])
    col = np.empty(9 * nnzb, np.int32)
    val = np.empty(9 * nnzb, np.float64)
    blk_row = rows_idx                                                              # block -> node
    q = np.arange(nnzb, dtype=np.int64) - brp[blk_row]                              # position of the block inside its row
    for r in range(3):
        base = ptr[3 * blk_row + r] + 3 * q
        for c in range(3):
            col[base + c] = (3 * bci + c).astype(np.int32)
            val[base + c] = bva[:, r, c]
    b = np.zeros(3 * N)
    free = np.ones((m, m, m), bool)
    free[:, :, 0] = False
    b[2::3][free.reshape(N)] = -1.0
    return ptr.astype(np.int32), col, val, b
