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
