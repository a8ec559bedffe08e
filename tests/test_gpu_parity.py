"""GPU parity tests (run with -m gpu on a B200). Every test calls the product through the C ABI
(polysolve_b200.Solver -> libpsb200.so) and compares with the CPU oracle on the same seeded inputs,
with the committed golden fixtures, or through size-independent properties at BASELINE sizes.

Bars: bit-exact for the analyze_pattern integer arrays; floating point within the tolerance written
next to each assertion (the reference's own acceptance is ||Ax-b|| < 1e-8,
reference tests/test_linear_solver.cpp:160-162)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def csc(o, i, v):
    n = len(o) - 1
    return sp.csc_matrix((v, i, o), shape=(n, n))


def make(psb, **kw):
    s = psb.Solver.create("CUDA", "")
    if kw:
        s.set_parameters({"CUDA": kw})
    return s


# ------------------------------------------------------------------ analyze_pattern: bit-exact integers
@pytest.mark.parametrize("case", ["poisson2d", "poisson3d", "convdiff", "random_unsym", "random_ragged"])
def test_analyze_pattern_bit_exact(psb, orc, case):
    rng = np.random.default_rng(3)
    if case == "poisson2d":
        o, i, v = orc.poisson2d(32)
    elif case == "poisson3d":
        o, i, v = orc.poisson3d(20)
    elif case == "convdiff":
        o, i, v = orc.convdiff2d(32, 0.5)
    else:
        n = 700
        A = sp.random(n, n, density=0.01, random_state=rng, format="csc")
        if case == "random_ragged":
            # empty rows/columns, one dense row and one dense column
            A = sp.lil_matrix(A)
            A[5, :] = 0
            A[:, 9] = 0
            A[17, :] = rng.standard_normal(n)
            A[:, 23] = rng.standard_normal((n, 1))
            A = sp.csc_matrix(A)
            A.eliminate_zeros()
        else:
            A = sp.csc_matrix(A + sp.eye(n))
        A.sort_indices()
        o, i, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    n = len(o) - 1
    s = make(psb)
    s.analyze_pattern_raw(n, o, i, n)
    rp, ci, perm = s.debug_get_csr(n, int(o[-1]))
    rp0, ci0, perm0 = orc.csc_to_csr(n, o, i)
    assert np.array_equal(rp, rp0)
    assert np.array_equal(ci, ci0)
    assert np.array_equal(perm, perm0)
    # and SpMV through the product kernel == oracle SpMV (<= 4 ulp-scaled: different summation order only)
    s.factorize_raw(n, o, i, v)
    x = orc.splitmix64(7, n)
    y = s.spmv(x)
    y0 = orc.spmv_csc(o, i, v, x)
    scale = orc.spmv_csc(o, i, np.abs(v), np.abs(x)) + 1e-300
    assert np.max(np.abs(y - y0) / scale) < 4 * np.finfo(float).eps


def test_analyze_pattern_is_idempotent(psb, orc):
    o, i, v = orc.poisson3d(16)
    n = 16 ** 3
    s = make(psb)
    s.analyze_pattern_raw(n, o, i, n)
    s.factorize_raw(n, o, i, v)
    assert not s.get_info()["analyze_skipped"]
    s.analyze_pattern_raw(n, o, i, n)  # Newton does this every iteration (Newton.cpp:189)
    assert s.get_info()["analyze_skipped"]
    o2, i2, _ = orc.poisson3d(15)
    s.analyze_pattern_raw(15 ** 3, o2, i2, 15 ** 3)
    assert not s.get_info()["analyze_skipped"]


# ------------------------------------------------------------------ SpMV schedules
@pytest.mark.parametrize("kernel", ["stream", "scalar", "vector4", "vector8", "vector32"])
def test_spmv_schedules_match_oracle(psb, orc, kernel):
    o, i, v = orc.poisson3d(40)
    n = 40 ** 3
    v = v * (1.0 + 0.1 * orc.splitmix64(11, len(v)))  # non-trivial values
    s = make(psb, spmv_kernel=kernel)
    s.factorize_raw(n, o, i, v)
    assert s.get_info()["spmv_kernel"] == {"scalar": "vector1"}.get(kernel, kernel)
    x = orc.splitmix64(5, n)
    y = s.spmv(x)
    rp, ci, perm = orc.csc_to_csr(n, o, i)
    y0 = orc.spmv_csr(rp, ci, v[perm], x)
    scale = orc.spmv_csr(rp, ci, np.abs(v[perm]), np.abs(x))
    assert np.max(np.abs(y - y0) / scale) < 4 * np.finfo(float).eps


@pytest.mark.parametrize("kernel", ["auto", "stream2", "stream4", "stream8", "stream16", "vector16"])
def test_spmv_wide_stream_long_rows(psb, orc, kernel):
    """Rows of ~25 nnz (a Galerkin-like operator): the TMA stream schedule with several lanes per row, every
    lane count, against scipy; auto must pick a multi-lane stream schedule."""
    o, i, v = orc.poisson3d(24)
    n = 24 ** 3
    A = csc(o, i, v * (1.0 + 0.1 * orc.splitmix64(11, len(v))))
    A2 = sp.csc_matrix(A @ A.T)
    A2.sort_indices()
    o2, i2, v2 = A2.indptr.astype(np.int32), A2.indices.astype(np.int32), A2.data.astype(np.float64)
    s = make(psb, spmv_kernel=kernel)
    s.factorize_raw(n, o2, i2, v2)
    name = s.get_info()["spmv_kernel"]
    if kernel == "auto":
        assert name.startswith("stream") and name != "stream", name
    else:
        assert name == kernel
    x = orc.splitmix64(5, n)
    y0 = A2 @ x
    scale = abs(A2) @ np.abs(x)
    assert np.max(np.abs(s.spmv(x) - y0) / scale) < 8 * np.finfo(float).eps
    # a ragged variant: every 7th row emptied, one very long row (tile overflow => direct-load path)
    L = sp.lil_matrix(A2)
    L[::7, :] = 0
    L[100, :] = 1.0
    B = sp.csc_matrix(L)
    B.eliminate_zeros()
    B.sort_indices()
    s.factorize_raw(n, B.indptr.astype(np.int32), B.indices.astype(np.int32), B.data.astype(np.float64))
    np.testing.assert_allclose(s.spmv(x), B @ x, rtol=1e-12, atol=1e-10)


def test_spmv_stream_falls_back_on_dense_tiles(psb, orc):
    """Tiles whose nnz exceed the shared-memory stage take the direct-load path."""
    rng = np.random.default_rng(5)
    n = 4096
    A = sp.random(n, n, density=0.004, random_state=rng, format="lil")
    A[300:340, :] = rng.standard_normal((40, n))  # 40 dense rows => one tile far above the cap
    A = sp.csc_matrix(A + sp.eye(n))
    A.sort_indices()
    o, i, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    s = make(psb, spmv_kernel="stream")
    s.factorize_raw(n, o, i, v)
    x = orc.splitmix64(9, n)
    y0 = A @ x
    np.testing.assert_allclose(s.spmv(x), y0, rtol=1e-12, atol=1e-12)


def test_spmv_eigenvector_full_size(psb):
    """BASELINE size (216^3, 10,077,696 DoF): A v = lambda v for an analytic eigenpair."""
    P = psb.problems
    n = 216
    o, i, v = P.poisson3d(n)
    k = np.arange(1, n + 1)
    sx, sy, sz = (np.sin(np.pi * m * k / (n + 1)) for m in (3, 7, 2))
    vec = np.einsum("i,j,k->kji", sx, sy, sz).ravel()
    lam = 4 * sum(np.sin(np.pi * m / (2 * (n + 1))) ** 2 for m in (3, 7, 2))
    s = make(psb)
    s.factorize_raw(n ** 3, o, i, v)
    assert s.get_info()["spmv_kernel"] == "stream"
    y = s.spmv(vec)
    assert np.max(np.abs(y - lam * vec)) < 1e-13 * 12  # |A|.|v| <= 12


# ------------------------------------------------------------------ Jacobi-PCG (Eigen ordering)
@pytest.mark.parametrize("graph", [True, False])
def test_c1_known_answer_and_golden(psb, orc, graph):
    """C1 of BASELINE.json: 32x32 Poisson, tol 1e-10 => 115 iterations, golden x."""
    g = np.load(os.path.join(GOLD, "c1_poisson2d_32.npz"))
    o, i, v = psb.problems.poisson2d(32)
    s = make(psb, tolerance=1e-10, max_iter=1000, use_graph=graph)
    s.analyze_pattern_raw(1024, o, i, 1024)
    s.factorize_raw(1024, o, i, v)
    x = np.zeros(1024)
    s.solve(g["b"], x)
    info = s.get_info()
    assert info["solver_iter"] == int(g["iters"]) == 115
    assert info["num_iterations"] == 115
    assert abs(info["solver_error"] - float(g["err"])) < 1e-13
    assert info["solver_status"] == "Converged"
    np.testing.assert_allclose(x, g["x"], rtol=0, atol=1e-12)          # oracle (Eigen restatement)
    np.testing.assert_allclose(x, g["x_direct"], rtol=0, atol=1e-9)    # independent direct solve
    assert np.linalg.norm(csc(o, i, v) @ x - g["b"]) < 1e-8           # the reference's acceptance bound
    # warm start: converged x in => 0 iterations, x untouched (test_linear_solver.cpp:432-450)
    x2 = x.copy()
    s.solve(g["b"], x2)
    assert s.get_info()["solver_iter"] == 0
    assert np.array_equal(x2, x)


@pytest.mark.parametrize("n,tol", [(24, 1e-8), (48, 1e-8)])
def test_pcg_iteration_parity_3d(psb, orc, n, tol):
    o, i, v = orc.poisson3d(n)
    N = n ** 3
    b = orc.spmv_csc(o, i, v, orc.splitmix64(42, N))
    x0, it0, err0, _ = orc.eigen_cg(o, i, v, b, tol=tol, max_iters=10000)
    s = make(psb, tolerance=tol, max_iter=10000)
    s.factorize_raw(N, o, i, v)
    x = np.zeros(N)
    s.solve(b, x)
    info = s.get_info()
    # same algorithm, different summation order: iteration counts within +-2 %
    assert abs(info["solver_iter"] - it0) <= max(1, 0.02 * it0)
    assert info["solver_error"] < tol
    assert np.linalg.norm(x - x0) / np.linalg.norm(x0) < 10 * tol
    assert np.linalg.norm(csc(o, i, v) @ x - b) / np.linalg.norm(b) < 2 * tol


@pytest.mark.parametrize("n,check_every", [(20, 16), (48, 6), (37, 2)])
def test_single_reduction_cg_matches_eigen_ordering(psb, orc, n, check_every):
    """krylov = cg1r (Chronopoulos-Gear single-reduction PCG: two kernels and one reduction per iteration) against the
    Eigen ordering: the same iterates in exact arithmetic -- iteration count within +-2 %, same x to the tolerance, same
    counting rule (warm start => 0 iterations), max_iter honoured."""
    o, i, v = orc.poisson3d(n)
    N = n ** 3
    v = v * (1.0 + 0.05 * orc.splitmix64(13, len(v)))
    v = 0.5 * (v + v[orc.csc_to_csr(N, o, i)[2]])  # keep it symmetric
    b = orc.splitmix64(42, N)
    out = {}
    for mode in ("cg1r", "cg"):
        s = make(psb, tolerance=1e-9, max_iter=5000, krylov=mode, check_every=check_every)
        s.factorize_raw(N, o, i, v)
        x = np.zeros(N)
        s.solve(b, x)
        info = s.get_info()
        assert info["krylov"] == mode and info["solver_status"] == "Converged" and info["solver_error"] < 1e-9
        out[mode] = (x, info["solver_iter"], info["solver_error"])
        x2 = x.copy()
        s.solve(b, x2)  # warm start => 0 iterations on both paths
        assert s.get_info()["solver_iter"] == 0 and np.array_equal(x2, x)
    assert abs(out["cg1r"][1] - out["cg"][1]) <= max(1, 0.02 * out["cg"][1])
    assert np.linalg.norm(out["cg1r"][0] - out["cg"][0]) / np.linalg.norm(out["cg"][0]) < 1e-7
    x0, it0, _, _ = orc.eigen_cg(o, i, v, b, tol=1e-9, max_iters=5000)
    assert abs(out["cg1r"][1] - it0) <= max(1, 0.02 * it0)
    assert np.linalg.norm(out["cg1r"][0] - x0) / np.linalg.norm(x0) < 1e-7
    assert np.linalg.norm(csc(o, i, v) @ out["cg1r"][0] - b) / np.linalg.norm(b) < 2e-9


def test_single_reduction_cg_matches_its_restatement(psb, orc):
    """krylov = cg1r on C1 against oracle/cg1r_oracle.py (the same recurrence in numpy, itself checked against the
    Eigen-ordering oracle in test_oracle.py): iterations within 2, x to 1e-9, reported error below the tolerance."""
    from oracle import cg1r_oracle

    o, i, v = orc.poisson2d(32)
    b = orc.splitmix64(42, 1024)
    A = csc(o, i, v).tocsr()
    xo, ito, erro, _ = cg1r_oracle.cg1r(A, b, dinv=1.0 / A.diagonal(), tol=1e-10, max_iters=1000)
    s = make(psb, tolerance=1e-10, max_iter=1000, krylov="cg1r")
    s.factorize_raw(1024, o, i, v)
    x = np.zeros(1024)
    s.solve(b, x)
    info = s.get_info()
    assert info["solver_status"] == "Converged" and abs(info["solver_iter"] - ito) <= 2
    assert info["solver_error"] < 1e-10 and np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-9


def test_single_reduction_cg_max_iter_zero_rhs_and_no_precond(psb, orc):
    o, i, v = orc.poisson3d(16)
    N = 16 ** 3
    b = orc.splitmix64(5, N)
    for mi in (1, 2, 3, 7, 8):
        s = make(psb, tolerance=1e-14, max_iter=mi, check_every=4, krylov="cg1r")
        s.factorize_raw(N, o, i, v)
        x = np.zeros(N)
        s.solve(b, x)
        info = s.get_info()
        assert info["solver_iter"] == mi and info["solver_status"] == "Reach max iterations"
    s = make(psb, tolerance=1e-10, krylov="cg1r")
    s.factorize_raw(N, o, i, v)
    x = orc.splitmix64(6, N)
    s.solve(np.zeros(N), x)   # Eigen: zero rhs => x = 0, 0 iterations
    assert s.get_info()["solver_iter"] == 0 and not x.any()
    s = make(psb, tolerance=1e-10, krylov="cg1r", precond="none", max_iter=2000)
    s.factorize_raw(N, o, i, v)
    x = np.zeros(N)
    s.solve(b, x)
    assert np.linalg.norm(csc(o, i, v) @ x - b) / np.linalg.norm(b) < 2e-10
    # constant diagonal: the unpreconditioned iteration takes the same steps as the Jacobi-preconditioned one
    _, it0, _, _ = orc.eigen_cg(o, i, v, b, tol=1e-10, max_iters=2000)
    assert abs(s.get_info()["solver_iter"] - it0) <= max(1, 0.02 * it0)


def test_set_parameters_after_factorize_invalidates_the_preconditioner(psb, orc):
    """Round-1 advisor finding: switching the preconditioner after factorize() must not keep solving with the old one."""
    o, i, v = orc.poisson3d(12)
    N = 12 ** 3
    b = orc.splitmix64(5, N)
    s = make(psb, tolerance=1e-10)
    s.factorize_raw(N, o, i, v)
    x = np.zeros(N)
    s.solve(b, x)
    it_jacobi = s.get_info()["solver_iter"]
    s.set_parameters({"CUDA": {"tolerance": 1e-9}})       # unrelated change: the factorization stays valid
    s.solve(b, np.zeros(N))
    s.set_parameters({"CUDA": {"precond": "none"}})
    with pytest.raises(RuntimeError, match="factorize"):
        s.solve(b, np.zeros(N))
    s.factorize_raw(N, o, i, v)
    x = np.zeros(N)
    s.solve(b, x)
    assert s.get_info()["precond"] == "none" and 0 < s.get_info()["solver_iter"] <= it_jacobi + 1  # looser tolerance now
    assert np.linalg.norm(csc(o, i, v) @ x - b) / np.linalg.norm(b) < 2e-9
    with pytest.raises(RuntimeError, match="spmv_kernel"):
        s.set_parameters({"CUDA": {"spmv_kernel": "streamX"}})
    s.solve(b, x)                                          # the rejected document left the solver usable


def test_pcg_identity_precond_and_max_iter(psb, orc):
    o, i, v = orc.poisson2d(32)
    b = orc.splitmix64(42, 1024)
    s = make(psb, tolerance=1e-10, precond="none")
    s.factorize_raw(1024, o, i, v)
    x = np.zeros(1024)
    s.solve(b, x)
    assert np.linalg.norm(csc(o, i, v) @ x - b) < 1e-8
    # max_iter cap: Eigen reports iterations() == maxIterations and does not throw
    s2 = make(psb, tolerance=1e-14, max_iter=10, check_every=4)
    s2.factorize_raw(1024, o, i, v)
    x = np.zeros(1024)
    s2.solve(b, x)
    info = s2.get_info()
    assert info["solver_iter"] == 10 and info["solver_status"] == "Reach max iterations"
    x0, it0, err0, _ = orc.eigen_cg(o, i, v, b, tol=1e-14, max_iters=10)
    assert it0 == 10
    np.testing.assert_allclose(x, x0, rtol=0, atol=1e-12)
    assert abs(info["solver_error"] - err0) < 1e-10


def test_zero_rhs_and_tiny_systems(psb, orc):
    o, i, v = orc.poisson2d(32)
    s = make(psb, tolerance=1e-10)
    s.factorize_raw(1024, o, i, v)
    x = orc.splitmix64(3, 1024)
    s.solve(np.zeros(1024), x)  # Eigen: ||b|| == 0 => x = 0, 0 iterations
    assert np.all(x == 0) and s.get_info()["solver_iter"] == 0
    # 1x1 and 3x3 systems (smaller than any tile / vector width)
    for n in (1, 2, 3):
        o, i, v = orc.poisson2d(n)
        N = n * n
        s = make(psb, tolerance=1e-12)
        s.factorize_raw(N, o, i, v)
        b = orc.splitmix64(1, N)
        x = np.zeros(N)
        s.solve(b, x)
        np.testing.assert_allclose(csc(o, i, v) @ x, b, atol=1e-12)


def test_pre_factor_protocol(psb, orc):
    """reference tests/test_linear_solver.cpp:241-307: analyze once, then 10x factorize(new values on
    the same pattern) + solve; every solve must reach ||Ax-b|| < 1e-8."""
    o, i, _ = orc.poisson2d(32)
    vals = orc.prefactor_values(o, i, rounds=10)
    s = make(psb, tolerance=1e-10)
    s.analyze_pattern_raw(1024, o, i, 1024)
    for r in range(10):
        b = orc.splitmix64(100 + r, 1024)
        x = np.zeros(1024)
        s.factorize_raw(1024, o, i, vals[r])
        s.solve(b, x)
        assert np.linalg.norm(csc(o, i, vals[r]) @ x - b) < 1e-8
        x0, it0, _, _ = orc.eigen_cg(o, i, vals[r], b, tol=1e-10)
        assert abs(s.get_info()["solver_iter"] - it0) <= 1
        np.testing.assert_allclose(x, x0, rtol=0, atol=1e-11)


def test_release_cached_memory_keeps_the_solver_usable(psb, orc):
    """Device buffers come from the stream-ordered pool; trimming it must not touch buffers in use."""
    o, i, v = orc.poisson3d(24)
    N = 24 ** 3
    b = orc.splitmix64(6, N)
    s = make(psb, tolerance=1e-9, precond="amg")
    s.factorize_raw(N, o, i, v)
    x = np.zeros(N)
    s.solve(b, x)
    it = s.get_info()["solver_iter"]
    s.release_cached_memory()
    x2 = np.zeros(N)
    s.solve(b, x2)
    assert s.get_info()["solver_iter"] == it and np.array_equal(x, x2)
    s.factorize_raw(N, o, i, v)      # rebuilds the hierarchy from a trimmed pool
    x3 = np.zeros(N)
    s.solve(b, x3)
    assert np.array_equal(x, x3)


def test_solve_device_pointers(psb, orc):
    import torch
    o, i, v = orc.poisson3d(24)
    N = 24 ** 3
    g = np.load(os.path.join(GOLD, "poisson3d_24.npz"))
    s = make(psb, tolerance=1e-8, max_iter=10000)
    s.factorize_raw(N, o, i, v)
    db = torch.from_numpy(g["b"]).cuda()
    dx = torch.zeros(N, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    s.solve_device(db.data_ptr(), dx.data_ptr(), N)
    x = dx.cpu().numpy()
    assert abs(s.get_info()["solver_iter"] - int(g["iters"])) <= 2
    assert np.linalg.norm(x - g["xstar"]) / np.linalg.norm(g["xstar"]) < 1e-6
    assert np.linalg.norm(x - g["x"]) / np.linalg.norm(g["x"]) < 1e-7


def test_device_resident_newton_step(psb, orc):
    """the Newton call sequence (Newton.cpp:173-214) with H, g and dx in GPU memory -- analyze once,
    factorize_csc_device(values, reg_weight), solve_device, residual_norm_device -- equals the host path."""
    import torch
    o, i, v = orc.poisson3d(20)
    N, nnz = 20 ** 3, len(v)
    v = v * (1.0 + 0.2 * orc.splitmix64(4, nnz))
    v = 0.5 * (v + v[orc.csc_to_csr(N, o, i)[2]])
    g = orc.splitmix64(8, N)
    s = make(psb, tolerance=1e-10, max_iter=5000)
    with pytest.raises(RuntimeError):
        s.factorize_device(N, nnz, 0, 0.0)  # no pattern yet
    s.analyze_pattern_raw(N, o, i, N)
    dv = torch.from_numpy(v).cuda()
    db = torch.from_numpy(-g).cuda()
    for shift in (0.0, 0.75):
        s.analyze_pattern_raw(N, o, i, N)            # Newton re-submits the pattern every iteration: skipped by hash
        assert s.get_info()["analyze_skipped"]
        s.factorize_device(N, nnz, dv.data_ptr(), shift)
        dx = torch.zeros(N, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        s.solve_device(db.data_ptr(), dx.data_ptr(), N)
        res = s.residual_norm_device(dx.data_ptr(), db.data_ptr(), N)
        x = dx.cpu().numpy()
        H = csc(o, i, v) + shift * sp.identity(N, format="csc")
        assert abs(res - np.linalg.norm(H @ x + g)) < 1e-12 * np.linalg.norm(g) + 1e-14
        assert res < 1e-9 * np.linalg.norm(g)
        # same iterates as the host entry points on the shifted matrix
        Hs = sp.csc_matrix(H)
        Hs.sort_indices()
        s2 = make(psb, tolerance=1e-10, max_iter=5000)
        s2.factorize_raw(N, Hs.indptr.astype(np.int32), Hs.indices.astype(np.int32), Hs.data.astype(np.float64))
        x2 = np.zeros(N)
        s2.solve(-g, x2)
        assert s2.get_info()["solver_iter"] == s.get_info()["solver_iter"]
        np.testing.assert_allclose(x, x2, rtol=0, atol=1e-13)
    with pytest.raises(RuntimeError):
        s.factorize_device(N, nnz - 1, dv.data_ptr(), 0.0)  # size differs from the analyzed pattern


# ------------------------------------------------------------------ BiCGSTAB (Eigen ordering)
def test_bicgstab_golden_unsymmetric(psb, orc):
    g = np.load(os.path.join(GOLD, "convdiff2d_32.npz"))
    o, i, v = orc.convdiff2d(32, 0.5)
    s = make(psb, krylov="bicgstab", tolerance=1e-10)
    s.analyze_pattern_raw(1024, o, i, 1024)
    assert s.get_info()["symmetric_pattern"]  # pattern symmetric, values not
    s.factorize_raw(1024, o, i, v)
    x = np.zeros(1024)
    s.solve(g["b"], x)
    info = s.get_info()
    # BiCGSTAB amplifies rounding differences; same algorithm => iteration count within a few
    assert abs(info["solver_iter"] - int(g["iters"])) <= 3
    assert info["solver_error"] < 1e-10
    np.testing.assert_allclose(x, g["x_direct"], rtol=0, atol=1e-8)
    assert np.linalg.norm(csc(o, i, v) @ x - g["b"]) < 1e-8
    x2 = x.copy()
    s.solve(g["b"], x2)
    assert s.get_info()["solver_iter"] == 0


def test_bicgstab_unsymmetric_pattern(psb, orc):
    rng = np.random.default_rng(11)
    n = 600
    A = sp.random(n, n, density=0.01, random_state=rng, format="csc")
    A = sp.csc_matrix(A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0))
    A.sort_indices()
    o, i, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    b = orc.splitmix64(2, n)
    s = make(psb, krylov="bicgstab", tolerance=1e-12)
    s.factorize_raw(n, o, i, v)
    assert not s.get_info()["symmetric_pattern"]
    x = np.zeros(n)
    s.solve(b, x)
    x0, it0, err0, _ = orc.eigen_bicgstab(o, i, v, b, tol=1e-12)
    assert abs(s.get_info()["solver_iter"] - it0) <= 2
    np.testing.assert_allclose(x, x0, rtol=0, atol=1e-10)
    assert np.linalg.norm(A @ x - b) < 1e-8


# ------------------------------------------------------------------ BASELINE-size property test
def test_pcg_full_size_manufactured_solution(psb):
    """C2 of BASELINE.json: 216^3 Poisson, b = A x*, Jacobi-PCG to 1e-8. Properties: converges in the
    ~3n iterations the oracle shows at smaller n, true residual and error are small."""
    P = psb.problems
    n = 216
    N = n ** 3
    o, i, v = P.poisson3d(n)
    xstar = P.splitmix64(42, N)
    b = P.spmv_csr(o, i, v, xstar)
    s = make(psb, tolerance=1e-8, max_iter=10000)
    s.analyze_pattern_raw(N, o, i, N)
    s.factorize_raw(N, o, i, v)
    x = np.zeros(N)
    s.solve(b, x)
    info = s.get_info()
    assert info["solver_status"] == "Converged"
    assert 400 <= info["solver_iter"] <= 560  # oracle (CPU, same input): see tests/golden/c2_oracle_iters.json
    r = P.spmv_csr(o, i, v, x) - b
    assert np.linalg.norm(r) / np.linalg.norm(b) < 2e-8
    assert np.linalg.norm(x - xstar) / np.linalg.norm(xstar) < 1e-5
