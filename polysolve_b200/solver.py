"""Host-side mirror of polysolve::linear::Solver for the "CUDA" backend.

Same method names, argument meaning and error behaviour as the reference interface
(reference src/polysolve/linear/Solver.hpp:31-132); every call forwards to the C ABI
(include/psb200.h) exactly as adapter/CUDASolver.cpp does on the C++ side. Matrices are
scipy.sparse CSC (the layout of StiffnessMatrix, reference src/polysolve/Types.hpp:11-15)."""
import ctypes as C
import json

import numpy as np

from . import _lib


class Solver:
    """polysolve::linear::Solver, "CUDA" flavour."""

    @staticmethod
    def available_solvers():
        # reference Solver.cpp:501-568 (we only provide the new entry)
        return ["CUDA"]

    @staticmethod
    def create(solver="CUDA", precond=""):
        """Solver::create(name, precond) -- reference Solver.cpp:307-496."""
        if isinstance(solver, dict):
            params = solver
            name = params.get("solver", "CUDA")
            if isinstance(name, (list, tuple)):
                # priority list (reference Solver.cpp:92-114 select_valid_solver)
                name = next((s for s in name if s in Solver.available_solvers()), None)
            if name != "CUDA":
                raise RuntimeError(f"Unrecognized solver type: {name}")
            s = Solver()
            s.set_parameters(params)
            return s
        if solver != "CUDA":
            raise RuntimeError(f"Unrecognized solver type: {solver}")  # Solver.cpp:495
        s = Solver()
        if precond:
            pmap = {"Eigen::DiagonalPreconditioner": "jacobi", "Eigen::IdentityPreconditioner": "none",
                    "jacobi": "jacobi", "amg": "amg", "none": "none"}
            if precond not in pmap:
                raise RuntimeError(f"Unrecognized preconditioner: {precond}")
            s.set_parameters({"CUDA": {"precond": pmap[precond]}})
        return s

    def __init__(self, _borrowed=None):
        self._L = _lib.lib()
        self._owned = _borrowed is None
        if _borrowed is not None:
            self._h = C.c_void_p(_borrowed)  # a handle owned by somebody else (the Newton driver's linear solvers)
        else:
            self._h = C.c_void_p()
            rc = self._L.psb200_create(C.byref(self._h), None)
            if rc:
                raise RuntimeError(self._L.psb200_last_error(None).decode())
        self._keep = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and getattr(self, "_owned", False):
            self._L.psb200_destroy(h)
        self._h = None

    def _check(self, rc):
        if rc:
            # the reference's error convention: std::runtime_error (Utils.cpp:65-69)
            raise RuntimeError(self._L.psb200_last_error(self._h).decode())

    # ---- polysolve::linear::Solver virtuals
    def set_parameters(self, params):
        self._check(self._L.psb200_set_parameters(self._h, json.dumps(params).encode()))

    def set_tolerance(self, tol):
        self._check(self._L.psb200_set_tolerance(self._h, float(tol)))

    def set_block_size(self, bs):
        self._check(self._L.psb200_set_block_size(self._h, int(bs)))

    @staticmethod
    def _csc(A):
        import scipy.sparse as sp
        if not sp.isspmatrix_csc(A):
            A = sp.csc_matrix(A)
        if not A.has_canonical_format:
            # StiffnessMatrix is a compressed Eigen matrix: row indices ascending inside a column, no duplicates
            A = A.copy()
            A.sum_duplicates()
        outer = np.ascontiguousarray(A.indptr, np.int32)
        inner = np.ascontiguousarray(A.indices, np.int32)
        vals = np.ascontiguousarray(A.data, np.float64)
        return A.shape[0], outer, inner, vals

    def analyze_pattern(self, A, precond_num):
        n, outer, inner, _ = self._csc(A)
        self.analyze_pattern_raw(n, outer, inner, precond_num)

    @staticmethod
    def _raw_sizes(n, outer, inner, vals=None):
        """The C ABI takes plain pointers and sizes: check here that the arrays are as long as those sizes say."""
        if n < 0 or len(outer) < n + 1:
            raise RuntimeError("outer must hold n + 1 column pointers")
        nnz = int(outer[n])
        if nnz < 0 or len(inner) < nnz or (vals is not None and len(vals) < nnz):
            raise RuntimeError("inner / values are shorter than outer[n] entries")
        return nnz

    def analyze_pattern_raw(self, n, outer, inner, precond_num):
        nnz = self._raw_sizes(n, outer, inner)
        self._check(self._L.psb200_analyze_pattern_csc(self._h, n, nnz, outer, inner, int(precond_num)))

    def factorize(self, A):
        n, outer, inner, vals = self._csc(A)
        self.factorize_raw(n, outer, inner, vals)

    def factorize_raw(self, n, outer, inner, vals):
        nnz = self._raw_sizes(n, outer, inner, vals)
        self._check(self._L.psb200_factorize_csc(self._h, n, nnz, outer, inner, vals))

    def factorize_device(self, n, nnz, vals_ptr, diag_shift=0.0):
        """factorize() with the values already in GPU memory (CSC order of the analyzed pattern); diag_shift is
        RegularizedNewton's reg_weight (reference Newton.cpp:287-290)."""
        self._check(self._L.psb200_factorize_csc_device(self._h, int(n), int(nnz), vals_ptr, float(diag_shift)))

    def residual_norm_device(self, x_ptr, b_ptr, n):
        """||A x - b||_2 for device-resident x, b (reference Newton.cpp:207)."""
        out = C.c_double()
        self._check(self._L.psb200_residual_norm_device(self._h, x_ptr, b_ptr, int(n), C.byref(out)))
        return out.value

    def solve(self, b, x):
        """x is in/out (initial guess), as in the reference (Solver.hpp:119-128)."""
        b = np.ascontiguousarray(b, np.float64)
        if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous):
            raise TypeError("x must be a contiguous float64 numpy array (it is updated in place)")
        if b.shape != x.shape:
            raise RuntimeError("psb200_solve: size mismatch")
        self._check(self._L.psb200_solve(self._h, b, x, b.shape[0]))
        return x

    def solve_device(self, b_ptr, x_ptr, n):
        """b, x resident on the solver's GPU (raw device pointers, e.g. torch.Tensor.data_ptr())."""
        self._check(self._L.psb200_solve_device(self._h, b_ptr, x_ptr, n))

    # ---- FEMSolver helpers (reference src/polysolve/linear/FEMSolver.cpp:97-372), Dirichlet masking + lifting on the GPU
    @staticmethod
    def _nodes(nodes):
        nodes = np.ascontiguousarray(nodes, np.int32)
        return nodes if nodes.size else np.zeros(1, np.int32), int(nodes.size)

    def dirichlet_solve(self, A, f, dirichlet_nodes, u, precond_num):
        """dirichlet_solve(solver, A, f, dirichlet_nodes, u, precond_num): f becomes g, u the solution (both in place)."""
        n, outer, inner, vals = self._csc(A)
        nodes, cnt = self._nodes(dirichlet_nodes)
        for v in (f, u):
            if not (isinstance(v, np.ndarray) and v.dtype == np.float64 and v.flags.c_contiguous and v.shape == (n,)):
                raise TypeError("f and u must be contiguous float64 arrays of length n (they are updated in place)")
        self._check(self._L.psb200_dirichlet_solve(self._h, n, int(outer[n]), outer, inner, vals, f, nodes, cnt, u, int(precond_num)))

    def prefactorize(self, A, dirichlet_nodes, precond_num):
        n, outer, inner, vals = self._csc(A)
        nodes, cnt = self._nodes(dirichlet_nodes)
        self._check(self._L.psb200_dirichlet_prefactorize(self._h, n, int(outer[n]), outer, inner, vals, nodes, cnt, int(precond_num)))

    def dirichlet_solve_prefactorized(self, A, f, u):
        """A: the matrix to lift with (None = the resident masked matrix)."""
        vals = None
        if A is not None:
            vals = self._csc(A)[3]
        self._check(self._L.psb200_dirichlet_solve_prefactorized(self._h, None if vals is None else vals.ctypes.data, f, u, f.shape[0]))

    def get_info(self):
        need = C.c_size_t()
        buf = C.create_string_buffer(1 << 16)
        rc = self._L.psb200_get_info(self._h, buf, len(buf), C.byref(need))
        if rc and need.value > len(buf):
            buf = C.create_string_buffer(need.value)
            rc = self._L.psb200_get_info(self._h, buf, len(buf), C.byref(need))
        self._check(rc)
        return json.loads(buf.value.decode())

    def name(self):
        return self._L.psb200_name(self._h).decode()

    def release_cached_memory(self):
        """Returns the unused part of the GPU's stream-ordered memory pool to the driver."""
        self._check(self._L.psb200_release_cached_memory(self._h))

    def is_dense(self):
        return False

    # ---- multi-GPU (one process per GPU): row-partitioned Jacobi-PCG over NVLink peer memory
    def dist_prepare(self, rank, world, halo_cap=1 << 20):
        """Allocates this rank's comm buffer; returns its 64-byte CUDA IPC handle."""
        buf = C.create_string_buffer(64)
        self._check(self._L.psb200_dist_prepare(self._h, rank, world, halo_cap, buf))
        return buf.raw

    def dist_connect(self, handles):
        """handles: the 64-byte handles of all ranks, concatenated in rank order."""
        self._check(self._L.psb200_dist_connect(self._h, bytes(handles)))

    def dist_setup_torch(self, halo_cap=1 << 20):
        """Plumbing through torch.distributed (any backend): all-gather the IPC handles and connect."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        mine = self.dist_prepare(rank, world, halo_cap)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        self.dist_connect(b"".join(gathered))
        dist.barrier()
        return rank, world

    def dist_allgather(self, x_full):
        """Every rank contributes its rows of x_full (in place) and receives everybody's. No-op on one GPU."""
        self._check(self._L.psb200_dist_allgather(self._h, x_full, x_full.shape[0]))

    def residual_norm(self, x, b):
        """||A x - b||_2 with full-length host vectors (global norm on a row partition)."""
        r = C.c_double()
        self._check(self._L.psb200_residual_norm(self._h, np.ascontiguousarray(x, np.float64), np.ascontiguousarray(b, np.float64), x.shape[0], C.byref(r)))
        return r.value

    def dist_reset(self):
        """Collective recovery after a communication timeout (barrier on the host before and after)."""
        self._check(self._L.psb200_dist_reset(self._h))

    def dist_local_range(self):
        a, b = C.c_int64(), C.c_int64()
        self._check(self._L.psb200_dist_local_range(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @staticmethod
    def dist_plan_host(n, outer, inner, rank, world, halo_cap=1 << 20, align=1):
        """Host-only partition / halo plan of one rank (no GPU needed)."""
        L = _lib.lib()
        if len(outer) != n + 1 or not 0 <= int(outer[n]) <= len(inner):
            raise RuntimeError("psb200_dist_plan_host: outer must hold n + 1 entries and outer[n] <= len(inner)")
        nnz = int(outer[n])
        offsets = np.zeros(world + 1, np.int64)
        counts = np.zeros(3, np.int64)
        rp = np.zeros(n + 1, np.int32)
        ci = np.zeros(max(nnz, 1), np.int32)
        perm = np.zeros(max(nnz, 1), np.int32)
        send_begin = np.zeros(world + 1, np.int32)
        # a row is sent to every rank that has an entry in its column: at most (world - 1) * n and at most one per
        # matrix entry (align per entry when nodes travel whole)
        send_rows = np.zeros(max(min((world - 1) * n, align * nnz), 1), np.int32)
        recv_count = np.zeros(world, np.int32)
        halo_cols = np.zeros(max(n, 1), np.int32)
        rc = L.psb200_dist_plan_host_aligned(n, nnz, outer, inner, rank, world, halo_cap, align, offsets, counts, rp, ci, perm,
                                     send_begin, send_rows, recv_count, halo_cols)
        if rc:
            raise RuntimeError("psb200_dist_plan_host failed")
        nl, lnnz, nh = (int(c) for c in counts)
        return dict(offsets=offsets, n_local=nl, rp=rp[:nl + 1], ci=ci[:lnnz], perm=perm[:lnnz], send_begin=send_begin,
                    send_rows=send_rows[:int(send_begin[-1])], recv_count=recv_count, halo_cols=halo_cols[:nh])

    # ---- test / bench hooks
    def debug_get_csr(self, n, nnz):
        rp = np.empty(n + 1, np.int32)
        ci = np.empty(max(nnz, 1), np.int32)
        perm = np.empty(max(nnz, 1), np.int32)
        self._check(self._L.psb200_debug_get_csr(self._h, rp, ci, perm))
        return rp, ci[:nnz], perm[:nnz]

    def spmv(self, x):
        x = np.ascontiguousarray(x, np.float64)
        y = np.empty_like(x)
        self._check(self._L.psb200_spmv(self._h, x, y, x.shape[0]))
        return y

    def bench_spmv(self, reps=20, kernel=""):
        ms = C.c_double()
        self._check(self._L.psb200_bench_spmv(self._h, kernel.encode(), reps, C.byref(ms)))
        return ms.value

    def stream(self):
        return self._L.psb200_get_stream(self._h)

    def debug_set_aggregates(self, level, agg):
        agg = np.ascontiguousarray(agg, np.int32)
        self._keep = agg
        self._check(self._L.psb200_debug_set_aggregates(self._h, level, agg.ctypes.data, agg.shape[0]))

    def debug_get_level(self, level, which):
        w = {"A": 0, "P": 1, "R": 2}[which]
        rows, cols, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self._L.psb200_debug_get_level(self._h, level, w, C.byref(rows), C.byref(cols), C.byref(nnz), None, None, None))
        rp = np.empty(rows.value + 1, np.int32)
        ci = np.empty(max(nnz.value, 1), np.int32)
        va = np.empty(max(nnz.value, 1), np.float64)
        self._check(self._L.psb200_debug_get_level(self._h, level, w, None, None, None, rp.ctypes.data, ci.ctypes.data, va.ctypes.data))
        return rows.value, cols.value, rp, ci[:nnz.value], va[:nnz.value]

    def debug_get_aggregates(self, level, n):
        agg = np.empty(n, np.int32)
        na = C.c_int64()
        self._check(self._L.psb200_debug_get_aggregates(self._h, level, agg.ctypes.data, n, C.byref(na)))
        return agg, na.value

    def precond_apply(self, r):
        r = np.ascontiguousarray(r, np.float64)
        z = np.empty_like(r)
        self._check(self._L.psb200_precond_apply(self._h, r, z, r.shape[0]))
        return z
