/* psb200 -- C ABI of the B200-native linear-solver backend for polysolve.
 *
 * This is the drop-in boundary: every entry point below is what a
 * polysolve::linear::Solver subclass ("CUDA", see adapter/CUDASolver.hpp) binds, one call per
 * virtual of the reference interface (reference src/polysolve/linear/Solver.hpp:90-131).
 * Plain pointers and sizes only; no C++/torch types. All functions return 0 on success and a
 * non-zero status otherwise; psb200_last_error() then holds the message. The C++ adapter turns a
 * non-zero status into std::runtime_error, the reference's error convention
 * (src/polysolve/Utils.cpp:65-69), so Newton's factorize fallback keeps working
 * (src/polysolve/nonlinear/descent_strategies/Newton.cpp:191-202).
 *
 * Matrices cross the boundary exactly as polysolve holds them: Eigen compressed-column
 * (StiffnessMatrix = Eigen::SparseMatrix<double, ColMajor, int>, src/polysolve/Types.hpp:11-15):
 * outer = outerIndexPtr() int32[n+1], inner = innerIndexPtr() int32[nnz], vals = valuePtr() f64[nnz].
 */
#ifndef PSB200_H
#define PSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct psb200_solver *psb200_handle;

#define PSB200_OK 0
#define PSB200_ERR_INVALID 1   /* bad argument / protocol violation */
#define PSB200_ERR_CUDA 2      /* CUDA runtime / kernel failure, out of memory */
#define PSB200_ERR_NUMERIC 3   /* zero / non-finite diagonal, breakdown detected in setup */
#define PSB200_ERR_COMM 4      /* NCCL / peer-memory failure */

/* Solver::create("CUDA", precond)  -- reference Solver.cpp:307-496 (MAS branch :402-404 is the template).
 * json_params may be NULL; otherwise the same document set_parameters() takes. */
int psb200_create(psb200_handle *out, const char *json_params);
int psb200_destroy(psb200_handle h);

/* Solver::set_parameters(const json&) -- Solver.hpp:93. Reads the "CUDA" object of the document:
 *   krylov: "cg"|"bicgstab"; precond: "jacobi"|"amg"|"none"; tolerance; max_iter; check_every;
 *   use_graph; spmv_kernel; device; amg: {max_levels, coarse_enough, ncycle, npre, npost, degree,
 *   power_iters, higher, lower, relax, eps_strong, ...} mirroring AMGCL.cpp:32-65.               */
int psb200_set_parameters(psb200_handle h, const char *json);
/* Solver::set_tolerance(double) -- Solver.hpp:116-117 (relative to ||b||, as Eigen/AMGCL). */
int psb200_set_tolerance(psb200_handle h, double tol);
/* Solver::set_block_size(int) -- Solver.hpp:110. */
int psb200_set_block_size(psb200_handle h, int block_size);

/* Solver::analyze_pattern(const StiffnessMatrix&, int precond_num) -- Solver.hpp:99.
 * Index work only: CSC -> CSR transpose map, SpMV tiling, (multi-GPU) row partition + halo lists.
 * Idempotent: an unchanged pattern (hash) is detected and skipped, because Newton calls this every
 * iteration (Newton.cpp:189). */
int psb200_analyze_pattern_csc(psb200_handle h, int64_t n, int64_t nnz, const int32_t *outer,
                               const int32_t *inner, int precond_num);
/* Solver::factorize(const StiffnessMatrix&) -- Solver.hpp:102. Values only when the pattern is
 * known (test "pre_factor", tests/test_linear_solver.cpp:241-307); builds the preconditioner
 * (Jacobi inverse diagonal / SA-AMG hierarchy). The matrix is borrowed only for this call. */
int psb200_factorize_csc(psb200_handle h, int64_t n, int64_t nnz, const int32_t *outer,
                         const int32_t *inner, const double *vals);
/* Solver::solve(b, x) -- Solver.hpp:119-128. x is in/out: its content is the initial guess
 * (tests/test_linear_solver.cpp:400-455 pin "converged guess => 0 iterations"). Host pointers. */
int psb200_solve(psb200_handle h, const double *b, double *x_inout, int64_t n);
/* Same, with b and x resident in device memory of the solver's GPU. */
int psb200_solve_device(psb200_handle h, const double *d_b, double *d_x_inout, int64_t n);

/* Device-resident Newton step (reference call site Newton::solve_sparse_linear_system,
 * Newton.cpp:173-214): the Hessian values are already in GPU memory, in the CSC order of the pattern given to
 * psb200_analyze_pattern_csc (which Newton re-submits unchanged every iteration, Newton.cpp:189). diag_shift is added
 * to every diagonal entry (RegularizedNewton: hessian += reg_weight * I, Newton.cpp:287-290; the diagonal must be
 * structurally present). Row partitions: d_vals is the full CSC value array on every rank. */
int psb200_factorize_csc_device(psb200_handle h, int64_t n, int64_t nnz, const double *d_vals, double diag_shift);
/* ||A x - b||_2 with x and b in device memory: the residual Newton checks after the solve
 * (objFunc.grad_norm(hessian * direction + grad), Newton.cpp:207, with b = -grad). Row partitions: local slices in,
 * the global norm out on every rank. */
int psb200_residual_norm_device(psb200_handle h, const double *d_x, const double *d_b, int64_t n, double *norm_out);
/* The same check with full-length HOST vectors; on a row partition every rank passes the full vectors and obtains the
 * global norm. */
int psb200_residual_norm(psb200_handle h, const double *x, const double *b, int64_t n, double *norm_out);

/* Solver::get_info(json&) -- Solver.hpp:96. Writes a JSON object with both key conventions:
 * "solver_iter","solver_error" (EigenSolver.tpp:88-89) and "num_iterations","final_res_norm"
 * (AMGCL.cpp:142-143), plus "solver_status" (MASSolver.cu:214-219) and timing/hierarchy details.
 * Returns PSB200_ERR_INVALID if cap is too small (required size in *needed when non-NULL). */
int psb200_get_info(psb200_handle h, char *json_out, size_t cap, size_t *needed);
/* Solver::name() -- Solver.hpp:131. Returns "CUDA". */
const char *psb200_name(psb200_handle h);
const char *psb200_last_error(psb200_handle h);
/* Device buffers come from the GPU's stream-ordered memory pool and stay cached there between factorize() calls
 * (policy precedent: one pool per solver, MASSolver.cu:154-156). This returns the cached, unused part to the driver. */
int psb200_release_cached_memory(psb200_handle h);

/* ---- Dirichlet pre-processing on the GPU: the reference's FEMSolver helpers, the step right before
 * the hot path in PolyFEM. N = diag(1 iff i is a Dirichlet dof).
 * psb200_dirichlet_solve = dirichlet_solve(solver, A, f, dirichlet_nodes, u, precond_num) with remove_zero_cols =
 * false (FEMSolver.cpp:97-300): A~ = A with the rows and columns of the Dirichlet dofs set to identity,
 * g = f - (I - N) A N f, analyze_pattern + factorize(A~) + solve(g, u); on return f holds g (":282 f = g") and u the
 * solution (u is also the initial guess). The host matrix is not modified (the reference rewrites A in place to save
 * host memory; here the masked copy lives in HBM). Single GPU. */
int psb200_dirichlet_solve(psb200_handle h, int64_t n, int64_t nnz, const int32_t *outer, const int32_t *inner,
                           const double *vals, double *f_inout, const int32_t *dirichlet_nodes, int64_t n_nodes,
                           double *u_inout, int precond_num);
/* prefactorize(solver, A, dirichlet_nodes, precond_num) -- FEMSolver.cpp:303-343. */
int psb200_dirichlet_prefactorize(psb200_handle h, int64_t n, int64_t nnz, const int32_t *outer, const int32_t *inner,
                                  const double *vals, const int32_t *dirichlet_nodes, int64_t n_nodes, int precond_num);
/* dirichlet_solve_prefactorized(solver, A, f, dirichlet_nodes, u) -- FEMSolver.cpp:345-372: g is computed with the
 * matrix values the caller passes (vals, CSC order of the prefactorized pattern); NULL = the resident masked matrix,
 * which is what the reference multiplies with when the caller hands back the A that prefactorize rewrote. */
int psb200_dirichlet_solve_prefactorized(psb200_handle h, const double *vals_or_null, double *f_inout, double *u_inout,
                                         int64_t n);

/* ---- multi-GPU (one process per GPU, up to 8 GPUs of one node). No reference counterpart.
 * The matrix is row-range partitioned (contiguous ranges balanced by nnz); halo x-entries and the
 * dot-product all-reduce travel over NVLink peer memory, fused into the solver's kernels.
 *   1. every rank: psb200_dist_prepare() -> a 64-byte CUDA IPC handle of its comm buffer
 *   2. the host application all-gathers the handles (torch.distributed / MPI), rank order
 *   3. every rank: psb200_dist_connect(all handles)
 * Afterwards analyze_pattern / factorize still receive the FULL CSC matrix on every rank (each rank
 * keeps its rows), psb200_solve receives full-length b / x (each rank reads and writes only its own
 * rows [row_begin, row_end)), psb200_solve_device receives the local slices. Every call is collective.
 * krylov = cg | cg1r (single-reduction CG: one all-reduce per iteration) with precond = jacobi | none, krylov = cg with
 * precond = amg. "amg": {"dist_mode": ...}: "partitioned" (default) = every level above amg.replicate_below non-zeros is
 * row-partitioned like the fine matrix (decoupled aggregation, rank-local P / R, distributed Galerkin product, per-level
 * halo exchange; scalar and block problems; device memory per rank ~ 1 / world), smaller levels are replicated;
 * "global" = one hierarchy of the whole matrix on every rank, fine level partitioned (scalar problems); "local" = a
 * rank-local hierarchy per GPU (block-Jacobi across ranks).
 * A wait for a peer that exceeds "comm_timeout_s" fails the call with PSB200_ERR_COMM and leaves the handle unusable
 * (every later call returns PSB200_ERR_COMM) until the application has synchronised the ranks on the host, every rank has
 * called psb200_dist_reset, and the ranks have synchronised again. */
int psb200_dist_prepare(psb200_handle h, int rank, int world, int64_t halo_cap_doubles, char handle_out[64]);
int psb200_dist_connect(psb200_handle h, const char *handles /* world * 64 bytes */);
int psb200_dist_reset(psb200_handle h);
/* Every rank contributes its own rows [row_begin, row_end) of x_full and receives everybody's (host vector of n values;
 * one fused all-gather over NVLink). After psb200_solve this turns the rank-local result into the whole solution on every
 * rank -- what a Newton driver that evaluates its Problem redundantly needs. No-op on a single GPU. */
int psb200_dist_allgather(psb200_handle h, double *x_full_inout, int64_t n);
int psb200_dist_local_range(psb200_handle h, int64_t *row_begin, int64_t *row_end);
/* Host-only plan of one rank (no GPU needed; used by the CPU tests to check partition offsets and halo
 * lists bit-exactly against the oracle). Caller-allocated arrays: offsets[world+1], counts[3] =
 * {local rows, local nnz, halo columns}, local_rp[n+1], local_ci[nnz], local_perm[nnz],
 * send_begin[world+1], send_rows[(world-1)*n] (a row can be sent to every other rank; send_begin[world] entries are
 * written), recv_count[world], halo_cols[n]. */
int psb200_dist_plan_host(int64_t n, int64_t nnz, const int32_t *outer, const int32_t *inner, int rank, int world,
                          int64_t halo_cap, int64_t *offsets, int64_t *counts, int32_t *local_rp, int32_t *local_ci,
                          int32_t *local_perm, int32_t *send_begin, int32_t *send_rows, int32_t *recv_count,
                          int32_t *halo_cols);
/* Same, with the row offsets rounded up to multiples of `align`: block problems (set_block_size(B), SURVEY 8e "for
 * block problems aligned to 3") keep the B rows of a node on one rank; align = 1 is psb200_dist_plan_host. */
int psb200_dist_plan_host_aligned(int64_t n, int64_t nnz, const int32_t *outer, const int32_t *inner, int rank, int world,
                                  int64_t halo_cap, int align, int64_t *offsets, int64_t *counts, int32_t *local_rp,
                                  int32_t *local_ci, int32_t *local_perm, int32_t *send_begin, int32_t *send_rows,
                                  int32_t *recv_count, int32_t *halo_cols);

/* ---- test / bench hooks (not part of the polysolve interface) */
/* CSR produced by analyze_pattern: row_ptr int32[n+1], col_idx int32[nnz], perm int32[nnz] with
 * vals_csr[k] = vals_csc[perm[k]]. Compared bit-exactly against the oracle's stable transpose. */
int psb200_debug_get_csr(psb200_handle h, int32_t *row_ptr, int32_t *col_idx, int32_t *perm);
/* y = A x through the product SpMV kernel (host buffers, full length). */
int psb200_spmv(psb200_handle h, const double *x, double *y, int64_t n);
/* reps launches of the product SpMV kernel on resident device vectors, timed with CUDA events on
 * the solver's stream; *ms_avg = average per launch. kernel: NULL/"" = the configured one. */
int psb200_bench_spmv(psb200_handle h, const char *kernel, int reps, double *ms_avg);
/* cudaStream_t the solver launches on (so callers can bracket it with their own events). */
void *psb200_get_stream(psb200_handle h);
/* AMG hooks: impose aggregates for level `level` before factorize (parity vs the oracle's greedy
 * aggregation); read back hierarchy matrices. which: 0=A 1=P 2=R. */
int psb200_debug_set_aggregates(psb200_handle h, int level, const int32_t *agg, int64_t n);
int psb200_debug_get_level(psb200_handle h, int level, int which, int64_t *rows, int64_t *cols,
                           int64_t *nnz, int32_t *row_ptr, int32_t *col_idx, double *vals);
/* aggregate id of every row of `level` (int32[n], -2 = removed node) and the aggregate count. */
int psb200_debug_get_aggregates(psb200_handle h, int level, int32_t *agg, int64_t n, int64_t *n_agg);
/* z = M^-1 r : one application of the configured preconditioner (host buffers). */
int psb200_precond_apply(psb200_handle h, const double *r, double *z, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* PSB200_H */
