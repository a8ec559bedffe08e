/* psb200_nl -- C ABI of the Newton / line-search driver that calls the psb200 linear solver the way
 * polysolve::nonlinear::Solver does.
 *
 * It mirrors, call for call, the reference's nonlinear path that ends in the hot path:
 *   nonlinear::Solver::create(solver_params, linear_solver_params, ...)   src/polysolve/nonlinear/Solver.cpp:124-186
 *   nonlinear::Solver::minimize(Problem&, TVector&)                       src/polysolve/nonlinear/Solver.cpp:255-582
 *   Newton / ProjectedNewton / RegularizedNewton (+ GradientDescent)      src/polysolve/nonlinear/descent_strategies/Newton.cpp:14-58,144-214,275-330
 *   LineSearch::line_search + Backtracking / Armijo / RobustArmijo /      src/polysolve/nonlinear/line_search/LineSearch.cpp:73-254, Backtracking.cpp:15-83, Armijo.cpp:13-32,
 *   ResidualBacktracking / NoLineSearch                                   RobustArmijo.cpp:16-46, ResidualBacktracking.cpp:15-28, NoLineSearch.cpp:11-22
 *   checkConvergence                                                      src/polysolve/nonlinear/Criteria.cpp:59-96
 * A polysolve::nonlinear::Problem subclass (src/polysolve/nonlinear/Problem.hpp:22-143) becomes a
 * table of callbacks. In a polysolve build nothing of this is needed -- nonlinear::Solver keeps calling
 * linear::Solver::create("CUDA") -- it exists so that the Newton call site (config 5 of BASELINE.json)
 * can be exercised and parity-tested in this repository, where polysolve itself cannot be built.
 */
#ifndef PSB200_NL_H
#define PSB200_NL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct psb200_nl_solver *psb200_nl_handle;

/* The solver's current Criteria (Criteria.hpp:34-57) as handed to Problem::callback. */
typedef struct psb200_nl_criteria
{
    int64_t iterations;
    double xDelta, fDelta, gradNorm, firstGradNorm, xDeltaDotGrad, relGradNorm, relXDelta, newtonDecrement;
    int64_t fDeltaCount;
    double energy, alpha, step;
} psb200_nl_criteria;

/* polysolve::nonlinear::Problem as callbacks. value, gradient and hessian are required, the rest may be NULL
 * (the defaults of Problem.hpp apply: steps are valid, max step size 1, no-ops). Vectors have length n. */
typedef struct psb200_nl_problem
{
    void *user;
    /* Problem::value(x) -- Problem.hpp:49 */
    double (*value)(void *user, const double *x, int64_t n);
    /* Problem::gradient(x, grad) -- Problem.hpp:54 */
    void (*gradient)(void *user, const double *x, int64_t n, double *grad_out);
    /* Problem::hessian(x, StiffnessMatrix&) -- Problem.hpp:67. The callee returns Eigen-style compressed-column
     * arrays that stay valid until its next call. project_to_psd mirrors Problem::set_project_to_psd
     * (Problem.hpp:97). Non-zero return = assembly failure (treated like a failed factorization). */
    int (*hessian)(void *user, const double *x, int64_t n, int project_to_psd, int64_t *nnz, const int32_t **outer,
                   const int32_t **inner, const double **vals);
    /* Problem::solution_changed -- Problem.hpp:101 */
    void (*solution_changed)(void *user, const double *x, int64_t n);
    /* Problem::is_step_valid -- Problem.hpp:73 */
    int (*is_step_valid)(void *user, const double *x0, const double *x1, int64_t n);
    /* Problem::max_step_size -- Problem.hpp:79 */
    double (*max_step_size)(void *user, const double *x0, const double *x1, int64_t n);
    /* Problem::line_search_begin / line_search_end -- Problem.hpp:86-89 */
    void (*line_search_begin)(void *user, const double *x0, const double *x1, int64_t n);
    void (*line_search_end)(void *user);
    /* Problem::post_step(PostStepData) -- Problem.hpp:93 */
    void (*post_step)(void *user, int iteration, const double *x, const double *grad, int64_t n);
    /* Problem::stop(x) -- Problem.hpp:114; non-zero stops with status "ObjectiveCustomStop" */
    int (*stop)(void *user, const double *x, int64_t n);
    /* Optional (may be NULL): Problem::hessian with the VALUES left in device memory -- the Hessian is assembled by the
     * caller's own CUDA kernel straight into a fixed compressed-column pattern (outer / inner: host arrays, the same on
     * every call) and goes to the linear solver through psb200_factorize_csc_device: nothing of the matrix crosses PCIe
     * in a Newton step (Newton.cpp:173-214 device-resident; RegularizedNewton's shift is applied on the device and needs
     * a structurally present diagonal). When set, it is used instead of `hessian`. */
    int (*hessian_device)(void *user, const double *x, int64_t n, int project_to_psd, int64_t *nnz, const int32_t **outer,
                          const int32_t **inner, const double **d_vals);
    /* ---- the remaining optional Problem virtuals (may be NULL: zero-initialise the struct). norm_type: 0 Euclidean,
     * 1 L2, 2 Linf -- "norm_type" of the solver parameters (Problem.hpp:14-19, Solver.cpp:117-121,224). */
    /* Problem::is_residual -- Problem.hpp:35: non-zero skips the descent-direction test (Solver.cpp:425) */
    int (*is_residual)(void *user);
    /* Problem::after_line_search_custom_operation -- Problem.hpp:103: non-zero => solution_changed(x1) (Solver.cpp:495-499) */
    int (*after_line_search_custom_operation)(void *user, const double *x0, const double *x1, int64_t n);
    /* Problem::callback(state, x) -- Problem.hpp:109: called at the end of every trip, zero ends the loop (Solver.cpp:558) */
    int (*callback)(void *user, const psb200_nl_criteria *state, const double *x, int64_t n);
    /* Problem::grad_norm / step_norm -- Problem.hpp:120-121 (defaults: Euclidean norm). Used for the stopping criteria,
     * by the line search (LineSearch.cpp:138-142, Backtracking.cpp:76-80). The Newton residual check ||H dx + g||
     * (Newton.cpp:207) is evaluated on the device and stays Euclidean. */
    double (*grad_norm)(void *user, const double *grad, int64_t n, int norm_type);
    double (*step_norm)(void *user, const double *dx, int64_t n, int norm_type);
    /* Problem::grad_norm_rescaling (which = 0) / step_norm_rescaling (1) / energy_norm_rescaling (2) -- Problem.hpp:116-118:
     * factors applied to the absolute tolerances (Solver.hpp:118-131) */
    double (*norm_rescaling)(void *user, int which, int norm_type);
} psb200_nl_problem;

/* nonlinear::Solver::create. solver_params: the reference's nonlinear JSON (keys and defaults of
 * nonlinear-solver-spec.json: "solver": "Newton" | "L-BFGS" | "GradientDescent" (each followed by the GradientDescent fallback,
 * Solver.cpp:156-181) or a LIST of strategies [{"type": "Newton"|"ProjectedNewton"|"RegularizedNewton"|
 * "RegularizedProjectedNewton"|"L-BFGS"|"GradientDescent", ...per-strategy parameters}] used in that order (Solver.cpp:147-154),
 * "line_search": {"method": "RobustArmijo"|"Armijo"|"Backtracking"|"ResidualBacktracking"|"None", ...},
 * "grad_norm_tol", "max_iterations", "norm_type", "newton_decrement_tol", "iterations_per_strategy" (one value or one per
 * strategy + 1, Solver.cpp:232-245), "Newton": {"residual_tolerance", "reg_weight_min", ...}, ...).
 * linear_params: the linear-solver JSON handed to every strategy's linear::Solver ({"solver": "CUDA", "CUDA": {...}}).
 * Either may be NULL for the defaults. */
int psb200_nl_create(psb200_nl_handle *out, const char *solver_params_json, const char *linear_params_json);
int psb200_nl_destroy(psb200_nl_handle h);
/* Calls hook(user, lin) once for every linear solver the driver owns (one per Newton strategy, Newton.cpp:32-52), in
 * creation order, before they are first used: a multi-GPU application connects them there (psb200_dist_prepare, exchange
 * of the IPC handles, psb200_dist_connect -- collectively, every rank in the same order). With connected solvers the
 * driver all-gathers every Newton step (psb200_dist_allgather), so each rank keeps the whole iterate and evaluates the
 * Problem redundantly; the Problem must then be deterministic across ranks. `lin` is the linear solver handle of psb200.h (a psb200_handle). */
int psb200_nl_set_linear_solver_hook(psb200_nl_handle h, void (*hook)(void *user, void *lin), void *user);
/* Solver::set_iteration_callback (Solver.hpp: "Iteration callback"; Solver.cpp:548-552): called once per completed
 * iteration with the current Criteria (alpha = the accepted line-search step); non-zero ends the solve with status
 * "ObjectiveCustomStop" (not an error). NULL removes it. The reference's test "iteration-callback"
 * (tests/test_nonlinear_solver.cpp:714-754) is replayed in tests/test_newton.py. */
int psb200_nl_set_iteration_callback(psb200_nl_handle h, int (*callback)(void *user, const psb200_nl_criteria *state), void *user);
/* Solver::set_direction_filter (Solver.hpp:80-86; Solver.cpp:353-358,392-403): filter(user, x, dx_inout, n) is applied to
 * every successfully computed update direction before it is vetted and searched, and to -grad when descent is measured
 * (dx . grad becomes -dx . filter(-grad)). NULL removes it. */
int psb200_nl_set_direction_filter(psb200_nl_handle h, void (*filter)(void *user, const double *x, double *dx_inout, int64_t n), void *user);
/* nonlinear::Solver::minimize(problem, x): x is in/out. Returns 0 when the loop ended with a converged status;
 * non-zero (message in psb200_nl_last_error) where the reference throws (NaN, iteration limit without
 * allow_out_of_iterations, failure on the last strategy). x always holds the last iterate. */
int psb200_nl_minimize(psb200_nl_handle h, const psb200_nl_problem *problem, double *x_inout, int64_t n);
/* solver_info of the reference (Solver.cpp:615-637, Newton.cpp:333-340): "status", "iterations", "energy",
 * "grad_norm", "line_search", "internal_solver": [get_info of every linear solve], timings. */
int psb200_nl_get_info(psb200_nl_handle h, char *json_out, size_t cap, size_t *needed);
const char *psb200_nl_last_error(psb200_nl_handle h);

/* ---- L-BFGS memory on device vectors. Mirrors what the reference's LBFGS strategy does with
 * LBFGSpp::BFGSMat (src/polysolve/nonlinear/descent_strategies/LBFGS.cpp:22-61): the first call after create / reset
 * returns -grad; every later call adds the correction (x - x_prev, grad - grad_prev) to a memory of history_size pairs
 * ("/L-BFGS/history_size", default 6, nonlinear-solver-spec.json) and returns direction = -H grad by the two-loop
 * recursion. psb200_nl_create with {"solver": "L-BFGS"} uses it as the first strategy, followed by GradientDescent
 * (Solver.cpp:83-85,175-181). device < 0: the current device. */
typedef struct psb200_lbfgs *psb200_lbfgs_handle;
int psb200_lbfgs_create(psb200_lbfgs_handle *out, int64_t n, int history_size, int device);
int psb200_lbfgs_destroy(psb200_lbfgs_handle h);
/* LBFGS::reset -- LBFGS.cpp:22-28 */
int psb200_lbfgs_reset(psb200_lbfgs_handle h);
/* LBFGS::compute_update_direction -- LBFGS.cpp:30-61; host vectors of length n */
int psb200_lbfgs_direction(psb200_lbfgs_handle h, const double *x, const double *grad, double *direction, int64_t n);
/* same with x, grad, direction resident on the GPU */
int psb200_lbfgs_direction_device(psb200_lbfgs_handle h, const double *d_x, const double *d_grad, double *d_direction,
                                  int64_t n);
const char *psb200_lbfgs_last_error(psb200_lbfgs_handle h);

#ifdef __cplusplus
}
#endif
#endif /* PSB200_NL_H */
