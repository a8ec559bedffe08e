/* psb200_problems -- a device-resident polysolve::nonlinear::Problem for BASELINE config 5.
 *
 * The reference's nonlinear solver minimises a user-supplied Problem subclass
 * (src/polysolve/nonlinear/Problem.hpp:22-143: value, gradient, hessian, ...); its tests define theirs inline
 * (tests/test_nonlinear_solver.cpp:30-325). Config 5 names "a nonlinear elasticity Problem subclass" whose Newton steps
 * end in the GPU linear solver: this is that Problem -- compressible Neo-Hookean P1 tetrahedra -- with its energy,
 * gradient and Hessian evaluated by CUDA kernels, the Hessian assembled straight into a fixed compressed-column pattern
 * in device memory, ready for psb200_factorize_csc_device (the device-resident form of
 * Newton::solve_sparse_linear_system, src/polysolve/nonlinear/descent_strategies/Newton.cpp:173-214).
 *
 *   W(F) = mu/2 (|F|^2 - 3) - mu ln J + lambda/2 (ln J)^2,  F = I + grad u,  unknowns = nodal displacements (3 per node)
 *
 * All evaluations are deterministic (gathers in a fixed order, no floating-point atomics): ranks of a multi-GPU run that
 * evaluate the same x take the same line-search decisions.
 */
#ifndef PSB200_PROBLEMS_H
#define PSB200_PROBLEMS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct psb200_nh *psb200_nh_handle;

/* X: node coordinates (3 per node), tets: 4 node ids per tetrahedron (positively oriented), fixed: one byte per dof
 * (non-zero = Dirichlet: the dof keeps its value; zero gradient, identity row and column in the Hessian).
 * device < 0: the current device. */
int psb200_nh_create(psb200_nh_handle *out, int64_t n_nodes, const double *X, int64_t n_tets, const int32_t *tets, double mu,
                     double lambda, const uint8_t *fixed, int device);
int psb200_nh_destroy(psb200_nh_handle h);
/* The Hessian's pattern: Eigen-style compressed-column arrays (host memory owned by the handle), full 3 x 3 blocks for
 * every pair of nodes that share a tetrahedron, rows ascending. Constant for the lifetime of the handle. */
int psb200_nh_pattern(psb200_nh_handle h, int64_t *n, int64_t *nnz, const int32_t **outer, const int32_t **inner);
/* Problem::value -- Problem.hpp:49. +inf when an element is inverted (J <= 0). x: host vector of 3 n_nodes displacements. */
int psb200_nh_value(psb200_nh_handle h, const double *x, double *value_out);
/* Problem::gradient -- Problem.hpp:54 */
int psb200_nh_gradient(psb200_nh_handle h, const double *x, double *grad_out);
/* Problem::hessian -- Problem.hpp:67, values only, left in DEVICE memory in the order of psb200_nh_pattern (valid until
 * the next call). Feed to psb200_factorize_csc_device. */
int psb200_nh_hessian_device(psb200_nh_handle h, const double *x, const double **d_vals_out);
/* the same values copied to the host (tests) */
int psb200_nh_hessian_host(psb200_nh_handle h, const double *x, double *vals_out);
const char *psb200_nh_last_error(psb200_nh_handle h);

#ifdef __cplusplus
}
#endif
#endif /* PSB200_PROBLEMS_H */
