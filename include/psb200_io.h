/* psb200_io -- Matrix Market replay of the reference's fixtures.
 *
 * The reference's tests read their matrices with Eigen::loadMarket (unsupported/Eigen/SparseExtra; e.g.
 * tests/test_linear_solver.cpp:55 "A_2.mat") and with the test-local loadSymmetric
 * (tests/test_linear_solver.cpp:25-50: a coordinate file that stores one triangle, mirrored on load), and write
 * with Eigen::saveMarket (FEMSolver.cpp:288, Newton.cpp:196). These entry points read / write the same files into /
 * from the compressed-column arrays psb200_analyze_pattern_csc takes (StiffnessMatrix layout, Types.hpp:11-15):
 * column pointers int32[cols+1], row indices int32[nnz] ascending inside a column, values f64[nnz]; duplicate
 * entries are summed like SparseMatrix::setFromTriplets does. Host-only code. */
#ifndef PSB200_IO_H
#define PSB200_IO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct psb200_market *psb200_market_handle;

/* symmetric_mode:  0 = entries as stored, whatever the header says (Eigen::loadMarket of Eigen 3.4: the header's
 *                      symmetry field is not interpreted);
 *                  1 = every off-diagonal entry (i, j) also yields (j, i) (loadSymmetric, test_linear_solver.cpp:25-50);
 *                 -1 = follow the header: mirror iff it says "symmetric" (skew-symmetric / hermitian are rejected).
 * Supported headers: "matrix coordinate real|integer|pattern general|symmetric" (pattern entries get the value 1).
 * Lines starting with '%' and blank lines are skipped anywhere. Out-of-range indices are an error. */
int psb200_market_load(const char *path, int symmetric_mode, psb200_market_handle *out, int64_t *rows, int64_t *cols,
                       int64_t *nnz);
/* copies the matrix out: outer int32[cols+1], inner int32[nnz], vals f64[nnz] */
int psb200_market_get_csc(psb200_market_handle m, int32_t *outer, int32_t *inner, double *vals);
int psb200_market_free(psb200_market_handle m);
/* Eigen::saveMarket(A, path): "%%MatrixMarket matrix coordinate real general", then "rows cols nnz", then one line
 * "i j value" (1-based) per stored entry in column-major order. symmetric != 0 writes the "symmetric" header and the
 * lower triangle only. Values are written with 17 significant digits (Eigen writes the stream default of 6: a replay
 * through this writer is lossless, a file written by the reference is read as it is). */
int psb200_market_save(const char *path, int64_t rows, int64_t cols, const int32_t *outer, const int32_t *inner,
                       const double *vals, int symmetric);
/* Eigen::loadMarketVector / saveMarketVector: "%%MatrixMarket matrix array real general", "n 1", n values. */
int psb200_market_load_vector(const char *path, double *out, int64_t cap, int64_t *n);
int psb200_market_save_vector(const char *path, const double *v, int64_t n);
/* message of the last failure on the calling thread */
const char *psb200_market_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* PSB200_IO_H */
