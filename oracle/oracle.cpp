// TEST INFRASTRUCTURE ONLY (see oracle_core.hpp header). PARITY UNPINNED.
//
// CPU restatement of the reference's Krylov path:
//   polysolve::linear::EigenIterative<Eigen::ConjugateGradient<StiffnessMatrix, Lower|Upper,
//   DiagonalPreconditioner>>  (reference src/polysolve/linear/Solver.cpp:270-276,433-435;
//   wrapper protocol src/polysolve/linear/EigenSolver.tpp:66-114) and
//   EigenIterative<Eigen::BiCGSTAB<...>> (Solver.cpp:437-439).
// The loops follow Eigen 5.0.1's published conjugate_gradient()/bicgstab():
// same update order, same stopping rules, same reported iterations()/error().
#include "oracle_core.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <random>
#include <omp.h>

namespace orc {

void spmv_csr(const Csr &A, const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < A.n; ++i)
    {
        double s = 0;
        for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
            s += A.val[k] * x[A.col[k]];
        y[i] = s;
    }
}

} // namespace orc

namespace {

// y = A x for a compressed-column matrix: the column-scatter product Eigen runs for a
// column-major operand (single thread; Types.hpp:14 makes StiffnessMatrix column-major).
void spmv_csc(int64_t n, const int32_t *outer, const int32_t *inner, const double *val, const double *x, double *y)
{
    std::memset(y, 0, sizeof(double) * n);
    for (int64_t j = 0; j < n; ++j)
    {
        const double xj = x[j];
        for (int32_t k = outer[j]; k < outer[j + 1]; ++k)
            y[inner[k]] += val[k] * xj;
    }
}

void spmv_csr_raw(int64_t n, const int32_t *ptr, const int32_t *col, const double *val, const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
    {
        double s = 0;
        for (int32_t k = ptr[i]; k < ptr[i + 1]; ++k)
            s += val[k] * x[col[k]];
        y[i] = s;
    }
}

double dot(int64_t n, const double *a, const double *b, bool par)
{
    double s = 0;
    if (par)
    {
#pragma omp parallel for reduction(+ : s) schedule(static)
        for (int64_t i = 0; i < n; ++i)
            s += a[i] * b[i];
    }
    else
    {
        for (int64_t i = 0; i < n; ++i)
            s += a[i] * b[i];
    }
    return s;
}

struct Op
{
    int64_t n;
    const int32_t *ptr, *idx;
    const double *val;
    int mode; // 0: CSC column scatter, 1 thread (Eigen-faithful). 1: arrays are CSR, OpenMP rows.
    void apply(const double *x, double *y) const
    {
        if (mode == 0)
            spmv_csc(n, ptr, idx, val, x, y);
        else
            spmv_csr_raw(n, ptr, idx, val, x, y);
    }
    // Eigen::DiagonalPreconditioner::factorize: invdiag = (A_jj != 0) ? 1/A_jj : 1
    void inv_diag(double *d) const
    {
        for (int64_t j = 0; j < n; ++j)
        {
            d[j] = 1.0;
            for (int32_t k = ptr[j]; k < ptr[j + 1]; ++k)
                if (idx[k] == j)
                {
                    // Eigen takes the first stored diagonal hit via InnerIterator; duplicates do not
                    // occur in compressed matrices produced by setFromTriplets.
                    d[j] = (val[k] != 0.0) ? 1.0 / val[k] : 1.0;
                    break;
                }
        }
    }
};

} // namespace

extern "C" {

// ------------------------------------------------------------------ RNG / vectors
void orc_splitmix64_fill(uint64_t seed, int64_t n, double *out)
{
    uint64_t s = seed;
    for (int64_t i = 0; i < n; ++i)
        out[i] = orc::splitmix64_unit(s);
}

// ------------------------------------------------------------------ generators (CSC == CSR for these; column-major naming)
// 2-D 5-point Dirichlet Laplacian on an n x n grid, lexicographic. diag 4, off -1.
int64_t orc_poisson2d_nnz(int n) { return 5ll * n * n - 4ll * n; }
void orc_poisson2d(int n, int32_t *outer, int32_t *inner, double *val)
{
    int64_t k = 0;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i)
        {
            const int64_t r = (int64_t)j * n + i;
            outer[r] = (int32_t)k;
            if (j > 0) { inner[k] = (int32_t)(r - n); val[k++] = -1; }
            if (i > 0) { inner[k] = (int32_t)(r - 1); val[k++] = -1; }
            inner[k] = (int32_t)r; val[k++] = 4;
            if (i < n - 1) { inner[k] = (int32_t)(r + 1); val[k++] = -1; }
            if (j < n - 1) { inner[k] = (int32_t)(r + n); val[k++] = -1; }
        }
    outer[(int64_t)n * n] = (int32_t)k;
}

// 3-D 7-point Dirichlet Laplacian on n^3, lexicographic. diag 6, off -1.
int64_t orc_poisson3d_nnz(int n) { return 7ll * n * n * n - 6ll * n * n; }
void orc_poisson3d(int n, int32_t *outer, int32_t *inner, double *val)
{
    const int64_t n2 = (int64_t)n * n;
    int64_t k = 0;
    for (int z = 0; z < n; ++z)
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x)
            {
                const int64_t r = z * n2 + (int64_t)y * n + x;
                outer[r] = (int32_t)k;
                if (z > 0) { inner[k] = (int32_t)(r - n2); val[k++] = -1; }
                if (y > 0) { inner[k] = (int32_t)(r - n); val[k++] = -1; }
                if (x > 0) { inner[k] = (int32_t)(r - 1); val[k++] = -1; }
                inner[k] = (int32_t)r; val[k++] = 6;
                if (x < n - 1) { inner[k] = (int32_t)(r + 1); val[k++] = -1; }
                if (y < n - 1) { inner[k] = (int32_t)(r + n); val[k++] = -1; }
                if (z < n - 1) { inner[k] = (int32_t)(r + n2); val[k++] = -1; }
            }
    outer[n2 * n] = (int32_t)k;
}

// Non-symmetric 2-D convection-diffusion (upwinded), same pattern as poisson2d, stored CSC:
// row r couples to west with -(1+c), east with -(1-c)  (c in [0,1)), north/south -1, diag 4.
// Used for the BiCGSTAB cases, where CSC != CSR matters (SURVEY 7 hard part g).
void orc_convdiff2d(int n, double c, int32_t *outer, int32_t *inner, double *val)
{
    // Build column by column: column q holds A(r, q) for rows r adjacent to q.
    int64_t k = 0;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i)
        {
            const int64_t q = (int64_t)j * n + i;
            outer[q] = (int32_t)k;
            if (j > 0) { inner[k] = (int32_t)(q - n); val[k++] = -1; }          // row q-n, its "north" = q
            if (i > 0) { inner[k] = (int32_t)(q - 1); val[k++] = -(1.0 - c); }   // row q-1, its east neighbour is q
            inner[k] = (int32_t)q; val[k++] = 4;
            if (i < n - 1) { inner[k] = (int32_t)(q + 1); val[k++] = -(1.0 + c); } // row q+1, its west neighbour is q
            if (j < n - 1) { inner[k] = (int32_t)(q + n); val[k++] = -1; }
        }
    outer[(int64_t)n * n] = (int32_t)k;
}

// Values of the reference's pattern-reuse test (tests/test_linear_solver.cpp:262-283):
// std::default_random_engine{42}, uniform_real_distribution(0.1, 5); per round, traverse the CSC
// pattern column by column; diagonal <- urd*100; strict upper (row<col) <- -urd mirrored to (col,row).
// Writes `rounds` value arrays back to back (rounds * nnz doubles). Pattern must be symmetric.
void orc_prefactor_values(int64_t n, const int32_t *outer, const int32_t *inner, int rounds, double *vals_out)
{
    std::default_random_engine eng{42};
    std::uniform_real_distribution<double> urd(0.1, 5);
    const int64_t nnz = outer[n];
    for (int rd = 0; rd < rounds; ++rd)
    {
        double *v = vals_out + (int64_t)rd * nnz;
        for (int64_t c = 0; c < n; ++c)
            for (int32_t k = outer[c]; k < outer[c + 1]; ++k)
            {
                const int32_t r = inner[k];
                if (r == c)
                    v[k] = urd(eng) * 100;
                else if (r < c)
                {
                    const double val = -urd(eng);
                    v[k] = val;
                    // mirrored entry (c, r) lives in column r
                    const int32_t *b = inner + outer[r], *e = inner + outer[r + 1];
                    const int32_t *p = std::lower_bound(b, e, (int32_t)c);
                    v[p - inner] = val;
                }
            }
    }
}

// ------------------------------------------------------------------ analyze_pattern oracle
// Stable counting transpose CSC -> CSR. perm[k_csr] = k_csc (so val_csr[k] = val_csc[perm[k]]).
// This is the bit-exact index oracle for psb200_analyze_pattern_csc; the idiom (indices once in
// analyze_pattern, values refreshed by the same traversal in factorize) is the reference's own
// at src/polysolve/linear/Pardiso.cpp:164-197,216-221.
void orc_csc_to_csr(int64_t nrows, int64_t ncols, const int32_t *outer, const int32_t *inner,
                    int32_t *row_ptr, int32_t *col_idx, int32_t *perm)
{
    const int64_t nnz = outer[ncols];
    std::fill(row_ptr, row_ptr + nrows + 1, 0);
    for (int64_t k = 0; k < nnz; ++k)
        row_ptr[inner[k] + 1]++;
    for (int64_t i = 0; i < nrows; ++i)
        row_ptr[i + 1] += row_ptr[i];
    std::vector<int32_t> cur(row_ptr, row_ptr + nrows);
    for (int64_t c = 0; c < ncols; ++c)
        for (int32_t k = outer[c]; k < outer[c + 1]; ++k)
        {
            const int32_t dst = cur[inner[k]]++;
            col_idx[dst] = (int32_t)c;
            perm[dst] = k;
        }
}

// Row-range partition oracle: contiguous ranges balanced by nnz, aligned to `align` rows.
// offsets has world+1 entries. Rule: offsets[g] = smallest aligned row r with row_ptr[r] >= g*nnz/world.
void orc_partition_rows(int64_t n, const int32_t *row_ptr, int world, int align, int64_t *offsets)
{
    const int64_t nnz = row_ptr[n];
    offsets[0] = 0;
    for (int g = 1; g < world; ++g)
    {
        const int64_t target = (int64_t)(((__int128)nnz * g) / world);
        const int32_t *p = std::lower_bound(row_ptr, row_ptr + n + 1, (int32_t)std::min<int64_t>(target, INT32_MAX));
        int64_t r = p - row_ptr;
        r = ((r + align - 1) / align) * align;
        r = std::min(r, n);
        r = std::max(r, offsets[g - 1]);
        offsets[g] = r;
    }
    offsets[world] = n;
}

// Halo oracle for one rank: given global CSR and [r0,r1), list the distinct off-rank columns in
// ascending global order (halo_cols) and the local column remap: owned c -> c-r0, halo c -> (r1-r0)+pos.
// Returns number of halo columns. local_col must hold row_ptr[r1]-row_ptr[r0] entries.
int64_t orc_halo_for_rank(int64_t n, const int32_t *row_ptr, const int32_t *col_idx, int64_t r0, int64_t r1,
                          int32_t *local_col, int32_t *halo_cols, int64_t halo_cap)
{
    std::vector<int32_t> h;
    for (int64_t k = row_ptr[r0]; k < row_ptr[r1]; ++k)
    {
        const int32_t c = col_idx[k];
        if (c < r0 || c >= r1)
            h.push_back(c);
    }
    std::sort(h.begin(), h.end());
    h.erase(std::unique(h.begin(), h.end()), h.end());
    if ((int64_t)h.size() > halo_cap)
        return -(int64_t)h.size();
    std::copy(h.begin(), h.end(), halo_cols);
    const int64_t base = row_ptr[r0];
    for (int64_t k = row_ptr[r0]; k < row_ptr[r1]; ++k)
    {
        const int32_t c = col_idx[k];
        if (c >= r0 && c < r1)
            local_col[k - base] = (int32_t)(c - r0);
        else
            local_col[k - base] = (int32_t)((r1 - r0) + (std::lower_bound(h.begin(), h.end(), c) - h.begin()));
    }
    (void)n;
    return (int64_t)h.size();
}

// ------------------------------------------------------------------ SpMV
void orc_spmv_csc(int64_t n, const int32_t *outer, const int32_t *inner, const double *val, const double *x, double *y)
{
    spmv_csc(n, outer, inner, val, x, y);
}
void orc_spmv_csr(int64_t n, const int32_t *ptr, const int32_t *col, const double *val, const double *x, double *y)
{
    spmv_csr_raw(n, ptr, col, val, x, y);
}

int orc_num_threads() { return omp_get_max_threads(); }
void orc_set_num_threads(int t) { omp_set_num_threads(t); }

// ------------------------------------------------------------------ Eigen::ConjugateGradient + DiagonalPreconditioner
// SURVEY A.1. mode 0 = Eigen-faithful (CSC, column scatter, 1 thread); mode 1 = arrays are CSR and
// SpMV/dots use OpenMP (context baseline only). `stop_after` > 0 caps the number of loop trips
// without changing anything else (bounded-sample timing for bench.py); pass 0 for the real solve.
// Returns 0; *iters = Eigen iterations(); *err = Eigen error() (relative to ||b||); *spmvs = SpMV count in the loop.
int orc_eigen_cg(int64_t n, const int32_t *ptr, const int32_t *idx, const double *val, const double *b, double *x,
                 double tol, int64_t max_iters, int mode, int64_t stop_after, int64_t *iters, double *err, int64_t *spmvs)
{
    Op A{n, ptr, idx, val, mode};
    const bool par = mode != 0;
    std::vector<double> invd(n), r(n), p(n), z(n), tmp(n);
    A.inv_diag(invd.data());
    *spmvs = 0;

    A.apply(x, tmp.data());
    for (int64_t i = 0; i < n; ++i)
        r[i] = b[i] - tmp[i];
    const double bn2 = dot(n, b, b, par);
    if (bn2 == 0)
    {
        std::fill(x, x + n, 0.0);
        *iters = 0;
        *err = 0;
        return 0;
    }
    const double thr = std::max(tol * tol * bn2, DBL_MIN);
    double rn2 = dot(n, r.data(), r.data(), par);
    if (rn2 < thr)
    {
        *iters = 0;
        *err = std::sqrt(rn2 / bn2);
        return 0;
    }
    for (int64_t i = 0; i < n; ++i)
        p[i] = invd[i] * r[i];
    double abs_new = dot(n, r.data(), p.data(), par);
    int64_t i = 0;
    while (i < max_iters)
    {
        A.apply(p.data(), tmp.data());
        ++*spmvs;
        const double alpha = abs_new / dot(n, p.data(), tmp.data(), par);
#pragma omp parallel for schedule(static) if (par)
        for (int64_t k = 0; k < n; ++k)
            x[k] += alpha * p[k];
#pragma omp parallel for schedule(static) if (par)
        for (int64_t k = 0; k < n; ++k)
            r[k] -= alpha * tmp[k];
        rn2 = dot(n, r.data(), r.data(), par);
        if (rn2 < thr)
            break;
#pragma omp parallel for schedule(static) if (par)
        for (int64_t k = 0; k < n; ++k)
            z[k] = invd[k] * r[k];
        const double abs_old = abs_new;
        abs_new = dot(n, r.data(), z.data(), par);
        const double beta = abs_new / abs_old;
#pragma omp parallel for schedule(static) if (par)
        for (int64_t k = 0; k < n; ++k)
            p[k] = z[k] + beta * p[k];
        i++;
        if (stop_after > 0 && i >= stop_after)
            break;
    }
    *err = std::sqrt(rn2 / bn2);
    *iters = i;
    return 0;
}

// ------------------------------------------------------------------ Eigen::BiCGSTAB + DiagonalPreconditioner
int orc_eigen_bicgstab(int64_t n, const int32_t *ptr, const int32_t *idx, const double *val, const double *b, double *x,
                       double tol, int64_t max_iters, int mode, int64_t *iters, double *err, int64_t *spmvs)
{
    Op A{n, ptr, idx, val, mode};
    const bool par = mode != 0;
    std::vector<double> invd(n), r(n), r0(n), v(n, 0.0), p(n, 0.0), y(n), z(n), s(n), t(n);
    A.inv_diag(invd.data());
    *spmvs = 0;

    A.apply(x, t.data());
    for (int64_t k = 0; k < n; ++k)
        r[k] = b[k] - t[k];
    r0 = r;
    double r0_sqnorm = dot(n, r0.data(), r0.data(), par);
    const double rhs_sqnorm = dot(n, b, b, par);
    if (rhs_sqnorm == 0)
    {
        std::fill(x, x + n, 0.0);
        *iters = 0;
        *err = 0;
        return 0;
    }
    double rho = 1, alpha = 1, w = 1;
    const double tol2 = tol * tol * rhs_sqnorm;
    const double eps2 = DBL_EPSILON * DBL_EPSILON;
    int64_t i = 0, restarts = 0;
    double rn2 = dot(n, r.data(), r.data(), par);
    while (rn2 > tol2 && i < max_iters)
    {
        const double rho_old = rho;
        rho = dot(n, r0.data(), r.data(), par);
        if (std::fabs(rho) < eps2 * r0_sqnorm)
        {
            A.apply(x, t.data());
            ++*spmvs;
            for (int64_t k = 0; k < n; ++k)
                r[k] = b[k] - t[k];
            r0 = r;
            rho = r0_sqnorm = dot(n, r.data(), r.data(), par);
            if (restarts++ == 0)
                i = 0;
        }
        const double beta = (rho / rho_old) * (alpha / w);
        for (int64_t k = 0; k < n; ++k)
            p[k] = r[k] + beta * (p[k] - w * v[k]);
        for (int64_t k = 0; k < n; ++k)
            y[k] = invd[k] * p[k];
        A.apply(y.data(), v.data());
        ++*spmvs;
        alpha = rho / dot(n, r0.data(), v.data(), par);
        for (int64_t k = 0; k < n; ++k)
            s[k] = r[k] - alpha * v[k];
        for (int64_t k = 0; k < n; ++k)
            z[k] = invd[k] * s[k];
        A.apply(z.data(), t.data());
        ++*spmvs;
        const double tt = dot(n, t.data(), t.data(), par);
        w = tt > 0 ? dot(n, t.data(), s.data(), par) / tt : 0.0;
        for (int64_t k = 0; k < n; ++k)
            x[k] += alpha * y[k] + w * z[k];
        for (int64_t k = 0; k < n; ++k)
            r[k] = s[k] - w * t[k];
        rn2 = dot(n, r.data(), r.data(), par);
        ++i;
    }
    *err = std::sqrt(rn2 / rhs_sqnorm);
    *iters = i;
    return 0;
}

} // extern "C"
