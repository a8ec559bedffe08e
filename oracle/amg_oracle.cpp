// TEST INFRASTRUCTURE ONLY (see oracle_core.hpp header). PARITY UNPINNED.
//
// CPU restatement of the reference's AMG-PCG path: polysolve::linear::AMGCL
// (reference src/polysolve/linear/AMGCL.cpp:32-65 default_params, :148-184 factorize,
// :190-212 solve) = amgcl::make_solver<amg<builtin<double>, smoothed_aggregation, chebyshev>, cg>
// from AMGCL 1.4.3 (cmake/recipes/amgcl.cmake:44-48; NOT in /root/reference, NOT in this image).
// The algorithms below restate AMGCL's published amg.hpp / plain_aggregates.hpp /
// smoothed_aggregation.hpp / relaxation/chebyshev.hpp / solver/cg.hpp as summarised in
// SURVEY.md Appendix A.3. Deviation (documented): the power-iteration start vector is a
// splitmix64 stream instead of per-thread mt19937(thread_id), so results do not depend on
// the OpenMP thread count.
#include "oracle_core.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <numeric>
#include <omp.h>
#include <stdexcept>

namespace {
using orc::Csr;

struct Params
{
    int max_levels = 6;         // AMGCL.cpp:44
    int coarse_enough = 3000;   // amgcl skyline_lu::coarse_enough() for scalar value type
    int direct_coarse = 0;      // AMGCL.cpp:45
    int ncycle = 2;             // AMGCL.cpp:46
    int npre = 1, npost = 1, pre_cycles = 1; // amgcl defaults
    int degree = 16;            // AMGCL.cpp:36
    int power_iters = 100;      // AMGCL.cpp:38
    double higher = 2.0;        // AMGCL.cpp:39
    double lower = 0.008333333333; // AMGCL.cpp:40
    int scale = 1;              // AMGCL.cpp:41
    double sa_relax = 1.0;      // AMGCL.cpp:50
    int estimate_spectral_radius = 1; // AMGCL.cpp:49
    double eps_strong = 0.0;    // AMGCL.cpp:52
    int relax_type = 0;         // 0 chebyshev (default), 1 damped jacobi (0.72), 2 spai0
};

struct Level
{
    Csr A, P, R;
    std::vector<double> M;       // inverted diagonal (chebyshev scale=true) or relaxation weights
    double cheb_d = 0, cheb_c = 0, rho = 0;
    std::vector<double> f, u, t; // work vectors
    std::vector<double> cr, cp;  // chebyshev r, p
    std::vector<int32_t> aggr;   // aggregate id per row (for level -> level+1), -2 removed
    std::vector<double> lu;      // dense LU for direct coarse
    std::vector<int> piv;
    bool direct = false;
};

struct Hierarchy
{
    Params prm;
    std::vector<Level> levels;
};

// Gershgorin bound or power iteration on (D^-1)A -- amgcl backend::spectral_radius<scale>.
double spectral_radius(const Csr &A, bool scale, int power_iters, uint64_t seed)
{
    const int64_t n = A.n;
    if (power_iters <= 0)
    {
        double emax = 0;
#pragma omp parallel for reduction(max : emax) schedule(static)
        for (int64_t i = 0; i < n; ++i)
        {
            double s = 0, dia = 1;
            for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
            {
                s += std::fabs(A.val[k]);
                if (A.col[k] == i)
                    dia = A.val[k];
            }
            if (scale)
                s *= std::fabs(1.0 / dia);
            emax = std::max(emax, s);
        }
        return emax;
    }
    std::vector<double> b0(n), b1(n);
    uint64_t s = seed;
    double nrm = 0;
    for (int64_t i = 0; i < n; ++i)
    {
        b0[i] = orc::splitmix64_unit(s);
        nrm += b0[i] * b0[i];
    }
    nrm = 1 / std::sqrt(nrm);
    for (int64_t i = 0; i < n; ++i)
        b0[i] *= nrm;
    double radius = 1;
    for (int it = 0; it < power_iters;)
    {
        double b1_norm = 0;
        radius = 0;
#pragma omp parallel for reduction(+ : b1_norm, radius) schedule(static)
        for (int64_t i = 0; i < n; ++i)
        {
            double sum = 0, dia = 1;
            for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
            {
                sum += A.val[k] * b0[A.col[k]];
                if (A.col[k] == i)
                    dia = A.val[k];
            }
            if (scale)
                sum *= 1.0 / dia;
            b1_norm += sum * sum;
            radius += std::fabs(sum * b0[i]);
            b1[i] = sum;
        }
        if (++it < power_iters)
        {
            const double inv = 1 / std::sqrt(b1_norm);
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i)
                b0[i] = b1[i] * inv;
        }
    }
    return radius < 0 ? 2.0 : radius;
}

// amgcl::coarsening::plain_aggregates (sequential greedy).
int64_t plain_aggregates(const Csr &A, double eps_strong, std::vector<char> &strong, std::vector<int32_t> &id)
{
    const int64_t n = A.n;
    const double eps2 = eps_strong * eps_strong;
    strong.assign(A.nnz(), 0);
    id.assign(n, -1);
    std::vector<double> dia(n, 0.0);
    for (int64_t i = 0; i < n; ++i)
        for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
            if (A.col[k] == i)
                dia[i] = A.val[k];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
    {
        const double eps_dia_i = eps2 * dia[i];
        for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
        {
            const int32_t c = A.col[k];
            const double v = A.val[k];
            strong[k] = (c != i) && (eps_dia_i * dia[c] < v * v);
        }
    }
    const int32_t undefined = -1, removed = -2;
    for (int64_t i = 0; i < n; ++i)
    {
        int32_t state = removed;
        for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
            if (strong[k])
            {
                state = undefined;
                break;
            }
        id[i] = state;
    }
    int64_t count = 0;
    std::vector<int32_t> neib;
    for (int64_t i = 0; i < n; ++i)
    {
        if (id[i] != undefined)
            continue;
        const int32_t cur = (int32_t)count++;
        id[i] = cur;
        neib.clear();
        for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
        {
            const int32_t c = A.col[k];
            if (strong[k] && id[c] != removed)
            {
                id[c] = cur;
                neib.push_back(c);
            }
        }
        for (int32_t c : neib)
            for (int32_t k = A.ptr[c]; k < A.ptr[c + 1]; ++k)
            {
                const int32_t cc = A.col[k];
                if (strong[k] && id[cc] == undefined)
                    id[cc] = cur;
            }
    }
    if (!count)
        return 0;
    std::vector<int32_t> cnt(count, 0);
    for (int64_t i = 0; i < n; ++i)
        if (id[i] >= 0)
            cnt[id[i]] = 1;
    std::partial_sum(cnt.begin(), cnt.end(), cnt.begin());
    if (count > cnt.back())
    {
        count = cnt.back();
        for (int64_t i = 0; i < n; ++i)
            if (id[i] >= 0)
                id[i] = cnt[id[i]] - 1;
    }
    return count;
}

void sort_rows(Csr &M)
{
#pragma omp parallel
    {
        std::vector<std::pair<int32_t, double>> tmp;
#pragma omp for schedule(dynamic, 1024)
        for (int64_t i = 0; i < M.n; ++i)
        {
            const int32_t b = M.ptr[i], e = M.ptr[i + 1];
            tmp.resize(e - b);
            for (int32_t k = b; k < e; ++k)
                tmp[k - b] = {M.col[k], M.val[k]};
            std::sort(tmp.begin(), tmp.end(), [](auto &a, auto &c) { return a.first < c.first; });
            for (int32_t k = b; k < e; ++k)
            {
                M.col[k] = tmp[k - b].first;
                M.val[k] = tmp[k - b].second;
            }
        }
    }
}

// Smoothed prolongation P = (I - omega D_f^-1 A_f) P_tent with P_tent(i, id[i]) = 1
// (amgcl smoothed_aggregation::transfer_operators without near-nullspace vectors).
Csr smoothed_prolongation(const Csr &A, const std::vector<char> &strong, const std::vector<int32_t> &id, int64_t nc, double omega)
{
    const int64_t n = A.n;
    Csr P;
    P.n = n;
    P.ptr.assign(n + 1, 0);
#pragma omp parallel
    {
        std::vector<int32_t> marker(nc, -1);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i)
        {
            int32_t cnt = 0;
            for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
            {
                const int32_t c = A.col[k];
                if (c != i && !strong[k])
                    continue;
                const int32_t g = id[c];
                if (g < 0)
                    continue;
                if (marker[g] != (int32_t)i)
                {
                    marker[g] = (int32_t)i;
                    ++cnt;
                }
            }
            P.ptr[i + 1] = cnt;
        }
    }
    for (int64_t i = 0; i < n; ++i)
        P.ptr[i + 1] += P.ptr[i];
    P.col.resize(P.ptr[n]);
    P.val.resize(P.ptr[n]);
#pragma omp parallel
    {
        std::vector<int32_t> marker(nc, -1);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i)
        {
            double dia = 0;
            for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
                if (A.col[k] == i || !strong[k])
                    dia += A.val[k];
            dia = -omega * (1.0 / dia);
            const int32_t row_beg = P.ptr[i];
            int32_t row_end = row_beg;
            for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
            {
                const int32_t c = A.col[k];
                if (c != i && !strong[k])
                    continue;
                const int32_t g = id[c];
                if (g < 0)
                    continue;
                const double va = (c == i) ? (1.0 - omega) : dia * A.val[k];
                if (marker[g] < row_beg)
                {
                    marker[g] = row_end;
                    P.col[row_end] = g;
                    P.val[row_end] = va;
                    ++row_end;
                }
                else
                    P.val[marker[g]] += va;
            }
        }
    }
    sort_rows(P);
    return P;
}

Csr transpose(const Csr &A, int64_t ncols)
{
    Csr T;
    T.n = ncols;
    T.ptr.assign(ncols + 1, 0);
    for (int64_t k = 0; k < A.nnz(); ++k)
        T.ptr[A.col[k] + 1]++;
    for (int64_t i = 0; i < ncols; ++i)
        T.ptr[i + 1] += T.ptr[i];
    T.col.resize(A.nnz());
    T.val.resize(A.nnz());
    std::vector<int32_t> cur(T.ptr.begin(), T.ptr.end() - 1);
    for (int64_t i = 0; i < A.n; ++i)
        for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
        {
            const int32_t d = cur[A.col[k]]++;
            T.col[d] = (int32_t)i;
            T.val[d] = A.val[k];
        }
    return T;
}

// Row-wise (Gustavson) C = A * B with B having ncolsB columns; rows sorted on output.
Csr spgemm(const Csr &A, const Csr &B, int64_t ncolsB)
{
    Csr C;
    C.n = A.n;
    C.ptr.assign(A.n + 1, 0);
#pragma omp parallel
    {
        std::vector<int32_t> marker(ncolsB, -1);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < A.n; ++i)
        {
            int32_t cnt = 0;
            for (int32_t ka = A.ptr[i]; ka < A.ptr[i + 1]; ++ka)
            {
                const int32_t a = A.col[ka];
                for (int32_t kb = B.ptr[a]; kb < B.ptr[a + 1]; ++kb)
                    if (marker[B.col[kb]] != (int32_t)i)
                    {
                        marker[B.col[kb]] = (int32_t)i;
                        ++cnt;
                    }
            }
            C.ptr[i + 1] = cnt;
        }
    }
    for (int64_t i = 0; i < A.n; ++i)
        C.ptr[i + 1] += C.ptr[i];
    C.col.resize(C.ptr[A.n]);
    C.val.resize(C.ptr[A.n]);
#pragma omp parallel
    {
        std::vector<int32_t> marker(ncolsB, -1);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < A.n; ++i)
        {
            const int32_t row_beg = C.ptr[i];
            int32_t row_end = row_beg;
            for (int32_t ka = A.ptr[i]; ka < A.ptr[i + 1]; ++ka)
            {
                const int32_t a = A.col[ka];
                const double va = A.val[ka];
                for (int32_t kb = B.ptr[a]; kb < B.ptr[a + 1]; ++kb)
                {
                    const int32_t c = B.col[kb];
                    if (marker[c] < row_beg)
                    {
                        marker[c] = row_end;
                        C.col[row_end] = c;
                        C.val[row_end] = va * B.val[kb];
                        ++row_end;
                    }
                    else
                        C.val[marker[c]] += va * B.val[kb];
                }
            }
        }
    }
    sort_rows(C);
    return C;
}

void setup_relax(Level &L, const Params &prm, uint64_t seed)
{
    const Csr &A = L.A;
    const int64_t n = A.n;
    L.M.assign(n, 1.0);
    for (int64_t i = 0; i < n; ++i)
        for (int32_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
            if (A.col[k] == i)
                L.M[i] = 1.0 / A.val[k];
    if (prm.relax_type == 0)
    {
        double hi = spectral_radius(A, prm.scale != 0, prm.power_iters, seed);
        L.rho = hi;
        const double lo = hi * prm.lower;
        hi *= prm.higher;
        L.cheb_d = 0.5 * (hi + lo);
        L.cheb_c = 0.5 * (hi - lo);
        if (!prm.scale)
            std::fill(L.M.begin(), L.M.end(), 1.0);
        L.cr.assign(n, 0.0);
        L.cp.assign(n, 0.0);
    }
    else if (prm.relax_type == 1)
    {
        for (auto &m : L.M)
            m *= 0.72; // amgcl damped_jacobi default damping
    }
    L.f.assign(n, 0.0);
    L.u.assign(n, 0.0);
    L.t.assign(n, 0.0);
}

// amgcl relaxation::chebyshev::solve  
void chebyshev_apply(Level &L, const Params &prm, const double *rhs, double *x)
{
    const Csr &A = L.A;
    const int64_t n = A.n;
    const double d = L.cheb_d, c = L.cheb_c;
    double alpha = 0, beta = 0;
    double *r = L.cr.data(), *p = L.cp.data();
    for (int k = 0; k < prm.degree; ++k)
    {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i)
        {
            double s = rhs[i];
            for (int32_t q = A.ptr[i]; q < A.ptr[i + 1]; ++q)
                s -= A.val[q] * x[A.col[q]];
            r[i] = L.M[i] * s;
        }
        if (k == 0)
        {
            alpha = 1.0 / d;
            beta = 0;
        }
        else if (k == 1)
        {
            alpha = 2 * d * (1.0 / (2 * d * d - c * c));
            beta = alpha * d - 1;
        }
        else
        {
            alpha = 1.0 / (d - 0.25 * alpha * c * c);
            beta = alpha * d - 1;
        }
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i)
        {
            p[i] = alpha * r[i] + beta * p[i];
            x[i] += p[i];
        }
    }
}

void relax_apply(Level &L, const Params &prm, const double *rhs, double *x)
{
    if (prm.relax_type == 0)
    {
        chebyshev_apply(L, prm, rhs, x);
        return;
    }
    // damped jacobi: x += M (rhs - A x)
    const Csr &A = L.A;
    double *t = L.t.data();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < A.n; ++i)
    {
        double s = rhs[i];
        for (int32_t q = A.ptr[i]; q < A.ptr[i + 1]; ++q)
            s -= A.val[q] * x[A.col[q]];
        t[i] = s;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < A.n; ++i)
        x[i] += L.M[i] * t[i];
}

void dense_lu_factor(Level &L)
{
    const int64_t n = L.A.n;
    L.lu.assign(n * n, 0.0);
    L.piv.resize(n);
    for (int64_t i = 0; i < n; ++i)
        for (int32_t k = L.A.ptr[i]; k < L.A.ptr[i + 1]; ++k)
            L.lu[i * n + L.A.col[k]] = L.A.val[k];
    double *a = L.lu.data();
    for (int64_t k = 0; k < n; ++k)
    {
        int64_t p = k;
        for (int64_t i = k + 1; i < n; ++i)
            if (std::fabs(a[i * n + k]) > std::fabs(a[p * n + k]))
                p = i;
        L.piv[k] = (int)p;
        if (p != k)
            for (int64_t j = 0; j < n; ++j)
                std::swap(a[k * n + j], a[p * n + j]);
        const double inv = 1.0 / a[k * n + k];
#pragma omp parallel for schedule(static)
        for (int64_t i = k + 1; i < n; ++i)
        {
            const double f = a[i * n + k] * inv;
            a[i * n + k] = f;
            for (int64_t j = k + 1; j < n; ++j)
                a[i * n + j] -= f * a[k * n + j];
        }
    }
    L.direct = true;
}

void dense_lu_solve(const Level &L, const double *rhs, double *x)
{
    const int64_t n = L.A.n;
    const double *a = L.lu.data();
    std::vector<double> y(rhs, rhs + n);
    for (int64_t k = 0; k < n; ++k)
    {
        std::swap(y[k], y[L.piv[k]]);
        for (int64_t i = k + 1; i < n; ++i)
            y[i] -= a[i * n + k] * y[k];
    }
    for (int64_t i = n - 1; i >= 0; --i)
    {
        double s = y[i];
        for (int64_t j = i + 1; j < n; ++j)
            s -= a[i * n + j] * y[j];
        y[i] = s / a[i * n + i];
    }
    std::copy(y.begin(), y.end(), x);
}

// amgcl::amg::do_init (SURVEY A.3 "Hierarchy build")
std::unique_ptr<Hierarchy> build(Csr A0, const Params &prm, const std::vector<std::vector<int32_t>> &imposed = {})
{
    auto H = std::make_unique<Hierarchy>();
    H->prm = prm;
    std::unique_ptr<Csr> A = std::make_unique<Csr>(std::move(A0));
    double eps_strong = prm.eps_strong;
    while (A && A->n > prm.coarse_enough)
    {
        H->levels.emplace_back();
        Level &L = H->levels.back();
        L.A = std::move(*A);
        setup_relax(L, prm, 1000 + H->levels.size());
        if ((int)H->levels.size() >= prm.max_levels)
        {
            A.reset();
            goto done; // last level is a plain (smoothing-only) level
        }
        // step_down: transfer operators + Galerkin coarse operator
        std::vector<char> strong;
        int64_t nc = plain_aggregates(L.A, eps_strong, strong, L.aggr);
        const size_t li = H->levels.size() - 1;
        if (li < imposed.size() && !imposed[li].empty())
        {
            // aggregates imposed by the test (e.g. the GPU's parallel MIS-2 result): same strength flags
            L.aggr = imposed[li];
            nc = 0;
            for (int32_t a : L.aggr)
                nc = std::max<int64_t>(nc, (int64_t)a + 1);
        }
        eps_strong *= 0.5;
        if (nc == 0)
        {
            A.reset();
            goto done;
        }
        double omega = prm.sa_relax;
        if (prm.estimate_spectral_radius)
            omega *= (4.0 / 3.0) / spectral_radius(L.A, true, 0, 0);
        else
            omega *= 2.0 / 3.0;
        L.P = smoothed_prolongation(L.A, strong, L.aggr, nc, omega);
        L.R = transpose(L.P, nc);
        Csr AP = spgemm(L.A, L.P, nc);
        A = std::make_unique<Csr>(spgemm(L.R, AP, nc));
    }
    if (A)
    {
        H->levels.emplace_back();
        Level &L = H->levels.back();
        L.A = std::move(*A);
        if (prm.direct_coarse)
        {
            dense_lu_factor(L);
            L.f.assign(L.A.n, 0.0);
            L.u.assign(L.A.n, 0.0);
            L.t.assign(L.A.n, 0.0);
        }
        else
            setup_relax(L, prm, 1000 + H->levels.size());
    }
done:
    return H;
}

void residual(const Csr &A, const double *f, const double *x, double *r)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < A.n; ++i)
    {
        double s = f[i];
        for (int32_t q = A.ptr[i]; q < A.ptr[i + 1]; ++q)
            s -= A.val[q] * x[A.col[q]];
        r[i] = s;
    }
}

// amgcl::amg::cycle
void cycle(Hierarchy &H, size_t l, const double *rhs, double *x)
{
    Level &L = H.levels[l];
    const Params &prm = H.prm;
    if (l + 1 == H.levels.size())
    {
        if (L.direct)
            dense_lu_solve(L, rhs, x);
        else
        {
            for (int i = 0; i < prm.npre; ++i)
                relax_apply(L, prm, rhs, x);
            for (int i = 0; i < prm.npost; ++i)
                relax_apply(L, prm, rhs, x);
        }
        return;
    }
    Level &N = H.levels[l + 1];
    for (int j = 0; j < prm.ncycle; ++j)
    {
        for (int i = 0; i < prm.npre; ++i)
            relax_apply(L, prm, rhs, x);
        residual(L.A, rhs, x, L.t.data());
        orc::spmv_csr(L.R, L.t.data(), N.f.data());
        std::fill(N.u.begin(), N.u.end(), 0.0);
        cycle(H, l + 1, N.f.data(), N.u.data());
        const Csr &P = L.P;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < P.n; ++i)
        {
            double s = 0;
            for (int32_t q = P.ptr[i]; q < P.ptr[i + 1]; ++q)
                s += P.val[q] * N.u[P.col[q]];
            x[i] += s;
        }
        for (int i = 0; i < prm.npost; ++i)
            relax_apply(L, prm, rhs, x);
    }
}

void precond_apply(Hierarchy &H, const double *rhs, double *x)
{
    const int64_t n = H.levels[0].A.n;
    if (H.prm.pre_cycles)
    {
        std::fill(x, x + n, 0.0);
        for (int i = 0; i < H.prm.pre_cycles; ++i)
            cycle(H, 0, rhs, x);
    }
    else
        std::copy(rhs, rhs + n, x);
}

double pdot(int64_t n, const double *a, const double *b)
{
    double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < n; ++i)
        s += a[i] * b[i];
    return s;
}

} // namespace

extern "C" {

struct orc_amg_params
{
    int max_levels, coarse_enough, direct_coarse, ncycle, npre, npost, pre_cycles;
    int degree, power_iters, scale, estimate_spectral_radius, relax_type;
    double higher, lower, sa_relax, eps_strong;
};

void orc_amg_default_params(orc_amg_params *p)
{
    Params d;
    p->max_levels = d.max_levels;
    p->coarse_enough = d.coarse_enough;
    p->direct_coarse = d.direct_coarse;
    p->ncycle = d.ncycle;
    p->npre = d.npre;
    p->npost = d.npost;
    p->pre_cycles = d.pre_cycles;
    p->degree = d.degree;
    p->power_iters = d.power_iters;
    p->scale = d.scale;
    p->estimate_spectral_radius = d.estimate_spectral_radius;
    p->relax_type = d.relax_type;
    p->higher = d.higher;
    p->lower = d.lower;
    p->sa_relax = d.sa_relax;
    p->eps_strong = d.eps_strong;
}

// The matrix is passed exactly as polysolve passes it to AMGCL (AMGCL.cpp:162-166,180):
// the CSC arrays reinterpreted as CSR (valid for symmetric A; AMGCL.hpp:36-43).
void *orc_amg_create(int64_t n, const int32_t *ptr, const int32_t *col, const double *val, const orc_amg_params *p)
{
    Params prm;
    prm.max_levels = p->max_levels;
    prm.coarse_enough = p->coarse_enough;
    prm.direct_coarse = p->direct_coarse;
    prm.ncycle = p->ncycle;
    prm.npre = p->npre;
    prm.npost = p->npost;
    prm.pre_cycles = p->pre_cycles;
    prm.degree = p->degree;
    prm.power_iters = p->power_iters;
    prm.scale = p->scale;
    prm.estimate_spectral_radius = p->estimate_spectral_radius;
    prm.relax_type = p->relax_type;
    prm.higher = p->higher;
    prm.lower = p->lower;
    prm.sa_relax = p->sa_relax;
    prm.eps_strong = p->eps_strong;
    Csr A;
    A.n = n;
    A.ptr.assign(ptr, ptr + n + 1);
    A.col.assign(col, col + ptr[n]);
    A.val.assign(val, val + ptr[n]);
    try
    {
        return build(std::move(A), prm).release();
    }
    catch (...)
    {
        return nullptr;
    }
}

// Same, with aggregates imposed for the first n_imposed levels (agg[l] has lens[l] entries; lens[l] = 0 => not imposed).
void *orc_amg_create_imposed(int64_t n, const int32_t *ptr, const int32_t *col, const double *val, const orc_amg_params *p,
                             int n_imposed, const int32_t *const *agg, const int64_t *lens)
{
    Hierarchy *tmp = (Hierarchy *)orc_amg_create(0, ptr, col, val, p); // parse params only (n = 0 => empty hierarchy)
    if (!tmp)
        return nullptr;
    Params prm = tmp->prm;
    delete tmp;
    std::vector<std::vector<int32_t>> imposed(n_imposed);
    for (int l = 0; l < n_imposed; ++l)
        if (lens[l] > 0)
            imposed[l].assign(agg[l], agg[l] + lens[l]);
    Csr A;
    A.n = n;
    A.ptr.assign(ptr, ptr + n + 1);
    A.col.assign(col, col + ptr[n]);
    A.val.assign(val, val + ptr[n]);
    try
    {
        return build(std::move(A), prm, imposed).release();
    }
    catch (...)
    {
        return nullptr;
    }
}

void orc_amg_destroy(void *h) { delete (Hierarchy *)h; }
int orc_amg_num_levels(void *h) { return (int)((Hierarchy *)h)->levels.size(); }
void orc_amg_level_info(void *h, int l, int64_t *rows, int64_t *nnz, int64_t *p_nnz, double *rho, double *d, double *c)
{
    Level &L = ((Hierarchy *)h)->levels[l];
    *rows = L.A.n;
    *nnz = L.A.nnz();
    *p_nnz = L.P.nnz();
    *rho = L.rho;
    *d = L.cheb_d;
    *c = L.cheb_c;
}
void orc_amg_get_aggregates(void *h, int l, int32_t *id)
{
    Level &L = ((Hierarchy *)h)->levels[l];
    std::copy(L.aggr.begin(), L.aggr.end(), id);
}
// which: 0 = A, 1 = P, 2 = R
void orc_amg_get_matrix(void *h, int l, int which, int32_t *ptr, int32_t *col, double *val)
{
    Level &L = ((Hierarchy *)h)->levels[l];
    const Csr &M = which == 0 ? L.A : which == 1 ? L.P : L.R;
    std::copy(M.ptr.begin(), M.ptr.end(), ptr);
    std::copy(M.col.begin(), M.col.end(), col);
    std::copy(M.val.begin(), M.val.end(), val);
}
int64_t orc_amg_matrix_rows(void *h, int l, int which)
{
    Level &L = ((Hierarchy *)h)->levels[l];
    const Csr &M = which == 0 ? L.A : which == 1 ? L.P : L.R;
    return M.n;
}
void orc_amg_apply(void *h, const double *rhs, double *x) { precond_apply(*(Hierarchy *)h, rhs, x); }

// amgcl::solver::cg with the AMG hierarchy as preconditioner (SURVEY A.3 "CG").
// Returns iterations; *rel_res = ||r|| / ||b||. res_hist (optional, cap entries) receives ||r||/||b|| per iteration.
int64_t orc_amg_cg(void *h, const double *b, double *x, double tol, int64_t maxiter, double *rel_res,
                   double *res_hist, int64_t hist_cap)
{
    Hierarchy &H = *(Hierarchy *)h;
    const Csr &A = H.levels[0].A;
    const int64_t n = A.n;
    std::vector<double> r(n), s(n), p(n), q(n);
    const double norm_rhs = std::sqrt(pdot(n, b, b));
    if (norm_rhs < std::numeric_limits<double>::epsilon())
    {
        std::fill(x, x + n, 0.0);
        *rel_res = norm_rhs;
        return 0;
    }
    const double eps = std::max(tol * norm_rhs, std::numeric_limits<double>::min());
    residual(A, b, x, r.data());
    double rho1 = 2, rho2 = 1;
    double res_norm = std::sqrt(pdot(n, r.data(), r.data()));
    int64_t iter = 0;
    for (; iter < maxiter && res_norm > eps; ++iter)
    {
        precond_apply(H, r.data(), s.data());
        rho2 = rho1;
        rho1 = pdot(n, r.data(), s.data());
        if (iter)
        {
            const double bta = rho1 / rho2;
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i)
                p[i] = s[i] + bta * p[i];
        }
        else
            p = s;
        orc::spmv_csr(A, p.data(), q.data());
        const double alpha = rho1 / pdot(n, q.data(), p.data());
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i)
        {
            x[i] += alpha * p[i];
            r[i] -= alpha * q[i];
        }
        res_norm = std::sqrt(pdot(n, r.data(), r.data()));
        if (res_hist && iter < hist_cap)
            res_hist[iter] = res_norm / norm_rhs;
    }
    *rel_res = res_norm / norm_rhs;
    return iter;
}

} // extern "C"
