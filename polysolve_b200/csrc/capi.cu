// extern "C" surface of libpsb200.so -- see include/psb200.h for the contract of every entry point.
#include "../../include/psb200.h"
#include "amg.hpp"
#include "solver.hpp"

#include <cstring>
#include <new>

#include "capi_internal.hpp"

namespace {
thread_local std::string g_create_error;

template <class F>
int guarded(psb200_handle h, F &&f)
{
    if (!h)
        return PSB200_ERR_INVALID;
    try
    {
        h->s.err.clear();
        psb::DeviceScope device_scope(h->s.device, h->s.ctx.stream != nullptr);
        psb::AllocScope alloc_scope(h->s.ctx.stream); // buffers are allocated / freed in order on the solver's stream
        h->s.check_not_poisoned();
        f(h->s);
        return PSB200_OK;
    }
    catch (const psb::CommError &e)
    {
        h->s.err = e.what();
        return PSB200_ERR_COMM;
    }
    catch (const psb::CudaError &e)
    {
        h->s.err = e.what();
        cudaGetLastError(); // clear sticky-free errors
        return PSB200_ERR_CUDA;
    }
    catch (const std::invalid_argument &e)
    {
        h->s.err = e.what();
        return PSB200_ERR_INVALID;
    }
    catch (const std::bad_alloc &)
    {
        h->s.err = "host out of memory";
        return PSB200_ERR_CUDA;
    }
    catch (const std::exception &e)
    {
        h->s.err = e.what();
        return PSB200_ERR_NUMERIC;
    }
    catch (...)
    {
        h->s.err = "unknown error";
        return PSB200_ERR_NUMERIC;
    }
}
} // namespace

extern "C" {

int psb200_create(psb200_handle *out, const char *json_params)
{
    if (!out)
        return PSB200_ERR_INVALID;
    *out = nullptr;
    psb200_solver *h = new (std::nothrow) psb200_solver();
    if (!h)
        return PSB200_ERR_CUDA;
    int rc = PSB200_OK;
    if (json_params && *json_params)
        rc = guarded(h, [&](psb::Solver &s) { s.set_parameters(json_params); });
    if (rc != PSB200_OK)
    {
        g_create_error = h->s.err;
        delete h;
        return rc;
    }
    *out = h;
    return PSB200_OK;
}

int psb200_destroy(psb200_handle h)
{
    if (h)
    {
        psb::DeviceScope device_scope(h->s.device, h->s.ctx.stream != nullptr);
        delete h;
    }
    return PSB200_OK;
}

int psb200_set_parameters(psb200_handle h, const char *json)
{
    return guarded(h, [&](psb::Solver &s) { s.set_parameters(json ? json : ""); });
}

int psb200_set_tolerance(psb200_handle h, double tol)
{
    return guarded(h, [&](psb::Solver &s) { s.prm.tolerance = tol; });
}

int psb200_set_block_size(psb200_handle h, int block_size)
{
    return guarded(h, [&](psb::Solver &s) {
        if (block_size < 1)
            throw std::invalid_argument("psb200_set_block_size: block_size < 1");
        s.prm.block_size = block_size;
    });
}

int psb200_analyze_pattern_csc(psb200_handle h, int64_t n, int64_t nnz, const int32_t *outer, const int32_t *inner, int precond_num)
{
    return guarded(h, [&](psb::Solver &s) { s.analyze_pattern(n, nnz, outer, inner, precond_num); });
}

int psb200_factorize_csc(psb200_handle h, int64_t n, int64_t nnz, const int32_t *outer, const int32_t *inner, const double *vals)
{
    return guarded(h, [&](psb::Solver &s) { s.factorize(n, nnz, outer, inner, vals); });
}

int psb200_factorize_csc_device(psb200_handle h, int64_t n, int64_t nnz, const double *d_vals, double diag_shift)
{
    return guarded(h, [&](psb::Solver &s) { s.factorize_device(n, nnz, d_vals, diag_shift); });
}

int psb200_residual_norm_device(psb200_handle h, const double *d_x, const double *d_b, int64_t n, double *norm_out)
{
    return guarded(h, [&](psb::Solver &s) {
        const double r = s.residual_norm_device(d_x, d_b, n);
        if (norm_out)
            *norm_out = r;
    });
}

int psb200_residual_norm(psb200_handle h, const double *x, const double *b, int64_t n, double *norm_out)
{
    return guarded(h, [&](psb::Solver &s) {
        const double r = s.residual_norm_host(x, b, n);
        if (norm_out)
            *norm_out = r;
    });
}

int psb200_dist_allgather(psb200_handle h, double *x_full_inout, int64_t n)
{
    return guarded(h, [&](psb::Solver &s) { s.dist_allgather_host(x_full_inout, n); });
}

int psb200_solve(psb200_handle h, const double *b, double *x, int64_t n)
{
    return guarded(h, [&](psb::Solver &s) { s.solve_host(b, x, n); });
}

int psb200_solve_device(psb200_handle h, const double *d_b, double *d_x, int64_t n)
{
    return guarded(h, [&](psb::Solver &s) { s.solve_device(d_b, d_x, n); });
}

int psb200_get_info(psb200_handle h, char *json_out, size_t cap, size_t *needed)
{
    if (!h)
        return PSB200_ERR_INVALID;
    h->s.build_info();
    const std::string &j = h->s.info_json;
    if (needed)
        *needed = j.size() + 1;
    if (!json_out || cap < j.size() + 1)
    {
        h->s.err = "psb200_get_info: buffer too small";
        return PSB200_ERR_INVALID;
    }
    std::memcpy(json_out, j.c_str(), j.size() + 1);
    return PSB200_OK;
}

// Returns the memory cached in the device's stream-ordered pool to the driver (buffers in use are unaffected).
int psb200_release_cached_memory(psb200_handle h)
{
    return guarded(h, [&](psb::Solver &s) {
        if (!s.ctx.stream)
            return;
        PSB_CUDA(cudaStreamSynchronize(s.ctx.stream));
        cudaMemPool_t pool = nullptr;
        PSB_CUDA(cudaDeviceGetDefaultMemPool(&pool, s.device));
        PSB_CUDA(cudaMemPoolTrimTo(pool, 0));
    });
}

const char *psb200_name(psb200_handle) { return "CUDA"; }

const char *psb200_last_error(psb200_handle h) { return h ? h->s.err.c_str() : g_create_error.c_str(); }

int psb200_debug_get_csr(psb200_handle h, int32_t *row_ptr, int32_t *col_idx, int32_t *perm)
{
    return guarded(h, [&](psb::Solver &s) {
        if (!s.analyzed)
            throw std::invalid_argument("psb200_debug_get_csr: analyze_pattern() first");
        PSB_CUDA(cudaStreamSynchronize(s.ctx.stream));
        if (row_ptr)
            PSB_CUDA(cudaMemcpy(row_ptr, s.A.rp.p, sizeof(int) * (s.n + 1), cudaMemcpyDeviceToHost));
        if (col_idx && s.nnz)
            PSB_CUDA(cudaMemcpy(col_idx, s.A.ci.p, sizeof(int) * s.nnz, cudaMemcpyDeviceToHost));
        if (perm && s.nnz)
            PSB_CUDA(cudaMemcpy(perm, s.perm.p, sizeof(int) * s.nnz, cudaMemcpyDeviceToHost));
    });
}

int psb200_spmv(psb200_handle h, const double *x, double *y, int64_t n)
{
    return guarded(h, [&](psb::Solver &s) { s.spmv_host(x, y, n); });
}

int psb200_bench_spmv(psb200_handle h, const char *kernel, int reps, double *ms_avg)
{
    return guarded(h, [&](psb::Solver &s) {
        const double ms = s.bench_spmv(kernel ? kernel : "", reps);
        if (ms_avg)
            *ms_avg = ms;
    });
}

void *psb200_get_stream(psb200_handle h) { return h ? (void *)h->s.ctx.stream : nullptr; }

int psb200_debug_set_aggregates(psb200_handle h, int level, const int32_t *agg, int64_t n)
{
    return guarded(h, [&](psb::Solver &s) {
        if (level < 0 || level > 32)
            throw std::invalid_argument("psb200_debug_set_aggregates: bad level");
        if ((int)s.imposed_aggregates.size() <= level)
            s.imposed_aggregates.resize(level + 1);
        if (agg && n > 0)
            s.imposed_aggregates[level].assign(agg, agg + n);
        else
            s.imposed_aggregates[level].clear();
    });
}

int psb200_debug_get_level(psb200_handle h, int level, int which, int64_t *rows, int64_t *cols, int64_t *nnz, int32_t *row_ptr,
                           int32_t *col_idx, double *vals)
{
    return guarded(h, [&](psb::Solver &s) {
        if (!s.amg)
            throw std::invalid_argument("psb200_debug_get_level: no AMG hierarchy");
        if (level < 0 || level >= s.amg->num_levels() || which < 0 || which > 2)
            throw std::invalid_argument("psb200_debug_get_level: bad level/which");
        const psb::CsrDev &M = s.amg->matrix(level, which);
        PSB_CUDA(cudaStreamSynchronize(s.ctx.stream));
        if (rows)
            *rows = M.n;
        if (cols)
            *cols = M.ncols;
        if (nnz)
            *nnz = M.nnz;
        if (row_ptr && M.n)
            PSB_CUDA(cudaMemcpy(row_ptr, M.rp.p, sizeof(int) * ((size_t)M.n + 1), cudaMemcpyDeviceToHost));
        if (col_idx && M.nnz)
            PSB_CUDA(cudaMemcpy(col_idx, M.ci.p, sizeof(int) * M.nnz, cudaMemcpyDeviceToHost));
        if (vals && M.nnz)
            PSB_CUDA(cudaMemcpy(vals, M.va.p, sizeof(double) * M.nnz, cudaMemcpyDeviceToHost));
    });
}

int psb200_debug_get_aggregates(psb200_handle h, int level, int32_t *agg, int64_t n, int64_t *n_agg)
{
    return guarded(h, [&](psb::Solver &s) {
        if (!s.amg || level < 0 || level >= s.amg->num_levels())
            throw std::invalid_argument("psb200_debug_get_aggregates: no AMG hierarchy / bad level");
        int na = 0;
        const int *d = s.amg->aggregates(level, &na);
        if (n_agg)
            *n_agg = na;
        PSB_CUDA(cudaStreamSynchronize(s.ctx.stream));
        if (agg && d && n > 0)
            PSB_CUDA(cudaMemcpy(agg, d, sizeof(int) * n, cudaMemcpyDeviceToHost));
    });
}

int psb200_precond_apply(psb200_handle h, const double *r, double *z, int64_t n)
{
    return guarded(h, [&](psb::Solver &s) { s.precond_apply_host(r, z, n); });
}

} // extern "C"
