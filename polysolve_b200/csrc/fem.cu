// Dirichlet pre-processing on the GPU: the three helpers of the reference's FEMSolver
// (src/polysolve/linear/FEMSolver.cpp:97-372) that sit immediately before the hot path in PolyFEM.
//   dirichlet_solve              :97-300   A~ = A with the rows and columns of the Dirichlet dofs set to identity,
//                                          g = f - (I - N) A N f, analyze + factorize + solve(g, u), f := g
//   prefactorize                 :303-343  A~ only, analyze + factorize
//   dirichlet_solve_prefactorized:345-372  g from the matrix the caller passes, solve(g, u), f := g
// The reference rebuilds the matrix from triplets on the host (dropping the masked entries, adding a full diagonal);
// here the values are masked in place on the device on the pattern of A (masked entries become explicit zeros, the
// diagonal of a Dirichlet row becomes 1), and the lifting is one masked SpMV with a fused epilogue. The numbers that
// reach the Krylov loop are the same: a zero entry contributes +0.0 to every product sum.
#include "../../include/psb200.h"
#include "capi_internal.hpp"
#include "solver.hpp"

#include <vector>

namespace psb {

namespace {

__global__ void mask_from_nodes_kernel(long long count, const int *__restrict__ nodes, int n, unsigned char *__restrict__ mask, int *bad)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count)
        return;
    const int k = nodes[i];
    if (k < 0 || k >= n)
        atomicExch(bad, 1);
    else
        mask[k] = 1;
}

// A~_ij = a_ij if neither i nor j is a Dirichlet dof, else (i == j ? 1 : 0)   (FEMSolver.cpp:131-150). 8 lanes per row.
__global__ void dirichlet_mask_kernel(CsrView A, double *__restrict__ va, const unsigned char *__restrict__ mask, int *missing)
{
    const int lane = threadIdx.x & 7;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (row >= A.n)
        return;
    const bool mr = mask[row] != 0;
    bool found = false;
    for (int k = A.rp[row] + lane; k < A.rp[row + 1]; k += 8)
    {
        const int j = A.ci[k];
        if (j == (int)row)
            found = true;
        if (mr || mask[j])
            va[k] = (j == (int)row) ? 1.0 : 0.0;
    }
    for (int o = 4; o > 0; o >>= 1)
        found |= (bool)__shfl_xor_sync(0xffffffffu, (int)found, o);
    if (lane == 0 && mr && !found)
        atomicExch(missing, 1); // the reference would insert the diagonal entry; the pattern here is fixed
}

// xN = N f (double2 lanes over the padded length; the mask is padded with zeros)
__global__ void masked_copy_kernel(long long n, const double *__restrict__ f, const unsigned char *__restrict__ mask, double *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = mask[i] ? f[i] : 0.0;
}

} // namespace

// g_i = f_i for a Dirichlet dof, f_i - (A N f)_i otherwise   (FEMSolver.cpp:113-124)
struct EpiDirichletLift
{
    static constexpr int NV = 0;
    using Pre = Pre2;
    double *g;
    const double *f;
    const unsigned char *mask;
    __device__ __forceinline__ Pre pre(int row) const { return {__ldg(f + row), mask[row] ? 1.0 : 0.0}; }
    __device__ __forceinline__ void operator()(int row, double s, Pre p, double (&)[1]) const { g[row] = p.b != 0.0 ? p.a : p.a - s; }
};

void Solver::dirichlet_set_nodes(const int *nodes, long long count)
{
    if (dist)
        throw std::runtime_error("psb200 dirichlet: not available on the row-partitioned path");
    if (count < 0 || (count > 0 && !nodes))
        throw std::invalid_argument("psb200 dirichlet: null node list");
    cudaStream_t st = ctx.stream;
    dmask.alloc((size_t)n_pad + 16, true);
    if (count == 0)
        return;
    DevBuf<int> d_nodes;
    d_nodes.alloc((size_t)count);
    PSB_CUDA(cudaMemcpyAsync(d_nodes.p, nodes, sizeof(int) * count, cudaMemcpyHostToDevice, st));
    int *d_bad = (int *)ctx.counter.p + 3;
    PSB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    mask_from_nodes_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(count, d_nodes.p, (int)n, dmask.p, d_bad);
    check_launch();
    int bad = 0;
    PSB_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaStreamSynchronize(st));
    if (bad)
        throw std::invalid_argument("psb200 dirichlet: node id out of range");
}

// values of the caller's matrix -> A.va (unmasked), on the analyzed pattern
void Solver::dirichlet_upload(long long n_, long long nnz_, const int *outer, const int *inner, const double *vals, int precond_num_)
{
    ensure_ctx(*this);
    if (dist)
        throw std::runtime_error("psb200 dirichlet: not available on the row-partitioned path");
    analyze_pattern(n_, nnz_, outer, inner, precond_num_);
    csc_vals.alloc(std::max<long long>(nnz_, 1));
    if (nnz_)
    {
        PSB_CUDA(cudaMemcpyAsync(csc_vals.p, vals, sizeof(double) * nnz_, cudaMemcpyHostToDevice, ctx.stream));
        gather_values_to_csr(csc_vals.p);
    }
    ensure_vectors();
}

// d_g = d_f - (I - N) A (N d_f) with the matrix currently in A.va; vp is the scratch for N f
void Solver::dirichlet_lift(const double *d_f, double *d_g)
{
    cudaStream_t st = ctx.stream;
    PSB_CUDA(cudaMemsetAsync(vp.p, 0, sizeof(double) * n_pad, st));
    if (n > 0)
    {
        masked_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, d_f, dmask.p, vp.p);
        check_launch();
        launch_spmv(ctx, "dirichlet_lift", A, vp.p, EpiDirichletLift{d_g, d_f, dmask.p}, FinNone{});
    }
}

void Solver::dirichlet_mask_matrix()
{
    cudaStream_t st = ctx.stream;
    if (n == 0)
        return;
    int *d_missing = (int *)ctx.counter.p + 3;
    PSB_CUDA(cudaMemsetAsync(d_missing, 0, sizeof(int), st));
    dirichlet_mask_kernel<<<(unsigned)((n * 8 + 255) / 256), 256, 0, st>>>(A.view(), A.va.p, dmask.p, d_missing);
    check_launch();
    int missing = 0;
    PSB_CUDA(cudaMemcpyAsync(&missing, d_missing, sizeof(int), cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaStreamSynchronize(st));
    if (missing)
        throw std::runtime_error("psb200 dirichlet: a Dirichlet row has no structural diagonal entry");
}

void Solver::dirichlet_solve(long long n_, long long nnz_, const int *outer, const int *inner, const double *vals, double *f, const int *nodes,
                             long long n_nodes, double *u, int precond_num_)
{
    if (!f || !u || (!vals && nnz_ > 0))
        throw std::invalid_argument("psb200_dirichlet_solve: null argument");
    dirichlet_upload(n_, nnz_, outer, inner, vals, precond_num_);
    dirichlet_set_nodes(nodes, n_nodes);
    cudaStream_t st = ctx.stream;
    const double t0 = 0;
    vz.alloc((size_t)n_pad, true);
    PSB_CUDA(cudaMemcpyAsync(vz.p, f, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    dirichlet_lift(vz.p, vb.p);      // g with the unmasked matrix
    dirichlet_mask_matrix();         // A~ in place
    factorize_tail(t0);
    PSB_CUDA(cudaMemcpyAsync(vx.p, u, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    run_solver(vb.p);
    PSB_CUDA(cudaMemcpyAsync(u, vx.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaMemcpyAsync(f, vb.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st)); // f = g (FEMSolver.cpp:282)
    PSB_CUDA(cudaStreamSynchronize(st));
    build_info();
}

void Solver::dirichlet_prefactorize(long long n_, long long nnz_, const int *outer, const int *inner, const double *vals, const int *nodes,
                                    long long n_nodes, int precond_num_)
{
    if (!vals && nnz_ > 0)
        throw std::invalid_argument("psb200_dirichlet_prefactorize: null values");
    dirichlet_upload(n_, nnz_, outer, inner, vals, precond_num_);
    dirichlet_set_nodes(nodes, n_nodes);
    dirichlet_mask_matrix();
    factorize_tail(0);
}

void Solver::dirichlet_solve_prefactorized(const double *vals, double *f, double *u, long long n_)
{
    if (!factorized || dmask.p == nullptr)
        throw std::runtime_error("psb200_dirichlet_solve_prefactorized: psb200_dirichlet_prefactorize() first");
    if (n_ != n || !f || !u)
        throw std::invalid_argument("psb200_dirichlet_solve_prefactorized: size mismatch or null vector");
    cudaStream_t st = ctx.stream;
    ensure_vectors();
    vz.alloc((size_t)n_pad, true);
    PSB_CUDA(cudaMemcpyAsync(vz.p, f, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    if (vals)
    {
        // lift with the matrix the caller passes (the reference multiplies with its `A` argument, FEMSolver.cpp:361):
        // its CSR values go to a scratch array, the factorized masked matrix stays untouched
        DevBuf<double> tmp_va;
        tmp_va.alloc(std::max<long long>(nnz, 1), false, 64);
        PSB_CUDA(cudaMemcpyAsync(csc_vals.p, vals, sizeof(double) * nnz_global, cudaMemcpyHostToDevice, st));
        std::swap(tmp_va.p, A.va.p);
        const bool bsr_was_ready = A.bsr_ready;
        A.bsr_ready = false; // the scratch values exist in scalar CSR form only: the lifting product takes the CSR schedule
        try
        {
            gather_values_to_csr(csc_vals.p);
            dirichlet_lift(vz.p, vb.p);
            PSB_CUDA(cudaStreamSynchronize(st));
        }
        catch (...)
        {
            std::swap(tmp_va.p, A.va.p);
            A.bsr_ready = bsr_was_ready;
            throw;
        }
        std::swap(tmp_va.p, A.va.p);
        A.bsr_ready = bsr_was_ready;
    }
    else
        dirichlet_lift(vz.p, vb.p); // the resident masked matrix: the lifting term vanishes, g = f
    PSB_CUDA(cudaMemcpyAsync(vx.p, u, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    run_solver(vb.p);
    PSB_CUDA(cudaMemcpyAsync(u, vx.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaMemcpyAsync(f, vb.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaStreamSynchronize(st));
    build_info();
}

} // namespace psb

// ====================================================================================== C ABI
namespace {
template <class F>
int fem_guarded(psb200_handle h, F &&f)
{
    if (!h)
        return PSB200_ERR_INVALID;
    try
    {
        h->s.err.clear();
        psb::DeviceScope device_scope(h->s.device, h->s.ctx.stream != nullptr);
        psb::AllocScope alloc_scope(h->s.ctx.stream);
        f(h->s);
        return PSB200_OK;
    }
    catch (const psb::CudaError &e)
    {
        h->s.err = e.what();
        cudaGetLastError();
        return PSB200_ERR_CUDA;
    }
    catch (const std::invalid_argument &e)
    {
        h->s.err = e.what();
        return PSB200_ERR_INVALID;
    }
    catch (const std::exception &e)
    {
        h->s.err = e.what();
        return PSB200_ERR_NUMERIC;
    }
}
} // namespace

extern "C" {

int psb200_dirichlet_solve(psb200_handle h, int64_t n, int64_t nnz, const int32_t *outer, const int32_t *inner, const double *vals,
                           double *f_inout, const int32_t *dirichlet_nodes, int64_t n_nodes, double *u_inout, int precond_num)
{
    return fem_guarded(h, [&](psb::Solver &s) { s.dirichlet_solve(n, nnz, outer, inner, vals, f_inout, dirichlet_nodes, n_nodes, u_inout, precond_num); });
}

int psb200_dirichlet_prefactorize(psb200_handle h, int64_t n, int64_t nnz, const int32_t *outer, const int32_t *inner, const double *vals,
                                  const int32_t *dirichlet_nodes, int64_t n_nodes, int precond_num)
{
    return fem_guarded(h, [&](psb::Solver &s) { s.dirichlet_prefactorize(n, nnz, outer, inner, vals, dirichlet_nodes, n_nodes, precond_num); });
}

int psb200_dirichlet_solve_prefactorized(psb200_handle h, const double *vals_or_null, double *f_inout, double *u_inout, int64_t n)
{
    return fem_guarded(h, [&](psb::Solver &s) { s.dirichlet_solve_prefactorized(vals_or_null, f_inout, u_inout, n); });
}

} // extern "C"
