// MEASUREMENT ONLY -- not part of the product path and not linked into libpsb200.so.
// The GPU PCG loop the reference already ships, timed on the same box beside ours (SURVEY 8 row a9, section 2b "the bar:
// beat cuSPARSE on the same box"):
//   * cusparseSpMV, CSR, fp64, CUSPARSE_SPMV_ALG_DEFAULT            (reference MASSolver.cu:245-290)
//   * one iteration of the unfused PCG loop of MASSolver::pcg_solve  (reference MASSolver.cu:469-595): SpMV, inner product
//     (one block reduction + one fp64 atomic per 128-thread block, mas_utils/InnerProduct.cu:16-45), scalar division
//     kernels, three axpby passes (MASSolver.cu:66-81), preconditioner apply (a diagonal scaling stands in for the
//     additive-Schwarz preconditioner, which is out of scope), device-resident scalars.
// Built as libpsb200_cmp.so (links libcusparse); bench.py --with-cusparse loads it through ctypes.
#include <cuda_runtime.h>
#include <cusparse.h>

#include <cstdint>
#include <cstdio>

namespace {

#define CMP_CUDA(x)                                                                      \
    do                                                                                   \
    {                                                                                    \
        cudaError_t e_ = (x);                                                            \
        if (e_ != cudaSuccess)                                                           \
        {                                                                                \
            std::snprintf(g_err, sizeof(g_err), "%s at line %d", cudaGetErrorString(e_), __LINE__); \
            return 1;                                                                    \
        }                                                                                \
    } while (0)
#define CMP_SP(x)                                                                        \
    do                                                                                   \
    {                                                                                    \
        cusparseStatus_t s_ = (x);                                                       \
        if (s_ != CUSPARSE_STATUS_SUCCESS)                                               \
        {                                                                                \
            std::snprintf(g_err, sizeof(g_err), "cusparse status %d at line %d", (int)s_, __LINE__); \
            return 2;                                                                    \
        }                                                                                \
    } while (0)

char g_err[256] = "";

__global__ void cmp_inner_product_kernel(long long n, const double *__restrict__ a, const double *__restrict__ b, double *out)
{
    __shared__ double sm[4];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double c = i < n ? a[i] * b[i] : 0.0;
    for (int o = 16; o > 0; o >>= 1)
        c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0)
        sm[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0)
        atomicAdd(out, sm[0] + sm[1] + sm[2] + sm[3]);
}
__global__ void cmp_scalar_division_kernel(const double *num, const double *den, double *out)
{
    out[0] = fabs(den[0]) < 1e-20 ? 0.0 : num[0] / den[0];
}
// y = (ha * *da) x + (hb * *db) y
__global__ void cmp_axpby_kernel(long long n, double ha, const double *da, double hb, const double *db, const double *__restrict__ x,
                                 double *__restrict__ y)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double alpha = ha * (da ? *da : 1.0), beta = hb * (db ? *db : 1.0);
    y[i] = alpha * x[i] + beta * y[i];
}
__global__ void cmp_diag_apply_kernel(long long n, const double *__restrict__ dinv, const double *__restrict__ r, double *__restrict__ z)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        z[i] = dinv[i] * r[i];
}
__global__ void cmp_fill_kernel(long long n, double *a, double v)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        a[i] = v;
}

inline unsigned nb(long long n, int t) { return (unsigned)((n + t - 1) / t); }

} // namespace

extern "C" {

const char *psb200_cmp_last_error() { return g_err; }

// Times `reps` cusparseSpMV calls and `iters` unfused PCG iterations on the CSR matrix (host arrays; x = 1, b = A 1).
// out[0] = ms per cusparseSpMV, out[1] = ms per unfused iteration, out[2] = cuSPARSE workspace bytes,
// out[3] = checksum sum(y) of y = A 1 (to compare with ours).
int psb200_cmp_run(int64_t n, int64_t nnz, const int32_t *rp, const int32_t *ci, const double *va, int reps, int iters, double *out)
{
    g_err[0] = 0;
    int *d_rp = nullptr, *d_ci = nullptr;
    double *d_va = nullptr, *vec[7] = {}, *scal = nullptr;
    cudaStream_t st;
    CMP_CUDA(cudaStreamCreate(&st));
    CMP_CUDA(cudaMalloc(&d_rp, sizeof(int) * (n + 1)));
    CMP_CUDA(cudaMalloc(&d_ci, sizeof(int) * nnz));
    CMP_CUDA(cudaMalloc(&d_va, sizeof(double) * nnz));
    CMP_CUDA(cudaMemcpy(d_rp, rp, sizeof(int) * (n + 1), cudaMemcpyHostToDevice));
    CMP_CUDA(cudaMemcpy(d_ci, ci, sizeof(int) * nnz, cudaMemcpyHostToDevice));
    CMP_CUDA(cudaMemcpy(d_va, va, sizeof(double) * nnz, cudaMemcpyHostToDevice));
    for (auto &v : vec)
        CMP_CUDA(cudaMalloc(&v, sizeof(double) * n));
    CMP_CUDA(cudaMalloc(&scal, sizeof(double) * 8));
    CMP_CUDA(cudaMemset(scal, 0, sizeof(double) * 8));
    double *x = vec[0], *r = vec[1], *z = vec[2], *p = vec[3], *Ap = vec[4], *b = vec[5], *dinv = vec[6];
    double *rz = scal, *rz_old = scal + 1, *pAp = scal + 2, *alpha = scal + 3, *beta = scal + 4;

    cusparseHandle_t h;
    CMP_SP(cusparseCreate(&h));
    CMP_SP(cusparseSetStream(h, st));
    cusparseSpMatDescr_t A;
    CMP_SP(cusparseCreateCsr(&A, n, n, nnz, d_rp, d_ci, d_va, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F));
    cusparseDnVecDescr_t vx, vy;
    CMP_SP(cusparseCreateDnVec(&vx, n, p, CUDA_R_64F));
    CMP_SP(cusparseCreateDnVec(&vy, n, Ap, CUDA_R_64F));
    const double one = 1.0, zero = 0.0;
    size_t ws = 0;
    CMP_SP(cusparseSpMV_bufferSize(h, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, A, vx, &zero, vy, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, &ws));
    void *d_ws = nullptr;
    CMP_CUDA(cudaMalloc(&d_ws, ws ? ws : 8));
    auto spmv = [&](double *in, double *o) -> int {
        CMP_SP(cusparseDnVecSetValues(vx, in));
        CMP_SP(cusparseDnVecSetValues(vy, o));
        CMP_SP(cusparseSpMV(h, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, A, vx, &zero, vy, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, d_ws));
        return 0;
    };
    auto inner = [&](const double *a, const double *c, double *o) {
        cudaMemsetAsync(o, 0, sizeof(double), st);
        cmp_inner_product_kernel<<<nb(n, 128), 128, 0, st>>>(n, a, c, o);
    };
    auto axpby = [&](double ha, const double *da, double hb, const double *db, const double *xx, double *yy) {
        cmp_axpby_kernel<<<nb(n, 256), 256, 0, st>>>(n, ha, da, hb, db, xx, yy);
    };

    cmp_fill_kernel<<<nb(n, 256), 256, 0, st>>>(n, p, 1.0);
    cmp_fill_kernel<<<nb(n, 256), 256, 0, st>>>(n, dinv, 1.0 / 6.0);
    cmp_fill_kernel<<<nb(n, 256), 256, 0, st>>>(n, x, 0.0);
    if (spmv(p, b))
        return 2; // b = A 1
    // checksum
    cmp_fill_kernel<<<nb(n, 256), 256, 0, st>>>(n, z, 1.0);
    inner(b, z, scal + 5);
    double chk = 0;
    CMP_CUDA(cudaMemcpyAsync(&chk, scal + 5, sizeof(double), cudaMemcpyDeviceToHost, st));
    CMP_CUDA(cudaStreamSynchronize(st));
    out[3] = chk;
    out[2] = (double)ws;

    cudaEvent_t e0, e1;
    CMP_CUDA(cudaEventCreate(&e0));
    CMP_CUDA(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i)
        if (spmv(p, Ap))
            return 2;
    CMP_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i)
        if (spmv(p, Ap))
            return 2;
    CMP_CUDA(cudaEventRecord(e1, st));
    CMP_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    CMP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    out[0] = (double)ms / (reps > 0 ? reps : 1);

    // MASSolver::pcg_solve prologue (MASSolver.cu:471-483): r = b - A x, z = M^-1 r, p = z, rz = r.z
    if (spmv(x, r))
        return 2;
    axpby(1.0, nullptr, -1.0, nullptr, b, r);
    cmp_diag_apply_kernel<<<nb(n, 256), 256, 0, st>>>(n, dinv, r, z);
    CMP_CUDA(cudaMemcpyAsync(p, z, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    inner(r, z, rz);
    auto iteration = [&]() -> int {
        if (spmv(p, Ap))
            return 2;
        inner(p, Ap, pAp);
        cmp_scalar_division_kernel<<<1, 1, 0, st>>>(rz, pAp, alpha);
        axpby(1.0, alpha, 1.0, nullptr, p, x);
        axpby(-1.0, alpha, 1.0, nullptr, Ap, r);
        cmp_diag_apply_kernel<<<nb(n, 256), 256, 0, st>>>(n, dinv, r, z);
        cudaMemcpyAsync(rz_old, rz, sizeof(double), cudaMemcpyDeviceToDevice, st);
        inner(r, z, rz);
        cmp_scalar_division_kernel<<<1, 1, 0, st>>>(rz, rz_old, beta);
        axpby(1.0, nullptr, 1.0, beta, z, p);
        return 0;
    };
    for (int i = 0; i < 3; ++i)
        if (iteration())
            return 2;
    CMP_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i)
        if (iteration())
            return 2;
    CMP_CUDA(cudaEventRecord(e1, st));
    CMP_CUDA(cudaEventSynchronize(e1));
    CMP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    out[1] = (double)ms / (iters > 0 ? iters : 1);
    CMP_CUDA(cudaGetLastError());

    cusparseDestroyDnVec(vx);
    cusparseDestroyDnVec(vy);
    cusparseDestroySpMat(A);
    cusparseDestroy(h);
    cudaFree(d_ws);
    cudaFree(d_rp);
    cudaFree(d_ci);
    cudaFree(d_va);
    for (auto &v : vec)
        cudaFree(v);
    cudaFree(scal);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    return 0;
}

} // extern "C"
