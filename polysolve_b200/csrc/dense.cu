// Dense coarse-grid solve of the AMG hierarchy ("direct_coarse": true; reference linear-solver-spec.json:363,
// AMGCL.cpp:45; AMGCL itself runs a skyline LU there). The coarsest matrix (<= coarse_enough = 3000 rows) is SPD for the
// problems the CG path accepts, so it is factorised A = L L^T by a blocked right-looking Cholesky whose panel and
// trailing updates, the triangular inverse X = L^-1 and the product Z = X^T X = A^-1 all run through ONE fp64 tensor-core
// GEMM kernel (mma.sync m8n8k4 DMMA -- the only dense, contraction-shaped work on this path, and the only place tensor
// cores are used; tcgen05 has no fp64 type). Every visit of the coarsest level is then a single dense GEMV with Z.
#include "amg_internal.hpp"

namespace psb {

namespace {

constexpr int NB = 64; // block size of the factorisation and tile size of the GEMM

__global__ void dense_from_csr_kernel(CsrView A, int np, double *__restrict__ D)
{
    const int row = blockIdx.x;
    if (row >= np)
        return;
    for (int j = threadIdx.x; j < np; j += blockDim.x)
        D[(size_t)row * np + j] = (row >= A.n && j == row) ? 1.0 : 0.0; // identity padding
    __syncthreads();
    if (row < A.n)
        for (int k = A.rp[row] + threadIdx.x; k < A.rp[row + 1]; k += blockDim.x)
            D[(size_t)row * np + A.ci[k]] = A.va[k];
}

// Cholesky of the NB x NB diagonal block k (in place, lower) and its inverse Linv (NB x NB, lower). One CTA of 256 threads.
__global__ void __launch_bounds__(256) potrf_diag_kernel(double *__restrict__ D, int np, int k, double *__restrict__ Linv, int *bad)
{
    __shared__ double a[NB][NB + 1];
    double *w = Linv + (size_t)k * NB * NB; // the inverse is built in place in global memory (a thread re-reads only its own column)
    double *blk = D + (size_t)k * NB * np + (size_t)k * NB;
    for (int e = threadIdx.x; e < NB * NB; e += 256)
        a[e / NB][e % NB] = blk[(size_t)(e / NB) * np + e % NB];
    __syncthreads();
    for (int j = 0; j < NB; ++j)
    {
        if (threadIdx.x == 0)
        {
            const double d = a[j][j];
            if (!(d > 0.0))
                *bad = 1;
            a[j][j] = sqrt(d > 0.0 ? d : 1.0);
        }
        __syncthreads();
        const double dj = a[j][j];
        for (int i = j + 1 + threadIdx.x; i < NB; i += 256)
            a[i][j] /= dj;
        __syncthreads();
        // trailing update of the lower triangle
        for (int e = threadIdx.x; e < (NB - j - 1) * (NB - j - 1); e += 256)
        {
            const int i = j + 1 + e / (NB - j - 1), c = j + 1 + e % (NB - j - 1);
            if (c <= i)
                a[i][c] -= a[i][j] * a[c][j];
        }
        __syncthreads();
    }
    // inverse of the lower factor: column c of W solves L y = e_c (forward substitution), one thread per column
    if (threadIdx.x < NB)
    {
        const int c = threadIdx.x;
        for (int i = 0; i < NB; ++i)
        {
            double s = (i == c) ? 1.0 : 0.0;
            for (int q = c; q < i; ++q)
                s -= a[i][q] * w[q * NB + c];
            w[i * NB + c] = i < c ? 0.0 : s / a[i][i];
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < NB * NB; e += 256)
    {
        const int i = e / NB, c = e % NB;
        blk[(size_t)i * np + c] = c <= i ? a[i][c] : 0.0;
    }
}

// C(i, j) = alpha C(i, j) + beta sum_k A(i, k) B(k, j) on 64 x 64 tiles, fp64 tensor cores (DMMA m8n8k4).
// A(i, k) = A[i * ars + k * acs], B(k, j) = B[k * brs + j * bcs], C row-major with leading dimension ldc. M, N multiples
// of 64, K a multiple of 16. lower_only: tiles above the diagonal are skipped. A tile may alias C when K covers it
// entirely inside one CTA (accumulators are written after the last read).
struct GemmArgs
{
    const double *A, *B;
    double *C;
    long long ars, acs, brs, bcs, ldc;
    int M, N, K;
    double alpha, beta;
    int lower_only;
};
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256) dmma_gemm_kernel(GemmArgs g)
{
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (g.lower_only && tj > ti)
        return;
    __shared__ double As[64][17];
    __shared__ double Bs[16][65];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = warp >> 1, wn = warp & 1; // warp tile: rows 16 wm .. +16, cols 32 wn .. +32
    double acc[2][4][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
            acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    const long long i0 = (long long)ti * 64, j0 = (long long)tj * 64;
    for (int k0 = 0; k0 < g.K; k0 += 16)
    {
        for (int e = threadIdx.x; e < 64 * 16; e += 256)
        {
            const int r = e / 16, kk = e % 16;
            As[r][kk] = g.A[(i0 + r) * g.ars + (long long)(k0 + kk) * g.acs];
            const int kr = e / 64, c = e % 64;
            Bs[kr][c] = g.B[(long long)(k0 + kr) * g.brs + (j0 + c) * g.bcs];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk += 4)
        {
            double a[2], b[4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
                a[mi] = As[16 * wm + 8 * mi + gid][kk + tig];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
                b[ni] = Bs[kk + tig][32 * wn + 8 * ni + gid];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
        {
            const long long r = i0 + 16 * wm + 8 * mi + gid, c = j0 + 32 * wn + 8 * ni + 2 * tig;
            double *p = g.C + r * g.ldc + c;
            p[0] = (g.alpha != 0.0 ? g.alpha * p[0] : 0.0) + g.beta * acc[mi][ni][0];
            p[1] = (g.alpha != 0.0 ? g.alpha * p[1] : 0.0) + g.beta * acc[mi][ni][1];
        }
}

void gemm(cudaStream_t st, const GemmArgs &g)
{
    if (g.M <= 0 || g.N <= 0 || g.K <= 0)
        return;
    dmma_gemm_kernel<<<dim3(g.N / 64, g.M / 64), 256, 0, st>>>(g);
    check_launch();
}

// x = Z f: one warp per row (Z is symmetric, n x n inside an np x np array)
__global__ void dense_gemv_kernel(int n, int np, const double *__restrict__ Z, const double *__restrict__ f, double *__restrict__ x, const int *done)
{
    if (done && *done)
        return;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n)
        return;
    double s = 0;
    for (int j = lane; j < n; j += 32)
        s += Z[(size_t)row * np + j] * f[j];
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0)
        x[row] = s;
}

} // namespace

// Z = A^-1 for the SPD matrix A (n <= a few thousand rows); Z is (np x np), np = n rounded up to 64
void dense_inverse_build(Ctx &c, const CsrDev &A, DevBuf<double> &Z, int &np_out)
{
    cudaStream_t st = c.stream;
    const int n = A.n;
    if (n <= 0)
        throw std::runtime_error("psb200 amg: direct coarse solve of an empty level");
    if (n > 8192)
        throw std::runtime_error("psb200 amg: direct_coarse supports at most 8192 coarse rows (lower coarse_enough)");
    const int np = (n + NB - 1) / NB * NB, nbk = np / NB;
    np_out = np;
    DevBuf<double> D, Linv, X, T;
    D.alloc((size_t)np * np);
    Linv.alloc((size_t)nbk * NB * NB);
    X.alloc((size_t)np * np, true);
    T.alloc((size_t)NB * np, true);
    Z.alloc((size_t)np * np);
    int *d_bad = (int *)c.counter.p + 3;
    PSB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    dense_from_csr_kernel<<<np, 256, 0, st>>>(A.view(), np, D.p);
    check_launch();
    const long long ld = np;
    for (int k = 0; k < nbk; ++k)
    {
        potrf_diag_kernel<<<1, 256, 0, st>>>(D.p, np, k, Linv.p, d_bad);
        check_launch();
        const int below = nbk - k - 1;
        if (below == 0)
            break;
        double *panel = D.p + (size_t)(k + 1) * NB * np + (size_t)k * NB; // A_ik, i > k
        // panel <- panel Linv_k^T : B(kk, j) = Linv_k[j][kk]
        gemm(st, GemmArgs{panel, Linv.p + (size_t)k * NB * NB, panel, ld, 1, 1, NB, ld, below * NB, NB, NB, 0.0, 1.0, 0});
        // trailing lower tiles A_ij -= L_ik L_jk^T : B(kk, j) = panel[j][kk]
        double *trail = D.p + (size_t)(k + 1) * NB * np + (size_t)(k + 1) * NB;
        gemm(st, GemmArgs{panel, panel, trail, ld, 1, 1, ld, ld, below * NB, below * NB, NB, 1.0, -1.0, 1});
    }
    if (d2h(c, d_bad))
        throw std::runtime_error("psb200 amg: direct_coarse needs a symmetric positive definite coarse matrix (Cholesky pivot <= 0)");
    // X = L^-1 by block rows: X_ii = Linv_i, X_i,<i = -Linv_i (L_i,<i X_<i,<i)
    for (int i = 0; i < nbk; ++i)
    {
        PSB_CUDA(cudaMemcpy2DAsync(X.p + (size_t)i * NB * np + (size_t)i * NB, sizeof(double) * np, Linv.p + (size_t)i * NB * NB, sizeof(double) * NB,
                                   sizeof(double) * NB, NB, cudaMemcpyDeviceToDevice, st));
        if (i == 0)
            continue;
        const double *Li = D.p + (size_t)i * NB * np; // L_i,0:i
        gemm(st, GemmArgs{Li, X.p, T.p, ld, 1, ld, 1, ld, NB, i * NB, i * NB, 0.0, 1.0, 0});
        gemm(st, GemmArgs{Linv.p + (size_t)i * NB * NB, T.p, X.p + (size_t)i * NB * np, NB, 1, ld, 1, ld, NB, i * NB, NB, 0.0, -1.0, 0});
    }
    // Z = X^T X : A(i, k) = X[k][i]
    gemm(st, GemmArgs{X.p, X.p, Z.p, 1, ld, ld, 1, ld, np, np, np, 0.0, 1.0, 0});
    PSB_CUDA(cudaStreamSynchronize(st));
}

void dense_inverse_apply(Ctx &c, int n, int np, const double *Z, const double *f, double *x, const int *done)
{
    c.prof_begin("coarse_direct");
    dense_gemv_kernel<<<(n * 32 + 255) / 256, 256, 0, c.stream>>>(n, np, Z, f, x, done);
    check_launch();
    c.prof_end();
}

} // namespace psb
