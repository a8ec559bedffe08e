// CSR SpMV kernels (fp64 values, int32 indices) with fused per-row epilogues and fused reductions.
// Two schedules:
//   * stream : TMA-staged tiles for short rows (stencils / FEM scalar problems, <= ~8 nnz per row)
//   * vector : LPR lanes per row (LPR = 1 .. 32) for everything else (coarse AMG levels, R, dense rows)
// Included from kernels.cuh (needs CsrView, RedCtx, grid_reduce).
#pragma once

namespace psb {

// ---------------------------------------------------------------------------------- epilogues
// An epilogue is applied once per row by the lane that holds the row sum:
//   Pre  pre(row)                      -- loads the per-row side inputs; issued BEFORE the gathers so
//                                         their latency overlaps with the x gathers
//   void operator()(row, sum, pre, acc) -- stores outputs, accumulates NV fused reduction values
struct NoPre
{
};
struct EpiStore
{
    static constexpr int NV = 0;
    using Pre = NoPre;
    double *y;
    __device__ __forceinline__ Pre pre(int) const { return {}; }
    __device__ __forceinline__ void operator()(int row, double s, Pre, double (&)[1]) const { y[row] = s; }
};
// y = A x and u.y  (PCG: p.Ap; BiCGSTAB: r0.v)
struct EpiDot
{
    static constexpr int NV = 1;
    using Pre = double;
    double *y;
    const double *u;
    __device__ __forceinline__ Pre pre(int row) const { return __ldg(u + row); }
    __device__ __forceinline__ void operator()(int row, double s, Pre ur, double (&acc)[1]) const
    {
        y[row] = s;
        acc[0] += ur * s;
    }
};
// t = A z, t.t and t.s (BiCGSTAB omega)
struct EpiDot2
{
    static constexpr int NV = 2;
    using Pre = double;
    double *y;
    const double *u;
    __device__ __forceinline__ Pre pre(int row) const { return __ldg(u + row); }
    __device__ __forceinline__ void operator()(int row, double s, Pre ur, double (&acc)[2]) const
    {
        y[row] = s;
        acc[0] += s * s;
        acc[1] += s * ur;
    }
};
struct Pre2
{
    double a, b;
};
struct Pre4
{
    double a, b, c, d;
};
// r = b - A x with ||r||^2, ||b||^2 and r.(dinv r)
struct EpiResidualNorms
{
    static constexpr int NV = 3;
    using Pre = Pre2;
    double *r;
    const double *b;
    const double *dinv;
    __device__ __forceinline__ Pre pre(int row) const { return {__ldg(b + row), __ldg(dinv + row)}; }
    __device__ __forceinline__ void operator()(int row, double s, Pre p, double (&acc)[3]) const
    {
        const double ri = p.a - s;
        r[row] = ri;
        acc[0] += ri * ri;
        acc[1] += p.a * p.a;
        acc[2] += ri * ri * p.b;
    }
};
// BiCGSTAB restart: r = b - A x, r0 = r, ||r||^2
struct EpiResidualRestart
{
    static constexpr int NV = 1;
    using Pre = double;
    double *r, *r0;
    const double *b;
    __device__ __forceinline__ Pre pre(int row) const { return __ldg(b + row); }
    __device__ __forceinline__ void operator()(int row, double s, Pre bi, double (&acc)[1]) const
    {
        const double ri = bi - s;
        r[row] = ri;
        r0[row] = ri;
        acc[0] += ri * ri;
    }
};
// r = b - A x (AMG residual before restriction)
struct EpiResidual
{
    static constexpr int NV = 0;
    using Pre = double;
    double *r;
    const double *b;
    __device__ __forceinline__ Pre pre(int row) const { return __ldg(b + row); }
    __device__ __forceinline__ void operator()(int row, double s, Pre bi, double (&)[1]) const { r[row] = bi - s; }
};
// One Chebyshev step fused into the SpMV (amgcl relaxation/chebyshev.hpp solve()):
//   res = M (b - A x_in);  p = alpha res + beta p;  x_out = x_in + p.     x_in != x_out (ping-pong).
struct EpiCheb
{
    static constexpr int NV = 0;
    using Pre = Pre4;
    const double *b, *dinv, *xin;
    double *p, *xout;
    double alpha, beta;
    __device__ __forceinline__ Pre pre(int row) const
    {
        return {__ldg(b + row), __ldg(dinv + row), __ldg(xin + row), beta != 0.0 ? p[row] : 0.0};
    }
    __device__ __forceinline__ void operator()(int row, double s, Pre q, double (&)[1]) const
    {
        const double res = q.b * (q.a - s);
        const double pn = alpha * res + beta * q.d;
        p[row] = pn;
        xout[row] = q.c + pn;
    }
};
// Damped Jacobi / generic diagonal relaxation: x_out = x_in + w[row] (b - A x_in)
struct EpiRelaxDiag
{
    static constexpr int NV = 0;
    using Pre = Pre4;
    const double *b, *w, *xin;
    double *xout;
    __device__ __forceinline__ Pre pre(int row) const { return {__ldg(b + row), __ldg(w + row), __ldg(xin + row), 0.0}; }
    __device__ __forceinline__ void operator()(int row, double s, Pre q, double (&)[1]) const { xout[row] = q.c + q.b * (q.a - s); }
};
// x += P u  (prolongation-and-correct)
struct EpiAddTo
{
    static constexpr int NV = 0;
    using Pre = double;
    double *x;
    __device__ __forceinline__ Pre pre(int row) const { return x[row]; }
    __device__ __forceinline__ void operator()(int row, double s, Pre xr, double (&)[1]) const { x[row] = xr + s; }
};
// power iteration on D^-1 A : b1 = dinv (A b0); ||b1||^2 and sum |b1_i b0_i| (amgcl spectral_radius)
struct EpiPower
{
    static constexpr int NV = 2;
    using Pre = Pre2;
    double *b1;
    const double *b0, *dinv;
    __device__ __forceinline__ Pre pre(int row) const { return {__ldg(b0 + row), __ldg(dinv + row)}; }
    __device__ __forceinline__ void operator()(int row, double s, Pre q, double (&acc)[2]) const
    {
        const double v = s * q.b;
        b1[row] = v;
        acc[0] += v * v;
        acc[1] += fabs(v * q.a);
    }
};

// ---------------------------------------------------------------------------------- vector schedule
// LPR lanes cooperate on one row (LPR = 1 is the scalar thread-per-row schedule). Grid-stride over rows.
template <class Epi, class Fin, int LPR, int THREADS>
__global__ void __launch_bounds__(THREADS) spmv_vector_kernel(CsrView A, const double *__restrict__ x, Epi epi, RedCtx rc,
                                                              Fin fin, const int *done, const int *only_if)
{
    griddep_launch_dependents();
    griddep_wait();
    if (done && *done)
        return;
    if (only_if && !*only_if)
        return;
    constexpr int NVA = Epi::NV > 0 ? Epi::NV : 1;
    double acc[NVA];
#pragma unroll
    for (int i = 0; i < NVA; ++i)
        acc[i] = 0;
    const double *xh = wait_halo(A, rc.comm);
    const int lane = threadIdx.x % LPR;
    const int rows_per_cta = THREADS / LPR;
    for (long long base = (long long)blockIdx.x * rows_per_cta; base < A.n; base += (long long)gridDim.x * rows_per_cta)
    {
        const int row = (int)base + threadIdx.x / LPR;
        double s = 0;
        typename Epi::Pre pre{};
        if (row < A.n)
        {
            const int kb = __ldg(A.rp + row), ke = __ldg(A.rp + row + 1);
            if (lane == 0)
                pre = epi.pre(row);
            int k = kb + lane;
            // two gathers in flight per lane
            for (; k + LPR < ke; k += 2 * LPR)
            {
                const int c0 = __ldg(A.ci + k), c1 = __ldg(A.ci + k + LPR);
                const double v0 = __ldg(A.va + k), v1 = __ldg(A.va + k + LPR);
                const double x0 = ldx(x, xh, A.nl, c0), x1 = ldx(x, xh, A.nl, c1);
                s += v0 * x0;
                s += v1 * x1;
            }
            if (k < ke)
                s += __ldg(A.va + k) * ldx(x, xh, A.nl, __ldg(A.ci + k));
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1)
            s += __shfl_xor_sync(0xffffffffu, s, o);
        if (row < A.n && lane == 0)
            epi(row, s, pre, acc);
    }
    double tot[NVA];
    if (grid_reduce<Epi::NV, THREADS>(acc, rc, tot) && threadIdx.x == 0)
        fin(tot);
}

// ---------------------------------------------------------------------------------- TMA-staged stream schedule
// A CTA walks tiles of THREADS consecutive rows (tile t -> CTA t mod grid, so the CTAs that run
// together work on neighbouring rows and share x through L2). For every tile one thread issues three
// 1-D TMA bulk copies (cp.async.bulk -> SASS UBLKCP) STAGES tiles ahead: the tile's slice of row_ptr,
// and its contiguous val / col ranges (16-byte aligned outward; the arrays are over-allocated).
// Completion is tracked by one mbarrier per stage (expect_tx bytes). So HBM streams the matrix at
// full line efficiency whatever the row length, nothing in the per-row dependency chain touches DRAM
// except the x gathers, and those are issued 8 at a time; each thread owns one row, which makes the
// gathers of a warp hit consecutive addresses for banded matrices.
// Tiles whose nnz exceed CAP fall back to direct global loads (correct for any matrix).
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar, unsigned long long policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

__device__ __forceinline__ double lds_f64(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds_s32(unsigned addr)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// Sum of the entries k, k + LPR, k + 2 LPR, ... < ke of one row from the staged tile (sva / sca: shared addresses of
// entry 0 of the value / column arrays), x gathered from global memory (HALO: columns >= nl come from the halo buffer).
template <int LPR, bool HALO>
__device__ __forceinline__ double wide_row_sum(unsigned sva, unsigned sca, int k, int ke, const double *__restrict__ x,
                                               const double *__restrict__ xh, int nl)
{
    auto gather = [&](int c) -> double {
        if constexpr (HALO)
            return c < nl ? __ldg(x + c) : __ldcg(xh + (c - nl));
        else
            return __ldg(x + c);
    };
    double sum = 0;
    for (; k + 3 * LPR < ke; k += 4 * LPR)
    {
        int c[4];
        double v[4], xx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
            c[u] = lds_s32(sca + 4u * (unsigned)(k + u * LPR));
            v[u] = lds_f64(sva + 8u * (unsigned)(k + u * LPR));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            xx[u] = gather(c[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            sum += v[u] * xx[u];
    }
    for (; k < ke; k += LPR)
        sum += lds_f64(sva + 8u * (unsigned)k) * gather(lds_s32(sca + 4u * (unsigned)k));
    return sum;
}

// LPR lanes cooperate on one row, so a tile holds THREADS / LPR rows: LPR = 1 for stencil-like rows (<= ~8 nnz),
// LPR = 2 .. 16 for the 30-80 nnz rows of Galerkin coarse levels and vector-valued FEM (the staged nnz per tile
// stays within CAP while every lane still has several gathers in flight).
template <int THREADS, int CAP, int STAGES, int LPR = 1>
struct StreamCfg
{
    static constexpr int threads = THREADS, cap = CAP, stages = STAGES, lpr = LPR, rows = THREADS / LPR;
    static constexpr int rp_ints = rows + 4; // rows + 1 row pointers, rounded to 16 bytes
    static constexpr size_t val_bytes = (size_t)STAGES * CAP * sizeof(double);
    static constexpr size_t col_bytes = (size_t)STAGES * CAP * sizeof(int);
    static constexpr size_t rp_bytes = (size_t)STAGES * rp_ints * sizeof(int);
    static constexpr size_t bytes = val_bytes + col_bytes + rp_bytes + 128;
};

template <class Epi, class Fin, class Cfg>
__global__ void __launch_bounds__(Cfg::threads) spmv_stream_kernel(CsrView A, const double *__restrict__ x, Epi epi, RedCtx rc, Fin fin,
                                                                   const int *done, const int *only_if)
{
    constexpr int THREADS = Cfg::threads, CAP = Cfg::cap, STAGES = Cfg::stages, LPR = Cfg::lpr, ROWS = Cfg::rows;
    griddep_launch_dependents();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sval = reinterpret_cast<double *>(smem_raw);
    int *scol = reinterpret_cast<int *>(smem_raw + Cfg::val_bytes);
    int *srp = reinterpret_cast<int *>(smem_raw + Cfg::val_bytes + Cfg::col_bytes);
    __shared__ __align__(8) unsigned long long bar[STAGES];

    constexpr int NVA = Epi::NV > 0 ? Epi::NV : 1;
    double acc[NVA];
#pragma unroll
    for (int i = 0; i < NVA; ++i)
        acc[i] = 0;

    const int ntiles = (A.n + ROWS - 1) / ROWS;
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
            mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // one thread: launch the three bulk copies of `tile` into stage s
    auto issue = [&](int tile, int s) {
        const int r0 = tile * ROWS;
        const int r1 = min(A.n, r0 + ROWS);
        const int k0 = __ldg(A.rp + r0), k1 = __ldg(A.rp + r1);
        const int ka = k0 & ~3;
        const int cnt4 = (k1 - ka + 3) & ~3;
        const unsigned rp_b = (unsigned)(((r1 - r0 + 1) + 3) & ~3) * 4u; // row_ptr array is over-allocated
        unsigned bytes = rp_b;
        const bool staged = cnt4 > 0 && cnt4 <= CAP;
        if (staged)
            bytes += (unsigned)cnt4 * 12u;
        mbar_expect_tx(&bar[s], bytes);
        tma_bulk_g2s(srp + (size_t)s * Cfg::rp_ints, A.rp + r0, rp_b, &bar[s], policy);
        if (staged)
        {
            tma_bulk_g2s(sval + (size_t)s * CAP, A.va + ka, (unsigned)cnt4 * 8u, &bar[s], policy);
            tma_bulk_g2s(scol + (size_t)s * CAP, A.ci + ka, (unsigned)cnt4 * 4u, &bar[s], policy);
        }
    };

    auto tile_at = [&](int pos) { return A.tile_order ? __ldg(A.tile_order + pos) : pos; };
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
        {
            const int pos = blockIdx.x + s * gridDim.x;
            if (pos < ntiles)
                issue(tile_at(pos), s);
        }
    }

    // Everything above touches only the matrix, which no kernel of the chain writes: under a programmatic dependent
    // launch it overlaps with the tail of the predecessor. From here on its results (x, the `done` flag) are needed.
    griddep_wait();
    if ((done && *done) || (only_if && !*only_if))
    {
        // the solve is over (or this launch is predicated off): drain the copies already in flight, then leave
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
            if ((int)blockIdx.x + s * (int)gridDim.x < ntiles)
                mbar_wait(&bar[s], 0);
        return;
    }

    // Row partitions: the halo values are needed by the boundary tiles only. With an interior-first order the wait
    // happens right before this CTA's first boundary tile; otherwise here (the TMA prefetch is already in flight).
    const double *xh = nullptr;
    bool halo_ready = A.halo_mask == 0;
    if (!halo_ready && A.tile_order == nullptr)
    {
        xh = wait_halo(A, rc.comm);
        halo_ready = true;
    }

    int it = 0;
    for (int pos = blockIdx.x; pos < ntiles; pos += gridDim.x, ++it)
    {
        const int tile = tile_at(pos);
        if (!halo_ready && pos >= A.n_interior)
        {
            xh = wait_halo(A, rc.comm);
            halo_ready = true;
        }
        const int s = it % STAGES;
        const unsigned parity = (it / STAGES) & 1;
        const int r0 = tile * ROWS;
        const int nrow = min(A.n - r0, ROWS);
        const int trow = threadIdx.x / LPR, lane = threadIdx.x % LPR;
        const int row = r0 + trow;
        const bool live = trow < nrow;
        typename Epi::Pre pre{};
        if (live && lane == 0)
            pre = epi.pre(row); // side inputs first: their latency hides behind the barrier wait and the gathers
        mbar_wait(&bar[s], parity);
        const int *rps = srp + (size_t)s * Cfg::rp_ints;
        const int k0 = rps[0], k1 = rps[nrow];
        const int ka = k0 & ~3;
        const bool staged = ((k1 - ka + 3) & ~3) <= CAP;
        int kb = 0, ke = 0;
        if (live)
        {
            kb = rps[trow];
            ke = rps[trow + 1];
        }
        double sum = 0;
        if (staged)
        {
            const double *sv = sval + (size_t)s * CAP - ka;
            const int *sc = scol + (size_t)s * CAP - ka;
            if constexpr (LPR == 1)
            {
                // 8 gathers in flight per trip; products are still added in k order
                for (int k = kb; k < ke; k += 8)
                {
                    double v[8], xx[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                    {
                        const bool ok = k + u < ke;
                        const int c = ok ? sc[k + u] : 0;
                        v[u] = ok ? sv[k + u] : 0.0;
                        xx[u] = ok ? ldx(x, xh, A.nl, c) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (k + u < ke)
                            sum += v[u] * xx[u];
                }
            }
            else
            {
                // lane l takes entries kb + l, kb + l + LPR, ...: conflict-free shared-memory reads, 4 gathers in flight.
                // The loop is lean on purpose (ncu: the first version issued ~27 instructions per entry and was
                // issue-bound at 0.62 of the HBM peak): 32-bit shared addresses computed once per tile, full trips
                // without predicates, one code path per halo mode. Products are added in the same order as before.
                const unsigned sva = smem_u32(sval + (size_t)s * CAP) - 8u * (unsigned)ka;
                const unsigned sca = smem_u32(scol + (size_t)s * CAP) - 4u * (unsigned)ka;
                int k = kb + lane;
                if (A.halo_mask == 0)
                    sum = wide_row_sum<LPR, false>(sva, sca, k, ke, x, xh, A.nl);
                else
                    sum = wide_row_sum<LPR, true>(sva, sca, k, ke, x, xh, A.nl);
            }
        }
        else
        {
            for (int k = kb + lane; k < ke; k += LPR)
                sum += __ldg(A.va + k) * ldx(x, xh, A.nl, __ldg(A.ci + k));
        }
        if constexpr (LPR > 1)
        {
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1)
                sum += __shfl_xor_sync(0xffffffffu, sum, o);
        }
        if (live && lane == 0)
            epi(row, sum, pre, acc);
        __syncthreads(); // every thread is done reading stage s
        if (threadIdx.x == 0)
        {
            const int next = pos + STAGES * gridDim.x;
            if (next < ntiles)
                issue(tile_at(next), s);
        }
    }
    double tot[NVA];
    if (grid_reduce<Epi::NV, THREADS>(acc, rc, tot) && threadIdx.x == 0)
        fin(tot);
}

// ---------------------------------------------------------------------------------- BSR-3 stream schedule
// Block problems (3 dofs per node: elasticity, AMGCL_Block<3>, reference AMGCL.cpp:246-298; the reference's own GPU path
// stores BSR too, mas_utils/BSRMatrix.cu:195-474, CuSparseWrapper.hpp:123-156). One 3 x 3 block = 72 B of values + ONE
// 4-byte block column: 76 B per block instead of 9 x 12 = 108 B of scalar CSR, and the three x values of a block column
// are gathered once and reused by the three rows (a third of the gather instructions of the scalar schedule).
// Same TMA pipeline as spmv_stream_kernel: a tile is ROWS block rows; the slice of the block row pointer, the value range
// (72 B per block) and the block-column range are staged by three bulk copies per tile, two tiles ahead. LPB lanes share
// a block row: lane l takes blocks kb + l, kb + l + LPB, ...; three row sums are reduced with shuffles and the lane that
// holds them applies the (scalar-row) epilogue three times.
struct BsrView
{
    const int *brp;    // block row pointer, nb + 1
    const int *bci;    // block column (= scalar column / 3)
    const double *bva; // 9 values per block, row-major
    int nb;            // block rows
    int nl;            // local SCALAR columns (row partitions: scalar columns >= nl are halo columns)
    unsigned halo_mask;
    const int *tile_order; // optional: sequence position -> tile (row partitions: boundary tiles first)
};

template <int THREADS, int CAPB, int STAGES, int LPB>
struct BsrCfg
{
    static constexpr int threads = THREADS, capb = CAPB, stages = STAGES, lpb = LPB, rows = THREADS / LPB;
    static constexpr int rp_ints = rows + 4;
    static constexpr size_t val_bytes = (size_t)STAGES * CAPB * 9 * sizeof(double);
    static constexpr size_t col_bytes = (size_t)STAGES * CAPB * sizeof(int);
    static constexpr size_t rp_bytes = (size_t)STAGES * rp_ints * sizeof(int);
    static constexpr size_t bytes = val_bytes + col_bytes + rp_bytes + 128;
};

template <class Epi, class Fin, class Cfg>
__global__ void __launch_bounds__(Cfg::threads) spmv_bsr3_kernel(BsrView A, const double *__restrict__ x, Epi epi, RedCtx rc, Fin fin,
                                                                 const int *done, const int *only_if)
{
    constexpr int THREADS = Cfg::threads, CAPB = Cfg::capb, STAGES = Cfg::stages, LPB = Cfg::lpb, ROWS = Cfg::rows;
    griddep_launch_dependents();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sval = reinterpret_cast<double *>(smem_raw);
    int *scol = reinterpret_cast<int *>(smem_raw + Cfg::val_bytes);
    int *srp = reinterpret_cast<int *>(smem_raw + Cfg::val_bytes + Cfg::col_bytes);
    __shared__ __align__(8) unsigned long long bar[STAGES];

    constexpr int NVA = Epi::NV > 0 ? Epi::NV : 1;
    double acc[NVA];
#pragma unroll
    for (int i = 0; i < NVA; ++i)
        acc[i] = 0;
    const int ntiles = (A.nb + ROWS - 1) / ROWS;
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
            mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int tile, int s) {
        const int r0 = tile * ROWS;
        const int r1 = min(A.nb, r0 + ROWS);
        const int k0 = __ldg(A.brp + r0), k1 = __ldg(A.brp + r1);
        const int ka = k0 & ~3;                    // 4 blocks = 288 B of values / 16 B of columns: both 16-byte aligned
        const int cnt4 = (k1 - ka + 3) & ~3;
        const unsigned rp_b = (unsigned)(((r1 - r0 + 1) + 3) & ~3) * 4u;
        unsigned bytes = rp_b;
        const bool staged = cnt4 > 0 && cnt4 <= CAPB;
        if (staged)
            bytes += (unsigned)cnt4 * 76u;
        mbar_expect_tx(&bar[s], bytes);
        tma_bulk_g2s(srp + (size_t)s * Cfg::rp_ints, A.brp + r0, rp_b, &bar[s], policy);
        if (staged)
        {
            tma_bulk_g2s(sval + (size_t)s * CAPB * 9, A.bva + (size_t)ka * 9, (unsigned)cnt4 * 72u, &bar[s], policy);
            tma_bulk_g2s(scol + (size_t)s * CAPB, A.bci + ka, (unsigned)cnt4 * 4u, &bar[s], policy);
        }
    };
    auto tile_at = [&](int pos) { return A.tile_order ? __ldg(A.tile_order + pos) : pos; };
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
        {
            const int pos = blockIdx.x + s * gridDim.x;
            if (pos < ntiles)
                issue(tile_at(pos), s);
        }
    }
    griddep_wait();
    if ((done && *done) || (only_if && !*only_if))
    {
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
            if ((int)blockIdx.x + s * (int)gridDim.x < ntiles)
                mbar_wait(&bar[s], 0);
        return;
    }
    const double *xh = nullptr;
    if (A.halo_mask != 0)
    {
        wait_pushes_landed(rc.comm);
        xh = rc.comm.halo(rc.comm.rank, (int)(*rc.comm.push_epoch % kHaloBufs), 0);
    }
    const int nl = A.nl;
    int it = 0;
    for (int pos = blockIdx.x; pos < ntiles; pos += gridDim.x, ++it)
    {
        const int s = it % STAGES;
        const unsigned parity = (it / STAGES) & 1;
        const int r0 = tile_at(pos) * ROWS;
        const int nrow = min(A.nb - r0, ROWS);
        const int trow = threadIdx.x / LPB, lane = threadIdx.x % LPB;
        const int brow = r0 + trow;
        const bool live = trow < nrow;
        typename Epi::Pre pre0{}, pre1{}, pre2{};
        if (live && lane == 0)
        {
            pre0 = epi.pre(3 * brow);
            pre1 = epi.pre(3 * brow + 1);
            pre2 = epi.pre(3 * brow + 2);
        }
        mbar_wait(&bar[s], parity);
        const int *rps = srp + (size_t)s * Cfg::rp_ints;
        const int k0 = rps[0], k1 = rps[nrow];
        const int ka = k0 & ~3;
        const bool staged = ((k1 - ka + 3) & ~3) <= CAPB;
        int kb = 0, ke = 0;
        if (live)
        {
            kb = rps[trow];
            ke = rps[trow + 1];
        }
        double s0 = 0, s1 = 0, s2 = 0;
        auto gather3 = [&](int bc, double &x0, double &x1, double &x2) {
            const int c = 3 * bc;
            const double *src = (c < nl) ? (x + c) : (xh + (c - nl));
            if (c < nl)
            {
                x0 = __ldg(src);
                x1 = __ldg(src + 1);
                x2 = __ldg(src + 2);
            }
            else
            {
                x0 = __ldcg(src);
                x1 = __ldcg(src + 1);
                x2 = __ldcg(src + 2);
            }
        };
        if (staged)
        {
            const double *sv = sval + (size_t)s * CAPB * 9 - (size_t)ka * 9;
            const int *sc = scol + (size_t)s * CAPB - ka;
            int k = kb + lane;
            for (; k + LPB < ke; k += 2 * LPB)
            {
                // two blocks in flight per lane
                const int c0 = sc[k], c1 = sc[k + LPB];
                double a0, a1, a2, b0, b1, b2;
                gather3(c0, a0, a1, a2);
                gather3(c1, b0, b1, b2);
                const double *v = sv + (size_t)k * 9, *w = sv + (size_t)(k + LPB) * 9;
                s0 += v[0] * a0 + v[1] * a1 + v[2] * a2;
                s1 += v[3] * a0 + v[4] * a1 + v[5] * a2;
                s2 += v[6] * a0 + v[7] * a1 + v[8] * a2;
                s0 += w[0] * b0 + w[1] * b1 + w[2] * b2;
                s1 += w[3] * b0 + w[4] * b1 + w[5] * b2;
                s2 += w[6] * b0 + w[7] * b1 + w[8] * b2;
            }
            if (k < ke)
            {
                double a0, a1, a2;
                gather3(sc[k], a0, a1, a2);
                const double *v = sv + (size_t)k * 9;
                s0 += v[0] * a0 + v[1] * a1 + v[2] * a2;
                s1 += v[3] * a0 + v[4] * a1 + v[5] * a2;
                s2 += v[6] * a0 + v[7] * a1 + v[8] * a2;
            }
        }
        else
        {
            for (int k = kb + lane; k < ke; k += LPB)
            {
                double a0, a1, a2;
                gather3(__ldg(A.bci + k), a0, a1, a2);
                const double *v = A.bva + (size_t)k * 9;
                s0 += __ldg(v) * a0 + __ldg(v + 1) * a1 + __ldg(v + 2) * a2;
                s1 += __ldg(v + 3) * a0 + __ldg(v + 4) * a1 + __ldg(v + 5) * a2;
                s2 += __ldg(v + 6) * a0 + __ldg(v + 7) * a1 + __ldg(v + 8) * a2;
            }
        }
#pragma unroll
        for (int o = LPB / 2; o > 0; o >>= 1)
        {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (live && lane == 0)
        {
            epi(3 * brow, s0, pre0, acc);
            epi(3 * brow + 1, s1, pre1, acc);
            epi(3 * brow + 2, s2, pre2, acc);
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            const int next = pos + STAGES * gridDim.x;
            if (next < ntiles)
                issue(tile_at(next), s);
        }
    }
    double tot[NVA];
    if (grid_reduce<Epi::NV, THREADS>(acc, rc, tot) && threadIdx.x == 0)
        fin(tot);
}

} // namespace psb
