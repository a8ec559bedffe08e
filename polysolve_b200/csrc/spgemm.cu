// Row-wise sparse matrix product C = A * B for the Galerkin triple product of the AMG setup (A P, then R (A P)):
// Gustavson's algorithm with one hash accumulator per row in SHARED memory, hand-written for sm_100a.
// (The reference obtains this from AMGCL's CPU spgemm inside amgcl::coarsening::galerkin; round 1 of this repo used
// expand -> cub radix sort -> compress, which moves every intermediate product through HBM about 14 times.)
//
//   symbolic : rows binned by their product count T_i (an upper bound of the row length); a group of G lanes owns one row
//              and a key-only open-addressing table of H >= T_i slots; distinct keys are counted -> row pointer of C
//   numeric  : rows binned by their exact length m_i; table of H >= 2 m_i (key, value) slots. The entries a_ik of the row
//              are walked SEQUENTIALLY, the lanes of the group spread over the entries of row k of B: within one step
//              all products have distinct columns, so the accumulation needs no floating-point atomics, and every output
//              entry is summed in ascending k -- the order of the sequential algorithm (and of the sort-based path, whose
//              results this reproduces bit for bit). The row is then ranked by column inside the table and written
//              sorted.
// Matrices with a row beyond the largest bin fall back to the sort-based product (amg.cu), any matrix is handled.
#include "amg_internal.hpp"

#include <cub/device/device_scan.cuh>

namespace psb {

namespace {

constexpr int kEmptyKey = -1;
constexpr int kSymBins = 5, kNumBins = 4;
// symbolic bins by T_i:      <= 32      <= 128     <= 512     <= 2048    <= 8192
// numeric bins by m_i:       <= 16      <= 64      <= 256     <= 1024
__device__ __constant__ int c_sym_limit[kSymBins] = {32, 128, 512, 2048, 8192};
__device__ __constant__ int c_num_limit[kNumBins] = {16, 64, 256, 1024};
const int h_sym_limit[kSymBins] = {32, 128, 512, 2048, 8192};
const int h_num_limit[kNumBins] = {16, 64, 256, 1024};

inline int nblk(long long n, int t = 256) { return (int)std::max<long long>(1, (n + t - 1) / t); }

__device__ __forceinline__ unsigned slot_of(int c, int H) { return ((unsigned)c * 2654435761u) & (unsigned)(H - 1); }

template <int G>
__device__ __forceinline__ unsigned group_mask()
{
    if (G >= 32)
        return 0xffffffffu;
    const int lane = threadIdx.x & 31;
    return ((1u << G) - 1u) << (lane / G * G);
}

// T_i = sum over the entries of row i of A of the length of the matching row of B (clamped to 2^30); max over rows
__global__ void product_count_kernel(CsrView A, const int *__restrict__ b_rp, int *__restrict__ T, int *maxT)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int t = 0;
    if (i < A.n)
    {
        long long c = 0;
        for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
        {
            const int a = A.ci[k];
            c += b_rp[a + 1] - b_rp[a];
        }
        t = (int)min(c, (long long)(1 << 30));
        T[i] = t;
    }
    for (int o = 16; o > 0; o >>= 1)
        t = max(t, __shfl_xor_sync(0xffffffffu, t, o));
    if ((threadIdx.x & 31) == 0 && t > 0)
        atomicMax(maxT, t);
}

template <int NB>
__device__ __forceinline__ int bin_of(int v, const int *limit)
{
    if (v <= 0)
        return -1;
#pragma unroll
    for (int b = 0; b < NB; ++b)
        if (v <= limit[b])
            return b;
    return NB; // beyond the largest bin
}
template <int NB, bool SYM>
__global__ void bin_count_kernel(int n, const int *__restrict__ key, int *__restrict__ counts)
{
    __shared__ int sc[NB + 1];
    if (threadIdx.x <= NB)
        sc[threadIdx.x] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        const int b = bin_of<NB>(key[i], SYM ? c_sym_limit : c_num_limit);
        if (b >= 0)
            atomicAdd(&sc[b], 1);
    }
    __syncthreads();
    if (threadIdx.x <= NB && sc[threadIdx.x])
        atomicAdd(&counts[threadIdx.x], sc[threadIdx.x]);
}
template <int NB, bool SYM>
__global__ void bin_fill_kernel(int n, const int *__restrict__ key, const int *__restrict__ offs, int *__restrict__ cursor, int *__restrict__ rows)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int b = bin_of<NB>(key[i], SYM ? c_sym_limit : c_num_limit);
    if (b >= 0 && b < NB)
        rows[offs[b] + atomicAdd(&cursor[b], 1)] = i;
}

// ---------------------------------------------------------------------------------- symbolic
template <int G, int H, int THREADS>
__global__ void __launch_bounds__(THREADS) spgemm_symbolic_kernel(CsrView A, const int *__restrict__ b_rp, const int *__restrict__ b_ci,
                                                                   const int *__restrict__ rows, int nrows, int *__restrict__ row_len)
{
    extern __shared__ int sh_keys[];
    constexpr int RPC = THREADS / G;
    const int g = threadIdx.x / G, lane = threadIdx.x % G;
    int *tab = sh_keys + (size_t)g * H;
    for (int s = lane; s < H; s += G)
        tab[s] = kEmptyKey;
    const unsigned mask = group_mask<G>();
    __syncwarp(mask);
    const int r = blockIdx.x * RPC + g;
    if (r >= nrows)
        return;
    const int i = rows[r];
    int count = 0;
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
    {
        const int a = __ldg(A.ci + k);
        const int q1 = __ldg(b_rp + a + 1);
        for (int q = __ldg(b_rp + a) + lane; q < q1; q += G)
        {
            const int c = __ldg(b_ci + q);
            unsigned s = slot_of(c, H);
            for (;;)
            {
                const int cur = tab[s];
                if (cur == c)
                    break;
                if (cur == kEmptyKey)
                {
                    const int old = atomicCAS(&tab[s], kEmptyKey, c);
                    if (old == kEmptyKey)
                    {
                        ++count;
                        break;
                    }
                    if (old == c)
                        break;
                }
                s = (s + 1) & (unsigned)(H - 1);
            }
        }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1)
        count += __shfl_xor_sync(mask, count, o);
    if (lane == 0)
        row_len[i] = count;
}

// ---------------------------------------------------------------------------------- numeric
template <int G, int H, int THREADS>
__global__ void __launch_bounds__(THREADS) spgemm_numeric_kernel(CsrView A, CsrView B, const int *__restrict__ rows, int nrows,
                                                                  const int *__restrict__ c_rp, int *__restrict__ c_ci, double *__restrict__ c_va)
{
    extern __shared__ __align__(16) unsigned char sh_raw[];
    constexpr int RPC = THREADS / G;
    double *all_vals = reinterpret_cast<double *>(sh_raw);
    int *all_keys = reinterpret_cast<int *>(sh_raw + sizeof(double) * (size_t)RPC * H);
    const int g = threadIdx.x / G, lane = threadIdx.x % G;
    int *keys = all_keys + (size_t)g * H;
    double *vals = all_vals + (size_t)g * H;
    for (int s = lane; s < H; s += G)
    {
        keys[s] = kEmptyKey;
        vals[s] = 0.0;
    }
    const unsigned mask = group_mask<G>();
    __syncwarp(mask);
    const int r = blockIdx.x * RPC + g;
    if (r >= nrows)
        return;
    const int i = rows[r];
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
    {
        const int a = __ldg(A.ci + k);
        const double va = __ldg(A.va + k);
        const int q1 = __ldg(B.rp + a + 1);
        for (int q = __ldg(B.rp + a) + lane; q < q1; q += G)
        {
            const int c = __ldg(B.ci + q);
            const double v = __dmul_rn(va, __ldg(B.va + q)); // product rounded on its own (no FMA contraction): the sort-based path and the CPU oracle round it too
            unsigned s = slot_of(c, H);
            for (;;)
            {
                const int cur = keys[s];
                if (cur == c)
                    break;
                if (cur == kEmptyKey)
                {
                    const int old = atomicCAS(&keys[s], kEmptyKey, c);
                    if (old == kEmptyKey || old == c)
                        break;
                }
                s = (s + 1) & (unsigned)(H - 1);
            }
            vals[s] += v; // the columns of one row of B are distinct: no other lane touches this slot in this step
        }
        __syncwarp(mask); // the next entry of the row may hit the same columns: steps are ordered
    }
    // rank the occupied slots by column and write the row sorted
    const int base = c_rp[i];
    for (int s = lane; s < H; s += G)
    {
        const int key = keys[s];
        if (key == kEmptyKey)
            continue;
        int rank = 0;
        for (int t = 0; t < H; ++t)
        {
            const int kt = keys[t];
            rank += (kt != kEmptyKey && kt < key);
        }
        c_ci[base + rank] = key;
        c_va[base + rank] = vals[s];
    }
}

template <int G, int H, int THREADS>
void launch_symbolic(cudaStream_t st, const CsrDev &A, const CsrDev &B, const int *rows, int nrows, int *row_len)
{
    if (nrows <= 0)
        return;
    constexpr int RPC = THREADS / G;
    const size_t smem = sizeof(int) * (size_t)RPC * H;
    auto kern = spgemm_symbolic_kernel<G, H, THREADS>;
    PSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(nrows + RPC - 1) / RPC, THREADS, smem, st>>>(A.view(), B.rp.p, B.ci.p, rows, nrows, row_len);
    check_launch();
}
template <int G, int H, int THREADS>
void launch_numeric(cudaStream_t st, const CsrDev &A, const CsrDev &B, const int *rows, int nrows, CsrDev &C)
{
    if (nrows <= 0)
        return;
    constexpr int RPC = THREADS / G;
    const size_t smem = (sizeof(int) + sizeof(double)) * (size_t)RPC * H;
    auto kern = spgemm_numeric_kernel<G, H, THREADS>;
    PSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(nrows + RPC - 1) / RPC, THREADS, smem, st>>>(A.view(), B.view(), rows, nrows, C.rp.p, C.ci.p, C.va.p);
    check_launch();
}

// rows grouped by bin: rows[offs[b] .. offs[b + 1]) are the rows of bin b (any order); returns the number of rows beyond
// the largest bin
template <int NB, bool SYM>
int bin_rows(Ctx &c, int n, const int *key, DevBuf<int> &rows, int (&offs)[NB + 2])
{
    cudaStream_t st = c.stream;
    DevBuf<int> counts, d_offs, cursor;
    counts.alloc(NB + 1, true);
    cursor.alloc(NB + 1, true);
    d_offs.alloc(NB + 2);
    bin_count_kernel<NB, SYM><<<nblk(n), 256, 0, st>>>(n, key, counts.p);
    check_launch();
    int h[NB + 1];
    PSB_CUDA(cudaMemcpyAsync(h, counts.p, sizeof(int) * (NB + 1), cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaStreamSynchronize(st));
    offs[0] = 0;
    for (int b = 0; b <= NB; ++b)
        offs[b + 1] = offs[b] + h[b];
    rows.alloc(std::max(1, offs[NB]));
    PSB_CUDA(cudaMemcpyAsync(d_offs.p, offs, sizeof(int) * (NB + 2), cudaMemcpyHostToDevice, st));
    bin_fill_kernel<NB, SYM><<<nblk(n), 256, 0, st>>>(n, key, d_offs.p, cursor.p, rows.p);
    check_launch();
    PSB_CUDA(cudaStreamSynchronize(st)); // offs is read by the copy above
    return h[NB];
}

} // namespace

// C = A * B by shared-memory hashing. Returns false (C untouched) when a row exceeds the largest bin: the caller then uses
// the sort-based product.
bool spgemm_hash(Ctx &c, Temp &tmp, const CsrDev &A, const CsrDev &B, int ncolsB, CsrDev &C)
{
    cudaStream_t st = c.stream;
    const int n = A.n;
    if (n == 0)
        return false;
    DevBuf<int> T, len, rows;
    T.alloc((size_t)n + 1, true);
    len.alloc((size_t)n + 1, true);
    int *d_max = (int *)c.counter.p + 3;
    PSB_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), st));
    product_count_kernel<<<nblk(n), 256, 0, st>>>(A.view(), B.rp.p, T.p, d_max);
    check_launch();
    const int maxT = d2h(c, d_max);
    if (maxT > h_sym_limit[kSymBins - 1])
        return false;
    // ---- symbolic
    int so[kSymBins + 2];
    bin_rows<kSymBins, true>(c, n, T.p, rows, so);
    launch_symbolic<4, 32, 256>(st, A, B, rows.p + so[0], so[1] - so[0], len.p);
    launch_symbolic<8, 128, 256>(st, A, B, rows.p + so[1], so[2] - so[1], len.p);
    launch_symbolic<16, 512, 256>(st, A, B, rows.p + so[2], so[3] - so[2], len.p);
    launch_symbolic<32, 2048, 256>(st, A, B, rows.p + so[3], so[4] - so[3], len.p);
    launch_symbolic<32, 8192, 128>(st, A, B, rows.p + so[4], so[5] - so[4], len.p);
    // ---- row pointer of C
    DevBuf<int> rp;
    rp.alloc((size_t)n + 1);
    exclusive_scan_int(c, tmp, len.p, rp.p, (long long)n + 1);
    const int nnzC = d2h(c, rp.p + n);
    int no[kNumBins + 2];
    const int beyond = bin_rows<kNumBins, false>(c, n, len.p, rows, no);
    if (beyond > 0)
        return false;
    C.n = n;
    C.ncols = ncolsB;
    C.nnz = nnzC;
    C.rp = std::move(rp);
    C.ci.alloc(std::max(1, nnzC), false, 64);
    C.va.alloc(std::max(1, nnzC), false, 64);
    // ---- numeric
    launch_numeric<4, 32, 256>(st, A, B, rows.p + no[0], no[1] - no[0], C);
    launch_numeric<8, 128, 256>(st, A, B, rows.p + no[1], no[2] - no[1], C);
    launch_numeric<16, 512, 256>(st, A, B, rows.p + no[2], no[3] - no[2], C);
    launch_numeric<32, 2048, 256>(st, A, B, rows.p + no[3], no[4] - no[3], C);
    PSB_CUDA(cudaStreamSynchronize(st));
    return true;
}

} // namespace psb
