// Smoothed-aggregation AMG on the device: setup (strength, parallel MIS-2 aggregation, smoothed
// prolongation, R = P^T, Galerkin RAP by expand-sort-compress) and the cycle (Chebyshev smoother
// fused into the SpMV, residual, restriction, prolongate-and-correct).
//
// Mirrors what the reference obtains from AMGCL 1.4.3 through polysolve::linear::AMGCL
// (reference src/polysolve/linear/AMGCL.cpp:32-65 parameters, :148-184 factorize = hierarchy build,
// :190-212 solve) -- see SURVEY.md Appendix A.3 for the algorithm this restates. The one deliberate
// difference: AMGCL's aggregation is a sequential greedy sweep; here it is a deterministic parallel
// MIS-2 (aggregates can also be imposed through psb200_debug_set_aggregates for parity tests).
#include "amg.hpp"
#include "amg_internal.hpp"
#include "push_epi.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <sstream>

namespace psb {

namespace {

inline int nblk(long long n, int t = 256) { return (int)std::max<long long>(1, (n + t - 1) / t); }

// ------------------------------------------------------------------------------ small utility kernels
__global__ void diag_kernel(CsrView A, double *__restrict__ diag)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n)
        return;
    double d = 0;
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
        if (A.ci[k] == (int)i)
            d = A.va[k];
    diag[i] = d;
}
__global__ void inv_kernel(long long n, const double *__restrict__ d, double *__restrict__ out, double scale)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = d[i] != 0.0 ? scale / d[i] : scale;
}
__global__ void fill_kernel(long long n, double *a, double v)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        a[i] = v;
}
// Gershgorin bound of D^-1 A: max_i sum_j |a_ij| / |a_ii|  (amgcl spectral_radius<true>, power_iters = 0).
// Positive doubles order like their bit patterns, so an integer atomicMax is exact and deterministic.
__global__ void gershgorin_kernel(CsrView A, unsigned long long *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0;
    if (i < A.n)
    {
        double d = 1;
        for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
        {
            s += fabs(A.va[k]);
            if (A.ci[k] == (int)i)
                d = A.va[k];
        }
        s *= fabs(1.0 / d);
    }
    for (int o = 16; o > 0; o >>= 1)
        s = fmax(s, __shfl_xor_sync(0xffffffffu, s, o));
    if ((threadIdx.x & 31) == 0 && s > 0)
        atomicMax(out, (unsigned long long)__double_as_longlong(s));
}
// counter-based splitmix64 -> U(-1,1): element i of the stream seeded with `seed`
__global__ void splitmix_kernel(long long n, unsigned long long seed, long long offset, double *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(offset + i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    out[i] = 2.0 * (double)(z >> 11) * (1.0 / 9007199254740992.0) - 1.0;
}


// ------------------------------------------------------------------------------ block mode (AMGCL_Block<B>, reference AMGCL.cpp:246-298)
// AMGCL's block value type (static_matrix<double,B,B>) is realised on the scalar CSR by its scalar expansion: the B rows of
// a node share one column list of full B x B blocks (made explicit by analyze_pattern, preserved by P, R and the Galerkin
// product), aggregation acts on nodes with Frobenius block norms, and every D^-1 of the scalar algorithm becomes the
// inverse of the B x B diagonal block, applied once to the matrix: Ahat = Dblk^-1 A is what the smoother, the
// prolongation smoothing and the spectral-radius estimate multiply with (and bhat = Dblk^-1 b per application).
template <int B>
__device__ __forceinline__ bool invert_small(const double *m, double *inv)
{
    if (B == 1)
    {
        inv[0] = 1.0 / m[0];
        return m[0] != 0.0;
    }
    if (B == 2)
    {
        const double det = m[0] * m[3] - m[1] * m[2];
        const double id = 1.0 / det;
        inv[0] = m[3] * id;
        inv[1] = -m[1] * id;
        inv[2] = -m[2] * id;
        inv[3] = m[0] * id;
        return det != 0.0 && det == det;
    }
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    const double id = 1.0 / det;
    inv[0] = c00 * id;
    inv[1] = (m[2] * m[7] - m[1] * m[8]) * id;
    inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    inv[3] = c01 * id;
    inv[4] = (m[0] * m[8] - m[2] * m[6]) * id;
    inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    inv[6] = c02 * id;
    inv[7] = (m[1] * m[6] - m[0] * m[7]) * id;
    inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    return det != 0.0 && det == det;
}
// dinvb[node] = inverse of the diagonal block; bad is raised for a missing / singular block or a broken block pattern
template <int B>
__global__ void block_diag_inv_kernel(CsrView A, double *__restrict__ dinvb, int *bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nb = A.n / B;
    if (i >= nb)
        return;
    double m[B * B];
    bool found = false;
    const int k0 = A.rp[B * i], len = A.rp[B * i + 1] - k0;
    for (int q = 0; q < len; q += B)
        if (A.ci[k0 + q] == B * i)
        {
            found = true;
            for (int r = 0; r < B; ++r)
                for (int c = 0; c < B; ++c)
                    m[r * B + c] = A.va[A.rp[B * i + r] + q + c];
        }
    for (int r = 1; r < B; ++r)
        if (A.rp[B * i + r + 1] - A.rp[B * i + r] != len)
            found = false;
    double inv[B * B];
    if (!found || !invert_small<B>(m, inv))
    {
        *bad = 1;
        for (int e = 0; e < B * B; ++e)
            inv[e] = (e % (B + 1) == 0) ? 1.0 : 0.0;
    }
    for (int e = 0; e < B * B; ++e)
        dinvb[(size_t)i * B * B + e] = inv[e];
}
// Ahat = Dblk^-1 A : entry k of row B i + c is sum_d Dinv[c][d] A[B i + d][same slot]
template <int B>
__global__ void block_scale_rows_kernel(CsrView A, const double *__restrict__ dinvb, double *__restrict__ out)
{
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= A.n)
        return;
    const int i = (int)(row / B), c = (int)(row % B);
    const int kb = A.rp[row], len = A.rp[row + 1] - kb;
    double w[B];
    int base[B];
    for (int d = 0; d < B; ++d)
    {
        w[d] = dinvb[(size_t)i * B * B + c * B + d];
        base[d] = A.rp[B * i + d];
    }
    for (int q = 0; q < len; ++q)
    {
        double s = 0;
        for (int d = 0; d < B; ++d)
            s += w[d] * A.va[base[d] + q];
        out[kb + q] = s;
    }
}
// out_node = Dinv_node in_node
template <int B>
__global__ void block_diag_apply_kernel(long long nb, const double *__restrict__ dinvb, const double *__restrict__ in, double *__restrict__ out,
                                        const int *done)
{
    if (done && *done)
        return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb)
        return;
    double v[B];
    for (int d = 0; d < B; ++d)
        v[d] = in[i * B + d];
    for (int c = 0; c < B; ++c)
    {
        double s = 0;
        for (int d = 0; d < B; ++d)
            s += dinvb[i * B * B + c * B + d] * v[d];
        out[i * B + c] = s;
    }
}
// node graph S: S.rp[i] = A.rp[B i] / B^2, S.ci = block columns, S.va = Frobenius norm of the block (math::norm of amgcl)
template <int B>
__global__ void block_norm_matrix_kernel(CsrView A, int *__restrict__ srp, int *__restrict__ sci, double *__restrict__ sva)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nb = A.n / B;
    if (i > nb)
        return;
    srp[i] = A.rp[B * i] / (B * B);
    if (i == nb)
        return;
    const int k0 = A.rp[B * i], len = (A.rp[B * i + 1] - k0) / B, o = k0 / (B * B);
    for (int q = 0; q < len; ++q)
    {
        double s = 0;
        for (int r = 0; r < B; ++r)
            for (int c = 0; c < B; ++c)
            {
                const double v = A.va[A.rp[B * i + r] + B * q + c];
                s += v * v;
            }
        sci[o + q] = A.ci[k0 + B * q] / B;
        sva[o + q] = sqrt(s);
    }
}
// amgcl spectral_radius<true> with power_iters = 0 on block values: max_i (sum_j ||A_ij||) ||D_i^-1||
template <int B>
__global__ void block_gershgorin_kernel(CsrView S, const double *__restrict__ dinvb, unsigned long long *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0;
    if (i < S.n)
    {
        for (int k = S.rp[i]; k < S.rp[i + 1]; ++k)
            s += S.va[k];
        double d = 0;
        for (int e = 0; e < B * B; ++e)
            d += dinvb[i * B * B + e] * dinvb[i * B * B + e];
        s *= sqrt(d);
    }
    for (int o = 16; o > 0; o >>= 1)
        s = fmax(s, __shfl_xor_sync(0xffffffffu, s, o));
    if ((threadIdx.x & 31) == 0 && s > 0)
        atomicMax(out, (unsigned long long)__double_as_longlong(s));
}
// amgcl spectral_radius power iteration, block values: radius = sum_nodes |<b1_node, b0_node>|
__global__ void block_radius_terms_kernel(int B, long long nb, const double *__restrict__ b1, const double *__restrict__ b0, double *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb)
        return;
    double s = 0;
    for (int c = 0; c < B; ++c)
        s += b1[i * B + c] * b0[i * B + c];
    out[i] = fabs(s);
}
// scalar aggregate ids from node aggregates: agg[B i + c] = B agg_node[i] + c (negative states are kept)
__global__ void block_expand_agg_kernel(int B, long long nb, const int *__restrict__ agg_node, int *__restrict__ agg)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb)
        return;
    const int g = agg_node[i];
    for (int c = 0; c < B; ++c)
        agg[i * B + c] = g >= 0 ? B * g + c : g;
}

// ------------------------------------------------------------------------------ strength + MIS-2 aggregation
// strong(i,j) <=> j != i and eps^2 a_ii a_jj < a_ij^2   (amgcl plain_aggregates; eps = 0 => every stored off-diagonal non-zero)
__device__ __forceinline__ bool is_strong(int i, int j, double aij, double eps2, const double *diag)
{
    return j != i && eps2 * diag[i] * diag[j] < aij * aij;
}
__device__ __forceinline__ unsigned long long node_prio(int i)
{
    unsigned long long z = (unsigned long long)(unsigned)i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (z & 0xFFFFFFFF00000000ull) | (unsigned)i; // unique, index breaks ties
}
enum : int
{
    ND_UNDECIDED = 0,
    ND_ROOT = 1,
    ND_EXCLUDED = 2,
    ND_REMOVED = 3
};
// nodes without any strong connection are "removed" (no row in P), as in amgcl
__global__ void agg_init_kernel(CsrView A, const double *__restrict__ diag, double eps2, int *__restrict__ state)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n)
        return;
    int st = ND_REMOVED;
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
        if (is_strong((int)i, A.ci[k], A.va[k], eps2, diag))
        {
            st = ND_UNDECIDED;
            break;
        }
    state[i] = st;
}
// m_out[i] = max over {i} U strongN(i) of m_in (m_in = priority of undecided nodes, else 0)
__global__ void agg_prio_kernel(long long n, const int *__restrict__ state, unsigned long long *__restrict__ m)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        m[i] = state[i] == ND_UNDECIDED ? node_prio((int)i) : 0ull;
}
__global__ void agg_max_kernel(CsrView A, const double *__restrict__ diag, double eps2, const unsigned long long *__restrict__ m_in,
                               unsigned long long *__restrict__ m_out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n)
        return;
    unsigned long long m = m_in[i];
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
    {
        const int j = A.ci[k];
        if (is_strong((int)i, j, A.va[k], eps2, diag))
            m = max(m, m_in[j]);
    }
    m_out[i] = m;
}
// undecided node whose priority is the maximum within distance 2 becomes a root
__global__ void agg_select_kernel(long long n, const unsigned long long *__restrict__ m2, int *__restrict__ state, int *changed)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    if (state[i] == ND_UNDECIDED && m2[i] == node_prio((int)i))
    {
        state[i] = ND_ROOT;
        *changed = 1;
    }
}
// flag_out[i] = flag_in[i] or any strong neighbour has flag_in; with flag_in = "is root" this marks distance-1, applied twice distance-2
__global__ void agg_spread_kernel(CsrView A, const double *__restrict__ diag, double eps2, const int *__restrict__ state,
                                  const unsigned char *__restrict__ f_in, unsigned char *__restrict__ f_out, int first)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n)
        return;
    unsigned char f = first ? (state[i] == ND_ROOT) : f_in[i];
    if (!f)
        for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
        {
            const int j = A.ci[k];
            if (is_strong((int)i, j, A.va[k], eps2, diag) && (first ? (state[j] == ND_ROOT) : f_in[j]))
            {
                f = 1;
                break;
            }
        }
    f_out[i] = f;
}
__global__ void agg_exclude_kernel(long long n, const unsigned char *__restrict__ near, int *__restrict__ state, int *remaining)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    if (state[i] == ND_UNDECIDED)
    {
        if (near[i])
            state[i] = ND_EXCLUDED;
        else
            *remaining = 1;
    }
}
__global__ void agg_rootflag_kernel(long long n, const int *__restrict__ state, int *__restrict__ flag)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        flag[i] = state[i] == ND_ROOT;
}
// pass A: roots take their scan id; distance-1 nodes join their (unique) root neighbour
__global__ void agg_assign1_kernel(CsrView A, const double *__restrict__ diag, double eps2, const int *__restrict__ state,
                                   const int *__restrict__ root_id, int *__restrict__ agg)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n)
        return;
    int a = -1;
    if (state[i] == ND_REMOVED)
        a = -2;
    else if (state[i] == ND_ROOT)
        a = root_id[i];
    else
        for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
        {
            const int j = A.ci[k];
            if (state[j] == ND_ROOT && is_strong((int)i, j, A.va[k], eps2, diag))
            {
                a = root_id[j];
                break;
            }
        }
    agg[i] = a;
}
// pass B: distance-2 nodes join the aggregate of their strongest already-assigned neighbour (ties: smallest column)
__global__ void agg_assign2_kernel(CsrView A, const double *__restrict__ diag, double eps2, const int *__restrict__ agg1, int *__restrict__ agg2)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n)
        return;
    int a = agg1[i];
    if (a == -1)
    {
        double best = -1;
        for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
        {
            const int j = A.ci[k];
            if (agg1[j] >= 0 && is_strong((int)i, j, A.va[k], eps2, diag) && fabs(A.va[k]) > best)
            {
                best = fabs(A.va[k]);
                a = agg1[j];
            }
        }
    }
    agg2[i] = a;
}

// ------------------------------------------------------------------------------ smoothed prolongation
// P = (I - omega D_f^-1 A_f) P_tent, P_tent(i, agg(i)) = 1 (amgcl smoothed_aggregation::transfer_operators).
// Each row is built in place inside the slot [A.rp[i], A.rp[i+1]) of scratch arrays (a row of P never has
// more entries than the row of A), merged by aggregate id and sorted by column; cnt[i] = entries.
// all_strong (block mode, eps_strong = 0): every stored entry takes part, so the B rows of a node keep one pattern.
__global__ void prolong_rows_kernel(CsrView A, const double *__restrict__ diagv, double eps2, const int *__restrict__ agg, double omega,
                                    int *__restrict__ scol, double *__restrict__ sval, int *__restrict__ cnt, bool all_strong)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n)
        return;
    const int kb = A.rp[i], ke = A.rp[i + 1];
    // filtered diagonal: diagonal plus weak connections
    double dia = 0;
    for (int k = kb; k < ke; ++k)
    {
        const int j = A.ci[k];
        if (j == (int)i || (!all_strong && !is_strong((int)i, j, A.va[k], eps2, diagv)))
            dia += A.va[k];
    }
    dia = -omega * (1.0 / dia);
    int m = 0;
    for (int k = kb; k < ke; ++k)
    {
        const int j = A.ci[k];
        if (j != (int)i && !all_strong && !is_strong((int)i, j, A.va[k], eps2, diagv))
            continue;
        const int g = agg[j];
        if (g < 0)
            continue;
        const double va = (j == (int)i) ? (1.0 - omega) : dia * A.va[k];
        // sorted insert / merge into scol[kb .. kb+m)
        int pos = 0;
        while (pos < m && scol[kb + pos] < g)
            ++pos;
        if (pos < m && scol[kb + pos] == g)
            sval[kb + pos] += va;
        else
        {
            for (int q = m; q > pos; --q)
            {
                scol[kb + q] = scol[kb + q - 1];
                sval[kb + q] = sval[kb + q - 1];
            }
            scol[kb + pos] = g;
            sval[kb + pos] = va;
            ++m;
        }
    }
    cnt[i] = m;
}
__global__ void compact_rows_kernel(long long n, const int *__restrict__ src_rp, const int *__restrict__ dst_rp, const int *__restrict__ scol,
                                    const double *__restrict__ sval, int *__restrict__ dcol, double *__restrict__ dval)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int s = src_rp[i], d = dst_rp[i], m = dst_rp[i + 1] - d;
    for (int q = 0; q < m; ++q)
    {
        dcol[d + q] = scol[s + q];
        dval[d + q] = sval[s + q];
    }
}

// ------------------------------------------------------------------------------ transpose helpers
__global__ void iota_kernel(long long n, int *a)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        a[i] = (int)i;
}
__global__ void expand_rows_kernel(int nrows, const int *__restrict__ rp, int *__restrict__ row_of)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows)
        return;
    for (int k = rp[i]; k < rp[i + 1]; ++k)
        row_of[k] = (int)i;
}
__global__ void lower_bound_rows_kernel(int nrows, long long nnz, const int *__restrict__ sorted_keys, int *__restrict__ rp)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nrows)
        return;
    long long lo = 0, hi = nnz;
    while (lo < hi)
    {
        const long long mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < (int)i)
            lo = mid + 1;
        else
            hi = mid;
    }
    rp[i] = (int)lo;
}
__global__ void gather_transpose_kernel(long long nnz, const int *__restrict__ perm, const int *__restrict__ row_of, const double *__restrict__ val,
                                        int *__restrict__ tcol, double *__restrict__ tval)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz)
        return;
    const int s = perm[k];
    tcol[k] = row_of[s];
    tval[k] = val[s];
}

// ------------------------------------------------------------------------------ SpGEMM by expand-sort-compress
__global__ void spgemm_count_kernel(CsrView A, const int *__restrict__ b_rp, long long *__restrict__ cnt)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n)
        return;
    long long c = 0;
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
    {
        const int a = A.ci[k];
        c += b_rp[a + 1] - b_rp[a];
    }
    cnt[i] = c;
}
__global__ void head_flags_kernel(long long n, const unsigned long long *__restrict__ keys, int *__restrict__ flag)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
// one thread per run head: sums the run in sorted (= generation) order -> deterministic
__global__ void compress_kernel(long long n, const unsigned long long *__restrict__ keys, const double *__restrict__ vals,
                                const int *__restrict__ flag, const int *__restrict__ pos, int *__restrict__ ccol, double *__restrict__ cval,
                                int *__restrict__ crow)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i])
        return;
    double s = vals[i];
    for (long long q = i + 1; q < n && !flag[q]; ++q)
        s += vals[q];
    const int p = pos[i];
    ccol[p] = (int)(keys[i] & 0xFFFFFFFFull);
    cval[p] = s;
    crow[p] = (int)(keys[i] >> 32);
}

// first Chebyshev step from a zero iterate: p = alpha M b; x = p
struct OpChebFirst
{
    static constexpr int NV = 0;
    const double *b, *dinv;
    double *p, *x;
    double alpha;
    __device__ __forceinline__ void prologue() {}
    template <int U>
    __device__ __forceinline__ void apply(long long j, long long stride, double (&)[1])
    {
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const long long i = j + u * stride;
            const double2 bv = ld2(b, i), dv = ld2(dinv, i);
            double2 o;
            o.x = alpha * dv.x * bv.x;
            o.y = alpha * dv.y * bv.y;
            st2(p, i, o);
            st2(x, i, o);
        }
    }
};
// A whole Chebyshev application (all `degree` steps) of a level that fits one SM, in ONE launch: the matrix is staged
// in shared memory once, the iterate and the Chebyshev direction live there too, steps are separated by __syncthreads.
// Same arithmetic as the step-per-launch path (EpiCheb / OpChebFirst) up to the summation order inside a row; it replaces 2 degree - 1 launches per visit of the coarsest level (the level is visited
// 2^(levels - 1) times per cycle and each of its SpMVs is pure launch latency).
constexpr int kSmallRows = 1024, kSmallNnz = 12288, kSmallDegree = 64;
struct ChebCoef
{
    double alpha[kSmallDegree], beta[kSmallDegree];
};
__global__ void __launch_bounds__(1024) cheb_sweep_small_kernel(CsrView A, const double *__restrict__ rhs, const double *__restrict__ dinv,
                                                                double *__restrict__ x, double *__restrict__ cp, ChebCoef cf, int degree,
                                                                int x_is_zero, const int *done)
{
    if (done && *done)
        return;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    double *sva = reinterpret_cast<double *>(sm_raw);
    double *sx = sva + kSmallNnz;
    double *sxn = sx + kSmallRows;
    int *sci = reinterpret_cast<int *>(sxn + kSmallRows);
    const int n = A.n, nnz = A.rp[n];
    for (int k = threadIdx.x; k < nnz; k += blockDim.x)
    {
        sva[k] = A.va[k];
        sci[k] = A.ci[k];
    }
    // LPR lanes per row (a power of two <= 32 with n * LPR <= 1024): the lanes of a row sit in one warp
    int lpr = 1;
    while (lpr < 32 && n * (lpr * 2) <= (int)blockDim.x)
        lpr *= 2;
    const int row = threadIdx.x / lpr, lane = threadIdx.x % lpr;
    const bool live = row < n;
    int kb = 0, ke = 0;
    double b = 0, d = 0, p = 0;
    if (live)
    {
        kb = A.rp[row];
        ke = A.rp[row + 1];
        b = rhs[row];
        d = dinv[row];
        if (lane == 0)
            sx[row] = x_is_zero ? 0.0 : x[row];
    }
    __syncthreads();
    for (int k = 0; k < degree; ++k)
    {
        const bool first = k == 0 && x_is_zero;
        double s = 0;
        if (live && !first)
            for (int q = kb + lane; q < ke; q += lpr)
                s += sva[q] * sx[sci[q]];
        for (int o = lpr >> 1; o > 0; o >>= 1)
            s += __shfl_xor_sync(0xffffffffu, s, o);
        if (live && lane == 0)
        {
            if (first)
                p = cf.alpha[0] * d * b; // OpChebFirst
            else
            {
                const double res = d * (b - s); // EpiCheb
                p = cf.alpha[k] * res + (cf.beta[k] != 0.0 ? cf.beta[k] * p : 0.0);
            }
            sxn[row] = sx[row] + p;
        }
        __syncthreads();
        double *t = sx;
        sx = sxn;
        sxn = t;
    }
    if (live && lane == 0)
    {
        x[row] = sx[row];
        cp[row] = p;
    }
}

// relaxation from a zero iterate with diagonal weights: x = w b
struct OpDiagFirst
{
    static constexpr int NV = 0;
    const double *b, *w;
    double *x;
    __device__ __forceinline__ void prologue() {}
    template <int U>
    __device__ __forceinline__ void apply(long long j, long long stride, double (&)[1])
    {
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const long long i = j + u * stride;
            const double2 bv = ld2(b, i), dv = ld2(w, i);
            double2 o;
            o.x = dv.x * bv.x;
            o.y = dv.y * bv.y;
            st2(x, i, o);
        }
    }
};

} // namespace

void exclusive_scan_int(Ctx &c, Temp &tmp, const int *in, int *out, long long n)
{
    size_t bytes = 0;
    PSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, c.stream));
    PSB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(bytes), bytes, in, out, n, c.stream));
}
void exclusive_scan_ll(Ctx &c, Temp &tmp, const long long *in, long long *out, long long n)
{
    size_t bytes = 0;
    PSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, c.stream));
    PSB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(bytes), bytes, in, out, n, c.stream));
}


// T = M^T for a CSR matrix with ncols columns. Stable sort by column keeps the fine-row order inside
// every row of the transpose (the same idiom as the CSC -> CSR ingest).
void transpose(Ctx &c, Temp &tmp, const CsrDev &M, CsrDev &T)
{
    T.n = M.ncols;
    T.ncols = M.n;
    T.nnz = M.nnz;
    T.rp.alloc((size_t)T.n + 1);
    T.ci.alloc(std::max<long long>(1, T.nnz), false, 64);
    T.va.alloc(std::max<long long>(1, T.nnz), false, 64);
    if (M.nnz == 0)
    {
        PSB_CUDA(cudaMemsetAsync(T.rp.p, 0, sizeof(int) * ((size_t)T.n + 1), c.stream));
        return;
    }
    DevBuf<int> iota, perm, sorted, row_of;
    iota.alloc(M.nnz);
    perm.alloc(M.nnz);
    sorted.alloc(M.nnz);
    row_of.alloc(M.nnz);
    iota_kernel<<<nblk(M.nnz), 256, 0, c.stream>>>(M.nnz, iota.p);
    expand_rows_kernel<<<nblk(M.n), 256, 0, c.stream>>>(M.n, M.rp.p, row_of.p);
    int end_bit = 1;
    while ((1ll << end_bit) < M.ncols && end_bit < 31)
        ++end_bit;
    size_t bytes = 0;
    PSB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, M.ci.p, sorted.p, iota.p, perm.p, (int)M.nnz, 0, end_bit, c.stream));
    PSB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(bytes), bytes, M.ci.p, sorted.p, iota.p, perm.p, (int)M.nnz, 0, end_bit, c.stream));
    lower_bound_rows_kernel<<<nblk((long long)T.n + 1), 256, 0, c.stream>>>(T.n, M.nnz, sorted.p, T.rp.p);
    gather_transpose_kernel<<<nblk(M.nnz), 256, 0, c.stream>>>(M.nnz, perm.p, row_of.p, M.va.p, T.ci.p, T.va.p);
    check_launch();
    PSB_CUDA(cudaStreamSynchronize(c.stream));
}

// first row r >= from whose product offset reaches `limit` (host-side chunking of the expansion)
__global__ void chunk_end_kernel(int n, const long long *__restrict__ off, int from, long long limit, int *out)
{
    if (blockIdx.x || threadIdx.x)
        return;
    int lo = from + 1, hi = n; // the chunk always takes at least one row
    while (lo < hi)
    {
        const int mid = (lo + hi + 1) >> 1;
        if (off[mid] - off[from] <= limit)
            lo = mid;
        else
            hi = mid - 1;
    }
    *out = lo;
}
__global__ void spgemm_expand_rows_kernel(CsrView A, CsrView B, const long long *__restrict__ off, int r0, int r1, unsigned long long *__restrict__ keys,
                                          double *__restrict__ vals)
{
    const long long i = (long long)r0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r1)
        return;
    long long o = off[i] - off[r0];
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
    {
        const int a = A.ci[k];
        const double va = A.va[k];
        for (int q = B.rp[a]; q < B.rp[a + 1]; ++q)
        {
            keys[o] = ((unsigned long long)(unsigned)i << 32) | (unsigned)B.ci[q];
            vals[o] = va * B.va[q];
            ++o;
        }
    }
}

// C = A * B (B has ncolsB columns). Expand all products, stable radix sort by (row, col), sum runs (deterministic: the
// products of an entry are added in generation order). Rows are processed in chunks of at most 2^29 products, so the
// temporaries stay bounded (C4: 5.4e9 products of A P on one GPU) whatever the size of the product.
void spgemm(Ctx &c, Temp &tmp, const CsrDev &A, const CsrDev &B, int ncolsB, CsrDev &C)
{
    if (spgemm_use_hash() && spgemm_hash(c, tmp, A, B, ncolsB, C))
        return; // the hand-written shared-memory hash product (spgemm.cu); what follows is the fallback for huge rows
    spgemm_sort(c, tmp, A, B, ncolsB, C);
}

void spgemm_sort(Ctx &c, Temp &tmp, const CsrDev &A, const CsrDev &B, int ncolsB, CsrDev &C)
{
    cudaStream_t st = c.stream;
    C.n = A.n;
    C.ncols = ncolsB;
    C.rp.alloc((size_t)C.n + 1);
    DevBuf<long long> cnt, off;
    cnt.alloc((size_t)A.n + 1, true);
    off.alloc((size_t)A.n + 1);
    if (A.n)
        spgemm_count_kernel<<<nblk(A.n), 256, 0, st>>>(A.view(), B.rp.p, cnt.p);
    exclusive_scan_ll(c, tmp, cnt.p, off.p, (long long)A.n + 1);
    const long long T = d2h(c, off.p + A.n);
    if (T == 0)
    {
        C.nnz = 0;
        C.ci.alloc(1, false, 64);
        C.va.alloc(1, false, 64);
        PSB_CUDA(cudaMemsetAsync(C.rp.p, 0, sizeof(int) * ((size_t)C.n + 1), st));
        return;
    }
    constexpr long long kChunk = 1ll << 29;
    int rbits = 1;
    while ((1ll << rbits) < A.n && rbits < 31)
        ++rbits;
    struct Piece
    {
        DevBuf<int> ci, row;
        DevBuf<double> va;
        int nnz = 0;
    };
    std::vector<Piece> pieces;
    long long total = 0;
    int *d_end = (int *)c.counter.p + 3;
    for (int r0 = 0; r0 < A.n;)
    {
        chunk_end_kernel<<<1, 1, 0, st>>>(A.n, off.p, r0, kChunk, d_end);
        const int r1 = d2h(c, d_end);
        long long o01[2];
        PSB_CUDA(cudaMemcpyAsync(&o01[0], off.p + r0, sizeof(long long), cudaMemcpyDeviceToHost, st));
        PSB_CUDA(cudaMemcpyAsync(&o01[1], off.p + r1, sizeof(long long), cudaMemcpyDeviceToHost, st));
        PSB_CUDA(cudaStreamSynchronize(st));
        const long long Tc = o01[1] - o01[0];
        if (Tc > 0x7fffffffLL)
            throw std::runtime_error("psb200 amg: one row of the Galerkin product has more than 2^31 intermediate products");
        pieces.emplace_back();
        Piece &pc = pieces.back();
        if (Tc > 0)
        {
            DevBuf<unsigned long long> keys, keys2;
            DevBuf<double> vals, vals2;
            keys.alloc(Tc);
            keys2.alloc(Tc);
            vals.alloc(Tc);
            vals2.alloc(Tc);
            spgemm_expand_rows_kernel<<<nblk(r1 - r0), 256, 0, st>>>(A.view(), B.view(), off.p, r0, r1, keys.p, vals.p);
            check_launch();
            size_t bytes = 0;
            PSB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.p, keys2.p, vals.p, vals2.p, (int)Tc, 0, 32 + rbits, st));
            PSB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(bytes), bytes, keys.p, keys2.p, vals.p, vals2.p, (int)Tc, 0, 32 + rbits, st));
            keys.release();
            vals.release();
            DevBuf<int> flag, pos;
            flag.alloc((size_t)Tc + 1, true);
            pos.alloc((size_t)Tc + 1);
            head_flags_kernel<<<nblk(Tc), 256, 0, st>>>(Tc, keys2.p, flag.p);
            exclusive_scan_int(c, tmp, flag.p, pos.p, Tc + 1);
            pc.nnz = d2h(c, pos.p + Tc);
            pc.ci.alloc(std::max(1, pc.nnz), false, 64);
            pc.va.alloc(std::max(1, pc.nnz), false, 64);
            pc.row.alloc(std::max(1, pc.nnz));
            compress_kernel<<<nblk(Tc), 256, 0, st>>>(Tc, keys2.p, vals2.p, flag.p, pos.p, pc.ci.p, pc.va.p, pc.row.p);
            check_launch();
            PSB_CUDA(cudaStreamSynchronize(st));
        }
        total += pc.nnz;
        r0 = r1;
    }
    if (total > 0x7fffffffLL - 1024)
        throw std::runtime_error("psb200 amg: a coarse matrix exceeds the int32 index range on one GPU");
    C.nnz = total;
    if (pieces.size() == 1 && total > 0)
    {
        // the common case: one chunk, its buffers become the result
        C.ci = std::move(pieces[0].ci);
        C.va = std::move(pieces[0].va);
        lower_bound_rows_kernel<<<nblk((long long)C.n + 1), 256, 0, st>>>(C.n, total, pieces[0].row.p, C.rp.p);
        check_launch();
        PSB_CUDA(cudaStreamSynchronize(st));
        return;
    }
    C.ci.alloc(std::max<long long>(1, total), false, 64);
    C.va.alloc(std::max<long long>(1, total), false, 64);
    DevBuf<int> crow;
    crow.alloc(std::max<long long>(1, total));
    long long o = 0;
    for (Piece &pc : pieces)
    {
        if (pc.nnz)
        {
            PSB_CUDA(cudaMemcpyAsync(C.ci.p + o, pc.ci.p, sizeof(int) * (size_t)pc.nnz, cudaMemcpyDeviceToDevice, st));
            PSB_CUDA(cudaMemcpyAsync(C.va.p + o, pc.va.p, sizeof(double) * (size_t)pc.nnz, cudaMemcpyDeviceToDevice, st));
            PSB_CUDA(cudaMemcpyAsync(crow.p + o, pc.row.p, sizeof(int) * (size_t)pc.nnz, cudaMemcpyDeviceToDevice, st));
        }
        o += pc.nnz;
    }
    lower_bound_rows_kernel<<<nblk((long long)C.n + 1), 256, 0, st>>>(C.n, total, crow.p, C.rp.p);
    check_launch();
    PSB_CUDA(cudaStreamSynchronize(st));
}

// ====================================================================================== level

AmgHierarchy::AmgHierarchy(Ctx &ctx, const AmgParams &prm) : ctx_(ctx), prm_(prm) {}
AmgHierarchy::~AmgHierarchy() {}
int AmgHierarchy::num_levels() const { return (int)levels_.size(); }

const CsrDev &AmgHierarchy::matrix(int level, int which) const
{
    const AmgLevel &L = *levels_.at(level);
    return which == 0 ? *L.A : which == 1 ? L.P : L.R;
}
int AmgHierarchy::matrix_cols(int level, int which) const { return matrix(level, which).ncols; }
const int *AmgHierarchy::aggregates(int level, int *n_agg) const
{
    const AmgLevel &L = *levels_.at(level);
    if (n_agg)
        *n_agg = L.n_agg;
    return L.agg.p;
}

static inline double bits_to_double(unsigned long long b)
{
    double d;
    std::memcpy(&d, &b, 8);
    return d;
}

static double gershgorin(Ctx &c, const CsrDev &A)
{
    unsigned long long *d_max = (unsigned long long *)c.partials.p;
    PSB_CUDA(cudaMemsetAsync(d_max, 0, 8, c.stream));
    gershgorin_kernel<<<nblk(A.n), 256, 0, c.stream>>>(A.view(), d_max);
    check_launch();
    return bits_to_double(d2h(c, d_max));
}

// node graph S of a block matrix: Frobenius norms of the B x B blocks (amgcl math::norm), all columns kept
static void block_norm_graph(Ctx &c, int B, const CsrDev &A, CsrDev &S)
{
    cudaStream_t st = c.stream;
    const long long nb = A.n / B;
    S.n = (int)nb;
    S.ncols = A.ncols / B;
    S.nnz = A.nnz / (B * B);
    S.rp.alloc((size_t)nb + 1);
    S.ci.alloc(std::max<long long>(1, S.nnz), false, 64);
    S.va.alloc(std::max<long long>(1, S.nnz), false, 64);
    if (B == 2)
        block_norm_matrix_kernel<2><<<nblk(nb + 1), 256, 0, st>>>(A.view(), S.rp.p, S.ci.p, S.va.p);
    else
        block_norm_matrix_kernel<3><<<nblk(nb + 1), 256, 0, st>>>(A.view(), S.rp.p, S.ci.p, S.va.p);
    check_launch();
}

// dinvb = inverted diagonal blocks of A, Ahat = Dblk^-1 A (same pattern)
static void block_scaled_matrix(Ctx &c, int B, const CsrDev &A, DevBuf<double> &dinvb, CsrDev &Ahat)
{
    cudaStream_t st = c.stream;
    if (A.n % B)
        throw std::runtime_error("psb200 amg: level size is not a multiple of the block size");
    const long long nb = A.n / B;
    dinvb.alloc((size_t)nb * B * B);
    int *d_bad = (int *)c.counter.p + 3;
    PSB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    if (B == 2)
        block_diag_inv_kernel<2><<<nblk(nb), 256, 0, st>>>(A.view(), dinvb.p, d_bad);
    else
        block_diag_inv_kernel<3><<<nblk(nb), 256, 0, st>>>(A.view(), dinvb.p, d_bad);
    check_launch();
    if (d2h(c, d_bad))
        throw std::runtime_error("psb200 amg: singular or missing diagonal block (block_size " + std::to_string(B) + ")");
    Ahat.n = A.n;
    Ahat.ncols = A.ncols;
    Ahat.nnz = A.nnz;
    Ahat.nl = A.nl;
    Ahat.halo_mask = A.halo_mask;
    Ahat.rp.alloc((size_t)A.n + 1);
    Ahat.ci.alloc(std::max<long long>(1, A.nnz), false, 64);
    Ahat.va.alloc(std::max<long long>(1, A.nnz), false, 64);
    PSB_CUDA(cudaMemcpyAsync(Ahat.rp.p, A.rp.p, sizeof(int) * ((size_t)A.n + 1), cudaMemcpyDeviceToDevice, st));
    if (A.nnz)
        PSB_CUDA(cudaMemcpyAsync(Ahat.ci.p, A.ci.p, sizeof(int) * (size_t)A.nnz, cudaMemcpyDeviceToDevice, st));
    if (B == 2)
        block_scale_rows_kernel<2><<<nblk(A.n), 256, 0, st>>>(A.view(), dinvb.p, Ahat.va.p);
    else
        block_scale_rows_kernel<3><<<nblk(A.n), 256, 0, st>>>(A.view(), dinvb.p, Ahat.va.p);
    check_launch();
    Ahat.kind = A.kind;
    Ahat.lpr = A.lpr;
    Ahat.narrow = A.narrow;
    Ahat.block = A.block;
    Ahat.use_bsr = A.use_bsr;
    Ahat.refresh_bsr(st);
}

void setup_relaxation(Ctx &c, const AmgParams &prm, AmgLevel &L, int seed_index, const SetupHooks *hooks)
{
    cudaStream_t st = c.stream;
    const CsrDev &A = *L.A;
    L.n = A.n;
    L.n_pad = ((long long)A.n + 3) & ~3ll;
    const size_t np = (size_t)std::max<long long>(L.n_pad, 4);
    L.dinv.alloc(np, true);
    L.f.alloc(np, true);
    L.u.alloc(np, true);
    L.ualt.alloc(np, true);
    L.t.alloc(np, true);
    L.cp.alloc(np, true);
    DevBuf<double> diag;
    diag.alloc(np, true);
    const int B = std::max(1, prm.block_size);
    L.Asm = &A;
    if (B > 1)
    {
        // Dblk^-1 once, Ahat = Dblk^-1 A; from here on the scalar code runs on Ahat with a unit diagonal scaling
        block_scaled_matrix(c, B, A, L.dinvb, L.Ahat);
        L.bh.alloc(np, true);
        L.Asm = &L.Ahat;
        if (A.n)
        {
            fill_kernel<<<nblk(A.n), 256, 0, st>>>(A.n, L.dinv.p, 1.0);
            fill_kernel<<<nblk(A.n), 256, 0, st>>>(A.n, diag.p, 1.0);
        }
    }
    else if (A.n)
    {
        diag_kernel<<<nblk(A.n), 256, 0, st>>>(A.view(), diag.p);
        inv_kernel<<<nblk(A.n), 256, 0, st>>>(A.n, diag.p, L.dinv.p, 1.0);
    }
    check_launch();
    const CsrDev &As = *L.Asm;
    if (prm.relax_type == "chebyshev")
    {
        // amgcl relaxation::chebyshev ctor: rho(D^-1 A) by power iteration (or Gershgorin), hi = higher rho, lo = lower rho
        if (!prm.scale)
            throw std::runtime_error("psb200 amg: chebyshev with scale=false is not supported");
        double rho;
        if (prm.power_iters <= 0)
        {
            rho = A.n ? gershgorin(c, As) : 0.0;
            if (hooks && hooks->allmax)
                rho = hooks->allmax(rho);
        }
        else
        {
            // b0 = the same counter-based splitmix64 stream the CPU restatement starts from (indexed by the GLOBAL row)
            DevBuf<double> b0, b1, scal;
            b0.alloc(np, true);
            b1.alloc(np, true);
            scal.alloc(4, true);
            const long long row0 = hooks ? hooks->row0 : 0;
            if (A.n)
                splitmix_kernel<<<nblk(A.n), 256, 0, st>>>(A.n, 1000ull + (unsigned long long)(seed_index + 1), row0, b0.p);
            check_launch();
            launch_vec(c, "amg_setup", np, OpDot{b0.p, b0.p}, FinStore{scal.p, 1});
            launch_vec(c, "amg_setup", np, OpScale{b0.p, b0.p, scal.p, 0.0}, FinNone{});
            for (int it = 0; it < prm.power_iters; ++it)
            {
                if (hooks && hooks->push)
                    hooks->push(b0.p);
                launch_spmv(c, "amg_setup", As, b0.p, EpiPower{b1.p, b0.p, L.dinv.p}, FinStore{scal.p, 2});
                if (it + 1 < prm.power_iters)
                    launch_vec(c, "amg_setup", np, OpScale{b0.p, b1.p, scal.p, 0.0}, FinNone{});
            }
            if (B > 1)
            {
                // the block algorithm takes |.| of the per-node inner product, not of every scalar product
                DevBuf<double> terms;
                terms.alloc(np, true);
                if (A.n)
                    block_radius_terms_kernel<<<nblk(A.n / B), 256, 0, st>>>(B, A.n / B, b1.p, b0.p, terms.p);
                check_launch();
                launch_vec(c, "amg_setup", np, OpDot{terms.p, L.dinv.p}, FinStore{scal.p + 1, 1});
            }
            double h[2];
            PSB_CUDA(cudaMemcpyAsync(h, scal.p, 16, cudaMemcpyDeviceToHost, st));
            PSB_CUDA(cudaStreamSynchronize(st));
            rho = h[1] < 0 ? 2.0 : h[1];
        }
        L.rho = rho;
        const double lo = rho * prm.lower, hi = rho * prm.higher;
        L.cheb_d = 0.5 * (hi + lo);
        L.cheb_c = 0.5 * (hi - lo);
        L.alpha.assign(prm.degree, 0.0);
        L.beta.assign(prm.degree, 0.0);
        double alpha = 0, beta = 0;
        const double d = L.cheb_d, cc = L.cheb_c;
        for (int k = 0; k < prm.degree; ++k)
        {
            if (k == 0)
            {
                alpha = 1.0 / d;
                beta = 0;
            }
            else if (k == 1)
            {
                alpha = 2 * d * (1.0 / (2 * d * d - cc * cc));
                beta = alpha * d - 1;
            }
            else
            {
                alpha = 1.0 / (d - 0.25 * alpha * cc * cc);
                beta = alpha * d - 1;
            }
            L.alpha[k] = alpha;
            L.beta[k] = beta;
        }
    }
    else
    {
        // damped Jacobi: w = damping / a_ii
        L.w.alloc(np, true);
        if (A.n)
            inv_kernel<<<nblk(A.n), 256, 0, st>>>(A.n, diag.p, L.w.p, prm.damping);
        check_launch();
    }
    PSB_CUDA(cudaStreamSynchronize(st));
}

// amgcl spectral_radius<true> with power_iters = 0 (the bound omega of the prolongation smoothing is made from)
double gershgorin_rho(Ctx &c, const AmgParams &prm, const AmgLevel &L)
{
    const int B = std::max(1, prm.block_size);
    const CsrDev &A = *L.A;
    if (A.n == 0)
        return 0.0;
    if (B == 1)
        return gershgorin(c, A);
    CsrDev S;
    block_norm_graph(c, B, A, S);
    unsigned long long *d_max = (unsigned long long *)c.partials.p;
    PSB_CUDA(cudaMemsetAsync(d_max, 0, 8, c.stream));
    if (B == 2)
        block_gershgorin_kernel<2><<<nblk(S.n), 256, 0, c.stream>>>(S.view(), L.dinvb.p, d_max);
    else
        block_gershgorin_kernel<3><<<nblk(S.n), 256, 0, c.stream>>>(S.view(), L.dinvb.p, d_max);
    check_launch();
    return bits_to_double(d2h(c, d_max));
}

// Deterministic parallel MIS-2 aggregation. Returns the number of aggregates; agg[i] in [0, n_agg) or -2 (removed).
static int aggregate_mis2(Ctx &c, Temp &tmp, const CsrDev &A, const double *diag, double eps_strong, DevBuf<int> &agg, int &rounds)
{
    cudaStream_t st = c.stream;
    const long long n = A.n;
    const double eps2 = eps_strong * eps_strong;
    DevBuf<int> state, flag, root_id, agg1;
    DevBuf<unsigned long long> m0, m1;
    DevBuf<unsigned char> f0, f1;
    state.alloc(n);
    flag.alloc(n + 1, true);
    root_id.alloc(n + 1);
    agg1.alloc(n);
    agg.alloc(n);
    rounds = 0;
    if (n == 0)
        return 0;
    m0.alloc(n);
    m1.alloc(n);
    f0.alloc(n);
    f1.alloc(n);
    int *d_flags = (int *)c.counter.p + 2; // [changed, remaining]
    agg_init_kernel<<<nblk(n), 256, 0, st>>>(A.view(), diag, eps2, state.p);
    for (;;)
    {
        ++rounds;
        PSB_CUDA(cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), st));
        agg_prio_kernel<<<nblk(n), 256, 0, st>>>(n, state.p, m0.p);
        agg_max_kernel<<<nblk(n), 256, 0, st>>>(A.view(), diag, eps2, m0.p, m1.p);
        agg_max_kernel<<<nblk(n), 256, 0, st>>>(A.view(), diag, eps2, m1.p, m0.p);
        agg_select_kernel<<<nblk(n), 256, 0, st>>>(n, m0.p, state.p, d_flags);
        agg_spread_kernel<<<nblk(n), 256, 0, st>>>(A.view(), diag, eps2, state.p, nullptr, f0.p, 1);
        agg_spread_kernel<<<nblk(n), 256, 0, st>>>(A.view(), diag, eps2, state.p, f0.p, f1.p, 0);
        agg_exclude_kernel<<<nblk(n), 256, 0, st>>>(n, f1.p, state.p, d_flags + 1);
        check_launch();
        int h[2];
        PSB_CUDA(cudaMemcpyAsync(h, d_flags, sizeof(h), cudaMemcpyDeviceToHost, st));
        PSB_CUDA(cudaStreamSynchronize(st));
        if (!h[1])
            break;
        if (rounds > 200)
            throw std::runtime_error("psb200 amg: MIS-2 aggregation did not terminate");
    }
    agg_rootflag_kernel<<<nblk(n), 256, 0, st>>>(n, state.p, flag.p);
    exclusive_scan_int(c, tmp, flag.p, root_id.p, n + 1);
    const int n_agg = d2h(c, root_id.p + n);
    agg_assign1_kernel<<<nblk(n), 256, 0, st>>>(A.view(), diag, eps2, state.p, root_id.p, agg1.p);
    agg_assign2_kernel<<<nblk(n), 256, 0, st>>>(A.view(), diag, eps2, agg1.p, agg.p);
    check_launch();
    PSB_CUDA(cudaStreamSynchronize(st));
    return n_agg;
}

// PSB200_SMALL_SWEEP=off in the environment disables the one-launch Chebyshev sweep of tiny levels (A/B runs)
static bool small_sweep_enabled()
{
    static bool on = std::getenv("PSB200_SMALL_SWEEP") == nullptr || std::string(std::getenv("PSB200_SMALL_SWEEP")) != "off";
    return on;
}

bool &spgemm_use_hash()
{
    static bool on = std::getenv("PSB200_SPGEMM") == nullptr || std::string(std::getenv("PSB200_SPGEMM")) != "sort";
    return on;
}

double wall_ms(cudaStream_t st)
{
    cudaStreamSynchronize(st);
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// amgcl plain_aggregates on the square matrix Asq (block mode: on the node graph with Frobenius block norms)
void build_aggregates(Ctx &c, Temp &tmp, const AmgParams &prm, const CsrDev &Asq, double eps_strong, const std::vector<int> *imposed, AmgLevel &L)
{
    cudaStream_t st = c.stream;
    const long long n = Asq.n;
    const int B = std::max(1, prm.block_size);
    if (B > 1 && eps_strong != 0.0)
        throw std::runtime_error("psb200 amg: block_size > 1 supports eps_strong = 0 only (polysolve's default, AMGCL.cpp:52)");
    if (imposed && !imposed->empty())
    {
        // imposed ids are per node in block mode, per row otherwise
        if ((long long)imposed->size() != n / B)
            throw std::invalid_argument("psb200 amg: imposed aggregate array has the wrong length");
        DevBuf<int> &dst = B > 1 ? L.agg_node : L.agg;
        dst.alloc(n / B);
        PSB_CUDA(cudaMemcpyAsync(dst.p, imposed->data(), sizeof(int) * (n / B), cudaMemcpyHostToDevice, st));
        PSB_CUDA(cudaStreamSynchronize(st));
        int mx = -1;
        for (int a : *imposed)
            mx = std::max(mx, a);
        L.n_agg = mx + 1;
    }
    else if (B > 1)
    {
        CsrDev S; // node graph with Frobenius block norms
        DevBuf<double> sdiag;
        block_norm_graph(c, B, Asq, S);
        sdiag.alloc((size_t)std::max<long long>(1, S.n));
        if (S.n)
            diag_kernel<<<nblk(S.n), 256, 0, st>>>(S.view(), sdiag.p);
        check_launch();
        L.n_agg = aggregate_mis2(c, tmp, S, sdiag.p, eps_strong, L.agg_node, L.mis_rounds);
    }
    else
    {
        DevBuf<double> diag;
        diag.alloc((size_t)std::max<long long>(1, n));
        if (n)
            diag_kernel<<<nblk(n), 256, 0, st>>>(Asq.view(), diag.p);
        check_launch();
        L.n_agg = aggregate_mis2(c, tmp, Asq, diag.p, eps_strong, L.agg, L.mis_rounds);
    }
    if (B > 1)
    {
        L.agg.alloc(n);
        if (n)
            block_expand_agg_kernel<<<nblk(n / B), 256, 0, st>>>(B, n / B, L.agg_node.p, L.agg.p);
        check_launch();
        L.n_agg *= B;
    }
    PSB_CUDA(cudaStreamSynchronize(st));
}

// amgcl smoothed_aggregation::transfer_operators: P = (I - omega D_f^-1 A_f) P_tent on the square matrix Asq
void build_prolongation(Ctx &c, Temp &tmp, const AmgParams &prm, const CsrDev &Asq, const CsrDev *Ahat, double eps_strong, double omega, AmgLevel &L)
{
    cudaStream_t st = c.stream;
    const long long n = Asq.n;
    const int B = std::max(1, prm.block_size);
    const double eps2 = eps_strong * eps_strong;
    CsrDev Ahat_own;
    DevBuf<double> dinvb_own;
    const CsrDev *Asm = &Asq;
    if (B > 1)
    {
        if (!Ahat)
        {
            block_scaled_matrix(c, B, Asq, dinvb_own, Ahat_own);
            Ahat = &Ahat_own;
        }
        Asm = Ahat;
    }
    DevBuf<double> diag;
    diag.alloc((size_t)std::max<long long>(1, n));
    if (n)
    {
        if (B > 1)
            fill_kernel<<<nblk(n), 256, 0, st>>>(n, diag.p, 1.0);
        else
            diag_kernel<<<nblk(n), 256, 0, st>>>(Asq.view(), diag.p);
    }
    check_launch();
    DevBuf<int> scol, cnt;
    DevBuf<double> sval;
    scol.alloc(std::max<long long>(1, Asm->nnz));
    sval.alloc(std::max<long long>(1, Asm->nnz));
    cnt.alloc(n + 1, true);
    if (n)
        prolong_rows_kernel<<<nblk(n), 256, 0, st>>>(Asm->view(), diag.p, eps2, L.agg.p, omega, scol.p, sval.p, cnt.p, B > 1);
    check_launch();
    L.P.n = (int)n;
    L.P.ncols = L.n_agg;
    L.P.rp.alloc(n + 1);
    exclusive_scan_int(c, tmp, cnt.p, L.P.rp.p, n + 1);
    L.P.nnz = d2h(c, L.P.rp.p + n);
    L.P.ci.alloc(std::max<long long>(1, L.P.nnz), false, 64);
    L.P.va.alloc(std::max<long long>(1, L.P.nnz), false, 64);
    if (n)
        compact_rows_kernel<<<nblk(n), 256, 0, st>>>(n, Asm->rp.p, L.P.rp.p, scol.p, sval.p, L.P.ci.p, L.P.va.p);
    check_launch();
    PSB_CUDA(cudaStreamSynchronize(st));
}

// amgcl amg::do_init (SURVEY A.3 "Hierarchy build")
void AmgHierarchy::setup(const CsrDev &A0, const std::vector<std::vector<int>> &imposed, int level_base)
{
    cudaStream_t st = ctx_.stream;
    levels_.clear();
    A0_ = &A0;
    level_base_ = level_base;
    Temp tmp;
    auto cur = std::make_unique<AmgLevel>();
    cur->A = &A0;
    double eps_strong = prm_.eps_strong;
    for (int i = 0; i < level_base; ++i)
        eps_strong *= 0.5; // amgcl halves eps_strong after every level
    while (cur && cur->A->n > prm_.coarse_enough)
    {
        levels_.push_back(std::move(cur));
        AmgLevel &L = *levels_.back();
        const CsrDev &A = *L.A;
        const int li = (int)levels_.size() - 1;
        double tp = wall_ms(st);
        setup_relaxation(ctx_, prm_, L, level_base + li);
        L.t_relax = wall_ms(st) - tp;
        if (level_base + (int)levels_.size() >= prm_.max_levels)
            break; // last level is a plain smoothing-only level
        // ---- aggregates
        tp = wall_ms(st);
        const int gi = level_base + li;
        build_aggregates(ctx_, tmp, prm_, A, eps_strong, gi < (int)imposed.size() ? &imposed[gi] : nullptr, L);
        L.t_agg = wall_ms(st) - tp;
        if (L.n_agg <= 0)
            break;
        // ---- omega = relax * (4/3) / rho_Gershgorin(D^-1 A)
        double omega = prm_.sa_relax;
        if (prm_.estimate_spectral_radius)
            omega *= (4.0 / 3.0) / gershgorin_rho(ctx_, prm_, L);
        else
            omega *= 2.0 / 3.0;
        L.omega = omega;
        // ---- smoothed prolongation
        tp = wall_ms(st);
        build_prolongation(ctx_, tmp, prm_, A, prm_.block_size > 1 ? &L.Ahat : nullptr, eps_strong, omega, L);
        L.t_prolong = wall_ms(st) - tp;
        eps_strong *= 0.5;
        tp = wall_ms(st);
        transpose(ctx_, tmp, L.P, L.R);
        L.P.plan("auto", st);
        L.R.plan("auto", st);
        L.t_transpose = wall_ms(st) - tp;
        // ---- Galerkin coarse operator A_c = R (A P)
        auto next = std::make_unique<AmgLevel>();
        {
            CsrDev AP;
            tp = wall_ms(st);
            spgemm(ctx_, tmp, A, L.P, L.n_agg, AP);
            L.t_ap = wall_ms(st) - tp;
            tp = wall_ms(st);
            spgemm(ctx_, tmp, L.R, AP, L.n_agg, next->Aown);
            L.t_rap = wall_ms(st) - tp;
        }
        next->Aown.block = std::max(1, prm_.block_size); // P, R and the Galerkin product keep the full-block pattern
        next->Aown.plan("auto", st);
        next->Aown.refresh_bsr(st);
        next->A = &next->Aown;
        cur = std::move(next);
    }
    if (cur)
    {
        // coarsest level (rows <= coarse_enough): relaxation only, or the dense direct solve (direct_coarse = true)
        levels_.push_back(std::move(cur));
        const double tp = wall_ms(st);
        AmgLevel &Lc = *levels_.back();
        setup_relaxation(ctx_, prm_, Lc, level_base + (int)levels_.size() - 1);
        if (prm_.direct_coarse)
        {
            dense_inverse_build(ctx_, *Lc.A, Lc.Zinv, Lc.z_np);
            Lc.direct = true;
        }
        Lc.t_relax = wall_ms(st) - tp;
    }
    PSB_CUDA(cudaStreamSynchronize(st));
}

// ====================================================================================== cycle
// One smoother application (amgcl relaxation apply_pre == apply_post for chebyshev / damped jacobi).
// x and x_alt ping-pong: the fused SpMV step reads x (gathered) and writes x_alt. On return `x` points
// at the buffer that holds the result. after_step (row partitions): called with every new iterate, pushes its halo.
void relax_level(Ctx &ctx, const AmgParams &prm, AmgLevel &L, bool fine, const double *rhs, double *&x, double *&x_alt, bool x_is_zero,
                 const int *done, const std::function<void(const double *)> *after_step, const FusedPush *fused)
{
    const CsrDev &A = *L.Asm;
    const int B = std::max(1, prm.block_size);
    if (B > 1)
    {
        // M (b - A x) = Dblk^-1 b - Ahat x : scale the right-hand side once per application
        const long long nb = L.n / B;
        ctx.prof_begin("block_diag_apply");
        if (B == 2)
            block_diag_apply_kernel<2><<<nblk(nb), 256, 0, ctx.stream>>>(nb, L.dinvb.p, rhs, L.bh.p, done);
        else
            block_diag_apply_kernel<3><<<nblk(nb), 256, 0, ctx.stream>>>(nb, L.dinvb.p, rhs, L.bh.p, done);
        check_launch();
        ctx.prof_end();
        rhs = L.bh.p;
    }
    if (prm.relax_type == "chebyshev" && !after_step && A.halo_mask == 0 && A.n <= kSmallRows && A.nnz <= kSmallNnz && prm.degree <= kSmallDegree &&
        small_sweep_enabled())
    {
        // the whole application in one launch (levels that fit one SM; never on a row partition)
        static bool attr_set[16] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        constexpr size_t smem = sizeof(double) * (kSmallNnz + 2 * kSmallRows) + sizeof(int) * kSmallNnz;
        if (!attr_set[dev & 15])
        {
            PSB_CUDA(cudaFuncSetAttribute(cheb_sweep_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set[dev & 15] = true;
        }
        ChebCoef cf;
        for (int k = 0; k < prm.degree; ++k)
        {
            cf.alpha[k] = L.alpha[k];
            cf.beta[k] = L.beta[k];
        }
        ctx.prof_begin("cheb_sweep_small");
        cheb_sweep_small_kernel<<<1, 1024, smem, ctx.stream>>>(A.view(), rhs, L.dinv.p, x, L.cp.p, cf, prm.degree, x_is_zero ? 1 : 0, done);
        check_launch();
        ctx.prof_end();
        return;
    }
    if (prm.relax_type == "chebyshev")
    {
        for (int k = 0; k < prm.degree; ++k)
        {
            if (k == 0 && x_is_zero)
                launch_vec(ctx, "cheb_first", L.n_pad, OpChebFirst{rhs, L.dinv.p, L.cp.p, x, L.alpha[0]}, FinNone{}, done);
            else if (fused)
            {
                // row partition: the step pushes its own boundary rows (no separate push launch for this iterate)
                launch_spmv(ctx, fine ? "spmv_cheb_l0" : "spmv_cheb_coarse", A, x,
                            EpiChebPush{rhs, L.dinv.p, x, L.cp.p, x_alt, L.alpha[k], L.beta[k], fused->push_epoch, fused->pm},
                            FinPushDone{fused->push_epoch, fused->halo_expect, fused->world, fused->pm}, done);
                std::swap(x, x_alt);
                continue;
            }
            else
            {
                launch_spmv(ctx, fine ? "spmv_cheb_l0" : "spmv_cheb_coarse", A, x, EpiCheb{rhs, L.dinv.p, x, L.cp.p, x_alt, L.alpha[k], L.beta[k]},
                            FinNone{}, done);
                std::swap(x, x_alt);
            }
            if (after_step)
                (*after_step)(x);
        }
    }
    else
    {
        if (x_is_zero)
            launch_vec(ctx, "jacobi_first", L.n_pad, OpDiagFirst{rhs, L.w.p, x}, FinNone{}, done);
        else
        {
            launch_spmv(ctx, "spmv_jacobi", A, x, EpiRelaxDiag{rhs, L.w.p, x, x_alt}, FinNone{}, done);
            std::swap(x, x_alt);
        }
        if (after_step)
            (*after_step)(x);
    }
}

void AmgHierarchy::relax(int l, const double *rhs, double *&x, double *&x_alt, bool x_is_zero, const int *done)
{
    relax_level(ctx_, prm_, *levels_[l], l == 0 && level_base_ == 0, rhs, x, x_alt, x_is_zero, done, nullptr, nullptr);
}

// amgcl amg::cycle (SURVEY A.3 "Cycle"). x/x_alt are this level's iterate buffers; returns with the
// result in `x` (pointers may have been swapped).
void AmgHierarchy::cycle(int l, const double *rhs, double *&x, double *&x_alt, bool x_is_zero, const int *done)
{
    AmgLevel &L = *levels_[l];
    if (l + 1 == (int)levels_.size() && L.direct)
    {
        dense_inverse_apply(ctx_, L.n, L.z_np, L.Zinv.p, rhs, x, done); // amgcl: (*lvl->solve)(rhs, x)
        return;
    }
    if (l + 1 == (int)levels_.size())
    {
        bool zero = x_is_zero;
        for (int i = 0; i < prm_.npre; ++i)
        {
            relax(l, rhs, x, x_alt, zero, done);
            zero = false;
        }
        for (int i = 0; i < prm_.npost; ++i)
        {
            relax(l, rhs, x, x_alt, zero, done);
            zero = false;
        }
        if (zero) // npre = npost = 0
            PSB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * L.n_pad, ctx_.stream));
        return;
    }
    AmgLevel &N = *levels_[l + 1];
    bool zero = x_is_zero;
    for (int j = 0; j < prm_.ncycle; ++j)
    {
        for (int i = 0; i < prm_.npre; ++i)
        {
            relax(l, rhs, x, x_alt, zero, done);
            zero = false;
        }
        if (zero)
        {
            // no pre-smoothing and x == 0: t = rhs
            PSB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * L.n_pad, ctx_.stream));
            zero = false;
        }
        launch_spmv(ctx_, "spmv_residual", *L.A, x, EpiResidual{L.t.p, rhs}, FinNone{}, done);
        launch_spmv(ctx_, "spmv_restrict", L.R, L.t.p, EpiStore{N.f.p}, FinNone{}, done);
        double *nu = N.u.p, *nalt = N.ualt.p;
        cycle(l + 1, N.f.p, nu, nalt, true, done);
        launch_spmv(ctx_, "spmv_prolong", L.P, nu, EpiAddTo{x}, FinNone{}, done);
        for (int i = 0; i < prm_.npost; ++i)
            relax(l, rhs, x, x_alt, false, done);
    }
}

void AmgHierarchy::apply(const double *rhs, double *out, const int *done)
{
    if (levels_.empty())
        throw std::runtime_error("psb200 amg: empty hierarchy");
    AmgLevel &L0 = *levels_[0];
    const size_t bytes = sizeof(double) * (size_t)L0.n_pad;
    if (prm_.pre_cycles <= 0)
    {
        PSB_CUDA(cudaMemcpyAsync(out, rhs, bytes, cudaMemcpyDeviceToDevice, ctx_.stream));
        return;
    }
    double *x = L0.u.p, *alt = L0.ualt.p;
    bool zero = true;
    for (int i = 0; i < prm_.pre_cycles; ++i)
    {
        cycle(0, rhs, x, alt, zero, done);
        zero = false;
    }
    // note: after convergence (done set) the skipped kernels leave stale data here; callers ignore it
    PSB_CUDA(cudaMemcpyAsync(out, x, bytes, cudaMemcpyDeviceToDevice, ctx_.stream));
}

// ------------------------------------------------------------------------------------------ row-partitioned level 0
struct AmgDistFine
{
    const CsrDev *A = nullptr; // rank-local rows of the fine matrix, halo columns
    const double *dinv = nullptr;
    long long row0 = 0, nl = 0, nl_pad = 0;
    CsrDev P, R; // rows [row0, row0 + nl) of P_0 (nl x n_1) and their transpose (n_1 x nl)
    DevBuf<double> u, ualt, t, cp, partial;
    std::function<void(const double *, const int *)> push_halo;
    std::function<void(const double *, double *, long long, const int *)> allreduce;
};

__global__ void slice_rowptr_kernel(int nl, const int *__restrict__ rp, int row0, int *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nl)
        out[i] = rp[row0 + i] - rp[row0];
}

void AmgHierarchy::setup_dist_fine(const CsrDev &A_local, const double *dinv_local, long long row0,
                                   std::function<void(const double *, const int *)> push_halo,
                                   std::function<void(const double *, double *, long long, const int *)> allreduce)
{
    if (levels_.empty())
        throw std::runtime_error("psb200 amg: setup() before setup_dist_fine()");
    if (std::max(1, prm_.block_size) > 1)
        throw std::runtime_error("psb200 amg: the partitioned cycle is scalar (block problems use the rank-local hierarchy)");
    cudaStream_t st = ctx_.stream;
    auto D = std::make_shared<AmgDistFine>();
    D->A = &A_local;
    D->dinv = dinv_local;
    D->row0 = row0;
    D->nl = A_local.n;
    D->nl_pad = (D->nl + 3) & ~3ll;
    D->push_halo = std::move(push_halo);
    D->allreduce = std::move(allreduce);
    AmgLevel &L0 = *levels_[0];
    if (row0 < 0 || row0 + D->nl > L0.n)
        throw std::invalid_argument("psb200 amg: local row range outside the hierarchy's fine level");
    const size_t np = (size_t)std::max<long long>(D->nl_pad, 4);
    for (DevBuf<double> *b : {&D->u, &D->ualt, &D->t, &D->cp})
        b->alloc(np, true);
    if (levels_.size() > 1)
    {
        // P_local = rows of P_0, R_local = P_local^T (columns [row0, row0 + nl) of R_0 = P_0^T)
        const CsrDev &P0 = L0.P;
        const int nl = (int)D->nl;
        int k0 = 0, k1 = 0;
        PSB_CUDA(cudaMemcpyAsync(&k0, P0.rp.p + row0, sizeof(int), cudaMemcpyDeviceToHost, st));
        PSB_CUDA(cudaMemcpyAsync(&k1, P0.rp.p + row0 + nl, sizeof(int), cudaMemcpyDeviceToHost, st));
        PSB_CUDA(cudaStreamSynchronize(st));
        CsrDev &P = D->P;
        P.n = nl;
        P.ncols = P0.ncols;
        P.nnz = k1 - k0;
        P.rp.alloc((size_t)nl + 1);
        P.ci.alloc(std::max<long long>(1, P.nnz), false, 64);
        P.va.alloc(std::max<long long>(1, P.nnz), false, 64);
        slice_rowptr_kernel<<<nblk((long long)nl + 1), 256, 0, st>>>(nl, P0.rp.p, (int)row0, P.rp.p);
        check_launch();
        if (P.nnz)
        {
            PSB_CUDA(cudaMemcpyAsync(P.ci.p, P0.ci.p + k0, sizeof(int) * (size_t)P.nnz, cudaMemcpyDeviceToDevice, st));
            PSB_CUDA(cudaMemcpyAsync(P.va.p, P0.va.p + k0, sizeof(double) * (size_t)P.nnz, cudaMemcpyDeviceToDevice, st));
        }
        Temp tmp;
        transpose(ctx_, tmp, P, D->R);
        PSB_CUDA(cudaStreamSynchronize(st));
        P.plan("auto", st);
        D->R.plan("auto", st);
        D->partial.alloc((size_t)levels_[1]->n_pad + 4, true);
    }
    dist_ = D;
}

// AmgHierarchy::relax for the partitioned level 0: every vector that is multiplied next has its halo pushed first
void AmgHierarchy::relax_dist(const double *rhs, double *&x, double *&x_alt, bool x_is_zero, const int *done)
{
    AmgDistFine &D = *dist_;
    AmgLevel &L = *levels_[0];
    if (prm_.relax_type == "chebyshev")
    {
        for (int k = 0; k < prm_.degree; ++k)
        {
            if (k == 0 && x_is_zero)
                launch_vec(ctx_, "cheb_first", D.nl_pad, OpChebFirst{rhs, D.dinv, D.cp.p, x, L.alpha[0]}, FinNone{}, done);
            else
            {
                launch_spmv(ctx_, "spmv_cheb_l0", *D.A, x, EpiCheb{rhs, D.dinv, x, D.cp.p, x_alt, L.alpha[k], L.beta[k]}, FinNone{}, done);
                std::swap(x, x_alt);
            }
            D.push_halo(x, done);
        }
    }
    else
    {
        // damped Jacobi: w = damping * D^-1 lives in the full-length level; the local slice starts at row0
        // (EpiRelaxDiag / OpDiagFirst read it element-wise, so an odd row0 is fine for the SpMV epilogue only)
        throw std::runtime_error("psb200 amg: the partitioned cycle provides the Chebyshev smoother (polysolve's default, AMGCL.cpp:36-47)");
    }
}

// amgcl amg::cycle at the partitioned level 0; levels >= 1 run replicated through cycle(1, ...)
void AmgHierarchy::cycle_dist(const double *rhs, double *&x, double *&x_alt, bool x_is_zero, const int *done)
{
    AmgDistFine &D = *dist_;
    if (levels_.size() == 1)
    {
        bool zero = x_is_zero;
        for (int i = 0; i < prm_.npre + prm_.npost; ++i)
        {
            relax_dist(rhs, x, x_alt, zero, done);
            zero = false;
        }
        if (zero)
            PSB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * D.nl_pad, ctx_.stream));
        return;
    }
    AmgLevel &N = *levels_[1];
    bool zero = x_is_zero;
    for (int j = 0; j < prm_.ncycle; ++j)
    {
        for (int i = 0; i < prm_.npre; ++i)
        {
            relax_dist(rhs, x, x_alt, zero, done);
            zero = false;
        }
        if (zero)
        {
            PSB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * D.nl_pad, ctx_.stream));
            D.push_halo(x, done);
            zero = false;
        }
        launch_spmv(ctx_, "spmv_residual", *D.A, x, EpiResidual{D.t.p, rhs}, FinNone{}, done);
        // f_1 = R_0 t = sum over ranks of R_local t_local
        launch_spmv(ctx_, "spmv_restrict", D.R, D.t.p, EpiStore{D.partial.p}, FinNone{}, done);
        D.allreduce(D.partial.p, N.f.p, N.n, done);
        double *nu = N.u.p, *nalt = N.ualt.p;
        cycle(1, N.f.p, nu, nalt, true, done);
        launch_spmv(ctx_, "spmv_prolong", D.P, nu, EpiAddTo{x}, FinNone{}, done);
        D.push_halo(x, done);
        for (int i = 0; i < prm_.npost; ++i)
            relax_dist(rhs, x, x_alt, false, done);
    }
}

void AmgHierarchy::apply_dist(const double *rhs, double *out, const int *done)
{
    if (!dist_)
        throw std::runtime_error("psb200 amg: setup_dist_fine() has not been called");
    AmgDistFine &D = *dist_;
    const size_t bytes = sizeof(double) * (size_t)D.nl_pad;
    if (prm_.pre_cycles <= 0)
    {
        PSB_CUDA(cudaMemcpyAsync(out, rhs, bytes, cudaMemcpyDeviceToDevice, ctx_.stream));
        return;
    }
    double *x = D.u.p, *alt = D.ualt.p;
    bool zero = true;
    for (int i = 0; i < prm_.pre_cycles; ++i)
    {
        cycle_dist(rhs, x, alt, zero, done);
        zero = false;
    }
    PSB_CUDA(cudaMemcpyAsync(out, x, bytes, cudaMemcpyDeviceToDevice, ctx_.stream));
}

double AmgHierarchy::total_nnz() const
{
    double tot = 0;
    for (auto &L : levels_)
        tot += (double)L->A->nnz;
    return tot;
}

std::string AmgHierarchy::levels_json() const
{
    std::ostringstream o;
    for (size_t l = 0; l < levels_.size(); ++l)
    {
        const AmgLevel &L = *levels_[l];
        if (l)
            o << ",";
        o << "{\"rows\":" << L.A->n << ",\"nnz\":" << L.A->nnz << ",\"p_nnz\":" << L.P.nnz << ",\"aggregates\":" << L.n_agg
          << ",\"rho\":" << jnum(L.rho) << ",\"omega\":" << jnum(L.omega) << ",\"mis_rounds\":" << L.mis_rounds << ",\"setup_ms\":{\"relax\":" << jnum(L.t_relax)
          << ",\"aggregate\":" << jnum(L.t_agg) << ",\"prolong\":" << jnum(L.t_prolong) << ",\"transpose\":" << jnum(L.t_transpose)
          << ",\"AP\":" << jnum(L.t_ap) << ",\"RAP\":" << jnum(L.t_rap) << "}"
          << ",\"spmv_kernel\":" << jstr(L.A->kernel_name()) << "}";
    }
    return o.str();
}

std::string AmgHierarchy::info_json() const
{
    std::ostringstream o;
    const double fine_nnz = levels_.empty() ? 1 : (double)levels_[0]->A->nnz;
    o << "{\"levels\":[" << levels_json();
    o << "],\"block_size\":" << std::max(1, prm_.block_size) << ",\"operator_complexity\":" << jnum(total_nnz() / fine_nnz) << ",\"ncycle\":" << prm_.ncycle << ",\"degree\":" << prm_.degree
      << ",\"relax\":" << jstr(prm_.relax_type) << "}";
    return o.str();
}

double *AmgHierarchy::level0_f() { return levels_.at(0)->f.p; }

double *AmgHierarchy::cycle0(const int *done)
{
    AmgLevel &L0 = *levels_.at(0);
    double *x = L0.u.p, *alt = L0.ualt.p;
    cycle(0, L0.f.p, x, alt, true, done);
    return x;
}

} // namespace psb
