// Device kernels of the Krylov hot path: CSR SpMV (three schedules) with fused epilogues, fused
// vector updates with deterministic reductions, and the device-resident solver state.
// Everything is fp64 values + int32 indices (reference Types.hpp:11-15).
#pragma once
#include "common.cuh"

#include <cfloat>

namespace psb {

// ---------------------------------------------------------------------------------- solver state
// Lives in device memory; the host reads a copy when it polls. Scalars never round-trip through the
// host inside the iteration (policy precedent: reference MASSolver.cu:46-56 keeps alpha/beta on device).
struct KState
{
    double rz, rz_new, pAp, rn2, bn2, thr, tol;
    double rho, rho_old, alpha, omega, r0v, tt, ts, r0n2;
    double c_beta; // cg1r: beta of the current trip
    int iter, done, status, max_iter, restart, restarts, c_started, pad1;
};
enum : int
{
    ST_RUNNING = 0,
    ST_CONVERGED = 1,
    ST_MAXITER = 2,
    ST_BREAKDOWN = 3,
    ST_ZERO_RHS = 4,
    ST_COMM = 5 // a grid-wide or cross-GPU wait timed out (persistent kernel)
};

struct CsrView
{
    const int *rp;
    const int *ci;
    const double *va;
    int n;
    // row-partitioned (multi-GPU) matrices: columns >= nl address the halo values the peers pushed
    // into this rank's comm buffer instead of the local x. Single GPU: nl = INT_MAX, halo_mask = 0.
    int nl;
    unsigned halo_mask; // non-zero: the matrix has halo columns (the ranks to wait for are CommDev::nbr_mask)
    // stream schedule on a row partition: sequence position -> tile with the tiles that touch no halo column first
    // (positions < n_interior), so the multiplication starts while the neighbours' pushes are still on the wire
    const int *tile_order;
    int n_interior;
};

// x[c] for a local column, xh[c - nl] for a halo column. Halo values were written by a peer GPU: they are
// read at L2 (ld.global.cg), never through the non-coherent L1.
__device__ __forceinline__ double ldx(const double *__restrict__ x, const double *__restrict__ xh, int nl, int c)
{
    return c < nl ? __ldg(x + c) : __ldcg(xh + (c - nl));
}

// Waits until everything the neighbours have pushed so far has landed completely: the flag of source q counts landed
// push chunks and halo_expect[q] is the running total this rank's own push kernels have announced (common.cuh, CommDev).
// Called by all threads of a CTA.
__device__ __forceinline__ void wait_pushes_landed(const CommDev &c)
{
    if ((int)threadIdx.x < c.world && ((c.nbr_mask >> threadIdx.x) & 1u))
    {
        if (!spin_ge(c.halo_flag(c.rank, threadIdx.x), c.halo_expect[threadIdx.x], c.error, c.spin_limit))
            *c.error = 1;
        fence_acq_rel_sys();
    }
    __syncthreads();
}
// Kernel-start form for a multiplying kernel: returns the halo base pointer of the vector pushed last.
__device__ __forceinline__ const double *wait_halo(const CsrView &A, const CommDev &c)
{
    if (A.halo_mask == 0)
        return nullptr;
    wait_pushes_landed(c);
    return c.halo(c.rank, (int)(*c.push_epoch % kHaloBufs), 0);
}

// ---------------------------------------------------------------------------------- finalizers
// Run by thread 0 of the last CTA with the grid totals.
struct FinNone
{
    __device__ __forceinline__ void operator()(const double *) const {}
};
struct FinStore
{
    double *out;
    int nv;
    __device__ __forceinline__ void operator()(const double *t) const
    {
        for (int i = 0; i < nv; ++i)
            out[i] = t[i];
    }
};
struct FinPAp
{
    KState *st;
    __device__ __forceinline__ void operator()(const double *t) const { st->pAp = t[0]; }
};
// Eigen conjugate_gradient() prologue: thresholds and the two early exits.
struct FinInitEigen
{
    KState *st;
    __device__ __forceinline__ void operator()(const double *t) const
    {
        st->rn2 = t[0];
        st->bn2 = t[1];
        st->rz = t[2];
        st->thr = fmax(st->tol * st->tol * t[1], DBL_MIN);
        if (t[1] == 0.0)
        {
            st->done = 1;
            st->status = ST_ZERO_RHS;
        }
        else if (t[0] < st->thr)
        {
            st->done = 1;
            st->status = ST_CONVERGED;
        }
        else if (!(t[0] == t[0]) || isinf(t[0]))
        {
            st->done = 1;
            st->status = ST_BREAKDOWN;
        }
    }
};
// amgcl::solver::cg prologue (SURVEY A.3 "CG"): eps = max(tol ||b||, min), loop condition ||r|| > eps.
struct FinInitAmgcl
{
    KState *st;
    __device__ __forceinline__ void operator()(const double *t) const
    {
        st->rn2 = t[0];
        st->bn2 = t[1];
        const double nb = sqrt(t[1]);
        st->thr = fmax(st->tol * nb, DBL_MIN);
        if (nb < DBL_EPSILON)
        {
            st->done = 1;
            st->status = ST_ZERO_RHS;
        }
        else if (!(sqrt(t[0]) > st->thr))
        {
            st->done = 1;
            st->status = (t[0] == t[0]) ? ST_CONVERGED : ST_BREAKDOWN;
        }
    }
};
struct FinCgUpdateEigen
{
    KState *st;
    __device__ __forceinline__ void operator()(const double *t) const
    {
        st->rn2 = t[0];
        st->rz_new = t[1];
        if (t[0] < st->thr) // Eigen breaks before i++
        {
            st->done = 1;
            st->status = ST_CONVERGED;
        }
        else if (!(t[0] == t[0]) || isinf(t[0]))
        {
            st->done = 1;
            st->status = ST_BREAKDOWN;
        }
    }
};
struct FinCgDirEigen
{
    KState *st;
    __device__ __forceinline__ void operator()(const double *) const
    {
        st->rz = st->rz_new;
        st->iter += 1;
        if (st->iter >= st->max_iter)
        {
            st->done = 1;
            st->status = ST_MAXITER;
        }
    }
};
struct FinRhoAmgcl
{
    KState *st;
    __device__ __forceinline__ void operator()(const double *t) const
    {
        st->rho_old = st->rho;
        st->rho = t[0];
    }
};
struct FinCgUpdateAmgcl
{
    KState *st;
    __device__ __forceinline__ void operator()(const double *t) const
    {
        st->rn2 = t[0];
        st->iter += 1;
        if (!(sqrt(t[0]) > st->thr))
        {
            st->done = 1;
            st->status = (t[0] == t[0]) ? ST_CONVERGED : ST_BREAKDOWN;
        }
        else if (st->iter >= st->max_iter)
        {
            st->done = 1;
            st->status = ST_MAXITER;
        }
    }
};
// Eigen bicgstab()
struct FinInitBicg
{
    KState *st;
    __device__ __forceinline__ void operator()(const double *t) const
    {
        st->rn2 = t[0];
        st->bn2 = t[1];
        st->r0n2 = t[0];
        st->thr = st->tol * st->tol * t[1];
        st->rho = 1;
        st->alpha = 1;
        st->omega = 1;
        if (t[1] == 0.0)
        {
            st->done = 1;
            st->status = ST_ZERO_RHS;
        }
        else if (!(t[0] > st->thr))
        {
            st->done = 1;
            st->status = (t[0] == t[0]) ? ST_CONVERGED : ST_BREAKDOWN;
        }
        else
        {
            // first loop trip: rho = r0.r = ||r0||^2
            st->rho_old = 1;
            st->rho = t[0];
            st->restart = (fabs(t[0]) < DBL_EPSILON * DBL_EPSILON * t[0]) ? 1 : 0;
        }
    }
};
struct FinBicgAlpha
{
    KState *st;
    __device__ __forceinline__ void operator()(const double *t) const
    {
        st->r0v = t[0];
        st->alpha = st->rho / t[0];
    }
};
struct FinBicgOmega
{
    KState *st;
    __device__ __forceinline__ void operator()(const double *t) const
    {
        st->tt = t[0];
        st->ts = t[1];
        st->omega = t[0] > 0.0 ? t[1] / t[0] : 0.0;
    }
};
// end of a BiCGSTAB trip: t[0] = ||r||^2, t[1] = r0.r for the next trip
struct FinBicgEnd
{
    KState *st;
    __device__ __forceinline__ void operator()(const double *t) const
    {
        st->rn2 = t[0];
        st->iter += 1;
        if (!(t[0] > st->thr))
        {
            st->done = 1;
            st->status = (t[0] == t[0]) ? ST_CONVERGED : ST_BREAKDOWN;
        }
        else if (st->iter >= st->max_iter)
        {
            st->done = 1;
            st->status = ST_MAXITER;
        }
        else
        {
            st->rho_old = st->rho;
            st->rho = t[1];
            st->restart = (fabs(t[1]) < DBL_EPSILON * DBL_EPSILON * st->r0n2) ? 1 : 0;
        }
    }
};
// restart: r = b - A x recomputed, r0 = r; t[0] = ||r||^2
struct FinBicgRestart
{
    KState *st;
    __device__ __forceinline__ void operator()(const double *t) const
    {
        st->rho = t[0];
        st->r0n2 = t[0];
        st->rn2 = t[0];
        if (st->restarts++ == 0)
            st->iter = 0;
        st->restart = 0;
    }
};

} // namespace psb

#include "spmv.cuh"

namespace psb {

// ---------------------------------------------------------------------------------- fused vector kernels
// Element-wise ops over double2 (128-bit) lanes with up to kMaxRed fused reductions. Vectors are
// padded with zeros to an even length, so no tail handling is needed. n2 = padded_length / 2.
template <class Op, class Fin, int THREADS>
__global__ void __launch_bounds__(THREADS) vec_kernel(long long n2, Op op, RedCtx rc, Fin fin, const int *done, const int *only_if)
{
    griddep_launch_dependents();
    griddep_wait();
    if (done && *done)
        return;
    if (only_if && !*only_if)
        return;
    constexpr int NVA = Op::NV > 0 ? Op::NV : 1;
    double acc[NVA];
#pragma unroll
    for (int i = 0; i < NVA; ++i)
        acc[i] = 0;
    op.prologue();
    const long long stride = (long long)gridDim.x * THREADS;
    long long j = (long long)blockIdx.x * THREADS + threadIdx.x;
    // two independent 128-bit lanes per trip for memory-level parallelism
    for (; j + stride < n2; j += 2 * stride)
    {
        op.template apply<2>(j, stride, acc);
    }
    if (j < n2)
        op.template apply<1>(j, stride, acc);
    double tot[NVA];
    if (grid_reduce<Op::NV, THREADS>(acc, rc, tot) && threadIdx.x == 0)
        fin(tot);
}

__device__ __forceinline__ double2 ld2(const double *p, long long j) { return reinterpret_cast<const double2 *>(p)[j]; }
__device__ __forceinline__ void st2(double *p, long long j, double2 v) { reinterpret_cast<double2 *>(p)[j] = v; }

// Eigen CG: x += alpha p; r -= alpha q; ||r||^2; r.(dinv r)   (SURVEY A.1 loop body, first half)
struct OpCgUpdateEigen
{
    static constexpr int NV = 2;
    double *x, *r;
    const double *p, *q, *dinv;
    const KState *st;
    double alpha;
    __device__ __forceinline__ void prologue() { alpha = st->rz / st->pAp; }
    template <int U>
    __device__ __forceinline__ void apply(long long j, long long stride, double (&acc)[2])
    {
        double2 xv[U], rv[U], pv[U], qv[U], dv[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            xv[u] = ld2(x, j + u * stride);
            rv[u] = ld2(r, j + u * stride);
            pv[u] = ld2(p, j + u * stride);
            qv[u] = ld2(q, j + u * stride);
            dv[u] = ld2(dinv, j + u * stride);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            xv[u].x += alpha * pv[u].x;
            xv[u].y += alpha * pv[u].y;
            rv[u].x -= alpha * qv[u].x;
            rv[u].y -= alpha * qv[u].y;
            st2(x, j + u * stride, xv[u]);
            st2(r, j + u * stride, rv[u]);
            acc[0] += rv[u].x * rv[u].x + rv[u].y * rv[u].y;
            acc[1] += rv[u].x * rv[u].x * dv[u].x + rv[u].y * rv[u].y * dv[u].y;
        }
    }
};
// Eigen CG: p = dinv r + beta p   (FIRST: p = dinv r)
template <bool FIRST>
struct OpCgDirEigen
{
    static constexpr int NV = 0;
    double *p;
    const double *r, *dinv;
    const KState *st;
    double beta;
    __device__ __forceinline__ void prologue() { beta = FIRST ? 0.0 : st->rz_new / st->rz; }
    template <int U>
    __device__ __forceinline__ void apply(long long j, long long stride, double (&)[1])
    {
        double2 rv[U], pv[U], dv[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            rv[u] = ld2(r, j + u * stride);
            dv[u] = ld2(dinv, j + u * stride);
            if (!FIRST)
                pv[u] = ld2(p, j + u * stride);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            double2 o;
            o.x = dv[u].x * rv[u].x;
            o.y = dv[u].y * rv[u].y;
            if (!FIRST)
            {
                o.x += beta * pv[u].x;
                o.y += beta * pv[u].y;
            }
            st2(p, j + u * stride, o);
        }
    }
};
// generic dot a.b
struct OpDot
{
    static constexpr int NV = 1;
    const double *a, *b;
    __device__ __forceinline__ void prologue() {}
    template <int U>
    __device__ __forceinline__ void apply(long long j, long long stride, double (&acc)[1])
    {
        double2 av[U], bv[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            av[u] = ld2(a, j + u * stride);
            bv[u] = ld2(b, j + u * stride);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            acc[0] += av[u].x * bv[u].x + av[u].y * bv[u].y;
    }
};
// amgcl cg: p = s + (rho/rho_old) p, or p = s on the first trip
struct OpCgDirAmgcl
{
    static constexpr int NV = 0;
    double *p;
    const double *s;
    const KState *st;
    double beta;
    __device__ __forceinline__ void prologue() { beta = st->iter ? st->rho / st->rho_old : 0.0; }
    template <int U>
    __device__ __forceinline__ void apply(long long j, long long stride, double (&)[1])
    {
        double2 sv[U], pv[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            sv[u] = ld2(s, j + u * stride);
            pv[u] = ld2(p, j + u * stride);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            double2 o;
            o.x = sv[u].x + (beta != 0.0 ? beta * pv[u].x : 0.0);
            o.y = sv[u].y + (beta != 0.0 ? beta * pv[u].y : 0.0);
            st2(p, j + u * stride, o);
        }
    }
};
// amgcl cg: x += alpha p; r -= alpha q; ||r||^2   with alpha = rho / p.q
struct OpCgUpdateAmgcl
{
    static constexpr int NV = 1;
    double *x, *r;
    const double *p, *q;
    const KState *st;
    double alpha;
    __device__ __forceinline__ void prologue() { alpha = st->rho / st->pAp; }
    template <int U>
    __device__ __forceinline__ void apply(long long j, long long stride, double (&acc)[1])
    {
        double2 xv[U], rv[U], pv[U], qv[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            xv[u] = ld2(x, j + u * stride);
            rv[u] = ld2(r, j + u * stride);
            pv[u] = ld2(p, j + u * stride);
            qv[u] = ld2(q, j + u * stride);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            xv[u].x += alpha * pv[u].x;
            xv[u].y += alpha * pv[u].y;
            rv[u].x -= alpha * qv[u].x;
            rv[u].y -= alpha * qv[u].y;
            st2(x, j + u * stride, xv[u]);
            st2(r, j + u * stride, rv[u]);
            acc[0] += rv[u].x * rv[u].x + rv[u].y * rv[u].y;
        }
    }
};
// BiCGSTAB: p = r + beta (p - omega v);  y = dinv p      beta = (rho/rho_old)(alpha/omega)
struct OpBicgP
{
    static constexpr int NV = 0;
    double *p, *y;
    const double *r, *v, *dinv;
    const KState *st;
    double beta, omega;
    __device__ __forceinline__ void prologue()
    {
        beta = (st->rho / st->rho_old) * (st->alpha / st->omega);
        omega = st->omega;
    }
    template <int U>
    __device__ __forceinline__ void apply(long long j, long long stride, double (&)[1])
    {
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const long long i = j + u * stride;
            const double2 rv = ld2(r, i), pv = ld2(p, i), vv = ld2(v, i), dv = ld2(dinv, i);
            double2 o, yo;
            o.x = rv.x + beta * (pv.x - omega * vv.x);
            o.y = rv.y + beta * (pv.y - omega * vv.y);
            yo.x = dv.x * o.x;
            yo.y = dv.y * o.y;
            st2(p, i, o);
            st2(y, i, yo);
        }
    }
};
// BiCGSTAB: s = r - alpha v (in place in r);  z = dinv s
struct OpBicgS
{
    static constexpr int NV = 0;
    double *r, *z;
    const double *v, *dinv;
    const KState *st;
    double alpha;
    __device__ __forceinline__ void prologue() { alpha = st->alpha; }
    template <int U>
    __device__ __forceinline__ void apply(long long j, long long stride, double (&)[1])
    {
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const long long i = j + u * stride;
            const double2 rv = ld2(r, i), vv = ld2(v, i), dv = ld2(dinv, i);
            double2 s, zo;
            s.x = rv.x - alpha * vv.x;
            s.y = rv.y - alpha * vv.y;
            zo.x = dv.x * s.x;
            zo.y = dv.y * s.y;
            st2(r, i, s);
            st2(z, i, zo);
        }
    }
};
// BiCGSTAB: x += alpha y + omega z;  r = s - omega t;  ||r||^2;  r0.r
struct OpBicgEnd
{
    static constexpr int NV = 2;
    double *x, *r;
    const double *y, *z, *t, *r0;
    const KState *st;
    double alpha, omega;
    __device__ __forceinline__ void prologue()
    {
        alpha = st->alpha;
        omega = st->omega;
    }
    template <int U>
    __device__ __forceinline__ void apply(long long j, long long stride, double (&acc)[2])
    {
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const long long i = j + u * stride;
            double2 xv = ld2(x, i), rv = ld2(r, i);
            const double2 yv = ld2(y, i), zv = ld2(z, i), tv = ld2(t, i), r0v = ld2(r0, i);
            xv.x += alpha * yv.x + omega * zv.x;
            xv.y += alpha * yv.y + omega * zv.y;
            rv.x -= omega * tv.x;
            rv.y -= omega * tv.y;
            st2(x, i, xv);
            st2(r, i, rv);
            acc[0] += rv.x * rv.x + rv.y * rv.y;
            acc[1] += r0v.x * rv.x + r0v.y * rv.y;
        }
    }
};
// y = a (copy), used for r0 = r
struct OpCopy
{
    static constexpr int NV = 0;
    double *y;
    const double *a;
    __device__ __forceinline__ void prologue() {}
    template <int U>
    __device__ __forceinline__ void apply(long long j, long long stride, double (&)[1])
    {
#pragma unroll
        for (int u = 0; u < U; ++u)
            st2(y, j + u * stride, ld2(a, j + u * stride));
    }
};
// y = s * a
struct OpScale
{
    static constexpr int NV = 0;
    double *y;
    const double *a;
    const double *scal; // device scalar: y = a / sqrt(*scal)
    double f;
    __device__ __forceinline__ void prologue() { f = rsqrt(*scal); }
    template <int U>
    __device__ __forceinline__ void apply(long long j, long long stride, double (&)[1])
    {
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            double2 v = ld2(a, j + u * stride);
            v.x *= f;
            v.y *= f;
            st2(y, j + u * stride, v);
        }
    }
};

} // namespace psb
