// Persistent Jacobi-PCG kernel (Eigen's ordering; reference path EigenSolver.tpp:108-114 ->
// Eigen conjugate_gradient()): ONE cooperative launch runs a whole batch of iterations. The three phases of
// an iteration -- q = A p with p.q | x += a p, r -= a q with ||r||^2, r.D^-1 r | p = D^-1 r + b p with the halo
// push -- are separated by a grid-wide sync that carries the reduction: every CTA deposits its partial sums and
// takes a ticket; the CTA that arrives last adds the partials in CTA order, exchanges the totals with the peer
// GPUs over NVLink (tagged words, no fences), runs the scalar finalizer on the device-resident KState and
// releases the grid. Compared with one kernel per phase this removes the launch/drain gaps (3 per iteration),
// keeps the TMA pipeline of the SpMV primed across phases (the matrix tiles of the next SpMV are prefetched
// while the vector phases run) and lets the SpMV start on the tiles that need no halo value while the
// neighbours' pushes are still on the wire (interior-first tile order; the wait happens before the first
// boundary tile). Arithmetic, update order and stopping rule are those of the split path (kernels.cuh), so
// both produce the same iterates; reductions add the per-CTA partials in CTA order (bit-reproducible for a
// fixed grid).
#include "dist.hpp"
#include "solver.hpp"

namespace psb {

struct GridBar
{
    unsigned long long count;    // CTA arrivals since the solve started (sync m completes at m * gridDim.x)
    unsigned long long released; // index of the last completed sync
};

struct PersistArgs
{
    CsrView A;
    const int *order; // sequence position -> tile (interior tiles first); nullptr: identity
    int n_interior;   // sequence positions below this touch no halo column
    double *x, *r, *q, *p0, *p1; // p0 holds the direction at kernel entry, p1 is the ping-pong partner
    const double *dinv;
    long long n2; // padded length / 2 (double2 lanes)
    KState *st;
    double *partials;
    int pstride;
    GridBar *bar;
    CommDev comm;
    PushList push;
    int iters; // iterations per launch (even, so p0 is current again at the next launch)
    long long *timing; // optional: SM cycles CTA 0 spent in [spmv, sync, update, sync, dir+push, sync] (accumulated)
};

using PCfg = StreamProd;
constexpr int kPersistCtasPerSm = 4;

// grid-wide sync + reduction + finalizer. m = index of this sync (1-based since the solve started), seq = index of
// the cross-GPU reduction it carries. Returns false on a spin timeout (lost CTA / lost peer).
template <int NV, class Fin>
__device__ __forceinline__ bool persist_sync(double (&v)[NV > 0 ? NV : 1], const PersistArgs &a, unsigned long long m, unsigned long long seq, Fin fin)
{
    constexpr int T = PCfg::threads;
    __shared__ double sm[NV > 0 ? NV : 1][T / 32];
    __shared__ int s_role; // 1: last arriver, 0: waiter, -1: timeout
    if constexpr (NV > 0)
    {
        block_sum<NV, T>(v, sm);
        if (threadIdx.x == 0)
        {
#pragma unroll
            for (int i = 0; i < NV; ++i)
                a.partials[i * a.pstride + blockIdx.x] = v[i];
        }
    }
    else
        __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence();
        const unsigned long long old = atomicAdd(&a.bar->count, 1ull);
        s_role = (old == m * gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_role == 1)
    {
        __threadfence();
        double tot[NV > 0 ? NV : 1];
        if constexpr (NV > 0)
        {
            double s[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i)
            {
                s[i] = 0;
                for (int b = threadIdx.x; b < (int)gridDim.x; b += T)
                    s[i] += __ldcg(a.partials + i * a.pstride + b);
            }
            block_sum<NV, T>(s, sm);
#pragma unroll
            for (int i = 0; i < NV; ++i)
                tot[i] = s[i];
            if (a.comm.world > 1)
                comm_allreduce_seq<NV>(a.comm, tot, seq);
        }
        if (threadIdx.x == 0)
        {
            fin(tot);
            if (a.comm.world > 1 && *a.comm.error)
            {
                a.st->done = 1;
                a.st->status = ST_COMM;
            }
            __threadfence();
            st_release_gpu(&a.bar->released, m);
        }
    }
    else if (threadIdx.x == 0)
    {
        const long long t0 = clock64();
        while (ld_acquire_gpu(&a.bar->released) < m)
        {
            if (clock64() - t0 > kSpinLimit)
            {
                s_role = -1;
                a.st->done = 1;
                a.st->status = ST_COMM;
                break;
            }
            __nanosleep(32);
        }
    }
    __syncthreads();
    return s_role >= 0;
}

__global__ void __launch_bounds__(PCfg::threads, kPersistCtasPerSm) cg_persist_kernel(PersistArgs a)
{
    constexpr int T = PCfg::threads, CAP = PCfg::cap, STAGES = PCfg::stages;
    KState *st = a.st;
    if (*(volatile int *)&st->done)
        return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sval = reinterpret_cast<double *>(smem_raw);
    int *scol = reinterpret_cast<int *>(smem_raw + PCfg::val_bytes);
    int *srp = reinterpret_cast<int *>(smem_raw + PCfg::val_bytes + PCfg::col_bytes);
    __shared__ __align__(8) unsigned long long bar[STAGES];

    const CsrView &A = a.A;
    const int G = (int)gridDim.x;
    const int ntiles = (A.n + T - 1) / T;
    const int nj = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + G - 1) / G : 0; // tiles of this CTA per SpMV
    const long long total_seq = (long long)a.iters * nj;
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

    auto tile_at = [&](int j) {
        const int pos = (int)blockIdx.x + j * G;
        return a.order ? __ldg(a.order + pos) : pos;
    };
    // one thread: the three bulk copies of `tile` into stage s (same staging as spmv_stream_kernel)
    auto issue = [&](int tile, int s) {
        const int r0 = tile * T;
        const int r1 = min(A.n, r0 + T);
        const int k0 = __ldg(A.rp + r0), k1 = __ldg(A.rp + r1);
        const int ka = k0 & ~3;
        const int cnt4 = (k1 - ka + 3) & ~3;
        const unsigned rp_b = (unsigned)(((r1 - r0 + 1) + 3) & ~3) * 4u;
        unsigned bytes = rp_b;
        const bool staged = cnt4 > 0 && cnt4 <= CAP;
        if (staged)
            bytes += (unsigned)cnt4 * 12u;
        mbar_expect_tx(&bar[s], bytes);
        tma_bulk_g2s(srp + (size_t)s * PCfg::rp_ints, A.rp + r0, rp_b, &bar[s], policy);
        if (staged)
        {
            tma_bulk_g2s(sval + (size_t)s * CAP, A.va + ka, (unsigned)cnt4 * 8u, &bar[s], policy);
            tma_bulk_g2s(scol + (size_t)s * CAP, A.ci + ka, (unsigned)cnt4 * 4u, &bar[s], policy);
        }
    };

    // the matrix never changes: the TMA pipeline runs STAGES tiles ahead across phase and iteration boundaries
    long long consumed = 0; // tiles this CTA has consumed since kernel entry; issued == min(total_seq, consumed + STAGES)
    int j_issue = 0;
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES && s < total_seq; ++s)
        {
            issue(tile_at(j_issue), s);
            j_issue = (j_issue + 1 == nj) ? 0 : j_issue + 1;
        }

    unsigned long long m = ld_acquire_gpu(&a.bar->released); // syncs completed by earlier launches of this solve
    unsigned long long seq = 0, epoch = 0;
    if (a.comm.world > 1)
    {
        seq = *a.comm.red_seq;
        epoch = *a.comm.push_epoch;
    }
    const long long vstride = (long long)G * T;
    bool ok = true;
    long long tph[6] = {0, 0, 0, 0, 0, 0};
    long long tc = clock64();
    auto lap = [&](int i) {
        const long long t = clock64();
        tph[i] += t - tc;
        tc = t;
    };

    for (int it = 0; it < a.iters && ok; ++it)
    {
        if (*(volatile int *)&st->done)
            break;
        double *pc = (it & 1) ? a.p1 : a.p0, *pn = (it & 1) ? a.p0 : a.p1;

        // ---------------------------------------------------------------- phase A: q = A p, p.q
        double acc1[1] = {0};
        {
            const double *xh = nullptr;
            bool halo_ready = A.halo_mask == 0;
            for (int j = 0; j < nj; ++j)
            {
                const int pos = (int)blockIdx.x + j * G;
                const int tile = a.order ? __ldg(a.order + pos) : pos;
                if (!halo_ready && pos >= a.n_interior)
                {
                    xh = wait_halo_epoch(A.halo_mask, a.comm, epoch);
                    halo_ready = true;
                }
                const int s = (int)(consumed % STAGES);
                const unsigned parity = (unsigned)((consumed / STAGES) & 1);
                const int r0 = tile * T;
                const int nrow = min(A.n - r0, T);
                const int row = r0 + threadIdx.x;
                const bool live = (int)threadIdx.x < nrow;
                double prow = 0;
                if (live)
                    prow = pc[row];
                mbar_wait(&bar[s], parity);
                const int *rps = srp + (size_t)s * PCfg::rp_ints;
                const int k0 = rps[0], k1 = rps[nrow];
                const int ka = k0 & ~3;
                const bool staged = ((k1 - ka + 3) & ~3) <= CAP;
                int kb = 0, ke = 0;
                if (live)
                {
                    kb = rps[threadIdx.x];
                    ke = rps[threadIdx.x + 1];
                }
                double sum = 0;
                if (staged)
                {
                    const double *sv = sval + (size_t)s * CAP - ka;
                    const int *sc = scol + (size_t)s * CAP - ka;
                    for (int k = kb; k < ke; k += 8)
                    {
                        double v[8], xx[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                        {
                            const bool in = k + u < ke;
                            const int c = in ? sc[k + u] : 0;
                            v[u] = in ? sv[k + u] : 0.0;
                            // p changes inside this kernel: coherent loads only (never ld.global.nc)
                            xx[u] = in ? (c < A.nl ? pc[c] : __ldcg(xh + (c - A.nl))) : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (k + u < ke)
                                sum += v[u] * xx[u];
                    }
                }
                else
                {
                    for (int k = kb; k < ke; ++k)
                    {
                        const int c = __ldg(A.ci + k);
                        sum += __ldg(A.va + k) * (c < A.nl ? pc[c] : __ldcg(xh + (c - A.nl)));
                    }
                }
                if (live)
                {
                    a.q[row] = sum;
                    acc1[0] += prow * sum;
                }
                __syncthreads(); // every thread is done reading stage s
                ++consumed;
                if (threadIdx.x == 0 && consumed + STAGES - 1 < total_seq)
                {
                    issue(tile_at(j_issue), s);
                    j_issue = (j_issue + 1 == nj) ? 0 : j_issue + 1;
                }
            }
        }
        lap(0);
        ++m;
        ++seq;
        ok = persist_sync<1>(acc1, a, m, seq, FinPAp{st});
        lap(1);
        if (!ok)
            break;

        // ---------------------------------------------------------------- phase B: x += alpha p, r -= alpha q, ||r||^2, r.D^-1 r
        double acc2[2] = {0, 0};
        {
            // all loads of both lanes are issued before the first store (memory-level parallelism at 4 CTAs / SM)
            OpCgUpdateEigen op{a.x, a.r, pc, a.q, a.dinv, st, 0.0};
            op.prologue();
            long long j = (long long)blockIdx.x * T + threadIdx.x;
            for (; j + vstride < a.n2; j += 2 * vstride)
                op.apply<2>(j, vstride, acc2);
            if (j < a.n2)
                op.apply<1>(j, vstride, acc2);
        }
        lap(2);
        ++m;
        ++seq;
        ok = persist_sync<2>(acc2, a, m, seq, FinCgUpdateEigen{st});
        lap(3);
        if (!ok || *(volatile int *)&st->done)
            break; // Eigen leaves the loop before the direction update and before i++

        // ---------------------------------------------------------------- phase C: p = D^-1 r + beta p, halo push
        {
            const double beta = st->rz_new / st->rz;
            if (a.comm.world > 1)
            {
                const int par = (int)((epoch + 1) & 1);
                for (int ch = blockIdx.x; ch < a.push.nchunks; ch += G)
                {
                    const int peer = a.push.chunk_peer[ch], start = a.push.chunk_start[ch], cnt = a.push.chunk_cnt[ch];
                    double *dst = a.comm.halo(peer, par, a.comm.rank) + a.push.chunk_off[ch];
                    for (int e = threadIdx.x; e < cnt; e += T)
                    {
                        const int row = a.push.rows[start + e];
                        dst[e] = a.dinv[row] * a.r[row] + beta * pc[row];
                    }
                    __syncthreads();
                    if (threadIdx.x == 0)
                        red_release_sys_add(a.comm.halo_flag(peer, a.comm.rank), 1ull);
                }
                ++epoch;
            }
            long long j = (long long)blockIdx.x * T + threadIdx.x;
            for (; j + vstride < a.n2; j += 2 * vstride)
            {
                double2 rv[2], dv[2], pv[2];
#pragma unroll
                for (int u = 0; u < 2; ++u)
                {
                    rv[u] = ld2(a.r, j + u * vstride);
                    dv[u] = ld2(a.dinv, j + u * vstride);
                    pv[u] = ld2(pc, j + u * vstride);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u)
                {
                    double2 o;
                    o.x = dv[u].x * rv[u].x + beta * pv[u].x;
                    o.y = dv[u].y * rv[u].y + beta * pv[u].y;
                    st2(pn, j + u * vstride, o);
                }
            }
            if (j < a.n2)
            {
                const double2 rv = ld2(a.r, j), dv = ld2(a.dinv, j), pv = ld2(pc, j);
                double2 o;
                o.x = dv.x * rv.x + beta * pv.x;
                o.y = dv.y * rv.y + beta * pv.y;
                st2(pn, j, o);
            }
        }
        lap(4);
        double acc0[1] = {0};
        ++m;
        ok = persist_sync<0>(acc0, a, m, 0, FinCgDirEigen{st});
        lap(5);
    }

    // drain the TMA copies that were issued ahead but never consumed (a CTA must not exit with copies in flight)
    {
        const long long issued = total_seq < consumed + STAGES ? total_seq : consumed + STAGES;
        for (long long s = consumed; s < issued; ++s)
            mbar_wait(&bar[(int)(s % STAGES)], (unsigned)((s / STAGES) & 1));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.timing)
        for (int i = 0; i < 6; ++i)
            atomicAdd((unsigned long long *)a.timing + i, (unsigned long long)tph[i]);
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.comm.world > 1)
    {
        *a.comm.red_seq = seq;
        *a.comm.push_epoch = epoch;
    }
}

// ------------------------------------------------------------------------------------------ host side
int Solver::persist_grid()
{
    static int per_sm_dev[16] = {0};
    int &per_sm = per_sm_dev[device & 15];
    if (!per_sm)
    {
        PSB_CUDA(cudaFuncSetAttribute(cg_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PCfg::bytes));
        PSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_persist_kernel, PCfg::threads, PCfg::bytes));
        per_sm = std::max(1, std::min(per_sm, kPersistCtasPerSm));
    }
    int sms = kSMs;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const long long need = std::max<long long>(1, (n + PCfg::threads - 1) / PCfg::threads);
    return (int)std::min<long long>(need, (long long)sms * per_sm);
}

void Solver::persist_reset()
{
    gbar.alloc(16, false);
    PSB_CUDA(cudaMemsetAsync(gbar.p, 0, 16 * sizeof(unsigned long long), ctx.stream));
}

// one cooperative launch = `iters` iterations starting from the direction in p_cur
void Solver::launch_cg_persist(double *p_cur, double *p_other, int iters)
{
    PersistArgs a{};
    a.A = A.view();
    a.order = a.A.tile_order;
    a.n_interior = a.A.tile_order ? a.A.n_interior : 0; // no order: every tile may touch halo columns, wait before the first
    a.push = PushList{nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    a.comm = ctx.comm;
    if (dist)
    {
        a.push = make_push(*dist);
    }
    else
        a.comm.world = 1;
    a.x = vx.p;
    a.r = vr.p;
    a.q = vq.p;
    a.p0 = p_cur;
    a.p1 = p_other;
    a.dinv = dinv.p;
    a.n2 = n_pad / 2;
    a.st = d_state;
    a.partials = ctx.partials.p;
    a.pstride = kMaxBlocks;
    a.bar = reinterpret_cast<GridBar *>(gbar.p);
    a.iters = iters;
    a.timing = reinterpret_cast<long long *>(gbar.p + 4);
    const int grid = persist_grid();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(PCfg::threads);
    cfg.dynamicSmemBytes = PCfg::bytes;
    cfg.stream = ctx.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ctx.prof_begin("cg_persist");
    PSB_CUDA(cudaLaunchKernelEx(&cfg, cg_persist_kernel, a));
    ctx.prof_end();
}

void Solver::persist_collect()
{
    unsigned long long h[6];
    PSB_CUDA(cudaMemcpy(h, gbar.p + 4, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 6; ++i)
        persist_cycles[i] = (double)h[i];
}

} // namespace psb
