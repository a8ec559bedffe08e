// Row-partitioned multi-GPU Jacobi-PCG: one process per GPU, NVLink peer memory for everything that
// crosses ranks. There is no reference counterpart (the reference is single-device, MASSolver.cu:193);
// the partition / halo index arrays are checked bit-exactly against the host oracle.
//
//  * partition : contiguous row ranges balanced by nnz (offsets[g] = first row r with row_ptr[r] >= g nnz / world)
//  * halo      : the owner PUSHES: the direction-update kernel recomputes the boundary entries of the new
//                search direction and stores them straight into the consumers' comm buffers (st.global on
//                IPC-mapped peer pointers), then raises a per-source epoch flag. The SpMV of the next
//                iteration waits on the flags of its neighbours while its TMA prefetch is already running.
//  * dots      : the last CTA of every reducing kernel writes its totals into every peer's slot and sums
//                the `world` slots in rank order (comm_allreduce, common.cuh) -- no extra launch, no NCCL.
#include "../../include/psb200.h"
#include "dist.hpp"
#include "capi_internal.hpp"
#include "solver.hpp"
#include "amg.hpp"

#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstring>
#include <sstream>

namespace psb {

// ====================================================================================== host plan
void DistPlanHost::build(long long n, long long nnz, const int *outer, const int *inner, int rank_, int world_, long long halo_cap_, int align)
{
    if (align < 1 || n % align != 0)
        throw std::invalid_argument("psb200 dist: the matrix size is not a multiple of the block size");
    if (world_ < 1 || world_ > kMaxRanks || rank_ < 0 || rank_ >= world_)
        throw std::invalid_argument("psb200 dist: rank/world out of range (world <= 8)");
    rank = rank_;
    world = world_;
    n_global = n;
    nnz_global = nnz;
    halo_cap = halo_cap_;
    // CSR row pointer of the whole matrix (counting pass over the CSC row indices)
    std::vector<int> row_ptr(n + 1, 0);
    for (long long k = 0; k < nnz; ++k)
    {
        const int i = inner[k];
        if (i < 0 || i >= n)
            throw std::invalid_argument("psb200 dist: inner index out of range");
        row_ptr[i + 1]++;
    }
    for (long long i = 0; i < n; ++i)
        row_ptr[i + 1] += row_ptr[i];
    // contiguous ranges balanced by nnz
    offsets.assign(world + 1, 0);
    for (int g = 1; g < world; ++g)
    {
        const long long target = (long long)(((__int128)nnz * g) / world);
        long long r = std::lower_bound(row_ptr.begin(), row_ptr.end(), (int)std::min<long long>(target, 0x7fffffff)) - row_ptr.begin();
        r = ((r + align - 1) / align) * align;
        r = std::min(r, n);
        r = std::max(r, offsets[g - 1]);
        offsets[g] = r;
    }
    offsets[world] = n;
    const long long a = offsets[rank], b = offsets[rank + 1];
    const int nl = (int)(b - a);
    auto owner = [&](int i) {
        int q = 0;
        while (i >= offsets[q + 1])
            ++q;
        return q;
    };
    // local rows by a counting transpose restricted to [a, b): columns come out ascending
    rp.assign(nl + 1, 0);
    for (int i = 0; i <= nl; ++i)
        rp[i] = row_ptr[a + i] - row_ptr[a];
    const int lnnz = rp[nl];
    std::vector<int> gcol(lnnz);
    perm.assign(lnnz, 0);
    {
        std::vector<int> cur(rp.begin(), rp.end() - 1);
        for (long long c = 0; c < n; ++c)
            for (int k = outer[c]; k < outer[c + 1]; ++k)
            {
                const int i = inner[k];
                if (i >= a && i < b)
                {
                    const int pos = cur[i - a]++;
                    gcol[pos] = (int)c;
                    perm[pos] = k;
                }
            }
    }
    // halo columns: distinct off-range columns, ascending (=> grouped by owner)
    halo_cols.clear();
    for (int k = 0; k < lnnz; ++k)
        if (gcol[k] < a || gcol[k] >= b)
            halo_cols.push_back(gcol[k]);
    std::sort(halo_cols.begin(), halo_cols.end());
    halo_cols.erase(std::unique(halo_cols.begin(), halo_cols.end()), halo_cols.end());
    recv_count.assign(world, 0);
    std::vector<int> seg_start(world + 1, 0);
    for (int c : halo_cols)
        recv_count[owner(c)]++;
    for (int q = 0; q < world; ++q)
    {
        seg_start[q + 1] = seg_start[q] + recv_count[q];
        if (recv_count[q] > halo_cap)
            throw std::runtime_error("psb200 dist: halo from rank " + std::to_string(q) + " needs " + std::to_string(recv_count[q]) +
                                     " values, capacity is " + std::to_string(halo_cap) + " (raise halo_cap in psb200_dist_prepare)");
    }
    if ((long long)nl + (long long)world * halo_cap > 0x7fffffffLL)
        throw std::runtime_error("psb200 dist: local rows + halo regions exceed the int32 column range");
    // remap columns
    ci.assign(lnnz, 0);
    for (int k = 0; k < lnnz; ++k)
    {
        const int c = gcol[k];
        if (c >= a && c < b)
            ci[k] = (int)(c - a);
        else
        {
            const int q = owner(c);
            const int pos = (int)(std::lower_bound(halo_cols.begin(), halo_cols.end(), c) - halo_cols.begin()) - seg_start[q];
            ci[k] = (int)(nl + (long long)q * halo_cap + pos);
        }
    }
    // send lists: my column j is needed by rank q iff column j has a row owned by q (CSC gives this directly)
    std::vector<std::vector<int>> send(world);
    for (long long j = a; j < b; ++j)
        for (int k = outer[j]; k < outer[j + 1]; ++k)
        {
            const int q = owner(inner[k]);
            if (q != rank && (send[q].empty() || send[q].back() != (int)(j - a)))
                send[q].push_back((int)(j - a)); // all rows of column j are visited consecutively, so back() dedups
        }
    send_begin.assign(world + 1, 0);
    send_rows.clear();
    for (int q = 0; q < world; ++q)
    {
        send_begin[q + 1] = send_begin[q] + (int)send[q].size();
        send_rows.insert(send_rows.end(), send[q].begin(), send[q].end());
    }
}

DistState::~DistState()
{
    for (int q = 0; q < world; ++q)
        if (q != rank && peer[q])
            cudaIpcCloseMemHandle(peer[q]);
    if (comm_buf)
        cudaFree(comm_buf);
    if (counters)
        cudaFree(counters);
}

// ====================================================================================== kernels
// The push part shared by the kernels below: value(row) -> halo region of the consumer. The send list is cut
// into chunks of kPushChunk entries per destination; a CTA stores a chunk and then adds 1 to the consumer's
// flag with release semantics (no grid-level ticket, no second fence). push_no = number of this push (1-based).
template <class ValueFn>
__device__ __forceinline__ void push_section(const PushList &pl, const CommDev &c, unsigned long long push_no, int first_block, int nblocks,
                                             ValueFn value)
{
    const int par = (int)(push_no & 1);
    for (int ch = (int)blockIdx.x - first_block; ch < pl.nchunks; ch += nblocks)
    {
        const int peer = pl.chunk_peer[ch], start = pl.chunk_start[ch], cnt = pl.chunk_cnt[ch];
        double *dst = c.halo(peer, par, c.rank) + pl.chunk_off[ch];
        for (int e = threadIdx.x; e < cnt; e += blockDim.x)
            dst[e] = value(pl.rows[start + e]);
        __syncthreads();
        if (threadIdx.x == 0)
            red_release_sys_add(c.halo_flag(peer, c.rank), 1ull);
    }
}

// p_new = dinv r + beta p_old (Eigen CG direction update) with the halo push fused in:
// CTAs [0, push_blocks) recompute the boundary entries and store them into the neighbours' halo regions
// (scheduled first, so the halo is on the wire while the bulk of the vector is still being updated);
// CTAs [push_blocks, grid) update the local vector. p_new != p_old (ping-pong), so the two never race.
template <bool FIRST, int THREADS>
__global__ void __launch_bounds__(THREADS) cg_dir_dist_kernel(long long n2, double *__restrict__ p_new, const double *__restrict__ p_old,
                                                              const double *__restrict__ r, const double *__restrict__ dinv, KState *st,
                                                              RedCtx rc, PushList pl, int vec_blocks, const int *done)
{
    griddep_launch_dependents();
    griddep_wait();
    if (done && *done)
        return;
    const double beta = FIRST ? 0.0 : st->rz_new / st->rz;
    const int push_blocks = (int)gridDim.x - vec_blocks;
    const unsigned long long push_no = *rc.comm.push_epoch + 1;
    if ((int)blockIdx.x >= push_blocks)
    {
        const long long stride = (long long)vec_blocks * THREADS;
        for (long long j = (long long)(blockIdx.x - push_blocks) * THREADS + threadIdx.x; j < n2; j += stride)
        {
            const double2 rv = ld2(r, j), dv = ld2(dinv, j);
            double2 o;
            o.x = dv.x * rv.x;
            o.y = dv.y * rv.y;
            if (!FIRST)
            {
                const double2 pv = ld2(p_old, j);
                o.x += beta * pv.x;
                o.y += beta * pv.y;
            }
            st2(p_new, j, o);
        }
    }
    else
        push_section(pl, rc.comm, push_no, 0, push_blocks, [&](int row) { return FIRST ? dinv[row] * r[row] : dinv[row] * r[row] + beta * p_old[row]; });
    double acc[1] = {0}, tot[1];
    if (grid_reduce<0, THREADS>(acc, rc, tot) && threadIdx.x == 0)
    {
        if (push_blocks > 0)
            *rc.comm.push_epoch = push_no;
        if (!FIRST)
        {
            // FinCgDirEigen
            st->rz = st->rz_new;
            st->iter += 1;
            if (st->iter >= st->max_iter)
            {
                st->done = 1;
                st->status = ST_MAXITER;
            }
        }
    }
}

// push the boundary entries of an arbitrary local vector (initial guess x0)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) halo_push_kernel(const double *__restrict__ v, RedCtx rc, PushList pl, const int *done)
{
    griddep_launch_dependents();
    griddep_wait();
    if (done && *done)
        return;
    const unsigned long long push_no = *rc.comm.push_epoch + 1;
    push_section(pl, rc.comm, push_no, 0, (int)gridDim.x, [&](int row) { return v[row]; });
    double acc[1] = {0}, tot[1];
    if (grid_reduce<0, THREADS>(acc, rc, tot) && threadIdx.x == 0)
        *rc.comm.push_epoch = push_no;
}

// ====================================================================================== Solver (dist mode)
void Solver::dist_prepare(int rank, int world, long long halo_cap, char handle_out[64])
{
    if (analyzed)
        throw std::runtime_error("psb200_dist_prepare: call before analyze_pattern");
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world)
        throw std::invalid_argument("psb200_dist_prepare: rank/world out of range (world <= 8)");
    if (halo_cap <= 0)
        halo_cap = 1 << 20;
    ensure_ctx(*this);
    dist = std::make_unique<DistState>();
    DistState &d = *dist;
    d.rank = rank;
    d.world = world;
    d.halo_cap = (halo_cap + 1) & ~1ll;
    d.comm_bytes = kCommHaloOff + sizeof(double) * 4 * kMaxRanks * (size_t)d.halo_cap; // halo + bulk regions, 2 parities each
    PSB_CUDA(cudaMalloc(&d.comm_buf, d.comm_bytes));
    PSB_CUDA(cudaMemset(d.comm_buf, 0, d.comm_bytes));
    PSB_CUDA(cudaMalloc(&d.counters, 256));
    PSB_CUDA(cudaMemset(d.counters, 0, 256));
    cudaIpcMemHandle_t hnd;
    PSB_CUDA(cudaIpcGetMemHandle(&hnd, d.comm_buf));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::memcpy(handle_out, &hnd, 64);
    PSB_CUDA(cudaDeviceSynchronize());
}

void Solver::dist_connect(const char *handles)
{
    if (!dist)
        throw std::runtime_error("psb200_dist_connect: psb200_dist_prepare first");
    DistState &d = *dist;
    for (int q = 0; q < d.world; ++q)
    {
        if (q == d.rank)
        {
            d.peer[q] = d.comm_buf;
            continue;
        }
        cudaIpcMemHandle_t hnd;
        std::memcpy(&hnd, handles + 64 * q, 64);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            throw CudaError(std::string("psb200_dist_connect: cudaIpcOpenMemHandle(rank ") + std::to_string(q) + "): " + cudaGetErrorString(e));
        d.peer[q] = p;
    }
    CommDev c;
    c.world = d.world;
    c.rank = d.rank;
    c.halo_cap = d.halo_cap;
    for (int q = 0; q < kMaxRanks; ++q)
        c.peer[q] = (unsigned char *)(q < d.world ? d.peer[q] : nullptr);
    c.red_seq = d.counters;
    c.push_epoch = d.counters + 1;
    c.error = (int *)(d.counters + 2);
    c.bulk_epoch = d.counters + 3;
    c.bulk_expect = d.counters + 8;
    for (int q = 0; q < kMaxRanks; ++q)
        c.in_chunks[q] = 0;
    ctx.comm = c;
    d.connected = true;
}

void Solver::analyze_pattern_dist(long long n_, long long nnz_, const int *outer, const int *inner)
{
    DistState &d = *dist;
    if (!d.connected && d.world > 1)
        throw std::runtime_error("psb200 dist: psb200_dist_connect has not been called");
    const int B = std::max(1, prm.block_size);
    if (B > 3)
        throw std::invalid_argument("psb200: block_size must be 1, 2 or 3 (reference AMGCL.cpp:111-123)");
    if (prm.krylov != "cg")
        throw std::runtime_error("psb200 dist: the row-partitioned path provides PCG (krylov=cg) with precond = jacobi | none | amg");
    d.plan.build(n_, nnz_, outer, inner, d.rank, d.world, d.halo_cap, B);
    d.A_diag.n = 0; // new pattern: the rank-local diagonal block is rebuilt at the next AMG factorize
    pattern_block = B; // the Krylov loop works on the scalar rows; the rank-local AMG sees B x B blocks (build_diag_block_dist)
    const DistPlanHost &P = d.plan;
    cudaStream_t st = ctx.stream;
    n = P.r1() - P.r0();
    nnz = (long long)P.ci.size();
    n_pad = (n + 3) & ~3ll;
    sym_pattern = false;
    A.n = (int)n;
    A.ncols = (int)n;
    A.nnz = nnz;
    A.nl = (int)n;
    A.rp.alloc(n + 1);
    A.ci.alloc(std::max<long long>(nnz, 1), false, 64);
    A.va.alloc(std::max<long long>(nnz, 1), false, 64);
    PSB_CUDA(cudaMemcpyAsync(A.rp.p, P.rp.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, st));
    if (nnz)
        PSB_CUDA(cudaMemcpyAsync(A.ci.p, P.ci.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice, st));
    d.recv_mask = 0;
    d.send_mask = 0;
    for (int q = 0; q < d.world; ++q)
    {
        if (P.recv_count[q] > 0)
            d.recv_mask |= 1u << q;
        if (P.send_begin[q + 1] > P.send_begin[q])
            d.send_mask |= 1u << q;
    }
    A.halo_mask = d.recv_mask;
    // device push list: the send rows of every destination cut into chunks of kPushChunk entries; the consumer
    // derives the number of chunks it will see from its own recv_count (same formula on both sides)
    d.n_push = (int)P.send_rows.size();
    std::vector<int> cpeer, cstart, ccnt, coff;
    for (int q = 0; q < d.world; ++q)
    {
        const int cnt = P.send_begin[q + 1] - P.send_begin[q];
        for (int o = 0; o < cnt; o += kPushChunk)
        {
            cpeer.push_back(q);
            cstart.push_back(P.send_begin[q] + o);
            ccnt.push_back(std::min(kPushChunk, cnt - o));
            coff.push_back(o);
        }
        ctx.comm.in_chunks[q] = (P.recv_count[q] + kPushChunk - 1) / kPushChunk;
    }
    d.n_chunks = (int)cpeer.size();
    d.push_rows.alloc(std::max(1, d.n_push));
    d.chunk_tab.alloc(std::max(1, 4 * d.n_chunks));
    if (d.n_push)
    {
        std::vector<int> tab;
        tab.insert(tab.end(), cpeer.begin(), cpeer.end());
        tab.insert(tab.end(), cstart.begin(), cstart.end());
        tab.insert(tab.end(), ccnt.begin(), ccnt.end());
        tab.insert(tab.end(), coff.begin(), coff.end());
        PSB_CUDA(cudaMemcpyAsync(d.push_rows.p, P.send_rows.data(), sizeof(int) * d.n_push, cudaMemcpyHostToDevice, st));
        PSB_CUDA(cudaMemcpyAsync(d.chunk_tab.p, tab.data(), sizeof(int) * tab.size(), cudaMemcpyHostToDevice, st));
        PSB_CUDA(cudaStreamSynchronize(st)); // tab is a stack-scoped staging vector
    }
    // values window of this rank inside the CSC value array (banded matrices: ~ the local share)
    d.val_lo = 0;
    d.val_hi = 0;
    if (nnz)
    {
        int lo = P.perm[0], hi = P.perm[0];
        for (int v : P.perm)
        {
            lo = std::min(lo, v);
            hi = std::max(hi, v);
        }
        d.val_lo = lo;
        d.val_hi = (long long)hi + 1;
        std::vector<int> rel(P.perm.size());
        for (size_t k = 0; k < rel.size(); ++k)
            rel[k] = P.perm[k] - lo;
        d.d_perm.alloc(nnz);
        PSB_CUDA(cudaMemcpyAsync(d.d_perm.p, rel.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice, st));
        PSB_CUDA(cudaStreamSynchronize(st)); // rel is a stack-scoped staging vector
    }
    PSB_CUDA(cudaStreamSynchronize(st));
    A.plan(prm.spmv_kernel, st);
    // interior-first tile order of the stream schedule: tiles that touch no halo column come first, so the SpMV
    // multiplies them while the neighbours' pushes are still on the wire (both the split and the persistent path)
    if (A.kind == SPMV_STREAM)
    {
        const int T = A.stream_rows();
        const int ntiles = (int)((n + T - 1) / T);
        std::vector<int> order, boundary;
        order.reserve(ntiles);
        for (int t = 0; t < ntiles; ++t)
        {
            const int k0 = P.rp[(size_t)t * T], k1 = P.rp[std::min<long long>(n, (long long)(t + 1) * T)];
            bool halo = false;
            for (int k = k0; k < k1 && !halo; ++k)
                halo = P.ci[k] >= (int)n;
            (halo ? boundary : order).push_back(t);
        }
        A.n_interior = (int)order.size();
        A.order_rows = T;
        order.insert(order.end(), boundary.begin(), boundary.end());
        A.tile_order.alloc(std::max(1, ntiles));
        if (ntiles)
            PSB_CUDA(cudaMemcpyAsync(A.tile_order.p, order.data(), sizeof(int) * ntiles, cudaMemcpyHostToDevice, st));
        PSB_CUDA(cudaStreamSynchronize(st)); // order is a stack-scoped staging vector
    }
    PSB_CUDA(cudaStreamSynchronize(st));
}

__global__ void gather_window_kernel(long long n, const double *__restrict__ src, const int *__restrict__ perm, double *__restrict__ dst)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n)
        dst[k] = src[perm[k]];
}

// values only: one contiguous H2D copy of this rank's CSC window + a device gather (no host-side gather)
void Solver::factorize_values_dist(const double *vals)
{
    DistState &d = *dist;
    if (nnz)
    {
        const long long w = d.val_hi - d.val_lo;
        d.d_csc_window.alloc((size_t)w);
        PSB_CUDA(cudaMemcpyAsync(d.d_csc_window.p, vals + d.val_lo, sizeof(double) * w, cudaMemcpyHostToDevice, ctx.stream));
        gather_window_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, ctx.stream>>>(nnz, d.d_csc_window.p, d.d_perm.p, A.va.p);
        check_launch();
    }
    PSB_CUDA(cudaStreamSynchronize(ctx.stream));
}

void Solver::check_comm_error()
{
    if (!dist)
        return;
    int e = 0;
    PSB_CUDA(cudaMemcpy(&e, dist->counters + 2, sizeof(int), cudaMemcpyDeviceToHost));
    if (e)
    {
        PSB_CUDA(cudaMemset(dist->counters + 2, 0, sizeof(int)));
        throw std::runtime_error("psb200 dist: a peer did not answer within the spin limit (lost rank or mismatched call sequence)");
    }
}

PushList make_push(DistState &d)
{
    const int *t = d.chunk_tab.p;
    const int nc = d.n_chunks;
    return PushList{d.push_rows.p, t, t + nc, t + 2 * nc, t + 3 * nc, nc};
}
// CTAs that push: one per chunk up to 128 (a CTA loops over chunks beyond that); 0 on a single rank. Every rank
// of a multi-rank run launches at least one so the epochs advance in lockstep.
static int push_ctas(const DistState &d) { return d.world > 1 ? std::max(1, std::min(128, d.n_chunks)) : 0; }

void Solver::push_halo_of(const double *d_v, const int *done)
{
    DistState &d = *dist;
    const int push_blocks = push_ctas(d);
    if (!push_blocks)
        return;
    launch_chain(ctx, halo_push_kernel<kVecThreads>, push_blocks, kVecThreads, 0, d_v, ctx.red(), make_push(d), done);
    check_launch();
}

// Sum of a vector across the ranks (the restriction of the distributed AMG cycle): every rank stores its partial
// into region [parity][rank] of EVERY rank's bulk area (chunks of kPushChunk entries, one release-add per chunk on the
// consumer's bulk flag), waits until all sources have delivered this segment and adds the `world` regions in rank order,
// so every rank obtains the bit-identical sum. One kernel; a CTA first pushes its chunks, then waits, then sums them.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) bulk_allreduce_kernel(const double *__restrict__ partial, double *__restrict__ out, int len,
                                                                 RedCtx rc, const int *done)
{
    if (done && *done)
        return;
    const CommDev &c = rc.comm;
    const int nchunks = (len + kPushChunk - 1) / kPushChunk;
    const unsigned long long epoch = *c.bulk_epoch + 1;
    const int par = (int)(epoch & 1);
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x)
    {
        const int off = ch * kPushChunk, cnt = min(kPushChunk, len - off);
        for (int e = threadIdx.x; e < cnt; e += THREADS)
        {
            const double v = partial[off + e];
            for (int q = 0; q < c.world; ++q)
                c.bulk(q, par, c.rank)[off + e] = v;
        }
        __syncthreads();
        if ((int)threadIdx.x < c.world)
            red_release_sys_add(c.bulk_flag(threadIdx.x, c.rank), 1ull);
    }
    if ((int)threadIdx.x < c.world)
    {
        if (!spin_ge(c.bulk_flag(c.rank, threadIdx.x), c.bulk_expect[threadIdx.x] + (unsigned long long)nchunks, c.error))
            *c.error = 1;
        fence_acq_rel_sys();
    }
    __syncthreads();
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x)
    {
        const int off = ch * kPushChunk, cnt = min(kPushChunk, len - off);
        for (int e = threadIdx.x; e < cnt; e += THREADS)
        {
            double s = 0;
            for (int q = 0; q < c.world; ++q)
                s += __ldcg(c.bulk(c.rank, par, q) + off + e);
            out[off + e] = s;
        }
    }
    double acc[1] = {0}, tot[1];
    if (grid_reduce<0, THREADS>(acc, rc, tot) && (int)threadIdx.x < c.world)
    {
        c.bulk_expect[threadIdx.x] += (unsigned long long)nchunks;
        if (threadIdx.x == 0)
            *c.bulk_epoch = epoch;
    }
}

// out = sum over ranks of partial (length len, the same on every rank); segments of at most halo_cap entries
void Solver::bulk_allreduce(const double *d_partial, double *d_out, long long len, const int *done)
{
    DistState &d = *dist;
    if (d.world == 1)
    {
        if (d_out != d_partial)
            PSB_CUDA(cudaMemcpyAsync(d_out, d_partial, sizeof(double) * len, cudaMemcpyDeviceToDevice, ctx.stream));
        return;
    }
    for (long long off = 0; off < len; off += d.halo_cap)
    {
        const int seg = (int)std::min<long long>(d.halo_cap, len - off);
        const int nchunks = (seg + kPushChunk - 1) / kPushChunk;
        const int grid = std::max(1, std::min(nchunks, 2 * kSMs));
        ctx.prof_begin("bulk_allreduce");
        bulk_allreduce_kernel<kVecThreads><<<grid, kVecThreads, 0, ctx.stream>>>(d_partial + off, d_out + off, seg, ctx.red(), done);
        check_launch();
        ctx.prof_end();
    }
}

// ---------------------------------------------------------------------------------- rank-local AMG
__global__ void diag_count_kernel(int n, int nl, const int *__restrict__ rp, const int *__restrict__ ci, int *__restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n)
        return;
    int c = 0;
    if (i < n)
        for (int k = rp[i]; k < rp[i + 1]; ++k)
            c += ci[k] < nl;
    cnt[i] = c;
}
__global__ void diag_fill_kernel(int n, int nl, const int *__restrict__ rp, const int *__restrict__ ci, const int *__restrict__ drp,
                                 int *__restrict__ dci, int *__restrict__ src)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    int o = drp[i];
    for (int k = rp[i]; k < rp[i + 1]; ++k)
        if (ci[k] < nl)
        {
            dci[o] = ci[k];
            src[o] = k;
            ++o;
        }
}
__global__ void diag_vals_kernel(long long nnz, const double *__restrict__ va, const int *__restrict__ src, double *__restrict__ out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz)
        out[k] = src[k] >= 0 ? va[src[k]] : 0.0;
}

// A_diag = A[local rows, local columns]: pattern once per analysis, values on every factorize
void Solver::build_diag_block_dist()
{
    DistState &d = *dist;
    cudaStream_t st = ctx.stream;
    CsrDev &D = d.A_diag;
    if (D.n != (int)n || D.rp.p == nullptr || d.diag_src.n == 0)
    {
        DevBuf<int> cnt;
        cnt.alloc((size_t)n + 1, true);
        D.n = (int)n;
        D.ncols = (int)n;
        D.rp.alloc((size_t)n + 1);
        const unsigned blocks = (unsigned)((n + 256) / 256);
        diag_count_kernel<<<blocks, 256, 0, st>>>((int)n, (int)n, A.rp.p, A.ci.p, cnt.p);
        size_t bytes = 0;
        PSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt.p, D.rp.p, (int)n + 1, st));
        DevBuf<unsigned char> tmp;
        tmp.alloc(bytes);
        PSB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, cnt.p, D.rp.p, (int)n + 1, st));
        int dn = 0;
        PSB_CUDA(cudaMemcpyAsync(&dn, D.rp.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
        PSB_CUDA(cudaStreamSynchronize(st));
        D.nnz = dn;
        D.ci.alloc(std::max(1, dn), false, 64);
        D.va.alloc(std::max(1, dn), false, 64);
        d.diag_src.alloc(std::max(1, dn));
        diag_fill_kernel<<<blocks, 256, 0, st>>>((int)n, (int)n, A.rp.p, A.ci.p, D.rp.p, D.ci.p, d.diag_src.p);
        check_launch();
        if (pattern_block > 1 && n > 0)
        {
            // block mode: the B rows of a node get the full B x B block pattern (diag_src = -1 marks the fill-in),
            // the invariant the block AMG kernels rely on (same expansion as the single-GPU analyze_pattern)
            D.nnz = expand_block_pattern(ctx, pattern_block, n, D.rp, D.ci, d.diag_src);
            D.va.alloc(std::max<long long>(1, D.nnz), false, 64);
        }
        PSB_CUDA(cudaStreamSynchronize(st));
        D.plan("auto", st);
    }
    if (D.nnz)
        diag_vals_kernel<<<(unsigned)((D.nnz + 255) / 256), 256, 0, st>>>(D.nnz, A.va.p, d.diag_src.p, D.va.p);
    check_launch();
    PSB_CUDA(cudaStreamSynchronize(st));
}

// p_new = s + beta p_old (amgcl cg direction update) with the halo push fused in (same layout as
// cg_dir_dist_kernel: push CTAs first).
template <int THREADS>
__global__ void __launch_bounds__(THREADS) cg_dir_amgcl_dist_kernel(long long n2, double *__restrict__ p_new, const double *__restrict__ p_old,
                                                                    const double *__restrict__ s, const KState *st, RedCtx rc, PushList pl,
                                                                    int vec_blocks, const int *done)
{
    griddep_launch_dependents();
    griddep_wait();
    if (done && *done)
        return;
    const double beta = st->iter ? st->rho / st->rho_old : 0.0;
    const int push_blocks = (int)gridDim.x - vec_blocks;
    const unsigned long long push_no = *rc.comm.push_epoch + 1;
    if ((int)blockIdx.x >= push_blocks)
    {
        const long long stride = (long long)vec_blocks * THREADS;
        for (long long j = (long long)(blockIdx.x - push_blocks) * THREADS + threadIdx.x; j < n2; j += stride)
        {
            const double2 sv = ld2(s, j), pv = ld2(p_old, j);
            double2 o;
            o.x = sv.x + (beta != 0.0 ? beta * pv.x : 0.0);
            o.y = sv.y + (beta != 0.0 ? beta * pv.y : 0.0);
            st2(p_new, j, o);
        }
    }
    else
        push_section(pl, rc.comm, push_no, 0, push_blocks, [&](int row) { return s[row] + (beta != 0.0 ? beta * p_old[row] : 0.0); });
    double acc[1] = {0}, tot[1];
    if (grid_reduce<0, THREADS>(acc, rc, tot) && threadIdx.x == 0 && push_blocks > 0)
        *rc.comm.push_epoch = push_no;
}

// AMG-PCG on the row partition (amgcl cg ordering). The preconditioner is rank-local: every rank
// applies the SA-AMG cycle of its own diagonal block (block-Jacobi across ranks, no communication inside the
// cycle); the outer CG is the global one -- halo push of p, three fused all-reduces per iteration.
void Solver::run_cg_amgcl_dist(const double *d_b)
{
    if (!amg)
        throw std::runtime_error("psb200_solve: AMG hierarchy missing (factorize with precond=amg first)");
    DistState &d = *dist;
    KState *S = d_state;
    const int *done = &S->done;
    d.vp2.alloc((size_t)n_pad, true);
    init_state(*this, prm.tolerance, prm.max_iter);
    const long long n2 = n_pad / 2;
    const int vec_blocks = vec_grid(n2);
    const int push_blocks = push_ctas(d);
    PushList pl = make_push(d);
    RedCtx rc = ctx.red();
    PSB_CUDA(cudaMemsetAsync(vp.p, 0, sizeof(double) * n_pad, ctx.stream));
    if (push_blocks)
    {
        launch_chain(ctx, halo_push_kernel<kVecThreads>, push_blocks, kVecThreads, 0, vx.p, rc, pl, (const int *)nullptr);
        check_launch();
    }
    launch_spmv(ctx, "spmv_residual", A, vx.p, EpiResidualNorms{vr.p, d_b, dinv.p}, FinInitAmgcl{S});
    auto batch = [&]() {
        double *pc = vp.p, *pn = d.vp2.p;
        for (int i = 0; i < 2; ++i)
        {
            if (amg->has_dist_fine())
                amg->apply_dist(vr.p, vz.p, done); // level 0 partitioned (halo pushes, bulk all-reduce), coarse levels replicated
            else
            {
                LocalScope local(ctx);
                amg->apply(vr.p, vz.p, done);
            }
            launch_vec(ctx, "dot", n_pad, OpDot{vr.p, vz.p}, FinRhoAmgcl{S}, done);
            ctx.prof_begin("cg_dir");
            launch_chain(ctx, cg_dir_amgcl_dist_kernel<kVecThreads>, vec_blocks + push_blocks, kVecThreads, 0, n2, pn, pc, vz.p, S, rc, pl, vec_blocks, done);
            check_launch();
            ctx.prof_end();
            launch_spmv(ctx, "spmv_dot", A, pn, EpiDot{vq.p, pn}, FinPAp{S}, done);
            launch_vec(ctx, "cg_update", n_pad, OpCgUpdateAmgcl{vx.p, vr.p, pn, vq.p, S, 0.0}, FinCgUpdateAmgcl{S}, done);
            std::swap(pc, pn);
        }
    };
    std::ostringstream key;
    key << "cg_amgcl_dist/" << n << "/" << (void *)vx.p << "/" << (void *)amg.get() << "/" << (void *)d_b;
    drive(batch, 2, key.str());
    finish_solve();
    check_comm_error();
}

// Jacobi-PCG in Eigen's ordering on the row partition. Same kernels as the single-GPU path for the
// SpMV and the x/r update (their reductions all-reduce inside the kernel); the direction update is the
// fused update + halo push. p ping-pongs between vp and vp2.
void Solver::run_cg_eigen_dist(const double *d_b)
{
    DistState &d = *dist;
    KState *S = d_state;
    const int *done = &S->done;
    d.vp2.alloc((size_t)n_pad, false);
    init_state(*this, prm.tolerance, prm.max_iter);
    const long long n2 = n_pad / 2;
    const int vec_blocks = vec_grid(n2);
    // every rank pushes at every push point (even an empty list) so the epochs advance in lockstep
    const int push_blocks = push_ctas(d);
    PushList pl = make_push(d);
    RedCtx rc = ctx.red();
    if (push_blocks)
    {
        ctx.prof_begin("halo_push");
        launch_chain(ctx, halo_push_kernel<kVecThreads>, push_blocks, kVecThreads, 0, vx.p, rc, pl, (const int *)nullptr);
        check_launch();
        ctx.prof_end();
    }
    launch_spmv(ctx, "spmv_residual", A, vx.p, EpiResidualNorms{vr.p, d_b, dinv.p}, FinInitEigen{S});
    ctx.prof_begin("cg_dir");
    launch_chain(ctx, cg_dir_dist_kernel<true, kVecThreads>, vec_blocks + push_blocks, kVecThreads, 0, n2, vp.p, vp.p, vr.p, dinv.p, S, rc, pl, vec_blocks, done);
    check_launch();
    ctx.prof_end();
    const int batch_iters = std::max(2, prm.check_every & ~1);
    if (use_persist())
    {
        persist_reset();
        auto pbatch = [&]() { launch_cg_persist(vp.p, d.vp2.p, batch_iters); };
        std::ostringstream pkey;
        pkey << "cg_persist_dist/" << n << "/" << (void *)vx.p << "/" << (void *)A.va.p << "/" << (void *)d.vp2.p << "/" << batch_iters;
        drive(pbatch, batch_iters, pkey.str());
        finish_solve();
        persist_collect();
        check_comm_error();
        return;
    }
    auto batch = [&]() {
        double *pc = vp.p, *pn = d.vp2.p;
        for (int i = 0; i < batch_iters; ++i)
        {
            launch_spmv(ctx, "spmv_dot", A, pc, EpiDot{vq.p, pc}, FinPAp{S}, done);
            launch_vec(ctx, "cg_update", n_pad, OpCgUpdateEigen{vx.p, vr.p, pc, vq.p, dinv.p, S, 0.0}, FinCgUpdateEigen{S}, done);
            ctx.prof_begin("cg_dir");
            launch_chain(ctx, cg_dir_dist_kernel<false, kVecThreads>, vec_blocks + push_blocks, kVecThreads, 0, n2, pn, pc, vr.p, dinv.p, S, rc, pl, vec_blocks, done);
            check_launch();
            ctx.prof_end();
            std::swap(pc, pn);
        }
    };
    std::ostringstream key;
    key << "cg_eigen_dist/" << A.kind << "/" << n << "/" << (void *)vx.p << "/" << (void *)A.va.p << "/" << (void *)d_b << "/" << batch_iters;
    drive(batch, batch_iters, key.str());
    finish_solve();
    check_comm_error();
}

} // namespace psb

// ====================================================================================== C ABI
extern "C" {

int psb200_dist_prepare(psb200_handle h, int rank, int world, int64_t halo_cap, char handle_out[64])
{
    if (!h || !handle_out)
        return PSB200_ERR_INVALID;
    try
    {
        h->s.err.clear();
        psb::DeviceScope device_scope(h->s.device, h->s.ctx.stream != nullptr);
        psb::AllocScope alloc_scope(h->s.ctx.stream);
        h->s.dist_prepare(rank, world, halo_cap, handle_out);
        return PSB200_OK;
    }
    catch (const std::invalid_argument &e)
    {
        h->s.err = e.what();
        return PSB200_ERR_INVALID;
    }
    catch (const std::exception &e)
    {
        h->s.err = e.what();
        return PSB200_ERR_COMM;
    }
}

int psb200_dist_connect(psb200_handle h, const char *handles)
{
    if (!h || !handles)
        return PSB200_ERR_INVALID;
    try
    {
        h->s.err.clear();
        psb::DeviceScope device_scope(h->s.device, h->s.ctx.stream != nullptr);
        psb::AllocScope alloc_scope(h->s.ctx.stream);
        h->s.dist_connect(handles);
        return PSB200_OK;
    }
    catch (const std::exception &e)
    {
        h->s.err = e.what();
        return PSB200_ERR_COMM;
    }
}

int psb200_dist_local_range(psb200_handle h, int64_t *row_begin, int64_t *row_end)
{
    if (!h || !h->s.dist || !h->s.analyzed)
        return PSB200_ERR_INVALID;
    if (row_begin)
        *row_begin = h->s.dist->plan.r0();
    if (row_end)
        *row_end = h->s.dist->plan.r1();
    return PSB200_OK;
}

// Host-only (no GPU needed): the partition / halo plan of one rank. Arrays are caller-allocated:
// offsets[world+1], local_rp[n+1], local_ci[nnz], local_perm[nnz], send_begin[world+1], send_rows[n],
// recv_count[world], halo_cols[n]; counts[0..2] = {local rows, local nnz, halo columns}.
int psb200_dist_plan_host(int64_t n, int64_t nnz, const int32_t *outer, const int32_t *inner, int rank, int world, int64_t halo_cap,
                          int64_t *offsets, int64_t *counts, int32_t *local_rp, int32_t *local_ci, int32_t *local_perm,
                          int32_t *send_begin, int32_t *send_rows, int32_t *recv_count, int32_t *halo_cols)
{
    return psb200_dist_plan_host_aligned(n, nnz, outer, inner, rank, world, halo_cap, 1, offsets, counts, local_rp, local_ci, local_perm,
                                         send_begin, send_rows, recv_count, halo_cols);
}

// Same with the row offsets rounded up to multiples of `align` (block problems: align = block size).
int psb200_dist_plan_host_aligned(int64_t n, int64_t nnz, const int32_t *outer, const int32_t *inner, int rank, int world,
                                  int64_t halo_cap, int align, int64_t *offsets, int64_t *counts, int32_t *local_rp, int32_t *local_ci,
                                  int32_t *local_perm, int32_t *send_begin, int32_t *send_rows, int32_t *recv_count, int32_t *halo_cols)
{
    try
    {
        psb::DistPlanHost P;
        P.build(n, nnz, outer, inner, rank, world, halo_cap, align);
        std::copy(P.offsets.begin(), P.offsets.end(), offsets);
        counts[0] = P.r1() - P.r0();
        counts[1] = (int64_t)P.ci.size();
        counts[2] = (int64_t)P.halo_cols.size();
        std::copy(P.rp.begin(), P.rp.end(), local_rp);
        std::copy(P.ci.begin(), P.ci.end(), local_ci);
        std::copy(P.perm.begin(), P.perm.end(), local_perm);
        std::copy(P.send_begin.begin(), P.send_begin.end(), send_begin);
        std::copy(P.send_rows.begin(), P.send_rows.end(), send_rows);
        std::copy(P.recv_count.begin(), P.recv_count.end(), recv_count);
        std::copy(P.halo_cols.begin(), P.halo_cols.end(), halo_cols);
        return PSB200_OK;
    }
    catch (...)
    {
        return PSB200_ERR_INVALID;
    }
}

} // extern "C"
