// Row-partitioned multi-GPU Krylov loops: one process per GPU, NVLink peer memory for everything that
// crosses ranks. There is no reference counterpart (the reference is single-device, MASSolver.cu:193);
// the partition / halo index arrays are checked bit-exactly against the host oracle.
//
//  * partition : contiguous row ranges balanced by nnz (offsets[g] = first row r with row_ptr[r] >= g nnz / world)
//  * halo      : the owner PUSHES: the direction-update kernel recomputes the boundary entries of the new
//                search direction and stores them straight into the consumers' comm buffers (st.global on
//                IPC-mapped peer pointers), then raises per-chunk flags. The SpMV of the next iteration waits
//                on the flags of its neighbours while its TMA prefetch is already running (common.cuh, CommDev).
//  * dots      : the last CTA of every reducing kernel writes its totals into every peer's slot and sums
//                the `world` slots in rank order (comm_allreduce, common.cuh) -- no extra launch, no NCCL.
//  * cg1r      : single-reduction (Chronopoulos-Gear) PCG: two kernels and ONE all-reduce per iteration.
//  * setup     : host-level collectives (barrier, all-gather of 8 doubles, all-to-all-v through the staging arena)
//                for the distributed AMG setup (amg_dist.cu).
#include "../../include/psb200.h"
#include "dist.hpp"
#include "capi_internal.hpp"
#include "solver.hpp"
#include "amg.hpp"
#include "amg_dist.hpp"

#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstring>
#include <functional>
#include <sstream>
#include <thread>

namespace psb {

// ====================================================================================== host plan
void DistPlanHost::build(long long n, long long nnz, const int *outer, const int *inner, int rank_, int world_, long long halo_cap_, int align)
{
    if (align < 1 || n % align != 0)
        throw std::invalid_argument("psb200 dist: the matrix size is not a multiple of the block size");
    if (world_ < 1 || world_ > kMaxRanks || rank_ < 0 || rank_ >= world_)
        throw std::invalid_argument("psb200 dist: rank/world out of range (world <= 8)");
    // the plan indexes `inner` and its own arrays with these values: reject what is not a compressed-column pattern before
    // anything is allocated (Solver::analyze_pattern makes the same checks; the host-only plan entry point arrives here
    // directly)
    if (n < 0 || nnz < 0 || !outer || (!inner && nnz > 0) || n > 0x7fffffffLL - 1024 || nnz > 0x7fffffffLL - 1024)
        throw std::invalid_argument("psb200 dist: null, negative or int32-overflowing pattern arguments");
    if (outer[0] != 0 || outer[n] != nnz)
        throw std::invalid_argument("psb200 dist: matrix is not compressed (outer[0] != 0 or outer[n] != nnz)");
    for (long long c = 0; c < n; ++c)
        if (outer[c] > outer[c + 1] || outer[c] < 0)
            throw std::invalid_argument("psb200 dist: outer index array is not non-decreasing");
    rank = rank_;
    world = world_;
    n_global = n;
    nnz_global = nnz;
    halo_cap = halo_cap_;
    // Host threads for the two passes over all nnz entries (the plan of a 70 M-nnz matrix took 0.3-0.6 s on one thread,
    // more than the GPU needs for everything else in analyze_pattern). Column ranges are dealt in order, so the result
    // does not depend on the thread count.
    const unsigned T = (unsigned)std::max<long long>(1, std::min<long long>({(long long)std::thread::hardware_concurrency(), 16ll, nnz / (1 << 20) + 1}));
    auto col_range = [&](unsigned t, long long &c0, long long &c1) {
        // split the columns so that every thread gets about the same number of entries
        const long long k0 = nnz * t / T, k1 = nnz * (t + 1) / T;
        c0 = std::upper_bound(outer, outer + n + 1, (int)k0) - outer - 1;
        c1 = t + 1 == T ? n : std::upper_bound(outer, outer + n + 1, (int)k1) - outer - 1;
        c0 = std::max<long long>(0, std::min(c0, n));
        c1 = std::max(c0, std::min(c1, n));
        if (t == 0)
            c0 = 0;
    };
    auto run_threads = [&](const std::function<void(unsigned)> &fn) {
        std::vector<std::thread> th;
        std::vector<std::exception_ptr> errs(T);
        for (unsigned t = 1; t < T; ++t)
            th.emplace_back([&, t]() {
                try
                {
                    fn(t);
                }
                catch (...)
                {
                    errs[t] = std::current_exception();
                }
            });
        try
        {
            fn(0);
        }
        catch (...)
        {
            errs[0] = std::current_exception();
        }
        for (auto &x : th)
            x.join();
        for (auto &e : errs)
            if (e)
                std::rethrow_exception(e);
    };
    // CSR row pointer of the whole matrix (counting pass over the CSC row indices): per-thread histograms, summed
    std::vector<int> row_ptr(n + 1, 0);
    {
        std::vector<std::vector<int>> hist(T);
        run_threads([&](unsigned t) {
            long long c0, c1;
            col_range(t, c0, c1);
            std::vector<int> &h = hist[t];
            if (t > 0)
                h.assign(n + 1, 0);
            int *dst = t == 0 ? row_ptr.data() : h.data();
            for (long long k = outer[c0]; k < outer[c1]; ++k)
            {
                const int i = inner[k];
                if (i < 0 || i >= n)
                    throw std::invalid_argument("psb200 dist: inner index out of range");
                dst[i + 1]++;
            }
        });
        for (unsigned t = 1; t < T; ++t)
            for (long long i = 0; i <= n; ++i)
                row_ptr[i] += hist[t][i];
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
    // local rows by a counting transpose restricted to [a, b): columns come out ascending. Every thread first counts the
    // entries of its column range per local row, then writes them behind the entries of the lower ranges.
    rp.assign(nl + 1, 0);
    for (int i = 0; i <= nl; ++i)
        rp[i] = row_ptr[a + i] - row_ptr[a];
    const int lnnz = rp[nl];
    std::vector<int> gcol(lnnz);
    perm.assign(lnnz, 0);
    {
        std::vector<std::vector<int>> cnt(T);
        run_threads([&](unsigned t) {
            long long c0, c1;
            col_range(t, c0, c1);
            std::vector<int> &h = cnt[t];
            h.assign(nl, 0);
            for (long long k = outer[c0]; k < outer[c1]; ++k)
            {
                const int i = inner[k];
                if (i >= a && i < b)
                    h[i - a]++;
            }
        });
        // cnt[t][i] -> first position of thread t in local row i
        for (int i = 0; i < nl; ++i)
        {
            int pos = rp[i];
            for (unsigned t = 0; t < T; ++t)
            {
                const int c = cnt[t][i];
                cnt[t][i] = pos;
                pos += c;
            }
        }
        run_threads([&](unsigned t) {
            long long c0, c1;
            col_range(t, c0, c1);
            std::vector<int> &cur = cnt[t];
            for (long long c = c0; c < c1; ++c)
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
        });
    }
    // halo columns: distinct off-range columns, ascending (=> grouped by owner); block problems: whole nodes
    halo_cols.clear();
    for (int k = 0; k < lnnz; ++k)
        if (gcol[k] < a || gcol[k] >= b)
        {
            if (align == 1)
                halo_cols.push_back(gcol[k]);
            else
                for (int d = 0; d < align; ++d)
                    halo_cols.push_back(gcol[k] / align * align + d);
        }
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
    // send lists: my column j is needed by rank q iff column j has a row owned by q (CSC gives this directly); block
    // problems: a node travels whole, i.e. the union over its `align` columns
    std::vector<std::vector<int>> send(world);
    for (long long j0 = a; j0 < b; j0 += align)
    {
        unsigned need = 0;
        for (long long j = j0; j < j0 + align; ++j)
            for (int k = outer[j]; k < outer[j + 1]; ++k)
            {
                const int q = owner(inner[k]);
                if (q != rank)
                    need |= 1u << q;
            }
        for (int q = 0; q < world; ++q)
            if ((need >> q) & 1u)
                for (int d = 0; d < align; ++d)
                    send[q].push_back((int)(j0 - a) + d);
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

// ====================================================================================== halo plans
unsigned HaloPlan::mask() const
{
    unsigned m = 0;
    for (int q = 0; q < world; ++q)
    {
        if (q < (int)recv_count.size() && recv_count[q] > 0)
            m |= 1u << q;
        if (q + 1 < (int)send_begin.size() && send_begin[q + 1] > send_begin[q])
            m |= 1u << q;
    }
    return m;
}

void HaloPlan::finalize(unsigned nbr_mask, const CommDev &comm, cudaStream_t st)
{
    n_push = (int)send_rows.size();
    std::vector<int> cpeer, cstart, ccnt, coff;
    for (int q = 0; q < world; ++q)
    {
        in_chunks[q] = 0;
        if (!((nbr_mask >> q) & 1u))
        {
            if ((q < (int)recv_count.size() && recv_count[q] > 0) || (q + 1 < (int)send_begin.size() && send_begin[q + 1] > send_begin[q]))
                throw std::logic_error("psb200 dist: halo plan has a neighbour outside the neighbour mask");
            continue;
        }
        const int cnt = send_begin[q + 1] - send_begin[q];
        int o = 0;
        do
        {
            cpeer.push_back(q);
            cstart.push_back(send_begin[q] + o);
            ccnt.push_back(std::min(kPushChunk, cnt - o));
            coff.push_back(o);
            o += kPushChunk;
        } while (o < cnt);
        in_chunks[q] = std::max(1, (recv_count[q] + kPushChunk - 1) / kPushChunk);
    }
    n_chunks = (int)cpeer.size();
    push_rows.alloc(std::max(1, n_push));
    chunk_tab.alloc(std::max(1, 4 * n_chunks));
    std::vector<int> tab;
    tab.insert(tab.end(), cpeer.begin(), cpeer.end());
    tab.insert(tab.end(), cstart.begin(), cstart.end());
    tab.insert(tab.end(), ccnt.begin(), ccnt.end());
    tab.insert(tab.end(), coff.begin(), coff.end());
    if (n_push)
        PSB_CUDA(cudaMemcpyAsync(push_rows.p, send_rows.data(), sizeof(int) * n_push, cudaMemcpyHostToDevice, st));
    if (!tab.empty())
        PSB_CUDA(cudaMemcpyAsync(chunk_tab.p, tab.data(), sizeof(int) * tab.size(), cudaMemcpyHostToDevice, st));
    // ---- maps for the fused push: row -> (destination, position, chunk)
    std::vector<int> first_chunk(world + 1, 0);
    {
        int ch = 0;
        for (int q = 0; q < world; ++q)
        {
            first_chunk[q] = ch;
            while (ch < n_chunks && cpeer[ch] == q)
                ++ch;
        }
        first_chunk[world] = ch;
    }
    struct Slot
    {
        int row, peer, off, chunk;
    };
    std::vector<Slot> slots;
    slots.reserve(n_push);
    for (int q = 0; q < world; ++q)
        for (int e = send_begin[q]; e < send_begin[q + 1]; ++e)
        {
            const int off = e - send_begin[q];
            slots.push_back(Slot{send_rows[e], q, off, first_chunk[q] + off / kPushChunk});
        }
    std::stable_sort(slots.begin(), slots.end(), [](const Slot &a, const Slot &b) { return a.row < b.row; });
    std::vector<int> h_brow, h_bptr, h_chunk(slots.size());
    std::vector<unsigned long long> h_dst(slots.size()), h_flag(std::max(1, n_chunks));
    std::vector<unsigned> h_bits((size_t)(n_local + 31) / 32 + 1, 0u);
    for (size_t k = 0; k < slots.size(); ++k)
    {
        if (k == 0 || slots[k].row != slots[k - 1].row)
        {
            h_brow.push_back(slots[k].row);
            h_bptr.push_back((int)k);
            if (slots[k].row < 0 || slots[k].row >= n_local)
                throw std::logic_error("psb200 dist: send row outside the local range");
            h_bits[slots[k].row >> 5] |= 1u << (slots[k].row & 31);
        }
        h_dst[k] = (unsigned long long)(uintptr_t)(comm.halo(slots[k].peer, 0, comm.rank) + slots[k].off);
        h_chunk[k] = slots[k].chunk;
    }
    std::vector<unsigned long long> h_empty;
    for (int ch = 0; ch < n_chunks; ++ch)
    {
        h_flag[ch] = (unsigned long long)(uintptr_t)comm.halo_flag(cpeer[ch], comm.rank);
        if (ccnt[ch] == 0)
            h_empty.push_back(h_flag[ch]);
    }
    std::vector<int> h_prefix(h_bits.size(), 0);
    for (size_t w = 1; w < h_bits.size(); ++w)
        h_prefix[w] = h_prefix[w - 1] + __builtin_popcount(h_bits[w - 1]);
    bits_prefix.alloc(h_prefix.size());
    PSB_CUDA(cudaMemcpyAsync(bits_prefix.p, h_prefix.data(), sizeof(int) * h_prefix.size(), cudaMemcpyHostToDevice, st));
    n_empty = (int)h_empty.size();
    empty_flag.alloc(std::max(1, n_empty));
    if (n_empty)
        PSB_CUDA(cudaMemcpyAsync(empty_flag.p, h_empty.data(), sizeof(unsigned long long) * n_empty, cudaMemcpyHostToDevice, st));
    h_bptr.push_back((int)slots.size());
    n_brow = (int)h_brow.size();
    n_slots = (int)slots.size();
    buf_stride = (long long)kMaxRanks * comm.halo_cap;
    send_bits.alloc(h_bits.size());
    brow.alloc(std::max(1, n_brow));
    bptr.alloc((size_t)n_brow + 1);
    slot_chunk.alloc(std::max<size_t>(1, h_chunk.size()));
    slot_dst.alloc(std::max<size_t>(1, h_dst.size()));
    chunk_flag.alloc(h_flag.size());
    in_chunks_dev.alloc(kMaxRanks);
    chunk_done.alloc((size_t)n_chunks + 1, true); // counters restart with every finalize (a collective point of the setup)
    PSB_CUDA(cudaMemcpyAsync(send_bits.p, h_bits.data(), sizeof(unsigned) * h_bits.size(), cudaMemcpyHostToDevice, st));
    if (n_brow)
        PSB_CUDA(cudaMemcpyAsync(brow.p, h_brow.data(), sizeof(int) * n_brow, cudaMemcpyHostToDevice, st));
    PSB_CUDA(cudaMemcpyAsync(bptr.p, h_bptr.data(), sizeof(int) * h_bptr.size(), cudaMemcpyHostToDevice, st));
    if (!h_chunk.empty())
    {
        PSB_CUDA(cudaMemcpyAsync(slot_chunk.p, h_chunk.data(), sizeof(int) * h_chunk.size(), cudaMemcpyHostToDevice, st));
        PSB_CUDA(cudaMemcpyAsync(slot_dst.p, h_dst.data(), sizeof(unsigned long long) * h_dst.size(), cudaMemcpyHostToDevice, st));
    }
    PSB_CUDA(cudaMemcpyAsync(chunk_flag.p, h_flag.data(), sizeof(unsigned long long) * h_flag.size(), cudaMemcpyHostToDevice, st));
    PSB_CUDA(cudaMemcpyAsync(in_chunks_dev.p, in_chunks, sizeof(int) * kMaxRanks, cudaMemcpyHostToDevice, st));
    PSB_CUDA(cudaStreamSynchronize(st)); // staging vectors are stack-scoped
}

PushMap HaloPlan::push_map() const
{
    const int nc = n_chunks;
    return PushMap{send_bits.p,
                   bits_prefix.p,
                   brow.p,
                   bptr.p,
                   slot_chunk.p,
                   reinterpret_cast<double *const *>(slot_dst.p),
                   chunk_tab.p + 2 * nc,
                   reinterpret_cast<unsigned long long *const *>(chunk_flag.p),
                   reinterpret_cast<unsigned long long *const *>(empty_flag.p),
                   chunk_done.p,
                   chunk_done.p + nc,
                   in_chunks_dev.p,
                   buf_stride,
                   n_brow,
                   nc,
                   n_empty};
}

PushList HaloPlan::push() const
{
    const int *t = chunk_tab.p;
    const int nc = n_chunks;
    PushList pl{push_rows.p, t, t + nc, t + 2 * nc, t + 3 * nc, nc, {}};
    for (int q = 0; q < kMaxRanks; ++q)
        pl.in_chunks[q] = in_chunks[q];
    return pl;
}

// ====================================================================================== kernels
// The push part shared by the kernels below: value(row) -> halo region of the consumer. The send list is cut
// into chunks of kPushChunk entries per destination; a CTA stores a chunk and then adds 1 to the consumer's
// flag with release semantics (no grid-level ticket, no second fence). push_no = number of this push (1-based).
// A neighbour issues push e only after it has finished reading the parity buffer that push e + 1 overwrites, so a rank may
// store as soon as the neighbours' push e has landed here. Kernels whose previous epoch was consumed by an SpMV of this
// rank (the Krylov loops: the SpMV waited for exactly that) pass WAIT_PREV = false; the generic push (AMG smoother
// iterates, where two pushes can follow each other without a consumer in between) waits itself -- with relaxed loads only:
// nothing is read from the peers here, the wait merely orders this CTA's stores after the observation.
template <bool WAIT_PREV, class ValueFn>
__device__ __forceinline__ void push_section(const PushList &pl, const CommDev &c, unsigned long long push_no, int first_block, int nblocks,
                                             ValueFn value)
{
    if (WAIT_PREV)
    {
        if ((int)threadIdx.x < c.world && ((c.nbr_mask >> threadIdx.x) & 1u))
            if (!spin_ge(c.halo_flag(c.rank, threadIdx.x), c.halo_expect[threadIdx.x], c.error, c.spin_limit))
                *c.error = 1;
        __syncthreads();
    }
    const int par = (int)(push_no % kHaloBufs);
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
// Bookkeeping of a completed push, run by the last CTA of the pushing kernel (all threads): the running totals the
// waiters compare the flags with, and the epoch (parity of the halo buffers).
__device__ __forceinline__ void push_complete(const PushList &pl, const CommDev &c, unsigned long long push_no)
{
    if ((int)threadIdx.x < c.world)
        c.halo_expect[threadIdx.x] += (unsigned long long)pl.in_chunks[threadIdx.x];
    if (threadIdx.x == 0)
        *c.push_epoch = push_no;
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
    const unsigned long long push_no = push_blocks > 0 ? *rc.comm.push_epoch + 1 : 0ull; // single GPU: no comm state
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
        push_section<false>(pl, rc.comm, push_no, 0, push_blocks, [&](int row) { return FIRST ? dinv[row] * r[row] : dinv[row] * r[row] + beta * p_old[row]; });
    double acc[1] = {0}, tot[1];
    if (grid_reduce<0, THREADS>(acc, rc, tot))
    {
        if (push_blocks > 0)
            push_complete(pl, rc.comm, push_no);
        if (!FIRST && threadIdx.x == 0)
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

// push the boundary entries of an arbitrary local vector (initial guess, the iterates of the AMG smoother)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) halo_push_kernel(const double *__restrict__ v, RedCtx rc, PushList pl, const int *done)
{
    griddep_launch_dependents();
    griddep_wait();
    if (done && *done)
        return;
    const unsigned long long push_no = *rc.comm.push_epoch + 1;
    push_section<true>(pl, rc.comm, push_no, 0, (int)gridDim.x, [&](int row) { return v[row]; });
    double acc[1] = {0}, tot[1];
    if (grid_reduce<0, THREADS>(acc, rc, tot))
        push_complete(pl, rc.comm, push_no);
}

// ---------------------------------------------------------------------------------- host-level collectives (setup)
// All-gather of 8 doubles per rank with barrier semantics: everything this rank stored to peer memory before the call
// (earlier kernels of the stream) is visible to a peer once the peer has seen this rank's words. One CTA.
__global__ void wide_allgather_kernel(CommDev c, const double *__restrict__ in, double *__restrict__ out)
{
    __shared__ unsigned long long sseq;
    if (threadIdx.x == 0)
        sseq = *c.wide_seq + 1;
    __syncthreads();
    const unsigned long long seq = sseq;
    const int par = (int)(seq & 1);
    const unsigned tag = (unsigned)seq;
    const int q = threadIdx.x / kWide, i = threadIdx.x % kWide;
    if (q < c.world)
    {
        fence_acq_rel_sys();
        st_ll(&c.wide(q, par, c.rank)->w[i], in[i], tag);
        const WideSlot *src = c.wide(c.rank, par, q);
        const long long t0 = clock64();
        const bool failed_before = *(const volatile int *)c.error != 0;
        uint4 v = ld_ll(&src->w[i]);
        while (v.y != tag || v.w != tag)
        {
            if (failed_before || clock64() - t0 > c.spin_limit)
            {
                *c.error = 1;
                break;
            }
            __nanosleep(100);
            v = ld_ll(&src->w[i]);
        }
        fence_acq_rel_sys();
        out[q * kWide + i] = __longlong_as_double((long long)(((unsigned long long)v.z << 32) | v.x));
    }
    __syncthreads();
    if (threadIdx.x == 0)
        *c.wide_seq = seq;
}

struct ArenaXfer
{
    const unsigned char *src[kMaxRanks]; // put: my send buffer for rank q (already offset to this round's chunk)
    unsigned char *dst[kMaxRanks];       // get: my receive buffer for source q (already offset)
    unsigned long long bytes[kMaxRanks]; // bytes of this round's chunk (multiple of 4)
};
// put: chunk for rank q -> slot [my rank] of rank q's arena.  get: slot [q] of my arena -> receive buffer of source q.
__global__ void arena_put_kernel(CommDev c, ArenaXfer x)
{
    for (int q = 0; q < c.world; ++q)
    {
        const unsigned long long words = x.bytes[q] / 4;
        const unsigned *s = reinterpret_cast<const unsigned *>(x.src[q]);
        unsigned *d = reinterpret_cast<unsigned *>(c.arena(q, c.rank));
        for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (unsigned long long)gridDim.x * blockDim.x)
            d[i] = s[i];
    }
}
__global__ void arena_get_kernel(CommDev c, ArenaXfer x)
{
    for (int q = 0; q < c.world; ++q)
    {
        const unsigned long long words = x.bytes[q] / 4;
        const unsigned *s = reinterpret_cast<const unsigned *>(c.arena(c.rank, q));
        unsigned *d = reinterpret_cast<unsigned *>(x.dst[q]);
        for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (unsigned long long)gridDim.x * blockDim.x)
            d[i] = __ldcg(s + i);
    }
}

void Solver::dist_allgather8(const double in[8], double out[64])
{
    DistState &d = *dist;
    if (d.world == 1)
    {
        for (int i = 0; i < 8; ++i)
            out[i] = in[i];
        return;
    }
    DevBuf<double> buf;
    buf.alloc(8 + 64, true);
    PSB_CUDA(cudaMemcpyAsync(buf.p, in, sizeof(double) * 8, cudaMemcpyHostToDevice, ctx.stream));
    wide_allgather_kernel<<<1, kMaxRanks * kWide, 0, ctx.stream>>>(ctx.comm, buf.p, buf.p + 8);
    check_launch();
    PSB_CUDA(cudaMemcpyAsync(out, buf.p + 8, sizeof(double) * 64, cudaMemcpyDeviceToHost, ctx.stream));
    PSB_CUDA(cudaStreamSynchronize(ctx.stream));
    check_comm_error();
}

void Solver::dist_barrier()
{
    double in[8] = {0, 0, 0, 0, 0, 0, 0, 0}, out[64];
    dist_allgather8(in, out);
}

double Solver::dist_max(double v)
{
    double in[8] = {v, 0, 0, 0, 0, 0, 0, 0}, out[64];
    dist_allgather8(in, out);
    double m = out[0];
    for (int q = 1; q < dist->world; ++q)
        m = std::max(m, out[q * 8]);
    return m;
}

void Solver::dist_gather_ll(long long v, long long out[kMaxRanks])
{
    double in[8] = {(double)v, 0, 0, 0, 0, 0, 0, 0}, o[64];
    dist_allgather8(in, o);
    for (int q = 0; q < kMaxRanks; ++q)
        out[q] = q < dist->world ? (long long)o[q * 8] : 0;
}

// Every rank sends send_bytes[q] bytes (device memory, multiples of 4) to every rank q and receives recv_bytes[q] from it
// (filled from the senders' sizes). expect_recv (optional): the sizes the receive buffers were allocated for -- a
// disagreement is reported BEFORE anything is copied. Chunked through the staging arena.
void Solver::dist_alltoallv(const void *const send[kMaxRanks], const size_t send_bytes[kMaxRanks], void *const recv[kMaxRanks], size_t recv_bytes[kMaxRanks],
                            const size_t *expect_recv)
{
    DistState &d = *dist;
    const int W = d.world;
    double in[8] = {0, 0, 0, 0, 0, 0, 0, 0}, all[64];
    for (int q = 0; q < W; ++q)
    {
        if (send_bytes[q] % 4)
            throw std::logic_error("psb200 dist: exchange sizes must be multiples of 4 bytes");
        in[q] = (double)send_bytes[q];
    }
    dist_allgather8(in, all); // all[s * 8 + q] = bytes rank s sends to rank q
    size_t mx = 0;
    for (int s = 0; s < W; ++s)
        for (int q = 0; q < W; ++q)
            mx = std::max(mx, (size_t)all[s * 8 + q]);
    for (int q = 0; q < W; ++q)
        recv_bytes[q] = (size_t)all[q * 8 + d.rank];
    if (expect_recv)
    {
        for (int q = 0; q < W; ++q)
            if (recv_bytes[q] != expect_recv[q])
                throw std::logic_error("psb200 dist: exchange size mismatch between ranks (rank " + std::to_string(q) + " sends " +
                                       std::to_string(recv_bytes[q]) + " bytes, " + std::to_string(expect_recv[q]) + " expected)");
    }
    const size_t slot = ctx.comm.arena_slot_bytes();
    for (size_t off = 0; off < mx; off += slot)
    {
        ArenaXfer x{};
        bool any_put = false, any_get = false;
        for (int q = 0; q < W; ++q)
        {
            const size_t sb = send_bytes[q] > off ? std::min(slot, send_bytes[q] - off) : 0;
            x.src[q] = (const unsigned char *)send[q] + off;
            x.bytes[q] = sb;
            any_put |= sb > 0;
        }
        if (any_put)
        {
            arena_put_kernel<<<2 * kSMs, 256, 0, ctx.stream>>>(ctx.comm, x);
            check_launch();
        }
        dist_barrier();
        for (int q = 0; q < W; ++q)
        {
            const size_t rb = recv_bytes[q] > off ? std::min(slot, recv_bytes[q] - off) : 0;
            x.dst[q] = (unsigned char *)recv[q] + off;
            x.bytes[q] = rb;
            any_get |= rb > 0;
        }
        if (any_get)
        {
            arena_get_kernel<<<2 * kSMs, 256, 0, ctx.stream>>>(ctx.comm, x);
            check_launch();
        }
        dist_barrier(); // nobody overwrites a slot before its reader is done
    }
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
    d.halo_cap = (halo_cap + 5) / 6 * 6; // a multiple of 2 and 3: halo column ids keep their dof index modulo the block size
    d.comm_bytes = kCommHaloOff + sizeof(double) * (kHaloBufs + 2) * kMaxRanks * (size_t)d.halo_cap; // 3 halo buffers + 2 bulk parities per source
    PSB_CUDA(cudaMalloc(&d.comm_buf, d.comm_bytes));
    PSB_CUDA(cudaMemset(d.comm_buf, 0, d.comm_bytes));
    PSB_CUDA(cudaMalloc(&d.counters, 512));
    PSB_CUDA(cudaMemset(d.counters, 0, 512));
    cudaIpcMemHandle_t hnd;
    PSB_CUDA(cudaIpcGetMemHandle(&hnd, d.comm_buf));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::memcpy(handle_out, &hnd, 64);
    PSB_CUDA(cudaDeviceSynchronize());
}

static long long spin_clocks(double seconds) { return (long long)std::max(1.0, seconds * 1.9e9); }

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
    c.spin_limit = spin_clocks(prm.comm_timeout_s);
    for (int q = 0; q < kMaxRanks; ++q)
        c.peer[q] = (unsigned char *)(q < d.world ? d.peer[q] : nullptr);
    // counters (device, 512 bytes): [0] red_seq [1] push_epoch [2] error [3] bulk_epoch [4] wide_seq [8..16) bulk_expect [16..24) halo_expect
    c.red_seq = d.counters;
    c.push_epoch = d.counters + 1;
    c.error = (int *)(d.counters + 2);
    c.bulk_epoch = d.counters + 3;
    c.wide_seq = d.counters + 4;
    c.bulk_expect = d.counters + 8;
    c.halo_expect = d.counters + 16;
    c.nbr_mask = 0;
    ctx.comm = c;
    d.connected = true;
}

// Collective recovery after a communication timeout: the caller synchronises the ranks on the host (no rank may still be
// inside a call), every rank resets, the caller synchronises again. Counters, flags and slots restart from zero.
void Solver::dist_reset()
{
    if (!dist)
        throw std::runtime_error("psb200_dist_reset: psb200_dist_prepare first");
    DistState &d = *dist;
    PSB_CUDA(cudaDeviceSynchronize());
    PSB_CUDA(cudaMemset(d.comm_buf, 0, kCommHaloOff));
    PSB_CUDA(cudaMemset(d.counters, 0, 512));
    PSB_CUDA(cudaDeviceSynchronize());
    d.poisoned = false;
    if (graph_exec)
    {
        cudaGraphExecDestroy(graph_exec);
        graph_exec = nullptr;
        graph_key.clear();
    }
}

// Neighbour mask of the solver = union over the fine matrix and all partitioned AMG levels; every halo plan is cut
// into chunks for the same mask (CommDev).
void Solver::dist_set_nbr_mask(unsigned mask)
{
    DistState &d = *dist;
    mask &= ~(1u << d.rank);
    d.nbr_mask = mask;
    ctx.comm.nbr_mask = mask;
    d.fine.finalize(mask, ctx.comm, ctx.stream);
}

void Solver::analyze_pattern_dist(long long n_, long long nnz_, const int *outer, const int *inner)
{
    DistState &d = *dist;
    if (!d.connected && d.world > 1)
        throw std::runtime_error("psb200 dist: psb200_dist_connect has not been called");
    const int B = std::max(1, prm.block_size);
    if (B > 3)
        throw std::invalid_argument("psb200: block_size must be 1, 2 or 3 (reference AMGCL.cpp:111-123)");
    if (prm.krylov == "bicgstab")
        throw std::runtime_error("psb200 dist: the row-partitioned path provides PCG (krylov = cg | cg1r) with precond = jacobi | none | amg");
    if (n_ < (long long)d.world * B)
        throw std::invalid_argument("psb200 dist: fewer rows than ranks");
    d.plan.build(n_, nnz_, outer, inner, d.rank, d.world, d.halo_cap, B);
    d.A_diag.n = 0; // new pattern: the rank-local diagonal block is rebuilt at the next AMG factorize
    const DistPlanHost &P = d.plan;
    if (P.r1() == P.r0())
        throw std::invalid_argument("psb200 dist: a rank owns no rows (matrix too small for this rank count)");
    cudaStream_t st = ctx.stream;
    n = P.r1() - P.r0();
    nnz = (long long)P.ci.size();
    n_pad = (n + 3) & ~3ll;
    sym_pattern = false;
    A.n = (int)n;
    A.ncols = (int)n;
    A.nl = (int)n;
    A.rp.alloc(n + 1);
    A.ci.alloc(std::max<long long>(nnz, 1), false, 64);
    PSB_CUDA(cudaMemcpyAsync(A.rp.p, P.rp.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, st));
    if (nnz)
        PSB_CUDA(cudaMemcpyAsync(A.ci.p, P.ci.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice, st));
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
    pattern_block = B;
    if (B > 1 && amg_partitioned())
    {
        // the partitioned block AMG smooths with Dblk^-1 A on the rank's rows INCLUDING their halo columns: all B rows of a
        // node need one column list (halo lists are whole nodes, DistPlanHost::build), fill-in entries gather 0
        sort_rows_by_column(ctx, n, A.rp, A.ci, d.d_perm); // the expansion walks sorted rows
        nnz = expand_block_pattern(ctx, B, n, A.rp, A.ci, d.d_perm);
    }
    A.nnz = nnz;
    A.va.alloc(std::max<long long>(nnz, 1), false, 64);
    A.halo_mask = P.halo_cols.empty() ? 0u : 1u;
    A.block = (B > 1 && amg_partitioned()) ? B : 1; // full-block pattern only where the rows were expanded
    d.fine.world = d.world;
    d.fine.n_local = n;
    d.fine.send_begin = P.send_begin;
    d.fine.send_rows = P.send_rows;
    d.fine.recv_count = P.recv_count;
    dist_set_nbr_mask(d.fine.mask());
    A.plan(prm.spmv_kernel, st);
    // interior-first tile order of the stream schedule: tiles that touch no halo column come first, so the SpMV
    // multiplies them while the neighbours' pushes are still on the wire
    A.n_interior = 0;
    A.order_rows = 0;
    if (A.kind == SPMV_STREAM && prm.interior_first && nnz == (long long)P.ci.size())
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
    {
        const int p = perm[k];
        dst[k] = p >= 0 ? src[p] : 0.0; // p < 0: explicit zero added by the block expansion
    }
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
        // the ranks' sequence counters may now disagree: nothing collective can be trusted until psb200_dist_reset
        dist->poisoned = true;
        throw CommError("psb200 dist: a peer did not answer within comm_timeout_s (lost rank or mismatched call sequence); "
                        "the handle stays unusable until every rank calls psb200_dist_reset");
    }
}

void Solver::check_not_poisoned() const
{
    if (dist && dist->poisoned)
        throw CommError("psb200 dist: an earlier communication timeout left the ranks out of step; call psb200_dist_reset on every rank");
}

void Solver::push_halo(const HaloPlan &hp, const double *d_v, const int *done)
{
    const int push_blocks = hp.push_ctas();
    if (!push_blocks)
        return;
    ctx.prof_begin("halo_push");
    launch_chain(ctx, halo_push_kernel<kVecThreads>, push_blocks, kVecThreads, 0, d_v, ctx.red(), hp.push(), done);
    check_launch();
    ctx.prof_end();
}

void Solver::push_halo_of(const double *d_v, const int *done) { push_halo(dist->fine, d_v, done); }

// Sum of a vector across the ranks (the restriction of the "global" AMG cycle): every rank stores its partial
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
        if (!spin_ge(c.bulk_flag(c.rank, threadIdx.x), c.bulk_expect[threadIdx.x] + (unsigned long long)nchunks, c.error, c.spin_limit))
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

// All-gather of the ranks' slices of a partitioned vector into a full-length vector on every rank (transition from the
// partitioned levels of the AMG cycle to the replicated ones): rank r owns [off[r], off[r+1]). Same flow control as the
// bulk all-reduce: chunk-counted release flags per source, two parities.
struct GatherOffsets
{
    int off[kMaxRanks]; // where the segment of rank q starts in the output
    int len[kMaxRanks]; // its length (<= halo_cap)
};
template <int THREADS>
__global__ void __launch_bounds__(THREADS) bulk_allgather_kernel(const double *__restrict__ mine, double *__restrict__ out, GatherOffsets go,
                                                                 RedCtx rc, const int *done)
{
    if (done && *done)
        return;
    const CommDev &c = rc.comm;
    const int len = go.len[c.rank];
    const int nchunks = (len + kPushChunk - 1) / kPushChunk;
    const unsigned long long epoch = *c.bulk_epoch + 1;
    const int par = (int)(epoch & 1);
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x)
    {
        const int off = ch * kPushChunk, cnt = min(kPushChunk, len - off);
        for (int e = threadIdx.x; e < cnt; e += THREADS)
        {
            const double v = mine[off + e];
            out[go.off[c.rank] + off + e] = v;
            for (int q = 0; q < c.world; ++q)
                if (q != c.rank)
                    c.bulk(q, par, c.rank)[off + e] = v;
        }
        __syncthreads();
        if ((int)threadIdx.x < c.world && (int)threadIdx.x != c.rank)
            red_release_sys_add(c.bulk_flag(threadIdx.x, c.rank), 1ull);
    }
    __shared__ int s_in[kMaxRanks];
    if ((int)threadIdx.x < c.world)
    {
        const int q = threadIdx.x;
        const int inq = q == c.rank ? 0 : (go.len[q] + kPushChunk - 1) / kPushChunk;
        s_in[q] = inq;
        if (inq > 0)
        {
            if (!spin_ge(c.bulk_flag(c.rank, q), c.bulk_expect[q] + (unsigned long long)inq, c.error, c.spin_limit))
                *c.error = 1;
            fence_acq_rel_sys();
        }
    }
    __syncthreads();
    for (int q = 0; q < c.world; ++q)
    {
        if (q == c.rank)
            continue;
        const int lq = go.len[q];
        const double *src = c.bulk(c.rank, par, q);
        for (int e = blockIdx.x * THREADS + threadIdx.x; e < lq; e += gridDim.x * THREADS)
            out[go.off[q] + e] = __ldcg(src + e);
    }
    double acc[1] = {0}, tot[1];
    if (grid_reduce<0, THREADS>(acc, rc, tot) && (int)threadIdx.x < c.world)
    {
        c.bulk_expect[threadIdx.x] += (unsigned long long)s_in[threadIdx.x];
        if (threadIdx.x == 0)
            *c.bulk_epoch = epoch;
    }
}

// out[offsets[q] .. offsets[q+1]) = the slice of rank q, on every rank; slices longer than halo_cap travel in segments.
// Every rank must have a non-empty slice (parity flow control: everybody hears from everybody in every exchange).
void Solver::bulk_allgather(const double *d_mine, double *d_out, const long long *offsets, const int *done)
{
    DistState &d = *dist;
    long long mx = 0, mn = 1ll << 62;
    for (int q = 0; q < d.world; ++q)
    {
        mx = std::max(mx, offsets[q + 1] - offsets[q]);
        mn = std::min(mn, offsets[q + 1] - offsets[q]);
    }
    if (mn <= 0)
        throw std::runtime_error("psb200 dist: all-gather with an empty slice");
    const long long cap = d.halo_cap;
    const long long nseg = (mx + cap - 1) / cap;
    for (long long sg = 0; sg < nseg; ++sg)
    {
        GatherOffsets go{};
        long long smx = 0;
        for (int q = 0; q < d.world; ++q)
        {
            const long long lenq = offsets[q + 1] - offsets[q];
            // spread every slice evenly over the segments so that no rank has an empty one
            const long long a = lenq * sg / nseg, b = lenq * (sg + 1) / nseg;
            go.off[q] = (int)(offsets[q] + a);
            go.len[q] = (int)(b - a);
            smx = std::max(smx, b - a);
            if (b - a <= 0)
                throw std::runtime_error("psb200 dist: all-gather segment became empty (slices too unbalanced for halo_cap)");
        }
        const long long mylen = offsets[d.rank + 1] - offsets[d.rank];
        const long long my_a = mylen * sg / nseg;
        const int nchunks = (int)((smx + kPushChunk - 1) / kPushChunk);
        const int grid = std::max(1, std::min(nchunks, 2 * kSMs));
        ctx.prof_begin("bulk_allgather");
        bulk_allgather_kernel<kVecThreads><<<grid, kVecThreads, 0, ctx.stream>>>(d_mine + my_a, d_out, go, ctx.red(), done);
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

// D = M[local rows, local columns] of a row-partitioned matrix M (columns >= M.nl dropped); src[k] = position in M
void extract_diag_block(Ctx &ctx, const CsrDev &M, CsrDev &D, DevBuf<int> &src)
{
    cudaStream_t st = ctx.stream;
    const int n = M.n;
    DevBuf<int> cnt;
    cnt.alloc((size_t)n + 1, true);
    D.n = n;
    D.ncols = n;
    D.rp.alloc((size_t)n + 1);
    const unsigned blocks = (unsigned)((n + 256) / 256);
    diag_count_kernel<<<blocks, 256, 0, st>>>(n, n, M.rp.p, M.ci.p, cnt.p);
    size_t bytes = 0;
    PSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt.p, D.rp.p, n + 1, st));
    DevBuf<unsigned char> tmp;
    tmp.alloc(bytes);
    PSB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, cnt.p, D.rp.p, n + 1, st));
    int dn = 0;
    PSB_CUDA(cudaMemcpyAsync(&dn, D.rp.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaStreamSynchronize(st));
    D.nnz = dn;
    D.ci.alloc(std::max(1, dn), false, 64);
    D.va.alloc(std::max(1, dn), false, 64);
    src.alloc(std::max(1, dn));
    diag_fill_kernel<<<blocks, 256, 0, st>>>(n, n, M.rp.p, M.ci.p, D.rp.p, D.ci.p, src.p);
    check_launch();
    PSB_CUDA(cudaStreamSynchronize(st));
}

// A_diag = A[local rows, local columns]: pattern once per analysis, values on every factorize
void Solver::build_diag_block_dist()
{
    DistState &d = *dist;
    cudaStream_t st = ctx.stream;
    CsrDev &D = d.A_diag;
    if (D.n != (int)n || D.rp.p == nullptr || d.diag_src.n == 0)
    {
        extract_diag_block(ctx, A, D, d.diag_src);
        if (pattern_block > 1 && n > 0)
        {
            // block mode: the B rows of a node get the full B x B block pattern (diag_src = -1 marks the fill-in),
            // the invariant the block AMG kernels rely on (same expansion as the single-GPU analyze_pattern)
            D.nnz = expand_block_pattern(ctx, pattern_block, n, D.rp, D.ci, d.diag_src);
            D.va.alloc(std::max<long long>(1, D.nnz), false, 64);
            D.block = pattern_block;
        }
        PSB_CUDA(cudaStreamSynchronize(st));
        D.plan("auto", st);
    }
    if (D.nnz)
        diag_vals_kernel<<<(unsigned)((D.nnz + 255) / 256), 256, 0, st>>>(D.nnz, A.va.p, d.diag_src.p, D.va.p);
    check_launch();
    D.refresh_bsr(st);
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
    const unsigned long long push_no = push_blocks > 0 ? *rc.comm.push_epoch + 1 : 0ull; // single GPU: no comm state
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
        push_section<true>(pl, rc.comm, push_no, 0, push_blocks, [&](int row) { return s[row] + (beta != 0.0 ? beta * p_old[row] : 0.0); });
    double acc[1] = {0}, tot[1];
    if (grid_reduce<0, THREADS>(acc, rc, tot) && push_blocks > 0)
        push_complete(pl, rc.comm, push_no);
}

// AMG-PCG on the row partition (amgcl cg ordering): the global CG with halo push of p and three fused
// all-reduces per iteration; the preconditioner is the partitioned hierarchy (amg_dist.cu), the "global" hierarchy with
// a partitioned fine level, or the rank-local one (block-Jacobi across ranks).
void Solver::run_cg_amgcl_dist(const double *d_b)
{
    if (!amg && !amg_dist)
        throw std::runtime_error("psb200_solve: AMG hierarchy missing (factorize with precond=amg first)");
    DistState &d = *dist;
    KState *S = d_state;
    const int *done = &S->done;
    d.vp2.alloc((size_t)n_pad, true);
    init_state(*this, prm.tolerance, prm.max_iter);
    const long long n2 = n_pad / 2;
    const int vec_blocks = vec_grid(n2);
    const int push_blocks = d.fine.push_ctas();
    PushList pl = d.fine.push();
    RedCtx rc = ctx.red();
    PSB_CUDA(cudaMemsetAsync(vp.p, 0, sizeof(double) * n_pad, ctx.stream));
    push_halo(d.fine, vx.p, nullptr);
    launch_spmv(ctx, "spmv_residual", A, vx.p, EpiResidualNorms{vr.p, d_b, dinv.p}, FinInitAmgcl{S});
    auto batch = [&]() {
        double *pc = vp.p, *pn = d.vp2.p;
        for (int i = 0; i < 2; ++i)
        {
            if (amg_dist)
                amg_dist->apply(vr.p, vz.p, done); // every large level partitioned, small levels replicated
            else if (amg->has_dist_fine())
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
    key << "cg_amgcl_dist/" << n << "/" << (void *)vx.p << "/" << (void *)amg.get() << "/" << (void *)amg_dist.get() << "/" << (void *)d_b;
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
    const int push_blocks = d.fine.push_ctas();
    PushList pl = d.fine.push();
    RedCtx rc = ctx.red();
    push_halo(d.fine, vx.p, nullptr);
    launch_spmv(ctx, "spmv_residual", A, vx.p, EpiResidualNorms{vr.p, d_b, dinv.p}, FinInitEigen{S});
    ctx.prof_begin("cg_dir");
    launch_chain(ctx, cg_dir_dist_kernel<true, kVecThreads>, vec_blocks + push_blocks, kVecThreads, 0, n2, vp.p, vp.p, vr.p, dinv.p, S, rc, pl, vec_blocks, done);
    check_launch();
    ctx.prof_end();
    const int batch_iters = std::max(2, prm.check_every & ~1);
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

// ---------------------------------------------------------------------------------- single-reduction PCG (krylov = cg1r)
// Chronopoulos-Gear form of the preconditioned CG (M = D^-1 or I): with u = M r and w = A u,
//     gamma = r.u   delta = w.u   beta = gamma / gamma_old   alpha = gamma / (delta - beta gamma / alpha_old)
//     p = u + beta p   s = w + beta s   x += alpha p   r -= alpha s
// all three dot products of an iteration (gamma, delta, ||r||^2) are taken from the same vectors, so they travel in ONE
// all-reduce, fused into the SpMV that produces w; the vector update is fused with the halo push of the new u. Two
// kernels and two synchronisation points (halo wait, all-reduce) per iteration instead of three kernels and three.
// In exact arithmetic the iterates equal those of the Eigen ordering; the iteration counter follows the same rule
// (incremented after a trip that did not converge).
struct EpiCg1r
{
    static constexpr int NV = 3;
    using Pre = Pre2;
    double *w;
    const double *r, *u;
    __device__ __forceinline__ Pre pre(int row) const { return {__ldg(r + row), __ldg(u + row)}; }
    __device__ __forceinline__ void operator()(int row, double s, Pre q, double (&acc)[3]) const
    {
        w[row] = s;
        acc[0] += q.a * q.b;
        acc[1] += s * q.b;
        acc[2] += q.a * q.a;
    }
};
struct FinCg1r
{
    KState *st;
    __device__ __forceinline__ void operator()(const double *t) const
    {
        const double gamma = t[0], delta = t[1], rn2 = t[2];
        st->rn2 = rn2;
        if (st->c_started && rn2 < st->thr)
        {
            st->done = 1;
            st->status = ST_CONVERGED;
            return;
        }
        if (!(rn2 == rn2) || isinf(rn2) || !(delta == delta))
        {
            st->done = 1;
            st->status = ST_BREAKDOWN;
            return;
        }
        double beta = 0.0, alpha;
        if (!st->c_started)
            alpha = gamma / delta;
        else
        {
            beta = gamma / st->rz;
            alpha = gamma / (delta - beta * gamma / st->alpha);
            st->iter += 1;
            if (st->iter >= st->max_iter)
            {
                st->done = 1;
                st->status = ST_MAXITER;
            }
        }
        st->c_started = 1;
        st->rz = gamma;
        st->alpha = alpha;
        st->c_beta = beta;
    }
};
// u = dinv r (and its halo push): the start of the recurrence
template <int THREADS>
__global__ void __launch_bounds__(THREADS) cg1r_start_kernel(long long n2, double *__restrict__ u, const double *__restrict__ r,
                                                             const double *__restrict__ dinv, RedCtx rc, PushList pl, int vec_blocks, const int *done)
{
    if (done && *done)
        return;
    const int push_blocks = (int)gridDim.x - vec_blocks;
    const unsigned long long push_no = push_blocks > 0 ? *rc.comm.push_epoch + 1 : 0ull; // single GPU: no comm state
    if ((int)blockIdx.x >= push_blocks)
    {
        const long long stride = (long long)vec_blocks * THREADS;
        for (long long j = (long long)(blockIdx.x - push_blocks) * THREADS + threadIdx.x; j < n2; j += stride)
        {
            const double2 rv = ld2(r, j), dv = ld2(dinv, j);
            double2 o;
            o.x = dv.x * rv.x;
            o.y = dv.y * rv.y;
            st2(u, j, o);
        }
    }
    else
        push_section<true>(pl, rc.comm, push_no, 0, push_blocks, [&](int row) { return dinv[row] * r[row]; });
    double acc[1] = {0}, tot[1];
    if (grid_reduce<0, THREADS>(acc, rc, tot) && push_blocks > 0)
        push_complete(pl, rc.comm, push_no);
}
// p = u + beta p; s_new = w + beta s; x += alpha p; r_new = r - alpha s_new; u_new = dinv r_new   (+ halo push of u_new).
// r, s and u ping-pong between two buffers each: the pushing CTAs recompute the boundary entries of u_new from the OLD
// r / s while the other CTAs write the new ones, so nothing is read and written in the same launch.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) cg1r_update_kernel(long long n2, double *__restrict__ x, double *__restrict__ p,
                                                              double *__restrict__ r_new, const double *__restrict__ r,
                                                              double *__restrict__ s_new, const double *__restrict__ s,
                                                              double *__restrict__ u_new, const double *__restrict__ u,
                                                              const double *__restrict__ w, const double *__restrict__ dinv, const KState *st,
                                                              RedCtx rc, PushList pl, int vec_blocks, const int *done)
{
    if (done && *done)
        return;
    const double alpha = st->alpha, beta = st->c_beta;
    const int push_blocks = (int)gridDim.x - vec_blocks;
    const unsigned long long push_no = push_blocks > 0 ? *rc.comm.push_epoch + 1 : 0ull; // single GPU: no comm state
    if ((int)blockIdx.x >= push_blocks)
    {
        const long long stride = (long long)vec_blocks * THREADS;
        for (long long j = (long long)(blockIdx.x - push_blocks) * THREADS + threadIdx.x; j < n2; j += stride)
        {
            const double2 uv = ld2(u, j), wv = ld2(w, j), dv = ld2(dinv, j);
            double2 pv = ld2(p, j), sv = ld2(s, j), xv = ld2(x, j), rv = ld2(r, j);
            pv.x = uv.x + beta * pv.x;
            pv.y = uv.y + beta * pv.y;
            sv.x = wv.x + beta * sv.x;
            sv.y = wv.y + beta * sv.y;
            xv.x += alpha * pv.x;
            xv.y += alpha * pv.y;
            rv.x -= alpha * sv.x;
            rv.y -= alpha * sv.y;
            st2(p, j, pv);
            st2(s_new, j, sv);
            st2(x, j, xv);
            st2(r_new, j, rv);
            double2 un;
            un.x = dv.x * rv.x;
            un.y = dv.y * rv.y;
            st2(u_new, j, un);
        }
    }
    else
        push_section<false>(pl, rc.comm, push_no, 0, push_blocks,
                     [&](int row) { return dinv[row] * (r[row] - alpha * (w[row] + beta * s[row])); });
    double acc[1] = {0}, tot[1];
    if (grid_reduce<0, THREADS>(acc, rc, tot) && push_blocks > 0)
        push_complete(pl, rc.comm, push_no);
}

void Solver::run_cg1r(const double *d_b)
{
    KState *S = d_state;
    const int *done = &S->done;
    // ping-pong pairs: r = vr / vz, s = vt / vr0, u = vy / vv; w = vq. After an even number of trips the current r, s, u
    // are back in vr, vt, vy (the batch is unrolled over an even count and re-enters with the same pointers).
    for (DevBuf<double> *v : {&vy, &vv, &vt, &vr0, &vz})
        v->alloc((size_t)n_pad, true);
    init_state(*this, prm.tolerance, prm.max_iter);
    const long long n2 = n_pad / 2;
    const int vec_blocks = vec_grid(n2);
    const int push_blocks = dist ? dist->fine.push_ctas() : 0;
    PushList pl{};
    if (dist)
        pl = dist->fine.push();
    RedCtx rc = ctx.red();
    cudaStream_t st = ctx.stream;
    PSB_CUDA(cudaMemsetAsync(vp.p, 0, sizeof(double) * n_pad, st));
    PSB_CUDA(cudaMemsetAsync(vt.p, 0, sizeof(double) * n_pad, st));
    if (dist)
        push_halo(dist->fine, vx.p, nullptr);
    launch_spmv(ctx, "spmv_residual", A, vx.p, EpiResidualNorms{vr.p, d_b, dinv.p}, FinInitEigen{S});
    ctx.prof_begin("cg1r_start");
    cg1r_start_kernel<kVecThreads><<<vec_blocks + push_blocks, kVecThreads, 0, st>>>(n2, vy.p, vr.p, dinv.p, rc, pl, vec_blocks, done);
    check_launch();
    ctx.prof_end();
    launch_spmv(ctx, "spmv_cg1r", A, vy.p, EpiCg1r{vq.p, vr.p, vy.p}, FinCg1r{S}, done);
    const int batch_iters = std::max(2, prm.check_every & ~1);
    auto batch = [&]() {
        double *rc_ = vr.p, *rn = vz.p, *sc = vt.p, *sn = vr0.p, *uc = vy.p, *un = vv.p;
        for (int i = 0; i < batch_iters; ++i)
        {
            ctx.prof_begin("cg1r_update");
            cg1r_update_kernel<kVecThreads><<<vec_blocks + push_blocks, kVecThreads, 0, st>>>(n2, vx.p, vp.p, rn, rc_, sn, sc, un, uc, vq.p, dinv.p, S,
                                                                                             rc, pl, vec_blocks, done);
            check_launch();
            ctx.prof_end();
            launch_spmv(ctx, "spmv_cg1r", A, un, EpiCg1r{vq.p, rn, un}, FinCg1r{S}, done);
            std::swap(rc_, rn);
            std::swap(sc, sn);
            std::swap(uc, un);
        }
    };
    std::ostringstream key;
    key << "cg1r/" << A.kind << "/" << A.lpr << "/" << n << "/" << (void *)vx.p << "/" << (void *)A.va.p << "/" << (void *)d_b << "/" << batch_iters << "/" << (dist ? dist->world : 1);
    drive(batch, batch_iters, key.str());
    finish_solve();
    if (dist)
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

int psb200_dist_reset(psb200_handle h)
{
    if (!h)
        return PSB200_ERR_INVALID;
    try
    {
        h->s.err.clear();
        psb::DeviceScope device_scope(h->s.device, h->s.ctx.stream != nullptr);
        h->s.dist_reset();
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
