// Row-partitioned smoothed-aggregation AMG: the multi-GPU form of amg.cu (the reference has no counterpart,
// its AMGCL path is single-process, AMGCL.cpp:148-212 / block form :246-298).
//
//  * aggregation is DECOUPLED: every rank aggregates its own rows on the diagonal block of the level matrix, so an
//    aggregate never crosses a rank boundary and the coarse unknowns of a rank are its own aggregates (coarse row
//    offsets = prefix sums of the aggregate counts);
//  * the prolongation is smoothed with the rank-local FILTERED matrix: connections that leave the rank are treated as
//    amgcl treats weak connections (dropped from the off-diagonal part and lumped into the diagonal -- for block problems
//    the B x B blocks are lumped into the diagonal block), so P and R = P^T are rank-local: restriction and
//    prolongation need no exchange, and constants stay in the range of P;
//  * the Galerkin product A_c = R (A P) is the only step that needs remote data: ONE exchange of the P rows of the halo
//    columns (row lengths, then global coarse column ids and values through the staging arena), then two local SpGEMMs;
//  * every partitioned level has its own halo plan (send lists obtained from the consumers' requests), the smoother
//    pushes the halo of every iterate exactly like the fine level does;
//  * levels with fewer than amg.replicate_below stored non-zeros (summed over the ranks) are replicated: their matrix is all-gathered once at setup, the
//    cycle all-gathers the coarse right-hand side (one fused kernel over NVLink), runs the remaining levels redundantly on
//    every rank -- no latency-bound exchanges on tiny levels -- and every rank prolongs from its own slice.
//
// Same Chebyshev smoother, cycle shape (ncycle, npre, npost) and parameters as the single-GPU hierarchy.
#include "amg_dist.hpp"
#include "amg_internal.hpp"
#include "push_epi.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include <algorithm>
#include <numeric>
#include <sstream>

namespace psb {

namespace {

inline int nblk(long long n, int t = 256) { return (int)std::max<long long>(1, (n + t - 1) / t); }

// D holds the diagonal block of M (src = positions in M). Adds every halo entry of row i to the entry (i, B (i / B) + c % B):
// the connection is dropped from the off-diagonal part and lumped into the diagonal (block), as amgcl's filtered matrix does
// with weak connections.
__global__ void lump_halo_kernel(int B, CsrView M, const int *__restrict__ drp, const int *__restrict__ dci, const int *__restrict__ src,
                                 double *__restrict__ dva)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M.n)
        return;
    const int d0 = drp[i], d1 = drp[i + 1];
    for (int k = d0; k < d1; ++k)
        dva[k] = src[k] >= 0 ? M.va[src[k]] : 0.0;
    const int node0 = B * (i / B);
    for (int k = M.rp[i]; k < M.rp[i + 1]; ++k)
    {
        const int c = M.ci[k];
        if (c < M.nl)
            continue;
        const int target = node0 + c % B;
        for (int q = d0; q < d1; ++q)
            if (dci[q] == target)
            {
                dva[q] += M.va[k];
                break;
            }
    }
}

// lens[e] = entries of row rows[e] of P
__global__ void row_lengths_kernel(int cnt, const int *__restrict__ rows, const int *__restrict__ rp, int *__restrict__ lens)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < cnt)
        lens[e] = rp[rows[e] + 1] - rp[rows[e]];
    else if (e == cnt)
        lens[e] = 0;
}
// packs the rows rows[e] of P (columns shifted to global coarse ids) at off[e]
__global__ void pack_rows_kernel(int cnt, const int *__restrict__ rows, const int *__restrict__ off, const int *__restrict__ rp,
                                 const int *__restrict__ ci, const double *__restrict__ va, int col_shift, int *__restrict__ oc,
                                 double *__restrict__ ov)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= cnt)
        return;
    int o = off[e];
    for (int k = rp[rows[e]]; k < rp[rows[e] + 1]; ++k, ++o)
    {
        oc[o] = ci[k] + col_shift;
        ov[o] = va[k];
    }
}
__global__ void shift_copy_int_kernel(long long n, const int *__restrict__ in, int shift, int *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = in[i] + shift;
}
// column ids of a partitioned matrix (nl + q * halo_cap + pos for halo columns) -> compact ids nl + seg_start[q] + pos
struct SegTable
{
    int seg_start[kMaxRanks + 1];
    long long off[kMaxRanks + 1];
};
__global__ void compact_cols_kernel(long long nnz, const int *__restrict__ ci, int nl, long long halo_cap, SegTable t, int *__restrict__ out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz)
        return;
    const int c = ci[k];
    if (c < nl)
        out[k] = c;
    else
    {
        const long long h = c - nl;
        const int q = (int)(h / halo_cap);
        out[k] = nl + t.seg_start[q] + (int)(h % halo_cap);
    }
}
// flags the entries whose global column lies outside [lo, hi)
__global__ void flag_offrange_kernel(long long nnz, const int *__restrict__ ci, int lo, int hi, unsigned char *__restrict__ flag)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz)
        flag[k] = ci[k] < lo || ci[k] >= hi;
}
// global column -> [local | halo] numbering of the next level
__global__ void remap_cols_kernel(long long nnz, const int *__restrict__ gci, int world, int rank, SegTable t, long long halo_cap,
                                  const int *__restrict__ halo_ids, int nh, int *__restrict__ out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz)
        return;
    const int c = gci[k];
    const int lo = (int)t.off[rank], hi = (int)t.off[rank + 1];
    if (c >= lo && c < hi)
    {
        out[k] = c - lo;
        return;
    }
    int a = 0, b = nh;
    while (a < b)
    {
        const int m = (a + b) >> 1;
        if (halo_ids[m] < c)
            a = m + 1;
        else
            b = m;
    }
    int q = 0;
    while (q + 1 < world && c >= (int)t.off[q + 1])
        ++q;
    out[k] = (int)((hi - lo) + q * halo_cap + (a - t.seg_start[q]));
}
// lens[i] = rp[i + 1] - rp[i], lens[n] = 0
__global__ void row_diff_kernel(int n, const int *__restrict__ rp, int *__restrict__ lens)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        lens[i] = rp[i + 1] - rp[i];
    else if (i == n)
        lens[i] = 0;
}
// flag[t] = 1 when tile t (rows_per_tile consecutive rows) holds a row that reads a halo column or is sent to a neighbour
__global__ void boundary_tiles_kernel(CsrView A, const unsigned *__restrict__ send_bits, int rows_per_tile, int ntiles, int *__restrict__ flag)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles)
        return;
    const int r0 = t * rows_per_tile, r1 = min(A.n, r0 + rows_per_tile);
    int f = 0;
    for (int r = r0; r < r1 && !f; ++r)
    {
        if ((send_bits[r >> 5] >> (r & 31)) & 1u)
            f = 1;
        for (int k = A.rp[r]; k < A.rp[r + 1] && !f; ++k)
            f = A.ci[k] >= A.nl;
    }
    flag[t] = f;
}
__global__ void rp_from_lengths_kernel(int n, const int *__restrict__ scan, int base, int *__restrict__ rp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n)
        rp[i] = scan[i] + base;
}

} // namespace

// Boundary-first tile order of a partitioned level matrix: the tiles whose rows are read by or sent to a neighbour come
// first, so their new values are pushed (EpiChebPush) while the kernel still works on the interior tiles, and the
// neighbours' next kernel finds its halo complete. One order for the stream schedule of the scalar CSR, one for BSR-3.
static void build_boundary_first(Ctx &ctx, CsrDev &M, const HaloPlan &hp)
{
    cudaStream_t st = ctx.stream;
    auto make = [&](int rows_per_tile, DevBuf<int> &out) {
        const int ntiles = (M.n + rows_per_tile - 1) / rows_per_tile;
        DevBuf<int> flag;
        flag.alloc(std::max(1, ntiles));
        boundary_tiles_kernel<<<nblk(ntiles), 256, 0, st>>>(M.view(), hp.send_bits.p, rows_per_tile, ntiles, flag.p);
        check_launch();
        std::vector<int> h(ntiles), order;
        if (ntiles)
            PSB_CUDA(cudaMemcpyAsync(h.data(), flag.p, sizeof(int) * ntiles, cudaMemcpyDeviceToHost, st));
        PSB_CUDA(cudaStreamSynchronize(st));
        order.reserve(ntiles);
        for (int t = 0; t < ntiles; ++t)
            if (h[t])
                order.push_back(t);
        for (int t = 0; t < ntiles; ++t)
            if (!h[t])
                order.push_back(t);
        out.alloc(std::max(1, ntiles));
        if (ntiles)
            PSB_CUDA(cudaMemcpyAsync(out.p, order.data(), sizeof(int) * ntiles, cudaMemcpyHostToDevice, st));
        PSB_CUDA(cudaStreamSynchronize(st));
    };
    M.use_order = false;
    M.use_bsr_order = false;
    if (M.n == 0 || M.halo_mask == 0)
        return;
    if (M.kind == SPMV_STREAM)
    {
        make(M.stream_rows(), M.tile_order);
        M.n_interior = 0; // the halo is needed from the first tile on
        M.order_rows = M.stream_rows();
        M.use_order = true;
    }
    if (M.use_bsr && M.block == 3)
    {
        make(3 * BsrProd::rows, M.bsr_tile_order);
        M.use_bsr_order = true;
    }
}

// ====================================================================================== level
struct DistAmgLevel
{
    AmgLevel L;                     // local rows: A (with halo columns), rank-local P / R, smoother data, work vectors
    HaloPlan plan_own;              // halo exchange of this level's vectors (levels >= 1)
    HaloPlan *plan = nullptr;
    std::vector<long long> offsets; // world + 1 global row offsets of this level
    long long n_global = 0, nnz_global = 0, p_nnz_global = 0, agg_global = 0;
    DevBuf<double> fc;              // restricted residual of the rank's coarse unknowns (input of the next level / the tail)
    double t_total = 0, t_exchange = 0, t_plan = 0;
};

// amg.fused_push (or PSB200_FUSED_PUSH=on in the environment, for A/B runs of an unchanged caller)
static bool env_fused_push()
{
    static bool on = std::getenv("PSB200_FUSED_PUSH") != nullptr && std::string(std::getenv("PSB200_FUSED_PUSH")) == "on";
    return on;
}

AmgDist::AmgDist(Solver &s, const AmgParams &prm) : s_(s), prm_(prm) {}
AmgDist::~AmgDist() {}
int AmgDist::num_levels() const { return (int)levels_.size() + (tail_ ? tail_->num_levels() : 0); }

void AmgDist::refinalize_plans()
{
    unsigned mask = s_.dist->fine.mask();
    for (auto &lv : levels_)
        mask |= lv->plan->mask();
    s_.dist_set_nbr_mask(mask); // re-cuts the fine plan
    for (auto &lv : levels_)
        if (lv->plan != &s_.dist->fine)
            lv->plan->finalize(s_.dist->nbr_mask, s_.ctx.comm, s_.ctx.stream);
}

// ====================================================================================== setup
void AmgDist::setup(const std::vector<std::vector<int>> &imposed)
{
    // direct_coarse applies to the coarsest level, which always lives in the replicated tail
    if (prm_.relax_type != "chebyshev")
        throw std::runtime_error("psb200 amg: the partitioned cycle provides the Chebyshev smoother (polysolve's default, AMGCL.cpp:36-47)");
    Ctx &ctx = s_.ctx;
    cudaStream_t st = ctx.stream;
    DistState &D = *s_.dist;
    const int W = D.world, me = D.rank;
    const int B = std::max(1, prm_.block_size);
    const double t_begin = wall_ms(st);
    levels_.clear();
    tail_.reset();
    Temp tmp;

    auto cur = std::make_unique<DistAmgLevel>();
    cur->L.A = &s_.A;
    cur->plan = &D.fine;
    cur->offsets = D.plan.offsets;
    cur->n_global = s_.n_global;
    {
        long long all[kMaxRanks];
        s_.dist_gather_ll(s_.A.nnz, all);
        cur->nnz_global = std::accumulate(all, all + W, 0ll);
    }
    double eps_strong = prm_.eps_strong;

    while (cur)
    {
        levels_.push_back(std::move(cur));
        DistAmgLevel &lv = *levels_.back();
        AmgLevel &L = lv.L;
        const CsrDev &A = *L.A;
        const int li = (int)levels_.size() - 1;
        const double t_lv = wall_ms(st);
        refinalize_plans();
        SetupHooks hooks;
        hooks.row0 = lv.offsets[me];
        hooks.push = [this, &lv](const double *v) { s_.push_halo(*lv.plan, v, nullptr); };
        hooks.allmax = [this](double v) { return s_.dist_max(v); };
        double tp = wall_ms(st);
        setup_relaxation(ctx, prm_, L, li, &hooks); // the power iteration multiplies with the partitioned matrix, dots all-reduce
        L.t_relax = wall_ms(st) - tp;
        if ((prm_.fused_push || env_fused_push()) && W > 1)
        {
            build_boundary_first(ctx, li == 0 ? s_.A : L.Aown, *lv.plan);
            if (B > 1)
                build_boundary_first(ctx, L.Ahat, *lv.plan);
        }
        if (li + 1 >= prm_.max_levels || lv.n_global <= prm_.coarse_enough)
            break; // relaxation-only last level (only when the whole hierarchy is a single level or max_levels is tiny)

        // ---- decoupled aggregation on the lumped diagonal block
        tp = wall_ms(st);
        CsrDev Af;
        {
            DevBuf<int> src;
            extract_diag_block(ctx, A, Af, src);
            if (Af.nnz)
                lump_halo_kernel<<<nblk(A.n), 256, 0, st>>>(B, A.view(), Af.rp.p, Af.ci.p, src.p, Af.va.p);
            check_launch();
            PSB_CUDA(cudaStreamSynchronize(st));
        }
        {
            LocalScope local(ctx);
            build_aggregates(ctx, tmp, prm_, Af, eps_strong, nullptr, L);
        }
        L.t_agg = wall_ms(st) - tp;
        long long nagg[kMaxRanks];
        s_.dist_gather_ll(L.n_agg, nagg);
        std::vector<long long> coff(W + 1, 0);
        for (int q = 0; q < W; ++q)
            coff[q + 1] = coff[q] + nagg[q];
        const long long ncg = coff[W];
        lv.agg_global = ncg;
        if (ncg <= 0)
            break;
        if (ncg > 0x7fffffffLL - 1024)
            throw std::runtime_error("psb200 amg: coarse level exceeds the int32 index range");
        for (int q = 0; q < W; ++q)
            if (nagg[q] == 0)
                throw std::runtime_error("psb200 amg: rank " + std::to_string(q) + " has no aggregates at level " + std::to_string(li) +
                                         " (partition too fine for this matrix; raise amg.replicate_below)");
        // ---- omega from the Gershgorin bound of the TRUE level matrix (local rows incl. halo columns), max over ranks
        double omega = prm_.sa_relax;
        if (prm_.estimate_spectral_radius)
            omega *= (4.0 / 3.0) / s_.dist_max(gershgorin_rho(ctx, prm_, L));
        else
            omega *= 2.0 / 3.0;
        L.omega = omega;
        // ---- rank-local smoothed prolongation and its transpose
        tp = wall_ms(st);
        build_prolongation(ctx, tmp, prm_, Af, nullptr, eps_strong, omega, L);
        L.t_prolong = wall_ms(st) - tp;
        eps_strong *= 0.5;
        tp = wall_ms(st);
        transpose(ctx, tmp, L.P, L.R);
        L.P.plan("auto", st);
        L.R.plan("auto", st);
        L.t_transpose = wall_ms(st) - tp;
        Af = CsrDev();
        {
            long long all[kMaxRanks];
            s_.dist_gather_ll(L.P.nnz, all);
            lv.p_nnz_global = std::accumulate(all, all + W, 0ll);
        }

        // ---- P rows of the halo columns: lengths, then packed (global column, value) pairs
        tp = wall_ms(st);
        const HaloPlan &hp = *lv.plan;
        const int nsend = (int)hp.send_rows.size();
        std::vector<int> seg_start(W + 1, 0);
        for (int q = 0; q < W; ++q)
            seg_start[q + 1] = seg_start[q] + hp.recv_count[q];
        const int nh = seg_start[W];
        DevBuf<int> d_send_rows, lens, lscan, hl, hscan;
        d_send_rows.alloc(std::max(1, nsend));
        lens.alloc((size_t)nsend + 1, true);
        lscan.alloc((size_t)nsend + 1, true);
        hl.alloc((size_t)nh + 1, true);
        hscan.alloc((size_t)nh + 1, true);
        if (nsend)
            PSB_CUDA(cudaMemcpyAsync(d_send_rows.p, hp.send_rows.data(), sizeof(int) * nsend, cudaMemcpyHostToDevice, st));
        row_lengths_kernel<<<nblk(nsend + 1), 256, 0, st>>>(nsend, d_send_rows.p, L.P.rp.p, lens.p);
        check_launch();
        {
            const void *snd[kMaxRanks] = {};
            void *rcv[kMaxRanks] = {};
            size_t sb[kMaxRanks] = {}, rb[kMaxRanks] = {};
            size_t ex[kMaxRanks] = {};
            for (int q = 0; q < W; ++q)
            {
                snd[q] = lens.p + hp.send_begin[q];
                sb[q] = sizeof(int) * (size_t)(hp.send_begin[q + 1] - hp.send_begin[q]);
                rcv[q] = hl.p + seg_start[q];
                ex[q] = sizeof(int) * (size_t)hp.recv_count[q];
            }
            s_.dist_alltoallv(snd, sb, rcv, rb, ex);
            for (int q = 0; q < W; ++q)
                if (rb[q] != sizeof(int) * (size_t)hp.recv_count[q])
                    throw std::logic_error("psb200 amg: halo plan mismatch between ranks (row lengths)");
        }
        exclusive_scan_int(ctx, tmp, lens.p, lscan.p, (long long)nsend + 1);
        exclusive_scan_int(ctx, tmp, hl.p, hscan.p, (long long)nh + 1);
        std::vector<int> h_lscan((size_t)nsend + 1), h_hscan((size_t)nh + 1);
        PSB_CUDA(cudaMemcpyAsync(h_lscan.data(), lscan.p, sizeof(int) * ((size_t)nsend + 1), cudaMemcpyDeviceToHost, st));
        PSB_CUDA(cudaMemcpyAsync(h_hscan.data(), hscan.p, sizeof(int) * ((size_t)nh + 1), cudaMemcpyDeviceToHost, st));
        PSB_CUDA(cudaStreamSynchronize(st));
        const int tot_send = h_lscan[nsend], tot_recv = h_hscan[nh];
        // P_ext = [P with global columns ; P rows of the halo columns]
        CsrDev Pext;
        Pext.n = A.n + nh;
        Pext.ncols = (int)ncg;
        Pext.nnz = L.P.nnz + tot_recv;
        Pext.rp.alloc((size_t)Pext.n + 1);
        Pext.ci.alloc(std::max<long long>(1, Pext.nnz), false, 64);
        Pext.va.alloc(std::max<long long>(1, Pext.nnz), false, 64);
        PSB_CUDA(cudaMemcpyAsync(Pext.rp.p, L.P.rp.p, sizeof(int) * ((size_t)A.n + 1), cudaMemcpyDeviceToDevice, st));
        rp_from_lengths_kernel<<<nblk(nh + 1), 256, 0, st>>>(nh, hscan.p, (int)L.P.nnz, Pext.rp.p + A.n);
        if (L.P.nnz)
        {
            shift_copy_int_kernel<<<nblk(L.P.nnz), 256, 0, st>>>(L.P.nnz, L.P.ci.p, (int)coff[me], Pext.ci.p);
            PSB_CUDA(cudaMemcpyAsync(Pext.va.p, L.P.va.p, sizeof(double) * (size_t)L.P.nnz, cudaMemcpyDeviceToDevice, st));
        }
        check_launch();
        {
            DevBuf<int> pc;
            DevBuf<double> pv;
            pc.alloc(std::max(1, tot_send));
            pv.alloc(std::max(1, tot_send));
            if (nsend)
                pack_rows_kernel<<<nblk(nsend), 256, 0, st>>>(nsend, d_send_rows.p, lscan.p, L.P.rp.p, L.P.ci.p, L.P.va.p, (int)coff[me], pc.p, pv.p);
            check_launch();
            const void *snd[kMaxRanks] = {};
            void *rcv[kMaxRanks] = {};
            size_t sb[kMaxRanks] = {}, rb[kMaxRanks] = {};
            size_t ex[kMaxRanks] = {};
            for (int q = 0; q < W; ++q)
            {
                snd[q] = pc.p + h_lscan[hp.send_begin[q]];
                sb[q] = sizeof(int) * (size_t)(h_lscan[hp.send_begin[q + 1]] - h_lscan[hp.send_begin[q]]);
                rcv[q] = Pext.ci.p + L.P.nnz + h_hscan[seg_start[q]];
                ex[q] = sizeof(int) * (size_t)(h_hscan[seg_start[q + 1]] - h_hscan[seg_start[q]]);
            }
            s_.dist_alltoallv(snd, sb, rcv, rb, ex);
            for (int q = 0; q < W; ++q)
            {
                if (rb[q] != sizeof(int) * (size_t)(h_hscan[seg_start[q + 1]] - h_hscan[seg_start[q]]))
                    throw std::logic_error("psb200 amg: halo plan mismatch between ranks (P rows)");
                snd[q] = pv.p + h_lscan[hp.send_begin[q]];
                sb[q] *= 2;
                rcv[q] = Pext.va.p + L.P.nnz + h_hscan[seg_start[q]];
                ex[q] *= 2;
            }
            s_.dist_alltoallv(snd, sb, rcv, rb, ex);
        }
        lv.t_exchange = wall_ms(st) - tp;

        // ---- Galerkin product, local rows: A_c[my aggregates, :] = R (A_compact P_ext), global coarse columns
        CsrDev Acg;
        {
            CsrDev Ac; // A with compact halo numbering (shares nothing with A: ci is rewritten)
            Ac.n = A.n;
            Ac.ncols = A.n + nh;
            Ac.nnz = A.nnz;
            Ac.rp.alloc((size_t)A.n + 1);
            Ac.ci.alloc(std::max<long long>(1, A.nnz), false, 64);
            Ac.va.alloc(std::max<long long>(1, A.nnz), false, 64);
            PSB_CUDA(cudaMemcpyAsync(Ac.rp.p, A.rp.p, sizeof(int) * ((size_t)A.n + 1), cudaMemcpyDeviceToDevice, st));
            SegTable tb{};
            for (int q = 0; q <= kMaxRanks; ++q)
                tb.seg_start[q] = seg_start[std::min(q, W)];
            if (A.nnz)
            {
                compact_cols_kernel<<<nblk(A.nnz), 256, 0, st>>>(A.nnz, A.ci.p, A.n, D.halo_cap, tb, Ac.ci.p);
                PSB_CUDA(cudaMemcpyAsync(Ac.va.p, A.va.p, sizeof(double) * (size_t)A.nnz, cudaMemcpyDeviceToDevice, st));
            }
            check_launch();
            CsrDev AP;
            tp = wall_ms(st);
            spgemm(ctx, tmp, Ac, Pext, (int)ncg, AP);
            L.t_ap = wall_ms(st) - tp;
            Ac = CsrDev();
            Pext = CsrDev();
            tp = wall_ms(st);
            spgemm(ctx, tmp, L.R, AP, (int)ncg, Acg);
            L.t_rap = wall_ms(st) - tp;
        }
        lv.fc.alloc((size_t)std::max<long long>(4, (L.n_agg + 3) & ~3), true);
        long long cnnz[kMaxRanks];
        s_.dist_gather_ll(Acg.nnz, cnnz);
        const long long cnnz_global = std::accumulate(cnnz, cnnz + W, 0ll);

        const bool replicate = cnnz_global < prm_.replicate_below || li + 2 >= prm_.max_levels || ncg <= prm_.coarse_enough;
        tp = wall_ms(st);
        if (replicate)
        {
            // ---- all-gather the coarse matrix: row lengths, global columns, values; the rest of the hierarchy is replicated
            if (cnnz_global > 0x7fffffffLL - 1024)
                throw std::runtime_error("psb200 amg: replicated coarse level exceeds the int32 index range (lower amg.replicate_below)");
            CsrDev &T = tail_A_;
            T = CsrDev();
            T.n = (int)ncg;
            T.ncols = (int)ncg;
            T.nnz = cnnz_global;
            T.rp.alloc((size_t)ncg + 1);
            T.ci.alloc(std::max<long long>(1, T.nnz), false, 64);
            T.va.alloc(std::max<long long>(1, T.nnz), false, 64);
            DevBuf<int> mylen, alllen, allscan;
            mylen.alloc((size_t)L.n_agg + 1, true);
            alllen.alloc((size_t)ncg + 1, true);
            allscan.alloc((size_t)ncg + 1, true);
            row_diff_kernel<<<nblk(L.n_agg + 1), 256, 0, st>>>(L.n_agg, Acg.rp.p, mylen.p);
            check_launch();
            long long nnz_off[kMaxRanks + 1] = {0};
            for (int q = 0; q < W; ++q)
                nnz_off[q + 1] = nnz_off[q] + cnnz[q];
            auto gatherv = [&](const void *mine, size_t my_bytes, unsigned char *full, const long long *elem_off, size_t elem) {
                const void *snd[kMaxRanks] = {};
                void *rcv[kMaxRanks] = {};
                size_t sb[kMaxRanks] = {}, rb[kMaxRanks] = {}, ex[kMaxRanks] = {};
                for (int q = 0; q < W; ++q)
                {
                    snd[q] = mine;
                    sb[q] = my_bytes;
                    rcv[q] = full + (size_t)elem_off[q] * elem;
                    ex[q] = (size_t)(elem_off[q + 1] - elem_off[q]) * elem;
                }
                s_.dist_alltoallv(snd, sb, rcv, rb, ex);
                for (int q = 0; q < W; ++q)
                    if (rb[q] != (size_t)(elem_off[q + 1] - elem_off[q]) * elem)
                        throw std::logic_error("psb200 amg: all-gather size mismatch between ranks");
            };
            gatherv(mylen.p, sizeof(int) * (size_t)L.n_agg, (unsigned char *)alllen.p, coff.data(), sizeof(int));
            exclusive_scan_int(ctx, tmp, alllen.p, allscan.p, ncg + 1);
            PSB_CUDA(cudaMemcpyAsync(T.rp.p, allscan.p, sizeof(int) * ((size_t)ncg + 1), cudaMemcpyDeviceToDevice, st));
            gatherv(Acg.ci.p, sizeof(int) * (size_t)Acg.nnz, (unsigned char *)T.ci.p, nnz_off, sizeof(int));
            gatherv(Acg.va.p, sizeof(double) * (size_t)Acg.nnz, (unsigned char *)T.va.p, nnz_off, sizeof(double));
            PSB_CUDA(cudaStreamSynchronize(st));
            T.block = B;
            T.plan("auto", st);
            T.refresh_bsr(st);
            tail_offsets_ = coff;
            for (int q = 0; q < W; ++q)
                if (coff[q + 1] - coff[q] > D.halo_cap)
                    throw std::runtime_error("psb200 amg: a rank's slice of the replicated level exceeds halo_cap (raise halo_cap)");
            lv.t_plan = wall_ms(st) - tp;
            lv.t_total = wall_ms(st) - t_lv;
            tail_ = std::make_unique<AmgHierarchy>(ctx, prm_);
            {
                LocalScope local(ctx); // replicated: every rank computes the same numbers, nothing is reduced across ranks
                tail_->setup(T, imposed, li + 1);
            }
            break;
        }

        // ---- the next level stays partitioned: halo plan from the global columns of my coarse rows
        auto next = std::make_unique<DistAmgLevel>();
        next->offsets = coff;
        next->n_global = ncg;
        next->nnz_global = cnnz_global;
        const int lo = (int)coff[me], hi = (int)coff[me + 1];
        DevBuf<int> halo_ids;
        int nh2 = 0;
        {
            DevBuf<unsigned char> flag;
            DevBuf<int> cand, cand_sorted, nsel;
            flag.alloc((size_t)std::max<long long>(1, Acg.nnz));
            cand.alloc((size_t)std::max<long long>(1, Acg.nnz));
            nsel.alloc(4, true);
            if (Acg.nnz)
                flag_offrange_kernel<<<nblk(Acg.nnz), 256, 0, st>>>(Acg.nnz, Acg.ci.p, lo, hi, flag.p);
            check_launch();
            size_t bytes = 0;
            PSB_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, Acg.ci.p, flag.p, cand.p, nsel.p, (int)Acg.nnz, st));
            PSB_CUDA(cub::DeviceSelect::Flagged(tmp.get(bytes), bytes, Acg.ci.p, flag.p, cand.p, nsel.p, (int)Acg.nnz, st));
            const int ncand = d2h(ctx, nsel.p);
            cand_sorted.alloc((size_t)std::max(1, ncand));
            halo_ids.alloc((size_t)std::max(1, ncand));
            if (ncand)
            {
                PSB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, cand.p, cand_sorted.p, ncand, 0, 32, st));
                PSB_CUDA(cub::DeviceRadixSort::SortKeys(tmp.get(bytes), bytes, cand.p, cand_sorted.p, ncand, 0, 32, st));
                PSB_CUDA(cub::DeviceSelect::Unique(nullptr, bytes, cand_sorted.p, halo_ids.p, nsel.p, ncand, st));
                PSB_CUDA(cub::DeviceSelect::Unique(tmp.get(bytes), bytes, cand_sorted.p, halo_ids.p, nsel.p, ncand, st));
                nh2 = d2h(ctx, nsel.p);
            }
        }
        std::vector<int> h_halo(nh2);
        if (nh2)
            PSB_CUDA(cudaMemcpyAsync(h_halo.data(), halo_ids.p, sizeof(int) * nh2, cudaMemcpyDeviceToHost, st));
        PSB_CUDA(cudaStreamSynchronize(st));
        HaloPlan &np = next->plan_own;
        np.world = W;
        np.n_local = coff[me + 1] - coff[me];
        np.recv_count.assign(W, 0);
        {
            int q = 0;
            for (int c : h_halo)
            {
                while (c >= coff[q + 1])
                    ++q;
                np.recv_count[q]++;
            }
        }
        std::vector<int> seg2(W + 1, 0);
        for (int q = 0; q < W; ++q)
        {
            seg2[q + 1] = seg2[q] + np.recv_count[q];
            if (np.recv_count[q] > D.halo_cap)
                throw std::runtime_error("psb200 amg: halo of a coarse level exceeds halo_cap (raise halo_cap in psb200_dist_prepare)");
        }
        // requests: every rank tells the owners which of their coarse unknowns it reads; the owners' send lists follow
        {
            double in[8] = {0, 0, 0, 0, 0, 0, 0, 0}, all[64];
            for (int q = 0; q < W; ++q)
                in[q] = np.recv_count[q];
            s_.dist_allgather8(in, all); // all[s * 8 + q] = number of unknowns rank s reads from rank q
            np.send_begin.assign(W + 1, 0);
            for (int q = 0; q < W; ++q)
                np.send_begin[q + 1] = np.send_begin[q] + (int)all[q * 8 + me];
            const int nreq = np.send_begin[W];
            DevBuf<int> req;
            req.alloc(std::max(1, nreq));
            const void *snd[kMaxRanks] = {};
            void *rcv[kMaxRanks] = {};
            size_t sb[kMaxRanks] = {}, rb[kMaxRanks] = {};
            size_t ex[kMaxRanks] = {};
            for (int q = 0; q < W; ++q)
            {
                snd[q] = halo_ids.p + seg2[q];
                sb[q] = sizeof(int) * (size_t)np.recv_count[q];
                rcv[q] = req.p + np.send_begin[q];
                ex[q] = sizeof(int) * (size_t)(np.send_begin[q + 1] - np.send_begin[q]);
            }
            s_.dist_alltoallv(snd, sb, rcv, rb, ex);
            np.send_rows.assign(nreq, 0);
            if (nreq)
                PSB_CUDA(cudaMemcpyAsync(np.send_rows.data(), req.p, sizeof(int) * nreq, cudaMemcpyDeviceToHost, st));
            PSB_CUDA(cudaStreamSynchronize(st));
            for (int &r : np.send_rows)
            {
                r -= lo;
                if (r < 0 || r >= hi - lo)
                    throw std::logic_error("psb200 amg: a peer requested a coarse unknown this rank does not own");
            }
        }
        // local matrix of the next level: columns remapped to [local | halo]
        CsrDev &An = next->L.Aown;
        An.n = hi - lo;
        An.ncols = hi - lo;
        An.nl = hi - lo;
        An.nnz = Acg.nnz;
        An.halo_mask = nh2 ? 1u : 0u;
        An.rp = std::move(Acg.rp);
        An.va = std::move(Acg.va);
        An.ci.alloc(std::max<long long>(1, An.nnz), false, 64);
        if ((long long)An.n + (long long)W * D.halo_cap > 0x7fffffffLL)
            throw std::runtime_error("psb200 amg: local rows + halo regions exceed the int32 column range");
        {
            SegTable tb{};
            for (int q = 0; q <= kMaxRanks; ++q)
            {
                tb.seg_start[q] = seg2[std::min(q, W)];
                tb.off[q] = coff[std::min(q, W)];
            }
            if (An.nnz)
                remap_cols_kernel<<<nblk(An.nnz), 256, 0, st>>>(An.nnz, Acg.ci.p, W, me, tb, D.halo_cap, halo_ids.p, nh2, An.ci.p);
            check_launch();
            PSB_CUDA(cudaStreamSynchronize(st));
        }
        An.block = B;
        An.plan("auto", st);
        An.refresh_bsr(st);
        next->L.A = &next->L.Aown;
        next->plan = &next->plan_own;
        lv.t_plan = wall_ms(st) - tp;
        lv.t_total = wall_ms(st) - t_lv;
        cur = std::move(next);
    }
    refinalize_plans();
    PSB_CUDA(cudaStreamSynchronize(st));
    t_setup_ms_ = wall_ms(st) - t_begin;
}

// ====================================================================================== cycle
void AmgDist::relax(int l, const double *rhs, double *&x, double *&x_alt, bool x_is_zero, const int *done)
{
    DistAmgLevel &lv = *levels_[l];
    std::function<void(const double *)> push = [this, &lv, done](const double *v) { s_.push_halo(*lv.plan, v, done); };
    if ((prm_.fused_push || env_fused_push()) && s_.dist->world > 1)
    {
        const FusedPush fp{s_.ctx.comm.push_epoch, s_.ctx.comm.halo_expect, s_.ctx.comm.world, lv.plan->push_map()};
        relax_level(s_.ctx, prm_, lv.L, l == 0, rhs, x, x_alt, x_is_zero, done, &push, &fp);
    }
    else
        relax_level(s_.ctx, prm_, lv.L, l == 0, rhs, x, x_alt, x_is_zero, done, &push);
}

// amgcl amg::cycle on the partitioned levels; below them the replicated tail runs one cycle from its level 0
void AmgDist::cycle(int l, const double *rhs, double *&x, double *&x_alt, bool x_is_zero, const int *done)
{
    Ctx &ctx = s_.ctx;
    DistAmgLevel &lv = *levels_[l];
    AmgLevel &L = lv.L;
    const bool has_next = l + 1 < (int)levels_.size();
    if (!has_next && !tail_)
    {
        bool zero = x_is_zero;
        for (int i = 0; i < prm_.npre + prm_.npost; ++i)
        {
            relax(l, rhs, x, x_alt, zero, done);
            zero = false;
        }
        if (zero)
            PSB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * L.n_pad, ctx.stream));
        return;
    }
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
            PSB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * L.n_pad, ctx.stream));
            s_.push_halo(*lv.plan, x, done);
            zero = false;
        }
        launch_spmv(ctx, "spmv_residual", *L.A, x, EpiResidual{L.t.p, rhs}, FinNone{}, done);
        const double *uc = nullptr;
        if (has_next)
        {
            DistAmgLevel &nx = *levels_[l + 1];
            launch_spmv(ctx, "spmv_restrict", L.R, L.t.p, EpiStore{nx.L.f.p}, FinNone{}, done);
            double *nu = nx.L.u.p, *nalt = nx.L.ualt.p;
            cycle(l + 1, nx.L.f.p, nu, nalt, true, done);
            uc = nu;
        }
        else
        {
            launch_spmv(ctx, "spmv_restrict", L.R, L.t.p, EpiStore{lv.fc.p}, FinNone{}, done);
            s_.bulk_allgather(lv.fc.p, tail_->level0_f(), tail_offsets_.data(), done);
            LocalScope local(ctx);
            uc = tail_->cycle0(done) + tail_offsets_[s_.dist->rank];
        }
        launch_spmv(ctx, "spmv_prolong", L.P, uc, EpiAddTo{x}, FinNone{}, done);
        s_.push_halo(*lv.plan, x, done);
        for (int i = 0; i < prm_.npost; ++i)
            relax(l, rhs, x, x_alt, false, done);
    }
}

void AmgDist::apply(const double *rhs, double *out, const int *done)
{
    if (levels_.empty())
        throw std::runtime_error("psb200 amg: empty hierarchy");
    AmgLevel &L0 = levels_[0]->L;
    const size_t bytes = sizeof(double) * (size_t)L0.n_pad;
    if (prm_.pre_cycles <= 0)
    {
        PSB_CUDA(cudaMemcpyAsync(out, rhs, bytes, cudaMemcpyDeviceToDevice, s_.ctx.stream));
        return;
    }
    double *x = L0.u.p, *alt = L0.ualt.p;
    bool zero = true;
    for (int i = 0; i < prm_.pre_cycles; ++i)
    {
        cycle(0, rhs, x, alt, zero, done);
        zero = false;
    }
    PSB_CUDA(cudaMemcpyAsync(out, x, bytes, cudaMemcpyDeviceToDevice, s_.ctx.stream));
}

std::string AmgDist::info_json() const
{
    std::ostringstream o;
    double fine_nnz = levels_.empty() ? 1 : (double)levels_[0]->nnz_global, tot = 0;
    o << "{\"levels\":[";
    for (size_t l = 0; l < levels_.size(); ++l)
    {
        const DistAmgLevel &lv = *levels_[l];
        const AmgLevel &L = lv.L;
        tot += (double)lv.nnz_global;
        if (l)
            o << ",";
        o << "{\"rows\":" << lv.n_global << ",\"nnz\":" << lv.nnz_global << ",\"p_nnz\":" << lv.p_nnz_global << ",\"aggregates\":" << lv.agg_global
          << ",\"partitioned\":true,\"local_rows\":" << L.A->n << ",\"local_nnz\":" << L.A->nnz << ",\"halo_in\":"
          << std::accumulate(lv.plan->recv_count.begin(), lv.plan->recv_count.end(), 0) << ",\"halo_out\":" << lv.plan->send_rows.size()
          << ",\"rho\":" << jnum(L.rho) << ",\"omega\":" << jnum(L.omega) << ",\"mis_rounds\":" << L.mis_rounds << ",\"setup_ms\":{\"relax\":" << jnum(L.t_relax)
          << ",\"aggregate\":" << jnum(L.t_agg) << ",\"prolong\":" << jnum(L.t_prolong) << ",\"transpose\":" << jnum(L.t_transpose)
          << ",\"AP\":" << jnum(L.t_ap) << ",\"RAP\":" << jnum(L.t_rap) << ",\"p_halo_exchange\":" << jnum(lv.t_exchange)
          << ",\"next_level_plan\":" << jnum(lv.t_plan) << ",\"total\":" << jnum(lv.t_total) << "}"
          << ",\"spmv_kernel\":" << jstr(L.A->kernel_name()) << "}";
    }
    if (tail_)
    {
        tot += tail_->total_nnz();
        if (!levels_.empty())
            o << ",";
        o << tail_->levels_json();
    }
    o << "],\"block_size\":" << std::max(1, prm_.block_size) << ",\"operator_complexity\":" << jnum(tot / fine_nnz) << ",\"ncycle\":" << prm_.ncycle
      << ",\"degree\":" << prm_.degree << ",\"relax\":" << jstr(prm_.relax_type) << ",\"partitioned_levels\":" << levels_.size()
      << ",\"replicated_levels\":" << (tail_ ? tail_->num_levels() : 0) << ",\"setup_ms\":" << jnum(t_setup_ms_) << "}";
    return o.str();
}

} // namespace psb
