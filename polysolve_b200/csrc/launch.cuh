// Host-side launch layer: device CSR container, per-solver execution context (stream, reduction
// scratch, profiling) and the SpMV / vector-kernel dispatchers.
#pragma once
#include "kernels.cuh"

#include <map>
#include <string>
#include <vector>

namespace psb {

constexpr int kMaxBlocks = 4096;
constexpr int kVecThreads = 256;
constexpr int kSpmvThreads = 256;
// production tile shape of the stream schedule: rows per tile, staged nnz per tile, pipeline depth
using StreamProd = StreamCfg<256, 2048, 2>;
// long rows (30-80 nnz: Galerkin coarse levels, 3-dof elasticity): LPR lanes per row, 256 / LPR rows per tile
template <int LPR>
using StreamWide = StreamCfg<256, 3072, 2, LPR>;
// the same with smaller stages for more resident CTAs (the gathers of scattered rows are latency bound: ncu shows 37 %
// active warps at 3 CTAs / SM on the level-1 Galerkin matrix): LPR 4 -> 2048 entries (4 CTAs / SM), LPR 8 -> 1536 (6)
template <int LPR>
using StreamNarrow = StreamCfg<256, LPR == 4 ? 2048 : 1536, 2, LPR>;
// in between for 4 lanes per row: 2304 entries (55 KB) still gives 4 CTAs / SM and fits rows of up to ~35 nnz
// (the level-1 Galerkin matrix of the 7-point Laplacian has 33)
using StreamMid4 = StreamCfg<256, 2304, 2, 4>;

// number of tiles of `rows` rows whose staged nnz range would not fit `cap` entries
template <int DUMMY>
__global__ void tile_overflow_kernel(int n, const int *__restrict__ rp, int rows, int cap, int *count)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int ntiles = (n + rows - 1) / rows;
    if (t >= ntiles)
        return;
    const int r0 = t * rows, r1 = min(n, r0 + rows);
    const int ka = rp[r0] & ~3;
    if (((rp[r1] - ka + 3) & ~3) > cap)
        atomicAdd(count, 1);
}

enum SpmvKind : int
{
    SPMV_VECTOR = 0,
    SPMV_STREAM = 1
};
// BSR-3 tile shape: 16 block rows (48 scalar rows) per tile of 128 threads, 8 lanes per block row, 272 staged blocks
// (20.7 KB) per stage -> 5 CTAs / SM. Sweep on B200 (profiles/r02_bsr_sweep.txt, 72^3-node elasticity): 128:272:2 0.0792 ms,
// 256:544:1 0.0806, 256:544:2 0.0870, 512:1088:1 0.0893, 128:272:3 0.0988, 128:288:4 0.1244
using BsrProd = BsrCfg<128, 272, 2, 8>;

// block row pointer / block columns of a scalar CSR whose rows 3 i .. 3 i + 2 share one list of full 3 x 3 blocks
template <int DUMMY>
__global__ void bsr3_pattern_kernel(int nb, const int *__restrict__ rp, const int *__restrict__ ci, int *__restrict__ brp, int *__restrict__ bci)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nb)
        return;
    const int k0 = rp[3 * i];
    brp[i] = k0 / 9;
    if (i == nb)
        return;
    const int len = (rp[3 * i + 1] - k0) / 3;
    for (int q = 0; q < len; ++q)
        bci[k0 / 9 + q] = ci[k0 + 3 * q] / 3;
}
// bva[9 (brp[i] + q) + 3 r + c] = va[rp[3 i + r] + 3 q + c]; one thread per scalar row
template <int DUMMY>
__global__ void bsr3_values_kernel(int n, const int *__restrict__ rp, const double *__restrict__ va, double *__restrict__ bva)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n)
        return;
    const int i = row / 3, r = row % 3;
    const int kb = rp[row], len = rp[row + 1] - kb;
    const size_t base = (size_t)rp[3 * i] + 3 * r; // = 9 brp[i] + 3 r
    for (int e = 0; e < len; ++e)
        bva[base + 9 * (size_t)(e / 3) + e % 3] = va[kb + e];
}

struct CsrDev
{
    int n = 0, ncols = 0;
    long long nnz = 0;
    DevBuf<int> rp, ci;
    DevBuf<double> va;
    int kind = SPMV_VECTOR;
    int lpr = 1; // lanes per row for the vector schedule
    int narrow = 0; // stream schedule with LPR 4 / 8: 0 = 3072-entry stage, 1 = StreamNarrow, 2 = StreamMid4 (LPR 4 only)
    // BSR-3 form of a block-3 matrix (76 B per block instead of 108): kept BESIDE the scalar CSR, which stays the fallback
    // (and what the setup kernels read). block: 3 when rows 3 i .. 3 i + 2 share one list of full 3 x 3 blocks.
    int block = 1;
    bool use_bsr = false, bsr_ready = false;
    DevBuf<int> brp, bci;
    DevBuf<double> bva;
    DevBuf<int> bsr_tile_order; // row partitions: boundary tiles first (tiles of BsrProd::rows block rows)
    bool use_bsr_order = false;
    BsrView bview() const { return BsrView{brp.p, bci.p, bva.p, n / 3, nl, halo_mask, use_bsr_order && bsr_tile_order.p ? bsr_tile_order.p : nullptr}; }
    // (re)builds the BSR arrays from the scalar CSR: call whenever the values have changed
    void refresh_bsr(cudaStream_t st)
    {
        bsr_ready = false;
        if (!use_bsr || block != 3 || n <= 0 || n % 3)
            return;
        const int nb = n / 3;
        brp.alloc((size_t)nb + 1, false, 64);
        bci.alloc(std::max<long long>(1, nnz / 9), false, 64);
        bva.alloc(std::max<long long>(1, nnz), false, 64);
        bsr3_pattern_kernel<0><<<(nb + 256) / 256, 256, 0, st>>>(nb, rp.p, ci.p, brp.p, bci.p);
        bsr3_values_kernel<0><<<(n + 255) / 256, 256, 0, st>>>(n, rp.p, va.p, bva.p);
        bsr_ready = cudaGetLastError() == cudaSuccess;
    }
    int nl = 0x7fffffff;      // local columns (multi-GPU: columns >= nl are halo columns)
    unsigned halo_mask = 0;   // ranks that push halo values to this one
    DevBuf<int> tile_order;   // interior-first tile order of the stream schedule (row partitions only)
    int n_interior = 0, order_rows = 0; // order_rows: rows per tile the order was built for
    bool use_order = false;   // Params::interior_first
    int stream_rows() const { return kSpmvThreads / std::max(1, lpr); }
    CsrView view() const
    {
        const bool ord = use_order && kind == SPMV_STREAM && halo_mask != 0 && tile_order.p != nullptr && order_rows == stream_rows();
        return CsrView{rp.p, ci.p, va.p, n, nl, halo_mask, ord ? tile_order.p : nullptr, ord ? n_interior : 0};
    }
    // Chooses the schedule. auto: the TMA stream schedule with the smallest lanes-per-row whose tiles fit the staging
    // buffers (checked on the device against the actual row pointer: at most 2 % of the tiles may overflow to the
    // direct-load path), else the plain vector schedule. `st` is the stream the row pointer was produced on.
    DevBuf<int> plan_scratch;
    double overflow_frac(int rows, int cap, cudaStream_t st)
    {
        if (n <= 0 || rp.p == nullptr)
            return 0.0;
        plan_scratch.alloc(4, false);
        PSB_CUDA(cudaMemsetAsync(plan_scratch.p, 0, sizeof(int), st));
        const int ntiles = (n + rows - 1) / rows;
        tile_overflow_kernel<0><<<(ntiles + 255) / 256, 256, 0, st>>>(n, rp.p, rows, cap, plan_scratch.p);
        int h = 0;
        PSB_CUDA(cudaMemcpyAsync(&h, plan_scratch.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        PSB_CUDA(cudaStreamSynchronize(st));
        return (double)h / ntiles;
    }
    void plan(const std::string &forced = "auto", cudaStream_t st = nullptr)
    {
        plan_csr(forced == "bsr" ? "auto" : forced, st);
        // BSR-3 on top when the matrix is a block-3 matrix large enough to stream and (almost) all its tiles fit the stage
        use_bsr = false;
        bsr_ready = false;
        if ((forced == "bsr" || forced == "auto") && block == 3 && n % 3 == 0 && n >= 3 * 1024 && rp.p != nullptr)
        {
            // the block row pointer is rp[3 i] / 9: check the tiles on a temporary copy of it
            const int nb = n / 3;
            DevBuf<int> tb, tc;
            tb.alloc((size_t)nb + 1, false, 64);
            tc.alloc(std::max<long long>(1, nnz / 9), false, 64);
            bsr3_pattern_kernel<0><<<(nb + 256) / 256, 256, 0, st>>>(nb, rp.p, ci.p, tb.p, tc.p);
            plan_scratch.alloc(4, false);
            PSB_CUDA(cudaMemsetAsync(plan_scratch.p, 0, sizeof(int), st));
            const int ntiles = (nb + BsrProd::rows - 1) / BsrProd::rows;
            tile_overflow_kernel<0><<<(ntiles + 255) / 256, 256, 0, st>>>(nb, tb.p, BsrProd::rows, BsrProd::capb, plan_scratch.p);
            int h = 0;
            PSB_CUDA(cudaMemcpyAsync(&h, plan_scratch.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            PSB_CUDA(cudaStreamSynchronize(st));
            use_bsr = forced == "bsr" || (double)h / ntiles <= 0.02;
        }
    }
    void plan_csr(const std::string &forced, cudaStream_t st)
    {
        const double avg = n > 0 ? (double)nnz / n : 0;
        const bool want_stream = forced.rfind("stream", 0) == 0;
        narrow = 0;
        if (want_stream && forced.size() > 6 && forced[6] != ':')
        {
            kind = SPMV_STREAM;
            lpr = std::stoi(forced.substr(6)); // "stream8n": stoi stops at the suffix
            narrow = (forced.back() == 'n' && (lpr == 4 || lpr == 8)) ? 1 : (forced.back() == 'm' && lpr == 4) ? 2 : 0;
            return;
        }
        if (want_stream || (forced == "auto" && n >= 4 * StreamProd::threads))
        {
            // candidates in order of preference (measured on B200, profiles/r02_spmv_schedules.txt): the narrow tile shapes
            // (more resident CTAs) beat the 3072-entry stage whenever their tiles fit -- 24.7 nnz/row: stream4n 0.76 of the
            // HBM peak vs stream4 0.67; 43 nnz/row (P1 elasticity): stream8n 0.95 vs stream4 0.89
            struct Cand
            {
                int L, cap;
                int narrow;
            };
            const Cand cands[] = {{1, StreamProd::cap, 0}, {2, StreamWide<2>::cap, 0}, {4, StreamNarrow<4>::cap, 1}, {4, StreamMid4::cap, 2},
                                  {8, StreamNarrow<8>::cap, 1}, {4, StreamWide<4>::cap, 0}, {8, StreamWide<8>::cap, 0},
                                  {16, StreamWide<16>::cap, 0}};
            for (const Cand &cd : cands)
            {
                const int rows = StreamProd::threads / cd.L;
                if (avg * rows + 16 > cd.cap)
                    continue;
                if (overflow_frac(rows, cd.cap, st) > 0.02)
                    continue;
                kind = SPMV_STREAM;
                lpr = cd.L;
                narrow = cd.narrow;
                return;
            }
            if (want_stream)
            {
                kind = SPMV_STREAM; // forced: correct for any matrix (overflowing tiles use direct loads)
                lpr = 1;
                return;
            }
        }
        kind = SPMV_VECTOR;
        if (forced == "scalar")
            lpr = 1;
        else if (avg <= 3)
            lpr = 1;
        else if (avg <= 6)
            lpr = 2;
        else if (avg <= 12)
            lpr = 4;
        else if (avg <= 24)
            lpr = 8;
        else if (avg <= 48)
            lpr = 16;
        else
            lpr = 32;
        if (forced.rfind("vector", 0) == 0 && forced.size() > 6)
            lpr = std::stoi(forced.substr(6));
    }
    std::string kernel_name() const
    {
        if (use_bsr)
            return "bsr3";
        return kind == SPMV_STREAM ? (lpr == 1 ? "stream" : "stream" + std::to_string(lpr) + (narrow == 1 ? "n" : narrow == 2 ? "m" : "")) : "vector" + std::to_string(lpr);
    }
};

struct ProfEntry
{
    double ms = 0;
    long long launches = 0;
};

// Execution context shared by every kernel a solver launches: one stream, one reduction scratch
// (reference precedent: one stream + one pool per solver, MASSolver.cu:154-156,193-195).
struct Ctx
{
    cudaStream_t stream = nullptr;
    DevBuf<double> partials;
    DevBuf<unsigned int> counter;
    long long launches = 0;
    // profiling (events around every launch; only when enabled, never inside graph capture)
    bool profile = false;
    bool capturing = false;
    bool pdl = false; // programmatic dependent launch between the kernels of the Krylov chain (Params::pdl)
    std::map<std::string, ProfEntry> prof;
    struct Pending
    {
        std::string name;
        cudaEvent_t a, b;
    };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> ev_pool;

    void init()
    {
        PSB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        alloc_stream() = stream; // restored by the AllocScope of the C-ABI call we are in
        partials.alloc((size_t)kMaxRed * kMaxBlocks, true);
        counter.alloc(4, true);
    }
    void destroy()
    {
        for (auto &p : pending)
        {
            cudaEventDestroy(p.a);
            cudaEventDestroy(p.b);
        }
        pending.clear();
        for (auto e : ev_pool)
            cudaEventDestroy(e);
        ev_pool.clear();
    }
    Ctx() = default;
    Ctx(const Ctx &) = delete;
    Ctx &operator=(const Ctx &) = delete;
    // The stream outlives every buffer that was allocated on it: Ctx is declared before all buffers of a Solver, so
    // this runs after their destructors have queued their cudaFreeAsync on it.
    ~Ctx()
    {
        destroy();
        partials.release();
        counter.release();
        if (stream)
        {
            cudaStreamSynchronize(stream);
            cudaStreamDestroy(stream);
        }
        stream = nullptr;
    }
    CommDev comm;             // world == 1 unless psb200_dist_connect() was called
    bool comm_local = false;  // true while rank-local work runs (AMG setup / cycle): reductions stay on this GPU
    RedCtx red() const
    {
        RedCtx r{partials.p, counter.p, kMaxBlocks, comm};
        if (comm_local)
            r.comm.world = 1;
        return r;
    }
    cudaEvent_t get_event()
    {
        if (!ev_pool.empty())
        {
            cudaEvent_t e = ev_pool.back();
            ev_pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        PSB_CUDA(cudaEventCreate(&e));
        return e;
    }
    void prof_begin(const char *name)
    {
        ++launches;
        if (!profile || capturing)
            return;
        Pending p{name, get_event(), get_event()};
        PSB_CUDA(cudaEventRecord(p.a, stream));
        pending.push_back(p);
    }
    void prof_end()
    {
        if (!profile || capturing)
            return;
        PSB_CUDA(cudaEventRecord(pending.back().b, stream));
    }
    // call after a stream synchronize
    void prof_collect()
    {
        for (auto &p : pending)
        {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess)
            {
                prof[p.name].ms += ms;
                prof[p.name].launches += 1;
            }
            ev_pool.push_back(p.a);
            ev_pool.push_back(p.b);
        }
        pending.clear();
    }
};

struct LocalScope
{
    Ctx &c;
    bool prev;
    explicit LocalScope(Ctx &ctx) : c(ctx), prev(ctx.comm_local) { c.comm_local = true; }
    ~LocalScope() { c.comm_local = prev; }
};

// Launch on the context's stream; with c.pdl the kernel may overlap its independent prologue with the tail of its
// predecessor (programmatic stream serialization; captured as a programmatic edge inside CUDA graphs). Every kernel
// launched through this helper executes griddep_wait() before it touches anything a predecessor produced.
template <class... KArgs, class... Args>
void launch_chain(Ctx &c, void (*kern)(KArgs...), int grid, int block, size_t smem, Args &&...args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = c.pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PSB_CUDA(cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...));
}

inline void check_launch()
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        throw CudaError(std::string("kernel launch failed: ") + cudaGetErrorString(e));
}

inline int vec_grid(long long n2)
{
    long long need = (n2 + kVecThreads - 1) / kVecThreads;
    need = (need + 1) / 2; // two lanes per thread per trip
    const long long cap = (long long)kSMs * 8;
    return (int)std::max<long long>(1, std::min(need, cap));
}

template <class Op, class Fin>
void launch_vec(Ctx &c, const char *name, long long n_pad, Op op, Fin fin, const int *done = nullptr, const int *only_if = nullptr)
{
    const long long n2 = n_pad / 2;
    if (n2 == 0)
        return;
    c.prof_begin(name);
    launch_chain(c, vec_kernel<Op, Fin, kVecThreads>, vec_grid(n2), kVecThreads, 0, n2, op, c.red(), fin, done, only_if);
    check_launch();
    c.prof_end();
}

template <class Epi, class Fin, int LPR>
void launch_spmv_vector(Ctx &c, const CsrDev &A, const double *x, Epi epi, Fin fin, const int *done, const int *only_if)
{
    const int rows_per_cta = kSpmvThreads / LPR;
    long long need = ((long long)A.n + rows_per_cta - 1) / rows_per_cta;
    const int grid = (int)std::max<long long>(1, std::min<long long>(need, (long long)kSMs * 16));
    launch_chain(c, spmv_vector_kernel<Epi, Fin, LPR, kSpmvThreads>, grid, kSpmvThreads, 0, A.view(), x, epi, c.red(), fin, done, only_if);
}

template <class Epi, class Fin, class Cfg>
void launch_spmv_stream(Ctx &c, const CsrDev &A, const double *x, Epi epi, Fin fin, const int *done, const int *only_if, int ctas_per_sm = 0)
{
    auto kern = spmv_stream_kernel<Epi, Fin, Cfg>;
    // the opt-in to > 48 KB of dynamic shared memory and the occupancy are per device (instances on several GPUs may
    // live in one process)
    static int max_ctas_dev[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int &max_ctas = max_ctas_dev[dev & 15];
    if (!max_ctas)
    {
        PSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::bytes));
        PSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_ctas, kern, Cfg::threads, Cfg::bytes));
        max_ctas = std::max(1, max_ctas);
    }
    const int ntiles = (A.n + Cfg::rows - 1) / Cfg::rows;
    const int per_sm = ctas_per_sm > 0 ? std::min(ctas_per_sm, max_ctas) : max_ctas;
    const int grid = std::min(ntiles, kSMs * per_sm); // persistent: every CTA resident, tiles dealt round-robin
    launch_chain(c, kern, grid, Cfg::threads, Cfg::bytes, A.view(), x, epi, c.red(), fin, done, only_if);
}

template <class Epi, class Fin, class Cfg = BsrProd>
void launch_spmv_bsr3(Ctx &c, const CsrDev &A, const double *x, Epi epi, Fin fin, const int *done, const int *only_if)
{
    auto kern = spmv_bsr3_kernel<Epi, Fin, Cfg>;
    static int max_ctas_dev[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int &max_ctas = max_ctas_dev[dev & 15];
    if (!max_ctas)
    {
        PSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::bytes));
        PSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_ctas, kern, Cfg::threads, Cfg::bytes));
        max_ctas = std::max(1, max_ctas);
    }
    const int nb = A.n / 3;
    const int ntiles = (nb + Cfg::rows - 1) / Cfg::rows;
    const int grid = std::min(ntiles, kSMs * max_ctas);
    launch_chain(c, kern, grid, Cfg::threads, Cfg::bytes, A.bview(), x, epi, c.red(), fin, done, only_if);
}

template <class Epi, class Fin>
void launch_spmv(Ctx &c, const char *name, const CsrDev &A, const double *x, Epi epi, Fin fin, const int *done = nullptr,
                 const int *only_if = nullptr)
{
    if (A.n == 0)
        return;
    c.prof_begin(name);
    if (A.use_bsr && A.bsr_ready)
        launch_spmv_bsr3<Epi, Fin>(c, A, x, epi, fin, done, only_if);
    else if (A.kind == SPMV_STREAM)
    {
        switch (A.lpr)
        {
        case 2: launch_spmv_stream<Epi, Fin, StreamWide<2>>(c, A, x, epi, fin, done, only_if); break;
        case 4:
            if (A.narrow == 1)
                launch_spmv_stream<Epi, Fin, StreamNarrow<4>>(c, A, x, epi, fin, done, only_if);
            else if (A.narrow == 2)
                launch_spmv_stream<Epi, Fin, StreamMid4>(c, A, x, epi, fin, done, only_if);
            else
                launch_spmv_stream<Epi, Fin, StreamWide<4>>(c, A, x, epi, fin, done, only_if);
            break;
        case 8:
            if (A.narrow)
                launch_spmv_stream<Epi, Fin, StreamNarrow<8>>(c, A, x, epi, fin, done, only_if);
            else
                launch_spmv_stream<Epi, Fin, StreamWide<8>>(c, A, x, epi, fin, done, only_if);
            break;
        case 16: launch_spmv_stream<Epi, Fin, StreamWide<16>>(c, A, x, epi, fin, done, only_if); break;
        default: launch_spmv_stream<Epi, Fin, StreamProd>(c, A, x, epi, fin, done, only_if); break;
        }
    }
    else
    {
        switch (A.lpr)
        {
        case 1: launch_spmv_vector<Epi, Fin, 1>(c, A, x, epi, fin, done, only_if); break;
        case 2: launch_spmv_vector<Epi, Fin, 2>(c, A, x, epi, fin, done, only_if); break;
        case 4: launch_spmv_vector<Epi, Fin, 4>(c, A, x, epi, fin, done, only_if); break;
        case 8: launch_spmv_vector<Epi, Fin, 8>(c, A, x, epi, fin, done, only_if); break;
        case 16: launch_spmv_vector<Epi, Fin, 16>(c, A, x, epi, fin, done, only_if); break;
        default: launch_spmv_vector<Epi, Fin, 32>(c, A, x, epi, fin, done, only_if); break;
        }
    }
    check_launch();
    c.prof_end();
}

} // namespace psb
