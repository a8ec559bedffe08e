// Host driver of the B200 linear-solver backend: ingest (analyze_pattern / factorize), Jacobi-PCG in
// Eigen's ordering, BiCGSTAB, AMG-PCG in AMGCL's ordering. Mirrors the call protocol of
// polysolve::linear::Solver (reference src/polysolve/linear/Solver.hpp:90-131) one method per virtual.
#include "solver.hpp"
#include "amg.hpp"
#include "dist.hpp"
#include "amg_dist.hpp"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#include <thread>
#include <cstring>
#include <functional>
#include <sstream>

namespace psb {

namespace {

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// 64-bit hash of the pattern arrays (4 independent lanes so it runs at memory speed on one core).
unsigned long long hash_chunk(const void *data, size_t bytes, unsigned long long seed)
{
    const unsigned long long *w = (const unsigned long long *)data;
    const size_t nw = bytes / 8;
    unsigned long long h[4] = {seed ^ 0x9E3779B97F4A7C15ull, seed ^ 0xBF58476D1CE4E5B9ull, seed ^ 0x94D049BB133111EBull, seed ^ 0xD6E8FEB86659FD93ull};
    size_t i = 0;
    for (; i + 4 <= nw; i += 4)
        for (int l = 0; l < 4; ++l)
        {
            unsigned long long v = w[i + l] * 0xFF51AFD7ED558CCDull;
            v ^= v >> 32;
            h[l] = (h[l] ^ v) * 0xC4CEB9FE1A85EC53ull;
            h[l] = (h[l] << 27) | (h[l] >> 37);
        }
    unsigned long long r = h[0] ^ (h[1] * 3) ^ (h[2] * 5) ^ (h[3] * 7);
    const unsigned char *tail = (const unsigned char *)data + i * 8;
    for (size_t k = i * 8; k < bytes; ++k)
        r = (r ^ *tail++) * 0x100000001B3ull;
    r ^= r >> 29;
    return r * 0xBF58476D1CE4E5B9ull;
}

// Fixed 8 MiB chunks hashed by up to 8 host threads and combined in chunk order, so the result does not
// depend on the thread count. Newton calls analyze_pattern + factorize every iteration (Newton.cpp:189-191):
// the unchanged-pattern check must cost a few ms, not a pass of one core over 300 MB.
unsigned long long hash_words(const void *data, size_t bytes, unsigned long long seed)
{
    constexpr size_t kChunk = 8u << 20;
    const size_t nchunks = (bytes + kChunk - 1) / kChunk;
    if (nchunks <= 1)
        return hash_chunk(data, bytes, seed);
    std::vector<unsigned long long> part(nchunks);
    const unsigned nthreads = (unsigned)std::min<size_t>(nchunks, std::max(1u, std::min(8u, std::thread::hardware_concurrency())));
    auto work = [&](unsigned t) {
        for (size_t c = t; c < nchunks; c += nthreads)
        {
            const size_t off = c * kChunk;
            part[c] = hash_chunk((const unsigned char *)data + off, std::min(kChunk, bytes - off), seed + 0x9E3779B97F4A7C15ull * (c + 1));
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nthreads; ++t)
        th.emplace_back(work, t);
    work(0);
    for (auto &t : th)
        t.join();
    return hash_chunk(part.data(), part.size() * 8, seed);
}

// ------------------------------------------------------------------ ingest kernels (analyze_pattern)
// Symmetric-pattern fast path. If the pattern is symmetric and sorted, CSR(row_ptr, col_idx) ==
// CSC(outer, inner) as integer arrays and only the value map is needed:
//   perm[k] (k in row i, col j = inner[k]) = position of row index i inside column j.
// 8 lanes per row; binary search in column j. Any miss / unsorted column clears *ok.
__global__ void sym_perm_kernel(int n, const int *__restrict__ outer, const int *__restrict__ inner, int *__restrict__ perm, int *ok)
{
    const int lane = threadIdx.x & 7;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (row >= n)
        return;
    const int kb = outer[row], ke = outer[row + 1];
    bool good = true;
    for (int k = kb + lane; k < ke; k += 8)
    {
        const int j = inner[k];
        if (k > kb && inner[k - 1] >= j)
            good = false; // not strictly ascending
        if (j < 0 || j >= n)
        {
            good = false;
            continue;
        }
        int lo = outer[j], hi = outer[j + 1];
        while (lo < hi)
        {
            const int mid = (lo + hi) >> 1;
            if (inner[mid] < (int)row)
                lo = mid + 1;
            else
                hi = mid;
        }
        if (lo < outer[j + 1] && inner[lo] == (int)row)
            perm[k] = lo;
        else
            good = false;
    }
    if (!good)
        atomicExch(ok, 0);
}

// General path helpers: column id of every CSC entry, iota, row_ptr from the sorted row keys.
__global__ void expand_cols_kernel(int ncols, const int *__restrict__ outer, int *__restrict__ col_of)
{
    const int lane = threadIdx.x & 7;
    const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (c >= ncols)
        return;
    for (int k = outer[c] + lane; k < outer[c + 1]; k += 8)
        col_of[k] = (int)c;
}
__global__ void iota_kernel(long long n, int *a)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        a[i] = (int)i;
}
__global__ void row_ptr_from_sorted_kernel(int nrows, long long nnz, const int *__restrict__ sorted_rows, int *__restrict__ row_ptr)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nrows)
        return;
    long long lo = 0, hi = nnz;
    while (lo < hi)
    {
        const long long mid = (lo + hi) >> 1;
        if (sorted_rows[mid] < (int)i)
            lo = mid + 1;
        else
            hi = mid;
    }
    row_ptr[i] = (int)lo;
}
__global__ void gather_int_kernel(long long n, const int *__restrict__ src, const int *__restrict__ idx, int *__restrict__ dst)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        dst[i] = src[idx[i]];
}

// ------------------------------------------------------------------ block expansion (block_size B > 1)
// amgcl::adapter::block_matrix (reference AMGCL.cpp:270-272) views the scalar matrix as B x B blocks: block row I holds
// rows B I .. B I + B - 1 and the union of their block columns, missing scalar entries being zero. The same view is
// made explicit here once per pattern: every block row gets the full B x B pattern (perm = -1 marks the fill-in), so
// all B rows of a node share one column list -- the invariant the block AMG kernels rely on.
// Walks the B sorted rows of node `node` in merged order; calls f(q, jb) for the q-th distinct block column jb.
template <class F>
__device__ __forceinline__ int merged_block_cols(int B, int node, const int *__restrict__ rp, const int *__restrict__ ci, F f)
{
    int head[4];
    for (int r = 0; r < B; ++r)
        head[r] = rp[B * node + r];
    int q = 0;
    for (;;)
    {
        int jb = 0x7fffffff;
        for (int r = 0; r < B; ++r)
            if (head[r] < rp[B * node + r + 1])
                jb = min(jb, ci[head[r]] / B);
        if (jb == 0x7fffffff)
            break;
        f(q, jb);
        for (int r = 0; r < B; ++r)
            while (head[r] < rp[B * node + r + 1] && ci[head[r]] / B == jb)
                ++head[r];
        ++q;
    }
    return q;
}
__global__ void block_count_kernel(int B, int nb, const int *__restrict__ rp, const int *__restrict__ ci, int *__restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nb)
        return;
    cnt[i] = i < nb ? merged_block_cols(B, i, rp, ci, [](int, int) {}) : 0;
}
__global__ void block_fill_kernel(int B, int nb, const int *__restrict__ rp, const int *__restrict__ ci, const int *__restrict__ perm,
                                  const int *__restrict__ boff, int *__restrict__ nrp, int *__restrict__ nci, int *__restrict__ nperm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nb)
        return;
    if (i == nb)
    {
        nrp[B * nb] = B * B * boff[nb];
        return;
    }
    const int len = boff[i + 1] - boff[i];
    const int base = B * B * boff[i];
    for (int r = 0; r < B; ++r)
        nrp[B * i + r] = base + r * B * len;
    merged_block_cols(B, i, rp, ci, [&](int q, int jb) {
        for (int r = 0; r < B; ++r)
            for (int c = 0; c < B; ++c)
            {
                nci[base + r * B * len + B * q + c] = B * jb + c;
                nperm[base + r * B * len + B * q + c] = -1;
            }
    });
    // scatter the existing entries: both lists are sorted, so one forward walk per row finds the block slot
    for (int r = 0; r < B; ++r)
    {
        int q = 0;
        const int rowbase = base + r * B * len;
        for (int k = rp[B * i + r]; k < rp[B * i + r + 1]; ++k)
        {
            const int j = ci[k];
            while (nci[rowbase + B * q] / B != j / B)
                ++q;
            nperm[rowbase + B * q + j % B] = perm[k];
        }
    }
}

// ------------------------------------------------------------------ factorize kernels
// vals_csr[k] = vals_csc[perm[k]]
__global__ void gather_vals_kernel(long long n, const double *__restrict__ src, const int *__restrict__ perm, double *__restrict__ dst)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        const int p = __ldg(perm + i);
        dst[i] = p >= 0 ? __ldg(src + p) : 0.0; // p < 0: explicit zero added by the block expansion
    }
}
// Eigen::DiagonalPreconditioner::factorize: invdiag = (A_ii != 0) ? 1/A_ii : 1. mode 0: ones.
__global__ void inv_diag_kernel(CsrView A, double *__restrict__ dinv, int mode, int *bad)
{
    const int lane = threadIdx.x & 7;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (row >= A.n)
        return;
    double d = 0;
    bool found = false;
    for (int k = A.rp[row] + lane; k < A.rp[row + 1]; k += 8)
        if (A.ci[k] == (int)row)
        {
            d = A.va[k];
            found = true;
        }
    // combine across the 8 lanes
    for (int o = 4; o > 0; o >>= 1)
    {
        const double od = __shfl_xor_sync(0xffffffffu, d, o);
        const bool of = __shfl_xor_sync(0xffffffffu, (int)found, o);
        if (of && !found)
        {
            d = od;
            found = true;
        }
    }
    if (lane == 0)
    {
        if (!(d == d) || isinf(d))
            atomicExch(bad, 1);
        dinv[row] = mode == 0 ? 1.0 : ((found && d != 0.0) ? 1.0 / d : 1.0);
    }
}

// A_ii += shift (8 lanes per row; rows of a row partition hold their diagonal at local column i)
__global__ void shift_diag_kernel(CsrView A, double *__restrict__ va, double shift, int *missing)
{
    const int lane = threadIdx.x & 7;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (row >= A.n)
        return;
    bool found = false;
    for (int k = A.rp[row] + lane; k < A.rp[row + 1]; k += 8)
        if (A.ci[k] == (int)row)
        {
            va[k] += shift;
            found = true;
        }
    for (int o = 4; o > 0; o >>= 1)
        found |= (bool)__shfl_xor_sync(0xffffffffu, (int)found, o);
    if (lane == 0 && !found)
        atomicExch(missing, 1);
}

inline int blocks_for(long long n, int threads) { return (int)std::max<long long>(1, (n + threads - 1) / threads); }

} // namespace

// Sorts every row of a CSR pattern by column (insertion sort, one thread per row; rows are short and nearly sorted: the
// local rows of a row partition are three ascending runs after the [local | halo] column remap). perm moves along.
__global__ void sort_rows_kernel(int n, const int *__restrict__ rp, int *__restrict__ ci, int *__restrict__ perm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int kb = rp[i], ke = rp[i + 1];
    for (int k = kb + 1; k < ke; ++k)
    {
        const int c = ci[k], p = perm[k];
        int q = k - 1;
        while (q >= kb && ci[q] > c)
        {
            ci[q + 1] = ci[q];
            perm[q + 1] = perm[q];
            --q;
        }
        ci[q + 1] = c;
        perm[q + 1] = p;
    }
}
void sort_rows_by_column(Ctx &ctx, long long n, DevBuf<int> &rp, DevBuf<int> &ci, DevBuf<int> &perm)
{
    if (n <= 0)
        return;
    sort_rows_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx.stream>>>((int)n, rp.p, ci.p, perm.p);
    check_launch();
}

// Full-block expansion of a sorted scalar CSR pattern (see the comment above merged_block_cols): rp / ci / perm are
// replaced by the expanded arrays (perm = -1 marks fill-in); returns the new nnz. n must be a multiple of B.
long long expand_block_pattern(Ctx &ctx, int B, long long n, DevBuf<int> &rp, DevBuf<int> &ci, DevBuf<int> &perm)
{
    cudaStream_t st = ctx.stream;
    if (B > 3)
        throw std::invalid_argument("psb200: block_size must be 1, 2 or 3 (reference AMGCL.cpp:111-123)");
    if (n % B != 0)
        throw std::invalid_argument("psb200: the matrix size is not a multiple of block_size");
    const int nb = (int)(n / B);
    DevBuf<int> cnt, boff;
    cnt.alloc((size_t)nb + 1, true);
    boff.alloc((size_t)nb + 1);
    block_count_kernel<<<blocks_for(nb + 1, 256), 256, 0, st>>>(B, nb, rp.p, ci.p, cnt.p);
    check_launch();
    size_t tmp_bytes = 0;
    PSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt.p, boff.p, nb + 1, st));
    DevBuf<unsigned char> tmp;
    tmp.alloc(tmp_bytes);
    PSB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cnt.p, boff.p, nb + 1, st));
    int nblocks = 0;
    PSB_CUDA(cudaMemcpyAsync(&nblocks, boff.p + nb, sizeof(int), cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaStreamSynchronize(st));
    const long long nnz_full = (long long)nblocks * B * B;
    if (nnz_full > 0x7fffffffLL - 1024)
        throw std::invalid_argument("psb200: int32 index range exceeded after block expansion");
    DevBuf<int> nrp, nci, nperm;
    nrp.alloc(n + 1);
    nci.alloc(std::max<long long>(nnz_full, 1), false, 64);
    nperm.alloc(std::max<long long>(nnz_full, 1));
    block_fill_kernel<<<blocks_for(nb + 1, 256), 256, 0, st>>>(B, nb, rp.p, ci.p, perm.p, boff.p, nrp.p, nci.p, nperm.p);
    check_launch();
    PSB_CUDA(cudaStreamSynchronize(st));
    rp = std::move(nrp);
    ci = std::move(nci);
    perm = std::move(nperm);
    return nnz_full;
}

// ==================================================================================== lifecycle
Solver::Solver() {}

Solver::~Solver()
{
    if (ctx.stream)
        cudaStreamSynchronize(ctx.stream); // kernels in flight may still read the helper instance's matrix (full_)
    amg.reset();
    amg_dist.reset();
    if (graph_exec)
        cudaGraphExecDestroy(graph_exec);
    for (auto &e : ev)
        if (e)
            cudaEventDestroy(e);
    if (d_state)
        cudaFree(d_state);
    if (h_state)
        cudaFreeHost(h_state);
    // the stream is destroyed by ~Ctx, after the buffers declared below it have been returned to the pool
}

void ensure_ctx(Solver &s)
{
    if (s.ctx.stream)
        return;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw CudaError(std::string("psb200: no CUDA device available (") + cudaGetErrorString(e) +
                        "); the CUDA backend has no CPU fallback");
    if (s.prm.device >= 0)
        PSB_CUDA(cudaSetDevice(s.prm.device));
    PSB_CUDA(cudaGetDevice(&s.device));
    {
        // keep freed blocks cached in the device's stream-ordered pool (DevBuf allocates from it)
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, s.device) == cudaSuccess && pool)
        {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        else
            cudaGetLastError();
    }
    s.ctx.init();
    PSB_CUDA(cudaMalloc(&s.d_state, sizeof(KState)));
    PSB_CUDA(cudaMemset(s.d_state, 0, sizeof(KState)));
    PSB_CUDA(cudaMallocHost(&s.h_state, 4 * sizeof(KState)));
    PSB_CUDA(cudaEventCreateWithFlags(&s.ev[0], cudaEventDisableTiming));
    PSB_CUDA(cudaEventCreateWithFlags(&s.ev[1], cudaEventDisableTiming));
}

// ==================================================================================== parameters
static void read_amg(const JValue &j, AmgParams &a)
{
    auto num = [&](const JValue &o, const char *k, auto &dst) {
        if (o.contains(k))
            dst = (std::remove_reference_t<decltype(dst)>)o.at(k).as_num();
    };
    // flat keys
    num(j, "max_levels", a.max_levels);
    num(j, "coarse_enough", a.coarse_enough);
    num(j, "ncycle", a.ncycle);
    num(j, "npre", a.npre);
    num(j, "npost", a.npost);
    num(j, "pre_cycles", a.pre_cycles);
    if (j.contains("direct_coarse"))
        a.direct_coarse = j.at("direct_coarse").as_bool();
    if (j.contains("aggregation"))
        a.aggregation = j.at("aggregation").as_str();
    if (j.contains("dist_mode"))
    {
        a.dist_mode = j.at("dist_mode").as_str();
        if (a.dist_mode != "partitioned" && a.dist_mode != "global" && a.dist_mode != "local")
            throw std::runtime_error("psb200: amg dist_mode must be partitioned, global or local");
    }
    num(j, "replicate_below", a.replicate_below);
    if (j.contains("fused_push"))
        a.fused_push = j.at("fused_push").as_bool();
    if (a.aggregation != "mis2")
        throw std::runtime_error("psb200: unknown amg aggregation '" + a.aggregation + "' (mis2)");
    // AMGCL-shaped sub-objects (AMGCL.cpp:32-65): relax{type,degree,power_iters,higher,lower,scale}, coarsening{relax,estimate_spectral_radius,aggr{eps_strong}}
    if (j.contains("relax") && j.at("relax").is_obj())
    {
        const JValue &r = j.at("relax");
        if (r.contains("type"))
            a.relax_type = r.at("type").as_str();
        num(r, "degree", a.degree);
        num(r, "power_iters", a.power_iters);
        num(r, "higher", a.higher);
        num(r, "lower", a.lower);
        num(r, "damping", a.damping);
        if (r.contains("scale"))
            a.scale = r.at("scale").as_bool();
    }
    if (j.contains("coarsening") && j.at("coarsening").is_obj())
    {
        const JValue &c = j.at("coarsening");
        num(c, "relax", a.sa_relax);
        if (c.contains("estimate_spectral_radius"))
            a.estimate_spectral_radius = c.at("estimate_spectral_radius").as_bool();
        if (c.contains("aggr") && c.at("aggr").is_obj())
            num(c.at("aggr"), "eps_strong", a.eps_strong);
    }
    if (a.relax_type != "chebyshev" && a.relax_type != "damped_jacobi")
        throw std::runtime_error("psb200: unsupported amg relax type '" + a.relax_type + "'");
}

// names CsrDev::plan understands (validated before a parameter document is committed)
static bool spmv_kernel_name_ok(const std::string &k)
{
    if (k == "auto" || k == "scalar" || k == "stream" || k == "bsr")
        return true;
    for (const char *pre : {"stream", "vector"})
        if (k.rfind(pre, 0) == 0)
        {
            std::string num = k.substr(6);
            if (pre[0] == 's' && (num == "4n" || num == "8n" || num == "4m"))
                return true; // narrow tile shapes of the stream schedule
            if (num.empty() || num.size() > 2 || num.find_first_not_of("0123456789") != std::string::npos)
                return false;
            const int v = std::stoi(num);
            const bool pow2 = v > 0 && (v & (v - 1)) == 0;
            return pow2 && v <= (pre[0] == 's' ? 16 : 32);
        }
    return false;
}

void Solver::set_parameters(const std::string &json)
{
    if (json.empty())
        return;
    JValue doc = JParser::parse(json);
    if (!doc.is_obj() || !doc.contains("CUDA"))
        return; // parameters of other solvers are ignored, like every polysolve backend does
    const JValue &j = doc.at("CUDA");
    if (!j.is_obj())
        throw std::runtime_error("psb200: \"CUDA\" must be an object");
    Params np = prm; // transactional: a rejected document leaves the solver unchanged
    if (j.contains("krylov"))
        np.krylov = j.at("krylov").as_str();
    if (j.contains("precond"))
        np.precond = j.at("precond").as_str();
    if (j.contains("tolerance"))
        np.tolerance = j.at("tolerance").as_num();
    if (j.contains("max_iter"))
        np.max_iter = (int)j.at("max_iter").as_num();
    if (j.contains("check_every"))
        np.check_every = std::max(1, (int)j.at("check_every").as_num());
    if (j.contains("use_graph"))
        np.use_graph = j.at("use_graph").as_bool();
    if (j.contains("interior_first"))
        np.interior_first = j.at("interior_first").as_bool();
    if (j.contains("pdl"))
        np.pdl = j.at("pdl").as_bool();
    if (j.contains("spmv_kernel"))
        np.spmv_kernel = j.at("spmv_kernel").as_str();
    if (j.contains("cg_kernel"))
        np.cg_kernel = j.at("cg_kernel").as_str();
    if (j.contains("device"))
        np.device = (int)j.at("device").as_num();
    if (j.contains("block_size"))
        np.block_size = (int)j.at("block_size").as_num();
    if (j.contains("profile"))
        np.profile = j.at("profile").as_bool();
    if (j.contains("verify_pattern"))
        np.verify_pattern = j.at("verify_pattern").as_bool();
    if (j.contains("comm_timeout_s"))
        np.comm_timeout_s = j.at("comm_timeout_s").as_num();
    if (j.contains("amg") && j.at("amg").is_obj())
        read_amg(j.at("amg"), np.amg);
    if (np.krylov != "cg" && np.krylov != "cg1r" && np.krylov != "bicgstab")
        throw std::runtime_error("psb200: unknown krylov '" + np.krylov + "' (cg | cg1r | bicgstab)");
    if (np.krylov == "cg1r" && np.precond == "amg")
        throw std::runtime_error("psb200: cg1r is the single-reduction form of Jacobi-PCG (precond jacobi | none); AMG-PCG uses krylov=cg");
    if (!(np.comm_timeout_s >= 0))
        throw std::runtime_error("psb200: comm_timeout_s must be >= 0");
    if (!spmv_kernel_name_ok(np.spmv_kernel))
        throw std::runtime_error("psb200: unknown spmv_kernel '" + np.spmv_kernel + "' (auto | stream | stream<2|4|8|16> | stream4n | stream4m | stream8n | vector<1|2|4|8|16|32> | scalar | bsr)");
    if (np.precond != "jacobi" && np.precond != "amg" && np.precond != "none")
        throw std::runtime_error("psb200: unknown precond '" + np.precond + "' (jacobi | amg | none)");
    if (np.cg_kernel != "auto" && np.cg_kernel != "split")
        throw std::runtime_error("psb200: unknown cg_kernel '" + np.cg_kernel + "' (auto | split)");
    if (np.krylov == "bicgstab" && np.precond == "amg")
        throw std::runtime_error("psb200: bicgstab + amg is not available yet");
    // everything is validated: commit. A change of the preconditioner (or of anything the hierarchy / D^-1 was built from)
    // invalidates the factorization: solve() then asks for a new factorize() instead of silently using the old one.
    const bool refactor = np.precond != prm.precond || np.block_size != prm.block_size || !np.amg.same_as(prm.amg);
    prm = np;
    if (refactor && factorized)
    {
        factorized = false;
        amg.reset();
        amg_dist.reset();
    }
    if (dist && dist->connected)
        ctx.comm.spin_limit = (long long)std::max(1.0, prm.comm_timeout_s * 1.9e9);
    if (graph_exec)
    {
        cudaGraphExecDestroy(graph_exec);
        graph_exec = nullptr;
        graph_key.clear();
    }
    if (analyzed)
    {
        A.plan(prm.spmv_kernel, ctx.stream);
        if (factorized)
            A.refresh_bsr(ctx.stream);
    }
    ctx.profile = prm.profile;
    ctx.pdl = prm.pdl;
    A.use_order = prm.interior_first;
}

// Hash of the index arrays. On a row partition of W ranks every rank hashes only the W-th part of both arrays that
// corresponds to its rank (the parts cover the arrays; the ranks combine their verdicts, see analyze_pattern), so the
// per-factorize pattern check costs 1/W of a pass over the indices per rank instead of a full pass on every rank.
unsigned long long Solver::pattern_hash_of(long long n_, long long nnz_, const int *outer, const int *inner) const
{
    const unsigned long long seed = 0x5bd1e995ull + (unsigned long long)n_;
    if (dist && dist->connected && dist->world > 1)
    {
        const long long W = dist->world, r = dist->rank;
        const long long o0 = (n_ + 1) * r / W, o1 = (n_ + 1) * (r + 1) / W, i0 = nnz_ * r / W, i1 = nnz_ * (r + 1) / W;
        unsigned long long h = hash_words(outer + o0, sizeof(int) * (size_t)(o1 - o0), seed + (unsigned long long)nnz_);
        return hash_words(inner + i0, sizeof(int) * (size_t)(i1 - i0), h);
    }
    unsigned long long h = hash_words(outer, sizeof(int) * (size_t)(n_ + 1), seed);
    return hash_words(inner, sizeof(int) * (size_t)nnz_, h);
}

// ==================================================================================== analyze_pattern
void Solver::analyze_pattern(long long n_, long long nnz_, const int *outer, const int *inner, int precond_num_)
{
    if (n_ < 0 || nnz_ < 0 || !outer || (!inner && nnz_ > 0))
        throw std::invalid_argument("psb200_analyze_pattern_csc: null or negative argument");
    if (n_ > 0x7fffffffLL - 1024 || nnz_ > 0x7fffffffLL - 1024)
        throw std::invalid_argument("psb200_analyze_pattern_csc: int32 index range exceeded (reference limit, BSRMatrix.cu:439-442)");
    if (n_ > 0 && (outer[0] != 0 || outer[n_] != nnz_))
        throw std::invalid_argument("psb200_analyze_pattern_csc: matrix is not compressed (outer[0] != 0 or outer[n] != nnz); call makeCompressed() first (cf. BSRMatrix.cu:444-452)");
    // O(n) host check before any kernel indexes `inner` with these values (an out-of-range read would poison the context)
    for (long long c = 0; c < n_; ++c)
        if (outer[c] > outer[c + 1] || outer[c] < 0)
            throw std::invalid_argument("psb200_analyze_pattern_csc: outer index array is not non-decreasing");
    ensure_ctx(*this);
    const double t0 = now_ms();
    precond_num = precond_num_;
    const unsigned long long h = pattern_hash_of(n_, nnz_, outer, inner);
    bool same = analyzed && n_global == n_ && nnz_global == nnz_ && h == pattern_hash && pattern_block == std::max(1, prm.block_size);
    if (dist && dist->connected && dist->world > 1)
        same = dist_max(same ? 0.0 : 1.0) == 0.0; // collective decision: a change seen by any rank re-analyses on all
    if (same)
    {
        analyze_skipped = true; // Newton calls analyze_pattern every iteration with an unchanged pattern (Newton.cpp:189)
        t_analyze_ms = now_ms() - t0;
        return;
    }
    analyze_skipped = false;
    analyzed = false;
    factorized = false;
    amg.reset();
    amg_dist.reset();
    if (graph_exec)
    {
        cudaGraphExecDestroy(graph_exec);
        graph_exec = nullptr;
        graph_key.clear();
    }
    n_global = n_;
    nnz_global = nnz_;
    pattern_hash = h;
    if (dist)
    {
        // row-partitioned mode: build the local part on the host (dist.cu) and stop here
        analyze_pattern_dist(n_, nnz_, outer, inner);
        if (amg_global())
        {
            if (!full_)
                full_ = std::make_unique<Solver>();
            full_->prm = prm;
            full_->prm.precond = "jacobi";
            full_->prm.krylov = "cg";
            full_->prm.verify_pattern = false; // this instance has just hashed the same arrays
            full_->prm.device = device;
            AllocScope guard(full_->ctx.stream);
            full_->analyze_pattern(n_, nnz_, outer, inner, precond_num_);
        }
        else
            full_.reset();
        analyzed = true;
        t_analyze_ms = now_ms() - t0;
        return;
    }
    n = n_;
    nnz = nnz_;
    n_pad = (n + 3) & ~3ll;
    cudaStream_t st = ctx.stream;

    csc_outer.alloc(n + 1);
    csc_inner.alloc(std::max<long long>(nnz, 1));
    perm.alloc(std::max<long long>(nnz, 1));
    PSB_CUDA(cudaMemcpyAsync(csc_outer.p, outer, sizeof(int) * (n + 1), cudaMemcpyHostToDevice, st));
    if (nnz)
        PSB_CUDA(cudaMemcpyAsync(csc_inner.p, inner, sizeof(int) * nnz, cudaMemcpyHostToDevice, st));

    A.n = (int)n;
    A.ncols = (int)n;
    A.nnz = nnz;
    A.rp.alloc(n + 1);
    A.ci.alloc(std::max<long long>(nnz, 1), false, 64);
    A.va.alloc(std::max<long long>(nnz, 1), false, 64);

    // 1) symmetric-pattern fast path
    int *d_ok = (int *)ctx.counter.p + 2;
    int one = 1;
    PSB_CUDA(cudaMemcpyAsync(d_ok, &one, sizeof(int), cudaMemcpyHostToDevice, st));
    if (n > 0)
    {
        sym_perm_kernel<<<blocks_for(n * 8, 256), 256, 0, st>>>((int)n, csc_outer.p, csc_inner.p, perm.p, d_ok);
        check_launch();
    }
    int ok = 0;
    PSB_CUDA(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaStreamSynchronize(st));
    sym_pattern = ok != 0;
    if (sym_pattern)
    {
        PSB_CUDA(cudaMemcpyAsync(A.rp.p, csc_outer.p, sizeof(int) * (n + 1), cudaMemcpyDeviceToDevice, st));
        if (nnz)
            PSB_CUDA(cudaMemcpyAsync(A.ci.p, csc_inner.p, sizeof(int) * nnz, cudaMemcpyDeviceToDevice, st));
    }
    else if (nnz > 0)
    {
        // 2) general path: stable radix sort of the CSC entries by row index (the reference's own ingest
        //    sorts with cub too, mas_utils/BSRMatrix.cu:294-309). Stability keeps columns ascending
        //    inside every row, which is exactly the oracle's counting transpose.
        for (long long k = 0; k < nnz; ++k)
            if (inner[k] < 0 || inner[k] >= n)
                throw std::invalid_argument("psb200_analyze_pattern_csc: inner index out of range");
        DevBuf<int> iota, sorted_rows, col_of;
        iota.alloc(nnz);
        sorted_rows.alloc(nnz);
        col_of.alloc(nnz);
        iota_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(nnz, iota.p);
        expand_cols_kernel<<<blocks_for(n * 8, 256), 256, 0, st>>>((int)n, csc_outer.p, col_of.p);
        check_launch();
        int end_bit = 1;
        while ((1ll << end_bit) < n && end_bit < 31)
            ++end_bit;
        size_t tmp_bytes = 0;
        PSB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, csc_inner.p, sorted_rows.p, iota.p, perm.p, (int)nnz, 0, end_bit, st));
        DevBuf<unsigned char> tmp;
        tmp.alloc(tmp_bytes);
        PSB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, csc_inner.p, sorted_rows.p, iota.p, perm.p, (int)nnz, 0, end_bit, st));
        row_ptr_from_sorted_kernel<<<blocks_for(n + 1, 256), 256, 0, st>>>((int)n, nnz, sorted_rows.p, A.rp.p);
        gather_int_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(nnz, col_of.p, perm.p, A.ci.p);
        check_launch();
        PSB_CUDA(cudaStreamSynchronize(st));
    }
    else
    {
        PSB_CUDA(cudaMemsetAsync(A.rp.p, 0, sizeof(int) * (n + 1), st));
    }
    pattern_block = 1;
    if (prm.block_size > 1 && n > 0)
    {
        const int B = prm.block_size;
        nnz = expand_block_pattern(ctx, B, n, A.rp, A.ci, perm);
        A.nnz = nnz;
        A.va.alloc(std::max<long long>(nnz, 1), false, 64);
        pattern_block = B;
        sym_pattern = false; // CSR arrays no longer alias the CSC arrays
    }
    A.block = pattern_block;
    A.plan(prm.spmv_kernel, st);
    PSB_CUDA(cudaStreamSynchronize(st));
    analyzed = true;
    t_analyze_ms = now_ms() - t0;
}

// ==================================================================================== factorize
void Solver::factorize(long long n_, long long nnz_, const int *outer, const int *inner, const double *vals)
{
    if (!vals && nnz_ > 0)
        throw std::invalid_argument("psb200_factorize_csc: null values");
    ensure_ctx(*this);
    // factorize() without (or with a stale) analyze_pattern(): analyze now. The Eigen iterative wrappers
    // accept this order too (EigenSolver.tpp:100-105 only needs the matrix).
    bool need = !analyzed || n_global != n_ || nnz_global != nnz_ || pattern_block != std::max(1, prm.block_size);
    // The raw CSC values go to the device first (valid whatever the pattern turns out to be): with pinned host
    // memory the copy runs while the host hashes the index arrays below.
    if (!dist)
    {
        csc_vals.alloc(std::max<long long>(nnz_, 1));
        if (nnz_)
            PSB_CUDA(cudaMemcpyAsync(csc_vals.p, vals, sizeof(double) * nnz_, cudaMemcpyHostToDevice, ctx.stream));
    }
    try
    {
        if (!need && prm.verify_pattern && outer && inner)
        {
            // guard against a silently changed pattern (Newton re-assembles every iteration, Newton.cpp:189-191)
            need = pattern_hash_of(n_, nnz_, outer, inner) != pattern_hash;
            if (dist && dist->connected && dist->world > 1)
                need = dist_max(need ? 1.0 : 0.0) != 0.0;
        }
        if (need)
        {
            if (!outer || (!inner && nnz_ > 0))
                throw std::invalid_argument("psb200_factorize_csc: pattern unknown and no index arrays given");
            analyze_pattern(n_, nnz_, outer, inner, precond_num > 0 ? precond_num : (int)n_);
        }
    }
    catch (...)
    {
        cudaStreamSynchronize(ctx.stream); // the value copy above may still be reading the caller's buffer
        throw;
    }
    const double t0 = now_ms();
    cudaStream_t st = ctx.stream;
    if (dist)
    {
        factorize_values_dist(vals);
        if (amg_global())
            factorize_full_for_amg(vals, nullptr, 0.0);
    }
    else if (nnz)
    {
        gather_vals_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(nnz, csc_vals.p, perm.p, A.va.p);
        check_launch();
    }
    factorize_tail(t0);
}

// Row partition + AMG in "global" mode (scalar problems): the hierarchy is that of the whole matrix.
bool Solver::amg_global() const
{
    return dist && prm.precond == "amg" && prm.amg.dist_mode == "global" && std::max(1, prm.block_size) == 1;
}
// Row partition + AMG with every large level partitioned (amg_dist.cu): scalar and block problems
bool Solver::amg_partitioned() const { return dist && prm.precond == "amg" && prm.amg.dist_mode == "partitioned"; }

// values of the whole matrix into the helper instance (pattern analysed in analyze_pattern)
void Solver::factorize_full_for_amg(const double *h_vals, const double *d_vals, double diag_shift)
{
    if (!full_ || !full_->analyzed)
        throw std::runtime_error("psb200 dist: the whole-matrix helper has no pattern (analyze_pattern with precond=amg first)");
    // earlier solves on our stream may still read the previous values / hierarchy
    PSB_CUDA(cudaStreamSynchronize(ctx.stream));
    amg.reset();
    AllocScope guard(full_->ctx.stream);
    if (d_vals)
        full_->factorize_device(n_global, nnz_global, d_vals, diag_shift);
    else
        full_->factorize(n_global, nnz_global, nullptr, nullptr, h_vals);
}

// Newton step with the Hessian values already on the device (reference call site Newton.cpp:173-214):
// d_vals holds the values in the CSC order of the analyzed pattern, nothing crosses PCIe. diag_shift is added to the
// diagonal after the gather (RegularizedNewton: hessian += reg_weight * I, Newton.cpp:287-290).
void Solver::factorize_device(long long n_, long long nnz_, const double *d_vals, double diag_shift)
{
    if (!analyzed)
        throw std::invalid_argument("psb200_factorize_csc_device: analyze_pattern() first (the device path has no index arrays to analyze)");
    if (n_ != n_global || nnz_ != nnz_global || pattern_block != std::max(1, prm.block_size))
        throw std::invalid_argument("psb200_factorize_csc_device: size differs from the analyzed pattern");
    if (!d_vals && nnz_ > 0)
        throw std::invalid_argument("psb200_factorize_csc_device: null values");
    ensure_ctx(*this);
    const double t0 = now_ms();
    cudaStream_t st = ctx.stream;
    if (nnz)
    {
        if (dist)
            gather_vals_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(nnz, d_vals + dist->val_lo, dist->d_perm.p, A.va.p);
        else
            gather_vals_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(nnz, d_vals, perm.p, A.va.p);
        check_launch();
    }
    if (diag_shift != 0.0 && n > 0)
    {
        int *d_missing = (int *)ctx.counter.p + 3;
        PSB_CUDA(cudaMemsetAsync(d_missing, 0, sizeof(int), st));
        shift_diag_kernel<<<blocks_for(n * 8, 256), 256, 0, st>>>(A.view(), A.va.p, diag_shift, d_missing);
        check_launch();
        int missing = 0;
        PSB_CUDA(cudaMemcpyAsync(&missing, d_missing, sizeof(int), cudaMemcpyDeviceToHost, st));
        PSB_CUDA(cudaStreamSynchronize(st));
        if (missing)
            throw std::runtime_error("psb200_factorize_csc_device: diag_shift needs a structurally present diagonal");
    }
    if (amg_global())
        factorize_full_for_amg(nullptr, d_vals, diag_shift);
    factorize_tail(t0);
}

// preconditioner setup on the values now in A.va (shared by the host and the device entry)
void Solver::factorize_tail(double t0)
{
    cudaStream_t st = ctx.stream;
    A.refresh_bsr(st); // block-3 matrices: the BSR form follows every change of the values
    dinv.alloc(n_pad, true);
    int *d_bad = (int *)ctx.counter.p + 3;
    PSB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    if (n > 0)
    {
        inv_diag_kernel<<<blocks_for(n * 8, 256), 256, 0, st>>>(A.view(), dinv.p, prm.precond == "none" ? 0 : 1, d_bad);
        check_launch();
    }
    int bad = 0;
    PSB_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaStreamSynchronize(st));
    if (bad)
        throw std::runtime_error("psb200_factorize_csc: non-finite diagonal entry");
    ensure_vectors();
    t_setup_precond_ms = 0;
    if (prm.precond == "amg")
    {
        const double t1 = now_ms();
        AmgParams ap = prm.amg;
        ap.block_size = pattern_block;
        amg.reset();
        amg_dist.reset();
        if (amg_partitioned())
        {
            // multi-GPU: every level above amg.replicate_below non-zeros is row-partitioned (decoupled aggregation, rank-local
            // P / R, distributed Galerkin product), the small levels are replicated; memory per rank ~ 1 / world
            amg_dist = std::make_unique<AmgDist>(*this, ap);
            amg_dist->setup(imposed_aggregates);
        }
        else if (dist && amg_global())
        {
            // multi-GPU, scalar problems: the hierarchy of the WHOLE matrix on every rank (setup is redundant, reductions
            // stay local), level 0 of the cycle partitioned, coarse levels replicated -- the iteration counts of 1 GPU
            amg = std::make_unique<AmgHierarchy>(ctx, ap);
            {
                LocalScope local(ctx);
                amg->setup(full_->A, imposed_aggregates);
            }
            amg->setup_dist_fine(
                A, dinv.p, dist->plan.r0(), [this](const double *v, const int *done) { push_halo_of(v, done); },
                [this](const double *partial, double *out, long long len, const int *done) { bulk_allreduce(partial, out, len, done); });
        }
        else if (dist)
        {
            // multi-GPU, amg.dist_mode = local (and block problems in global mode): every rank builds the hierarchy of its
            // own diagonal block (block-Jacobi across ranks, no communication in the cycle)
            amg = std::make_unique<AmgHierarchy>(ctx, ap);
            LocalScope local(ctx);
            build_diag_block_dist();
            amg->setup(dist->A_diag, imposed_aggregates);
        }
        else
        {
            amg = std::make_unique<AmgHierarchy>(ctx, ap);
            amg->setup(A, imposed_aggregates);
        }
        PSB_CUDA(cudaStreamSynchronize(st));
        t_setup_precond_ms = now_ms() - t1;
        if (graph_exec)
        {
            cudaGraphExecDestroy(graph_exec);
            graph_exec = nullptr;
            graph_key.clear();
        }
    }
    else
    {
        amg.reset();
        amg_dist.reset();
    }
    factorized = true;
    last_iters = 0;
    last_error = 0;
    last_status = 0;
    t_factorize_ms = t0 > 0 ? now_ms() - t0 : 0.0;
    build_info();
}

void Solver::gather_values_to_csr(const double *d_csc_vals)
{
    if (!nnz)
        return;
    gather_vals_kernel<<<blocks_for(nnz, 256), 256, 0, ctx.stream>>>(nnz, d_csc_vals, perm.p, A.va.p);
    check_launch();
}

void Solver::run_solver(const double *d_b)
{
    if (prm.krylov == "cg1r")
        run_cg1r(d_b);
    else if (dist)
        prm.precond == "amg" ? run_cg_amgcl_dist(d_b) : run_cg_eigen_dist(d_b);
    else if (prm.krylov == "bicgstab")
        run_bicgstab(d_b);
    else if (prm.precond == "amg")
        run_cg_amgcl(d_b);
    else
        run_cg_eigen(d_b);
}

void Solver::ensure_vectors()
{
    const size_t np = (size_t)n_pad;
    bool realloc = vx.capacity() < np + 16;
    for (DevBuf<double> *v : {&vb, &vx, &vr, &vp, &vq})
        v->alloc(np, false);
    if (prm.krylov == "bicgstab")
        for (DevBuf<double> *v : {&vz, &vy, &vv, &vt, &vr0})
            v->alloc(np, false);
    if (prm.precond == "amg")
        vz.alloc(np, false);
    if (realloc && graph_exec)
    {
        cudaGraphExecDestroy(graph_exec);
        graph_exec = nullptr;
        graph_key.clear();
    }
}

// ==================================================================================== solve
void Solver::solve_host(const double *b, double *x, long long n_)
{
    if (!factorized)
        throw std::runtime_error("psb200_solve: factorize() has not been called");
    if (n_ != n_global || !b || !x)
        throw std::invalid_argument("psb200_solve: size mismatch or null vector");
    const double t0 = now_ms();
    cudaStream_t st = ctx.stream;
    ensure_vectors();
    if (prm.profile)
        ctx.prof.clear();
    // row-partitioned mode: b and x are full-length on every rank; this rank reads / writes its own rows
    const long long row0 = dist ? dist->plan.r0() : 0;
    b += row0;
    x += row0;
    PSB_CUDA(cudaMemcpyAsync(vb.p, b, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    PSB_CUDA(cudaMemcpyAsync(vx.p, x, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    run_solver(vb.p);
    PSB_CUDA(cudaMemcpyAsync(x, vx.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaStreamSynchronize(st));
    t_solve_ms = now_ms() - t0;
    build_info();
}

void Solver::solve_device(const double *d_b, double *d_x, long long n_)
{
    if (!factorized)
        throw std::runtime_error("psb200_solve_device: factorize() has not been called");
    if (n_ != n || !d_b || !d_x)
        throw std::invalid_argument("psb200_solve_device: size mismatch or null vector");
    const double t0 = now_ms();
    cudaStream_t st = ctx.stream;
    ensure_vectors();
    if (prm.profile)
        ctx.prof.clear();
    // x lives in the solver's own buffer so captured graphs keep stable pointers
    PSB_CUDA(cudaMemcpyAsync(vx.p, d_x, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    run_solver(d_b);
    PSB_CUDA(cudaMemcpyAsync(d_x, vx.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    PSB_CUDA(cudaStreamSynchronize(st));
    t_solve_ms = now_ms() - t0;
    build_info();
}

// ||A x - b||_2 with x, b resident on the device (the Newton residual check H dx + g, Newton.cpp:207, with b = -g)
double Solver::residual_norm_device(const double *d_x, const double *d_b, long long n_)
{
    if (!factorized)
        throw std::runtime_error("psb200_residual_norm_device: factorize() has not been called");
    if (n_ != n || !d_x || !d_b)
        throw std::invalid_argument("psb200_residual_norm_device: size mismatch or null vector");
    ensure_vectors();
    cudaStream_t st = ctx.stream;
    // the solver's own buffers: padded (zero tails) and, on a row partition, the ones the halo push reads
    PSB_CUDA(cudaMemcpyAsync(vx.p, d_x, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    if (dist)
        push_halo_of(vx.p);
    DevBuf<double> out;
    out.alloc(4, true);
    launch_spmv(ctx, "spmv_residual", A, vx.p, EpiResidualNorms{vr.p, d_b, dinv.p}, FinStore{out.p, 3});
    double h[3] = {0, 0, 0};
    PSB_CUDA(cudaMemcpyAsync(h, out.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaStreamSynchronize(st));
    if (dist)
        check_comm_error();
    return std::sqrt(h[0]);
}

// ||A x - b||_2 with full-length host vectors (on a row partition every rank passes the full vectors and obtains the
// global norm: the reduction all-reduces inside its kernel). The Newton residual check of Newton.cpp:207.
double Solver::residual_norm_host(const double *x, const double *b, long long n_)
{
    if (!factorized)
        throw std::runtime_error("psb200_residual_norm: factorize() has not been called");
    if (n_ != n_global || !x || !b)
        throw std::invalid_argument("psb200_residual_norm: size mismatch or null vector");
    ensure_vectors();
    const long long row0 = dist ? dist->plan.r0() : 0;
    DevBuf<double> dx, db;
    dx.alloc((size_t)n_pad, true);
    db.alloc((size_t)n_pad, true);
    PSB_CUDA(cudaMemcpyAsync(dx.p, x + row0, sizeof(double) * n, cudaMemcpyHostToDevice, ctx.stream));
    PSB_CUDA(cudaMemcpyAsync(db.p, b + row0, sizeof(double) * n, cudaMemcpyHostToDevice, ctx.stream));
    return residual_norm_device(dx.p, db.p, n);
}

// Row partition: every rank contributes its own rows of x_full and receives everybody's (the Newton driver needs the whole
// step on every rank). One fused all-gather kernel over NVLink.
void Solver::dist_allgather_host(double *x_full, long long n_)
{
    if (!dist)
        return; // single GPU: x is complete already
    if (!analyzed || n_ != n_global || !x_full)
        throw std::invalid_argument("psb200_dist_allgather: analyze_pattern first / size mismatch");
    DevBuf<double> mine, full;
    mine.alloc((size_t)n_pad, true);
    full.alloc((size_t)n_global + 4, true);
    const long long row0 = dist->plan.r0();
    PSB_CUDA(cudaMemcpyAsync(mine.p, x_full + row0, sizeof(double) * n, cudaMemcpyHostToDevice, ctx.stream));
    if (dist->world > 1)
        bulk_allgather(mine.p, full.p, dist->plan.offsets.data(), nullptr);
    else
        PSB_CUDA(cudaMemcpyAsync(full.p, mine.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx.stream));
    PSB_CUDA(cudaMemcpyAsync(x_full, full.p, sizeof(double) * n_global, cudaMemcpyDeviceToHost, ctx.stream));
    PSB_CUDA(cudaStreamSynchronize(ctx.stream));
    check_comm_error();
}

// Launches batches of iterations until the device-side `done` flag is seen. Two batches are kept in
// flight so the GPU never waits for the host; kernels launched after convergence exit immediately
// (they read st->done first), so x is exactly the iterate at the converged iteration.
void Solver::drive(const std::function<void()> &enqueue_batch, int batch_iters, const std::string &key)
{
    cudaStream_t st = ctx.stream;
    const bool graph = prm.use_graph && !prm.profile;
    if (graph && (graph_exec == nullptr || graph_key != key))
    {
        if (graph_exec)
        {
            cudaGraphExecDestroy(graph_exec);
            graph_exec = nullptr;
        }
        cudaGraph_t g = nullptr;
        ctx.capturing = true;
        const long long launches_before = ctx.launches;
        PSB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
        try
        {
            enqueue_batch();
        }
        catch (...)
        {
            cudaStreamEndCapture(st, &g);
            if (g)
                cudaGraphDestroy(g);
            ctx.capturing = false;
            throw;
        }
        PSB_CUDA(cudaStreamEndCapture(st, &g));
        ctx.capturing = false;
        graph_launches_per_batch = ctx.launches - launches_before; // launch accounting for gpu_launches
        ctx.launches = launches_before;
        PSB_CUDA(cudaGraphInstantiate(&graph_exec, g, 0));
        PSB_CUDA(cudaGraphDestroy(g));
        graph_key = key;
    }
    const long long max_batches = (long long)prm.max_iter / std::max(1, batch_iters) + 3;
    bool finished = false;
    for (long long k = 0; k < max_batches && !finished; ++k)
    {
        if (graph)
        {
            PSB_CUDA(cudaGraphLaunch(graph_exec, st));
            ctx.launches += graph_launches_per_batch;
        }
        else
            enqueue_batch();
        PSB_CUDA(cudaMemcpyAsync(&h_state[k & 1], d_state, sizeof(KState), cudaMemcpyDeviceToHost, st));
        PSB_CUDA(cudaEventRecord(ev[k & 1], st));
        if (k >= 1)
        {
            PSB_CUDA(cudaEventSynchronize(ev[(k - 1) & 1]));
            if (h_state[(k - 1) & 1].done)
                finished = true;
        }
    }
}

void Solver::finish_solve()
{
    cudaStream_t st = ctx.stream;
    PSB_CUDA(cudaMemcpyAsync(&h_state[2], d_state, sizeof(KState), cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaStreamSynchronize(st));
    if (prm.profile)
        ctx.prof_collect();
    const KState &s = h_state[2];
    last_iters = s.iter;
    last_status = s.done ? s.status : ST_MAXITER;
    if (s.status == ST_ZERO_RHS)
    {
        // Eigen / AMGCL: zero right-hand side => x = 0, 0 iterations, error 0
        PSB_CUDA(cudaMemsetAsync(vx.p, 0, sizeof(double) * n_pad, st));
        last_error = 0;
        last_iters = 0;
        last_status = ST_CONVERGED;
    }
    else
        last_error = s.bn2 > 0 ? std::sqrt(s.rn2 / s.bn2) : 0.0;
}

void init_state(Solver &s, double tol, int max_iter)
{
    KState &h = s.h_state[3];
    std::memset(&h, 0, sizeof(KState));
    h.tol = tol;
    h.max_iter = max_iter;
    h.rho = 1;
    h.rho_old = 1;
    h.alpha = 1;
    h.omega = 1;
    PSB_CUDA(cudaMemcpyAsync(s.d_state, &h, sizeof(KState), cudaMemcpyHostToDevice, s.ctx.stream));
}

// Jacobi-PCG in Eigen's ordering (reference path EigenSolver.tpp:108-114 -> Eigen
// conjugate_gradient()). Per iteration: 1 fused SpMV+dot, 1 fused x/r update + ||r||^2 + r.z,
// 1 direction update = B_spmv + 88 N bytes of compulsory traffic.
void Solver::run_cg_eigen(const double *d_b)
{
    KState *S = d_state;
    const int *done = &S->done;
    init_state(*this, prm.tolerance, prm.max_iter);
    launch_spmv(ctx, "spmv_residual", A, vx.p, EpiResidualNorms{vr.p, d_b, dinv.p}, FinInitEigen{S});
    launch_vec(ctx, "cg_dir", n_pad, OpCgDirEigen<true>{vp.p, vr.p, dinv.p, S, 0.0}, FinNone{}, done);
    auto batch = [&]() {
        for (int i = 0; i < prm.check_every; ++i)
        {
            launch_spmv(ctx, "spmv_dot", A, vp.p, EpiDot{vq.p, vp.p}, FinPAp{S}, done);
            launch_vec(ctx, "cg_update", n_pad, OpCgUpdateEigen{vx.p, vr.p, vp.p, vq.p, dinv.p, S, 0.0}, FinCgUpdateEigen{S}, done);
            launch_vec(ctx, "cg_dir", n_pad, OpCgDirEigen<false>{vp.p, vr.p, dinv.p, S, 0.0}, FinCgDirEigen{S}, done);
        }
    };
    std::ostringstream key;
    key << "cg_eigen/" << A.kind << "/" << A.lpr << "/" << n << "/" << (void *)vx.p << "/" << (void *)A.va.p << "/" << prm.check_every;
    drive(batch, prm.check_every, key.str());
    finish_solve();
}

// AMG-PCG in AMGCL's ordering (SURVEY A.3 "CG"; reference path AMGCL.cpp:190-212 -> amgcl::solver::cg).
void Solver::run_cg_amgcl(const double *d_b)
{
    if (!amg)
        throw std::runtime_error("psb200_solve: AMG hierarchy missing (factorize with precond=amg first)");
    KState *S = d_state;
    const int *done = &S->done;
    init_state(*this, prm.tolerance, prm.max_iter);
    PSB_CUDA(cudaMemsetAsync(vp.p, 0, sizeof(double) * n_pad, ctx.stream));
    launch_spmv(ctx, "spmv_residual", A, vx.p, EpiResidualNorms{vr.p, d_b, dinv.p}, FinInitAmgcl{S});
    auto batch = [&]() {
        amg->apply(vr.p, vz.p, done);
        launch_vec(ctx, "dot", n_pad, OpDot{vr.p, vz.p}, FinRhoAmgcl{S}, done);
        launch_vec(ctx, "cg_dir", n_pad, OpCgDirAmgcl{vp.p, vz.p, S, 0.0}, FinNone{}, done);
        launch_spmv(ctx, "spmv_dot", A, vp.p, EpiDot{vq.p, vp.p}, FinPAp{S}, done);
        launch_vec(ctx, "cg_update", n_pad, OpCgUpdateAmgcl{vx.p, vr.p, vp.p, vq.p, S, 0.0}, FinCgUpdateAmgcl{S}, done);
    };
    std::ostringstream key;
    key << "cg_amgcl/" << n << "/" << (void *)vx.p << "/" << (void *)amg.get();
    drive(batch, 1, key.str());
    finish_solve();
}

// Jacobi-BiCGSTAB in Eigen's ordering (reference Solver.cpp:437-439).
void Solver::run_bicgstab(const double *d_b)
{
    KState *S = d_state;
    const int *done = &S->done;
    init_state(*this, prm.tolerance, prm.max_iter);
    cudaStream_t st = ctx.stream;
    PSB_CUDA(cudaMemsetAsync(vp.p, 0, sizeof(double) * n_pad, st));
    PSB_CUDA(cudaMemsetAsync(vv.p, 0, sizeof(double) * n_pad, st));
    launch_spmv(ctx, "spmv_residual", A, vx.p, EpiResidualNorms{vr.p, d_b, dinv.p}, FinInitBicg{S});
    launch_vec(ctx, "copy", n_pad, OpCopy{vr0.p, vr.p}, FinNone{}, done);
    auto batch = [&]() {
        for (int i = 0; i < prm.check_every; ++i)
        {
            // restart branch (rare): only runs when the previous trip flagged |rho| < eps^2 ||r0||^2
            launch_spmv(ctx, "spmv_restart", A, vx.p, EpiResidualRestart{vr.p, vr0.p, d_b}, FinBicgRestart{S}, done, &S->restart);
            launch_vec(ctx, "bicg_p", n_pad, OpBicgP{vp.p, vy.p, vr.p, vv.p, dinv.p, S, 0.0, 0.0}, FinNone{}, done);
            launch_spmv(ctx, "spmv_dot", A, vy.p, EpiDot{vv.p, vr0.p}, FinBicgAlpha{S}, done);
            launch_vec(ctx, "bicg_s", n_pad, OpBicgS{vr.p, vz.p, vv.p, dinv.p, S, 0.0}, FinNone{}, done);
            launch_spmv(ctx, "spmv_dot2", A, vz.p, EpiDot2{vt.p, vr.p}, FinBicgOmega{S}, done);
            launch_vec(ctx, "bicg_end", n_pad, OpBicgEnd{vx.p, vr.p, vy.p, vz.p, vt.p, vr0.p, S, 0.0, 0.0}, FinBicgEnd{S}, done);
        }
    };
    std::ostringstream key;
    key << "bicgstab/" << A.kind << "/" << A.lpr << "/" << n << "/" << (void *)vx.p << "/" << (void *)d_b << "/" << prm.check_every;
    drive(batch, prm.check_every, key.str());
    finish_solve();
}

// ==================================================================================== hooks
void Solver::spmv_host(const double *x, double *y, long long n_)
{
    if (!factorized || n_ != n)
        throw std::invalid_argument("psb200_spmv: not factorized or size mismatch");
    ensure_vectors();
    cudaStream_t st = ctx.stream;
    PSB_CUDA(cudaMemcpyAsync(vp.p, x, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    launch_spmv(ctx, "spmv", A, vp.p, EpiStore{vq.p}, FinNone{});
    PSB_CUDA(cudaMemcpyAsync(y, vq.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaStreamSynchronize(st));
}

double Solver::bench_spmv(const std::string &kernel, int reps)
{
    if (!factorized)
        throw std::runtime_error("psb200_bench_spmv: factorize() first");
    ensure_vectors();
    cudaStream_t st = ctx.stream;
    const int kind0 = A.kind, lpr0 = A.lpr;
    const int narrow0 = A.narrow;
    const bool bsr0 = A.use_bsr;
    // tile-shape exploration of the stream schedule: "stream:<threads>:<cap>:<stages>[:<ctas_per_sm>]"
    std::function<void()> one = [&]() { launch_spmv(ctx, "spmv", A, vp.p, EpiStore{vq.p}, FinNone{}); };
    if (kernel.rfind("stream:", 0) == 0)
    {
        int t = 0, cap = 0, stg = 0, ctas = 0;
        if (std::sscanf(kernel.c_str(), "stream:%d:%d:%d:%d", &t, &cap, &stg, &ctas) < 3)
            throw std::invalid_argument("psb200_bench_spmv: expected stream:<threads>:<cap>:<stages>[:<ctas>]");
        bool found = false;
#define PSB_STREAM_VARIANT(T, C, S)                                                                                          \
    if (t == T && cap == C && stg == S)                                                                                      \
    {                                                                                                                        \
        found = true;                                                                                                        \
        one = [&, ctas]() { launch_spmv_stream<EpiStore, FinNone, StreamCfg<T, C, S>>(ctx, A, vp.p, EpiStore{vq.p}, FinNone{}, nullptr, nullptr, ctas); }; \
    }
        PSB_STREAM_VARIANT(256, 2560, 3)
        PSB_STREAM_VARIANT(256, 2048, 2)
        PSB_STREAM_VARIANT(256, 2048, 3)
        PSB_STREAM_VARIANT(256, 2048, 4)
        PSB_STREAM_VARIANT(128, 1024, 2)
        PSB_STREAM_VARIANT(128, 1024, 3)
        PSB_STREAM_VARIANT(128, 1024, 4)
        PSB_STREAM_VARIANT(512, 4096, 2)
        PSB_STREAM_VARIANT(512, 4096, 3)
        PSB_STREAM_VARIANT(64, 512, 4)
#undef PSB_STREAM_VARIANT
        if (!found)
            throw std::invalid_argument("psb200_bench_spmv: stream variant not compiled: " + kernel);
    }
    else if (kernel.rfind("bsr:", 0) == 0)
    {
        // tile-shape exploration of the BSR-3 schedule: "bsr:<threads>:<capb>:<stages>"
        int t = 0, cap = 0, stg = 0;
        if (std::sscanf(kernel.c_str(), "bsr:%d:%d:%d", &t, &cap, &stg) < 3)
            throw std::invalid_argument("psb200_bench_spmv: expected bsr:<threads>:<capb>:<stages>");
        if (!(A.use_bsr && A.bsr_ready))
            throw std::invalid_argument("psb200_bench_spmv: the matrix has no BSR-3 form");
        bool found = false;
#define PSB_BSR_VARIANT(T, C, S)                                                                                              \
    if (t == T && cap == C && stg == S)                                                                                       \
    {                                                                                                                         \
        found = true;                                                                                                         \
        one = [&]() { launch_spmv_bsr3<EpiStore, FinNone, BsrCfg<T, C, S, 8>>(ctx, A, vp.p, EpiStore{vq.p}, FinNone{}, nullptr, nullptr); }; \
    }
        PSB_BSR_VARIANT(256, 544, 2)
        PSB_BSR_VARIANT(256, 544, 1)
        PSB_BSR_VARIANT(128, 272, 2)
        PSB_BSR_VARIANT(128, 272, 3)
        PSB_BSR_VARIANT(128, 288, 4)
        PSB_BSR_VARIANT(512, 1088, 1)
#undef PSB_BSR_VARIANT
        if (!found)
            throw std::invalid_argument("psb200_bench_spmv: BSR variant not compiled: " + kernel);
    }
    else if (!kernel.empty())
    {
        A.plan(kernel, st);
        A.refresh_bsr(st);
    }
    // a non-trivial resident x
    launch_vec(ctx, "copy", n_pad, OpCopy{vp.p, dinv.p}, FinNone{});
    for (int i = 0; i < 3; ++i)
        one();
    check_launch();
    cudaEvent_t a, b;
    PSB_CUDA(cudaEventCreate(&a));
    PSB_CUDA(cudaEventCreate(&b));
    PSB_CUDA(cudaEventRecord(a, st));
    for (int i = 0; i < reps; ++i)
        one();
    PSB_CUDA(cudaEventRecord(b, st));
    PSB_CUDA(cudaEventSynchronize(b));
    float ms = 0;
    PSB_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    A.kind = kind0;
    A.lpr = lpr0;
    A.narrow = narrow0;
    if (A.use_bsr != bsr0)
    {
        A.use_bsr = bsr0;
        A.refresh_bsr(st);
    }
    return (double)ms / std::max(1, reps);
}

void Solver::precond_apply_host(const double *r, double *z, long long n_)
{
    if (!factorized || n_ != n)
        throw std::invalid_argument("psb200_precond_apply: not factorized or size mismatch");
    ensure_vectors();
    cudaStream_t st = ctx.stream;
    PSB_CUDA(cudaMemcpyAsync(vr.p, r, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    if (amg)
    {
        vz.alloc((size_t)n_pad, false);
        amg->apply(vr.p, vz.p, nullptr);
        PSB_CUDA(cudaMemcpyAsync(z, vz.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    }
    else
    {
        init_state(*this, 0, 0);
        launch_vec(ctx, "cg_dir", n_pad, OpCgDirEigen<true>{vp.p, vr.p, dinv.p, d_state, 0.0}, FinNone{});
        PSB_CUDA(cudaMemcpyAsync(z, vp.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    }
    PSB_CUDA(cudaStreamSynchronize(st));
}

void Solver::build_info()
{
    static const char *status_str[] = {"Running", "Converged", "Reach max iterations", "Breakdown (non-finite residual)", "Zero right-hand side",
                                       "Communication timeout"};
    std::ostringstream o;
    o << "{";
    // both conventions: Eigen/MAS (EigenSolver.tpp:88-89, MASSolver.cu:214-219) and AMGCL/Hypre (AMGCL.cpp:142-143)
    o << "\"solver_iter\":" << last_iters << ",\"solver_error\":" << jnum(last_error);
    o << ",\"num_iterations\":" << last_iters << ",\"final_res_norm\":" << jnum(last_error);
    o << ",\"solver_status\":" << jstr(status_str[std::min(std::max(last_status, 0), 5)]);
    o << ",\"krylov\":" << jstr(prm.krylov) << ",\"precond\":" << jstr(prm.precond);
    o << ",\"n\":" << n_global << ",\"nnz\":" << nnz_global;
    if (dist)
        o << ",\"dist\":{\"rank\":" << dist->rank << ",\"world\":" << dist->world << ",\"row_begin\":" << dist->plan.r0()
          << ",\"row_end\":" << dist->plan.r1() << ",\"local_nnz\":" << nnz << ",\"halo_in\":" << dist->plan.halo_cols.size()
          << ",\"halo_out\":" << dist->fine.send_rows.size() << ",\"nbr_mask\":" << dist->nbr_mask << "}";
    o << ",\"symmetric_pattern\":" << (sym_pattern ? "true" : "false");
    o << ",\"analyze_skipped\":" << (analyze_skipped ? "true" : "false");
    o << ",\"spmv_kernel\":" << jstr(A.kernel_name());
    o << ",\"cg_kernel\":\"split\"";
    o << ",\"time_analyze_ms\":" << jnum(t_analyze_ms) << ",\"time_factorize_ms\":" << jnum(t_factorize_ms);
    o << ",\"time_precond_setup_ms\":" << jnum(t_setup_precond_ms) << ",\"time_solve_ms\":" << jnum(t_solve_ms);
    o << ",\"gpu_launches\":" << ctx.launches;
    if (amg)
        o << ",\"amg\":" << amg->info_json();
    if (amg_dist)
        o << ",\"amg\":" << amg_dist->info_json();
    if ((amg || amg_dist) && dist)
        o << ",\"amg_dist_mode\":" << jstr(amg_dist ? "partitioned" : amg->has_dist_fine() ? "global" : "local");
    if (!ctx.prof.empty())
    {
        o << ",\"profile\":{";
        bool first = true;
        for (auto &kv : ctx.prof)
        {
            if (!first)
                o << ",";
            first = false;
            o << jstr(kv.first) << ":{\"ms\":" << jnum(kv.second.ms) << ",\"launches\":" << kv.second.launches << "}";
        }
        o << "}";
    }
    o << "}";
    info_json = o.str();
}

} // namespace psb
