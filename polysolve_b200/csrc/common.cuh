// Shared device/host helpers for the psb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>

namespace psb {

struct CudaError : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

// a cross-GPU wait timed out (PSB200_ERR_COMM)
struct CommError : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

#define PSB_CUDA(expr)                                                                              \
    do                                                                                              \
    {                                                                                               \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw ::psb::CudaError(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + \
                                   __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")");    \
    } while (0)

constexpr int kSMs = 148; // B200: 2 dies x 74 SMs
constexpr int kMaxRed = 4; // max simultaneous reduction values per kernel

// Stream the calling thread's device allocations are ordered on (set by the C-ABI entry points for the duration
// of a call: every buffer of a solver is produced and consumed on that solver's stream). nullptr: plain cudaMalloc.
inline cudaStream_t &alloc_stream()
{
    static thread_local cudaStream_t s = nullptr;
    return s;
}
struct AllocScope
{
    cudaStream_t prev;
    explicit AllocScope(cudaStream_t s) : prev(alloc_stream())
    {
        if (s)
            alloc_stream() = s;
    }
    ~AllocScope() { alloc_stream() = prev; }
};

// Makes `device` current for the duration of a C-ABI call and restores the caller's device afterwards: solver
// instances on different GPUs may be driven from one host thread (the reference keeps several solver instances alive
// at once, Newton.cpp:32-52), and streams / pool allocations are only valid with their own device current.
struct DeviceScope
{
    int prev = -1;
    bool active = false;
    explicit DeviceScope(int device, bool enable)
    {
        if (!enable)
            return;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device)
            active = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceScope()
    {
        if (active)
            cudaSetDevice(prev);
    }
};

// Owning device buffer. Memory comes from the device's stream-ordered pool (cudaMallocAsync on the solver's stream;
// the pool's release threshold is raised in ensure_ctx so freed blocks stay cached): a factorize that rebuilds an AMG
// hierarchy -- dozens of buffers, GBs of temporaries -- never pays cudaMalloc / cudaFree after the first time
// (reference precedent: one stream + one pool per solver, MASSolver.cu:154-156,193-195). Outside an AllocScope, or
// while the stream is being captured, it falls back to cudaMalloc. Over-allocates `pad` elements so 16-byte TMA
// bulk copies that round a range outward never leave the allocation.
template <typename T>
struct DevBuf
{
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept { *this = std::move(o); }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        if (this != &o)
        {
            release();
            p = o.p;
            n = o.n;
            cap_ = o.cap_;
            st_ = o.st_;
            pooled_ = o.pooled_;
            o.p = nullptr;
            o.n = 0;
            o.cap_ = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release()
    {
        if (p)
        {
            if (pooled_)
                cudaFreeAsync(p, st_); // ordered after every use: all of them were enqueued on st_
            else
                cudaFree(p);
        }
        p = nullptr;
        n = 0;
        cap_ = 0;
    }
    // Grow-only (re)allocation; contents are zeroed when `zero` (always after a fresh allocation).
    void alloc(size_t count, bool zero = false, size_t pad = 16)
    {
        cudaStream_t st = alloc_stream();
        if (count + pad > cap_)
        {
            release();
            const size_t bytes = (count + pad) * sizeof(T);
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            if (st)
                cudaStreamIsCapturing(st, &cs);
            pooled_ = st != nullptr && cs == cudaStreamCaptureStatusNone;
            cudaError_t e = pooled_ ? cudaMallocAsync(&p, bytes, st) : cudaMalloc(&p, bytes);
            if (e != cudaSuccess)
            {
                size_t fr = 0, tot = 0;
                cudaMemGetInfo(&fr, &tot);
                p = nullptr;
                cudaGetLastError();
                throw CudaError("CUDA out of memory: requested " + std::to_string(bytes >> 20) + " MiB, free " + std::to_string(fr >> 20) +
                                " MiB of " + std::to_string(tot >> 20) + " MiB");
            }
            st_ = st;
            cap_ = count + pad;
            zero = true;
        }
        n = count;
        if (zero)
        {
            if (pooled_)
                PSB_CUDA(cudaMemsetAsync(p, 0, cap_ * sizeof(T), st_));
            else
                PSB_CUDA(cudaMemset(p, 0, cap_ * sizeof(T)));
        }
    }
    size_t capacity() const { return cap_; }

private:
    size_t cap_ = 0;
    cudaStream_t st_ = nullptr;
    bool pooled_ = false;
};

// ------------------------------------------------------------------------------------------------
// Deterministic grid-wide reduction: warp shuffle -> shared -> one partial per CTA -> the CTA that
// takes the last ticket sums the partials in a fixed order. No floating-point atomics, so results
// are bit-reproducible for a fixed grid (contrast reference mas_utils/InnerProduct.cu:16-33, which
// does one fp64 atomic per block).
// ------------------------------------------------------------------------------------------------
// Multi-GPU communication over NVLink peer memory (one process per GPU; buffers exchanged as CUDA
// IPC handles). Every rank owns one "comm buffer" with the same layout; kernels store straight into
// the peers' buffers (st.global on mapped peer pointers) and poll their own.
//   [0, 1024)     reduction slots  RedSlot[2 parity][8 source ranks]   (flag-in-data words, no fences)
//   [1024, 3072)  wide slots       WideSlot[2 parity][8 source ranks]  (8 doubles per rank: host-level all-gather, setup only)
//   [4096, 8192)  halo flags       unsigned long long[8 source ranks]  (monotonic count of landed push chunks)
//                 bulk flags       the same for the bulk all-reduce / all-gather, 1024 bytes further
//   [8192, ...)   halo data        double[3 buffers][8 source ranks][halo_cap]  (buffer = push number mod 3)
//                 bulk data        double[2 parity][8 source ranks][halo_cap]   (vector all-reduce / all-gather of the
//                                  partitioned AMG cycle; during setup the same area is the staging arena of the
//                                  host-level exchanges, one slot of 2 * halo_cap doubles per source rank)
constexpr int kMaxRanks = 8;
// One fp64 travels as two 8-byte words {32 payload bits, 32-bit sequence tag}: an aligned 8-byte store is
// single-copy atomic over NVLink, so a word whose tag matches carries valid payload -- the reader needs no
// fence and the writer no flag store behind a fence (the "LL" idea of NCCL's low-latency protocol).
struct RedSlot
{
    uint4 w[kMaxRed]; // {lo, tag, hi, tag}
};
static_assert(sizeof(RedSlot) == 64, "RedSlot must be 64 bytes");
constexpr int kWide = 8;
struct WideSlot
{
    uint4 w[kWide];
};
constexpr size_t kCommWideOff = 1024, kCommFlagsOff = 4096, kCommBulkFlagsOff = 4096 + 1024, kCommHaloOff = 8192;
constexpr int kPushChunk = 512; // halo entries per push chunk (one release-add on the consumer's flag per chunk)
// Halo buffers per source. Two would do when every push is its own kernel (a rank pushes epoch e + 2 only after its
// consumer of e + 1 has waited for the neighbours' e + 1, which they issued after reading e). The smoother of the
// partitioned AMG levels pushes from INSIDE the multiplying kernel (EpiChebPush, amg_dist.cu): a neighbour may then
// start pushing epoch e + 2 while this rank's kernel that reads epoch e is still running -- with three buffers the
// writer of e + 3 is the first to reuse the buffer of e, and it cannot start before this rank's kernel e + 2 has pushed.
constexpr int kHaloBufs = 3;

// Flow control of the halo exchange. Every rank executes the same sequence of pushes (one per multiplied vector, at
// every level of the AMG cycle); push number e writes the parity-(e & 1) halo regions of its consumers in chunks and
// adds 1 to the consumer's flag per chunk. A consumer does not count epochs times a fixed chunk number (levels differ
// in their halo sizes, and a new pattern changes them): its own push kernel of the same epoch adds the number of chunks
// it is about to receive from every source to halo_expect[], and waiters compare the flag with that running total.
// Neighbour relations are symmetric (nbr_mask: union over all levels of "sends to" and "receives from"), every push
// delivers at least one (possibly empty) chunk to every neighbour, and a push first waits until the previous epoch has
// landed completely -- so a rank can never overwrite a parity buffer a neighbour is still reading.
struct CommDev
{
    int world = 1, rank = 0;
    unsigned nbr_mask = 0;            // ranks this one exchanges halo values with (symmetric)
    long long halo_cap = 0;           // doubles per (parity, source rank) region
    long long spin_limit = 6000000000ll; // SM clocks a wait may take before the solve fails (Params::comm_timeout_s)
    unsigned char *peer[kMaxRanks] = {}; // comm buffer base of every rank (peer[rank] is local)
    unsigned long long *red_seq = nullptr;     // local: number of all-reduces completed
    unsigned long long *push_epoch = nullptr;  // local: number of halo pushes completed
    unsigned long long *halo_expect = nullptr; // local [8]: push chunks expected so far from every source (all pushes up to push_epoch)
    unsigned long long *bulk_epoch = nullptr;  // local: number of bulk all-reduce segments completed
    unsigned long long *bulk_expect = nullptr; // local [8]: chunks expected so far from every source (bulk exchanges)
    unsigned long long *wide_seq = nullptr;    // local: number of wide all-gathers completed
    int *error = nullptr;                      // local: set to 1 on a spin-wait timeout; stays set until psb200_dist_reset
    __host__ __device__ RedSlot *slot(int owner, int parity, int src) const
    {
        return reinterpret_cast<RedSlot *>(peer[owner]) + parity * kMaxRanks + src;
    }
    __host__ __device__ WideSlot *wide(int owner, int parity, int src) const
    {
        return reinterpret_cast<WideSlot *>(peer[owner] + kCommWideOff) + parity * kMaxRanks + src;
    }
    __host__ __device__ unsigned long long *halo_flag(int owner, int src) const
    {
        return reinterpret_cast<unsigned long long *>(peer[owner] + kCommFlagsOff) + src;
    }
    __host__ __device__ double *halo(int owner, int buf, int src) const
    {
        return reinterpret_cast<double *>(peer[owner] + kCommHaloOff) + ((long long)buf * kMaxRanks + src) * halo_cap;
    }
    __host__ __device__ unsigned long long *bulk_flag(int owner, int src) const
    {
        return reinterpret_cast<unsigned long long *>(peer[owner] + kCommBulkFlagsOff) + src;
    }
    __host__ __device__ double *bulk(int owner, int parity, int src) const
    {
        return reinterpret_cast<double *>(peer[owner] + kCommHaloOff) + ((long long)(kHaloBufs + parity) * kMaxRanks + src) * halo_cap;
    }
    // setup-time staging arena: the bulk area seen as one slot of 2 * halo_cap doubles per source rank
    __host__ __device__ unsigned char *arena(int owner, int src) const
    {
        return reinterpret_cast<unsigned char *>(reinterpret_cast<double *>(peer[owner] + kCommHaloOff) + (long long)(kHaloBufs * kMaxRanks + 2 * src) * halo_cap);
    }
    __host__ __device__ size_t arena_slot_bytes() const { return (size_t)(2 * halo_cap) * sizeof(double); }
};


__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
// flag += v on a (peer) counter with release semantics at system scope: every store this thread has observed
// (its own and, through the preceding bar.sync, its CTA's) is visible before the increment
__device__ __forceinline__ void red_release_sys_add(unsigned long long *p, unsigned long long v)
{
    asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_ll(uint4 *p, double v, unsigned tag)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)b), "r"(tag), "r"((unsigned)(b >> 32)), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_ll(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// spin until *p >= want (system scope); false on timeout. Once a wait has timed out on this rank (*err set) every later
// wait gives up at once: a lost peer costs one spin limit, not one per kernel of a captured graph.
__device__ __forceinline__ bool spin_ge(const unsigned long long *p, unsigned long long want, const int *err, long long limit)
{
    if (err && *(const volatile int *)err)
        return false;
    const long long t0 = clock64();
    while (ld_sys(p) < want)
    {
        if (clock64() - t0 > limit)
            return false;
        __nanosleep(20);
    }
    return true;
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-serialization attribute may become
// resident while its predecessor drains; everything it does before griddep_wait() must be independent of the
// predecessor (barrier init, TMA prefetch of the constant matrix), everything after sees the predecessor's memory.
// Both are no-ops for an ordinary launch.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct RedCtx
{
    double *partials;      // [kMaxRed][max_blocks]
    unsigned int *counter; // ticket, self-resetting (atomicInc wraps)
    int stride;            // max_blocks
    CommDev comm;          // world == 1: single GPU, no exchange
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sums of NV values. Result valid in thread 0. THREADS must be a multiple of 32, <= 1024.
template <int NV, int THREADS>
__device__ __forceinline__ void block_sum(double (&v)[NV], double (*sm)[THREADS / 32])
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i)
    {
        const double s = warp_sum(v[i]);
        if (lane == 0)
            sm[i][w] = s;
    }
    __syncthreads();
    if (w == 0)
    {
#pragma unroll
        for (int i = 0; i < NV; ++i)
        {
            double s = (lane < THREADS / 32) ? sm[i][lane] : 0.0;
            s = warp_sum(s);
            v[i] = s;
        }
    }
    __syncthreads();
}

// All-reduce (sum) of NV doubles across the ranks, fused into the reducing kernel: the last CTA of
// every rank stores its totals as tagged words into slot[parity][rank] of EVERY peer's comm buffer over
// NVLink, then polls the `world` slots of its own buffer until their tags carry this reduction's sequence
// number and sums them in rank order (so every rank obtains the bit-identical result). One NVLink store
// latency per all-reduce; no fence on either side. Two slot parities are enough: a rank cannot start
// reduction s+2 before every peer has finished reading reduction s.
// Called by all threads of the reducing CTA (>= 32 threads); tot[] is valid in thread 0 on entry and on
// return. seq = index of this reduction (1-based, identical on every rank).
template <int NV>
__device__ __forceinline__ void comm_allreduce_seq(const CommDev &c, double (&tot)[NV], unsigned long long seq)
{
    __shared__ double sh[kMaxRed];
    __shared__ double got[kMaxRed][kMaxRanks];
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int i = 0; i < NV; ++i)
            sh[i] = tot[i];
    }
    __syncthreads();
    const int par = (int)(seq & 1);
    const unsigned tag = (unsigned)seq;
    if ((int)threadIdx.x < c.world)
    {
        RedSlot *dst = c.slot(threadIdx.x, par, c.rank);
#pragma unroll
        for (int i = 0; i < NV; ++i)
            st_ll(&dst->w[i], sh[i], tag);
        const RedSlot *src = c.slot(c.rank, par, threadIdx.x);
        const long long t0 = clock64();
        const bool failed_before = *(const volatile int *)c.error != 0;
#pragma unroll
        for (int i = 0; i < NV; ++i)
        {
            uint4 v = ld_ll(&src->w[i]);
            while (v.y != tag || v.w != tag)
            {
                if (failed_before || clock64() - t0 > c.spin_limit)
                {
                    *c.error = 1;
                    break;
                }
                v = ld_ll(&src->w[i]);
            }
            got[i][threadIdx.x] = __longlong_as_double((long long)(((unsigned long long)v.z << 32) | v.x));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int i = 0; i < NV; ++i)
        {
            double s = 0;
            for (int q = 0; q < c.world; ++q)
                s += got[i][q];
            tot[i] = s;
        }
    }
}
template <int NV>
__device__ __forceinline__ void comm_allreduce(const CommDev &c, double (&tot)[NV])
{
    __shared__ unsigned long long sseq;
    if (threadIdx.x == 0)
        sseq = *c.red_seq + 1;
    __syncthreads();
    comm_allreduce_seq<NV>(c, tot, sseq);
    if (threadIdx.x == 0)
        *c.red_seq = sseq;
}

// Returns true (for all threads of the CTA) in the CTA that finished last; then tot[] (thread 0) holds
// the grid totals. With NV == 0 this is only the "who is last" ticket.
template <int NV, int THREADS>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV > 0 ? NV : 1], const RedCtx &rc, double (&tot)[NV > 0 ? NV : 1])
{
    __shared__ double sm[NV > 0 ? NV : 1][THREADS / 32];
    __shared__ int is_last;
    if constexpr (NV > 0)
    {
        block_sum<NV, THREADS>(v, sm);
        if (threadIdx.x == 0)
        {
#pragma unroll
            for (int i = 0; i < NV; ++i)
                rc.partials[i * rc.stride + blockIdx.x] = v[i];
        }
    }
    else
        __syncthreads(); // all reads of this CTA are done before the ticket is taken
    if (threadIdx.x == 0)
    {
        __threadfence();
        const unsigned t = atomicInc(rc.counter, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last)
        return false;
    if constexpr (NV > 0)
    {
        __threadfence();
        double s[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i)
        {
            s[i] = 0;
            for (int b = threadIdx.x; b < (int)gridDim.x; b += THREADS)
                s[i] += __ldcg(rc.partials + i * rc.stride + b);
        }
        block_sum<NV, THREADS>(s, sm);
#pragma unroll
        for (int i = 0; i < NV; ++i)
            tot[i] = s[i];
        if (rc.comm.world > 1)
            comm_allreduce<NV>(rc.comm, tot);
    }
    return true;
}

} // namespace psb
