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

// Owning device buffer (cudaMalloc). Over-allocates `pad` elements so 16-byte TMA bulk copies that
// round a range outward never leave the allocation.
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
            o.p = nullptr;
            o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        n = 0;
    }
    // Grow-only (re)allocation; contents are zeroed when `zero`.
    void alloc(size_t count, bool zero = false, size_t pad = 16)
    {
        if (count + pad > cap_)
        {
            release();
            cudaError_t e = cudaMalloc(&p, (count + pad) * sizeof(T));
            if (e != cudaSuccess)
            {
                size_t fr = 0, tot = 0;
                cudaMemGetInfo(&fr, &tot);
                p = nullptr;
                throw CudaError("CUDA out of memory: requested " + std::to_string((count + pad) * sizeof(T) >> 20) +
                                " MiB, free " + std::to_string(fr >> 20) + " MiB of " + std::to_string(tot >> 20) + " MiB");
            }
            cap_ = count + pad;
            zero = true;
        }
        n = count;
        if (zero)
            PSB_CUDA(cudaMemset(p, 0, cap_ * sizeof(T)));
    }
    size_t capacity() const { return cap_; }

private:
    size_t cap_ = 0;
};

// ------------------------------------------------------------------------------------------------
// Deterministic grid-wide reduction: warp shuffle -> shared -> one partial per CTA -> the CTA that
// takes the last ticket sums the partials in a fixed order. No floating-point atomics, so results
// are bit-reproducible for a fixed grid (contrast reference mas_utils/InnerProduct.cu:16-33, which
// does one fp64 atomic per block).
struct RedCtx
{
    double *partials;      // [kMaxRed][max_blocks]
    unsigned int *counter; // ticket, self-resetting (atomicInc wraps)
    int stride;            // max_blocks
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
    }
    return true;
}

} // namespace psb
