// Fused halo push: the Chebyshev step of a row-partitioned AMG level stores the boundary entries of the new iterate
// into the consumers' halo buffers from the EPILOGUE of the multiplying kernel (no separate push launch), and with the
// boundary-first tile order the values are on the wire while the kernel is still working on its interior rows.
// See CommDev (common.cuh) for the flow control and kHaloBufs for why three halo buffers make this safe.
#pragma once
#include "dist.hpp"

namespace psb {

struct FusedPush
{
    unsigned long long *push_epoch, *halo_expect;
    int world;
    PushMap pm;
};

// One Chebyshev step (EpiCheb, spmv.cuh) + push of the rows other ranks read
struct EpiChebPush
{
    static constexpr int NV = 0;
    using Pre = Pre4;
    const double *b, *dinv, *xin;
    double *p, *xout;
    double alpha, beta;
    const unsigned long long *push_epoch;
    PushMap pm;
    __device__ __forceinline__ Pre pre(int row) const
    {
        return {__ldg(b + row), __ldg(dinv + row), __ldg(xin + row), beta != 0.0 ? p[row] : 0.0};
    }
    __device__ __forceinline__ void operator()(int row, double s, Pre q, double (&)[1]) const
    {
        const double res = q.b * (q.a - s);
        const double pn = alpha * res + beta * q.d;
        p[row] = pn;
        const double xn = q.c + pn;
        xout[row] = xn;
        if ((__ldg(pm.bits + (row >> 5)) >> (row & 31)) & 1u)
        {
            // the epoch and the launch counter are updated by the last CTA of this kernel, after every epilogue has run
            const unsigned long long epoch = __ldcg(push_epoch), seq = __ldcg(pm.fused_seq);
            const long long shift = (long long)((epoch + 1) % kHaloBufs) * pm.buf_stride;
            // index of this row among the sent rows: rank of its bit (prefix count of the word + bits below it)
            const unsigned word = __ldg(pm.bits + (row >> 5));
            const int lo = __ldg(pm.bits_prefix + (row >> 5)) + __popc(word & ((1u << (row & 31)) - 1u));
            const int s1 = __ldg(pm.bptr + lo + 1);
            for (int sl = __ldg(pm.bptr + lo); sl < s1; ++sl)
            {
                const int ch = __ldg(pm.slot_chunk + sl);
                double *dst = pm.slot_dst[sl] + shift;
                *dst = xn;
                __threadfence(); // the value is ordered before the count
                const unsigned long long old = atomicAdd(pm.chunk_done + ch, 1ull);
                if (old + 1 == (unsigned long long)__ldg(pm.chunk_cnt + ch) * (seq + 1))
                {
                    // this store completed the chunk: everything counted before (by any thread of this GPU) is released
                    // to the consumer with the flag
                    fence_acq_rel_sys();
                    red_release_sys_add(pm.chunk_flag[ch], 1ull);
                }
            }
        }
    }
};

// Run by thread 0 of the last CTA: the bookkeeping of a completed push (running totals of expected chunks, epoch), the
// release of the empty chunks (neighbours that receive nothing from this level still get their one chunk per push)
struct FinPushDone
{
    unsigned long long *push_epoch, *halo_expect;
    int world;
    PushMap pm;
    __device__ __forceinline__ void operator()(const double *) const
    {
        int inc[kMaxRanks];
#pragma unroll
        for (int q = 0; q < kMaxRanks; ++q)
            inc[q] = pm.in_chunks[q]; // independent loads, issued together
#pragma unroll
        for (int q = 0; q < kMaxRanks; ++q)
            if (inc[q])
                halo_expect[q] += (unsigned long long)inc[q];
        for (int e = 0; e < pm.n_empty; ++e) // usually none: a neighbour of another level that receives nothing from this one
            red_release_sys_add(pm.empty_flag[e], 1ull);
        *pm.fused_seq += 1;
        *push_epoch += 1;
    }
};

} // namespace psb
