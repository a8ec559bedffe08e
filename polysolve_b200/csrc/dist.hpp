// Row-partitioned multi-GPU state (one process per GPU). See dist.cu.
#pragma once
#include "launch.cuh"

#include <vector>

namespace psb {

// Host-side plan for one rank: contiguous row range balanced by nnz, local CSR with columns remapped
// to [local | halo regions], send lists. Pure host code (testable without a GPU).
struct DistPlanHost
{
    int rank = 0, world = 1;
    long long n_global = 0, nnz_global = 0, halo_cap = 0;
    std::vector<long long> offsets; // world + 1 row offsets; rank g owns [offsets[g], offsets[g+1])
    std::vector<int> rp, ci, perm;  // local CSR: ci remapped (c - r0 for owned columns, nl + q*halo_cap + pos for halo columns); vals_local[k] = vals_csc[perm[k]]
    std::vector<int> send_begin;    // world + 1: send_rows[send_begin[q] .. send_begin[q+1]) go to rank q, in ascending global order
    std::vector<int> send_rows;     // local row ids
    std::vector<int> recv_count;    // world: number of halo values received from each rank
    std::vector<int> halo_cols;     // global ids of the halo columns, ascending (grouped by owner automatically)
    long long r0() const { return offsets[rank]; }
    long long r1() const { return offsets[rank + 1]; }
    // align: row offsets are rounded up to a multiple of `align` (block problems keep the rows of a node together), and
    // halo / send lists are completed to whole nodes (all `align` dofs of a node travel together), so that the local
    // matrix can be expanded to full align x align blocks including its halo columns
    void build(long long n, long long nnz, const int *outer, const int *inner, int rank_, int world_, long long halo_cap_, int align = 1);
};

// Device view of the send side of one halo exchange (see push_section, dist.cu)
struct PushList
{
    const int *rows;                                             // local row ids, grouped by destination
    const int *chunk_peer, *chunk_start, *chunk_cnt, *chunk_off; // per chunk: destination, first entry in rows, entries, offset in the destination's region
    int nchunks;
    int in_chunks[kMaxRanks]; // chunks this rank receives from every source in the same push (0: not a neighbour)
};

// Device view of the send side for pushes issued from the EPILOGUE of the multiplying kernel (fused push): which local
// rows travel where. bits: one bit per local row; brow: the sent rows, ascending; bptr: CSR over brow into the slot
// arrays: slot_dst = address of the entry in the destination's halo buffer 0 (the other buffers follow at buf_stride
// doubles), slot_chunk = chunk id. chunk_done counts the entries stored per chunk, monotonically over the launches
// (fused_seq = launches completed): the thread whose increment completes a chunk releases the consumer's flag
// (chunk_flag). Everything is reached through pointers with constant offsets: a kernel parameter struct that is indexed
// dynamically (or whose address escapes) is copied to local memory by every thread.
struct PushMap
{
    const unsigned *bits;
    const int *bits_prefix; // number of set bits in the words before this one
    const int *brow, *bptr, *slot_chunk;
    double *const *slot_dst;
    const int *chunk_cnt;
    unsigned long long *const *chunk_flag;
    unsigned long long *const *empty_flag; // flags of the chunks without entries (released by the finalizer), n_empty of them
    unsigned long long *chunk_done, *fused_seq;
    const int *in_chunks; // device copy of HaloPlan::in_chunks
    long long buf_stride;
    int n_brow, nchunks, n_empty;
};

// Halo exchange plan of one row-partitioned matrix (the fine matrix, or one level of the partitioned AMG hierarchy):
// host lists + the device push list. finalize() cuts the send lists into chunks for a given neighbour mask: every
// neighbour gets at least one (possibly empty) chunk per push, see CommDev.
struct HaloPlan
{
    std::vector<int> send_begin, send_rows, recv_count;
    DevBuf<int> push_rows, chunk_tab;
    int n_push = 0, n_chunks = 0, world = 1;
    int in_chunks[kMaxRanks] = {};
    unsigned mask() const; // ranks this plan sends to or receives from
    void finalize(unsigned nbr_mask, const CommDev &comm, cudaStream_t st);
    PushList push() const;
    // fused push (see PushMap); built by finalize()
    DevBuf<unsigned> send_bits;
    DevBuf<int> brow, bptr, slot_chunk, in_chunks_dev, bits_prefix;
    DevBuf<unsigned long long> slot_dst, chunk_flag, empty_flag; // device addresses stored as 64-bit integers
    int n_empty = 0;
    DevBuf<unsigned long long> chunk_done; // [n_chunks] + fused_seq at the end
    long long buf_stride = 0;
    int n_brow = 0, n_slots = 0;
    long long n_local = 0; // local rows of the matrix this plan belongs to (bitmap length)
    PushMap push_map() const;
    // CTAs that push: one per chunk up to 128 (a CTA loops over chunks beyond that); at least one on a multi-rank run so
    // that the epochs advance in lockstep
    int push_ctas() const { return world > 1 ? std::max(1, std::min(128, n_chunks)) : 0; }
};

struct DistState
{
    int rank = 0, world = 1;
    long long halo_cap = 0;
    unsigned char *comm_buf = nullptr;
    size_t comm_bytes = 0;
    void *peer[kMaxRanks] = {};
    bool connected = false;
    bool poisoned = false; // a wait timed out: the ranks' counters may disagree; every call fails until psb200_dist_reset
    unsigned long long *counters = nullptr; // device: see dist_connect
    DistPlanHost plan;
    HaloPlan fine;         // halo exchange of the fine matrix
    unsigned nbr_mask = 0; // union of the neighbour masks of the fine matrix and of every partitioned AMG level
    DevBuf<double> vp2;    // second direction buffer (p ping-pongs so pushed values never race with the update)
    // values ingest: the CSC window [val_lo, val_hi) that holds every local entry is uploaded as one
    // contiguous copy and gathered on the device (d_perm[k] = perm[k] - val_lo)
    long long val_lo = 0, val_hi = 0;
    DevBuf<int> d_perm;
    DevBuf<double> d_csc_window;
    // rank-local diagonal block A[r0:r1, r0:r1] (halo columns dropped): the operator the rank-local AMG is built on
    CsrDev A_diag;
    DevBuf<int> diag_src; // A_diag.va[k] = A.va[diag_src[k]]
    ~DistState();
};

} // namespace psb
