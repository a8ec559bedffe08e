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
    // align: row offsets are rounded up to a multiple of `align` (block problems keep the rows of a node together)
    void build(long long n, long long nnz, const int *outer, const int *inner, int rank_, int world_, long long halo_cap_, int align = 1);
};

// Device view of the send side of the halo exchange (see push_section, dist.cu)
struct PushList
{
    const int *rows;                                          // local row ids, grouped by destination
    const int *chunk_peer, *chunk_start, *chunk_cnt, *chunk_off; // per chunk: destination, first entry in rows, entries, offset in the destination's region
    int nchunks;
};

struct DistState
{
    int rank = 0, world = 1;
    long long halo_cap = 0;
    unsigned char *comm_buf = nullptr;
    size_t comm_bytes = 0;
    void *peer[kMaxRanks] = {};
    bool connected = false;
    unsigned long long *counters = nullptr; // device: [0] red_seq, [1] push_epoch, [2] error (int)
    DistPlanHost plan;
    // device push list: send rows of all destinations concatenated + the chunk table [peer | start | count | offset]
    DevBuf<int> push_rows, chunk_tab;
    int n_push = 0, n_chunks = 0;
    unsigned send_mask = 0, recv_mask = 0;
    DevBuf<double> vp2; // second direction buffer (p ping-pongs so pushed values never race with the update)
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

PushList make_push(DistState &d);

} // namespace psb
