// Internal solver object behind the C ABI (include/psb200.h).
#pragma once
#include "json_mini.hpp"
#include "launch.cuh"

#include <functional>
#include <memory>
#include <string>
#include <vector>

namespace psb {

// Mirrors the AMGCL parameter tree polysolve passes (reference AMGCL.cpp:32-65 default_params()).
struct AmgParams
{
    int max_levels = 6;
    int coarse_enough = 3000;
    bool direct_coarse = false;
    int ncycle = 2;
    int npre = 1, npost = 1, pre_cycles = 1;
    std::string relax_type = "chebyshev"; // chebyshev | damped_jacobi
    int degree = 16;
    int power_iters = 100;
    double higher = 2.0, lower = 0.008333333333;
    bool scale = true;
    double damping = 0.72;
    double sa_relax = 1.0;
    bool estimate_spectral_radius = true;
    double eps_strong = 0.0;
    int block_size = 1; // AMGCL_Block<B> (reference AMGCL.cpp:246-298): B x B value type
    std::string aggregation = "mis2"; // mis2 (parallel, deterministic) | imposed via debug hook
    // row partitions: "partitioned" = every level is row-partitioned (decoupled aggregation, rank-local P / R, distributed
    // Galerkin product, per-level halo exchange) down to the levels below replicate_below non-zeros, which are replicated;
    // "global" = one hierarchy of the whole matrix on every rank, only level 0 partitioned (the iteration counts of the
    // 1-GPU run, no memory partitioning); "local" = every rank its own hierarchy of its diagonal block (block-Jacobi)
    std::string dist_mode = "partitioned";
    // partitioned mode: levels with fewer stored non-zeros (summed over the ranks) than this are replicated on every rank.
    // One SpMV of a replicated level costs its whole matrix on every GPU, a partitioned one 1 / world of it plus a halo
    // push and wait (~10 us): the two break even near 7 M non-zeros on B200 (12 B per non-zero at ~6.5 TB/s).
    long long replicate_below = 8000000;
    // partitioned mode: the Chebyshev steps push their boundary rows from the SpMV epilogue (boundary tiles first, three
    // halo buffers) instead of a separate push kernel. Off by default: measured on 2 B200s it LOSES (C3 solve 0.0958 s vs
    // 0.0882 s, C4 1.04 s vs 0.95 s: the per-row device fence + atomic + NVLink store sit on the critical path of the
    // boundary tiles, ~12 us per launch, more than the 4 us push kernel they replace)
    bool fused_push = false;
    bool same_as(const AmgParams &o) const
    {
        return max_levels == o.max_levels && coarse_enough == o.coarse_enough && direct_coarse == o.direct_coarse && ncycle == o.ncycle &&
               npre == o.npre && npost == o.npost && pre_cycles == o.pre_cycles && relax_type == o.relax_type && degree == o.degree &&
               power_iters == o.power_iters && higher == o.higher && lower == o.lower && scale == o.scale && damping == o.damping &&
               sa_relax == o.sa_relax && estimate_spectral_radius == o.estimate_spectral_radius && eps_strong == o.eps_strong &&
               aggregation == o.aggregation && dist_mode == o.dist_mode && replicate_below == o.replicate_below && fused_push == o.fused_push;
    }
};

struct Params
{
    std::string krylov = "cg";      // cg | cg1r (single-reduction CG, row partitions) | bicgstab
    std::string precond = "jacobi"; // jacobi | amg | none
    double tolerance = 1e-12;       // relative to ||b|| (Eigen / AMGCL semantics); spec default of Eigen iterative solvers
    int max_iter = 1000;
    int check_every = 16;
    bool use_graph = true;
    // row partitions: the stream SpMV walks the tiles without halo columns first and waits for the halo before its first
    // boundary tile. Off by default: A/B on 2 B200s (profiles/r01_interior_first.txt) 5780 vs 5515 it/s -- the extra
    // dependent load of the tile order in the TMA issue path costs more than the hidden halo latency at that size
    bool interior_first = false;
    // programmatic dependent launch between the kernels of the Krylov chain. Off by default: measured on B200
    // (profiles/r01_pdl.txt) it gains nothing at 1.26 M rows and loses 9 % at 10 M rows (dependents that become
    // resident while the predecessor drains land unevenly on the SMs, and the vector kernels are statically partitioned)
    bool pdl = false;
    std::string spmv_kernel = "auto";
    std::string cg_kernel = "auto"; // Jacobi-PCG schedule: split (one kernel per phase) | auto (= split; the persistent cooperative kernel of round 1 measured slower and was removed)
    int device = -1; // -1: current device
    int block_size = 1;
    bool profile = false;
    bool verify_pattern = true;
    double comm_timeout_s = 3.0; // row partitions: a kernel that waits longer than this for a peer fails the solve
    AmgParams amg;
};

class AmgHierarchy; // amg.cu
class AmgDist;      // amg_dist.cu

struct DistState; // dist.hpp / dist.cu
struct HaloPlan;

struct Solver
{
    Params prm;
    std::string err;
    std::string info_json = "{}";
    int device = 0;
    Ctx ctx;

    // multi-GPU (row-partitioned) mode: set by dist_prepare/dist_connect. In this mode n, nnz, A and all
    // vectors describe the LOCAL row range; n_global / nnz_global the whole system.
    std::unique_ptr<DistState> dist;
    long long n_global = 0, nnz_global = 0;
    void dist_prepare(int rank, int world, long long halo_cap, char handle_out[64]);
    void dist_connect(const char *handles);
    void analyze_pattern_dist(long long n, long long nnz, const int *outer, const int *inner);
    void factorize_values_dist(const double *vals);
    void run_cg_eigen_dist(const double *d_b);
    void run_cg_amgcl_dist(const double *d_b);
    void run_cg1r(const double *d_b); // single-reduction PCG (any world size)
    void build_diag_block_dist();
    void check_comm_error();
    void check_not_poisoned() const;
    void dist_reset();
    void dist_set_nbr_mask(unsigned mask);
    // host-level collectives over the comm buffer (setup time; every rank calls them in the same order)
    void dist_allgather8(const double in[8], double out[64]); // out[q * 8 + i] = in[i] of rank q; barrier semantics
    void dist_barrier();
    double dist_max(double v);
    void dist_gather_ll(long long v, long long out[kMaxRanks]);
    void dist_alltoallv(const void *const send[kMaxRanks], const size_t send_bytes[kMaxRanks], void *const recv[kMaxRanks], size_t recv_bytes[kMaxRanks],
                        const size_t *expect_recv = nullptr);
    void push_halo(const HaloPlan &hp, const double *d_v, const int *done);
    void bulk_allgather(const double *d_mine, double *d_out, const long long *offsets, const int *done);

    // pattern state (analyze_pattern)
    long long n = 0, nnz = 0;
    unsigned long long pattern_hash = 0;
    bool analyzed = false, factorized = false;
    bool sym_pattern = false;
    int pattern_block = 1;   // block size the stored CSR pattern was expanded for (analyze_pattern)
    long long nnz_input = 0; // nnz of the matrix as received (nnz counts the expanded pattern in block mode)
    int precond_num = 0;
    DevBuf<int> csc_outer, csc_inner, perm;
    DevBuf<double> csc_vals;
    CsrDev A;
    DevBuf<double> dinv;
    long long n_pad = 0;

    // work vectors (padded, zero tails)
    DevBuf<double> vb, vx, vr, vp, vq, vz, vy, vv, vt, vr0;
    KState *d_state = nullptr;
    KState *h_state = nullptr; // pinned, 4 slots
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaGraphExec_t graph_exec = nullptr;
    std::string graph_key;
    long long graph_launches_per_batch = 0;

    // Row partition + AMG, "global" mode: a helper instance (not partitioned, same device) that holds the WHOLE matrix in
    // CSR form; the hierarchy is built from it on every rank and only level 0 of the cycle is partitioned (amg.hpp).
    std::unique_ptr<Solver> full_;
    bool amg_global() const;
    bool amg_partitioned() const;
    std::unique_ptr<AmgDist> amg_dist; // row partition, amg.dist_mode = partitioned
    void factorize_full_for_amg(const double *h_vals, const double *d_vals, double diag_shift);
    std::unique_ptr<AmgHierarchy> amg;
    std::vector<std::vector<int>> imposed_aggregates;

    // last solve
    long long last_iters = 0;
    double last_error = 0;
    int last_status = 0;
    double t_analyze_ms = 0, t_factorize_ms = 0, t_solve_ms = 0, t_setup_precond_ms = 0;
    bool analyze_skipped = false;

    Solver();
    ~Solver();
    void set_parameters(const std::string &json);
    unsigned long long pattern_hash_of(long long n, long long nnz, const int *outer, const int *inner) const;
    void analyze_pattern(long long n, long long nnz, const int *outer, const int *inner, int precond_num);
    void factorize(long long n, long long nnz, const int *outer, const int *inner, const double *vals);
    void factorize_device(long long n, long long nnz, const double *d_vals, double diag_shift);
    void factorize_tail(double t0);
    double residual_norm_device(const double *d_x, const double *d_b, long long n);
    double residual_norm_host(const double *x, const double *b, long long n);
    void dist_allgather_host(double *x_full, long long n);
    void push_halo_of(const double *d_v, const int *done = nullptr); // row partition: push the boundary entries of a local vector (dist.cu)
    void bulk_allreduce(const double *d_partial, double *d_out, long long len, const int *done = nullptr); // sum of a vector across ranks
    // Dirichlet pre-processing (fem.cu; reference FEMSolver.cpp:97-372)
    DevBuf<unsigned char> dmask; // N: 1 for a Dirichlet dof
    void dirichlet_set_nodes(const int *nodes, long long count);
    void dirichlet_upload(long long n, long long nnz, const int *outer, const int *inner, const double *vals, int precond_num);
    void dirichlet_lift(const double *d_f, double *d_g);
    void dirichlet_mask_matrix();
    void dirichlet_solve(long long n, long long nnz, const int *outer, const int *inner, const double *vals, double *f, const int *nodes,
                         long long n_nodes, double *u, int precond_num);
    void dirichlet_prefactorize(long long n, long long nnz, const int *outer, const int *inner, const double *vals, const int *nodes,
                                long long n_nodes, int precond_num);
    void dirichlet_solve_prefactorized(const double *vals, double *f, double *u, long long n);
    void gather_values_to_csr(const double *d_csc_vals); // A.va[k] = d_csc_vals[perm[k]] on the solver's stream
    void run_solver(const double *d_b);                  // the configured Krylov method on vx (in/out) and d_b
    void solve_host(const double *b, double *x, long long n);
    void solve_device(const double *d_b, double *d_x, long long n);
    void spmv_host(const double *x, double *y, long long n);
    double bench_spmv(const std::string &kernel, int reps);
    void precond_apply_host(const double *r, double *z, long long n);
    void build_info();

    void ensure_vectors();
    void run_cg_eigen(const double *d_b);
    void run_cg_amgcl(const double *d_b);
    void run_bicgstab(const double *d_b);
    void drive(const std::function<void()> &enqueue_batch, int batch_iters, const std::string &key);
    void finish_solve();
};

void ensure_ctx(Solver &s);
long long expand_block_pattern(Ctx &ctx, int B, long long n, DevBuf<int> &rp, DevBuf<int> &ci, DevBuf<int> &perm);
void sort_rows_by_column(Ctx &ctx, long long n, DevBuf<int> &rp, DevBuf<int> &ci, DevBuf<int> &perm);
void init_state(Solver &s, double tol, int max_iter);

} // namespace psb
