// Smoothed-aggregation AMG hierarchy on the device (setup + cycle). See amg.cu.
#pragma once
#include "solver.hpp"

#include <functional>

namespace psb {

struct AmgLevel;
struct AmgDistFine;

class AmgHierarchy
{
public:
    AmgHierarchy(Ctx &ctx, const AmgParams &prm);
    ~AmgHierarchy();
    // Builds the hierarchy for the fine matrix A (borrowed: must outlive the hierarchy). level_base > 0: A is level
    // `level_base` of a larger hierarchy whose upper levels are row-partitioned (amg_dist.cu); it counts against
    // max_levels, selects the imposed aggregates and the power-iteration seeds, and halves eps_strong accordingly.
    void setup(const CsrDev &A, const std::vector<std::vector<int>> &imposed_aggregates, int level_base = 0);
    // x = M^-1 rhs : pre_cycles cycles from a zero initial guess (amgcl amg::apply).
    void apply(const double *rhs, double *x, const int *done);
    int num_levels() const;
    std::string info_json() const;
    std::string levels_json() const; // the comma-separated level objects of info_json
    double total_nnz() const;
    // one cycle from level 0 on caller-visible buffers (the replicated tail of a partitioned hierarchy):
    // f / u / u_alt are the level-0 work vectors; after cycle0() the result is in the returned pointer
    double *level0_f();
    double *cycle0(const int *done);
    const CsrDev &matrix(int level, int which) const; // 0 A, 1 P, 2 R
    int matrix_cols(int level, int which) const;
    // aggregate id per row of `level` (device pointer, rows(level) entries; -2 = removed); n_agg out
    const int *aggregates(int level, int *n_agg) const;

    // Row-partitioned use (multi-GPU): the hierarchy was built from the WHOLE matrix on every rank, the cycle runs
    // level 0 on the rank's own rows -- smoothing and residual with the partitioned matrix A_local (halo columns, one
    // halo push per multiplied vector), restriction as a partial product summed across the ranks, prolongation of the
    // rank's rows -- and the levels below replicated on every rank. Same hierarchy, same arithmetic as the single-GPU
    // cycle up to summation order, so the iteration counts are those of the 1-GPU run.
    //   row0           first global row of this rank; A_local.n rows
    //   dinv_local     D^-1 of the local rows (padded, zero tail)
    //   push_halo(v, done)                  pushes the boundary entries of the local vector v to the neighbours
    //   allreduce(partial, out, len, done)  out = sum over ranks of partial
    void setup_dist_fine(const CsrDev &A_local, const double *dinv_local, long long row0,
                         std::function<void(const double *, const int *)> push_halo,
                         std::function<void(const double *, double *, long long, const int *)> allreduce);
    bool has_dist_fine() const { return (bool)dist_; }
    // x_local = M^-1 rhs_local (collective: every rank calls it with its slice)
    void apply_dist(const double *rhs_local, double *x_local, const int *done);

private:
    void cycle(int l, const double *rhs, double *&x, double *&x_alt, bool x_is_zero, const int *done);
    void relax(int l, const double *rhs, double *&x, double *&x_alt, bool x_is_zero, const int *done);
    Ctx &ctx_;
    AmgParams prm_;
    const CsrDev *A0_ = nullptr;
    int level_base_ = 0;
    std::vector<std::unique_ptr<AmgLevel>> levels_;
    std::shared_ptr<AmgDistFine> dist_;
    void relax_dist(const double *rhs, double *&x, double *&x_alt, bool x_is_zero, const int *done);
    void cycle_dist(const double *rhs, double *&x, double *&x_alt, bool x_is_zero, const int *done);
};

} // namespace psb
