// Row-partitioned smoothed-aggregation AMG: every level above `replicate_below` non-zeros is partitioned across
// the ranks like the fine matrix, the small levels below are replicated. See amg_dist.cu.
#pragma once
#include "amg.hpp"
#include "dist.hpp"

#include <memory>
#include <string>
#include <vector>

namespace psb {

struct DistAmgLevel;

class AmgDist
{
public:
    AmgDist(Solver &s, const AmgParams &prm);
    ~AmgDist();
    // Collective: builds the hierarchy from the solver's local rows (s.A, halo plan s.dist->fine).
    void setup(const std::vector<std::vector<int>> &imposed_aggregates);
    // x_local = M^-1 rhs_local (collective)
    void apply(const double *rhs_local, double *x_local, const int *done);
    std::string info_json() const;
    int num_levels() const;

private:
    void cycle(int l, const double *rhs, double *&x, double *&x_alt, bool x_is_zero, const int *done);
    void relax(int l, const double *rhs, double *&x, double *&x_alt, bool x_is_zero, const int *done);
    void refinalize_plans();
    Solver &s_;
    AmgParams prm_;
    std::vector<std::unique_ptr<DistAmgLevel>> levels_; // partitioned levels
    std::unique_ptr<AmgHierarchy> tail_;                // replicated levels (level index base = levels_.size())
    CsrDev tail_A_;                                     // the replicated matrix the tail starts from
    std::vector<long long> tail_offsets_;               // world + 1: slice of every rank in the tail's level-0 vectors
    double t_setup_ms_ = 0;
};

} // namespace psb
