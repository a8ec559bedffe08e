// Smoothed-aggregation AMG hierarchy on the device (setup + cycle). See amg.cu.
#pragma once
#include "solver.hpp"

namespace psb {

struct AmgLevel;

class AmgHierarchy
{
public:
    AmgHierarchy(Ctx &ctx, const AmgParams &prm);
    ~AmgHierarchy();
    // Builds the hierarchy for the fine matrix A (borrowed: must outlive the hierarchy).
    void setup(const CsrDev &A, const std::vector<std::vector<int>> &imposed_aggregates);
    // x = M^-1 rhs : pre_cycles cycles from a zero initial guess (amgcl amg::apply).
    void apply(const double *rhs, double *x, const int *done);
    int num_levels() const;
    std::string info_json() const;
    const CsrDev &matrix(int level, int which) const; // 0 A, 1 P, 2 R
    int matrix_cols(int level, int which) const;
    // aggregate id per row of `level` (device pointer, rows(level) entries; -2 = removed); n_agg out
    const int *aggregates(int level, int *n_agg) const;

private:
    void cycle(int l, const double *rhs, double *&x, double *&x_alt, bool x_is_zero, const int *done);
    void relax(int l, const double *rhs, double *&x, double *&x_alt, bool x_is_zero, const int *done);
    Ctx &ctx_;
    AmgParams prm_;
    const CsrDev *A0_ = nullptr;
    std::vector<std::unique_ptr<AmgLevel>> levels_;
};

} // namespace psb
