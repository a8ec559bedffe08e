// TEST INFRASTRUCTURE ONLY -- CPU oracle for the polysolve hot path.
// Nothing under oracle/ may be imported, linked or executed by the product
// (polysolve_b200/, include/, adapter/). Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
//
// PARITY UNPINNED: Eigen 5.0.1 and AMGCL 1.4.3 (the libraries that hold the
// arithmetic of this path, fetched by CPM in the reference:
// cmake/recipes/eigen.cmake:23-28, cmake/recipes/amgcl.cmake:44-48) are absent
// from /root/reference and from this image, and the reference's tests hold no
// golden vectors (residual bounds only, tests/test_linear_solver.cpp:160-162).
// This file restates the published algorithms; it is pinned only against
// (i) analytic eigenpairs of the Dirichlet Laplacian, (ii) scipy, and
// (iii) the known answers recorded in SURVEY.md Appendix A.5.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>

namespace orc {

struct Csr {
    int64_t n = 0;
    std::vector<int32_t> ptr, col;
    std::vector<double> val;
    int64_t nnz() const { return (int64_t)col.size(); }
};

// splitmix64 stream -> U(-1,1); the generator SURVEY.md section 8(d) fixes for b / x*.
inline uint64_t splitmix64_next(uint64_t &s)
{
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
inline double splitmix64_unit(uint64_t &s)
{
    return 2.0 * (double)(splitmix64_next(s) >> 11) * (1.0 / 9007199254740992.0) - 1.0;
}

void spmv_csr(const Csr &A, const double *x, double *y);

} // namespace orc
