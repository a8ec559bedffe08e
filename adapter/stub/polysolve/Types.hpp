// compile-check stub mirroring reference src/polysolve/Types.hpp:11-17
#pragma once
#include <Eigen/Sparse>
#include <nlohmann/json.hpp>
namespace polysolve
{
    typedef Eigen::SparseMatrix<double, Eigen::ColMajor, int> StiffnessMatrix;
    using json = nlohmann::json;
} // namespace polysolve
