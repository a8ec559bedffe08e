// Executed C++ test of the drop-in boundary: links adapter/CUDASolver.cpp (the polysolve::linear::Solver subclass) against
// libpsb200.so and drives it exactly like the reference's own test does (tests/test_linear_solver.cpp:103-164: create by
// name, set_parameters, analyze_pattern, factorize, solve, get_info, residual check). Eigen and polysolve are not in this
// image, so the matrix/vector types are the API stubs of adapter/stub (same member names as Eigen's); nlohmann::json is
// the real header. The `create` below is the registration branch of INTEGRATION.md section 2, verbatim.
//
// Config 1 of BASELINE.json: 32 x 32 2-D Poisson, b = splitmix64(42), Jacobi-PCG, tol 1e-10 -> 115 iterations.
#include "CUDASolver.hpp"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>

using polysolve::json;
using polysolve::StiffnessMatrix;
using namespace polysolve::linear;

// --- INTEGRATION.md section 2, Solver.cpp:400-405 (the branch added next to "MAS")
static std::unique_ptr<Solver> create(const std::string &solver, const std::string &precond)
{
    if (solver == "CUDA")
    {
        auto s = std::make_unique<CUDASolver>();
        if (!precond.empty())
            s->set_parameters(json{{"CUDA", {{"precond", CUDASolver::map_precond(precond)}}}});
        return s;
    }
    throw std::runtime_error("Unrecognized solver type: " + solver); // Solver.cpp:495
}

static StiffnessMatrix poisson2d(int m)
{
    StiffnessMatrix A;
    const int n = m * m;
    A.rows_ = A.cols_ = n;
    A.outer_.push_back(0);
    for (int c = 0; c < n; ++c)
    {
        const int i = c / m, j = c % m;
        auto put = [&](int r, double v) { A.inner_.push_back(r); A.vals_.push_back(v); };
        if (i > 0) put(c - m, -1);
        if (j > 0) put(c - 1, -1);
        put(c, 4);
        if (j < m - 1) put(c + 1, -1);
        if (i < m - 1) put(c + m, -1);
        A.outer_.push_back((int)A.inner_.size());
    }
    return A;
}

static double splitmix_next(uint64_t &s)
{
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return 2.0 * (double)(z >> 11) * (1.0 / 9007199254740992.0) - 1.0;
}

static double residual(const StiffnessMatrix &A, const Eigen::VectorXd &x, const Eigen::VectorXd &b)
{
    std::vector<double> y(b.v.size(), 0.0);
    for (int c = 0; c < (int)A.cols(); ++c)
        for (int k = A.outer_[c]; k < A.outer_[c + 1]; ++k)
            y[A.inner_[k]] += A.vals_[k] * x.v[c];
    double s = 0;
    for (size_t i = 0; i < y.size(); ++i)
        s += (y[i] - b.v[i]) * (y[i] - b.v[i]);
    return std::sqrt(s);
}

#define REQUIRE(cond)                                                        \
    do                                                                       \
    {                                                                        \
        if (!(cond))                                                         \
        {                                                                    \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);    \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main()
{
    REQUIRE(CUDASolver::map_precond("Eigen::DiagonalPreconditioner") == "jacobi");
    REQUIRE(CUDASolver::map_precond("Eigen::IdentityPreconditioner") == "none");
    REQUIRE(CUDASolver::map_precond("amg") == "amg");
    bool threw = false;
    try { create("NoSuchSolver", ""); } catch (const std::runtime_error &) { threw = true; }
    REQUIRE(threw);

    const StiffnessMatrix A = poisson2d(32);
    Eigen::VectorXd b, x;
    b.resize(A.rows());
    x.resize(A.rows());
    x.setZero();
    uint64_t seed = 42;
    for (auto &v : b.v)
        v = splitmix_next(seed);
    REQUIRE(std::fabs(b.v[0] - 0.4831297575436466) < 1e-15);

    // test_linear_solver.cpp:141-158
    auto solver = create("CUDA", "Eigen::DiagonalPreconditioner");
    REQUIRE(solver->name() == "CUDA");
    json params;
    params["CUDA"]["tolerance"] = 1e-10;
    params["CUDA"]["max_iter"] = 1000;
    solver->set_parameters(params);
    solver->analyze_pattern(A, (int)A.rows());
    solver->factorize(A);
    solver->solve(b, x);
    json info;
    solver->get_info(info);
    std::printf("solver_iter=%d solver_error=%.6e status=%s\n", info["solver_iter"].get<int>(), info["solver_error"].get<double>(),
                info["solver_status"].get<std::string>().c_str());
    REQUIRE(info["solver_iter"].get<int>() == 115);       // SURVEY A.5 known answer
    REQUIRE(info["num_iterations"].get<int>() == 115);    // AMGCL-style key too (AMGCL.cpp:142-143)
    REQUIRE(info["solver_error"].get<double>() < 1e-10);
    const double err = residual(A, x, b);
    std::printf("||Ax-b|| = %.6e\n", err);
    REQUIRE(err < 1e-8);                                  // test_linear_solver.cpp:160-162

    // warm start from the converged x: 0 iterations (test_linear_solver.cpp:432-450)
    solver->solve(b, x);
    solver->get_info(info);
    REQUIRE(info["solver_iter"].get<int>() == 0);

    // the AMG preconditioner through the same boundary, AMGCL-style keys (test_linear_solver.cpp:400-455)
    auto amg = create("CUDA", "amg");
    json p2;
    p2["CUDA"]["tolerance"] = 1e-10;
    p2["CUDA"]["amg"]["coarse_enough"] = 100;
    amg->set_parameters(p2);
    amg->analyze_pattern(A, (int)A.rows());
    amg->factorize(A);
    x.setZero();
    amg->solve(b, x);
    amg->get_info(info);
    std::printf("amg num_iterations=%d final_res_norm=%.3e\n", info["num_iterations"].get<int>(), info["final_res_norm"].get<double>());
    REQUIRE(info["num_iterations"].get<int>() > 0);
    REQUIRE(info["final_res_norm"].get<double>() < 1e-10);
    REQUIRE(residual(A, x, b) < 1e-8);
    amg->solve(b, x);
    amg->get_info(info);
    REQUIRE(info["num_iterations"].get<int>() == 0);

    // errors surface as std::runtime_error (Utils.cpp:65-69), recoverable by Newton (Newton.cpp:191-202)
    threw = false;
    try
    {
        json bad;
        bad["CUDA"]["krylov"] = "gmres";
        solver->set_parameters(bad);
    }
    catch (const std::runtime_error &) { threw = true; }
    REQUIRE(threw);
    std::printf("adapter test OK\n");
    return 0;
}
