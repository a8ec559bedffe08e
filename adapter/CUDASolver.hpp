#pragma once
// polysolve::linear::CUDASolver -- the "CUDA" entry of Solver::create: a thin PIMPL adapter from
// polysolve's C++ interface to the C ABI of libpsb200 (include/psb200.h).
//
// Drop-in template: reference src/polysolve/linear/MASSolver.hpp:35-72 (PIMPL behind Solver,
// registered by name in Solver.cpp:402-404). Integration steps: see INTEGRATION.md.

#include <polysolve/linear/Solver.hpp>

#include <memory>
#include <string>
#include <vector>

namespace polysolve::linear
{
    class CUDASolver : public Solver
    {
    public:
        CUDASolver();
        ~CUDASolver() override;
        POLYSOLVE_DELETE_MOVE_COPY(CUDASolver)

    public:
        // Reads params["CUDA"] (krylov, precond, tolerance, max_iter, amg{...}, ...) -- Solver.hpp:93
        void set_parameters(const json &params) override;

        // "solver_iter","solver_error" (EigenSolver.tpp:88-89) and "num_iterations","final_res_norm"
        // (AMGCL.cpp:142-143) plus "solver_status" -- Solver.hpp:96
        void get_info(json &params) const override;

        // Index work only (CSC->CSR map, SpMV tiling); cached on an unchanged pattern -- Solver.hpp:99
        void analyze_pattern(const StiffnessMatrix &A, const int precond_num) override;

        // Values upload + preconditioner build (Jacobi / SA-AMG hierarchy) -- Solver.hpp:102
        void factorize(const StiffnessMatrix &A) override;

        // x is in/out: initial guess on entry -- Solver.hpp:119-128
        void solve(const Ref<const VectorXd> b, Ref<VectorXd> x) override;

        void set_block_size(int block_size) override; // Solver.hpp:110
        void set_tolerance(const double tol) override; // Solver.hpp:116-117

        std::string name() const override { return "CUDA"; } // Solver.hpp:131

        // Second argument of Solver::create(name, precond) (Solver.cpp:307) -> value of params["CUDA"]["precond"].
        // The reference's names select an Eigen preconditioner class (Solver.cpp:230-305 PrecondHelper*): Diagonal ->
        // "jacobi", Identity -> "none"; "amg" / "AMGCL" select the SA-AMG V-cycle. Anything else (including the empty
        // string and Eigen's incomplete factorizations, which have no counterpart here) falls back to the default, as
        // PrecondHelper::create does.
        static std::string map_precond(const std::string &precond);

    public:
        // ---- beyond the Solver virtuals (optional). The reference's FEMSolver free functions
        // (FEMSolver.cpp:97-372) call these instead of their host-side triplet rebuild when `solver` is a CUDASolver
        // (see INTEGRATION.md section 5): masking and rhs lifting run on the GPU, A is not rewritten on the host.
        void dirichlet_solve(const StiffnessMatrix &A, Eigen::VectorXd &f, const std::vector<int> &dirichlet_nodes, Eigen::VectorXd &u,
                             const int precond_num);
        void prefactorize(const StiffnessMatrix &A, const std::vector<int> &dirichlet_nodes, const int precond_num);
        void dirichlet_solve_prefactorized(const StiffnessMatrix *A_or_null, Eigen::VectorXd &f, Eigen::VectorXd &u);

        // Newton step with device-resident data (Newton.cpp:173-214): values in the CSC order of the analyzed pattern,
        // reg_weight of RegularizedNewton (Newton.cpp:287-290), the residual ||H dx + g|| (Newton.cpp:207) with b = -g.
        void factorize_device(const long n, const long nnz, const double *d_vals, const double reg_weight = 0.0);
        void solve_device(const double *d_b, double *d_x, const long n);
        double residual_norm_device(const double *d_x, const double *d_b, const long n);

    private:
        struct Impl;
        std::unique_ptr<Impl> impl_;
    };
} // namespace polysolve::linear
