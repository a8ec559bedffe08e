// polysolve::linear::CUDASolver: forwards every virtual to the C ABI of libpsb200.
// Error convention: any non-zero status becomes std::runtime_error (reference
// src/polysolve/Utils.cpp:65-69), so Newton's factorize fallback keeps working
// (reference src/polysolve/nonlinear/descent_strategies/Newton.cpp:191-202).
#include "CUDASolver.hpp"

#include <psb200.h>

#include <stdexcept>
#include <vector>

namespace polysolve::linear
{
    struct CUDASolver::Impl
    {
        psb200_handle h = nullptr;

        void check(int rc, const char *where) const
        {
            if (rc != PSB200_OK)
                throw std::runtime_error(std::string("[CUDA] ") + where + ": " + psb200_last_error(h));
        }
    };

    CUDASolver::CUDASolver() : impl_(std::make_unique<Impl>())
    {
        if (psb200_create(&impl_->h, nullptr) != PSB200_OK)
            throw std::runtime_error(std::string("[CUDA] psb200_create: ") + psb200_last_error(nullptr));
    }

    CUDASolver::~CUDASolver()
    {
        if (impl_ && impl_->h)
            psb200_destroy(impl_->h);
    }

    std::string CUDASolver::map_precond(const std::string &precond)
    {
        if (precond == "Eigen::IdentityPreconditioner" || precond == "none")
            return "none";
        if (precond == "amg" || precond == "AMGCL" || precond == "AMG")
            return "amg";
        return "jacobi"; // "Eigen::DiagonalPreconditioner", "", unknown names
    }

    void CUDASolver::set_parameters(const json &params)
    {
        // polysolve namespaces parameters by solver name (params["CUDA"]...), cf. MASSolver.cu:605-614
        if (params.contains("CUDA"))
        {
            json sub;
            sub["CUDA"] = params["CUDA"];
            impl_->check(psb200_set_parameters(impl_->h, sub.dump().c_str()), "set_parameters");
        }
    }

    void CUDASolver::get_info(json &params) const
    {
        size_t need = 0;
        std::vector<char> buf(1 << 14);
        int rc = psb200_get_info(impl_->h, buf.data(), buf.size(), &need);
        if (rc != PSB200_OK && need > buf.size())
        {
            buf.resize(need);
            rc = psb200_get_info(impl_->h, buf.data(), buf.size(), &need);
        }
        impl_->check(rc, "get_info");
        const json info = json::parse(buf.data());
        for (auto it = info.begin(); it != info.end(); ++it)
            params[it.key()] = it.value();
    }

    void CUDASolver::analyze_pattern(const StiffnessMatrix &A, const int precond_num)
    {
        // uncompressed matrices must be compressed first (reference mas_utils/BSRMatrix.cu:444-452)
        if (!A.isCompressed())
        {
            StiffnessMatrix Ac = A;
            Ac.makeCompressed();
            impl_->check(psb200_analyze_pattern_csc(impl_->h, Ac.rows(), Ac.nonZeros(), Ac.outerIndexPtr(), Ac.innerIndexPtr(), precond_num), "analyze_pattern");
            return;
        }
        impl_->check(psb200_analyze_pattern_csc(impl_->h, A.rows(), A.nonZeros(), A.outerIndexPtr(), A.innerIndexPtr(), precond_num), "analyze_pattern");
    }

    void CUDASolver::factorize(const StiffnessMatrix &A)
    {
        // The matrix is borrowed only for this call; values (and, if unknown, the pattern) are copied to the device.
        if (!A.isCompressed())
        {
            StiffnessMatrix Ac = A;
            Ac.makeCompressed();
            impl_->check(psb200_factorize_csc(impl_->h, Ac.rows(), Ac.nonZeros(), Ac.outerIndexPtr(), Ac.innerIndexPtr(), Ac.valuePtr()), "factorize");
            return;
        }
        impl_->check(psb200_factorize_csc(impl_->h, A.rows(), A.nonZeros(), A.outerIndexPtr(), A.innerIndexPtr(), A.valuePtr()), "factorize");
    }

    void CUDASolver::solve(const Ref<const VectorXd> b, Ref<VectorXd> x)
    {
        if (x.size() != b.size())
            throw std::runtime_error("[CUDA] solve: x and b differ in size");
        // Ref<VectorXd> has unit inner stride, so data() is a plain contiguous array
        impl_->check(psb200_solve(impl_->h, b.data(), x.data(), b.size()), "solve");
    }

    void CUDASolver::dirichlet_solve(const StiffnessMatrix &A, Eigen::VectorXd &f, const std::vector<int> &dirichlet_nodes, Eigen::VectorXd &u,
                                     const int precond_num)
    {
        if (!A.isCompressed())
            throw std::runtime_error("[CUDA] dirichlet_solve: the matrix must be compressed");
        if (u.size() != A.rows())
        {
            u.resize(A.rows()); // FEMSolver.cpp:268-272
            u.setZero();
        }
        impl_->check(psb200_dirichlet_solve(impl_->h, A.rows(), A.nonZeros(), A.outerIndexPtr(), A.innerIndexPtr(), A.valuePtr(), f.data(),
                                            dirichlet_nodes.data(), (int64_t)dirichlet_nodes.size(), u.data(), precond_num),
                     "dirichlet_solve");
    }

    void CUDASolver::prefactorize(const StiffnessMatrix &A, const std::vector<int> &dirichlet_nodes, const int precond_num)
    {
        if (!A.isCompressed())
            throw std::runtime_error("[CUDA] prefactorize: the matrix must be compressed");
        impl_->check(psb200_dirichlet_prefactorize(impl_->h, A.rows(), A.nonZeros(), A.outerIndexPtr(), A.innerIndexPtr(), A.valuePtr(),
                                                   dirichlet_nodes.data(), (int64_t)dirichlet_nodes.size(), precond_num),
                     "prefactorize");
    }

    void CUDASolver::dirichlet_solve_prefactorized(const StiffnessMatrix *A_or_null, Eigen::VectorXd &f, Eigen::VectorXd &u)
    {
        if (u.size() != f.size())
        {
            u.resize(f.size()); // FEMSolver.cpp:363-367
            u.setZero();
        }
        impl_->check(psb200_dirichlet_solve_prefactorized(impl_->h, A_or_null ? A_or_null->valuePtr() : nullptr, f.data(), u.data(), f.size()),
                     "dirichlet_solve_prefactorized");
    }

    void CUDASolver::factorize_device(const long n, const long nnz, const double *d_vals, const double reg_weight)
    {
        impl_->check(psb200_factorize_csc_device(impl_->h, n, nnz, d_vals, reg_weight), "factorize_device");
    }

    void CUDASolver::solve_device(const double *d_b, double *d_x, const long n)
    {
        impl_->check(psb200_solve_device(impl_->h, d_b, d_x, n), "solve_device");
    }

    double CUDASolver::residual_norm_device(const double *d_x, const double *d_b, const long n)
    {
        double r = 0;
        impl_->check(psb200_residual_norm_device(impl_->h, d_x, d_b, n, &r), "residual_norm_device");
        return r;
    }

    void CUDASolver::set_block_size(int block_size)
    {
        impl_->check(psb200_set_block_size(impl_->h, block_size), "set_block_size");
    }

    void CUDASolver::set_tolerance(const double tol)
    {
        impl_->check(psb200_set_tolerance(impl_->h, tol), "set_tolerance");
    }
} // namespace polysolve::linear
