// compile-check stub mirroring the virtual interface of reference src/polysolve/linear/Solver.hpp:31-132
#pragma once
#include <polysolve/Types.hpp>
#include <memory>
#include <string>
#define POLYSOLVE_DELETE_MOVE_COPY(Base) \
    Base(Base &&) = delete;              \
    Base &operator=(Base &&) = delete;   \
    Base(const Base &) = delete;         \
    Base &operator=(const Base &) = delete;
namespace polysolve::linear
{
    class Solver
    {
    public:
        typedef Eigen::VectorXd VectorXd;
        template <typename T>
        using Ref = Eigen::Ref<T>;
        virtual ~Solver() = default;
    protected:
        Solver() = default;
    public:
        virtual void set_parameters(const json &params) {}
        virtual void get_info(json &params) const {};
        virtual void analyze_pattern(const StiffnessMatrix &A, const int precond_num) {}
        virtual void factorize(const StiffnessMatrix &A) {}
        virtual void analyze_pattern_dense(const Eigen::MatrixXd &A, const int precond_num) {}
        virtual void factorize_dense(const Eigen::MatrixXd &A) {}
        virtual bool is_dense() const { return false; }
        virtual void set_block_size(int block_size) {}
        virtual void set_is_nullspace(const VectorXd &x) {}
        virtual void set_tolerance(const double tol) {}
        virtual void solve(const Ref<const VectorXd> b, Ref<VectorXd> x) = 0;
        virtual std::string name() const { return ""; }
    };
} // namespace polysolve::linear
