// Newton / line-search driver over the psb200 C ABI (include/psb200_nl.h).
//
// Host-side control flow only -- the reference's nonlinear loop is host code too. It restates
//   polysolve::nonlinear::Solver::minimize            (reference src/polysolve/nonlinear/Solver.cpp:255-582)
//   Newton::create_solver / compute_update_direction  (descent_strategies/Newton.cpp:14-58,144-214)
//   RegularizedNewton                                 (Newton.cpp:275-291 hessian += w I, :326-330 handle_error)
//   GradientDescent fallback                          (Solver.cpp:175-181, GradientDescent.cpp:18-34)
//   LineSearch::line_search, Backtracking, Armijo     (line_search/LineSearch.cpp:73-254, Backtracking.cpp:15-83, Armijo.cpp:13-32)
//   checkConvergence                                  (Criteria.cpp:59-96)
// and talks to the linear solver exclusively through psb200_* entry points, i.e. through the same
// analyze_pattern -> factorize -> solve -> get_info sequence Newton.cpp:189-211 issues.
#include "../../include/psb200.h"
#include "../../include/psb200_nl.h"
#include "json_mini.hpp"

#include <algorithm>
#include <cfenv>
#include <chrono>
#include <cmath>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace {

using psb::JValue;
using psb::jnum;
using psb::jstr;
constexpr double NaN = std::numeric_limits<double>::quiet_NaN();

double now_s()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

struct ScopedTimer
{
    double &acc;
    double t0;
    explicit ScopedTimer(double &a) : acc(a), t0(now_s()) {}
    ~ScopedTimer() { acc += now_s() - t0; }
};

// reference Criteria.hpp:12-30 (Status) and :34-57 (Criteria)
enum class Status
{
    NotStarted,
    Continue,
    IterationLimit,
    XDeltaTolerance,
    FDeltaTolerance,
    GradNormTolerance,
    RelGradNormTolerance,
    RelXDeltaTolerance,
    NewtonDecrementTolerance,
    ObjectiveCustomStop,
    NanEncountered,
    NotDescentDirection,
    LineSearchFailed,
    UpdateDirectionFailed
};

const char *status_message(Status s)
{
    switch (s)
    {
    case Status::NotStarted: return "Solver not started";
    case Status::Continue: return "Convergence criteria not reached";
    case Status::IterationLimit: return "Iteration limit reached";
    case Status::XDeltaTolerance: return "Change in parameter vector too small";
    case Status::FDeltaTolerance: return "Change in cost function value too small";
    case Status::GradNormTolerance: return "Gradient vector norm too small";
    case Status::RelGradNormTolerance: return "Relative gradient vector too small";
    case Status::RelXDeltaTolerance: return "Relative change in parameter vector too small";
    case Status::NewtonDecrementTolerance: return "Newton decrement too small";
    case Status::ObjectiveCustomStop: return "Objective function specified to stop";
    case Status::NanEncountered: return "Objective or gradient function returned NaN";
    case Status::NotDescentDirection: return "Search direction not a descent direction";
    case Status::LineSearchFailed: return "Line search failed";
    case Status::UpdateDirectionFailed: return "Update direction could not be computed";
    }
    return "Unknown status";
}

bool is_converged_status(Status s)
{
    return s == Status::XDeltaTolerance || s == Status::FDeltaTolerance || s == Status::GradNormTolerance ||
           s == Status::RelGradNormTolerance || s == Status::RelXDeltaTolerance || s == Status::NewtonDecrementTolerance;
}

struct Criteria
{
    long iterations = 0;
    double xDelta = 0, fDelta = 0, gradNorm = 0, firstGradNorm = 0, xDeltaDotGrad = 0, relGradNorm = 0, relXDelta = 0,
           newtonDecrement = 0;
    int fDeltaCount = 0;
    double energy = 0, alpha = 0, step = 0;
};

psb200_nl_criteria as_c(const Criteria &c)
{
    return {c.iterations, c.xDelta, c.fDelta, c.gradNorm, c.firstGradNorm, c.xDeltaDotGrad, c.relGradNorm, c.relXDelta, c.newtonDecrement,
            c.fDeltaCount, c.energy, c.alpha, c.step};
}

// Criteria.cpp:59-96
Status check_convergence(const Criteria &stop, const Criteria &cur)
{
    if (stop.iterations > 0 && cur.iterations > stop.iterations)
        return Status::IterationLimit;
    const double stop_grad = cur.iterations == 0 ? stop.firstGradNorm : stop.gradNorm;
    if (stop_grad > 0 && cur.gradNorm < stop_grad)
        return Status::GradNormTolerance;
    if (stop.relXDelta > 0 && cur.relXDelta < stop.relXDelta)
        return Status::RelXDeltaTolerance;
    if (stop.relGradNorm > 0 && cur.relGradNorm < stop.relGradNorm)
        return Status::RelGradNormTolerance;
    if (stop.newtonDecrement > 0 && cur.newtonDecrement < stop.newtonDecrement)
        return Status::NewtonDecrementTolerance;
    if (stop.xDelta > 0 && cur.xDelta < stop.xDelta)
        return Status::XDeltaTolerance;
    if (stop.fDelta > 0 && cur.fDelta < stop.fDelta && cur.fDeltaCount >= stop.fDeltaCount)
        return Status::FDeltaTolerance;
    if (stop.xDeltaDotGrad < 0 && cur.xDeltaDotGrad > stop.xDeltaDotGrad)
        return Status::NotDescentDirection;
    return Status::Continue;
}

double norm2(const std::vector<double> &v)
{
    double s = 0;
    for (double a : v)
        s += a * a;
    return std::sqrt(s);
}
double dot(const std::vector<double> &a, const std::vector<double> &b)
{
    double s = 0;
    for (size_t i = 0; i < a.size(); ++i)
        s += a[i] * b[i];
    return s;
}
bool all_finite(const std::vector<double> &v)
{
    for (double a : v)
        if (!std::isfinite(a))
            return false;
    return true;
}

// JSON lookups with the defaults of nonlinear-solver-spec.json (jse::inject_defaults in the reference)
double jget(const JValue &j, const char *k, double def) { return j.contains(k) ? j.at(k).as_num() : def; }
bool jgetb(const JValue &j, const char *k, bool def) { return j.contains(k) ? j.at(k).as_bool() : def; }
std::string jgets(const JValue &j, const char *k, const std::string &def) { return j.contains(k) ? j.at(k).as_str() : def; }
JValue empty_obj()
{
    JValue v;
    v.kind = JValue::Obj;
    return v;
}
JValue jsub(const JValue &j, const char *k) { return j.contains(k) && j.at(k).is_obj() ? j.at(k) : empty_obj(); }

struct Problem
{
    const psb200_nl_problem *p;
    int64_t n;
    double value(const std::vector<double> &x) const { return p->value(p->user, x.data(), n); }
    void gradient(const std::vector<double> &x, std::vector<double> &g) const
    {
        g.resize((size_t)n);
        p->gradient(p->user, x.data(), n, g.data());
    }
    void solution_changed(const std::vector<double> &x) const
    {
        if (p->solution_changed)
            p->solution_changed(p->user, x.data(), n);
    }
    bool is_step_valid(const std::vector<double> &x0, const std::vector<double> &x1) const
    {
        return p->is_step_valid ? p->is_step_valid(p->user, x0.data(), x1.data(), n) != 0 : true;
    }
    double max_step_size(const std::vector<double> &x0, const std::vector<double> &x1) const
    {
        return p->max_step_size ? p->max_step_size(p->user, x0.data(), x1.data(), n) : 1.0;
    }
    void line_search_begin(const std::vector<double> &x0, const std::vector<double> &x1) const
    {
        if (p->line_search_begin)
            p->line_search_begin(p->user, x0.data(), x1.data(), n);
    }
    void line_search_end() const
    {
        if (p->line_search_end)
            p->line_search_end(p->user);
    }
    void post_step(int it, const std::vector<double> &x, const std::vector<double> &g) const
    {
        if (p->post_step)
            p->post_step(p->user, it, x.data(), g.data(), n);
    }
    bool stop(const std::vector<double> &x) const { return p->stop ? p->stop(p->user, x.data(), n) != 0 : false; }
    // Problem.hpp:35,103,109,116-121 -- optional, with the reference's defaults
    int norm_type = 1; // "norm_type": Euclidean 0, L2 1 (spec default), Linf 2
    bool is_residual() const { return p->is_residual ? p->is_residual(p->user) != 0 : false; }
    bool after_line_search_custom_operation(const std::vector<double> &x0, const std::vector<double> &x1) const
    {
        return p->after_line_search_custom_operation ? p->after_line_search_custom_operation(p->user, x0.data(), x1.data(), n) != 0 : false;
    }
    bool callback(const psb200_nl_criteria &state, const std::vector<double> &x) const
    {
        return p->callback ? p->callback(p->user, &state, x.data(), n) != 0 : true;
    }
    double grad_norm(const std::vector<double> &g) const { return p->grad_norm ? p->grad_norm(p->user, g.data(), n, norm_type) : norm2(g); }
    double step_norm(const std::vector<double> &d) const { return p->step_norm ? p->step_norm(p->user, d.data(), n, norm_type) : norm2(d); }
    double rescaling(int which) const { return p->norm_rescaling ? p->norm_rescaling(p->user, which, norm_type) : 1.0; }
    // 1/2 x^T H x with the host Hessian (Solver.cpp:409-423); NaN when the assembly fails
    double half_xHx(const std::vector<double> &x) const
    {
        int64_t nnz = 0;
        const int32_t *outer = nullptr, *inner = nullptr;
        const double *vals = nullptr;
        if (!p->hessian || p->hessian(p->user, x.data(), n, 0, &nnz, &outer, &inner, &vals) != 0 || !outer || (nnz > 0 && (!inner || !vals)))
            return NaN;
        double s = 0;
        for (int64_t c = 0; c < n; ++c)
            for (int32_t k = outer[c]; k < outer[c + 1]; ++k)
                s += x[(size_t)inner[k]] * vals[k] * x[(size_t)c];
        return 0.5 * s;
    }
};

void axpy(std::vector<double> &out, const std::vector<double> &x, double a, const std::vector<double> &d)
{
    out.resize(x.size());
    for (size_t i = 0; i < x.size(); ++i)
        out[i] = x[i] + a * d[i];
}

// ------------------------------------------------------------------------------------------ line search
// LineSearch.cpp:60-254 (template method) with the Backtracking / Armijo criteria.
struct LineSearch
{
    std::string method = "RobustArmijo";
    double min_step_size = 1e-10, min_step_size_final = 1e-20, default_init_step_size = 1, step_ratio = 0.5;
    int max_step_size_iter = 30, max_step_size_iter_final = 100;
    double use_grad_norm_tol = 1e-6, armijo_c = 1e-4;
    double delta_relative_tolerance = 0.1; // RobustArmijo (nonlinear-solver-spec.json:683-688)
    bool is_final_strategy = false;
    int cur_iter = 0;
    long total_iterations = 0;
    double armijo_criteria = 0;

    double cur_min_step() const { return is_final_strategy ? min_step_size_final : min_step_size; }
    int cur_max_iter() const { return is_final_strategy ? max_step_size_iter_final : max_step_size_iter; }

    // LineSearch.cpp:189-223
    double nan_free_step(const std::vector<double> &x, const std::vector<double> &dx, const Problem &f, double step) const
    {
        std::vector<double> nx;
        axpy(nx, x, step, dx);
        int &it = const_cast<int &>(cur_iter);
        while (step > cur_min_step() && it < cur_max_iter())
        {
            if (!f.is_step_valid(x, nx) || !std::isfinite(f.value(nx)))
            {
                step *= step_ratio;
                axpy(nx, x, step, dx);
            }
            else
                break;
            ++it;
        }
        if (it >= cur_max_iter() || step <= cur_min_step())
            return NaN;
        return step;
    }

    // Backtracking.cpp:15-64 with criteria() of Backtracking.cpp:66-83 / Armijo.cpp:20-32
    double descent_step(const std::vector<double> &x, const std::vector<double> &dx, const Problem &f, bool use_grad_norm,
                        double old_energy, const std::vector<double> &old_grad, double step)
    {
        const bool armijo = method == "Armijo" || method == "RobustArmijo";
        if (armijo)
            armijo_criteria = armijo_c * dot(dx, old_grad); // Armijo.cpp:13-18
        std::vector<double> nx, ng;
        for (; step > cur_min_step() && cur_iter < cur_max_iter(); step *= step_ratio, ++cur_iter)
        {
            axpy(nx, x, step, dx);
            f.solution_changed(nx);
            if (!f.is_step_valid(x, nx))
                continue;
            const double e = f.value(nx);
            if (!std::isfinite(e))
                continue;
            bool ok;
            if (armijo)
            {
                ok = e <= old_energy + step * armijo_criteria; // Armijo.cpp:20-32 (RobustArmijo tries it first)
                if (!ok && method == "RobustArmijo" && std::fabs(e - old_energy) <= delta_relative_tolerance * std::fabs(old_energy))
                {
                    // RobustArmijo.cpp:30-44 (Longva et al. 2023): when the energy difference drowns in rounding error,
                    // estimate it from the gradients at both ends (trapezoid rule) plus a bound on the estimate's error
                    f.gradient(nx, ng);
                    double dsum = 0, ddif = 0;
                    for (size_t i = 0; i < dx.size(); ++i)
                    {
                        dsum += dx[i] * (ng[i] + old_grad[i]);
                        ddif += dx[i] * (ng[i] - old_grad[i]);
                    }
                    const double deltaE_approx = step / 2 * dsum;
                    const double abs_eps_est = step / 2 * std::fabs(ddif);
                    ok = deltaE_approx + abs_eps_est <= step * armijo_criteria;
                }
            }
            else if (use_grad_norm || method == "ResidualBacktracking") // ResidualBacktracking.cpp:15-28: always the gradient norm
            {
                f.gradient(nx, ng);
                ok = f.grad_norm(ng) < f.grad_norm(old_grad); // Backtracking.cpp:76-80
            }
            else
                ok = e < old_energy;
            if (ok)
                break;
        }
        return step;
    }

    // LineSearch.cpp:73-187
    double line_search(const std::vector<double> &x, const std::vector<double> &dx, const Problem &f)
    {
        cur_iter = 0;
        const double e0 = f.value(x);
        if (std::isnan(e0))
            return NaN;
        std::vector<double> g0;
        f.gradient(x, g0);
        if (!all_finite(g0))
            return NaN;
        double step = default_init_step_size;
        step = nan_free_step(x, dx, f, step);
        if (std::isnan(step))
            return NaN;
        std::vector<double> nx;
        axpy(nx, x, step, dx);
        f.line_search_begin(x, nx);
        {
            // LineSearch.cpp:226-254 compute_max_step_size: rounds down when scaling by the CCD step
            const double mx = f.max_step_size(x, nx);
            if (mx == 0)
            {
                f.line_search_end();
                return NaN;
            }
            const int rnd = std::fegetround();
            std::fesetround(FE_DOWNWARD);
            step *= mx;
            std::fesetround(rnd);
        }
        const double gn = f.grad_norm(g0);
        if (gn < 1e-30)
        {
            total_iterations += cur_iter;
            return step;
        }
        const bool use_grad_norm = gn < use_grad_norm_tol * f.rescaling(0); // LineSearch.cpp:142
        if (method == "None")
        {
            // NoLineSearch.cpp:11-22: the starting step, announced through solution_changed; the checks below still apply
            axpy(nx, x, step, dx);
            f.solution_changed(nx);
        }
        else
            step = descent_step(x, dx, f, use_grad_norm, e0, g0, step);
        total_iterations += cur_iter;
        if (std::isnan(step))
            return NaN;
        if (cur_iter >= cur_max_iter() || step <= cur_min_step())
        {
            f.solution_changed(x);
            f.line_search_end();
            return NaN;
        }
        f.line_search_end();
        return step;
    }
};

// ------------------------------------------------------------------------------------------ strategies
struct Strategy
{
    virtual ~Strategy() {}
    virtual std::string name() const = 0;
    virtual void reset() {}
    virtual bool is_direction_descent() { return true; }
    virtual bool handle_error() { return false; }
    virtual bool direction(const Problem &f, const std::vector<double> &x, const std::vector<double> &g, std::vector<double> &dx,
                           std::string &err) = 0;
    virtual void info(std::ostringstream &) const {}
};

// Newton.cpp:144-214 (sparse branch). kind: 0 Newton, 1 ProjectedNewton, 2 RegularizedNewton.
struct NewtonStrategy : Strategy
{
    int kind;
    bool project_to_psd;
    double residual_tolerance;
    double reg_weight_min = 0, reg_weight_max = 0, reg_weight_inc = 0, reg_weight = 0;
    psb200_handle lin = nullptr;
    std::vector<std::string> internal_info;
    double assembly_time = 0, inverting_time = 0;
    // regularised copy of the Hessian (diagonal entries inserted when the pattern lacks them)
    std::vector<int32_t> r_outer, r_inner;
    std::vector<double> r_vals;
    std::vector<double> hd, resid;

    NewtonStrategy(int kind_, bool psd, double res_tol, const std::string &lin_json) : kind(kind_), project_to_psd(psd), residual_tolerance(res_tol)
    {
        // Newton.cpp:70: linear_solver = linear::Solver::create(linear_solver_params, logger)
        if (psb200_create(&lin, lin_json.empty() ? nullptr : lin_json.c_str()) != PSB200_OK)
            throw std::runtime_error(std::string("psb200_create: ") + psb200_last_error(nullptr));
        if (!(res_tol > 0))
            throw std::runtime_error("Newton residual_tolerance must be > 0");
    }
    ~NewtonStrategy() override
    {
        if (lin)
            psb200_destroy(lin);
    }
    std::string name() const override { return kind == 0 ? "Newton" : kind == 1 ? "ProjectedNewton" : "RegularizedNewton"; }
    void reset() override
    {
        internal_info.clear();
        reg_weight = reg_weight_min;
        assembly_time = inverting_time = 0;
    }
    bool handle_error() override
    {
        if (kind != 2)
            return false;
        reg_weight *= reg_weight_inc; // Newton.cpp:326-330
        return reg_weight < reg_weight_max;
    }

    // hessian += w I (Newton.cpp:287-290; Eigen's sparse sum creates missing diagonal entries)
    void regularize(int64_t n, const int32_t *&outer, const int32_t *&inner, const double *&vals, int64_t &nnz)
    {
        r_outer.assign((size_t)n + 1, 0);
        r_inner.clear();
        r_vals.clear();
        r_inner.reserve((size_t)nnz + n);
        r_vals.reserve((size_t)nnz + n);
        for (int64_t c = 0; c < n; ++c)
        {
            bool placed = false;
            for (int32_t k = outer[c]; k < outer[c + 1]; ++k)
            {
                const int32_t r = inner[k];
                if (!placed && r >= c)
                {
                    if (r == c)
                    {
                        r_inner.push_back(r);
                        r_vals.push_back(vals[k] + reg_weight);
                        placed = true;
                        continue;
                    }
                    r_inner.push_back((int32_t)c);
                    r_vals.push_back(reg_weight);
                    placed = true;
                }
                r_inner.push_back(r);
                r_vals.push_back(vals[k]);
            }
            if (!placed)
            {
                r_inner.push_back((int32_t)c);
                r_vals.push_back(reg_weight);
            }
            r_outer[c + 1] = (int32_t)r_inner.size();
        }
        outer = r_outer.data();
        inner = r_inner.data();
        vals = r_vals.data();
        nnz = (int64_t)r_inner.size();
    }

    bool direction(const Problem &f, const std::vector<double> &x, const std::vector<double> &g, std::vector<double> &dx,
                   std::string &err) override
    {
        const int64_t n = f.n;
        int64_t nnz = 0;
        const int32_t *outer = nullptr, *inner = nullptr;
        const double *vals = nullptr;
        const double *d_vals = nullptr; // device-resident values (Problem::hessian_device)
        {
            ScopedTimer t(assembly_time);
            if (f.p->hessian_device)
            {
                if (f.p->hessian_device(f.p->user, x.data(), n, project_to_psd ? 1 : 0, &nnz, &outer, &inner, &d_vals) != 0)
                    return false;
            }
            else
            {
                if (f.p->hessian(f.p->user, x.data(), n, project_to_psd ? 1 : 0, &nnz, &outer, &inner, &vals) != 0)
                    return false;
                if (kind == 2 && reg_weight > 0)
                    regularize(n, outer, inner, vals, nnz);
            }
        }
        std::vector<double> rhs(g.size());
        for (size_t i = 0; i < g.size(); ++i)
            rhs[i] = -g[i];
        {
            ScopedTimer t(inverting_time);
            // Newton.cpp:189: analyze_pattern every iteration (cached by hash on the solver side)
            // A missing device / CUDA failure is never "recoverable": falling through to GradientDescent would be a
            // silent CPU fallback. Only numeric failures (PSB200_ERR_NUMERIC) take the reference's fallback route.
            int rc = psb200_analyze_pattern_csc(lin, n, nnz, outer, inner, (int)n);
            if (rc == PSB200_ERR_CUDA || rc == PSB200_ERR_COMM || rc == PSB200_ERR_INVALID)
                throw std::runtime_error(std::string("analyze_pattern: ") + psb200_last_error(lin));
            if (rc != PSB200_OK)
                return false;
            // Newton.cpp:191-202: a failing factorize is recoverable (NaN residual -> next strategy)
            if (d_vals) // RegularizedNewton: hessian += reg_weight I on the device (Newton.cpp:287-290)
                rc = psb200_factorize_csc_device(lin, n, nnz, d_vals, kind == 2 ? reg_weight : 0.0);
            else
                rc = psb200_factorize_csc(lin, n, nnz, outer, inner, vals);
            if (rc == PSB200_ERR_CUDA || rc == PSB200_ERR_COMM || rc == PSB200_ERR_INVALID)
                throw std::runtime_error(std::string("factorize: ") + psb200_last_error(lin));
            if (rc != PSB200_OK)
                return false;
            // Newton.cpp:204: solve(-grad, direction); direction carries the previous step in as the initial guess
            dx.resize(g.size());
            if (psb200_solve(lin, rhs.data(), dx.data(), n) != PSB200_OK)
            {
                err = psb200_last_error(lin); // an exception from solve() propagates in the reference (Newton.cpp:204)
                throw std::runtime_error(err);
            }
            // row-partitioned solver: every rank needs the whole step (no-op on one GPU)
            if (psb200_dist_allgather(lin, dx.data(), n) != PSB200_OK)
                throw std::runtime_error(std::string("dist_allgather: ") + psb200_last_error(lin));
        }
        // Newton.cpp:207: residual = ||H dx + g|| = ||H dx - rhs||, through the product SpMV kernel (global norm on a partition)
        double residual = 0;
        if (psb200_residual_norm(lin, dx.data(), rhs.data(), n, &residual) != PSB200_OK)
            throw std::runtime_error(psb200_last_error(lin));
        // Newton.cpp:209-211
        std::vector<char> buf(1 << 15);
        size_t need = 0;
        if (psb200_get_info(lin, buf.data(), buf.size(), &need) != PSB200_OK && need > buf.size())
        {
            buf.resize(need);
            psb200_get_info(lin, buf.data(), buf.size(), &need);
        }
        internal_info.emplace_back(buf.data());
        // Newton.cpp:156-162
        if (std::isnan(residual) || residual > residual_tolerance)
            return false;
        return true;
    }
    void info(std::ostringstream &o) const override
    {
        o << ",\"time_assembly\":" << jnum(assembly_time) << ",\"time_inverting\":" << jnum(inverting_time);
    }
};

// LBFGS.cpp:22-61: the memory lives on the GPU (lbfgs.cu), created on first use for the problem size at hand
struct LbfgsStrategy : Strategy
{
    int history = 6;
    int device = -1;
    psb200_lbfgs_handle h = nullptr;
    long long n_alloc = -1;
    explicit LbfgsStrategy(int history_size, int dev) : history(history_size), device(dev) {}
    ~LbfgsStrategy() override
    {
        if (h)
            psb200_lbfgs_destroy(h);
    }
    std::string name() const override { return "L-BFGS"; }
    void reset() override
    {
        if (h)
            psb200_lbfgs_reset(h);
    }
    bool direction(const Problem &, const std::vector<double> &x, const std::vector<double> &g, std::vector<double> &dx, std::string &err) override
    {
        if (!h || n_alloc != (long long)x.size())
        {
            if (h)
                psb200_lbfgs_destroy(h);
            h = nullptr;
            if (psb200_lbfgs_create(&h, (int64_t)x.size(), history, device) != PSB200_OK)
                throw std::runtime_error(std::string("L-BFGS: ") + psb200_lbfgs_last_error(nullptr)); // no device: never "recoverable"
            n_alloc = (long long)x.size();
        }
        dx.resize(x.size());
        if (psb200_lbfgs_direction(h, x.data(), g.data(), dx.data(), (int64_t)x.size()) != PSB200_OK)
            throw std::runtime_error(std::string("L-BFGS: ") + psb200_lbfgs_last_error(h));
        (void)err;
        return true;
    }
    void info(std::ostringstream &o) const override { o << ",\"history_size\":" << history; }
};

struct GradientDescent : Strategy
{
    std::string name() const override { return "GradientDescent"; }
    bool direction(const Problem &, const std::vector<double> &, const std::vector<double> &g, std::vector<double> &dx, std::string &) override
    {
        dx.resize(g.size());
        for (size_t i = 0; i < g.size(); ++i)
            dx[i] = -g[i];
        return true;
    }
};

} // namespace

struct psb200_nl_solver
{
    std::string err;
    std::string info_json = "{}";
    Criteria stop;
    bool allow_out_of_iterations = false, allow_non_grad_convergence = false;
    std::vector<std::unique_ptr<Strategy>> strategies;
    std::vector<int> iter_per_strategy;
    LineSearch ls;
    std::string solver_name = "Newton";
    int norm_type = 1;
    Status status = Status::NotStarted;
    // Solver::set_iteration_callback / set_direction_filter (Solver.hpp:76-86)
    int (*iteration_callback)(void *, const psb200_nl_criteria *) = nullptr;
    void *iteration_user = nullptr;
    void (*direction_filter)(void *, const double *, double *, int64_t) = nullptr;
    void *filter_user = nullptr;
};

namespace {

void build(psb200_nl_solver &S, const std::string &solver_json, const std::string &linear_json)
{
    const JValue j = solver_json.empty() ? empty_obj() : psb::JParser::parse(solver_json);
    const JValue adv = jsub(j, "advanced");
    // Solver.cpp:200-230 with the defaults of nonlinear-solver-spec.json
    S.stop.xDelta = jget(j, "x_delta_tol", 0);
    S.stop.fDelta = jget(adv, "f_delta_tol", 0);
    S.stop.gradNorm = jget(j, "grad_norm_tol", 1e-10);
    S.stop.firstGradNorm = jget(j, "first_grad_norm_tol", 1e-12);
    S.stop.xDeltaDotGrad = -jget(adv, "derivative_along_delta_x_tol", 0);
    S.stop.relGradNorm = jget(j, "rel_grad_norm_tol", 1e-10);
    S.stop.relXDelta = jget(j, "rel_x_delta_tol", 0);
    S.stop.newtonDecrement = jget(j, "newton_decrement_tol", 0);
    S.stop.iterations = (long)jget(j, "max_iterations", 500);
    S.stop.fDeltaCount = (int)jget(adv, "f_delta_step_tol", 100);
    S.allow_out_of_iterations = jgetb(j, "allow_out_of_iterations", false);
    S.allow_non_grad_convergence = jgetb(j, "allow_non_grad_convergence", false);
    {
        // Solver.cpp:117-121,224; spec default "L2" (nonlinear-solver-spec.json:91-99)
        const std::string nt = jgets(j, "norm_type", "L2");
        if (nt == "Euclidean")
            S.norm_type = 0;
        else if (nt == "L2")
            S.norm_type = 1;
        else if (nt == "Linf")
            S.norm_type = 2;
        else
            throw std::runtime_error("Unknown norm_type " + nt + " (Euclidean, L2, Linf)");
    }

    const JValue lsj = jsub(j, "line_search");
    S.ls.method = jgets(lsj, "method", "RobustArmijo");
    if (S.ls.method != "Backtracking" && S.ls.method != "Armijo" && S.ls.method != "RobustArmijo" && S.ls.method != "ResidualBacktracking" &&
        S.ls.method != "None") // LineSearch.cpp:24-52
        throw std::runtime_error("Unknown line search " + S.ls.method + "!");
    S.ls.use_grad_norm_tol = jget(lsj, "use_grad_norm_tol", 1e-6);
    S.ls.min_step_size = jget(lsj, "min_step_size", 1e-10);
    S.ls.max_step_size_iter = (int)jget(lsj, "max_step_size_iter", 30);
    S.ls.min_step_size_final = jget(lsj, "min_step_size_final", 1e-20);
    S.ls.max_step_size_iter_final = (int)jget(lsj, "max_step_size_iter_final", 100);
    S.ls.default_init_step_size = jget(lsj, "default_init_step_size", 1);
    S.ls.step_ratio = jget(lsj, "step_ratio", 0.5);
    S.ls.armijo_c = jget(jsub(lsj, "Armijo"), "c", 1e-4);
    S.ls.delta_relative_tolerance = jget(jsub(lsj, "RobustArmijo"), "delta_relative_tolerance", 0.1);

    if (j.contains("solver") && j.at("solver").kind == JValue::Arr)
    {
        // Solver.cpp:147-154: "solver": [{"type": "Newton", ...}, {"type": "L-BFGS"}, ...] -- the strategies in the order
        // given, each built from its own entry (create_solver, Solver.cpp:34-106), and NO automatic GradientDescent fallback.
        // Parameters follow extract_param (Utils.cpp:78-84): entry[Type][name] when the entry has a Type object, else
        // entry[name], else the default of /solver/*/name (nonlinear-solver-spec.json).
        S.solver_name = "list";
        int device = -1;
        if (!linear_json.empty())
        {
            const JValue lj = psb::JParser::parse(linear_json);
            if (lj.is_obj() && lj.contains("CUDA") && lj.at("CUDA").is_obj() && lj.at("CUDA").contains("device"))
                device = (int)lj.at("CUDA").at("device").as_num();
        }
        for (const JValue &e : j.at("solver").arr)
        {
            if (!e.is_obj() || !e.contains("type"))
                throw std::runtime_error("every entry of the \"solver\" list needs a \"type\"");
            const std::string t = e.at("type").as_str();
            auto ex = [&](const char *key, const char *name, double def) {
                if (e.contains(key) && e.at(key).is_obj() && e.at(key).contains(name))
                    return e.at(key).at(name).as_num();
                return jget(e, name, def);
            };
            if (t == "Newton" || t == "SparseNewton" || t == "sparse_newton")
                S.strategies.push_back(std::make_unique<NewtonStrategy>(0, false, ex("Newton", "residual_tolerance", 1e-5), linear_json));
            else if (t == "ProjectedNewton")
                S.strategies.push_back(std::make_unique<NewtonStrategy>(1, true, ex("ProjectedNewton", "residual_tolerance", 1e-5), linear_json));
            else if (t == "RegularizedNewton" || t == "RegularizedProjectedNewton")
            {
                const char *k = "RegularizedNewton";
                const double wmin = ex(k, "reg_weight_min", 1e-8), wmax = ex(k, "reg_weight_max", 1e8), winc = ex(k, "reg_weight_inc", 10);
                if (wmin <= 0) // Newton.cpp:117-118
                    throw std::runtime_error("Newton reg_weight_min must be  > 0");
                if (winc <= 1)
                    throw std::runtime_error("Newton reg_weight_inc must be > 1");
                if (wmax <= wmin)
                    throw std::runtime_error("Newton reg_weight_max must be > reg_weight_min");
                auto r = std::make_unique<NewtonStrategy>(2, t == "RegularizedProjectedNewton", ex(k, "residual_tolerance", 1e-5), linear_json);
                r->reg_weight_min = wmin;
                r->reg_weight_max = wmax;
                r->reg_weight_inc = winc;
                r->reg_weight = wmin;
                S.strategies.push_back(std::move(r));
            }
            else if (t == "LBFGS" || t == "L-BFGS")
            {
                const int history = (int)ex("L-BFGS", "history_size", 6);
                if (history <= 0)
                    throw std::runtime_error("L-BFGS history_size must be >=1, instead got " + std::to_string(history));
                S.strategies.push_back(std::make_unique<LbfgsStrategy>(history, device));
            }
            else if (t == "GradientDescent" || t == "gradient_descent")
                S.strategies.push_back(std::make_unique<GradientDescent>());
            else
                throw std::runtime_error("Unrecognized solver type: " + t +
                                         " (this driver provides Newton, ProjectedNewton, Regularized[Projected]Newton, L-BFGS and GradientDescent)");
        }
        if (S.strategies.empty())
            throw std::runtime_error("the \"solver\" list is empty");
    }
    else if ((S.solver_name = jgets(j, "solver", "Newton")) == "Newton" || S.solver_name == "SparseNewton" || S.solver_name == "sparse_newton")
    {
        // Newton.cpp:14-58
        const JValue nj = jsub(j, "Newton");
        const double res_tol = jget(nj, "residual_tolerance", 1e-5);
        const double wmin = jget(nj, "reg_weight_min", 1e-8), wmax = jget(nj, "reg_weight_max", 1e8), winc = jget(nj, "reg_weight_inc", 10);
        if (!jgetb(nj, "force_psd_projection", false))
            S.strategies.push_back(std::make_unique<NewtonStrategy>(0, false, res_tol, linear_json));
        if (jgetb(nj, "use_psd_projection", true))
            S.strategies.push_back(std::make_unique<NewtonStrategy>(1, true, res_tol, linear_json));
        if (wmin > 0)
        {
            if (winc <= 1)
                throw std::runtime_error("Newton reg_weight_inc must be > 1");
            if (wmax <= wmin)
                throw std::runtime_error("Newton reg_weight_max must be > reg_weight_min");
            auto r = std::make_unique<NewtonStrategy>(2, jgetb(nj, "use_psd_projection_in_regularized", true), res_tol, linear_json);
            r->reg_weight_min = wmin;
            r->reg_weight_max = wmax;
            r->reg_weight_inc = winc;
            r->reg_weight = wmin;
            S.strategies.push_back(std::move(r));
        }
        if (S.strategies.empty())
            throw std::runtime_error("Newton needs to have at least one of force_psd_projection=false, reg_weight_min>0, or use_psd_projection=true");
        S.strategies.push_back(std::make_unique<GradientDescent>()); // Solver.cpp:175-181
    }
    else if (S.solver_name == "LBFGS" || S.solver_name == "L-BFGS")
    {
        // Solver.cpp:83-85 + the GradientDescent fallback every non-GD solver gets (:175-181)
        const int history = (int)jget(jsub(j, "L-BFGS"), "history_size", 6);
        if (history <= 0)
            throw std::runtime_error("L-BFGS history_size must be >=1, instead got " + std::to_string(history)); // LBFGS.cpp:17-18
        int device = -1;
        if (!linear_json.empty())
        {
            const JValue lj = psb::JParser::parse(linear_json);
            if (lj.is_obj() && lj.contains("CUDA") && lj.at("CUDA").is_obj() && lj.at("CUDA").contains("device"))
                device = (int)lj.at("CUDA").at("device").as_num();
        }
        S.strategies.push_back(std::make_unique<LbfgsStrategy>(history, device));
        S.strategies.push_back(std::make_unique<GradientDescent>());
    }
    else if (S.solver_name == "GradientDescent" || S.solver_name == "gradient_descent")
        S.strategies.push_back(std::make_unique<GradientDescent>());
    else
        throw std::runtime_error("Unrecognized solver type: " + S.solver_name + " (this driver provides Newton, L-BFGS and GradientDescent)");
    // Solver.cpp:232-245
    if (j.contains("iterations_per_strategy") && j.at("iterations_per_strategy").kind == JValue::Arr)
    {
        const auto &a = j.at("iterations_per_strategy").arr;
        if (a.size() != S.strategies.size() + 1)
            throw std::runtime_error("Invalit iter_per_strategy size: " + std::to_string(a.size()) + "!=" + std::to_string(S.strategies.size() + 1));
        S.iter_per_strategy.clear();
        for (const JValue &v : a)
            S.iter_per_strategy.push_back((int)v.as_num());
    }
    else
        S.iter_per_strategy.assign(S.strategies.size() + 1, (int)jget(j, "iterations_per_strategy", 5));
}

// Solver.cpp:255-582. Returns false where the reference throws; err holds the message.
bool minimize(psb200_nl_solver &S, const Problem &f, std::vector<double> &x)
{
    // Solver.hpp:118-131 reset_stopping_criteria: absolute tolerances scaled by the Problem's rescalings (default 1)
    Criteria stop = S.stop;
    stop.xDelta *= f.rescaling(1);
    stop.fDelta *= f.rescaling(2);
    stop.gradNorm *= f.rescaling(0);
    stop.firstGradNorm *= f.rescaling(0);
    stop.xDeltaDotGrad *= f.rescaling(2);
    stop.newtonDecrement *= f.rescaling(2);
    Criteria cur;
    size_t strategy = 0, previous_strategy = 0;
    int current_strategy_iter = 0;
    for (auto &s : S.strategies)
        s->reset();
    S.ls.total_iterations = 0;
    S.status = Status::NotStarted;
    std::vector<double> grad(x.size(), 0.0), dx(x.size(), 0.0), x1;
    double old_energy = NaN, total_time = 0, obj_time = 0, grad_time = 0, dir_time = 0, ls_time = 0;
    double initial_grad_norm = NaN, initial_dx_norm = NaN;
    const double t_start = now_s();
    f.solution_changed(x);
    f.post_step(0, x, grad);
    bool ok = true;
    auto name = [&]() { return S.strategies[std::min(strategy, S.strategies.size() - 1)]->name(); };
    auto fail = [&](Status st, const std::string &msg) {
        S.status = st;
        S.err = "[" + name() + "][" + S.ls.method + "] " + msg;
        ok = false;
    };
    do
    {
        S.ls.is_final_strategy = strategy == S.strategies.size() - 1;
        double energy;
        {
            ScopedTimer t(obj_time);
            energy = f.value(x);
        }
        cur.energy = energy;
        if (!std::isfinite(energy))
        {
            fail(Status::NanEncountered, "f(x) is nan or inf; stopping");
            break;
        }
        cur.fDelta = std::abs(old_energy - energy);
        {
            ScopedTimer t(grad_time);
            f.gradient(x, grad);
        }
        cur.gradNorm = f.grad_norm(grad);
        if (cur.iterations == 0)
        {
            initial_grad_norm = cur.gradNorm;
            cur.relGradNorm = NaN;
        }
        else
            cur.relGradNorm = cur.gradNorm / initial_grad_norm;
        if (std::isnan(cur.gradNorm))
        {
            fail(Status::NanEncountered, "Gradient is nan; stopping");
            break;
        }
        cur.xDelta = cur.xDeltaDotGrad = cur.relXDelta = cur.newtonDecrement = NaN;
        S.status = check_convergence(stop, cur);
        if (S.status != Status::Continue)
            break;

        bool dir_ok;
        std::string derr;
        {
            ScopedTimer t(dir_time);
            try
            {
                dir_ok = S.strategies[strategy]->direction(f, x, grad, dx, derr);
            }
            catch (const std::exception &e)
            {
                fail(Status::UpdateDirectionFailed, std::string("linear solve failed: ") + e.what());
                break;
            }
        }
        if (S.direction_filter && dir_ok) // Solver.cpp:353-358
            S.direction_filter(S.filter_user, x.data(), dx.data(), f.n);
        cur.xDelta = f.step_norm(dx);
        if (cur.iterations == 0)
        {
            initial_dx_norm = cur.xDelta;
            cur.relXDelta = NaN;
        }
        else
            cur.relXDelta = cur.xDelta / initial_dx_norm;
        if (!dir_ok || std::isnan(cur.xDelta))
        {
            if (!S.strategies[strategy]->handle_error())
                ++strategy;
            if (strategy >= S.strategies.size())
            {
                fail(Status::UpdateDirectionFailed, std::string(status_message(Status::UpdateDirectionFailed)) + " on last strategy; stopping");
                break;
            }
            S.status = Status::Continue;
            continue;
        }
        if (S.direction_filter)
        {
            // Solver.cpp:392-403: with a filter, descent is measured against the filtered steepest-descent direction
            std::vector<double> neg_grad(grad.size());
            for (size_t i = 0; i < grad.size(); ++i)
                neg_grad[i] = -grad[i];
            S.direction_filter(S.filter_user, x.data(), neg_grad.data(), f.n);
            cur.xDeltaDotGrad = -dot(dx, neg_grad);
        }
        else
            cur.xDeltaDotGrad = dot(dx, grad);
        if (stop.newtonDecrement > 0)
            cur.newtonDecrement = f.half_xHx(x); // Solver.cpp:409-423 (x^T H x / 2, NaN when the Hessian cannot be assembled)
        if (!f.is_residual() && S.strategies[strategy]->is_direction_descent() && cur.gradNorm != 0 && cur.xDeltaDotGrad >= 0)
        {
            if (!S.strategies[strategy]->handle_error())
                ++strategy;
            if (strategy >= S.strategies.size())
            {
                fail(Status::NotDescentDirection, std::string(status_message(Status::NotDescentDirection)) + " on last strategy; stopping");
                break;
            }
            S.status = Status::Continue;
            continue;
        }
        S.status = check_convergence(stop, cur);
        if (S.status != Status::Continue)
            break;

        double rate;
        {
            ScopedTimer t(ls_time);
            rate = S.ls.line_search(x, dx, f);
        }
        cur.alpha = rate;
        if (std::isnan(rate))
        {
            if (!S.strategies[strategy]->handle_error())
                ++strategy;
            if (strategy >= S.strategies.size())
            {
                fail(Status::LineSearchFailed, "Line search failed on last strategy; stopping");
                break;
            }
            continue;
        }
        axpy(x1, x, rate, dx);
        if (f.after_line_search_custom_operation(x, x1)) // Solver.cpp:495-499
            f.solution_changed(x1);
        x = x1;
        old_energy = energy;
        if (strategy != previous_strategy)
            current_strategy_iter = 0;
        if (strategy != 0 && current_strategy_iter >= S.iter_per_strategy[strategy])
        {
            strategy = 0;
            for (auto &s : S.strategies)
                s->reset();
        }
        previous_strategy = strategy;
        ++current_strategy_iter;
        cur.step = std::abs(rate) * norm2(dx); // Solver.cpp:528: (rate * delta_x).norm()
        f.post_step((int)cur.iterations, x, grad);
        if (f.stop(x))
            S.status = Status::ObjectiveCustomStop;
        cur.fDeltaCount = (cur.fDelta < stop.fDelta) ? cur.fDeltaCount + 1 : 0;
        if (S.iteration_callback)
        {
            const psb200_nl_criteria c = as_c(cur);
            if (S.iteration_callback(S.iteration_user, &c)) // Solver.cpp:548-552
                S.status = Status::ObjectiveCustomStop;
        }
        if (++cur.iterations >= stop.iterations)
            S.status = Status::IterationLimit;
    } while (f.callback(as_c(cur), x) && S.status == Status::Continue); // Solver.cpp:558
    total_time = now_s() - t_start;

    if (ok && !S.allow_out_of_iterations && S.status == Status::IterationLimit)
        fail(Status::IterationLimit, "Reached iteration limit (limit=" + std::to_string(stop.iterations) + ")");
    const bool succeeded = S.status == Status::GradNormTolerance || S.status == Status::RelGradNormTolerance ||
                           (S.allow_non_grad_convergence && is_converged_status(S.status));

    // solver_info (Solver.cpp:615-637, Newton.cpp:333-340)
    std::ostringstream o;
    o << "{\"status\":" << jstr(status_message(S.status)) << ",\"succeeded\":" << (succeeded ? "true" : "false")
      << ",\"iterations\":" << cur.iterations << ",\"energy\":" << jnum(f.value(x)) << ",\"grad_norm\":" << jnum(cur.gradNorm)
      << ",\"x_delta\":" << jnum(cur.xDelta) << ",\"line_search\":" << jstr(S.ls.method)
      << ",\"line_search_iterations\":" << S.ls.total_iterations << ",\"solver\":" << jstr(S.solver_name)
      << ",\"final_strategy\":" << jstr(name()) << ",\"total_time\":" << jnum(total_time) << ",\"time_obj_fun\":" << jnum(obj_time)
      << ",\"time_grad\":" << jnum(grad_time) << ",\"time_update_direction\":" << jnum(dir_time) << ",\"time_line_search\":" << jnum(ls_time);
    double t_asm = 0, t_inv = 0;
    o << ",\"internal_solver\":[";
    bool first = true;
    for (auto &s : S.strategies)
        if (auto *nw = dynamic_cast<NewtonStrategy *>(s.get()))
        {
            t_asm += nw->assembly_time;
            t_inv += nw->inverting_time;
            for (const std::string &i : nw->internal_info)
            {
                if (!first)
                    o << ",";
                o << i;
                first = false;
            }
        }
    o << "],\"time_assembly\":" << jnum(t_asm) << ",\"time_inverting\":" << jnum(t_inv) << "}";
    S.info_json = o.str();
    return ok;
}

thread_local std::string g_nl_create_error;

} // namespace

extern "C" {

int psb200_nl_create(psb200_nl_handle *out, const char *solver_params_json, const char *linear_params_json)
{
    if (!out)
        return PSB200_ERR_INVALID;
    *out = nullptr;
    try
    {
        auto s = std::make_unique<psb200_nl_solver>();
        build(*s, solver_params_json ? solver_params_json : "", linear_params_json ? linear_params_json : "");
        *out = s.release();
        return PSB200_OK;
    }
    catch (const std::exception &e)
    {
        g_nl_create_error = e.what();
        return PSB200_ERR_INVALID;
    }
}

int psb200_nl_destroy(psb200_nl_handle h)
{
    delete h;
    return PSB200_OK;
}

int psb200_nl_set_linear_solver_hook(psb200_nl_handle h, void (*hook)(void *user, void *lin), void *user)
{
    if (!h || !hook)
        return PSB200_ERR_INVALID;
    for (auto &st : h->strategies)
        if (auto *ns = dynamic_cast<NewtonStrategy *>(st.get()))
            hook(user, ns->lin);
    return PSB200_OK;
}

int psb200_nl_set_iteration_callback(psb200_nl_handle h, int (*callback)(void *user, const psb200_nl_criteria *state), void *user)
{
    if (!h)
        return PSB200_ERR_INVALID;
    h->iteration_callback = callback;
    h->iteration_user = user;
    return PSB200_OK;
}

int psb200_nl_set_direction_filter(psb200_nl_handle h, void (*filter)(void *user, const double *x, double *dx_inout, int64_t n), void *user)
{
    if (!h)
        return PSB200_ERR_INVALID;
    h->direction_filter = filter;
    h->filter_user = user;
    return PSB200_OK;
}

int psb200_nl_minimize(psb200_nl_handle h, const psb200_nl_problem *problem, double *x_inout, int64_t n)
{
    if (!h || !problem || !x_inout || n < 0 || !problem->value || !problem->gradient || !problem->hessian)
        return PSB200_ERR_INVALID;
    try
    {
        h->err.clear();
        std::vector<double> x(x_inout, x_inout + n);
        Problem f{problem, n};
        f.norm_type = h->norm_type;
        const bool ok = minimize(*h, f, x);
        std::copy(x.begin(), x.end(), x_inout);
        return ok ? PSB200_OK : PSB200_ERR_NUMERIC;
    }
    catch (const std::exception &e)
    {
        h->err = e.what();
        return PSB200_ERR_NUMERIC;
    }
}

int psb200_nl_get_info(psb200_nl_handle h, char *json_out, size_t cap, size_t *needed)
{
    if (!h)
        return PSB200_ERR_INVALID;
    const size_t need = h->info_json.size() + 1;
    if (needed)
        *needed = need;
    if (!json_out || cap < need)
        return PSB200_ERR_INVALID;
    std::memcpy(json_out, h->info_json.c_str(), need);
    return PSB200_OK;
}

const char *psb200_nl_last_error(psb200_nl_handle h) { return h ? h->err.c_str() : g_nl_create_error.c_str(); }

} // extern "C"
