"""TEST INFRASTRUCTURE ONLY. PARITY UNPINNED (see oracle_core.hpp header).

Python restatement of the reference's Newton + line-search control flow, used to check the C++ driver
(polysolve_b200/csrc/newton.cpp) step for step on small problems:
  nonlinear::Solver::minimize          reference src/polysolve/nonlinear/Solver.cpp:255-582
  Newton chain + residual check        descent_strategies/Newton.cpp:14-58,144-214,275-291,326-330
  LineSearch / Backtracking / Armijo   line_search/LineSearch.cpp:73-254, Backtracking.cpp:15-83, Armijo.cpp:13-32
  checkConvergence                     Criteria.cpp:59-96
The linear solve is pluggable (`linsolve(H_csc, rhs, x0) -> x`): the tests pass the oracle's Eigen-faithful
Jacobi-PCG or a scipy direct solve."""
import math

import numpy as np
import scipy.sparse as sp

NaN = float("nan")


def check_convergence(stop, cur):
    if stop["iterations"] > 0 and cur["iterations"] > stop["iterations"]:
        return "IterationLimit"
    sg = stop["firstGradNorm"] if cur["iterations"] == 0 else stop["gradNorm"]
    if sg > 0 and cur["gradNorm"] < sg:
        return "GradNormTolerance"
    if stop["relXDelta"] > 0 and cur["relXDelta"] < stop["relXDelta"]:
        return "RelXDeltaTolerance"
    if stop["relGradNorm"] > 0 and cur["relGradNorm"] < stop["relGradNorm"]:
        return "RelGradNormTolerance"
    if stop.get("newtonDecrement", 0) > 0 and cur["newtonDecrement"] < stop["newtonDecrement"]:
        return "NewtonDecrementTolerance"
    if stop["xDelta"] > 0 and cur["xDelta"] < stop["xDelta"]:
        return "XDeltaTolerance"
    if stop["fDelta"] > 0 and cur["fDelta"] < stop["fDelta"] and cur["fDeltaCount"] >= stop["fDeltaCount"]:
        return "FDeltaTolerance"
    if stop["xDeltaDotGrad"] < 0 and cur["xDeltaDotGrad"] > stop["xDeltaDotGrad"]:
        return "NotDescentDirection"
    return "Continue"


def _grad_norm(f, g, nt):
    """Problem::grad_norm (Problem.hpp:120), default Euclidean"""
    return float(f.grad_norm(g, nt)) if hasattr(f, "grad_norm") else float(np.linalg.norm(g))


def _step_norm(f, d, nt):
    """Problem::step_norm (Problem.hpp:121)"""
    return float(f.step_norm(d, nt)) if hasattr(f, "step_norm") else float(np.linalg.norm(d))


def _rescaling(f, which, nt):
    """Problem::grad_norm_rescaling / step_norm_rescaling / energy_norm_rescaling (Problem.hpp:116-118), default 1"""
    fn = getattr(f, which + "_norm_rescaling", None)
    return float(fn(nt)) if fn else 1.0


class LineSearch:
    def __init__(self, p):
        self.norm_type = p.get("norm_type", "L2")        # Solver.cpp:253
        ls = p.get("line_search", {})
        self.method = ls.get("method", "RobustArmijo")
        self.min_step = ls.get("min_step_size", 1e-10)
        self.max_iter = ls.get("max_step_size_iter", 30)
        self.min_step_final = ls.get("min_step_size_final", 1e-20)
        self.max_iter_final = ls.get("max_step_size_iter_final", 100)
        self.init = ls.get("default_init_step_size", 1.0)
        self.ratio = ls.get("step_ratio", 0.5)
        self.use_grad_norm_tol = ls.get("use_grad_norm_tol", 1e-6)
        self.c = ls.get("Armijo", {}).get("c", 1e-4)
        self.delta_rel_tol = ls.get("RobustArmijo", {}).get("delta_relative_tolerance", 0.1)  # nonlinear-solver-spec.json:683-688
        self.final = False
        self.total = 0

    def lim(self):
        return (self.min_step_final, self.max_iter_final) if self.final else (self.min_step, self.max_iter)

    def line_search(self, x, dx, f):
        it = 0
        e0 = f.value(x)
        if math.isnan(e0):
            return NaN
        g0 = np.asarray(f.gradient(x), float)
        if not np.all(np.isfinite(g0)):
            return NaN
        step = self.init
        mn, mx = self.lim()
        while step > mn and it < mx:
            nx = x + step * dx
            if (not f.is_step_valid(x, nx)) or (not math.isfinite(f.value(nx))):
                step *= self.ratio
            else:
                break
            it += 1
        if it >= mx or step <= mn:
            return NaN
        f.line_search_begin(x, x + step * dx)
        ms = f.max_step_size(x, x + step * dx)
        if ms == 0:
            f.line_search_end()
            return NaN
        step = step * ms  # the reference rounds this product downward (LineSearch.cpp:243-248); exact for ms == 1
        gn = _grad_norm(f, g0, self.norm_type)
        if gn < 1e-30:
            self.total += it
            return step
        use_gn = gn < self.use_grad_norm_tol * _rescaling(f, "grad", self.norm_type)   # LineSearch.cpp:142
        arm = self.c * float(dx @ g0)
        if self.method == "None":
            f.solution_changed(x + step * dx)                  # NoLineSearch.cpp:11-22; the checks below still apply
        while self.method != "None" and step > mn and it < mx:
            nx = x + step * dx
            f.solution_changed(nx)
            ok = False
            if f.is_step_valid(x, nx):
                e = f.value(nx)
                if math.isfinite(e):
                    if self.method in ("Armijo", "RobustArmijo"):
                        ok = e <= e0 + step * arm  # Armijo.cpp:20-32; RobustArmijo.cpp:27 tries it first
                        if not ok and self.method == "RobustArmijo" and abs(e - e0) <= self.delta_rel_tol * abs(e0):
                            # RobustArmijo.cpp:30-44
                            ng = np.asarray(f.gradient(nx), float)
                            dE = step / 2 * float(dx @ (ng + g0))
                            eps = step / 2 * abs(float(dx @ (ng - g0)))
                            ok = dE + eps <= step * arm
                    elif use_gn or self.method == "ResidualBacktracking":   # ResidualBacktracking.cpp:15-28
                        # Backtracking.cpp:76-80 evaluates both norms at every trial ("TODO cache old grad norm")
                        ok = _grad_norm(f, np.asarray(f.gradient(nx), float), self.norm_type) < _grad_norm(f, g0, self.norm_type)
                    else:
                        ok = e < e0
            if ok:
                break
            step *= self.ratio
            it += 1
        self.total += it
        if it >= mx or step <= mn:
            f.solution_changed(x)
            f.line_search_end()
            return NaN
        f.line_search_end()
        return step


class BFGSMat:
    """LBFGSpp::BFGSMat as the reference's LBFGS strategy uses it (descent_strategies/LBFGS.hpp:38, LBFGS.cpp:48-51;
    LBFGSpp is un-vendored, cmake/recipes/LBFGSpp.cmake -- restated from its published algorithm):
    add_correction stores (s, y), ys = s.y, theta = y.y / s.y in a ring of m pairs; apply_Hv is the two-loop recursion
    with H0 = I / theta."""

    def __init__(self, n, m):
        self.m, self.S, self.Y = m, np.zeros((m, n)), np.zeros((m, n))
        self.ys, self.alpha = np.zeros(m), np.zeros(m)
        self.theta, self.ncorr, self.ptr = 1.0, 0, 0

    def add_correction(self, s, y):
        loc = self.ptr % self.m
        self.S[loc], self.Y[loc] = s, y
        self.ys[loc] = float(s @ y)
        self.theta = float(y @ y) / self.ys[loc]
        if self.ncorr < self.m:
            self.ncorr += 1
        self.ptr = loc + 1

    def apply_Hv(self, v, a):
        res = a * v
        j = self.ptr % self.m
        for _ in range(self.ncorr):
            j = (j + self.m - 1) % self.m
            self.alpha[j] = float(self.S[j] @ res) / self.ys[j]
            res = res - self.alpha[j] * self.Y[j]
        res = res / self.theta
        for _ in range(self.ncorr):
            beta = float(self.Y[j] @ res) / self.ys[j]
            res = res + (self.alpha[j] - beta) * self.S[j]
            j = (j + 1) % self.m
        return res


class LbfgsStrategy:
    """LBFGS::reset / compute_update_direction (LBFGS.cpp:22-61)."""

    def __init__(self, history):
        self.history, self.mat, self.prev_x, self.prev_g = history, None, None, None

    def reset(self):
        self.mat, self.prev_x, self.prev_g = None, None, None

    def direction(self, x, grad):
        if self.prev_x is None:
            d = -grad
            self.mat = BFGSMat(x.size, self.history)
        else:
            self.mat.add_correction(x - self.prev_x, grad - self.prev_g)
            d = self.mat.apply_Hv(grad, -1.0)
        self.prev_x, self.prev_g = x.copy(), grad.copy()
        return d


def minimize(problem, x, params, linsolve, iteration_callback=None, direction_filter=None):
    """Returns (x, info). Raises RuntimeError where the reference throws. iteration_callback(state) -> bool and
    direction_filter(x, dx) (in place) are Solver::set_iteration_callback / set_direction_filter (Solver.hpp:76-86)."""
    adv = params.get("advanced", {})
    nt = params.get("norm_type", "L2")
    rg, rs, re = _rescaling(problem, "grad", nt), _rescaling(problem, "step", nt), _rescaling(problem, "energy", nt)
    # Solver.cpp:206-219 + Solver.hpp:118-131 (reset_stopping_criteria)
    stop = dict(xDelta=params.get("x_delta_tol", 0) * rs, fDelta=adv.get("f_delta_tol", 0) * re,
                gradNorm=params.get("grad_norm_tol", 1e-10) * rg, firstGradNorm=params.get("first_grad_norm_tol", 1e-12) * rg,
                xDeltaDotGrad=-adv.get("derivative_along_delta_x_tol", 0) * re,
                relGradNorm=params.get("rel_grad_norm_tol", 1e-10), relXDelta=params.get("rel_x_delta_tol", 0),
                newtonDecrement=params.get("newton_decrement_tol", 0) * re,
                iterations=params.get("max_iterations", 500), fDeltaCount=adv.get("f_delta_step_tol", 100))
    nw = params.get("Newton", {})
    res_tol = nw.get("residual_tolerance", 1e-5)
    wmin, wmax, winc = nw.get("reg_weight_min", 1e-8), nw.get("reg_weight_max", 1e8), nw.get("reg_weight_inc", 10)
    # a strategy: [name, project_to_psd, reg_weight, residual_tolerance, reg_weight_min, reg_weight_max, reg_weight_inc]
    strategies = []
    lbfgs = None
    if isinstance(params.get("solver"), (list, tuple)):
        # Solver.cpp:147-154: the strategies of the list in order, each from its own entry, no GradientDescent fallback;
        # parameters by extract_param (Utils.cpp:78-84): entry[Type][name], else entry[name], else the /solver/*/ default
        for e in params["solver"]:
            t = e["type"]

            def ex(key, name, default, e=e):
                return e[key][name] if isinstance(e.get(key), dict) and name in e[key] else e.get(name, default)
            if t in ("Newton", "SparseNewton", "sparse_newton"):
                strategies.append(["Newton", False, 0.0, ex("Newton", "residual_tolerance", 1e-5), 0, 0, 0])
            elif t == "ProjectedNewton":
                strategies.append(["ProjectedNewton", True, 0.0, ex(t, "residual_tolerance", 1e-5), 0, 0, 0])
            elif t in ("RegularizedNewton", "RegularizedProjectedNewton"):
                k = "RegularizedNewton"
                w0 = ex(k, "reg_weight_min", 1e-8)
                strategies.append(["RegularizedNewton", t == "RegularizedProjectedNewton", w0, ex(k, "residual_tolerance", 1e-5), w0,
                                   ex(k, "reg_weight_max", 1e8), ex(k, "reg_weight_inc", 10)])
            elif t in ("L-BFGS", "LBFGS"):
                lbfgs = LbfgsStrategy(ex("L-BFGS", "history_size", 6))
                strategies.append(["L-BFGS", False, 0.0, 0, 0, 0, 0])
            elif t in ("GradientDescent", "gradient_descent"):
                strategies.append(["GradientDescent", False, 0.0, 0, 0, 0, 0])
            else:
                raise RuntimeError("Unrecognized solver type: " + t)
    else:
        if params.get("solver", "Newton") in ("L-BFGS", "LBFGS"):       # Solver.cpp:83-85
            lbfgs = LbfgsStrategy(params.get("L-BFGS", {}).get("history_size", 6))
            strategies.append(["L-BFGS", False, 0.0, 0, 0, 0, 0])
        elif params.get("solver", "Newton") in ("GradientDescent", "gradient_descent"):   # Solver.cpp:92-94: GD alone
            pass
        else:
            if not nw.get("force_psd_projection", False):
                strategies.append(["Newton", False, 0.0, res_tol, 0, 0, 0])
            if nw.get("use_psd_projection", True):
                strategies.append(["ProjectedNewton", True, 0.0, res_tol, 0, 0, 0])
            if wmin > 0:
                strategies.append(["RegularizedNewton", nw.get("use_psd_projection_in_regularized", True), wmin, res_tol, wmin, wmax, winc])
        strategies.append(["GradientDescent", False, 0.0, 0, 0, 0, 0])   # Solver.cpp:176-181
    per = params.get("iterations_per_strategy", 5)                  # Solver.cpp:232-245: one value or one per strategy + 1
    if isinstance(per, (list, tuple)):
        if len(per) != len(strategies) + 1:
            raise RuntimeError(f"Invalit iter_per_strategy size: {len(per)}!={len(strategies) + 1}")
        per = list(per)
    else:
        per = [per] * (len(strategies) + 1)
    ls = LineSearch(params)
    x = np.array(x, float)
    n = x.size
    cur = dict(iterations=0, xDelta=0.0, fDelta=0.0, gradNorm=0.0, xDeltaDotGrad=0.0, relGradNorm=0.0, relXDelta=0.0, fDeltaCount=0)
    strategy = prev = 0
    cur_iter = 0
    dx = np.zeros(n)
    old_energy = NaN
    status = "NotStarted"
    lin_iters = []
    problem.solution_changed(x)
    problem.post_step(0, x, np.zeros(n))   # Solver.cpp:285-286: once before the loop, gradient still zero
    g0n = dx0n = NaN

    def handle_error(s):
        if s[0] != "RegularizedNewton":
            return False
        s[2] *= s[6]
        return s[2] < s[5]

    def reset():
        for s in strategies:
            if s[0] == "RegularizedNewton":
                s[2] = s[4]
        if lbfgs is not None:
            lbfgs.reset()

    def keep_going():
        # Problem::callback(state, x) (Problem.hpp:109) is the first operand of the do-while condition (Solver.cpp:558):
        # it runs at the end of EVERY trip, also after the `continue`s of a strategy fallback
        return bool(problem.callback(dict(cur), x)) if hasattr(problem, "callback") else True

    is_residual = bool(problem.is_residual()) if hasattr(problem, "is_residual") else False
    cur.update(energy=NaN, alpha=NaN, step=NaN, newtonDecrement=0.0, firstGradNorm=0.0)
    while True:
        ls.final = strategy == len(strategies) - 1
        energy = problem.value(x)
        cur["energy"] = energy
        if not math.isfinite(energy):
            raise RuntimeError("f(x) is nan or inf; stopping")
        cur["fDelta"] = abs(old_energy - energy)
        grad = np.asarray(problem.gradient(x), float)
        cur["gradNorm"] = _grad_norm(problem, grad, nt)
        if cur["iterations"] == 0:
            g0n = cur["gradNorm"]
            cur["relGradNorm"] = NaN
        else:
            cur["relGradNorm"] = cur["gradNorm"] / g0n
        cur["xDelta"] = cur["xDeltaDotGrad"] = cur["relXDelta"] = cur["newtonDecrement"] = NaN
        status = check_convergence(stop, cur)
        if status != "Continue":
            break
        s = strategies[strategy]
        ok = True
        if s[0] == "GradientDescent":
            dx = -grad
        elif s[0] == "L-BFGS":
            dx = lbfgs.direction(x, grad)
        else:
            H = sp.csc_matrix(problem.hessian(x, s[1]))
            if s[0] == "RegularizedNewton" and s[2] > 0:
                H = (H + s[2] * sp.identity(n, format="csc")).tocsc()
            try:
                dx, it = linsolve(H, -grad, dx.copy())
                lin_iters.append(it)
                r = float(np.linalg.norm(H @ dx + grad))
                ok = not (math.isnan(r) or r > s[3])
            except ArithmeticError:
                ok = False
        if direction_filter is not None and ok:               # Solver.cpp:353-358
            dx = np.array(dx, float)
            direction_filter(x, dx)
        cur["xDelta"] = _step_norm(problem, dx, nt)
        if cur["iterations"] == 0:
            dx0n = cur["xDelta"]
            cur["relXDelta"] = NaN
        else:
            cur["relXDelta"] = cur["xDelta"] / dx0n
        if (not ok) or math.isnan(cur["xDelta"]):
            if not handle_error(s):
                strategy += 1
            if strategy >= len(strategies):
                raise RuntimeError("Update direction could not be computed on last strategy; stopping")
            if not keep_going():
                break
            continue
        if direction_filter is not None:                      # Solver.cpp:392-403
            ng = -grad
            direction_filter(x, ng)
            cur["xDeltaDotGrad"] = -float(dx @ ng)
        else:
            cur["xDeltaDotGrad"] = float(dx @ grad)
        if stop["newtonDecrement"] > 0:                       # Solver.cpp:409-423: 1/2 x^T H x (sic), NaN on failure
            try:
                cur["newtonDecrement"] = 0.5 * float(x @ (sp.csc_matrix(problem.hessian(x, False)) @ x))
            except RuntimeError:
                cur["newtonDecrement"] = NaN
        if (not is_residual) and cur["gradNorm"] != 0 and cur["xDeltaDotGrad"] >= 0:
            if not handle_error(s):
                strategy += 1
            if strategy >= len(strategies):
                raise RuntimeError("Search direction not a descent direction on last strategy; stopping")
            if not keep_going():
                break
            continue
        status = check_convergence(stop, cur)
        if status != "Continue":
            break
        rate = ls.line_search(x, dx, problem)
        cur["alpha"] = rate
        if math.isnan(rate):
            if not handle_error(s):
                strategy += 1
            if strategy >= len(strategies):
                raise RuntimeError("Line search failed on last strategy; stopping")
            if not keep_going():
                break
            continue
        x1 = x + rate * dx
        if hasattr(problem, "after_line_search_custom_operation") and problem.after_line_search_custom_operation(x, x1):
            problem.solution_changed(x1)                      # Solver.cpp:495-499
        x = x1
        old_energy = energy
        if strategy != prev:
            cur_iter = 0
        if strategy != 0 and cur_iter >= per[strategy]:
            strategy = 0
            reset()
        prev = strategy
        cur_iter += 1
        cur["step"] = float(np.linalg.norm(rate * dx))
        problem.post_step(cur["iterations"], x, grad)
        if problem.stop(x):
            status = "ObjectiveCustomStop"
        cur["fDeltaCount"] = cur["fDeltaCount"] + 1 if cur["fDelta"] < stop["fDelta"] else 0
        if iteration_callback is not None and iteration_callback(dict(cur)):   # Solver.cpp:548-552
            status = "ObjectiveCustomStop"
        cur["iterations"] += 1
        if cur["iterations"] >= stop["iterations"]:
            status = "IterationLimit"
        if (not keep_going()) or status != "Continue":
            break
    if status == "IterationLimit" and not params.get("allow_out_of_iterations", False):
        raise RuntimeError("Reached iteration limit")
    return x, dict(status=status, iterations=cur["iterations"], energy=problem.value(x), grad_norm=cur["gradNorm"],
                   linear_iterations=lin_iters, final_strategy=strategies[min(strategy, len(strategies) - 1)][0],
                   line_search_iterations=ls.total)
