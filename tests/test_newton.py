"""Newton + line-search driver (include/psb200_nl.h, polysolve_b200/csrc/newton.cpp) against the Python
restatement of the reference's control flow (oracle/newton_oracle.py) -- SURVEY 8 rows a11/a12, config 5.

The analytic problems and the acceptance rule (||x - x*|| < 1e-7 or ||grad|| < 1e-7) are those of the reference's
tests/test_nonlinear_solver.cpp:30-325,422-426."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class Base:
    def solution_changed(self, x): pass
    def is_step_valid(self, x0, x1): return True
    def max_step_size(self, x0, x1): return 1.0
    def line_search_begin(self, x0, x1): pass
    def line_search_end(self): pass
    def post_step(self, it, x, g): pass
    def stop(self, x): return False


class Rosenbrock(Base):
    """tests/test_nonlinear_solver.cpp: Rosenbrock(N), minimum at 1."""
    def __init__(self, n=10): self.n = n
    def value(self, x): return float(np.sum(100 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2))
    def gradient(self, x):
        g = np.zeros(self.n)
        g[:-1] += -400 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2 * (1 - x[:-1])
        g[1:] += 200 * (x[1:] - x[:-1] ** 2)
        return g
    def hessian(self, x, psd=False):
        d = np.zeros(self.n)
        d[:-1] += 1200 * x[:-1] ** 2 - 400 * x[1:] + 2
        d[1:] += 200
        o = -400 * x[:-1]
        H = sp.diags([o, d, o], [-1, 0, 1], format="csc")
        if psd:
            w, v = np.linalg.eigh(H.toarray())
            H = sp.csc_matrix((v * np.maximum(w, 1e-8)) @ v.T)
        return H
    def solutions(self): return [np.ones(self.n)]


class Sphere(Base):
    def __init__(self, n=10): self.n = n
    def value(self, x): return float(x @ x)
    def gradient(self, x): return 2 * x
    def hessian(self, x, psd=False): return 2 * sp.identity(self.n, format="csc")
    def solutions(self): return [np.zeros(self.n)]


class Quadratic(Base):
    """f = sum_i (i+1) (x_i - 1)^2 + coupling; SPD tridiagonal Hessian."""
    def __init__(self, n=12):
        self.n = n
        self.A = sp.diags([-np.ones(n - 1), 2.5 + np.arange(n) * 0.1, -np.ones(n - 1)], [-1, 0, 1], format="csc")
        self.b = np.linspace(-1, 1, n)
    def value(self, x): return float(0.5 * x @ (self.A @ x) - self.b @ x)
    def gradient(self, x): return self.A @ x - self.b
    def hessian(self, x, psd=False): return self.A
    def solutions(self): return [spla.spsolve(self.A, self.b)]


PARAMS = {"solver": "Newton", "line_search": {"method": "Backtracking"}, "grad_norm_tol": 1e-8, "rel_grad_norm_tol": 0,
          "max_iterations": 200, "Newton": {"residual_tolerance": 1e-5}}


def direct(H, rhs, x0):
    return spla.spsolve(H.tocsc(), rhs), 1


def test_newton_oracle_converges_on_reference_problems():
    """CPU: the restatement itself reaches the minima the reference's tests require."""
    from oracle import newton_oracle as NO
    rng = np.random.default_rng(0)
    for prob in (Rosenbrock(10), Sphere(10), Quadratic(12)):
        for _ in range(3):
            x0 = rng.uniform(-1, 1, prob.n)
            x, info = NO.minimize(prob, x0, PARAMS, direct)
            assert info["status"] == "GradNormTolerance", info
            assert min(np.linalg.norm(x - s) for s in prob.solutions()) < 1e-7 or info["grad_norm"] < 1e-7


class _Recorder:
    """wraps a problem and records what post_step is handed"""
    def __init__(self, p):
        self.p, self.n, self.steps = p, p.n, []
    def __getattr__(self, k): return getattr(self.p, k)
    def post_step(self, it, x, g): self.steps.append((int(it), np.array(x, float), np.array(g, float)))


# Criteria.cpp:98-133 (status_message): the text the driver reports for the restatement's status names
_STATUS_TEXT = {"GradNormTolerance": "Gradient vector norm too small", "XDeltaTolerance": "Change in parameter vector too small",
                "RelXDeltaTolerance": "Relative change in parameter vector too small",
                "RelGradNormTolerance": "Relative gradient vector too small",
                "FDeltaTolerance": "Change in cost function value too small", "IterationLimit": "Iteration limit reached"}
_GD_VARIANTS = [
    {}, {"x_delta_tol": 1e-3}, {"rel_x_delta_tol": 1e-2}, {"rel_grad_norm_tol": 1e-2}, {"first_grad_norm_tol": 1e3},
    {"advanced": {"f_delta_tol": 1e-4, "f_delta_step_tol": 3}}, {"max_iterations": 5, "allow_out_of_iterations": True},
    {"max_iterations": 5}, {"line_search": {"method": "Backtracking", "use_grad_norm_tol": 1e-4}},
    {"line_search": {"method": "Armijo", "Armijo": {"c": 0.3}}},
    {"line_search": {"method": "RobustArmijo", "RobustArmijo": {"delta_relative_tolerance": 0.1}}},
    {"line_search": {"method": "Backtracking", "default_init_step_size": 0.5, "step_ratio": 0.3}},
    {"line_search": {"method": "Armijo", "min_step_size": 1e-3, "max_step_size_iter": 4}},
    {"line_search": {"method": "None"}},
    {"line_search": {"method": "ResidualBacktracking"}},
    {"line_search": {"method": "ResidualBacktracking", "step_ratio": 0.7, "max_step_size_iter": 12}},
    {"line_search": {"method": "None", "default_init_step_size": 0.05}},
]


@pytest.mark.filterwarnings("ignore::RuntimeWarning")
@pytest.mark.parametrize("variant", range(len(_GD_VARIANTS)))
def test_driver_outer_loop_matches_oracle_on_cpu(psb, variant):
    """CPU (no linear solve involved): with "solver": "GradientDescent" (Solver.cpp:92-94) the C++ driver runs without a
    GPU, so its outer loop, every stopping criterion (Criteria.cpp:59-96), the iteration-limit error and the three line
    searches are compared with the restatement step for step: identical status, iteration count, error text and
    BIT-IDENTICAL iterates / gradients at every post_step call (Solver.cpp:286,536)."""
    from oracle import newton_oracle as NO
    v = _GD_VARIANTS[variant]
    base = {"solver": "GradientDescent", "grad_norm_tol": 1e-6, "rel_grad_norm_tol": 0, "max_iterations": 300}
    for method in (["Backtracking", "Armijo", "RobustArmijo"] if "line_search" not in v else [None]):
        P = dict(base, **v)
        if method:
            P["line_search"] = {"method": method}
        for prob in (Quadratic(12), Rosenbrock(4)):
            x0 = np.random.default_rng(1).uniform(-1, 1, prob.n)
            po, pd = _Recorder(prob), _Recorder(prob)
            xo = eo = ed = None
            io = info = {}
            try:
                xo, io = NO.minimize(po, x0.copy(), P, direct)
            except RuntimeError as e:
                eo = str(e)
            x = x0.copy()
            s = psb.NonlinearSolver.create(P, {"solver": "CUDA"})
            try:
                s.minimize(pd, x)
                info = s.get_info()
            except RuntimeError as e:
                ed = str(e)
            assert (eo is None) == (ed is None), (P, eo, ed)
            if eo is not None:
                assert eo.split(";")[0] in ed                 # the driver prefixes "[strategy][line search] "
            else:
                assert info["iterations"] == io["iterations"] and np.array_equal(x, xo), (P, info, io)
                assert info["status"] == _STATUS_TEXT[io["status"]], (P, info["status"], io["status"])
            assert len(po.steps) == len(pd.steps) >= 1
            for (ia, xa, ga), (ib, xb, gb) in zip(po.steps, pd.steps):
                assert ia == ib and np.array_equal(xa, xb) and np.array_equal(ga, gb), (P, ia)


class _Hooked(Quadratic):
    """Quadratic with every optional Problem callback active (Problem.hpp:79-143): a step-size cap, a validity box, a
    custom stop, a region where the energy is NaN; records the order in which the solver calls them."""
    def __init__(self, cap=1.0, box=None, stop_at=None, nan_region=None):
        super().__init__(12)
        self.cap, self.box, self.stop_at, self.nan_region, self.calls = cap, box, stop_at, nan_region, []
    def value(self, x):
        if self.nan_region is not None and np.abs(x).max() > self.nan_region:
            return float("nan")
        return super().value(x)
    def solution_changed(self, x): self.calls.append("changed")
    def max_step_size(self, x0, x1):
        self.calls.append("max")
        return self.cap
    def is_step_valid(self, x0, x1):
        self.calls.append("valid")
        return True if self.box is None else bool(np.all(np.abs(x1) < self.box))
    def line_search_begin(self, x0, x1): self.calls.append("begin")
    def line_search_end(self): self.calls.append("end")
    def stop(self, x):
        self.calls.append("stop")
        return self.stop_at is not None and super().value(x) < self.stop_at


@pytest.mark.filterwarnings("ignore::RuntimeWarning")
@pytest.mark.parametrize("hooks", [dict(cap=0.3), dict(box=0.9), dict(stop_at=-1.0), dict(nan_region=1.5),
                                   dict(cap=0.5, box=1.2, stop_at=-1.2), dict(cap=1e-11)])
def test_driver_calls_the_problem_hooks_like_the_oracle_on_cpu(psb, hooks):
    """CPU: max_step_size / is_step_valid / line_search_begin / line_search_end / solution_changed / stop
    (LineSearch.cpp:73-254, Solver.cpp:255-582) -- same status or error, bit-identical iterates and the SAME SEQUENCE of
    callback invocations as the restatement, for the three line searches."""
    from oracle import newton_oracle as NO
    for method in ("Backtracking", "Armijo", "RobustArmijo", "ResidualBacktracking", "None"):
        P = {"solver": "GradientDescent", "grad_norm_tol": 1e-6, "rel_grad_norm_tol": 0, "max_iterations": 300,
             "line_search": {"method": method, "default_init_step_size": 0.1 if method == "None" else 1.0}}
        x0 = np.random.default_rng(1).uniform(-0.8, 0.8, 12)
        if "nan_region" in hooks:
            x0 = np.clip(3.0 * x0, -1.4, 1.4)
        po, pd = _Recorder(_Hooked(**hooks)), _Recorder(_Hooked(**hooks))
        xo = eo = ed = None
        io = info = {}
        try:
            xo, io = NO.minimize(po, x0.copy(), P, direct)
        except RuntimeError as e:
            eo = str(e)
        x = x0.copy()
        s = psb.NonlinearSolver.create(P, {"solver": "CUDA"})
        try:
            s.minimize(pd, x)
            info = s.get_info()
        except RuntimeError as e:
            ed = str(e)
        assert (eo is None) == (ed is None), (hooks, method, eo, ed)
        if eo is None:
            assert info["iterations"] == io["iterations"] and np.array_equal(x, xo)
            if io["status"] == "ObjectiveCustomStop":
                assert info["status"] == "Objective function specified to stop"   # Criteria.cpp:120-121
        else:
            assert eo.split(";")[0] in ed
        assert po.p.calls == pd.p.calls and len(po.p.calls) > 10
        assert len(po.steps) == len(pd.steps)
        assert all(a[0] == b[0] and np.array_equal(a[1], b[1]) for a, b in zip(po.steps, pd.steps))


class _Normed(Quadratic):
    """Quadratic with the remaining optional Problem virtuals (Problem.hpp:35,103-121): custom gradient / step norms per
    norm type and tolerance rescalings (what PolyFEM's problems do with their mass matrix), callback(state, x),
    after_line_search_custom_operation, is_residual. Records calls and the Criteria handed to callback."""
    def __init__(self, resc=(1, 1, 1), cb_stop=None, after=False, residual=False):
        super().__init__(12)
        self.resc, self.cb_stop, self.after, self.residual, self.calls, self.states = resc, cb_stop, after, residual, [], []
    @staticmethod
    def _norm(v, nt):
        return float(np.abs(v).max()) if nt == "Linf" else float(np.sqrt(np.sum(v * v) / (len(v) if nt == "L2" else 1)))
    def grad_norm(self, g, nt):
        self.calls.append(("gn", nt))
        return self._norm(g, nt)
    def step_norm(self, d, nt):
        self.calls.append(("sn", nt))
        return self._norm(d, nt)
    def grad_norm_rescaling(self, nt): return self.resc[0]
    def step_norm_rescaling(self, nt): return self.resc[1]
    def energy_norm_rescaling(self, nt): return self.resc[2]
    def callback(self, state, x):
        self.states.append(dict(state))
        return not (self.cb_stop is not None and state["iterations"] >= self.cb_stop)
    def after_line_search_custom_operation(self, x0, x1):
        self.calls.append("after")
        return self.after
    def solution_changed(self, x): self.calls.append("changed")
    def is_residual(self): return self.residual


_NORMED_CASES = [
    (dict(), {"norm_type": "L2"}), (dict(), {"norm_type": "Linf"}), (dict(), {"norm_type": "Euclidean"}), (dict(), {}),
    (dict(resc=(10, 1, 1)), {"norm_type": "Linf"}), (dict(resc=(1, 100, 1)), {"x_delta_tol": 1e-4}),
    (dict(resc=(1, 1, 50)), {"advanced": {"f_delta_tol": 1e-6, "f_delta_step_tol": 2}}),
    (dict(cb_stop=3), {}), (dict(after=True), {}), (dict(residual=True), {}), (dict(), {"newton_decrement_tol": 0.5}),
    (dict(), {"line_search": {"method": "Backtracking", "use_grad_norm_tol": 1e-2}}),
    (dict(resc=(1e4, 1, 1)), {"line_search": {"method": "Backtracking", "use_grad_norm_tol": 1e-5}}),
]


@pytest.mark.parametrize("case", range(len(_NORMED_CASES)))
def test_driver_norms_rescalings_callback_match_oracle_on_cpu(psb, case):
    """CPU: "norm_type" with Problem::grad_norm / step_norm, the tolerance rescalings (Solver.hpp:118-131), the line
    search's use_grad_norm switch (LineSearch.cpp:142, Backtracking.cpp:76-80), "newton_decrement_tol" (Solver.cpp:409-423),
    Problem::callback as the loop condition (Solver.cpp:558), after_line_search_custom_operation (:495-499), is_residual
    (:425): same status, iterates bit-identical, the same sequence of norm / solution_changed calls and the same Criteria
    at every callback as the restatement."""
    from oracle import newton_oracle as NO
    kw, pv = _NORMED_CASES[case]
    P = dict({"solver": "GradientDescent", "grad_norm_tol": 1e-6, "rel_grad_norm_tol": 0, "max_iterations": 300,
              "line_search": {"method": "Backtracking"}}, **pv)
    x0 = np.random.default_rng(1).uniform(-0.8, 0.8, 12)
    po, pd = _Recorder(_Normed(**kw)), _Recorder(_Normed(**kw))
    xo, io = NO.minimize(po, x0.copy(), P, direct)
    x = x0.copy()
    s = psb.NonlinearSolver.create(P, {"solver": "CUDA"})
    s.minimize(pd, x)
    info = s.get_info()
    assert info["iterations"] == io["iterations"] and np.array_equal(x, xo)
    if io["status"] in _STATUS_TEXT:
        assert info["status"] == _STATUS_TEXT[io["status"]]
    elif io["status"] == "NewtonDecrementTolerance":
        assert info["status"] == "Newton decrement too small"            # Criteria.cpp:118-119
    else:
        assert io["status"] == "Continue" and info["status"] == "Convergence criteria not reached"   # callback said stop
    assert po.p.calls == pd.p.calls and len(po.p.calls) > 5
    assert len(po.steps) == len(pd.steps) and all(a[0] == b[0] and np.array_equal(a[1], b[1]) for a, b in zip(po.steps, pd.steps))
    assert len(po.p.states) == len(pd.p.states) >= 1
    for a, b in zip(po.p.states, pd.p.states):
        for k in ("iterations", "fDeltaCount"):
            assert a[k] == b[k]
        for k in ("xDelta", "fDelta", "gradNorm", "xDeltaDotGrad", "relGradNorm", "relXDelta", "newtonDecrement", "energy", "alpha", "step"):
            u, v = float(a[k]), float(b[k])
            # numpy's dot and the driver's loop add in different orders: 4 ulp on the dot products, exact elsewhere
            assert (np.isnan(u) and np.isnan(v)) or abs(u - v) <= 1e-15 * max(abs(u), abs(v)) * 4, (k, u, v)
    with pytest.raises(RuntimeError, match="norm_type"):
        psb.NonlinearSolver.create(dict(P, norm_type="L3"), {"solver": "CUDA"})


def test_iterations_per_strategy_list_is_validated(psb):
    """Solver.cpp:232-245: a list needs one entry per strategy + 1 (the reference's message, typo included)."""
    with pytest.raises(RuntimeError, match="Invalit iter_per_strategy size: 2!=5"):
        psb.NonlinearSolver.create({"solver": "Newton", "iterations_per_strategy": [1, 2]}, {"solver": "CUDA"})
    psb.NonlinearSolver.create({"solver": "Newton", "iterations_per_strategy": [1, 2, 3, 4, 5]}, {"solver": "CUDA"})
    psb.NonlinearSolver.create({"solver": "GradientDescent", "iterations_per_strategy": [1, 2]}, {"solver": "CUDA"})


_GD = {"type": "GradientDescent"}
_LIST_CASES = [
    # non-final strategies may halve the step only twice: the line search fails on Rosenbrock, the next strategy takes over
    # (the last one with the *_final limits, LineSearch.hpp current_min_step_size / current_max_step_size_iter) and after
    # iterations_per_strategy[k] successful steps the solver returns to strategy 0 (Solver.cpp:512-522)
    ({"solver": [_GD, _GD], "line_search": {"method": "Backtracking", "max_step_size_iter": 2}, "iterations_per_strategy": [3, 2, 1]}, 4),
    ({"solver": [_GD, _GD, _GD], "line_search": {"method": "Armijo", "max_step_size_iter": 3, "max_step_size_iter_final": 40},
      "iterations_per_strategy": 2}, 4),
    ({"solver": [_GD, _GD], "line_search": {"method": "RobustArmijo", "min_step_size": 0.2, "min_step_size_final": 1e-12},
      "iterations_per_strategy": [1, 4, 1]}, 6),
    ({"solver": [_GD], "line_search": {"method": "Backtracking"}}, 5),
    ({"solver": [_GD, _GD], "line_search": {"method": "Backtracking", "max_step_size_iter": 1, "max_step_size_iter_final": 1}}, 4),
]


@pytest.mark.filterwarnings("ignore::RuntimeWarning")
@pytest.mark.parametrize("case", range(len(_LIST_CASES)))
def test_driver_strategy_fallback_matches_oracle_on_cpu(psb, case):
    """CPU: "solver" as a LIST of strategies (Solver.cpp:147-154; no automatic GradientDescent fallback) made of
    GradientDescent entries drives the strategy machinery without a linear solve: line-search failure -> next strategy,
    final-strategy limits, return to strategy 0 after iterations_per_strategy[k] steps, failure on the last strategy.
    Same status / error, iteration count, total line-search iterations and bit-identical iterates as the restatement."""
    from oracle import newton_oracle as NO
    pv, n = _LIST_CASES[case]
    prob = Rosenbrock(n)
    P = dict({"grad_norm_tol": 1e-5, "rel_grad_norm_tol": 0, "max_iterations": 60, "allow_out_of_iterations": True}, **pv)
    x0 = np.random.default_rng(2).uniform(-1, 1, n)
    po, pd = _Recorder(prob), _Recorder(prob)
    xo = eo = ed = None
    io = info = {}
    try:
        xo, io = NO.minimize(po, x0.copy(), P, direct)
    except RuntimeError as e:
        eo = str(e)
    x = x0.copy()
    s = psb.NonlinearSolver.create(P, {"solver": "CUDA"})
    try:
        s.minimize(pd, x)
        info = s.get_info()
    except RuntimeError as e:
        ed = str(e)
    assert (eo is None) == (ed is None)
    if eo is None:
        assert info["iterations"] == io["iterations"] and np.array_equal(x, xo)
        assert info["line_search_iterations"] == io["line_search_iterations"] and info["final_strategy"] == io["final_strategy"]
        assert info["status"] == _STATUS_TEXT[io["status"]]
    else:
        assert eo.split(";")[0] in ed
    assert len(po.steps) == len(pd.steps) and all(a[0] == b[0] and np.array_equal(a[1], b[1]) for a, b in zip(po.steps, pd.steps))


def test_solver_list_is_validated(psb):
    lin = {"solver": "CUDA"}
    for bad, msg in [([], "empty"), ([{"x": 1}], "type"), ([{"type": "ADAM"}], "Unrecognized solver type: ADAM"),
                     ([{"type": "RegularizedNewton", "reg_weight_min": 0}], "reg_weight_min"),
                     ([{"type": "L-BFGS", "history_size": 0}], "history_size"),
                     ([{"type": "Newton", "residual_tolerance": -1}], "residual_tolerance")]:
        with pytest.raises(RuntimeError, match=msg):
            psb.NonlinearSolver.create({"solver": bad}, lin)
    # every provided type, per-entry parameters in both spellings of extract_param (Utils.cpp:78-84)
    psb.NonlinearSolver.create({"solver": [{"type": "Newton", "residual_tolerance": 1e-6}, {"type": "ProjectedNewton"},
                                           {"type": "RegularizedNewton", "RegularizedNewton": {"reg_weight_min": 1e-6}},
                                           {"type": "RegularizedProjectedNewton"}, {"type": "L-BFGS", "history_size": 4},
                                           {"type": "GradientDescent"}], "iterations_per_strategy": [1, 2, 3, 4, 5, 6, 7]}, lin)
    with pytest.raises(RuntimeError, match="Invalit iter_per_strategy size: 2!=3"):
        psb.NonlinearSolver.create({"solver": [_GD, _GD], "iterations_per_strategy": [1, 2]}, lin)


class RefQuadratic(Base):
    """tests/test_nonlinear_solver.cpp:78-129 QuadraticProblem: (x0+2)^2 + (x1-3)^2 + (x2-1)^2, minimum (-2, 3, 1)."""
    n = 3
    def value(self, x): return float((x[0] + 2) ** 2 + (x[1] - 3) ** 2 + (x[2] - 1) ** 2)
    def gradient(self, x): return np.array([2 * (x[0] + 2), 2 * (x[1] - 3), 2 * (x[2] - 1)])
    def hessian(self, x, psd=False): return 2 * sp.identity(3, format="csc")
    def solutions(self): return [np.array([-2.0, 3.0, 1.0])]


class Beale(Base):
    """tests/test_nonlinear_solver.cpp:207-241 Beale, minimum (3, 0.5)."""
    n = 2
    def _t(self, x): return [c - x[0] + x[0] * x[1] ** k for k, c in ((1, 1.5), (2, 2.25), (3, 2.625))]
    def value(self, x): return float(sum(t * t for t in self._t(x)))
    def gradient(self, x):
        t = self._t(x)
        d0 = [-1 + x[1] ** k for k in (1, 2, 3)]
        d1 = [k * x[0] * x[1] ** (k - 1) for k in (1, 2, 3)]
        return np.array([2 * sum(a * b for a, b in zip(t, d0)), 2 * sum(a * b for a, b in zip(t, d1))])
    def hessian(self, x, psd=False):
        t = self._t(x)
        d0 = [-1 + x[1] ** k for k in (1, 2, 3)]
        d1 = [k * x[0] * x[1] ** (k - 1) for k in (1, 2, 3)]
        d01 = [k * x[1] ** (k - 1) for k in (1, 2, 3)]
        d11 = [k * (k - 1) * x[0] * x[1] ** (k - 2) if k > 1 else 0.0 for k in (1, 2, 3)]
        h00 = 2 * sum(a * a for a in d0)
        h01 = 2 * sum(a * b + c * d for a, b, c, d in zip(d0, d1, t, d01))
        h11 = 2 * sum(b * b + c * d for b, c, d in zip(d1, t, d11))
        return sp.csc_matrix(np.array([[h00, h01], [h01, h11]]))
    def solutions(self): return [np.array([3.0, 0.5])]


def test_reference_problem_derivatives():
    """the hand-written derivatives of the replayed problems against central differences"""
    rng = np.random.default_rng(4)
    for prob in (RefQuadratic(), Beale(), Rosenbrock(10), Sphere(10)):
        x = rng.uniform(-1, 1, prob.n)
        g = prob.gradient(x)
        H = prob.hessian(x).toarray()
        for i in range(prob.n):
            e = np.zeros(prob.n)
            e[i] = 1e-6
            assert abs((prob.value(x + e) - prob.value(x - e)) / 2e-6 - g[i]) < 1e-5 * max(1, abs(g[i]))
            assert np.abs((prob.gradient(x + e) - prob.gradient(x - e)) / 2e-6 - H[:, i]).max() < 1e-4 * max(1, np.abs(H).max())


@pytest.mark.filterwarnings("ignore::RuntimeWarning")
@pytest.mark.parametrize("method", ["Armijo", "RobustArmijo", "Backtracking", "ResidualBacktracking", "None"])
def test_reference_nonlinear_test_replayed_with_gradient_descent_on_cpu(psb, method):
    """The reference's TEST_CASE("nonlinear") (tests/test_nonlinear_solver.cpp:262-349,422-426) for the solver that needs no
    GPU: every problem x every available line search from x = 0 with max_iterations 1000, rel_grad_norm_tol 0; an
    exception (no convergence) is tolerated exactly as there, but a run that RETURNS must satisfy the reference's check
    min_sol |x - sol| < 1e-7 or |grad| < 1e-7 -- and the driver must agree with the restatement on which runs return."""
    from oracle import newton_oracle as NO
    P = {"solver": "GradientDescent", "line_search": {"method": method}, "max_iterations": 1000, "rel_grad_norm_tol": 0}
    returned = 0
    for prob in (RefQuadratic(), Rosenbrock(10), Sphere(10), Beale()):
        x = np.zeros(prob.n)
        s = psb.NonlinearSolver.create(P, {"solver": "CUDA"})
        try:
            s.minimize(prob, x)
            ok = True
        except RuntimeError:
            ok = False
        try:
            xo, _ = NO.minimize(prob, np.zeros(prob.n), P, direct)
            oko = True
        except RuntimeError:
            oko = False
        assert ok == oko
        if ok:
            returned += 1
            assert np.array_equal(x, xo)
            err = min(np.linalg.norm(x - sol) for sol in prob.solutions())
            if err >= 1e-7:
                err = np.linalg.norm(prob.gradient(x))
            assert err < 1e-7, (type(prob).__name__, method, err)
    assert returned >= (0 if method == "None" else 2)     # the quadratic bowls converge with every real line search


def test_reference_iteration_callback_test_replayed_on_cpu(psb):
    """The reference's TEST_CASE("iteration-callback") (tests/test_nonlinear_solver.cpp:714-754) with GradientDescent in
    place of Newton (the strategy that runs without a GPU; the callback sits in the shared outer loop): it fires every
    iteration with a finite positive line-search alpha, asks to stop at iterations >= 2, the solver returns WITHOUT
    throwing (ObjectiveCustomStop) before it has converged; driver == restatement."""
    from oracle import newton_oracle as NO
    P = {"solver": "GradientDescent", "line_search": {"method": "Backtracking"}, "max_iterations": 100, "rel_grad_norm_tol": 0}
    prob = Rosenbrock(10)
    calls, calls_o = [], []
    s = psb.NonlinearSolver.create(P, {"solver": "CUDA"})
    s.set_iteration_callback(lambda st: (calls.append(dict(st)), st["iterations"] >= 2)[1])
    x = np.zeros(prob.n)
    s.minimize(prob, x)                                            # REQUIRE_NOTHROW
    assert 1 <= len(calls) <= 4 and all(np.isfinite(c["alpha"]) and c["alpha"] > 0 for c in calls)
    assert np.linalg.norm(prob.gradient(x)) > 1e-7                 # stopped before converging
    assert s.get_info()["status"] == "Objective function specified to stop"
    xo, io = NO.minimize(prob, np.zeros(prob.n), P, direct, iteration_callback=lambda st: (calls_o.append(dict(st)), st["iterations"] >= 2)[1])
    assert io["status"] == "ObjectiveCustomStop" and np.array_equal(x, xo) and len(calls) == len(calls_o)
    assert [c["alpha"] for c in calls] == [c["alpha"] for c in calls_o] and [c["iterations"] for c in calls] == [c["iterations"] for c in calls_o]
    s.set_iteration_callback(None)                                 # removed: runs to the iteration limit, which throws
    with pytest.raises(RuntimeError, match="Reached iteration limit"):
        s.minimize(prob, np.zeros(prob.n))


def test_direction_filter_matches_oracle_on_cpu(psb):
    """Solver::set_direction_filter (Solver.hpp:80-86; Solver.cpp:353-358,392-403): a filter that pins the first three
    dofs. They never move, descent is measured on the filtered gradient, driver == restatement bit for bit."""
    from oracle import newton_oracle as NO
    P = {"solver": "GradientDescent", "line_search": {"method": "Backtracking"}, "max_iterations": 400, "grad_norm_tol": 1e-9,
         "rel_grad_norm_tol": 0, "x_delta_tol": 1e-6, "allow_non_grad_convergence": True}
    prob = Quadratic(12)
    x0 = np.random.default_rng(3).uniform(-1, 1, 12)

    def pin(x, dx):
        dx[:3] = 0.0
    s = psb.NonlinearSolver.create(P, {"solver": "CUDA"})
    s.set_direction_filter(pin)
    x = x0.copy()
    s.minimize(prob, x)
    xo, io = NO.minimize(prob, x0.copy(), P, direct, direction_filter=pin)
    assert np.array_equal(x[:3], x0[:3]) and np.array_equal(x, xo) and s.get_info()["iterations"] == io["iterations"]
    assert np.abs(prob.gradient(x)[3:]).max() < 1e-5               # minimised over the free dofs
    assert s.get_info()["status"] == _STATUS_TEXT[io["status"]]


class HugeOffset(Base):
    """f = C + 0.5 |x - 1|^2 with C = 1e17: the energy difference of a good step drowns in the rounding error of C, so plain
    Armijo (Armijo.cpp:20-32) rejects every step size while RobustArmijo's gradient-based estimate (RobustArmijo.cpp:30-44)
    accepts the full Newton step."""
    def __init__(self, n=6): self.n = n
    def value(self, x): return 1e17 + 0.5 * float((x - 1) @ (x - 1))
    def gradient(self, x): return x - 1
    def hessian(self, x, psd=False): return sp.identity(self.n, format="csc")
    def solutions(self): return [np.ones(self.n)]


def test_robust_armijo_oracle_accepts_steps_below_rounding_error():
    from oracle import newton_oracle as NO
    prob = HugeOffset()
    x0 = np.full(prob.n, 1.0 + 1e-3)
    ra = dict(PARAMS, line_search={"method": "RobustArmijo"})
    x, info = NO.minimize(prob, x0, ra, direct)
    assert info["status"] == "GradNormTolerance" and np.abs(x - 1).max() < 1e-7
    # the default line search of the reference is RobustArmijo (nonlinear-solver-spec.json:603-604)
    assert NO.LineSearch({}).method == "RobustArmijo"
    # plain Armijo cannot see a decrease of 3e-6 next to 1e17 when the rounding goes the wrong way: perturb so that it does
    ls = NO.LineSearch(ra)
    class Noisy(HugeOffset):
        def value(self, x): return 1e17 + (16.0 if np.abs(x - x0).max() > 0 else 0.0)  # every trial point "rises" by one ulp(1e17)
    step = ls.line_search(x0, -(x0 - 1), Noisy())
    # full step: dE_approx + eps = -|g|^2/2 + |g|^2/2 = 0 is not a sufficient decrease; half step: (-3/8 + 1/8)|g|^2 is
    assert step == 0.5
    ls_plain = NO.LineSearch(dict(PARAMS, line_search={"method": "Armijo"}))
    assert np.isnan(ls_plain.line_search(x0, -(x0 - 1), Noisy()))


@pytest.mark.gpu
def test_robust_armijo_driver_matches_oracle(psb):
    """The driver's RobustArmijo takes the same branch as the restatement on the rounding-error problem."""
    from oracle import newton_oracle as NO
    prob = HugeOffset()
    x0 = np.full(prob.n, 1.0 + 1e-3)
    ra = dict(PARAMS, line_search={"method": "RobustArmijo"})
    xo, io = NO.minimize(prob, x0, ra, direct)
    x = x0.copy()
    s = psb.NonlinearSolver.create(ra, _lin())
    s.minimize(prob, x)
    info = s.get_info()
    assert info["succeeded"] and info["iterations"] == io["iterations"] and info["line_search"] == "RobustArmijo"
    assert np.abs(x - xo).max() < 1e-10


def test_nl_create_rejects_unknown_solver(psb):
    with pytest.raises(RuntimeError, match="Unrecognized solver type"):
        psb.NonlinearSolver.create({"solver": "L-BFGS-B"}, {"solver": "CUDA"})
    with pytest.raises(RuntimeError, match="Unknown line search"):
        psb.NonlinearSolver.create({"solver": "Newton", "line_search": {"method": "Wolfe"}}, {"solver": "CUDA"})


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_newton_fails_loudly_without_gpu(psb):
    """No silent fallback to GradientDescent when the device is missing."""
    s = psb.NonlinearSolver.create(PARAMS, {"solver": "CUDA"})
    with pytest.raises(RuntimeError, match="(?i)cuda|device"):
        s.minimize(Sphere(10), np.full(10, 0.5))


def _lin(tol=1e-12, precond="jacobi"):
    return {"solver": "CUDA", "CUDA": {"tolerance": tol, "max_iter": 2000, "precond": precond}}


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["Backtracking", "Armijo", "RobustArmijo"])
def test_reference_problems_on_gpu(psb, method):
    """tests/test_nonlinear_solver.cpp:422-426 for the Newton chain: every start converges; indefinite Rosenbrock Hessians
    exercise the Newton -> ProjectedNewton -> RegularizedNewton fallback (Newton.cpp:156-162, Solver.cpp:375-394)."""
    rng = np.random.default_rng(1)
    p = dict(PARAMS, line_search={"method": method})
    for prob in (Rosenbrock(10), Sphere(10), Quadratic(12)):
        for _ in range(3):
            x = rng.uniform(-1, 1, prob.n)
            s = psb.NonlinearSolver.create(p, _lin())
            s.minimize(prob, x)
            info = s.get_info()
            assert info["succeeded"], info
            assert min(np.linalg.norm(x - sol) for sol in prob.solutions()) < 1e-7 or info["grad_norm"] < 1e-7
            assert info["line_search"] == method
            assert all("solver_iter" in i and "num_iterations" in i for i in info["internal_solver"])


@pytest.mark.gpu
def test_driver_matches_oracle_step_for_step(psb, orc):
    """Same Newton iterations, same inner CG iteration counts and the same minimiser as the restatement driven by the
    oracle's Eigen-faithful Jacobi-PCG."""
    from oracle import newton_oracle as NO
    P = psb.problems
    prob = P.QuarticSpringGrid3D(12, kappa=10.0)
    x0 = 0.2 * P.splitmix64(9, prob.N)

    def cg(H, rhs, guess):
        H = H.tocsc()
        H.sort_indices()
        x, it, err, _ = orc.eigen_cg(H.indptr.astype(np.int32), H.indices.astype(np.int32), H.data, rhs, x0=guess, tol=1e-10, max_iters=2000)
        return x, it

    xo, io = NO.minimize(prob, x0, PARAMS, cg)
    x = x0.copy()
    s = psb.NonlinearSolver.create(PARAMS, _lin(1e-10))
    s.minimize(prob, x)
    info = s.get_info()
    assert info["status"] == "Gradient vector norm too small" and io["status"] == "GradNormTolerance"
    assert info["iterations"] == io["iterations"]
    gi = [i["solver_iter"] for i in info["internal_solver"]]
    assert len(gi) == len(io["linear_iterations"])
    assert all(abs(a - b) <= 1 for a, b in zip(gi, io["linear_iterations"])), (gi, io["linear_iterations"])
    assert np.abs(x - xo).max() < 1e-9
    # the pattern is analysed once, later Newton steps hit the hash (Newton.cpp:189 calls analyze_pattern every step)
    assert [i["analyze_skipped"] for i in info["internal_solver"]][1:] == [True] * (len(gi) - 1)


@pytest.mark.gpu
def test_newton_amg_million_dof(psb):
    """Config 5 shape on one GPU: 1,000,000-DoF nonlinear problem, inner solve = GPU SA-AMG-PCG."""
    P = psb.problems
    prob = P.QuarticSpringGrid3D(100, kappa=10.0)
    x = np.zeros(prob.N)
    p = dict(PARAMS, grad_norm_tol=1e-8)
    s = psb.NonlinearSolver.create(p, {"solver": "CUDA", "CUDA": {"precond": "amg", "tolerance": 1e-10, "max_iter": 200}})
    s.minimize(prob, x)
    info = s.get_info()
    assert info["succeeded"] and info["grad_norm"] < 1e-8, info
    assert np.linalg.norm(prob.gradient(x)) < 1e-8
    assert 2 <= info["iterations"] <= 30
    assert all(i["precond"] == "amg" and i["solver_status"] == "Converged" for i in info["internal_solver"])


# ------------------------------------------------------------------------------------------ L-BFGS
def test_lbfgs_oracle_two_loop_equals_explicit_bfgs_matrix():
    """The restated BFGSMat (LBFGSpp, un-vendored) against the textbook definition it implements: with H0 = I / theta,
    H_{k+1} = (I - rho s y^T) H_k (I - rho y s^T) + rho s s^T over the stored pairs, oldest first."""
    from oracle import newton_oracle as NO
    rng = np.random.default_rng(2)
    n, m = 9, 4
    B = NO.BFGSMat(n, m)
    pairs = []
    for k in range(7):  # more corrections than the memory holds: the ring wraps
        s = rng.standard_normal(n)
        y = s * rng.uniform(0.5, 2.0, n) + 0.05 * rng.standard_normal(n)
        B.add_correction(s, y)
        pairs = (pairs + [(s, y)])[-m:]
        H = np.eye(n) / B.theta
        for s_, y_ in pairs:
            rho = 1.0 / (s_ @ y_)
            V = np.eye(n) - rho * np.outer(y_, s_)
            H = V.T @ H @ V + rho * np.outer(s_, s_)
        v = rng.standard_normal(n)
        np.testing.assert_allclose(B.apply_Hv(v, -1.0), -H @ v, rtol=1e-11, atol=1e-12)


def test_lbfgs_oracle_converges_on_reference_problems():
    """tests/test_nonlinear_solver.cpp:422-426 with solver = L-BFGS (the restatement alone, no GPU)."""
    from oracle import newton_oracle as NO
    rng = np.random.default_rng(3)
    p = dict(PARAMS, solver="L-BFGS", max_iterations=2000)
    for prob in (Rosenbrock(10), Sphere(10), Quadratic(12)):
        x, info = NO.minimize(prob, rng.uniform(-1, 1, prob.n), p, direct)
        assert min(np.linalg.norm(x - sol) for sol in prob.solutions()) < 1e-6 or info["grad_norm"] < 1e-7, (type(prob), info)


@pytest.mark.gpu
@pytest.mark.parametrize("n,m", [(1, 3), (10, 6), (1003, 4), (40000, 6)])
def test_lbfgs_device_direction_matches_oracle(psb, n, m):
    """Every direction of a sequence longer than the memory (ring wrap-around), sizes that are not multiples of the
    vector width, reset, and the device-pointer entry: equal to the oracle to rounding (different summation order)."""
    import torch
    from oracle import newton_oracle as NO
    rng = np.random.default_rng(n + m)
    L = psb.Lbfgs(n, m)
    O = NO.LbfgsStrategy(m)
    A = rng.uniform(0.5, 2.0, n)           # gradients of a convex quadratic keep s.y > 0
    x = rng.standard_normal(n)
    for k in range(2 * m + 3):
        if k == m + 2:
            L.reset()
            O.reset()
        g = A * x + 0.01 * np.sin(x)
        d0 = O.direction(x, g)
        d = L.compute_update_direction(x, g) if k % 2 == 0 else None
        if d is None:
            dx, dg, dd = torch.from_numpy(x).cuda(), torch.from_numpy(g).cuda(), torch.zeros(n, dtype=torch.float64, device="cuda")
            torch.cuda.synchronize()
            L.compute_update_direction_device(dx.data_ptr(), dg.data_ptr(), dd.data_ptr())
            d = dd.cpu().numpy()
        assert np.linalg.norm(d - d0) <= 1e-11 * np.linalg.norm(d0), (k, np.linalg.norm(d - d0), np.linalg.norm(d0))
        assert d @ g < 0
        x = x + 0.5 * d
    with pytest.raises(RuntimeError):
        L.compute_update_direction(np.zeros(n + 1), np.zeros(n + 1))
    with pytest.raises(RuntimeError):
        psb.Lbfgs(5, 0)   # LBFGS.cpp:17-18: history_size must be >= 1


@pytest.mark.gpu
def test_lbfgs_driver_matches_oracle(psb):
    """solver = "L-BFGS" through the nonlinear driver: strategy chain [L-BFGS, GradientDescent] (Solver.cpp:83-85,175-181),
    same iteration count and minimiser as the restatement; the reference problems converge from random starts."""
    from oracle import newton_oracle as NO
    rng = np.random.default_rng(4)
    p = dict(PARAMS, solver="L-BFGS", max_iterations=2000)
    p["L-BFGS"] = {"history_size": 5}
    for prob in (Quadratic(12), Sphere(10), Rosenbrock(10)):
        x0 = rng.uniform(-1, 1, prob.n)
        xo, io = NO.minimize(prob, x0, p, direct)
        x = x0.copy()
        s = psb.NonlinearSolver.create(p, _lin())
        s.minimize(prob, x)
        info = s.get_info()
        assert info["succeeded"], info
        assert info["solver"] == "L-BFGS"
        assert min(np.linalg.norm(x - sol) for sol in prob.solutions()) < 1e-6 or info["grad_norm"] < 1e-7
        if not isinstance(prob, Rosenbrock):  # chaotic line-search path on Rosenbrock: only convergence is compared
            assert abs(info["iterations"] - io["iterations"]) <= 1
            assert np.abs(x - xo).max() < 1e-8
