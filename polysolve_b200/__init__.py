"""polysolve_b200 -- B200-native linear-solver backend for polysolve (hot path only).

`linear.Solver.create("CUDA", "")` mirrors polysolve::linear::Solver::create; everything runs in
hand-written sm_100a CUDA behind the C ABI of include/psb200.h."""
from . import _lib, io, neohookean, problems  # noqa: F401
from .solver import Solver  # noqa: F401
from .nonlinear import Lbfgs, NonlinearSolver, Problem  # noqa: F401
