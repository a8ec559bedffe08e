// The opaque handle type of the C ABI.
#pragma once
#include "solver.hpp"

struct psb200_solver
{
    psb::Solver s;
};
