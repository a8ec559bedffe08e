"""Writes adapter/cuda-solver-spec.json: the jse rules a maintainer appends to the reference's linear-solver-spec.json
(next to the /MAS block, linear-solver-spec.json:456-509) so that strict validation accepts params["CUDA"].
Names and defaults of /CUDA/amg/* are those of /AMGCL/precond/* (linear-solver-spec.json:294-454, AMGCL.cpp:32-65)."""
import json
import os

R = []


def obj(ptr, optional, doc):
    R.append({"pointer": ptr, "default": None, "type": "object", "optional": optional, "doc": doc})


def leaf(ptr, default, typ, doc, options=None, **kw):
    r = {"pointer": ptr, "default": default, "type": typ}
    if options is not None:
        r["options"] = options
    r.update(kw)
    r["doc"] = doc
    R.append(r)


obj("/CUDA", ["krylov", "precond", "tolerance", "max_iter", "check_every", "use_graph", "cg_kernel", "spmv_kernel", "block_size",
              "device", "interior_first", "pdl", "profile", "verify_pattern", "comm_timeout_s", "amg"],
    "Settings for the B200 CUDA solver (libpsb200).")
leaf("/CUDA/krylov", "cg", "string", "Krylov method: cg = Eigen::ConjugateGradient ordering (AMGCL's cg with precond amg), "
     "cg1r = single-reduction (Chronopoulos-Gear) CG for multi-GPU runs, bicgstab = Eigen::BiCGSTAB ordering.",
     options=["cg", "cg1r", "bicgstab"])
leaf("/CUDA/precond", "jacobi", "string", "Preconditioner: Jacobi (Eigen::DiagonalPreconditioner), smoothed-aggregation AMG "
     "(AMGCL defaults) or none (Eigen::IdentityPreconditioner).", options=["jacobi", "amg", "none"])
leaf("/CUDA/tolerance", 1e-12, "float", "Convergence tolerance relative to ||b|| (as /Eigen::ConjugateGradient/tolerance).")
leaf("/CUDA/max_iter", 1000, "int", "Maximum number of iterations.")
leaf("/CUDA/check_every", 16, "int", "Iterations per CUDA-graph batch between host polls of the device-side stop flag.", min=1)
leaf("/CUDA/use_graph", True, "bool", "Replay each batch of iterations as one CUDA graph.")
leaf("/CUDA/cg_kernel", "auto", "string", "Jacobi-PCG schedule: one kernel per phase (the persistent cooperative kernel of round 1 was removed).",
     options=["auto", "split"])
leaf("/CUDA/spmv_kernel", "auto", "string", "SpMV schedule: auto | stream | stream<2|4|8|16> | vector<1..32> | scalar | bsr.")
leaf("/CUDA/block_size", 1, "int", "Block size of vector-valued problems (AMGCL_Block<B>, AMGCL.cpp:111-123).", options=[1, 2, 3])
leaf("/CUDA/device", -1, "int", "CUDA device ordinal; -1 = the current device.")
leaf("/CUDA/interior_first", False, "bool", "Row partitions: multiply tiles without halo columns first.")
leaf("/CUDA/pdl", False, "bool", "Programmatic dependent launch between the kernels of the Krylov chain.")
leaf("/CUDA/profile", False, "bool", "Time every kernel with CUDA events and report the totals in get_info.")
leaf("/CUDA/verify_pattern", True, "bool", "factorize() re-hashes the index arrays to detect a silently changed pattern.")
leaf("/CUDA/comm_timeout_s", 3.0, "float", "Row partitions: seconds a kernel waits for a peer GPU before the solve fails.", min=0)
obj("/CUDA/amg", ["max_levels", "coarse_enough", "direct_coarse", "ncycle", "npre", "npost", "pre_cycles", "aggregation", "dist_mode",
                  "replicate_below", "fused_push", "relax", "coarsening"], "SA-AMG preconditioner settings; mirrors /AMGCL/precond.")
leaf("/CUDA/amg/max_levels", 6, "int", "Maximum number of levels.")
leaf("/CUDA/amg/coarse_enough", 3000, "int", "Stop coarsening below this many rows (AMGCL default 3000 / block size).")
leaf("/CUDA/amg/direct_coarse", False, "bool", "Use a direct solver for the coarsest level.")
leaf("/CUDA/amg/ncycle", 2, "int", "Number of cycles (1 = V, 2 = W).")
leaf("/CUDA/amg/npre", 1, "int", "Pre-relaxations.")
leaf("/CUDA/amg/npost", 1, "int", "Post-relaxations.")
leaf("/CUDA/amg/pre_cycles", 1, "int", "Cycles per preconditioner application.")
leaf("/CUDA/amg/aggregation", "mis2", "string", "Aggregation algorithm (deterministic parallel MIS-2).", options=["mis2"])
leaf("/CUDA/amg/dist_mode", "partitioned", "string", "Row partitions: partitioned = every level above replicate_below is "
     "row-partitioned (decoupled aggregation, distributed Galerkin product); global = one hierarchy of the whole matrix on every "
     "rank with only level 0 partitioned; local = rank-local hierarchy of the diagonal block (block-Jacobi).",
     options=["partitioned", "global", "local"])
leaf("/CUDA/amg/replicate_below", 8000000, "int", "Row partitions: levels with fewer stored non-zeros (all ranks together) are replicated on every rank.")
leaf("/CUDA/amg/fused_push", False, "bool", "Row partitions: push the halo of the smoother iterates from the SpMV epilogue (boundary tiles first).")
obj("/CUDA/amg/relax", ["type", "degree", "power_iters", "higher", "lower", "scale", "damping"], "Smoother settings.")
leaf("/CUDA/amg/relax/type", "chebyshev", "string", "Type of relaxation to use.", options=["chebyshev", "damped_jacobi"])
leaf("/CUDA/amg/relax/degree", 16, "int", "Degree of the polynomial.")
leaf("/CUDA/amg/relax/power_iters", 100, "int", "Number of power iterations.")
leaf("/CUDA/amg/relax/higher", 2, "float", "Higher level relaxation.")
leaf("/CUDA/amg/relax/lower", 0.008333333333, "float", "Lower level relaxation.")
leaf("/CUDA/amg/relax/scale", True, "bool", "Scale.")
leaf("/CUDA/amg/relax/damping", 0.72, "float", "Damping of damped_jacobi.")
obj("/CUDA/amg/coarsening", ["relax", "estimate_spectral_radius", "aggr"], "Coarsening parameters (smoothed aggregation).")
leaf("/CUDA/amg/coarsening/relax", 1, "float", "Coarsening relaxation.")
leaf("/CUDA/amg/coarsening/estimate_spectral_radius", True, "bool", "Should the spectral radius be estimated.")
obj("/CUDA/amg/coarsening/aggr", ["eps_strong"], "Aggregation settings.")
leaf("/CUDA/amg/coarsening/aggr/eps_strong", 0, "float", "Aggregation epsilon strong.")

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "adapter", "cuda-solver-spec.json")
json.dump(R, open(out, "w"), indent=4)
print(out, len(R), "rules")
