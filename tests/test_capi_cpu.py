"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/psb200.h
declares, host-side parameter logic works, and -- with no GPU -- compute entry points fail loudly
instead of falling back to anything."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(psb):
    L = psb._lib.lib()
    header = "".join(open(os.path.join(ROOT, "include", h)).read() for h in sorted(os.listdir(os.path.join(ROOT, "include"))) if h.endswith(".h"))
    declared = sorted(set(re.findall(r"\b(psb200_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    assert sorted(psb._lib.SYMBOLS) == declared
    for name in declared:
        assert hasattr(L, name), name


def test_product_never_imports_oracle():
    """The product path must not import, link or execute the oracle (or any CPU fallback)."""
    pat = re.compile(r"import\s+oracle|from\s+oracle|from\s+\.+oracle|liboracle|orc_[a-z]|oracle\.py|oracle/(?!\))")
    files = []
    for d, _, fs in os.walk(os.path.join(ROOT, "polysolve_b200")):
        files += [os.path.join(d, f) for f in fs if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h", "Makefile"))]
    files += [os.path.join(ROOT, f) for f in ("adapter/CUDASolver.cpp", "adapter/CUDASolver.hpp", "include/psb200.h")]
    for p in files:
        if os.path.exists(p):
            src = open(p).read().replace("must not depend on oracle/)", "")
            assert not pat.search(src), p


def test_create_name_and_parameters(psb):
    s = psb.Solver.create("CUDA", "")
    assert s.name() == "CUDA"
    assert not s.is_dense()
    s.set_parameters({"CUDA": {"tolerance": 1e-10, "max_iter": 500, "krylov": "cg", "precond": "jacobi"},
                      "Eigen::ConjugateGradient": {"tolerance": 1.0}})  # other solvers' keys are ignored
    s.set_tolerance(1e-9)
    info = s.get_info()
    for k in ("solver_iter", "solver_error", "num_iterations", "final_res_norm", "solver_status"):
        assert k in info
    with pytest.raises(RuntimeError, match="Unrecognized solver type"):
        psb.Solver.create("Eigen::SimplicialLDLT", "")
    with pytest.raises(RuntimeError, match="unknown krylov"):
        s.set_parameters({"CUDA": {"krylov": "gmres"}})
    with pytest.raises(RuntimeError, match="unknown precond"):
        s.set_parameters({"CUDA": {"precond": "ilu"}})
    with pytest.raises(RuntimeError, match="json parse error"):
        s._check(s._L.psb200_set_parameters(s._h, b"{not json"))


def test_json_factory_priority_list(psb):
    # reference Solver.cpp:92-114: "solver" may be a priority list
    s = psb.Solver.create({"solver": ["Hypre", "CUDA", "Eigen::SimplicialLDLT"], "CUDA": {"tolerance": 1e-9}})
    assert s.name() == "CUDA"


def test_protocol_errors(psb):
    s = psb.Solver.create("CUDA", "")
    b = np.ones(4)
    x = np.zeros(4)
    with pytest.raises(RuntimeError):
        s.solve(b, x)  # no factorize yet (and, on a CPU box, no device)
    with pytest.raises(RuntimeError, match="not compressed|null|no CUDA device"):
        outer = np.array([1, 2, 3, 4, 5], np.int32)
        inner = np.zeros(5, np.int32)
        s.analyze_pattern_raw(4, outer, inner, 4)
    # the Python mirror checks the array lengths against the sizes it passes (the C ABI takes plain pointers)
    with pytest.raises(RuntimeError, match="shorter than outer"):
        s.analyze_pattern_raw(4, np.array([0, 1, 2, 3, 5], np.int32), np.zeros(4, np.int32), 4)
    with pytest.raises(RuntimeError, match="shorter than outer"):
        s.factorize_raw(4, np.array([0, 1, 2, 3, 4], np.int32), np.zeros(4, np.int32), np.zeros(3))
    with pytest.raises(RuntimeError, match="n \\+ 1 column pointers"):
        s.analyze_pattern_raw(4, np.array([0, 1, 2, 3], np.int32), np.zeros(4, np.int32), 4)


def test_malformed_outer_array_is_rejected_before_any_kernel(psb):
    """Round-1 advisor finding: a non-monotone outer array must come back as PSB200_ERR_INVALID from the host-side check,
    not index device memory out of bounds (which would poison the CUDA context)."""
    s = psb.Solver.create("CUDA", "")
    outer = np.array([0, 3, 2, 4, 4], np.int32)   # outer[0] = 0 and outer[n] = nnz hold, but it decreases in between
    inner = np.zeros(4, np.int32)
    with pytest.raises(RuntimeError, match="non-decreasing"):
        s.analyze_pattern_raw(4, outer, inner, 4)
    with pytest.raises(RuntimeError, match="spmv_kernel"):
        s.set_parameters({"CUDA": {"spmv_kernel": "vectorY"}})   # validated before the parameters are committed
    with pytest.raises(RuntimeError, match="comm_timeout_s"):
        s.set_parameters({"CUDA": {"comm_timeout_s": -1}})
    with pytest.raises(RuntimeError, match="cg1r"):
        s.set_parameters({"CUDA": {"krylov": "cg1r", "precond": "amg"}})
    s.set_parameters({"CUDA": {"krylov": "cg1r", "comm_timeout_s": 10.0, "amg": {"dist_mode": "partitioned", "replicate_below": 1000}}})


def test_parameter_text_is_parsed_defensively(psb):
    """The parameter document crosses the C ABI as text (nlohmann dump()): malformed or hostile text must come back as an
    error code, never crash -- truncated / mutated documents, and nesting deep enough to overflow a recursive parser."""
    import json
    import random
    s = psb.Solver.create("CUDA", "")

    def send(txt):
        return s._L.psb200_set_parameters(s._h, txt.encode())

    for depth in (129, 10 ** 4, 10 ** 6):
        assert send("[" * depth) != 0 and "nesting too deep" in s._L.psb200_last_error(s._h).decode()
        assert send('{"a":' * depth) != 0
    assert send(json.dumps({"CUDA": {"amg": {"relax": {"degree": 8}}}})) == 0
    rnd = random.Random(1)

    def doc(d=0):
        k = rnd.random()
        if d > 4 or k < 0.3:
            return rnd.choice([1, 2.5, -3e10, True, False, None, "a\"b\\c\n\t\u00e9", "", 1e308, -0.0])
        if k < 0.65:
            return {rnd.choice(["CUDA", "amg", "relax", "precond", "tolerance", "max_iter", "krylov", "x y"]): doc(d + 1)
                    for _ in range(rnd.randint(0, 4))}
        return [doc(d + 1) for _ in range(rnd.randint(0, 4))]

    for _ in range(400):
        txt = json.dumps(doc())
        if rnd.random() < 0.4 and len(txt) > 2:
            i = rnd.randrange(len(txt))
            txt = txt[:i] + rnd.choice(["", "{", "}", '"', ",", ":", "\\", "[", "]", "tru", "nul", "1e", "-"]) + txt[i + rnd.randint(0, 2):]
        assert send(txt) in (0, 1, 2, 3, 4, 5, 6)      # a status code, and the process is still here
    s.set_parameters({"CUDA": {"tolerance": 1e-9}})    # the handle still works


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu(psb):
    s = psb.Solver.create("CUDA", "")
    o, i, v = psb.problems.poisson2d(4)
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA error"):
        s.factorize_raw(16, o, i, v)


def test_generators_match_oracle(psb, orc):
    P = psb.problems
    for n in (1, 2, 7):
        for a, b in ((P.poisson2d(n), orc.poisson2d(n)), (P.poisson3d(n), orc.poisson3d(n))):
            assert all(np.array_equal(u, w) for u, w in zip(a, b))
    assert np.array_equal(P.splitmix64(42, 257), orc.splitmix64(42, 257))
    assert P.spmv_bytes(10077696, 70263936) == 1044721156
    assert P.pcg_iter_bytes(10077696, 70263936) == 1931558404


def test_headers_are_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: every header under include/ must compile as C99 (no C++ types in signatures)
    and a C translation unit that references every declared function must link against libpsb200.so."""
    import re
    import subprocess
    inc = os.path.join(ROOT, "include")
    headers = sorted(h for h in os.listdir(inc) if h.endswith(".h"))
    text = "".join(open(os.path.join(inc, h)).read() for h in headers)
    names = sorted(set(re.findall(r"\b(psb200_[a-z0-9_]+)\s*\(", text)))
    src = tmp_path / "abi.c"
    src.write_text("".join(f'#include "{h}"\n' for h in headers)
                   + "typedef void (*fn)(void);\nfn table[] = {\n" + "".join(f"    (fn){n},\n" for n in names) + "};\n"
                   + "int main(void) { return table[0] == 0; }\n")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-Wno-pedantic", "-I", inc, "-c", str(src), "-o",
                           str(tmp_path / "abi.o")])
    lib = os.path.join(ROOT, "polysolve_b200", "csrc")
    subprocess.check_call(["gcc", str(tmp_path / "abi.o"), "-L", lib, "-lpsb200", "-Wl,-rpath," + lib, "-o", str(tmp_path / "abi")])
