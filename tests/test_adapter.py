"""The C++ side of the drop-in boundary, executed: adapter/test_adapter links adapter/CUDASolver.cpp (the
polysolve::linear::Solver subclass a maintainer adds) against libpsb200.so and replays the reference's own
test flow (tests/test_linear_solver.cpp:103-164,400-455) on config 1 -- SURVEY 8 rows a2/a3/b."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "adapter")


def _build():
    subprocess.check_call(["make", "-C", ADAPTER, "test_adapter"], stdout=subprocess.DEVNULL)
    return os.path.join(ADAPTER, "test_adapter")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_adapter_links_and_fails_loudly_without_gpu():
    """Every symbol the adapter uses resolves against libpsb200.so; without a device the first compute call throws
    std::runtime_error naming the missing GPU (no CPU fallback)."""
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device available" in r.stderr and "[CUDA] analyze_pattern" in r.stderr


@pytest.mark.gpu
def test_adapter_solves_config1_through_cpp_boundary():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "solver_iter=115" in r.stdout and "adapter test OK" in r.stdout


def test_spec_fragment_covers_every_parameter():
    """adapter/cuda-solver-spec.json (the rules INTEGRATION.md adds to linear-solver-spec.json) is valid JSON, every rule
    hangs below /CUDA, every parent lists its children under "optional", every key the library's parser reads has a rule,
    and the AMG defaults are the ones polysolve passes to AMGCL (AMGCL.cpp:32-65, linear-solver-spec.json:294-454)."""
    import re
    rules = json.load(open(os.path.join(ADAPTER, "cuda-solver-spec.json")))
    by_ptr = {r["pointer"]: r for r in rules}
    assert "/CUDA" in by_ptr and all(p.startswith("/CUDA") for p in by_ptr)
    for ptr, r in by_ptr.items():
        assert "type" in r and "doc" in r, ptr
        if r["type"] == "object":
            for child in r.get("optional", []):
                assert ptr + "/" + child in by_ptr, (ptr, child)
        if ptr != "/CUDA":
            parent, leaf = ptr.rsplit("/", 1)
            assert leaf in by_ptr[parent]["optional"], ptr
    # keys read by Solver::set_parameters / read_amg
    src = open(os.path.join(ROOT, "polysolve_b200", "csrc", "solver.cu")).read()
    body = src[src.index("static void read_amg"):src.index("// ==================================================================================== analyze_pattern")]
    keys = set(re.findall(r'contains\("([a-z_]+)"\)', body)) | set(re.findall(r'num\([a-z.()"]+, "([a-z_]+)"', body))
    leaves = {p.rsplit("/", 1)[1] for p in by_ptr}
    assert keys - {"CUDA"} <= leaves, sorted(keys - leaves)
    d = lambda p: by_ptr[p]["default"]
    assert d("/CUDA/amg/max_levels") == 6 and d("/CUDA/amg/ncycle") == 2 and d("/CUDA/amg/direct_coarse") is False
    assert d("/CUDA/amg/relax/degree") == 16 and d("/CUDA/amg/relax/power_iters") == 100 and d("/CUDA/amg/relax/type") == "chebyshev"
    assert d("/CUDA/amg/relax/higher") == 2 and abs(d("/CUDA/amg/relax/lower") - 1 / 120) < 1e-9 and d("/CUDA/amg/relax/scale") is True
    assert d("/CUDA/amg/coarsening/relax") == 1 and d("/CUDA/amg/coarsening/estimate_spectral_radius") is True
    assert d("/CUDA/amg/coarsening/aggr/eps_strong") == 0


def test_spec_defaults_are_accepted_by_the_library(psb):
    """A document built from the fragment's defaults passes psb200_set_parameters (no device needed)."""
    rules = json.load(open(os.path.join(ADAPTER, "cuda-solver-spec.json")))
    doc = {}
    for r in sorted(rules, key=lambda r: r["pointer"].count("/")):
        parts = r["pointer"].strip("/").split("/")
        cur = doc
        for p in parts[:-1]:
            cur = cur.setdefault(p, {})
        cur[parts[-1]] = {} if r["type"] == "object" else r["default"]
    s = psb.Solver.create("CUDA", "")
    s.set_parameters(doc)
