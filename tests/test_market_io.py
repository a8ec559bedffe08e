"""Matrix Market replay (include/psb200_io.h): Eigen::loadMarket / saveMarket and the reference tests'
loadSymmetric (tests/test_linear_solver.cpp:25-50), checked against scipy.io's independent reader / writer and by
round trips. Host-only code: no GPU needed."""
import os

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp


def _rand(n, m, density, seed, sym=False):
    rng = np.random.default_rng(seed)
    A = sp.random(n, m, density=density, random_state=rng, format="csc")
    A.data = rng.standard_normal(A.nnz) * 10.0 ** rng.integers(-12, 12, A.nnz)  # values that need 17 digits
    if sym:
        A = sp.csc_matrix(A + A.T + sp.diags(rng.standard_normal(n)))
    A.sort_indices()
    return A


def _same(A, B):
    A, B = sp.csc_matrix(A), sp.csc_matrix(B)
    A.sort_indices()
    B.sort_indices()
    return A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices) and np.array_equal(A.data, B.data)


def test_load_files_written_by_scipy(psb, tmp_path):
    A = _rand(60, 45, 0.08, 1)
    p = tmp_path / "general.mtx"
    scipy.io.mmwrite(str(p), A, precision=17)
    assert _same(psb.io.load_market(p), A)
    S = _rand(50, 50, 0.05, 2, sym=True)
    p2 = tmp_path / "symmetric.mtx"
    scipy.io.mmwrite(str(p2), S, symmetry="symmetric", precision=17)
    # loadSymmetric mirrors the stored triangle; Eigen::loadMarket keeps the entries as stored
    assert _same(psb.io.load_symmetric(p2), S)
    assert _same(psb.io.load_market(p2, symmetric=-1), S)
    assert _same(psb.io.load_market(p2, symmetric=0), sp.tril(S))
    assert _same(scipy.io.mmread(str(p2)), psb.io.load_symmetric(p2))


def test_save_round_trips_bit_exact_and_scipy_reads_it(psb, tmp_path, orc):
    o, i, v = orc.poisson3d(9)
    A = sp.csc_matrix((v * (1 + 0.3 * orc.splitmix64(3, len(v))), i, o), shape=(729, 729))
    p = tmp_path / "a.mtx"
    psb.io.save_market(A, p)
    assert open(p).readline().strip() == "%%MatrixMarket matrix coordinate real general"   # Eigen::saveMarket's header
    assert _same(psb.io.load_market(p), A)
    assert _same(scipy.io.mmread(str(p)), A)
    S = sp.csc_matrix(A + A.T)
    ps = tmp_path / "s.mtx"
    psb.io.save_market(S, ps, symmetric=True)
    assert _same(psb.io.load_symmetric(ps), S)
    assert _same(scipy.io.mmread(str(ps)), S)
    x = orc.splitmix64(5, 1000) * 1e-7
    pv = tmp_path / "v.mtx"
    psb.io.save_market_vector(x, pv)
    assert np.array_equal(psb.io.load_market_vector(pv), x)
    assert np.array_equal(np.asarray(scipy.io.mmread(str(pv))).ravel(), x)


def test_ragged_inputs(psb, tmp_path):
    p = tmp_path / "ragged.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n% a comment\n\n%another\n4 5 6\n"
                 "1 1 1.5\n4 5 -2e3\n\n2 3 0.25\n% comment between entries\n2 3 0.75\n1 1 1\n3 1 7\n")
    A = psb.io.load_market(p)
    assert A.shape == (4, 5) and A.nnz == 4                     # duplicates are summed (setFromTriplets)
    assert A[0, 0] == 2.5 and A[1, 2] == 1.0 and A[3, 4] == -2000.0 and A[2, 0] == 7.0
    assert np.array_equal(A.indptr, [0, 2, 2, 3, 3, 4])         # empty columns stay empty
    pe = tmp_path / "empty.mtx"
    pe.write_text("%%MatrixMarket matrix coordinate real general\n3 3 0\n")
    E = psb.io.load_market(pe)
    assert E.shape == (3, 3) and E.nnz == 0
    pp = tmp_path / "pattern.mtx"
    pp.write_text("%%MatrixMarket matrix coordinate pattern symmetric\n3 3 2\n2 1\n3 3\n")
    P = psb.io.load_market(pp, symmetric=-1)
    assert np.array_equal(P.toarray(), [[0, 1, 0], [1, 0, 0], [0, 0, 1]])
    # no header at all: the reference's loadSymmetric only skips '%' lines (test_linear_solver.cpp:29-32)
    pn = tmp_path / "nohdr.mtx"
    pn.write_text("2 2 2\n1 1 4\n2 1 -1\n")
    assert np.array_equal(psb.io.load_symmetric(pn).toarray(), [[4, -1], [-1, 0]])


@pytest.mark.parametrize("body,msg", [
    ("%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n", "out of range"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n", "announces"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1\n2 2 1.0\n", "too few fields"),
    ("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1 0\n", "complex"),
    ("%%MatrixMarket matrix array real general\n2 1\n1\n2\n", "coordinate"),
])
def test_malformed_files_fail_loudly(psb, tmp_path, body, msg):
    p = tmp_path / "bad.mtx"
    p.write_text(body)
    with pytest.raises(RuntimeError, match=msg):
        psb.io.load_market(p)
    with pytest.raises(RuntimeError, match="cannot open"):
        psb.io.load_market(tmp_path / "missing.mtx")


@pytest.mark.gpu
def test_replay_fixture_through_the_solver(psb, orc, tmp_path):
    """The reference's test flow (tests/test_linear_solver.cpp:52-100): load a Matrix Market fixture, solve, check
    ||Ax - b|| < 1e-8 -- with a fixture this repository writes itself (polyfem-data is not vendored)."""
    o, i, v = orc.poisson2d(32)
    A = sp.csc_matrix((v, i, o), shape=(1024, 1024))
    psb.io.save_market(sp.tril(A), tmp_path / "A_sym.mtx", symmetric=True)
    B = psb.io.load_symmetric(tmp_path / "A_sym.mtx")
    assert _same(A, B)
    b = orc.splitmix64(42, 1024)
    s = psb.Solver.create("CUDA", "")
    s.set_parameters({"CUDA": {"tolerance": 1e-10}})
    s.analyze_pattern(B, 1024)
    s.factorize(B)
    x = np.zeros(1024)
    s.solve(b, x)
    assert s.get_info()["solver_iter"] == 115      # the C1 known answer
    assert np.linalg.norm(B @ x - b) < 1e-8


def test_randomised_files_match_scipy(psb, tmp_path):
    """150 generated coordinate files -- real / integer / pattern, general / symmetric, shuffled entries, tabs and runs of
    blanks as separators, LF and CRLF line ends, comments and blank lines before the size line, empty matrices, missing
    final newline -- read with header-following symmetry (-1) == scipy.io.mmread; mode 0 == the stored entries;
    mode 1 mirrors whatever the header says."""
    import random
    rnd = random.Random(3)
    rng = np.random.default_rng(3)
    for t in range(150):
        rows = rnd.randint(1, 12)
        cols = rows if rnd.random() < 0.6 else rnd.randint(1, 12)
        A = sp.random(rows, cols, density=rnd.choice([0, 0.05, 0.3, 1.0]), random_state=rng, format="coo")
        field = rnd.choice(["real", "integer", "pattern"])
        sym = rows == cols and rnd.random() < 0.4
        if sym:
            A = sp.tril(A + A.T).tocoo()
        if field == "integer":
            A.data = np.round(A.data * 100) + 1
        if field == "pattern":
            A.data = np.ones_like(A.data)
        lines = [f"%%MatrixMarket matrix coordinate {field} {'symmetric' if sym else 'general'}"]
        if rnd.random() < 0.5:
            lines.append("% a comment")
        if rnd.random() < 0.3:
            lines.append("")
        lines.append(f"{rows} {cols} {A.nnz}")
        ent = list(zip(A.row.tolist(), A.col.tolist(), A.data.tolist()))
        rnd.shuffle(ent)
        for i, j, v in ent:
            s = rnd.choice([" ", "  ", "\t"])
            if field == "pattern":
                lines.append(f"{i + 1}{s}{j + 1}")
            elif field == "integer":
                lines.append(f"{i + 1}{s}{j + 1}{s}{int(v)}")
            else:
                lines.append(f"{i + 1}{s}{j + 1}{s}{v!r}" if rnd.random() < 0.7 else f"{i + 1}{s}{j + 1}{s}{v:.17E}")
        eol = rnd.choice(["\n", "\r\n"])
        path = tmp_path / f"f{t}.mtx"
        with open(path, "w", newline="") as f:
            f.write(eol.join(lines) + (eol if rnd.random() < 0.8 else ""))
        ref = sp.csc_matrix(scipy.io.mmread(str(path)))
        got = psb.io.load_market(path, symmetric=-1)
        assert got.shape == ref.shape and (got.nnz + ref.nnz == 0 or abs(got - ref).max() == 0), t
        stored = sp.csc_matrix((A.data, (A.row, A.col)), shape=(rows, cols))
        got0 = psb.io.load_market(path, symmetric=0)
        assert got0.shape == stored.shape and (got0.nnz + stored.nnz == 0 or abs(got0 - stored).max() == 0), t
        if rows == cols:
            strict = sp.tril(stored, -1)
            got1 = psb.io.load_market(path, symmetric=1)
            want1 = stored + strict.T if sym else stored + sp.csc_matrix((A.data, (A.col, A.row)), shape=(rows, cols)) - sp.diags(stored.diagonal())
            assert got1.nnz + want1.nnz == 0 or abs(got1 - want1).max() <= 1e-15 * max(1.0, abs(want1).max()), t


def test_mutated_files_never_crash_the_reader(psb, tmp_path):
    """600 mutations of a valid file (deleted / inserted / corrupted lines, hostile size lines): every call returns a
    matrix or raises RuntimeError with the reader's message -- no exception crosses the C ABI, no reservation trusts the
    size line (a size line announcing 10^14 entries once ended the process in std::length_error)."""
    import random
    rnd = random.Random(11)
    base = ["%%MatrixMarket matrix coordinate real general", "% c", "4 4 5", "1 1 1.5", "2 2 -2e3", "3 1 4", "4 4 1", "2 3 .5"]
    junk = ["", "%", "0 0 0", "-1 2 3", "5 5 1", "1 1", "1", "a b c", "99999999999 1 1", "1 1 1e999", "2 2 nan",
            "4 4 99999999999999999999", "1 1 1 1 1", "\x00\x01", "3 3 3", "2147483646 2147483646 1", "4 4 99999999999999"]
    path = tmp_path / "m.mtx"
    ok = err = 0
    for _ in range(600):
        lines = list(base)
        for _ in range(rnd.randint(1, 3)):
            k = rnd.randrange(len(lines))
            op = rnd.random()
            if op < 0.2:
                del lines[k]
            elif op < 0.4:
                lines.insert(k, rnd.choice(junk))
            elif lines[k]:
                chars = list(lines[k])
                chars[rnd.randrange(len(chars))] = rnd.choice("0123456789 -.eE%\tx")
                lines[k] = "".join(chars)
            if not lines:
                lines = [""]
        path.write_text("\n".join(lines) + "\n")
        for mode in (0, 1, -1):
            try:
                A = psb.io.load_market(path, symmetric=mode)
                assert A.shape[0] >= 0 and A.nnz <= 10
                ok += 1
            except RuntimeError as e:
                assert "psb200_market_load" in str(e)
                err += 1
        try:
            psb.io.load_market_vector(path)
        except RuntimeError as e:
            assert "psb200_market_load_vector" in str(e)
    assert ok > 50 and err > 50
    path.write_text("%%MatrixMarket matrix coordinate real general\n3 3 99999999999999\n1 1 1\n")
    with pytest.raises(RuntimeError, match="too short"):
        psb.io.load_market(path)
    path.write_text("%%MatrixMarket matrix array real general\n99999999999 1\n1\n")
    with pytest.raises(RuntimeError, match="too short"):
        psb.io.load_market_vector(path)
