"""the opt-in correct Matrix-Market -> CSR ingest
(include/sblas_ingest.h) against the oracle's restatement of the reference's own correct loader
(sptrsv/sptrsv_v1/src/mmio_highlevel.h:139-298) and against scipy.io.mmread."""
import os

import numpy as np
import pytest
import scipy.io
import scipy.sparse

import oracle
import sblas_b200 as sb
from conftest import GOLDEN


def write_mtx(path, m, n, entries, field="real", symm="general", comments=2):
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate %s %s\n" % (field, symm))
        for k in range(comments):
            f.write("%% comment %d\n" % k)
        f.write("%d %d %d\n" % (m, n, len(entries)))
        for e in entries:
            if field == "pattern":
                f.write("%d %d\n" % (e[0] + 1, e[1] + 1))
            elif field == "integer":
                f.write("%d %d %d\n" % (e[0] + 1, e[1] + 1, int(e[2])))
            elif field == "complex":
                f.write("%d %d %.17g %.17g\n" % (e[0] + 1, e[1] + 1, e[2], 0.25))
            else:
                f.write("%d %d %.17g\n" % (e[0] + 1, e[1] + 1, e[2]))


def dense(m, n, rp, col, val):
    a = np.zeros((m, n))
    rows = np.repeat(np.arange(m), np.diff(rp))
    np.add.at(a, (rows, col), val)
    return a


CASES = [("real", "general"), ("real", "symmetric"), ("pattern", "general"), ("pattern", "symmetric"),
         ("integer", "general"), ("integer", "symmetric"), ("complex", "hermitian")]


@pytest.mark.parametrize("field,symm", CASES)
def test_ingest_matches_oracle_and_scipy(tmp_path, field, symm):
    rng = np.random.default_rng(hash((field, symm)) % 1000)
    m = n = 37
    ents = {}
    for _ in range(260):
        i, j = int(rng.integers(0, m)), int(rng.integers(0, n))
        if symm != "general" and j > i:
            i, j = j, i                              # lower triangle only, as the format requires
        ents[(i, j)] = float(rng.integers(-9, 10)) if field == "integer" else float(rng.standard_normal())
    entries = [(i, j, v) for (i, j), v in ents.items()]
    rng.shuffle(entries)                             # unsorted file: the case the harness loader gets wrong
    path = str(tmp_path / "a.mtx")
    write_mtx(path, m, n, entries, field, symm)
    gm, gn, rp, col, val, sym = sb.mtx_read_csr(path)
    om, on, orp, ocol, oval, osym = oracle.load_mtx_csr(path)
    assert (gm, gn, sym) == (om, on, osym) == (m, n, symm != "general")
    assert (rp == orp).all() and (col == ocol).all() and (val == oval).all()      # entry for entry
    assert rp[0] == 0 and rp[-1] == len(col) and (np.diff(rp) >= 0).all()
    if field != "complex":                           # scipy conjugates hermitian mirrors; the reference does not
        want = scipy.io.mmread(path).toarray()
        assert np.array_equal(dense(m, n, rp, col, val), want)


def test_ingest_equals_harness_loader_on_row_sorted_general_file(tmp_path):
    """For a general file whose entries are sorted by row the correct loader and the reference
    harness's loader (file order used as CSR) give the same arrays."""
    rng = np.random.default_rng(5)
    m, n = 50, 41
    entries = sorted(((int(i), int(j), float(rng.standard_normal())) for i, j in
                      {(int(rng.integers(0, m)), int(rng.integers(0, n))) for _ in range(300)}), key=lambda e: e[0])
    path = str(tmp_path / "s.mtx")
    write_mtx(path, m, n, entries)
    _, _, rp, col, val, _ = sb.mtx_read_csr(path)
    hm, hn, hr, hc, hv = oracle.load_mtx(path, "f")
    assert (rp == oracle.coo_to_rowptr(hm, hr)).all() and (col == hc).all() and (val == hv).all()


def test_ingest_errors(tmp_path):
    p = str(tmp_path / "bad.mtx")
    open(p, "w").write("not a banner\n1 1 1\n1 1 1.0\n")
    with pytest.raises(IOError):
        sb.mtx_read_csr(p)
    with pytest.raises(IOError):
        sb.mtx_read_csr(str(tmp_path / "missing.mtx"))
    open(p, "w").write("%%MatrixMarket matrix coordinate real general\n3 3 4\n1 1 1.0\n2 2 2.0\n")
    with pytest.raises(IOError):
        sb.mtx_read_csr(p)                           # fewer entries than announced
    open(p, "w").write("%%MatrixMarket matrix coordinate real general\n3 3 1\n4 1 1.0\n")
    with pytest.raises(IOError):
        sb.mtx_read_csr(p)                           # row index out of range


def test_sample_matrix_through_the_correct_loader():
    """qh768 is a `general` file sorted by COLUMN: the harness loader yields a permuted matrix
    (SURVEY F3), the correct loader the matrix scipy reads."""
    path = os.path.join(GOLDEN, "..", "..", "sample_matrix", "qh768.mtx")
    if not os.path.exists(path):
        path = "/root/reference/sample_matrix/qh768.mtx"
    if not os.path.exists(path):
        pytest.skip("sample matrix not present")
    m, n, rp, col, val, sym = sb.mtx_read_csr(path)
    assert (m, n, int(rp[-1]), sym) == (768, 768, 2934, False)
    assert np.array_equal(dense(m, n, rp, col, val), scipy.io.mmread(path).toarray())


GOLDEN_CASES = ["real_general", "real_symmetric", "pattern_general", "pattern_symmetric", "integer_general",
                "integer_symmetric", "complex_hermitian"]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_ingest_matches_the_reference_loader_golden(name):
    """Pinned: the CSR arrays the REFERENCE's own loader (mmio_highlevel.h, compiled by oracle/Makefile)
    produced for the committed fixtures (tests/golden/make_golden_ingest.py) -- entry for entry, for the
    product and for the oracle restatement."""
    g = np.load(os.path.join(GOLDEN, "ingest_expected.npz"))
    path = os.path.join(GOLDEN, "ingest_%s.mtx" % name)
    for loader in (sb.mtx_read_csr, oracle.load_mtx_csr):
        m, n, rp, col, val, sym = loader(path)
        assert [m, n, int(sym)] == g[name + "_mn"].tolist()
        assert (rp == g[name + "_rowptr"]).all() and (col == g[name + "_col"]).all() and (val == g[name + "_val"]).all()


def test_ingest_matches_the_compiled_reference_loader_live(tmp_path):
    """Same comparison against the reference loader itself when oracle/_ref/libref_ingest.so is present
    (it travels with the repo snapshot), on fresh random files."""
    ref = oracle.ref_ingest()
    if ref is None:
        pytest.skip("oracle/_ref/libref_ingest.so not built (needs the reference checkout at build time)")
    for k, (field, symm) in enumerate(CASES):
        rng = np.random.default_rng(500 + k)
        m, n = 61, 47 if symm == "general" else 61
        ents = {}
        for _ in range(500):
            i, j = int(rng.integers(0, m)), int(rng.integers(0, n))
            if symm != "general" and j > i:
                i, j = j, i
            ents[(i, j)] = float(rng.integers(-9, 10)) if field == "integer" else float(rng.standard_normal())
        entries = [(i, j, v) for (i, j), v in ents.items()]
        rng.shuffle(entries)
        path = str(tmp_path / ("r%d.mtx" % k))
        write_mtx(path, m, n, entries, field, symm)
        want = ref(path)
        for loader in (sb.mtx_read_csr, oracle.load_mtx_csr):
            got = loader(path)
            assert got[0:2] == want[0:2] and got[5] == want[5]
            assert (got[2] == want[2]).all() and (got[3] == want[3]).all() and (got[4] == want[4]).all()


def test_symmetric_file_must_be_square(tmp_path):
    """A file that declares itself symmetric with n > m would mirror entries into rows that do not exist:
    rejected with its own code (-7) instead of writing past the row pointer."""
    path = str(tmp_path / "rect_sym.mtx")
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real symmetric\n3 9 2\n1 8 1.5\n3 9 2.5\n")
    with pytest.raises(IOError, match="-7"):
        sb.mtx_read_csr(path)


def test_entry_count_that_wraps_size_t_is_refused(tmp_path):
    """Untrusted size line: 2^62 announced entries make `count * 4` wrap to 0 bytes; the loader must refuse (-6, out of
    memory) instead of allocating a short buffer and writing the file's entries past it."""
    import ctypes as C
    path = str(tmp_path / "huge_count.mtx")
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n4 4 4611686018427387904\n")
        for k in range(64):
            f.write("%d %d 1.0\n" % (k % 4 + 1, (k * 3) % 4 + 1))
    rp = np.zeros(5, np.int64)
    col = np.zeros(64, np.int32)
    val = np.zeros(64)
    rc = sb.lib().sblas_mtx_read_csr(path.encode(), rp.ctypes.data, col.ctypes.data, val.ctypes.data)
    assert rc == -6, rc


def test_mutated_files_never_crash_the_loader(tmp_path):
    """Byte-level mutations of a valid file (truncation, deleted / duplicated / garbled lines, negative and huge
    indices): the loader either reads a matrix or returns an error code; sizes it reports are what it writes."""
    rng = np.random.default_rng(2024)
    base = ["%%MatrixMarket matrix coordinate real symmetric", "% comment", "6 6 7",
            "1 1 2.0", "2 1 -1.0", "3 3 4.5", "5 2 1e-3", "6 6 7", "4 4 1", "6 1 3.25"]
    path = str(tmp_path / "mut.mtx")
    outcomes = {"ok": 0, "err": 0}
    for trial in range(300):
        lines = list(base)
        for _ in range(int(rng.integers(1, 4))):
            k = int(rng.integers(0, len(lines)))
            op = int(rng.integers(0, 6))
            if op == 0 and len(lines) > 1:
                del lines[k]
            elif op == 1:
                lines.insert(k, lines[k])
            elif op == 2:
                lines[k] = lines[k][: int(rng.integers(0, len(lines[k]) + 1))]
            elif op == 3:
                lines[k] = lines[k].replace("1", "-1", 1)
            elif op == 4:
                lines[k] = lines[k].replace("6", "999999999", 1)
            else:
                lines[k] = "".join(chr(int(c)) for c in rng.integers(32, 127, size=12))
        with open(path, "w") as f:
            f.write("\n".join(lines) + ("\n" if trial % 2 else ""))
        try:
            m, n, rp, col, val, sym = sb.mtx_read_csr(path)
        except IOError:
            outcomes["err"] += 1
            continue
        outcomes["ok"] += 1
        assert rp[0] == 0 and (np.diff(rp) >= 0).all() and rp[-1] == len(col) == len(val)
        assert len(col) == 0 or (col.min() >= 0 and col.max() < n)
    assert outcomes["ok"] > 0 and outcomes["err"] > 0, outcomes
