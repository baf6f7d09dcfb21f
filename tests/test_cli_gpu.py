"""The command line (test_spmv) and the Python-3 driver keep the reference's contract:
same argv, the `m:` / `Average` lines run_test.py scrapes, Y/Y pass columns.  Also runs the
reference's UNMODIFIED harness linked against libsblas_spmv.so when oracle/_ref has it."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "test_spmv")
REFH = os.path.join(ROOT, "oracle", "_ref", "test_spmv_refharness")


def write_mtx(path, qh768):
    with open(path, "w") as fh:
        fh.write("%%MatrixMarket matrix coordinate real general\n% test\n")
        fh.write("%d %d %d\n" % (qh768["m"], qh768["n"], qh768["nnz"]))
        for r, c, v in zip(qh768["row"], qh768["col"], qh768["val"]):
            fh.write("%d %d %s\n" % (r + 1, c + 1, repr(float(v))))


def run(cmd):
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    return p.returncode, p.stdout


def check_output(out, m, n, nnz, repeats, ngpu, kernel):
    sys.path.insert(0, ROOT)
    import run_test
    lines = out.strip().split("\n")
    assert lines[0] == "Using %d GPU(s)." % ngpu
    assert lines[1] == "Kernel #%d is selected." % kernel
    assert "m: %d n: %d nnz: %d" % (m, n, nnz) in lines
    assert any(l.startswith("Matrix space size: ") and l.endswith(" GB.") for l in lines)
    assert "Warming up GPU(s)..." in lines and "Starting tests..." in lines
    assert "  Test No.   Baseline    Version 1     Pass     Version 2     Pass" in lines
    pm, pn, pnnz, t1, t2, t3 = run_test.parse_spmv(out)          # the reference's scraper
    assert (pm, pn, pnnz) == (m, n, nnz)
    assert t1 > 0 and t2 > 0 and t3 > 0
    i0 = lines.index("=" * 71)
    i1 = lines.index("." * 71)
    rows = lines[i0 + 1:i1]                      # one line per repeat (setw columns may touch, as in the reference)
    assert len(rows) == repeats
    for r in rows:
        tok = r.split()
        assert tok.count("Y") == 2 and "N" not in tok and "Failed" not in r, r


def test_cli_file_mode(qh768, tmp_path):
    mtx = str(tmp_path / "qh768.mtx")
    write_mtx(mtx, qh768)
    for kernel in (1, 2, 3):
        rc, out = run([CLI, "f", mtx, "1", "2", str(kernel), "f"])
        assert rc == 0, out
        check_output(out, 768, 768, 2934, 2, 1, kernel)
        assert ("Loading input matrix from " + mtx) in out


def test_cli_correct_ingest_opt_in(tmp_path):
    """SBLAS_INGEST=csr: a SYMMETRIC, unsorted file is expanded and bucketed by
    row; the m:/nnz line shows the expanded count and all three entry points agree (Y Y)."""
    rng = np.random.default_rng(9)
    m = 300
    ents = {}
    for _ in range(4000):
        i, j = sorted((int(rng.integers(0, m)), int(rng.integers(0, m))), reverse=True)
        ents[(i, j)] = float(rng.uniform(0.5, 1.5))
    items = list(ents.items())
    rng.shuffle(items)
    mtx = str(tmp_path / "sym.mtx")
    with open(mtx, "w") as fh:
        fh.write("%%MatrixMarket matrix coordinate real symmetric\n")
        fh.write("%d %d %d\n" % (m, m, len(items)))
        for (i, j), v in items:
            fh.write("%d %d %r\n" % (i + 1, j + 1, v))
    expanded = sum(2 if i != j else 1 for (i, j), _ in items)
    p = subprocess.run([CLI, "f", mtx, "1", "2", "1", "f"], capture_output=True, text=True, timeout=300, cwd=ROOT,
                       env=dict(os.environ, SBLAS_INGEST="csr"))
    assert p.returncode == 0, p.stdout
    check_output(p.stdout, m, m, expanded, 2, 1, 1)


def test_cli_generator_mode():
    rc, out = run([CLI, "g", "200", "1", "1", "1"])          # spmv/INSTALL.md:78
    assert rc == 0, out
    check_output(out, 200, 200, 4850, 1, 1, 1)
    assert "Start generating data ........" in out and "Done generating data." in out
    rc, out = run([CLI, "g", "10000", "1", "2", "2"])        # spmv/test/batch_test.sh size
    assert rc == 0, out
    check_output(out, 10000, 10000, 12125000, 2, 1, 2)


def test_cli_argument_errors():
    rc, out = run([CLI, "g", "200"])
    assert rc != 0 and out.startswith("Incorrect number of arguments!")
    rc, out = run([CLI, "g", "200", "1", "1", "7"])
    assert "The kernel version can only be: 1, 2, or 3." in out
    rc, out = run([CLI, "g", "200", "0", "1", "1"])
    assert "Number of GPU(s) needs to be greater than 0." in out
    rc, out = run([CLI, "g", "200", "99", "1", "1"])
    assert "Not enough number of GPUs" in out


def test_unmodified_reference_harness_links_and_passes(qh768, tmp_path):
    if not os.path.exists(REFH):
        pytest.skip("oracle/_ref/test_spmv_refharness not built (needs the reference checkout at build time)")
    mtx = str(tmp_path / "qh768.mtx")
    write_mtx(mtx, qh768)
    rc, out = run([REFH, "f", mtx, "1", "2", "1", "f"])
    assert rc == 0, out
    check_output(out, 768, 768, 2934, 2, 1, 1)
    rc2, out2 = run([CLI, "f", mtx, "1", "2", "1", "f"])
    # identical text apart from the timings
    # same text apart from the numbers (setw columns may touch, so drop numbers and blanks altogether)
    strip = lambda s: re.sub(r"[0-9.e+\-\s]+", "", s)
    assert [strip(l) for l in out.split("\n")] == [strip(l) for l in out2.split("\n")]
    rc, out = run([REFH, "g", "200", "1", "1", "2"])
    assert rc == 0, out
    check_output(out, 200, 200, 4850, 1, 1, 2)


def test_run_test_py3_driver(tmp_path):
    env = dict(os.environ, SBLAS_NGPUS="1", SBLAS_MTXPATH=str(tmp_path) + "/")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "run_test.py")], capture_output=True, text=True,
                       timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stdout + p.stderr
    rows = [l for l in open(os.path.join(ROOT, "results.csv")).read().strip().split("\n") if l.startswith("spmv")]
    assert len(rows) == 3
    for r, label in zip(rows, ("V1", "V2", "V3")):
        f = [c.strip() for c in r.split(",")]
        assert f[:7] == ["spmv", "qh768.mtx", "1", "768", "768", "2934", label] and float(f[7]) > 0


def test_spmm_cli_contract(qh768, tmp_path):
    """test_spmm: the reference's argv and the three lines run_test.py scrapes (run_test.py:146-175), the PASS
    line of the multi-GPU vs single-GPU comparison (spmm/test/dspmm_baseline_test.cu:540-545)."""
    import torch
    sys.path.insert(0, ROOT)
    import run_test
    mtx = str(tmp_path / "qh768.mtx")
    write_mtx(mtx, qh768)
    cli = os.path.join(ROOT, "test_spmm")
    for ngpu in [g for g in (1, 2, 4, 8) if g <= torch.cuda.device_count()]:
        rc, out = run([cli, mtx, "128", str(ngpu), "1"])
        assert rc == 0, out
        lines = out.strip().split("\n")
        assert lines[0] == "Using %d GPU(s)." % ngpu
        assert "Matrix A -- #row: 768 #col: 768 nnz: 2934" in lines
        assert "Matrix B -- #row: 768 #col: 128 (dense)" in lines
        assert "mgpu check: PASS" in lines
        m, n, k, nnz, t = run_test.parse_spmm(out)
        assert (m, n, k, nnz) == (768, 768, 128, 2934) and t > 0
    rc, out = run([cli, mtx, "128"])
    assert rc != 0 and out.startswith("Usage: ./spmm")
