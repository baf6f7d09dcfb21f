"""The two command lines up to their first compute call, on the CPU.

test_spmv / test_spmm allocate their host arrays with cudaMallocHost and ask for the device count before they read
the matrix file; a three-function LD_PRELOAD shim (built here with gcc: one device reported, page-locked memory =
malloc) lets the argument checks and the Matrix-Market loaders of the harness run without a GPU.  No compute is
possible behind the shim: the library still answers "no CPU fallback", which the last test checks.  The loaders read
untrusted files: short entry lists and indices outside the matrix must stop the program before the row counters are
indexed with them (the reference harness, spmv/test/dspmv_test.cu:150-183, carries on)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = r"""
#include <stdlib.h>
int cudaGetDeviceCount(int *c) { *c = 1; return 0; }
int cudaMallocHost(void **p, size_t b) { *p = malloc(b ? b : 1); return *p ? 0 : 2; }
int cudaFreeHost(void *p) { free(p); return 0; }
"""
BANNER = "%%MatrixMarket matrix coordinate real general\n"


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    d = tmp_path_factory.mktemp("shim")
    src, so = str(d / "shim.c"), str(d / "shim.so")
    open(src, "w").write(SHIM)
    subprocess.run(["gcc", "-shared", "-fPIC", src, "-o", so], check=True)
    for exe in ("test_spmv", "test_spmm"):
        if not os.path.exists(os.path.join(ROOT, exe)):
            pytest.skip("%s not built (python -c 'import __graft_entry__ as g; g.build()')" % exe)
    return so


def run(shim, *argv):
    env = dict(os.environ, LD_PRELOAD=shim)
    p = subprocess.run([os.path.join(ROOT, argv[0])] + [str(a) for a in argv[1:]], capture_output=True, text=True,
                       env=env, cwd=ROOT, timeout=60)
    return p.returncode, p.stdout


def write(tmp_path, name, body):
    p = str(tmp_path / name)
    open(p, "w").write(body)
    return p


def test_spmv_cli_stops_on_a_short_entry_list(shim, tmp_path):
    p = write(tmp_path, "short.mtx", BANNER + "3 3 4\n1 1 1.0\n2 2 2.0\n")
    rc, out = run(shim, "test_spmv", "f", p, 1, 1, 1, "f")
    assert rc == 1 and "m: 3 n: 3 nnz: 4" in out and "holds 2 of the 4 entries" in out
    assert "Warming up" not in out


@pytest.mark.parametrize("entry", ["7 2 2.0", "2 9 2.0", "0 1 2.0", "2 0 2.0", "-3 1 2.0"])
def test_spmv_cli_stops_on_an_index_outside_the_matrix(shim, tmp_path, entry):
    p = write(tmp_path, "oob.mtx", BANNER + "3 3 2\n1 1 1.0\n" + entry + "\n")
    rc, out = run(shim, "test_spmv", "f", p, 1, 1, 1, "f")
    assert rc == 1 and "lies outside the 3 x 3 matrix" in out and "Warming up" not in out


@pytest.mark.parametrize("size", ["0 3 2", "3 0 2", "3 3 -1", "-4 3 2"])
def test_spmv_cli_refuses_a_bad_size_line(shim, tmp_path, size):
    p = write(tmp_path, "size.mtx", BANNER + size + "\n1 1 1.0\n2 1 2.0\n")
    rc, out = run(shim, "test_spmv", "f", p, 1, 1, 1, "f")
    assert rc == 1 and "size line" in out and "Warming up" not in out


def test_spmv_cli_bad_banner_and_arguments(shim, tmp_path):
    p = write(tmp_path, "banner.mtx", "not a banner\n3 3 1\n1 1 1.0\n")
    rc, out = run(shim, "test_spmv", "f", p, 1, 1, 1, "f")
    assert rc == 1 and "Could not process Matrix Market banner." in out       # dspmv_test.cu:117-120
    rc, out = run(shim, "test_spmv", "f", p, 1, 1)
    assert "Incorrect number of arguments!" in out                            # dspmv_test.cu:63-67
    rc, out = run(shim, "test_spmv", "f", p, 0, 1, 1, "f")
    assert "Number of GPU(s) needs to be greater than 0" in out
    rc, out = run(shim, "test_spmv", "f", p, 1, 1, 4, "f")
    assert "The kernel version can only be: 1, 2, or 3." in out
    rc, out = run(shim, "test_spmv", "f", p, 2, 1, 1, "f")                    # the shim reports one device
    assert "Not enough number of GPUs" in out
    rc, out = run(shim, "test_spmv", "g", 1001, 1, 1, 1)
    assert "n must be a positive multiple of 8" in out


def test_spmv_cli_reads_the_sample_matrix_and_reaches_the_library(shim):
    """A valid file passes the loader checks, and the first compute call then fails loudly: behind the shim there is
    no device, and the library has no CPU path to fall back to."""
    sample = os.path.join(ROOT, "tests", "golden", "qh768_coo.npz")
    import numpy as np
    g = np.load(sample)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "qh768.mtx")
        with open(p, "w") as fh:
            fh.write(BANNER + "%d %d %d\n" % (int(g["m"]), int(g["n"]), len(g["val"])))
            for r, c, v in zip(g["row"], g["col"], g["val"]):
                fh.write("%d %d %s\n" % (r + 1, c + 1, repr(float(v))))
        rc, out = run(shim, "test_spmv", "f", p, 1, 1, 1, "f")
    assert "m: 768 n: 768 nnz: 2934" in out and "Warming up GPU(s)..." in out
    row = [l for l in out.splitlines() if l.strip().startswith("1 ")]
    assert len(row) == 1 and row[0].split()[1:] == ["Failed", "Failed", "N/A", "Failed.", "N/A"], out   # dspmv_test.cu:400-418
    assert [l for l in out.splitlines() if l.strip().startswith("Average")][0].split()[1:] == ["Failed"] * 3


def test_spmm_cli_loader_checks(shim, tmp_path):
    p = write(tmp_path, "oob.mtx", BANNER + "3 3 2\n1 1 1.0\n7 2 2.0\n")
    rc, out = run(shim, "test_spmm", p, 8, 1, 1)
    assert rc == 1 and "Matrix A -- #row: 3 #col: 3 nnz: 2" in out and "i = 1 [6, 1]" in out
    p = write(tmp_path, "short.mtx", BANNER + "3 3 4\n1 1 1.0\n")
    rc, out = run(shim, "test_spmm", p, 8, 1, 1)
    assert rc == 1 and "Could not read Matrix Market format for matrix A." in out
    rc, out = run(shim, "test_spmm", p, 0, 1, 1)
    assert "number of columns of B" in out
    rc, out = run(shim, "test_spmm", p, 8, 2, 1)
    assert "Not enough number of GPUs" in out
    rc, out = run(shim, "test_spmm", p)
    assert "Usage: ./spmm" in out
