import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def qh768():
    """The reference's sample matrix as its harness loads it (file order)."""
    import oracle
    g = np.load(os.path.join(GOLDEN, "qh768_coo.npz"))
    m, n = int(g["m"]), int(g["n"])
    rp = oracle.coo_to_rowptr(m, g["row"])
    return dict(m=m, n=n, nnz=int(rp[-1]), rowptr=rp, col=np.ascontiguousarray(g["col"]),
                val=np.ascontiguousarray(g["val"]), row=np.ascontiguousarray(g["row"]))


def make_csr(rng, m, n, row_len, sort_cols=True, dup=False):
    """Random CSR with the given per-row lengths (int64 rowptr, int32 col, f64 val)."""
    row_len = np.asarray(row_len, np.int64)
    rp = np.zeros(m + 1, np.int64)
    rp[1:] = np.cumsum(row_len)
    nnz = int(rp[-1])
    col = rng.integers(0, n, size=nnz, dtype=np.int64).astype(np.int32)
    if sort_cols:
        for i in range(m):
            col[rp[i]:rp[i + 1]].sort()
    val = rng.uniform(-1.0, 1.0, size=nnz)
    return rp, col, val


def check_tol(y_gpu, y_ref, bound, what=""):
    """BASELINE.json tolerance: |y_gpu - y_oracle| <= 1e-12 * (|alpha| sum|a||x| + |beta||y|) per row."""
    err = np.abs(np.asarray(y_gpu) - np.asarray(y_ref))
    lim = 1e-12 * np.asarray(bound)
    bad = np.nonzero(~(err <= lim))[0]
    assert bad.size == 0, "%s: %d rows out of tolerance, first row %d err %.3e lim %.3e" % (
        what, bad.size, bad[0], err[bad[0]], lim[bad[0]])
