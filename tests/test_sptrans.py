"""Sparse transposition CSR -> CSC (reference sptrans/sptrans_v1/src/sptrans_kernal.h).
Integer / index work: everything is compared BIT FOR BIT.

  not gpu : the oracle (oracle_csr2csc) against the reference's own host transposition compiled where it lies
            (oracle/_ref/libref_sptrans.so, pure host code: sptrans/sptrans_v1/src/tranpose.h), against the
            committed golden arrays generated from it, and against scipy
  gpu     : the library (kernal_sptrans through the C-ABI) against the oracle on every visible GPU count:
            empty rows / columns, duplicates, unsorted rows, rectangular shapes, one hub column, a
            device-generated matrix of ~60 M entries; transposing twice gives the matrix back.
"""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, make_csr


def _mat(rng, m, n, lens, sort_cols=True, dup=False):
    rp64, col, val = make_csr(rng, m, n, lens, sort_cols=sort_cols)
    if dup and len(col) > 10:
        col[5:len(col):7] = col[4:len(col) - 1:7]           # duplicate (row, column) pairs here and there
    return rp64.astype(np.int32), col, val


CASES = {
    "small_mixed": lambda rng: (300, 211, np.concatenate([rng.integers(0, 9, size=290), [150, 0, 0, 90], rng.integers(0, 3, size=6)])),
    "tall": lambda rng: (5000, 37, rng.integers(0, 6, size=5000)),
    "wide_with_hub_column": lambda rng: (64, 100000, rng.integers(1000, 3000, size=64)),
    "empty": lambda rng: (50, 60, np.zeros(50, np.int64)),
    "one_entry": lambda rng: (1, 1, np.array([1], np.int64)),
}


def _build(name):
    rng = np.random.default_rng(700 + sorted(CASES).index(name))
    m, n, lens = CASES[name](rng)
    rp, col, val = _mat(rng, m, n, lens, sort_cols=(name != "small_mixed"), dup=(name == "small_mixed"))
    if name == "wide_with_hub_column":
        col[::3] = 4242                                      # one column holds a third of the matrix
    return m, n, rp, col, val


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_against_the_reference_host_transposition(name):
    import scipy.sparse as sp
    m, n, rp, col, val = _build(name)
    got = oracle.csr2csc(m, n, rp, col, val)
    ref = oracle.ref_sptrans()
    if ref is not None:
        want = ref(m, n, rp, col, val)
        for a, b in zip(got, want):
            assert a.dtype == b.dtype and (a == b).all(), name
    g = np.load(os.path.join(GOLDEN, "ref_sptrans.npz"))
    for key, a in zip(("colptr", "rowidx", "val"), got):
        assert (g["%s_%s" % (name, key)] == a).all(), (name, key)
    # same matrix as scipy's transpose (which sorts and sums duplicates: compare dense)
    if m * n <= 2_000_000:
        A = sp.csr_matrix((val, col, rp), shape=(m, n)).toarray()
        At = sp.csc_matrix((got[2], got[1], got[0]), shape=(m, n)).toarray()
        assert np.allclose(A, At, rtol=0, atol=1e-12 * (np.abs(A).max() + 1))


def _gpu_counts():
    import torch
    k = torch.cuda.device_count()
    want = int(os.environ.get("SBLAS_EXPECT_GPUS", "0"))
    assert k >= want, "SBLAS_EXPECT_GPUS=%d but only %d visible" % (want, k)
    return [g for g in (1, 2, 4, 8) if g <= k]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_library_against_oracle(name):
    import sblas_b200 as sb
    m, n, rp, col, val = _build(name)
    want = oracle.csr2csc(m, n, rp, col, val)
    for g in _gpu_counts():
        if g > m:
            continue
        rc, colptr, rowidx, v = sb.kernal_sptrans(m, n, int(rp[-1]), g, rp, col, val, ref=want)
        assert rc == 0, (rc, sb.last_error())
        assert (colptr == want[0]).all() and (rowidx == want[1]).all() and (v == want[2]).all(), (name, g)


@pytest.mark.gpu
def test_large_power_law_matrix_and_round_trip():
    """~60 M entries, Circuit5M-shaped rows (a few hub rows of 1e5..1e6 entries, median 5), columns 80 % banded /
    20 % uniform: bit-exact against the oracle, and transposing the result again gives the input back (rows of the
    input are column-sorted, so CSR(CSC(A)^T ... ) is the identity on the arrays)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import sblas_b200 as sb
    c = bench.host_problem("circuit5m")
    m, n, nnz = c["m"], c["n"], c["nnz"]
    rp = c["rp"].astype(np.int32)
    want = oracle.csr2csc(m, n, rp, c["col"], c["val"])
    for g in _gpu_counts():
        rc, colptr, rowidx, v = sb.kernal_sptrans(m, n, nnz, g, rp, c["col"], c["val"])
        assert rc == 0, sb.last_error()
        assert (colptr == want[0]).all() and (rowidx == want[1]).all() and (v == want[2]).all(), g
        print("sptrans circuit5m ngpu=%d: %.2f ms on the devices" % (g, sb.sptrans_last_device_ms()))
    rc, rp2, col2, v2 = sb.kernal_sptrans(n, m, nnz, 1, colptr, rowidx, v)      # transpose of the transpose
    assert rc == 0
    assert (rp2 == rp).all() and (col2 == c["col"]).all() and (v2 == c["val"]).all()
