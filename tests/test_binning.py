"""Plan-level adaptive row binning, host logic only: 4096-row blocks -> runs of
one class (0 general, 1 short, 2 medium with R rows per warp)."""
import ctypes as C

import numpy as np

import sblas_b200 as sb

RB = 4096


def bin_rows(lens, short_max=4, medium_on=1, min_nnz=1 << 20):
    lens = np.asarray(lens, np.int64)
    nrows = len(lens)
    nblk = (nrows + RB - 1) // RB
    rp = np.zeros(nrows + 1, np.int64)
    np.cumsum(lens, out=rp[1:])
    longest = np.array([lens[b * RB:(b + 1) * RB].max() for b in range(nblk)], np.int32)
    first = np.ascontiguousarray(rp[0:nrows:RB][:nblk].astype(np.int32))
    cls, R, beg = (np.zeros(nblk + 2, np.int32) for _ in range(3))
    L = sb.lib()
    L.sblas_bin_row_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
    n = L.sblas_bin_row_blocks(longest.ctypes.data, first.ctypes.data, nblk, nrows, int(rp[-1]), short_max, medium_on,
                               min_nnz, cls.ctypes.data, R.ctypes.data, beg.ctypes.data)
    return [(int(cls[i]), int(R[i]), int(beg[i]) * RB, min(int(beg[i + 1]) * RB, nrows)) for i in range(n)]


def test_bench_shapes():
    # config 2b shape scaled down 10x: long rows then rows of 100 -> general + medium R=2
    g = np.concatenate([np.full(12500, 9000), np.full(87500, 100)])
    runs = bin_rows(g)
    assert [(c, r) for c, r, _, _ in runs] == [(0, 0), (2, 2)]
    assert runs[0][3] == 16384 and runs[1][2] == 16384          # the block that mixes both stays general
    # config 5 shape scaled down: rows of 180 then rows of 2 -> medium R=1 + short
    b = np.concatenate([np.full(62500, 180), np.full(1437500, 2)])
    runs = bin_rows(b)
    assert [(c, r) for c, r, _, _ in runs] == [(2, 1), (0, 0), (1, 0)] or [(c, r) for c, r, _, _ in runs] == [(2, 1), (1, 0)]
    assert runs[-1][0] == 1 and runs[-1][3] == len(b)


def test_small_runs_join_their_neighbours_and_equal_neighbours_merge():
    lens = np.concatenate([np.full(3 * RB, 3000), np.full(RB, 2), np.full(3 * RB, 3000)])     # 8K-entry short run
    assert [(c, r) for c, r, _, _ in bin_rows(lens)] == [(0, 0)]
    assert [(c, r) for c, r, _, _ in bin_rows(lens, min_nnz=1000)] == [(0, 0), (1, 0), (0, 0)]
    # medium blocks with different longest rows merge and keep the smallest R
    lens = np.concatenate([np.full(40 * RB, 60), np.full(40 * RB, 120), np.full(40 * RB, 64)])
    assert [(c, r) for c, r, _, _ in bin_rows(lens)] == [(2, 2)]
    assert [(c, r) for c, r, _, _ in bin_rows(lens, medium_on=0)] == [(0, 0)]


def test_class_borders():
    n = 300 * RB
    assert bin_rows(np.full(n, 4))[0][:2] == (1, 0)
    assert bin_rows(np.full(n, 5))[0][:2] == (0, 0)             # too long for thread-per-row, too short for a window
    assert bin_rows(np.full(n, 16))[0][:2] == (2, 8)
    assert bin_rows(np.full(n, 32))[0][:2] == (2, 8)
    assert bin_rows(np.full(n, 33))[0][:2] == (2, 7)
    assert bin_rows(np.full(n, 128))[0][:2] == (2, 2)
    assert bin_rows(np.full(n, 129))[0][:2] == (2, 1)
    assert bin_rows(np.full(n, 256))[0][:2] == (2, 1)
    assert bin_rows(np.full(n, 257))[0][:2] == (0, 0)
    sparse = np.full(n, 2)
    sparse[::RB] = 200                                           # one long-ish row per block: windows mostly empty
    assert bin_rows(sparse)[0][:2] == (0, 0)
