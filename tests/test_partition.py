"""Host logic of the product (C, via the C-ABI): the three partitioners must be bit-exact
with the oracle / the reference's golden vectors.  CPU only -- no compute calls."""
import json
import os

import numpy as np
import pytest

import oracle
import sblas_b200 as sb
from conftest import GOLDEN

KEYS = ("start_idx", "end_idx", "start_row", "end_row", "start_flag", "end_flag", "dev_m", "dev_nnz")


def _rowptrs():
    rng = np.random.default_rng(99)
    out = []
    for t in range(25):
        m = int(rng.integers(1, 500))
        cnt = rng.integers(1, 10, size=m)
        if t % 5 == 0:
            cnt[rng.integers(0, m)] = int(rng.integers(100, 2000))       # a row spanning several shards
        rp = np.zeros(m + 1, np.int64)
        rp[1:] = np.cumsum(cnt)
        out.append(rp)
    return out


def test_get_row_from_index_bit_exact():
    g = np.load(os.path.join(GOLDEN, "ref_row_from_index.npz"))
    for i in range(40):
        rp, ans = g["rp%d" % i], g["ans%d" % i]
        got = [sb.get_row_from_index(len(rp) - 1, rp, k) for k in range(len(ans))]
        assert got == ans.tolist()


def test_v1_partition_bit_exact(qh768):
    for rp in [qh768["rowptr"]] + _rowptrs():
        for ngpu in (1, 2, 3, 4, 8):
            if ngpu > rp[-1]:
                continue
            a, b = sb.partition_v1(rp, ngpu), oracle.partition_v1(rp, ngpu)
            for k in KEYS:
                assert (a[k] == b[k]).all(), (k, ngpu)


def test_v2_tasks_bit_exact(qh768):
    for rp in [qh768["rowptr"]] + _rowptrs()[:10]:
        nnz = int(rp[-1])
        for d in (1, 2, 4, 8):
            for c in (1, 2, 4, 8):
                nb = nnz // (d * c)                 # the harness sweep, dspmv_test.cu:314-332
                if nb <= 0:
                    continue
                a, b = sb.generate_tasks_v2(rp, nb), oracle.generate_tasks_v2(rp, nb)
                for k in KEYS:
                    assert (a[k] == b[k]).all(), (k, nb)
                T = len(a["dev_m"])
                # fixed task->GPU map honours the reference quota (dspmv_mgpu_v2.cu:125-126)
                owners = [sb.v2_task_owner(T, d, t) for t in range(T)]
                assert owners == sorted(owners)
                for g in range(d):
                    assert owners.count(g) == oracle.lib().oracle_v2_quota(T, g, d)


def test_baseline_partition_bit_exact(qh768):
    for rp in [qh768["rowptr"]] + _rowptrs():
        for ngpu in (1, 2, 3, 4, 8):
            a, b = sb.partition_baseline(rp, ngpu), oracle.partition_baseline(rp, ngpu)
            for k in ("start_row", "end_row", "dev_m", "dev_nnz"):
                assert (a[k] == b[k]).all(), (k, ngpu)


def test_local_rowptr_bit_exact(qh768):
    rp = qh768["rowptr"]
    for ngpu in (2, 4, 8):
        p = sb.partition_v1(rp, ngpu)
        for d in range(ngpu):
            part = {k: p[k][d] for k in KEYS}
            want = oracle.local_rowptr_v1(rp, part["start_idx"], part["start_row"], part["dev_m"], part["dev_nnz"])
            assert (sb.local_rowptr(rp, part) == want).all()
        b = sb.partition_baseline(rp, ngpu)
        for d in range(ngpu):
            part = {k: b[k][d] for k in KEYS}
            want = oracle.local_rowptr_baseline(rp, part["start_row"], part["dev_m"])
            assert (sb.local_rowptr(rp, part, baseline=True) == want).all()


def test_golden_known_answers(qh768):
    gold = json.load(open(os.path.join(GOLDEN, "ref_partitions.json")))
    for g, rows in gold["qh768"]["v1_known_answers_survey_8c"].items():
        p = sb.partition_v1(qh768["rowptr"], int(g))
        got = [[int(p[k][i]) for k in KEYS[:6]] for i in range(int(g))]
        assert got == rows


def test_layout_only_plan_lists_one_panel_per_segment():
    """Without a GPU (SBLAS_LAYOUT_ONLY) a plan cannot bin rows (the statistics are computed on the
    device), so every live segment is exactly one panel that covers its rows and entries; the panel
    accessors of the C-ABI work on such plans."""
    import sblas_b200 as sb
    rng = np.random.default_rng(3)
    lens = np.concatenate([rng.integers(1, 9, size=400), [5000, 0, 0, 3000], rng.integers(50, 90, size=100)])
    rp = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=rp[1:])
    m, nnz = len(lens), int(rp[-1])
    col = np.zeros(nnz, np.int32)
    val = np.ones(nnz)
    for version, world, nb, q in ((sb.V1, 3, 0, 1), (sb.V2, 2, nnz // 7 + 1, 2), (sb.BASELINE, 4, 0, 1)):
        for rank in range(world):
            p = sb.Plan.create_rank(version, m, m, nnz, val, rp, col, world, rank, 0, kernel=1, nb=nb, q=q,
                                    flags=sb.LAYOUT_ONLY)
            segs = p.local_segments()
            units = p.units()
            assert len(units) == len(segs)
            for u, s in zip(units, segs):
                assert (u["row_lo"], u["row_hi"], u["nz0"], u["nz1"]) == (s["row_lo"], s["row_hi"], s["nz0"], s["nz1"])
                assert u["kind"] in (1, 3)          # lanes-per-row below 65,536 entries per GPU, else the TMA kernel
            p.destroy()


def test_byte_balanced_partition_properties():
    """The opt-in fourth version (not in the reference): contiguous nnz ranges covering [0, nnz) like v1, rows and flags
    from the same lookups, every shard within one row's weight of an equal share of 12*nnz + 28*rows bytes."""
    import sblas_b200 as sb
    rng = np.random.default_rng(5)
    for lens in (np.concatenate([np.full(6250, 180, np.int64), np.full(43750, 2, np.int64)]),
                 rng.integers(1, 50, size=20000), np.concatenate([[100000], rng.integers(1, 9, size=5000)])):
        rp = np.zeros(len(lens) + 1, np.int64)
        rp[1:] = np.cumsum(lens)
        m, nnz = len(lens), int(rp[-1])
        for g in (1, 2, 3, 8):
            p = sb.partition_bytes(rp, g)
            assert p["start_idx"][0] == 0 and p["end_idx"][-1] == nnz - 1
            assert (p["start_idx"][1:] == p["end_idx"][:-1] + 1).all()
            share = (12.0 * nnz + 28.0 * m) / g
            for i in range(g):
                s, e = int(p["start_idx"][i]), int(p["end_idx"][i])
                r0 = int(np.searchsorted(rp, s, side="right")) - 1
                r1 = int(np.searchsorted(rp, e + 1, side="right")) - 1
                w = 12.0 * (e + 1 - s) + 28.0 * (r1 - r0)
                assert abs(w - share) <= 12.0 * lens.max() + 56.0, (g, i, w, share)
                assert p["start_row"][i] == sb.get_row_from_index(m, rp, s) and p["end_row"][i] == sb.get_row_from_index(m, rp, e)


def test_partitions_bit_exact_on_the_full_size_row_pointers():
    """The partitioners on the row pointers of the BASELINE configs themselves (config 2b: 1 M rows / 1.2125 G
    entries; config 5: 50 M rows / 1.2125 G entries; config 3: the power-law row lengths), not only on qh768-sized
    inputs: int64 arithmetic near 2^31 per shard, 50 M-entry bisections."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    for name in ("g1m", "big50m", "circuit5m"):
        lens = bench.workload(name)["row_len"]()
        rp = np.zeros(len(lens) + 1, np.int64)
        np.cumsum(lens, out=rp[1:])
        nnz = int(rp[-1])
        for ngpu in (1, 2, 4, 8):
            a, b = sb.partition_v1(rp, ngpu), oracle.partition_v1(rp, ngpu)
            for k in KEYS:
                assert (a[k] == b[k]).all(), (name, "v1", k, ngpu)
            a, b = sb.partition_baseline(rp, ngpu), oracle.partition_baseline(rp, ngpu)
            for k in ("start_row", "end_row", "dev_m", "dev_nnz"):
                assert (a[k] == b[k]).all(), (name, "baseline", k, ngpu)
            nb = nnz // (ngpu * 8)                               # the harness's d = ngpu, c = 8 sweep point
            a, b = sb.generate_tasks_v2(rp, nb), oracle.generate_tasks_v2(rp, nb)
            for k in KEYS:
                assert (a[k] == b[k]).all(), (name, "v2", k, ngpu)
        # the reference's own compiled row lookup on the shard borders of the 8-way split
        ref = oracle.ref_helper()
        if ref is not None:
            p8 = sb.partition_v1(rp, 8)
            for i in range(8):
                for idx in (int(p8["start_idx"][i]), int(p8["end_idx"][i])):
                    assert ref(len(rp) - 1, rp, idx) == sb.get_row_from_index(len(rp) - 1, rp, idx), (name, idx)


def test_partitioners_bit_exact_on_arbitrary_row_pointers_incl_empty_rows():
    """Property test (hypothesis): row pointers with EMPTY rows (where the reference's bisection names a neighbouring
    row), single-entry matrices, one huge row, more GPUs than entries.  The product's partitioners must
    equal the oracle's field for field, and the oracle's row lookup is pinned on the reference's compiled helper."""
    from hypothesis import given, settings, strategies as st

    ref = oracle.ref_helper()

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.one_of(st.just(0), st.integers(0, 6), st.integers(0, 400)), min_size=1, max_size=60),
           st.integers(1, 9), st.integers(1, 8))
    def check(lens, ngpu, c):
        rp = np.zeros(len(lens) + 1, np.int64)
        rp[1:] = np.cumsum(lens)
        m, nnz = len(lens), int(rp[-1])
        for idx in {0, nnz // 2, max(nnz - 1, 0)}:
            got = sb.get_row_from_index(m, rp, idx)
            assert got == oracle.get_row_from_index(rp, idx)
            if ref is not None:
                assert got == ref(m, rp, idx)
        a, b = sb.partition_baseline(rp, ngpu), oracle.partition_baseline(rp, ngpu)
        for k in ("start_row", "end_row", "dev_m", "dev_nnz"):
            assert (a[k] == b[k]).all(), ("baseline", k)
        if nnz >= ngpu:
            a, b = sb.partition_v1(rp, ngpu), oracle.partition_v1(rp, ngpu)
            for k in KEYS:
                assert (a[k] == b[k]).all(), ("v1", k)
            assert a["start_idx"][0] == 0 and a["end_idx"][-1] == nnz - 1
            assert (a["start_idx"][1:] == a["end_idx"][:-1] + 1).all()
        nb = nnz // (ngpu * c)
        if nb > 0:
            a, b = sb.generate_tasks_v2(rp, nb), oracle.generate_tasks_v2(rp, nb)
            for k in KEYS:
                assert (a[k] == b[k]).all(), ("v2", k)
            assert int(a["dev_nnz"].sum()) == nnz

    check()
