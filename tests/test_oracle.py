"""The oracle against the reference's own artefacts: golden vectors produced from the
reference's compiled spmv_helper.cu and sample matrix (tests/golden/make_golden.py),
the known answers of SURVEY.md section 8c, and -- when oracle/_ref is present -- the
reference object itself.  CPU only."""
import json
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, make_csr


def test_get_row_from_index_matches_reference_vectors():
    g = np.load(os.path.join(GOLDEN, "ref_row_from_index.npz"))
    n = len([k for k in g.files if k.startswith("rp")])
    assert n == 40
    for i in range(n):
        rp, ans = g["rp%d" % i], g["ans%d" % i]
        got = np.array([oracle.get_row_from_index(rp, k) for k in range(len(ans))], np.int32)
        assert (got == ans).all(), i


def test_get_row_from_index_survey_probe():
    rp = np.array([0, 3, 4, 8], np.int64)
    assert [oracle.get_row_from_index(rp, i) for i in range(9)] == [0, 0, 0, 1, 2, 2, 2, 2, 3]
    # F8: next to empty rows the reference names a neighbouring row -- reproduced
    assert oracle.get_row_from_index(np.array([0, 2, 2, 2, 5, 5, 6], np.int64), 5) == 4
    assert oracle.get_row_from_index(np.array([0, 0, 0, 3], np.int64), 0) == 1


def test_get_row_from_index_against_compiled_reference():
    ref = oracle.ref_helper()
    if ref is None:
        pytest.skip("oracle/_ref not built (reference checkout absent)")
    rng = np.random.default_rng(7)
    for trial in range(60):
        m = int(rng.integers(1, 300))
        cnt = rng.integers(0 if trial % 3 == 0 else 1, 7, size=m)
        cnt[0] = max(cnt[0], 1)
        rp = np.zeros(m + 1, np.int64)
        rp[1:] = np.cumsum(cnt)
        for idx in range(int(rp[-1])):
            assert oracle.get_row_from_index(rp, idx) == ref(m, rp, idx)


def _cmp_parts(got, want):
    for i, w in enumerate(want):
        for k, v in w.items():
            assert int(got[k][i]) == v, (i, k, int(got[k][i]), v)


def test_partitions_match_golden(qh768):
    gold = json.load(open(os.path.join(GOLDEN, "ref_partitions.json")))
    q = gold["qh768"]
    assert (q["m"], q["n"], q["nnz"]) == (qh768["m"], qh768["n"], qh768["nnz"]) == (768, 768, 2934)
    for g, want in q["v1"].items():
        _cmp_parts(oracle.partition_v1(qh768["rowptr"], int(g)), want)
    for g, want in q["baseline"].items():
        _cmp_parts(oracle.partition_baseline(qh768["rowptr"], int(g)), want)
    for nb, want in q["v2"].items():
        got = oracle.generate_tasks_v2(qh768["rowptr"], int(nb))
        assert len(got["dev_m"]) == len(want)
        _cmp_parts(got, want)
    # d=1,c=8 -> nb=366 -> 9 tasks; d=2,c=8 -> nb=183 -> 17 tasks
    assert len(q["v2"]["366"]) == 9 and len(q["v2"]["183"]) == 17
    for name, ent in gold["synthetic"].items():
        rp = np.array(ent["rowptr"], np.int64)
        for g, want in ent["v1"].items():
            _cmp_parts(oracle.partition_v1(rp, int(g)), want)
        for g, want in ent["baseline"].items():
            _cmp_parts(oracle.partition_baseline(rp, int(g)), want)
        for nb, want in ent["v2"].items():
            _cmp_parts(oracle.generate_tasks_v2(rp, int(nb)), want)


def test_local_rowptr_shapes(qh768):
    rp = qh768["rowptr"]
    p = oracle.partition_v1(rp, 4)
    for d in range(4):
        lp = oracle.local_rowptr_v1(rp, p["start_idx"][d], p["start_row"][d], p["dev_m"][d], p["dev_nnz"][d])
        assert lp[0] == 0 and lp[-1] == p["dev_nnz"][d] and (np.diff(lp) >= 0).all()
    b = oracle.partition_baseline(rp, 4)
    assert b["dev_nnz"].sum() == qh768["nnz"] and b["dev_m"].sum() == qh768["m"]


def test_loader_reproduces_harness_order(qh768, tmp_path):
    """Write the fixture back as .mtx and load it with the restated harness loader:
    file order kept, 0-based, not row sorted (F3)."""
    p = tmp_path / "qh768.mtx"
    with open(p, "w") as fh:
        fh.write("%%MatrixMarket matrix coordinate real general\n% regenerated from tests/golden/qh768_coo.npz\n")
        fh.write("%d %d %d\n" % (qh768["m"], qh768["n"], qh768["nnz"]))
        for r, c, v in zip(qh768["row"], qh768["col"], qh768["val"]):
            fh.write("%d %d %s\n" % (r + 1, c + 1, repr(float(v))))
    m, n, r, c, v = oracle.load_mtx(str(p), "f")
    assert (m, n) == (768, 768)
    assert (r == qh768["row"]).all() and (c == qh768["col"]).all() and (v == qh768["val"]).all()
    assert (np.diff(c) >= 0).all() and not (np.diff(r) >= 0).all()      # column sorted file, rows scrambled
    # 'b' mode reads "%d %d" per line (pattern files) and sets every value to 0.00001
    pb = tmp_path / "qh768_pattern.mtx"
    with open(pb, "w") as fh:
        fh.write("%%MatrixMarket matrix coordinate pattern general\n")
        fh.write("%d %d %d\n" % (qh768["m"], qh768["n"], qh768["nnz"]))
        for r_, c_ in zip(qh768["row"], qh768["col"]):
            fh.write("%d %d\n" % (r_ + 1, c_ + 1))
    _, _, rb, cb, vb = oracle.load_mtx(str(pb), "b")
    assert (vb == 0.00001).all() and (rb == qh768["row"]).all() and (cb == qh768["col"]).all()
    cnt = np.bincount(r, minlength=m)
    assert cnt.min() == 1 and cnt.max() == 10


def test_generator_sizes_and_rand():
    gold = json.load(open(os.path.join(GOLDEN, "ref_partitions.json")))
    r, c, v, alpha, beta = oracle.gen_g(200)
    assert len(v) == 4850                       # g 200 -> 25 rows x 180 + 175 rows x 2
    cnt = np.bincount(r, minlength=200)
    assert (cnt[:25] == 180).all() and (cnt[25:] == 2).all()
    assert (c[:180] == np.arange(180)).all()
    assert v[0] == 1804289383 / 2147483647      # glibc rand() seed 1, first draw
    assert 0.0 <= alpha <= 1.0 and 0.0 <= beta <= 1.0
    # f mode: ALPHA/BETA are the first two draws
    oracle.lib().oracle_srand(1)
    assert [oracle.lib().oracle_rand_unit(), oracle.lib().oracle_rand_unit()] == gold["alpha_beta_f_mode"]
    with pytest.raises(ValueError):
        oracle.gen_g(4)


def test_csr_spmv_against_scipy():
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    m, n = 300, 257
    rp, col, val = make_csr(rng, m, n, rng.integers(0, 12, size=m))
    x, y = rng.standard_normal(n), rng.standard_normal(m)
    A = sp.csr_matrix((val, col, rp), shape=(m, n))
    got = oracle.csr_spmv(rp, col, val, x, 0.7, -1.3, y)
    want = 0.7 * (A @ x) - 1.3 * y
    bound = oracle.csr_spmv_bound(rp, col, val, x, 0.7, -1.3, y)
    assert (np.abs(got - want) <= 1e-13 * bound + 1e-300).all()


@pytest.mark.parametrize("ngpu", [1, 2, 3, 4, 8])
def test_multi_gpu_restatements_agree_with_single(qh768, ngpu):
    """The restated v1 / v2 / baseline flows (partition + per-shard csrmv + reference host
    merge) reproduce the single-matrix product, also with y != 0, beta != 0."""
    rng = np.random.default_rng(ngpu)
    x = rng.uniform(0.5, 1.5, qh768["n"])
    y = rng.standard_normal(qh768["m"]) * 1e6
    a, b = 0.8401877171547095, 0.39438292681909304
    args = (qh768["rowptr"], qh768["col"], qh768["val"], x, a, b, y)
    want = oracle.csr_spmv(*args)
    bound = oracle.csr_spmv_bound(*args)
    for got in (oracle.spmv_mgpu_v1(*args, ngpu), oracle.spmv_mgpu_baseline(*args, ngpu),
                oracle.spmv_mgpu_v2(*args, qh768["nnz"] // (ngpu * 4))):
        assert (np.abs(got - want) <= 1e-12 * bound).all()


def test_v2_single_y2_defect_of_the_reference_is_documented(qh768):
    """struct spmv_task has one y2 (dspmv_mgpu_v2.cu:249,263): with y != 0, beta != 0 a task
    split at both ends corrects its start row with the END row's original y.  Faithful
    restatement differs from the definition there; with y == 0 (the harness) it does not."""
    rng = np.random.default_rng(5)
    x = rng.uniform(0.5, 1.5, qh768["n"])
    a, b = 0.84, 0.39
    y = rng.standard_normal(qh768["m"]) * 1e6
    args = (qh768["rowptr"], qh768["col"], qh768["val"], x, a, b)
    want, bound = oracle.csr_spmv(*args, y), oracle.csr_spmv_bound(*args, y)
    nb = qh768["nnz"] // 4
    faithful = oracle.spmv_mgpu_v2(*args, y, nb, faithful_y2=True)
    assert (np.abs(faithful - want) > 1e-12 * bound).any()
    y0 = np.zeros(qh768["m"])
    assert (oracle.spmv_mgpu_v2(*args, y0, nb, faithful_y2=True) == oracle.spmv_mgpu_v2(*args, y0, nb)).all()


def test_multi_gpu_restatement_row_spanning_three_shards():
    rng = np.random.default_rng(11)
    rp, col, val = make_csr(rng, 4, 64, [1, 50, 1, 1])
    x, y = rng.standard_normal(64), rng.standard_normal(4)
    args = (rp, col, val, x, 1.25, 0.5, y)
    want, bound = oracle.csr_spmv(*args), oracle.csr_spmv_bound(*args)
    for g in (2, 3, 4, 8):
        assert (np.abs(oracle.spmv_mgpu_v1(*args, g) - want) <= 1e-12 * bound).all()
    for nb in (1, 2, 5, 7, 16):
        assert (np.abs(oracle.spmv_mgpu_v2(*args, nb) - want) <= 1e-12 * bound).all()
