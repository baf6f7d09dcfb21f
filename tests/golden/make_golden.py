#!/usr/bin/env python
"""Regenerates the golden fixtures in this directory.  Run in the build container
(needs /root/reference and oracle/_ref/libref_helper.so built by oracle/Makefile):

    python tests/golden/make_golden.py

Fixtures written
  qh768_coo.npz          the reference's bundled sample matrix (sample_matrix/qh768.mtx,
                         SuiteSparse Bai/qh768) as (m, n, row, col, val) IN FILE ORDER --
                         i.e. what the reference harness's loader produces
                         (spmv/test/dspmv_test.cu:101-136; not row sorted).
                         Parsed here with plain Python, independent of the oracle's loader.
  ref_row_from_index.npz outputs of the REFERENCE's compiled get_row_from_index
                         (spmv/src/spmv_helper.cu:16-39) on seeded row pointers, with and
                         without empty rows (incl. the F8 quirk cases).
  ref_partitions.json    v1 / v2 / baseline partition arrays for qh768 and small synthetic
                         row pointers, computed by THIS script from the reference formulas
                         (dspmv_mgpu_v1.cu:59-133, dspmv_mgpu_v2.cu:218-289,
                         dspmv_mgpu_baseline.cu:60-87) written independently in Python and
                         using the reference's compiled helper for every row lookup; plus the
                         hand-derived known answers of SURVEY.md section 8c, which the script
                         asserts against before writing.
"""
import ctypes as C
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_MTX = "/root/reference/sample_matrix/qh768.mtx"


def ref_fn():
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_helper.so"))
    f = getattr(R, "_Z18get_row_from_indexiPxx")
    f.argtypes = [C.c_int, np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS"), C.c_longlong]
    f.restype = C.c_int
    return f


def parse_mtx(path):
    rows, cols, vals = [], [], []
    with open(path) as fh:
        header = fh.readline()
        assert header.startswith("%%MatrixMarket")
        line = fh.readline()
        while line.startswith("%"):
            line = fh.readline()
        m, n, nnz = (int(t) for t in line.split())
        for _ in range(nnz):
            a, b, c = fh.readline().split()
            rows.append(int(a) - 1)
            cols.append(int(b) - 1)
            vals.append(float(c))
    return m, n, np.array(rows, np.int32), np.array(cols, np.int32), np.array(vals, np.float64)


def rowptr_of(m, rows):
    cnt = np.bincount(rows, minlength=m).astype(np.int64)
    rp = np.zeros(m + 1, np.int64)
    rp[1:] = np.cumsum(cnt)
    return rp


def v1_partition(rp, ngpu, R):
    m, nnz = len(rp) - 1, int(rp[-1])
    out = []
    for i in range(ngpu):
        s = int(math.floor(float(i * nnz) / ngpu))
        e = int(math.floor(float((i + 1) * nnz) / ngpu)) - 1
        sr, er = R(m, rp, s), R(m, rp, e)
        out.append(dict(start_idx=s, end_idx=e, start_row=sr, end_row=er,
                        start_flag=int(s > rp[sr]), end_flag=int(e < rp[er + 1] - 1),
                        dev_m=er - sr + 1, dev_nnz=e - s + 1))
    return out


def v2_tasks(rp, nb, R):
    m, nnz = len(rp) - 1, int(rp[-1])
    T = (nnz + nb - 1) // nb
    out = []
    for t in range(T):
        s = (t * nnz) // T
        e = ((t + 1) * nnz) // T - 1
        sr, er = R(m, rp, s), R(m, rp, e)
        out.append(dict(start_idx=s, end_idx=e, start_row=sr, end_row=er,
                        start_flag=int(s > rp[sr]), end_flag=int(e < rp[er + 1] - 1),
                        dev_m=er - sr + 1, dev_nnz=e - s + 1))
    return out


def baseline_partition(rp, ngpu):
    m = len(rp) - 1
    out = []
    for d in range(ngpu):
        sr, er = (d * m) // ngpu, ((d + 1) * m) // ngpu - 1
        out.append(dict(start_row=sr, end_row=er, dev_m=er - sr + 1, dev_nnz=int(rp[er + 1] - rp[sr])))
    return out


def main():
    if not os.path.exists(REF_MTX):
        sys.exit("reference checkout not present; fixtures are already committed")
    R = ref_fn()
    m, n, rows, cols, vals = parse_mtx(REF_MTX)
    np.savez_compressed(os.path.join(HERE, "qh768_coo.npz"), m=m, n=n, row=rows, col=cols, val=vals)
    rp = rowptr_of(m, rows)

    # --- get_row_from_index vectors from the reference object
    rng = np.random.default_rng(20261018)
    cases = []
    for k in range(40):
        mm = int(rng.integers(1, 400))
        lo = 0 if k % 2 else 1                      # odd cases contain empty rows
        cnt = rng.integers(lo, 9, size=mm)
        if cnt.sum() == 0:
            cnt[0] = 1
        r = np.zeros(mm + 1, np.int64)
        r[1:] = np.cumsum(cnt)
        idx = np.arange(int(r[-1]), dtype=np.int64)
        ans = np.array([R(mm, r, int(i)) for i in idx], np.int32)
        cases.append((r, ans))
    np.savez_compressed(os.path.join(HERE, "ref_row_from_index.npz"),
                        **{"rp%d" % i: c[0] for i, c in enumerate(cases)},
                        **{"ans%d" % i: c[1] for i, c in enumerate(cases)})

    # --- partitions
    gold = {"qh768": {"m": m, "n": n, "nnz": int(rp[-1]), "v1": {}, "v2": {}, "baseline": {}}}
    for g in (1, 2, 3, 4, 8):
        gold["qh768"]["v1"][str(g)] = v1_partition(rp, g, R)
        gold["qh768"]["baseline"][str(g)] = baseline_partition(rp, g)
    nnz = int(rp[-1])
    for d in (1, 2, 4, 8):
        for c in (1, 2, 4, 8):
            nb = nnz // (d * c)                      # harness sweep, dspmv_test.cu:314-332
            gold["qh768"]["v2"]["%d" % nb] = v2_tasks(rp, nb, R)
    # SURVEY.md section 8c known answers (start_idx,end_idx,start_row,end_row,start_flag,end_flag)
    known = {
        "1": [(0, 2933, 0, 767, 0, 0)],
        "2": [(0, 1466, 0, 435, 0, 1), (1467, 2933, 435, 767, 1, 0)],
        "4": [(0, 732, 0, 191, 0, 1), (733, 1466, 191, 435, 1, 1), (1467, 2199, 435, 566, 1, 1),
              (2200, 2933, 566, 767, 1, 0)],
        "8": [(0, 365, 0, 91, 0, 1), (366, 732, 91, 191, 1, 1), (733, 1099, 191, 308, 1, 0),
              (1100, 1466, 309, 435, 0, 1), (1467, 1832, 435, 504, 1, 1), (1833, 2199, 504, 566, 1, 1),
              (2200, 2566, 566, 651, 1, 1), (2567, 2933, 651, 767, 1, 0)],
    }
    for g, rows_ in known.items():
        got = [(p["start_idx"], p["end_idx"], p["start_row"], p["end_row"], p["start_flag"], p["end_flag"])
               for p in gold["qh768"]["v1"][g]]
        assert got == rows_, (g, got, rows_)
    gold["qh768"]["v1_known_answers_survey_8c"] = {k: [list(t) for t in v] for k, v in known.items()}

    # small synthetic row pointers (no empty rows) incl. rows spanning >= 3 shards
    synth = {}
    for name, cnt in {
        "one_long_row": [1, 50, 1, 1],
        "uniform": [3] * 20,
        "skew": [1, 1, 40, 2, 2, 1, 30, 1],
        "single_row": [17],
    }.items():
        r = np.zeros(len(cnt) + 1, np.int64)
        r[1:] = np.cumsum(cnt)
        ent = {"rowptr": r.tolist(), "v1": {}, "v2": {}, "baseline": {}}
        for g in (1, 2, 3, 4, 8):
            if g <= int(r[-1]):
                ent["v1"][str(g)] = v1_partition(r, g, R)
            if g <= len(cnt):
                ent["baseline"][str(g)] = baseline_partition(r, g)
        for nb in (1, 2, 5, 7, 16, int(r[-1])):
            ent["v2"][str(nb)] = v2_tasks(r, nb, R)
        synth[name] = ent
    gold["synthetic"] = synth
    gold["alpha_beta_f_mode"] = [1804289383 / 2147483647, 846930886 / 2147483647]   # glibc rand() seed 1
    with open(os.path.join(HERE, "ref_partitions.json"), "w") as fh:
        json.dump(gold, fh)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
