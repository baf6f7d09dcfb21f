#!/usr/bin/env python
"""Regenerates the Matrix-Market ingest fixtures.  Run in the build container (needs
oracle/_ref/libref_ingest.so, i.e. the reference's OWN loader sptrsv/sptrsv_v1/src/mmio_highlevel.h
compiled by oracle/Makefile):

    python tests/golden/make_golden_ingest.py

Writes tests/golden/ingest_<case>.mtx (small files: general / symmetric / pattern / integer /
hermitian-complex, unsorted entry order) and tests/golden/ingest_expected.npz with the CSR arrays the
REFERENCE's loader (mmio_info + mmio_data) produces for each of them, plus for the bundled sample
matrix qh768.  tests/test_ingest.py checks the oracle and the product against these everywhere
(the GPU box has no reference checkout)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

CASES = [("real", "general"), ("real", "symmetric"), ("pattern", "general"), ("pattern", "symmetric"),
         ("integer", "general"), ("integer", "symmetric"), ("complex", "hermitian")]


def write_mtx(path, m, n, entries, field, symm):
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate " + field + " " + symm + "\n")
        f.write("% s-blas_b200 ingest fixture (tests/golden/make_golden_ingest.py)\n")
        f.write("%d %d %d\n" % (m, n, len(entries)))
        for i, j, v in entries:
            if field == "pattern":
                f.write("%d %d\n" % (i + 1, j + 1))
            elif field == "integer":
                f.write("%d %d %d\n" % (i + 1, j + 1, int(v)))
            elif field == "complex":
                f.write("%d %d %.17g %.17g\n" % (i + 1, j + 1, v, 0.25))
            else:
                f.write("%d %d %.17g\n" % (i + 1, j + 1, v))


def main():
    ref = oracle.ref_ingest()
    assert ref is not None, "build oracle/_ref/libref_ingest.so first (make -C oracle)"
    out = {}
    for k, (field, symm) in enumerate(CASES):
        rng = np.random.default_rng(1000 + k)
        m = n = 29
        ents = {}
        for _ in range(170):
            i, j = int(rng.integers(0, m)), int(rng.integers(0, n))
            if symm != "general" and j > i:
                i, j = j, i
            ents[(i, j)] = float(rng.integers(-9, 10)) if field == "integer" else float(np.round(rng.standard_normal(), 6))
        entries = [(i, j, v) for (i, j), v in ents.items()]
        rng.shuffle(entries)
        name = "%s_%s" % (field, symm)
        path = os.path.join(HERE, "ingest_%s.mtx" % name)
        write_mtx(path, m, n, entries, field, symm)
        gm, gn, rp, col, val, sym = ref(path)
        out[name + "_mn"] = np.array([gm, gn, int(sym)], np.int64)
        out[name + "_rowptr"], out[name + "_col"], out[name + "_val"] = rp, col, val
    q = "/root/reference/sample_matrix/qh768.mtx"
    gm, gn, rp, col, val, sym = ref(q)
    out["qh768_mn"] = np.array([gm, gn, int(sym)], np.int64)
    out["qh768_rowptr"], out["qh768_col"], out["qh768_val"] = rp, col, val
    np.savez_compressed(os.path.join(HERE, "ingest_expected.npz"), **out)
    print("wrote", len(CASES), "fixtures + ingest_expected.npz")


if __name__ == "__main__":
    main()
