"""Generate tests/golden/ref_sptrans.npz from the reference's OWN host transposition
(sptrans/sptrans_v1/src/tranpose.h compiled where it lies into oracle/_ref/libref_sptrans.so by oracle/Makefile;
pure host code, runs in the build container) for the cases of tests/test_sptrans.py."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
import test_sptrans  # noqa: E402

ref = oracle.ref_sptrans()
assert ref is not None, "make -C oracle refsptrans first (needs /root/reference)"
out = {}
for name in sorted(test_sptrans.CASES):
    m, n, rp, col, val = test_sptrans._build(name)
    colptr, rowidx, v = ref(m, n, rp, col, val)
    out[name + "_colptr"], out[name + "_rowidx"], out[name + "_val"] = colptr, rowidx, v
np.savez_compressed(os.path.join(HERE, "ref_sptrans.npz"), **out)
print("wrote", len(out), "arrays")
