"""Generate tests/golden/ref_y.npz: the y vectors the REFERENCE'S OWN, UNMODIFIED entry points
(spmv/src/dspmv_mgpu_{baseline,v1,v2}.cu compiled where they lie into oracle/_ref/libref_spmv.so
with oracle/compat_csrmv.h mapping the removed cusparseDcsrmv[_mp] onto cusparseSpMV) produce on
a B200 for the cases of tests/ref_cases.py, ngpu = 1.

Run on the GPU box (no GPU in the build container):
    gpurun -- 'python tests/golden/make_golden_y.py gpurun_out/ref_y.npz'
then copy gpurun_out/ref_y.npz to tests/golden/ref_y.npz and commit it.  The run also prints how
far the oracle (CPU restatement) and this repo's library are from every vector, in units of the
north-star bound 1e-12 * (|alpha| sum|a||x| + |beta||y|).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(out_path):
    import oracle
    import ref_cases
    import sblas_b200 as sb
    import torch
    ref = oracle.ref_spmv()
    assert ref is not None, "oracle/_ref/libref_spmv.so missing (make -C oracle refspmv)"
    store, report = {}, {}
    for name, c in ref_cases.cases().items():
        want = oracle.csr_spmv(c["rp"], c["col"], c["val"], c["x"], c["alpha"], c["beta"], c["y0"])
        bound = np.maximum(oracle.csr_spmv_bound(c["rp"], c["col"], c["val"], c["x"], c["alpha"], c["beta"], c["y0"]), 1e-300)
        for entry in ref_cases.ENTRIES:
            rc, y = ref_cases.run_entry(ref, c, entry, 1)
            if not entry.startswith("v2"):
                assert rc == 0, (name, entry, rc)
            store["%s/%s" % (name, entry)] = y
            rc2, ylib = ref_cases.run_entry(sb, c, entry, 1)
            assert rc2 == 0, (name, entry, sb.last_error())
            sh = ref_cases.shared_rows(c, entry, 1)
            interior = np.ones(c["m"], bool)
            interior[sh] = False
            report["%s/%s" % (name, entry)] = {
                "ref_vs_oracle_all": float((np.abs(y - want) / bound).max() / 1e-12),
                "ref_vs_oracle_interior": float((np.abs(y - want) / bound)[interior].max() / 1e-12),
                "lib_vs_ref_all": float((np.abs(ylib - y) / bound).max() / 1e-12),
                "lib_vs_ref_interior": float((np.abs(ylib - y) / bound)[interior].max() / 1e-12),
                "lib_vs_oracle_all": float((np.abs(ylib - want) / bound).max() / 1e-12),
                "shared_rows": int(len(sh)),
            }
    meta = {"gpu": torch.cuda.get_device_name(0), "cuda_runtime": torch.version.cuda,
            "ref_lib": "oracle/_ref/libref_spmv.so (unmodified reference sources + oracle/compat_csrmv.h)",
            "units": "multiples of 1e-12 * (|alpha| sum|a||x| + |beta||y|)", "report": report}
    np.savez_compressed(out_path, meta=json.dumps(meta), **store)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_y.npz"))
