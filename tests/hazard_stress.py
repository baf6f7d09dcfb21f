"""Stress input for the stage-release hazard, run as a script by
tests/test_spmv_gpu.py::test_stage_release_hazard_regression with SBLAS_LIB pointing at the library
under test: ~226 M entries with FULLY SCATTERED columns (deep load/store-unit queues) through every
kernel that hands a bulk-copy stage back early -- path A and path W of the general kernel, the row-tile
and the row-split kernel -- compared over ALL rows with the independent register-tile kernel
(kernel = 3: no staging ring).  Prints one JSON line {"rows", "bad", "worst"}."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(products=4):
    import torch
    import sblas_b200 as sb
    rng = np.random.default_rng(5)
    lens = np.concatenate([np.full(8000, 9000, np.int64), np.full(300000, 180, np.int64), np.full(61440, 1000, np.int64),
                           rng.integers(100, 301, size=200000)]).astype(np.int64)
    m, n = len(lens), 1 << 24
    rp = np.zeros(m + 1, np.int64)
    np.cumsum(lens, out=rp[1:])
    nnz = int(rp[-1])
    d_val = torch.empty(nnz, dtype=torch.float64, device="cuda")
    d_col = torch.empty(nnz, dtype=torch.int32, device="cuda")
    d_rp = torch.from_numpy(rp).cuda()
    sb.synth_fill_csr(d_rp.data_ptr(), 0, m, 0, nnz, n, sb.COLS_UNIFORM, 0, 11, d_val.data_ptr(), d_col.data_ptr())
    torch.cuda.synchronize()
    plans = [sb.Plan.create_rank(sb.V1, m, n, nnz, d_val.data_ptr(), rp, d_col.data_ptr(), 1, 0, 0, kernel=k,
                                 flags=sb.SRC_DEVICE_SHARD) for k in (1, 3)]
    kinds = sorted(set(u["kind"] for u in plans[0].units()))
    x = rng.uniform(0.5, 1.0, n)
    bad, worst = 0, 0.0
    ys = []
    for p in plans:
        p.upload(x, None)
    for it in range(products):
        out = []
        for p in plans:
            p.execute_device(1.0, 0.0, sync=True)
            y = np.zeros(m)
            p.download(y)
            out.append(y)
        err = np.abs(out[0] - out[1]) / np.maximum(out[1], 1e-300)      # entries and x are positive: the bound is y itself
        bad += int((err > 2e-12).sum())
        worst = max(worst, float(err.max()))
    for p in plans:
        p.destroy()
    print(json.dumps({"rows": m, "nnz": nnz, "products": products, "bad": bad, "worst": worst, "kinds": kinds,
                      "lib": os.path.basename(sb.LIB_PATH)}))


if __name__ == "__main__":
    main()
