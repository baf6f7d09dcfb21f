"""Inputs shared by tests/golden/make_golden_y.py (which runs the reference's own, unmodified
entry points -- oracle/_ref/libref_spmv.so -- on a B200 and commits their y vectors) and by the
tests that compare the oracle (CPU) and this repo's library (GPU) with those vectors.

Every case is rebuilt from seeds / the committed qh768 COO fixture, so only y travels.
Entry names: "baseline", "v1k1", "v1k2" (kernel 1 = cusparseDcsrmv, 2 = cusparseDcsrmv_mp, the
reference's own choice, dspmv_mgpu_v1.cu:199-211), "v2k1" (nb = nnz/8, q = 8: the harness's
d=1,c=8 sweep point, dspmv_test.cu:314-332) and "v2k2" (nb = nnz/3, q = 2).
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_Y = os.path.join(GOLDEN, "ref_y.npz")
A, B = 0.8401877171547095, 0.39438292681909304        # harness ALPHA/BETA (glibc rand(), seed 1)
ENTRIES = ("baseline", "v1k1", "v1k2", "v2k1", "v2k2")


def _qh768():
    import oracle
    g = np.load(os.path.join(GOLDEN, "qh768_coo.npz"))
    m, n = int(g["m"]), int(g["n"])
    rp = oracle.coo_to_rowptr(m, g["row"])
    return m, n, rp, np.ascontiguousarray(g["col"]), np.ascontiguousarray(g["val"])


def _csr(rng, n, lens):
    lens = np.asarray(lens, np.int64)
    rp = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=rp[1:])
    nnz = int(rp[-1])
    col = rng.integers(0, n, size=nnz, dtype=np.int64).astype(np.int32)
    val = rng.uniform(-1.0, 1.0, size=nnz)
    return rp, col, val


def cases():
    """name -> dict(m, n, nnz, rp, col, val, x, y0, alpha, beta)."""
    import oracle
    out = {}
    m, n, rp, col, val = _qh768()
    out["qh768_harness"] = dict(m=m, n=n, rp=rp, col=col, val=val, x=np.ones(n), y0=np.zeros(m), alpha=A, beta=B)
    rng = np.random.default_rng(1)
    out["qh768_y"] = dict(m=m, n=n, rp=rp, col=col, val=val, x=rng.uniform(0.5, 1.5, n),
                          y0=rng.standard_normal(m) * 1e9, alpha=A, beta=B)
    for gn in (200, 10000):
        r, c, v, alpha, beta = oracle.gen_g(gn)
        out["g%d" % gn] = dict(m=gn, n=gn, rp=oracle.coo_to_rowptr(gn, r), col=c, val=v, x=np.ones(gn),
                               y0=np.zeros(gn), alpha=alpha, beta=beta)
    rng = np.random.default_rng(61)
    lens = np.concatenate([rng.integers(0, 6, size=3000), rng.integers(100, 300, size=500), [20000, 0, 9000],
                           np.full(5000, 2, np.int64), rng.integers(40, 120, size=5000)])
    rp2, col2, val2 = _csr(rng, 7001, lens)
    out["mixed_y"] = dict(m=len(lens), n=7001, rp=rp2, col=col2, val=val2, x=rng.uniform(0.5, 1.5, 7001),
                          y0=rng.standard_normal(len(lens)), alpha=-1.75, beta=0.625)
    out["mixed_beta0"] = dict(out["mixed_y"], alpha=2.0, beta=0.0)
    for c in out.values():
        c["nnz"] = int(c["rp"][-1])
    return out


def v2_params(nnz, entry):
    return (max(nnz // 8, 1), 8) if entry == "v2k1" else (max(nnz // 3, 1), 2)


def run_entry(api, c, entry, ngpu=1):
    """Call one entry point of `api` (oracle.ref_spmv() or the sblas_b200 module: same argument
    lists) on case c; returns (status, y)."""
    y = c["y0"].copy()
    args = (c["m"], c["n"], c["nnz"], c["alpha"], c["val"], c["rp"], c["col"], c["x"], c["beta"], y)
    name = {"baseline": ("baseline", "spMV_mgpu_baseline"), "v1": ("v1", "spMV_mgpu_v1"), "v2": ("v2", "spMV_mgpu_v2")}
    if entry == "baseline":
        fn = getattr(api, name["baseline"][0], None) or getattr(api, name["baseline"][1])
        rc = fn(*args, ngpu)
    elif entry.startswith("v1"):
        fn = getattr(api, name["v1"][0], None) or getattr(api, name["v1"][1])
        rc = fn(*args, ngpu, int(entry[-1]))
    else:
        fn = getattr(api, name["v2"][0], None) or getattr(api, name["v2"][1])
        nb, q = v2_params(c["nnz"], entry)
        rc = fn(*args, ngpu, int(entry[-1]), max(nb // ngpu, 1), q)
    return rc, y


def shared_rows(c, entry, ngpu=1):
    """Rows that lie on a shard / task border of the reference partition for this entry: where the
    reference's host merge arithmetic (and, for v2 with y != 0 and beta != 0, its single-y2 defect,
    SURVEY Appendix A) enters the result."""
    import oracle
    if entry == "baseline":
        return np.zeros(0, np.int64)
    if entry.startswith("v1"):
        p = oracle.partition_v1(c["rp"], ngpu)
    else:
        nb, _ = v2_params(c["nnz"], entry)
        p = oracle.generate_tasks_v2(c["rp"], max(nb // ngpu, 1))
    rows = [p["start_row"][t] for t in range(len(p["start_row"])) if p["start_flag"][t]]
    rows += [p["end_row"][t] for t in range(len(p["end_row"])) if p["end_flag"][t]]
    return np.unique(np.asarray(rows, np.int64))
