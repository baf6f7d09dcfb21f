"""The host layout of rank plans (partition -> live segments -> edge slots -> merge lists) on many more shapes than
the gloo test can afford: every rank's SBLAS_LAYOUT_ONLY plan is built in THIS process, the CPU stand-in of
tests/test_dist_cpu.py plays the kernels, the all-gather of the edge blocks is a concatenation, and the assembled y
must match the oracle with every row written exactly once.  Row pointers come from hypothesis: empty rows (where the
reference's row lookup names a neighbour), rows spanning many segments, more ranks than rows."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle
import sblas_b200 as sb
from test_dist_cpu import A, B, _segment_cpu


def _emulate(version, rp, col, val, n, world, nb, q, x, y0):
    m, nnz = len(rp) - 1, int(rp[-1])
    plans = [sb.Plan.create_rank(version, m, n, nnz, 0, rp, 0, world, r, 0, kernel=2, nb=nb, q=q, flags=sb.LAYOUT_ONLY)
             for r in range(world)]
    try:
        slots = max(max(p.edge_slots for p in plans), 1)
        assert all(max(p.edge_slots, 1) == slots for p in plans), "edge block size must agree on every rank"
        table = np.zeros(world * slots)
        y, owned = np.zeros(m), np.zeros(m, np.int64)
        for r, p in enumerate(plans):
            for seg in p.local_segments():
                rows, out, keep, e = _segment_cpu(rp, col, val, x, y0, seg, A, B)
                y[rows[keep]] = out[keep]
                owned[rows[keep]] += 1
                table[r * slots + seg["edge_slot"]] = e[0]
                table[r * slots + seg["edge_slot"] + 1] = e[1]
        for p in plans:
            segs = p.local_segments()
            first_row = segs[0]["dev_first_row"] if segs else 0
            for lrow, offs in p.merge_list():
                rr = first_row + lrow
                y[rr] = A * sum(float(table[o]) for o in offs) + B * y0[rr]
                owned[rr] += 1
        return y, owned
    finally:
        for p in plans:
            p.destroy()


@settings(max_examples=250, deadline=None)
@given(st.lists(st.one_of(st.just(0), st.integers(0, 5), st.integers(0, 300)), min_size=1, max_size=40),
       st.integers(1, 8), st.sampled_from(["v1", "v2", "baseline", "bytes"]), st.integers(1, 8), st.integers(0, 2 ** 31))
def test_rank_layouts_assemble_the_product(lens, world, version, c, seed):
    rng = np.random.default_rng(seed)
    m, n = len(lens), 50
    rp = np.zeros(m + 1, np.int64)
    rp[1:] = np.cumsum(lens)
    nnz = int(rp[-1])
    if nnz < world:                   # the reference's v1 needs at least one entry per GPU
        return
    col = rng.integers(0, n, size=nnz).astype(np.int32)
    val = rng.uniform(-1, 1, size=nnz)
    x, y0 = rng.uniform(0.5, 1.5, n), rng.standard_normal(m)
    ver = {"v1": sb.V1, "v2": sb.V2, "baseline": sb.BASELINE, "bytes": sb.V1_BYTES}[version]
    nb, q = (max(1, nnz // (world * c)), 2) if version == "v2" else (0, 1)
    y, owned = _emulate(ver, rp, col, val, n, world, nb, q, x, y0)
    assert (owned == 1).all(), ("every row must be written exactly once", version, world, lens, owned.tolist())
    want = oracle.csr_spmv(rp, col, val, x, A, B, y0)
    bound = oracle.csr_spmv_bound(rp, col, val, x, A, B, y0)
    assert (np.abs(y - want) <= 1e-12 * bound + 1e-300).all(), (version, world, lens)
