"""The N > 1 (one process per GPU) path on the CPU: world_size 2 and 3 over gloo.
Each rank builds the REAL host layout of its rank plan (sblas_spmv_plan_create_rank with
SBLAS_LAYOUT_ONLY: partition, segments, edge slots, merge lists -- no GPU), a CPU stand-in
plays the kernels (per-segment csrmv on the clamped rows, raw partial sums for split rows),
the edge blocks are all-gathered exactly as bench.py does it, the merge lists are applied,
and the assembled y must match the oracle.  Covers v1, v2 (several tasks per rank) and
baseline, rows spanning three ranks, y != 0 and beta != 0."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
import sblas_b200 as sb
from conftest import GOLDEN, make_csr

A, B = 0.8401877171547095, 0.39438292681909304


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cases():
    g = np.load(os.path.join(GOLDEN, "qh768_coo.npz"))
    rp = oracle.coo_to_rowptr(int(g["m"]), g["row"])
    yield "qh768", rp, np.ascontiguousarray(g["col"]), np.ascontiguousarray(g["val"]), int(g["n"])
    rng = np.random.default_rng(5)
    lens = np.array([3, 1, 500, 2, 2, 90, 1, 4, 4, 0, 0, 7], np.int64)
    rp2, col2, val2 = make_csr(rng, len(lens), 300, lens)
    yield "long_row", rp2, col2, val2, 300


def _segment_cpu(rp, col, val, x, y0, seg, alpha, beta):
    """What the kernels compute for one segment: finished rows + raw edge partials."""
    lo, hi, nz0, nz1 = seg["row_lo"], seg["row_hi"], seg["nz0"], seg["nz1"]
    rows = np.arange(lo, hi + 1)
    b = np.clip(rp[rows], nz0, nz1)
    e = np.clip(rp[rows + 1], nz0, nz1)
    raw = np.array([np.dot(val[bi:ei], x[col[bi:ei]]) if ei > bi else 0.0 for bi, ei in zip(b, e)])
    out = alpha * raw + beta * y0[rows]
    edge = [0.0, 0.0]
    keep = np.ones(len(rows), bool)
    if seg["shared_first"]:
        edge[0] = raw[0]
        keep[0] = False
    if seg["shared_last"] and not (seg["shared_first"] and lo == hi):
        edge[1] = raw[-1]
        keep[-1] = False
    return rows, out, keep, edge


def _worker(rank, world, port, version, nb, q, result_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        worst = 0.0
        for name, rp, col, val, n in _cases():
            m, nnz = len(rp) - 1, int(rp[-1])
            rng = np.random.default_rng(11)                       # same on every rank
            x, y0 = rng.uniform(0.5, 1.5, n), rng.standard_normal(m)
            plan = sb.Plan.create_rank(version, m, n, nnz, 0, rp, 0, world, rank, 0, kernel=2,
                                       nb=(nnz // nb if nb else 0), q=q, flags=sb.LAYOUT_ONLY)
            slots = max(plan.edge_slots, 1)
            edge = torch.zeros(slots, dtype=torch.float64)
            y = np.zeros(m)
            owned = np.zeros(m)
            for seg in plan.local_segments():
                rows, out, keep, e = _segment_cpu(rp, col, val, x, y0, seg, A, B)
                y[rows[keep]] = out[keep]
                owned[rows[keep]] += 1
                edge[seg["edge_slot"]] = e[0]
                edge[seg["edge_slot"] + 1] = e[1]
            table = torch.zeros(world * slots, dtype=torch.float64)
            dist.all_gather_into_tensor(table, edge)              # the exchange step of bench.py
            segs = plan.local_segments()
            first_row = segs[0]["dev_first_row"] if segs else 0
            for lrow, offs in plan.merge_list():
                r = first_row + lrow
                y[r] = A * sum(float(table[o]) for o in offs) + B * y0[r]
                owned[r] += 1
            ty, to = torch.from_numpy(y), torch.from_numpy(owned)
            dist.all_reduce(ty)
            dist.all_reduce(to)
            assert (to.numpy() == 1).all(), "%s: every row must be written by exactly one rank" % name
            want = oracle.csr_spmv(rp, col, val, x, A, B, y0)
            bound = oracle.csr_spmv_bound(rp, col, val, x, A, B, y0)
            worst = max(worst, float((np.abs(ty.numpy() - want) / bound).max()))
            plan.destroy()
        if rank == 0:
            result_q.put(worst)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("version,nb,q", [(sb.V1, 0, 1), (sb.V2, 7, 2), (sb.BASELINE, 0, 1), (sb.V1_BYTES, 0, 1)])
def test_rank_sharded_spmv_over_gloo(world, version, nb, q):
    ctx = mp.get_context("spawn")
    result_q = ctx.SimpleQueue()
    mp.spawn(_worker, args=(world, _free_port(), version, nb, q, result_q), nprocs=world, join=True)
    worst = result_q.get()
    assert worst <= 1e-12, worst


def test_in_process_x_upload_covers_every_window():
    """Host arithmetic of the in-process multi-GPU upload (sblas_plan.c: sblas_spmv_plan_upload): GPU li of `live`
    uploads slice li of x over its own PCIe link, then pulls from every peer only the part of the peer's slice that
    lies inside the window of columns its own shard reads.  Whatever the windows are, every column of a GPU's window
    must end up in its replica exactly once, and nothing outside the window may be pulled."""
    import ctypes as C
    import sblas_b200 as sb
    L = sb.lib()
    LL = C.c_longlong
    L.sblas_x_slice.argtypes = [LL, C.c_int, C.c_int, C.POINTER(LL), C.POINTER(LL)]
    L.sblas_x_slice.restype = None
    L.sblas_x_pull_range.argtypes = [LL, LL, LL, LL, C.POINTER(LL), C.POINTER(LL)]
    L.sblas_x_pull_range.restype = None
    rng = np.random.default_rng(12)
    for trial in range(200):
        n = int(rng.integers(1, 5000)) if trial % 4 else int(rng.integers(1, 9))
        live = int(rng.integers(2, 9))
        slices = []
        for li in range(live):
            a, b = LL(), LL()
            L.sblas_x_slice(n, li, live, C.byref(a), C.byref(b))
            slices.append((a.value, b.value))
        assert slices[0][0] == 0 and slices[-1][1] == n and all(slices[i][1] == slices[i + 1][0] for i in range(live - 1))
        for d in range(live):
            w = np.sort(rng.integers(0, n, size=2))
            if trial % 7 == 0:
                w = np.array([0, n - 1])                     # SBLAS_X_WINDOW=0: all of x
            cover = np.zeros(n, np.int32)
            cover[slices[d][0]:slices[d][1]] += 1             # its own slice comes from the host
            for o in range(live):
                if o == d or slices[o][1] <= slices[o][0]:
                    continue
                a, b = LL(), LL()
                L.sblas_x_pull_range(slices[o][0], slices[o][1], int(w[0]), int(w[1]), C.byref(a), C.byref(b))
                if b.value > a.value:
                    assert slices[o][0] <= a.value and b.value <= slices[o][1]
                    assert w[0] <= a.value and b.value <= w[1] + 1
                    cover[a.value:b.value] += 1
            assert (cover[w[0]:w[1] + 1] == 1).all(), (n, live, d, w.tolist())
            assert (cover <= 1).all()
