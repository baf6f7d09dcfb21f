"""Parity of the CUDA path with the oracle, through the C-ABI (ctypes on
libsblas_spmv.so).  Tolerance (BASELINE.json north_star): per row
|y_gpu - y_oracle| <= 1e-12 * (|alpha| sum_j |a_ij||x_j| + |beta||y_i|).
Run with `-m gpu` on a B200."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
import sblas_b200 as sb
from conftest import check_tol, make_csr

pytestmark = pytest.mark.gpu
A, B = 0.8401877171547095, 0.39438292681909304        # the harness's ALPHA/BETA in f mode


def ngpus():
    import torch
    return torch.cuda.device_count()


def run_all_versions(rp, col, val, x, alpha, beta, y0, ngpu_list=(1,), kernels=(1, 2, 3), what=""):
    m, n, nnz = len(rp) - 1, len(x), int(rp[-1])
    want = oracle.csr_spmv(rp, col, val, x, alpha, beta, y0)
    bound = oracle.csr_spmv_bound(rp, col, val, x, alpha, beta, y0)
    for g in ngpu_list:
        y = y0.copy()
        assert sb.spMV_mgpu_baseline(m, n, nnz, alpha, val, rp, col, x, beta, y, g) == 0, sb.last_error()
        check_tol(y, want, bound, "%s baseline ngpu=%d" % (what, g))
        for k in kernels:
            y = y0.copy()
            assert sb.spMV_mgpu_v1(m, n, nnz, alpha, val, rp, col, x, beta, y, g, k) == 0, sb.last_error()
            check_tol(y, want, bound, "%s v1 ngpu=%d kernel=%d" % (what, g, k))
            for c in (1, 2, 8):
                nb = nnz // (g * c)
                if nb <= 0:
                    continue
                y = y0.copy()
                assert sb.spMV_mgpu_v2(m, n, nnz, alpha, val, rp, col, x, beta, y, g, k, nb, c) == 0, sb.last_error()
                check_tol(y, want, bound, "%s v2 ngpu=%d kernel=%d nb=%d q=%d" % (what, g, k, nb, c))
    return want


def gpu_counts():
    """GPU counts the in-process entry points are driven with.  SBLAS_EXPECT_GPUS=N makes a lease with
    fewer visible GPUs an error instead of a silently smaller sweep (the multi-GPU CI leg sets it)."""
    want = int(os.environ.get("SBLAS_EXPECT_GPUS", "0"))
    assert ngpus() >= want, "SBLAS_EXPECT_GPUS=%d but only %d GPU(s) visible" % (want, ngpus())
    return [g for g in (1, 2, 4, 8) if g <= ngpus()]


def test_qh768_harness_conditions(qh768):
    """Config 1: sample matrix as the harness loads it, x = 1, y = 0, harness alpha/beta."""
    x = np.ones(qh768["n"])
    run_all_versions(qh768["rowptr"], qh768["col"], qh768["val"], x, A, B, np.zeros(qh768["m"]),
                     gpu_counts(), what="qh768")


def test_qh768_nonzero_y_and_beta(qh768):
    """What the reference never tests: y != 0, beta != 0 through the split-row merge."""
    rng = np.random.default_rng(1)
    x = rng.uniform(0.5, 1.5, qh768["n"])
    y0 = rng.standard_normal(qh768["m"]) * 1e9
    run_all_versions(qh768["rowptr"], qh768["col"], qh768["val"], x, A, B, y0, gpu_counts(), what="qh768 y!=0")


def test_generator_g200_and_g10000():
    """INSTALL.md smoke input `g 200` and a batch_test.sh-size input, harness vectors."""
    for n in (200, 10000):
        r, c, v, alpha, beta = oracle.gen_g(n)
        rp = oracle.coo_to_rowptr(n, r)
        run_all_versions(rp, c, v, np.ones(n), alpha, beta, np.zeros(n), gpu_counts(),
                         kernels=(1, 2) if n > 200 else (1, 2, 3), what="g %d" % n)


@pytest.mark.parametrize("kind,ipt", [("vec", 16), ("tile", 4), ("tile", 8), ("tile", 16), ("tma", 16), ("vecp", 4), ("vecp", 8)])
def test_random_shapes_each_kernel_family(kind, ipt, monkeypatch):
    """Short rows, empty rows, long rows, rows much longer than a tile, unsorted/duplicate columns."""
    monkeypatch.setenv("SBLAS_KIND", kind)
    monkeypatch.setenv("SBLAS_IPT", str(ipt))
    rng = np.random.default_rng(ipt)
    shapes = {
        "short": rng.integers(1, 9, size=5000),
        "with_empty": rng.integers(0, 4, size=7000),
        "mixed": np.concatenate([rng.integers(0, 6, size=3000), [20000, 1, 0, 0, 9000], rng.integers(50, 300, size=200)]),
        "power_law": np.minimum((rng.pareto(1.2, size=4000) * 3).astype(np.int64) + 1, 60000),
        "all_long": rng.integers(2000, 9000, size=40),
        "many_empty_then_one": np.concatenate([np.zeros(10000, np.int64), [5], np.zeros(9000, np.int64)]),
        "leading_trailing_empty": np.concatenate([[0, 0, 0], rng.integers(1, 50, size=500), [0, 0]]),
        "medium_64_500": rng.integers(64, 500, size=600),
        "medium_180": np.full(900, 180, np.int64),
        "medium_with_gaps": np.concatenate([rng.integers(100, 400, size=200), [0, 0, 3000, 1, 0], rng.integers(150, 260, size=300)]),
        "rows_of_two": np.full(40000, 2, np.int64),
        "rows_of_one_and_empty": rng.integers(0, 2, size=30000),
    }
    for name, lens in shapes.items():
        m, n = len(lens), 4099
        rp, col, val = make_csr(rng, m, n, lens, sort_cols=(name != "mixed"))
        x, y0 = rng.standard_normal(n), rng.standard_normal(m)
        run_all_versions(rp, col, val, x, -1.75, 0.625, y0, (1,), kernels=(1,), what="%s/%s/%d" % (name, kind, ipt))
        run_all_versions(rp, col, val, x, 2.0, 0.0, y0, (1,), kernels=(1,), what="%s/%s/%d beta=0" % (name, kind, ipt))


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_tma_kernel_reduction_paths(mode, monkeypatch):
    """Every per-tile reduction path of the TMA kernel on the same inputs: 0 = production choice
    (A long rows / W warp pieces / M merge / S block), bit 0 = W off, bit 1 = M replaced by S.
    Shapes sit on the eligibility borders (<= 8 row starts per 256-entry chunk) and v2's small tasks
    put partial tiles (masked ranges, split first/last rows) through each path."""
    monkeypatch.setenv("SBLAS_KIND", "tma")
    monkeypatch.setenv("SBLAS_TMA_MODE", str(mode))
    rng = np.random.default_rng(100 + mode)
    shapes = {
        "border_24_48": rng.integers(24, 49, size=4000),
        "border_30_34": rng.integers(30, 35, size=5000),
        "rows_180": np.full(1500, 180, np.int64),
        "rows_100": np.full(2500, 100, np.int64),
        "medium_64_2000": rng.integers(64, 2001, size=300),
        "short_then_medium": np.concatenate([rng.integers(1, 6, size=20000), rng.integers(100, 300, size=800)]),
        "medium_with_long": np.concatenate([rng.integers(100, 300, size=500), [30000], rng.integers(40, 90, size=900)]),
        "rows_256_aligned": np.full(1000, 256, np.int64),
        "rows_2": np.full(60000, 2, np.int64),
    }
    for name, lens in shapes.items():
        m, n = len(lens), 6151
        rp, col, val = make_csr(rng, m, n, lens)
        x, y0 = rng.standard_normal(n), rng.standard_normal(m)
        run_all_versions(rp, col, val, x, -1.75, 0.625, y0, (1,), kernels=(1,), what="%s/mode %d" % (name, mode))
        run_all_versions(rp, col, val, x, 2.0, 0.0, y0, (1,), kernels=(1,), what="%s/mode %d beta=0" % (name, mode))


@pytest.mark.parametrize("short_max", [2, 4, 8])
def test_row_panels_pick_a_kernel_per_panel(short_max, monkeypatch):
    """Adaptive row binning at plan level (kernel = 1): 4096-row blocks whose rows all hold <= short_max
    entries go to the thread-per-row kernel, the rest to the TMA kernel; panels are cut at row
    borders inside v1 shards and v2 tasks (split first/last rows stay with the outer panels)."""
    monkeypatch.setenv("SBLAS_SHORT_MAX", str(short_max))
    monkeypatch.setenv("SBLAS_PANEL_MIN_NNZ", "2048")
    monkeypatch.setenv("SBLAS_VEC_BELOW", "1")
    rng = np.random.default_rng(200 + short_max)
    lens = np.concatenate([np.full(10000, 2, np.int64), np.full(5000, 150, np.int64), rng.integers(0, 5, size=9000),
                           np.full(100, 300, np.int64), np.full(5000, 2, np.int64), rng.integers(1, 9, size=6000),
                           [40000], np.full(9000, 1, np.int64)])
    m, n = len(lens), 8191
    rp, col, val = make_csr(rng, m, n, lens)
    x, y0 = rng.standard_normal(n), rng.standard_normal(m)
    run_all_versions(rp, col, val, x, -1.75, 0.625, y0, gpu_counts(), kernels=(1,), what="panels short_max=%d" % short_max)
    run_all_versions(rp, col, val, x, 2.0, 0.0, y0, (1,), kernels=(1,), what="panels beta=0")
    # the plan really is cut into several launches
    p = sb.Plan.create_rank(sb.V1, m, n, int(rp[-1]), val, rp, col, 1, 0, 0, kernel=1)
    assert p.launches >= 5, p.launches
    p.destroy()


def test_row_tile_kernel_on_medium_panels(monkeypatch):
    """Panels of medium rows (longest row of a 4096-row block in [32, 256]) go to the row-tile kernel,
    warp per R = floor(256/longest) whole rows.  One block per R = 1..8, with empty and very short
    rows sprinkled in, unsorted columns, split rows at shard/task borders (v1 x ngpu, v2 tasks)."""
    monkeypatch.setenv("SBLAS_PANEL_MIN_NNZ", "2048")
    monkeypatch.setenv("SBLAS_VEC_BELOW", "1")
    rng = np.random.default_rng(77)
    blocks = []
    for longest in (256, 128, 85, 64, 51, 42, 36, 32, 200, 100):
        ln = rng.integers(longest // 2 + 1, longest + 1, size=4096)
        ln[rng.integers(0, 4096, size=40)] = 0
        ln[rng.integers(0, 4096, size=40)] = rng.integers(1, 4, size=40)
        ln[rng.integers(0, 4096)] = longest
        blocks.append(ln)
    blocks.insert(3, np.full(5000, 2, np.int64))            # a short panel in between
    blocks.append(np.array([3000, 1, 0, 700], np.int64))    # and a general tail
    lens = np.concatenate(blocks).astype(np.int64)
    m, n = len(lens), 16381
    rp, col, val = make_csr(rng, m, n, lens, sort_cols=False)
    x, y0 = rng.standard_normal(n), rng.standard_normal(m)
    run_all_versions(rp, col, val, x, -1.75, 0.625, y0, gpu_counts(), kernels=(1,), what="row tiles")
    run_all_versions(rp, col, val, x, 2.0, 0.0, y0, (1,), kernels=(1,), what="row tiles beta=0")
    monkeypatch.setenv("SBLAS_MEDIUM", "0")
    run_all_versions(rp, col, val, x, -1.75, 0.625, y0, (1,), kernels=(1,), what="row tiles off")


def test_row_split_kernel_on_long_medium_panels(monkeypatch):
    """Panels whose longest row lies in (256, 2048] and whose rows are mostly that long go to the row-split
    kernel, G = 2 / 4 / 8 warps per row (SBLAS_MEDIUM bit 1).  One block per G, with empty and short rows
    sprinkled in, split rows at shard/task borders (v1 x ngpu, v2 tasks)."""
    monkeypatch.setenv("SBLAS_MEDIUM", "3")
    monkeypatch.setenv("SBLAS_PANEL_MIN_NNZ", "2048")
    monkeypatch.setenv("SBLAS_VEC_BELOW", "1")
    rng = np.random.default_rng(91)
    blocks = []
    for longest in (512, 300, 1024, 700, 2048, 1500):
        ln = rng.integers(longest // 2 + 1, longest + 1, size=4096)
        ln[rng.integers(0, 4096, size=30)] = 0
        ln[rng.integers(0, 4096, size=30)] = rng.integers(1, 40, size=30)
        ln[rng.integers(0, 4096)] = longest
        blocks.append(ln)
    blocks.insert(2, np.full(5000, 2, np.int64))
    blocks.append(np.array([30000, 1, 0, 700], np.int64))
    lens = np.concatenate(blocks).astype(np.int64)
    m, n = len(lens), 32749
    rp, col, val = make_csr(rng, m, n, lens, sort_cols=False)
    x, y0 = rng.standard_normal(n), rng.standard_normal(m)
    p = sb.Plan.create(sb.V1, m, n, int(rp[-1]), val, rp, col, 1, kernel=1)
    assert 7 in [u["kind"] for u in p.units()], p.units()
    p.destroy()
    run_all_versions(rp, col, val, x, -1.75, 0.625, y0, gpu_counts(), kernels=(1,), what="row split")
    run_all_versions(rp, col, val, x, 2.0, 0.0, y0, (1,), kernels=(1,), what="row split beta=0")


def test_row_spanning_many_segments():
    """A row longer than nnz/ngpu (v1) and than nb (v2): >= 3 segments share it."""
    rng = np.random.default_rng(17)
    lens = np.array([3, 1, 100000, 2, 2, 70000, 1], np.int64)
    rp, col, val = make_csr(rng, len(lens), 5000, lens)
    x, y0 = rng.uniform(0.1, 1.0, 5000), rng.standard_normal(len(lens))
    m, n, nnz = len(lens), 5000, int(rp[-1])
    want = oracle.csr_spmv(rp, col, val, x, A, B, y0)
    bound = oracle.csr_spmv_bound(rp, col, val, x, A, B, y0)
    for nb in (1000, 4096, 30000, 170010):
        for q in (1, 4):
            y = y0.copy()
            assert sb.spMV_mgpu_v2(m, n, nnz, A, val, rp, col, x, B, y, 1, 2, nb, q) == 0, sb.last_error()
            check_tol(y, want, bound, "nb=%d q=%d" % (nb, q))


def test_rank_plans_emulate_multi_gpu_on_one_device(qh768):
    """One-process-per-GPU mode: `world` rank plans (all on cuda:0 here, executed one after
    the other), every rank's edge block copied into the rank-major table (an NCCL all-gather
    in bench.py), then each owner's merge kernel."""
    rng = np.random.default_rng(23)
    cases = [(qh768["rowptr"], qh768["col"], qh768["val"], qh768["n"])]
    lens = np.array([3, 1, 50000, 2, 2, 9000, 1, 4, 4], np.int64)
    rp2, col2, val2 = make_csr(rng, len(lens), 3000, lens)
    cases.append((rp2, col2, val2, 3000))
    for rp, col, val, n in cases:
        m, nnz = len(rp) - 1, int(rp[-1])
        x, y0 = rng.uniform(0.5, 1.5, n), rng.standard_normal(m)
        want = oracle.csr_spmv(rp, col, val, x, A, B, y0)
        bound = oracle.csr_spmv_bound(rp, col, val, x, A, B, y0)
        for version, world, nb, q in ((sb.V1, 2, 0, 1), (sb.V1, 4, 0, 1), (sb.V1, 8, 0, 1), (sb.V2, 4, nnz // 13, 2),
                                      (sb.BASELINE, 4, 0, 1)):
            plans = [sb.Plan.create_rank(version, m, n, nnz, val, rp, col, world, r, 0, kernel=2, nb=nb, q=q)
                     for r in range(world)]
            slots = plans[0].edge_slots
            table = np.zeros(world * max(slots, 1))
            y = y0.copy()
            for r, p in enumerate(plans):
                p.execute(A, x, B, y)        # writes the rows rank r owns; split rows wait for the merge
                if slots:
                    sb.memcpy(table[r * slots:], p.edge_ptr(), 8 * slots, 2)
            d_table = plans[0].x_ptr()       # any device scratch of >= world*slots doubles: reuse rank 0's x
            assert world * slots <= n
            sb.memcpy(d_table, table, 8 * world * slots, 1)
            for r, p in enumerate(plans):
                p.merge_gathered(d_table, A, B)
                sb.device_synchronize()
                ptr, first, rows = p.y_ptr()
                if rows == 0:
                    continue
                ybuf = np.zeros(rows)
                sb.memcpy(ybuf, ptr, 8 * rows, 2)
                mine = [s for s in (p.segment(i) for i in range(p.num_segments))
                        if s["device"] == r and s["end_idx"] >= s["start_idx"]]
                skip = 1 if (version != sb.BASELINE and mine and mine[0]["start_flag"]) else 0
                y[first + skip:first + rows] = ybuf[skip:]
            check_tol(y, want, bound, "rank plans version=%d world=%d" % (version, world))
            for p in plans:
                p.destroy()


def test_fused_peer_exchange_on_one_device(qh768):
    """The NCCL-free exchange (P2P stores + epoch flags, edge_publish / edge_merge_wait kernels):
    `world` rank plans on cuda:0 with ordinary device buffers standing in for the peer-mapped
    tables.  Stepped in phases (all products, all publishes, all merges) so that no kernel
    ever waits for a kernel that has not run; three products in a row exercise the parity
    double-buffering and the ack back-pressure."""
    import torch
    rng = np.random.default_rng(31)
    lens = np.array([3, 1, 50000, 2, 2, 9000, 1, 4, 4], np.int64)
    rp2, col2, val2 = make_csr(rng, len(lens), 3000, lens)
    for rp, col, val, n in ((qh768["rowptr"], qh768["col"], qh768["val"], qh768["n"]), (rp2, col2, val2, 3000)):
        m, nnz = len(rp) - 1, int(rp[-1])
        x = rng.uniform(0.5, 1.5, n)
        for version, world, nb, q in ((sb.V1, 4, 0, 1), (sb.V1, 8, 0, 1), (sb.V2, 4, nnz // 13, 2)):
            plans = [sb.Plan.create_rank(version, m, n, nnz, val, rp, col, world, r, 0, kernel=2, nb=nb, q=q)
                     for r in range(world)]
            slots = plans[0].edge_slots
            tw = world * max(slots, 1)
            bufs = [torch.zeros(2 * tw + 2 * world, dtype=torch.float64, device="cuda") for _ in range(world)]
            for p in plans:
                p.bind_peer_tables([b.data_ptr() for b in bufs], tw)
            y = rng.standard_normal(m)
            for it in range(3):
                want = oracle.csr_spmv(rp, col, val, x, A, B, y)
                bound = oracle.csr_spmv_bound(rp, col, val, x, A, B, y)
                for p in plans:
                    p.upload(x, y)
                    p.execute_device(A, B, sync=True)
                for p in plans:
                    p.exchange_merge(A, B, phase=1)
                sb.device_synchronize()
                for p in plans:
                    p.exchange_merge(A, B, phase=2)
                sb.device_synchronize()
                ynew = y.copy()
                for p in plans:
                    p.download(ynew)
                check_tol(ynew, want, bound, "fused exchange version=%d world=%d product %d" % (version, world, it))
                y = ynew
            for p in plans:
                p.destroy()


def _device_synth_plan(m, n, lens, cols_mode, band, seed=7):
    """A resident single-GPU plan over a synthetic matrix generated ON the device at a size
    the oracle cannot check row by row; returns (plan, rp, keepalive)."""
    import torch
    rp = np.zeros(m + 1, np.int64)
    np.cumsum(lens, out=rp[1:])
    nnz = int(rp[-1])
    d_val = torch.empty(nnz, dtype=torch.float64, device="cuda")
    d_col = torch.empty(nnz, dtype=torch.int32, device="cuda")
    d_rp = torch.from_numpy(rp).cuda()
    sb.synth_fill_csr(d_rp.data_ptr(), 0, m, 0, nnz, n, cols_mode, band, seed, d_val.data_ptr(), d_col.data_ptr())
    torch.cuda.synchronize()
    plan = sb.Plan.create_rank(sb.V1, m, n, nnz, d_val.data_ptr(), rp, d_col.data_ptr(), 1, 0, 0, kernel=1,
                               flags=sb.SRC_DEVICE_SHARD, keep=(d_val, d_col))
    return plan, rp, (d_val, d_col, d_rp)


def test_full_size_properties_on_device_generated_matrix():
    """Size-independent properties at a size beyond the oracle's reach (~150M nnz, mixed row
    lengths incl. rows of 2, 100 and 9,000 nnz): run-to-run bit reproducibility (no FP atomics),
    linearity in x, the alpha/beta contract, and a sample of rows against the oracle."""
    rng = np.random.default_rng(3)
    lens = np.concatenate([np.full(10000, 9000, np.int64), np.full(400000, 100, np.int64),
                           np.full(1000000, 2, np.int64), rng.integers(1, 300, size=200000)])
    m = len(lens)
    n = 1 << 20
    plan, rp, keep = _device_synth_plan(m, n, lens, sb.COLS_BANDRUN, 1 << 16)
    yptr, first, rows = plan.y_ptr()
    assert (first, rows) == (0, m)

    def run(x, y0, alpha, beta):
        y = y0.copy()
        plan.execute(alpha, x, beta, y)
        return y

    x1, x2 = rng.standard_normal(n), rng.standard_normal(n)
    y0 = rng.standard_normal(m)
    a1 = run(x1, y0, 1.0, 0.0)
    assert (run(x1, y0, 1.0, 0.0) == a1).all(), "two runs must agree bit for bit"
    a2 = run(x2, y0, 1.0, 0.0)
    a12 = run(x1 + x2, y0, 1.0, 0.0)
    scale = np.abs(a1) + np.abs(a2) + 1e-300
    # |A|(|x1|+|x2|) >= |a1|+|a2|: linearity to a few ulps of the row bound
    absb = run(np.abs(x1) + np.abs(x2), y0, 1.0, 0.0)       # not the bound itself, only an order of magnitude
    assert (np.abs(a12 - (a1 + a2)) <= 1e-12 * (np.abs(absb) + scale) + 1e-9 * scale).all()
    full = run(x1, y0, -0.75, 0.5)
    assert np.allclose(full, -0.75 * a1 + 0.5 * y0, rtol=0, atol=1e-12 * (np.abs(a1) + np.abs(y0)).max())
    # sampled rows against the oracle
    d_val, d_col, _ = keep
    pick = np.unique(np.concatenate([[0, 9999, 10000, m - 1], rng.integers(0, m, size=400)]))
    for r in pick:
        b, e = int(rp[r]), int(rp[r + 1])
        vv, cc = np.empty(e - b), np.empty(e - b, np.int32)
        if e > b:
            sb.memcpy(vv, d_val.data_ptr() + 8 * b, 8 * (e - b), 2)
            sb.memcpy(cc, d_col.data_ptr() + 4 * b, 4 * (e - b), 2)
        lrp = np.array([0, e - b], np.int64)
        want = oracle.csr_spmv(lrp, cc, vv, x1, -0.75, 0.5, y0[r:r + 1])[0]
        bound = oracle.csr_spmv_bound(lrp, cc, vv, x1, -0.75, 0.5, y0[r:r + 1])[0]
        assert abs(full[r] - want) <= 1e-12 * bound + 1e-300, r
    plan.destroy()


def test_versions_and_kernels_agree_on_generator_matrix():
    """baseline, v1 and v2 (several task sizes / streams) and kernels 1-3 on `g 4000`: all within
    the tolerance of the oracle and of each other (what the harness's Y/N column checks)."""
    r, c, v, alpha, beta = oracle.gen_g(4000)
    rp = oracle.coo_to_rowptr(4000, r)
    x, y0 = np.ones(4000), np.zeros(4000)
    run_all_versions(rp, c, v, x, alpha, beta, y0, gpu_counts(), kernels=(1, 2, 3), what="g 4000")


def test_device_row_pointer_is_the_reference_local_row_pointer(qh768):
    """The int32 row pointer a GPU works on is built ON the GPU from the int64 slice
    (rebase_rowptr_kernel); it must equal, entry for entry, the local row pointer the reference
    builds on the host for the same shard (dspmv_mgpu_v1.cu:125-133, dspmv_mgpu_baseline.cu:82-85)."""
    rp, col, val = qh768["rowptr"], qh768["col"], qh768["val"]
    m, n, nnz = qh768["m"], qh768["n"], qh768["nnz"]
    for world in (1, 2, 4, 8):
        parts = oracle.partition_v1(rp, world)
        bparts = oracle.partition_baseline(rp, world)
        for r in range(world):
            for version in (sb.V1, sb.BASELINE):
                p = sb.Plan.create_rank(version, m, n, nnz, val, rp, col, world, r, 0, kernel=2)
                ptr, cnt = p.rowptr_ptr()
                got = np.zeros(cnt, np.int32)
                sb.memcpy(got, ptr, 4 * cnt, 2)
                if version == sb.V1:
                    want = oracle.local_rowptr_v1(rp, parts["start_idx"][r], parts["start_row"][r],
                                                  parts["dev_m"][r], parts["dev_nnz"][r])
                else:
                    want = oracle.local_rowptr_baseline(rp, bparts["start_row"][r], bparts["dev_m"][r])
                assert cnt == len(want) and (got == want).all(), (world, r, version)
                p.destroy()


def test_one_shot_plan_cache(qh768, monkeypatch):
    """SBLAS_PLAN_CACHE=1: repeated one-shot calls on the same host arrays reuse the upload; a
    different x / y / alpha / beta still gives the right answer; an in-place edit of csrVal is
    picked up after cache_clear()."""
    monkeypatch.setenv("SBLAS_PLAN_CACHE", "1")
    rng = np.random.default_rng(41)
    r, c, v, _, _ = oracle.gen_g(2000)
    rp = oracle.coo_to_rowptr(2000, r)
    v = v.copy()
    for it in range(4):
        x, y0 = rng.standard_normal(2000), rng.standard_normal(2000)
        al, be = float(rng.uniform(0.1, 2)), float(rng.uniform(0.1, 2))
        for fn, extra in ((sb.spMV_mgpu_v1, (1, 2)), (sb.spMV_mgpu_v2, (1, 2, len(v) // 4, 2)), (sb.spMV_mgpu_baseline, (1,))):
            y = y0.copy()
            assert fn(2000, 2000, len(v), al, v, rp, c, x, be, y, *extra) == 0, sb.last_error()
            check_tol(y, oracle.csr_spmv(rp, c, v, x, al, be, y0), oracle.csr_spmv_bound(rp, c, v, x, al, be, y0), "cache it %d" % it)
    v[1000:2000] *= 3.0                       # in-place edit in the middle: invisible to the fingerprint
    sb.cache_clear()
    x, y0 = rng.standard_normal(2000), rng.standard_normal(2000)
    y = y0.copy()
    assert sb.spMV_mgpu_v1(2000, 2000, len(v), 1.0, v, rp, c, x, 0.5, y, 1, 1) == 0
    check_tol(y, oracle.csr_spmv(rp, c, v, x, 1.0, 0.5, y0), oracle.csr_spmv_bound(rp, c, v, x, 1.0, 0.5, y0), "after clear")
    sb.cache_clear()


def test_one_shot_calls_reuse_a_retained_allocation(monkeypatch):
    """The one-shot entry points keep one main allocation per GPU between calls (a cudaMalloc + cudaFree pair of a
    shard costs more than the product once peer access is on).  Matrices of growing and shrinking size, every version
    and GPU count, must give bit-for-bit what the same calls give with the pool off (SBLAS_POOL=0); cache_clear()
    hands the memory back."""
    import torch
    r, c, v, _, _ = oracle.gen_g(200)              # warm call: kernels loaded before free memory is read
    y = np.zeros(200)
    assert sb.spMV_mgpu_v1(200, 200, len(v), A, v, oracle.coo_to_rowptr(200, r), c, np.ones(200), B, y, 1, 1) == 0
    sb.cache_clear()
    results = {}
    for pool in ("1", "0"):
        monkeypatch.setenv("SBLAS_POOL", pool)
        for idx, n in enumerate((3000, 200, 9000, 56, 9000, 1200)):          # up, down, up, tiny, same again, mid
            r, c, v, _, _ = oracle.gen_g(n)
            rp = oracle.coo_to_rowptr(n, r)
            rs = np.random.default_rng(1000 + idx)
            x, y0 = rs.standard_normal(n), rs.standard_normal(n)
            for g in gpu_counts():
                for name, fn, extra in (("v1", sb.spMV_mgpu_v1, (g, 1)), ("v2", sb.spMV_mgpu_v2, (g, 1, max(1, len(v) // (3 * g)), 2)),
                                        ("base", sb.spMV_mgpu_baseline, (g,))):
                    y = y0.copy()
                    assert fn(n, n, len(v), A, v, rp, c, x, B, y, *extra) == 0, sb.last_error()
                    if pool == "1":
                        check_tol(y, oracle.csr_spmv(rp, c, v, x, A, B, y0), oracle.csr_spmv_bound(rp, c, v, x, A, B, y0),
                                  "pool n=%d g=%d %s" % (n, g, name))
                        results[(idx, g, name)] = y
                    else:
                        assert (y == results[(idx, g, name)]).all(), ("pooled and unpooled differ", n, g, name)
        if pool == "1":
            free_pooled = torch.cuda.mem_get_info(0)[0]
            sb.cache_clear()
            # GPU 0's retained block is the one-GPU shard of the 9,000-row matrix (9.8 M entries x 12 B = 118 MB)
            assert torch.cuda.mem_get_info(0)[0] - free_pooled >= (100 << 20), "cache_clear() must give the block back"


def test_against_cusparse_generic_spmv(qh768):
    """Third opinion on the arithmetic: the reference's csrmv is legacy cuSPARSE (removed in CUDA 11);
    its living successor, the generic cusparseSpMV that torch's sparse CSR mat-vec calls, must agree
    with this library (and with the oracle) to the same per-row bound.  Test-only use of a library."""
    import torch
    rng = np.random.default_rng(61)
    lens = np.concatenate([rng.integers(0, 6, size=3000), rng.integers(100, 300, size=500), [20000, 0, 9000],
                           np.full(5000, 2, np.int64), rng.integers(40, 120, size=5000)])
    rp2, col2, val2 = make_csr(rng, len(lens), 7001, lens)
    for rp, col, val, n in ((qh768["rowptr"], qh768["col"], qh768["val"], qh768["n"]), (rp2, col2, val2, 7001)):
        m, nnz = len(rp) - 1, int(rp[-1])
        x, y0 = rng.uniform(0.5, 1.5, n), rng.standard_normal(m)
        A_t = torch.sparse_csr_tensor(torch.from_numpy(rp), torch.from_numpy(col.astype(np.int64)), torch.from_numpy(val),
                                      size=(m, n), dtype=torch.float64, device="cuda")
        ax = (A_t @ torch.from_numpy(x).cuda()).cpu().numpy()
        lib_y = y0.copy()
        assert sb.spMV_mgpu_v1(m, n, nnz, A, val, rp, col, x, B, lib_y, 1, 1) == 0, sb.last_error()
        bound = oracle.csr_spmv_bound(rp, col, val, x, A, B, y0)
        check_tol(lib_y, A * ax + B * y0, 2.0 * bound, "vs cusparse generic SpMV")
        check_tol(A * ax + B * y0, oracle.csr_spmv(rp, col, val, x, A, B, y0), 2.0 * bound, "cusparse vs oracle")


def test_chained_products_stay_on_the_devices():
    """Iterative use: y -> x on every GPU with sblas_spmv_plan_chain (NVLink
    all-gather of the owned row slices, no host copy), then the next product.  Every product is
    checked against the oracle applied to the previous result; the gathered x is bit-identical to
    the y that was downloaded."""
    rng = np.random.default_rng(71)
    lens = np.concatenate([rng.integers(1, 9, size=3000), [40000, 3, 25000], rng.integers(60, 200, size=800)])
    m = len(lens)
    rp, col, val = make_csr(rng, m, m, lens)
    val *= 0.05
    nnz = int(rp[-1])
    for g in gpu_counts():
        for version, nb, q in ((sb.V1, 0, 1), (sb.V2, nnz // (3 * g) + 1, 2)):
            p = sb.Plan.create(version, m, m, nnz, val, rp, col, g, kernel=1, nb=nb, q=q)
            x = rng.uniform(0.5, 1.5, m)
            y = np.zeros(m)
            p.upload(x, None)
            for it in range(4):
                p.execute_device(1.25, 0.0, sync=False)
                y_dev = np.zeros(m)
                p.download(y_dev)
                want = oracle.csr_spmv(rp, col, val, x, 1.25, 0.0, y)
                check_tol(y_dev, want, oracle.csr_spmv_bound(rp, col, val, x, 1.25, 0.0, y), "chain it %d g %d" % (it, g))
                p.chain()
                sb.device_synchronize()
                for d in range(p.num_devices):
                    xd = np.zeros(m)
                    sb.memcpy(xd, p.x_ptr(d), 8 * m, 2)
                    assert (xd == y_dev).all(), (it, g, d)
                x = y_dev
            p.destroy()


def test_upload_moves_only_the_columns_a_shard_reads():
    """A one-GPU (or one-rank) plan copies x[first_col..last_col] of its shard only; stale x outside the
    window must not matter, and the window is exact."""
    rng = np.random.default_rng(83)
    m, n = 3000, 50000
    lens = rng.integers(1, 40, size=m)
    rp = np.zeros(m + 1, np.int64)
    np.cumsum(lens, out=rp[1:])
    nnz = int(rp[-1])
    col = rng.integers(12345, 23456, size=nnz, dtype=np.int64).astype(np.int32)
    col[7], col[nnz // 2] = 12000, 23999
    val = rng.uniform(-1, 1, size=nnz)
    p = sb.Plan.create(sb.V1, m, n, nnz, val, rp, col, 1, kernel=1)
    assert p.x_window() == (12000, 23999)
    for it in range(2):
        x, y0 = rng.standard_normal(n), rng.standard_normal(m)
        y = y0.copy()
        p.execute(A, x, B, y)
        check_tol(y, oracle.csr_spmv(rp, col, val, x, A, B, y0), oracle.csr_spmv_bound(rp, col, val, x, A, B, y0), "x window %d" % it)
    p.destroy()


def test_stage_release_hazard_regression():
    """A stage of the bulk-copy ring must not go back to the producer before the values read from it have
    arrived.  tests/hazard_stress.py drives every early-releasing kernel with fully
    scattered gathers at ~226 M entries and compares ALL rows with the independent kernel 3.  The production
    library must be clean; the same sources built with the pre-fix release (lib/libsblas_spmv_unsafe.so,
    -DSBLAS_UNSAFE_EARLY_RELEASE) are run too and the outcome is printed: when that build shows wrong rows
    the input is proven to catch the hazard (it is timing dependent, so a clean run of the unsafe build is
    not a failure of this test)."""
    import json
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    libdir = os.path.join(os.path.dirname(here), "s-blas_b200", "lib")

    def run(lib):
        env = dict(os.environ, SBLAS_LIB=os.path.join(libdir, lib))
        r = subprocess.run([sys.executable, os.path.join(here, "hazard_stress.py")], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        return json.loads(r.stdout.strip().splitlines()[-1])
    good = run("libsblas_spmv.so")
    print("production:", good)
    assert {3, 6, 7} <= set(good["kinds"]), good          # general (A, W), row-tile and row-split panels all ran
    assert good["bad"] == 0, good
    if os.path.exists(os.path.join(libdir, "libsblas_spmv_unsafe.so")):
        print("pre-fix build:", run("libsblas_spmv_unsafe.so"))


def test_graph_replay_of_a_product():
    """sblas_spmv_plan_step: the first product runs as plain launches, the second is captured into a CUDA graph,
    later ones replay it; y feeds back (beta != 0) so every product is checked against the oracle applied to the
    previous result; a change of alpha / beta re-captures.  SBLAS_GRAPH=0 must give bit-identical results."""
    rng = np.random.default_rng(97)
    lens = np.concatenate([rng.integers(1, 9, size=9000), [40000, 3, 25000], rng.integers(60, 200, size=8000),
                           np.full(8192, 2, np.int64)])
    m, n = len(lens), 20011
    rp, col, val = make_csr(rng, m, n, lens)
    val *= 0.01
    nnz = int(rp[-1])
    x = rng.uniform(0.5, 1.5, n)
    results = {}
    for graph in ("1", "0"):
        os.environ["SBLAS_GRAPH"] = graph
        try:
            p = sb.Plan.create(sb.V1, m, n, nnz, val, rp, col, 1, kernel=1)
            y = rng.standard_normal(m) if graph == "1" else results["y0"].copy()
            results.setdefault("y0", y.copy())
            p.upload(x, y)
            outs = []
            for it, (al, be) in enumerate([(A, B), (A, B), (A, B), (A, B), (1.5, 0.25), (1.5, 0.25), (A, B)]):
                want = oracle.csr_spmv(rp, col, val, x, al, be, y)
                bound = oracle.csr_spmv_bound(rp, col, val, x, al, be, y)
                p.step(al, be)
                got = np.zeros(m)
                p.download(got)
                check_tol(got, want, bound, "step %d graph=%s" % (it, graph))
                y = got
                outs.append(got)
            results[graph] = outs
            p.destroy()
        finally:
            os.environ.pop("SBLAS_GRAPH", None)
    for a, b in zip(results["1"], results["0"]):
        assert (a == b).all(), "graph replay and stream launches must agree bit for bit"


def test_rank_chain_over_peer_x_on_one_device(monkeypatch):
    """The one-process-per-GPU form of the chain, emulated with `world` rank plans on cuda:0 and
    ordinary device buffers standing in for the peer-mapped x / flag / exchange buffers.  Per product: the segments,
    the fused split-row exchange (publish, then merge -- stepped in phases, because on ONE device a kernel that
    spins on a flag must never wait for a kernel the host has not launched yet; graph replay is off for the same
    reason: instantiating a graph may block on the device) and sblas_spmv_plan_chain (all-gather of the owned y
    rows into EVERY rank's x with stores + epoch flags; its three small kernels per rank are launched back to back for
    all ranks).  After every product every rank's x equals the others' bit for bit and the oracle's y within
    tolerance.  The real multi-process path (graph replay included) is exercised by bench.py's rank_chain leg."""
    import torch
    monkeypatch.setenv("SBLAS_GRAPH", "0")
    rng = np.random.default_rng(113)
    lens = np.concatenate([rng.integers(1, 9, size=3000), [40000, 3, 25000], rng.integers(60, 200, size=800)])
    m = len(lens)
    rp, col, val = make_csr(rng, m, m, lens)
    val *= 0.05
    nnz = int(rp[-1])
    for world in (2, 4):
        plans = [sb.Plan.create_rank(sb.V1, m, m, nnz, val, rp, col, world, r, 0, kernel=1) for r in range(world)]
        slots = plans[0].edge_slots
        tw = world * max(slots, 1)
        tables = [torch.zeros(2 * tw + 2 * world, dtype=torch.float64, device="cuda") for _ in range(world)]
        xs = [torch.zeros(m, dtype=torch.float64, device="cuda") for _ in range(world)]
        flags = [torch.zeros(2 * world, dtype=torch.int64, device="cuda") for _ in range(world)]
        x = rng.uniform(0.5, 1.5, m)
        for p in plans:
            p.bind_peer_tables([t.data_ptr() for t in tables], tw)
            p.bind_peer_x([t.data_ptr() for t in xs], [f.data_ptr() for f in flags])
            p.upload(x, None)
        sb.device_synchronize()
        for it in range(4):
            want = oracle.csr_spmv(rp, col, val, x, 1.25, 0.0, np.zeros(m))
            bound = oracle.csr_spmv_bound(rp, col, val, x, 1.25, 0.0, np.zeros(m))
            for p in plans:
                p.execute_device(1.25, 0.0, sync=True)
            for p in plans:
                p.exchange_merge(1.25, 0.0, phase=1)
            sb.device_synchronize()
            for p in plans:
                p.exchange_merge(1.25, 0.0, phase=2)
            sb.device_synchronize()
            for p in plans:                          # kernel launches only: all ranks' flag kernels get to run together
                p.chain()
            sb.device_synchronize()
            got = [t.cpu().numpy() for t in xs]
            for r in range(1, world):
                assert (got[r] == got[0]).all(), (world, it, r)
            check_tol(got[0], want, bound, "rank chain world=%d product %d" % (world, it))
            x = got[0]
        for p in plans:
            p.destroy()


def test_byte_balanced_version(qh768):
    """SBLAS_V1_BYTES (opt-in, not in the reference): same machinery as v1 on differently placed cuts -- in-process
    plans on every visible GPU count and rank plans emulating 8 GPUs with the gathered-table merge."""
    rng = np.random.default_rng(131)
    lens = np.concatenate([np.full(3000, 180, np.int64), np.full(60000, 2, np.int64), [30000], rng.integers(1, 60, size=4000)])
    rp, col, val = make_csr(rng, len(lens), 9001, lens)
    m, n, nnz = len(lens), 9001, int(rp[-1])
    x, y0 = rng.standard_normal(n), rng.standard_normal(m)
    want = oracle.csr_spmv(rp, col, val, x, A, B, y0)
    bound = oracle.csr_spmv_bound(rp, col, val, x, A, B, y0)
    for g in gpu_counts():
        p = sb.Plan.create(sb.V1_BYTES, m, n, nnz, val, rp, col, g, kernel=1)
        y = y0.copy()
        p.execute(A, x, B, y)
        check_tol(y, want, bound, "byte-balanced in-process ngpu=%d" % g)
        p.destroy()
    world = 8
    plans = [sb.Plan.create_rank(sb.V1_BYTES, m, n, nnz, val, rp, col, world, r, 0, kernel=1) for r in range(world)]
    slots = plans[0].edge_slots
    table = np.zeros(world * max(slots, 1))
    y = y0.copy()
    for r, p in enumerate(plans):
        p.execute(A, x, B, y)
        if slots:
            sb.memcpy(table[r * slots:], p.edge_ptr(), 8 * slots, 2)
    d_table = plans[0].x_ptr()
    sb.memcpy(d_table, table, 8 * world * slots, 1)
    parts = sb.partition_bytes(rp, world)
    for r, p in enumerate(plans):
        p.merge_gathered(d_table, A, B)
        sb.device_synchronize()
        ptr, first, rows = p.y_ptr()
        if rows == 0:
            continue
        ybuf = np.zeros(rows)
        sb.memcpy(ybuf, ptr, 8 * rows, 2)
        skip = 1 if parts["start_flag"][r] else 0
        y[first + skip:first + rows] = ybuf[skip:]
    check_tol(y, want, bound, "byte-balanced rank plans world=8")
    for p in plans:
        p.destroy()
