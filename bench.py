#!/usr/bin/env python
"""bench.py -- double CSR SpMV (y = alpha*A*x + beta*y) GFLOP/s and achieved HBM GB/s on
1/2/4/8 B200, for the hot path of pnnl/s-blas named by BASELINE.json.

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]
    (N > 1: launched by torchrun, one rank per GPU, NCCL)

A step is one SpMV of the named synthetic matrix, sharded with the reference's v1
nnz-balanced partition (spmv/src/dspmv_mgpu_v1.cu:59-100) over the N ranks (strong
scaling: the matrix is fixed).  Per step every rank launches the tile kernel + its fix-up
on its resident shard; for N > 1 the raw partial sums of rows split between ranks (<= 2
doubles per rank) are all-gathered (NCCL) and each owner finishes its split rows in
ascending rank order.  value = 2*nnz / max-over-ranks step time.

Timing: CUDA events on the plan's own stream, W >= 3 warm-up steps, exactly K timed steps
between barrier + synchronize, max over ranks.  The inputs (>= 14 GB at the default
workload) are far larger than the 126 MB L2, so no flush is needed between iterations.

The oracle (oracle/) is used here only (a) to CHECK a sample of rows of the result before
timing and (b) as the timed CPU baseline (cpu_baseline, and the whole of --impl reference).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALPHA, BETA = 0.8401877171547095, 0.39438292681909304      # the harness's ALPHA/BETA (glibc rand(), seed 1)
if os.environ.get("SBLAS_BENCH_BETA"):                      # diagnostics only
    BETA = float(os.environ["SBLAS_BENCH_BETA"])
SEED = 20260318


# ----------------------------------------------------------------------------- workloads
def workload(name):
    """Returns dict(m, n, row_len (callable -> int64 array of length m), cols_mode, band, desc)."""
    import sblas_b200 as sb

    def two_block(m, n1, l1, l2):
        def f():
            a = np.empty(m, np.int64)
            a[:n1] = l1
            a[n1:] = l2
            return a
        return f

    if name == "g1m":          # BASELINE config 2, reading 2b
        m = 1_000_000
        return dict(m=m, n=m, row_len=two_block(m, m // 8, 9000, 100), cols_mode=sb.COLS_PREFIX, band=0,
                    desc="test_spmv g shape at n=1,000,000 rows, densities scaled 1/100 (125,000 rows x 9,000 nnz + "
                         "875,000 rows x 100 nnz = 1,212,500,000 nnz, cols 0..k-1 per row; the literal g 1000000 "
                         "would be 1.2e11 nnz = 1.46 TB)")
    if name == "g100000":      # BASELINE config 2, reading 2a: the literal generator at the largest feasible n
        m = 100_000
        return dict(m=m, n=m, row_len=two_block(m, m // 8, 90000, 1000), cols_mode=sb.COLS_PREFIX, band=0,
                    desc="test_spmv g 100000 (12,500 rows x 90,000 nnz + 87,500 rows x 1,000 nnz = 1,212,500,000 nnz)")
    if name == "big50m":       # BASELINE config 5
        m = 50_000_000
        return dict(m=m, n=m, row_len=two_block(m, m // 8, 180, 2), cols_mode=sb.COLS_BANDRUN, band=1 << 20,
                    desc="50M-row ~1.2B-nnz non-uniform (6.25M rows x 180 nnz + 43.75M rows x 2 nnz), banded columns "
                         "+-2^20 in runs of 16 consecutive columns")
    if name == "big50m_scatter":
        m = 50_000_000
        return dict(m=m, n=m, row_len=two_block(m, m // 8, 180, 2), cols_mode=sb.COLS_BANDED, band=1 << 20,
                    desc="50M-row ~1.2B-nnz non-uniform, banded columns +-2^20, every entry in its own cache line "
                         "(adversarial for the L1 gather path)")
    if name == "big50m_uniform":
        m = 50_000_000
        return dict(m=m, n=m, row_len=two_block(m, m // 8, 180, 2), cols_mode=sb.COLS_UNIFORM, band=0,
                    desc="50M-row ~1.2B-nnz non-uniform, uniform columns (adversarial x traffic)")
    if name == "rows180":      # diagnostic: the long-row block of big50m alone
        m = 6_250_000
        return dict(m=m, n=50_000_000, row_len=two_block(m, m, 180, 180), cols_mode=sb.COLS_BANDRUN, band=1 << 20,
                    desc="diagnostic: 6.25M rows x 180 nnz, banded runs")
    if name == "rows100":      # diagnostic: the short-row block of g1m alone (x12 rows to fill the GPU)
        m = 10_500_000
        return dict(m=m, n=1_000_000, row_len=two_block(m, m, 100, 100), cols_mode=sb.COLS_PREFIX, band=0,
                    desc="diagnostic: 10.5M rows x 100 nnz, prefix columns")
    if name == "rows9000":     # diagnostic: the long-row block of g1m alone
        m = 125_000
        return dict(m=m, n=1_000_000, row_len=two_block(m, m, 9000, 9000), cols_mode=sb.COLS_PREFIX, band=0,
                    desc="diagnostic: 125,000 rows x 9,000 nnz, prefix columns")
    if name == "rows9000b":    # diagnostic: long rows with banded-run columns (x from L2, not L1)
        m = 125_000
        return dict(m=m, n=50_000_000, row_len=two_block(m, m, 9000, 9000), cols_mode=sb.COLS_BANDRUN, band=1 << 20,
                    desc="diagnostic: 125,000 rows x 9,000 nnz, banded runs")
    if name == "rows180p":     # diagnostic: rows of 180 with prefix columns (x from L1)
        m = 6_250_000
        return dict(m=m, n=1_000_000, row_len=two_block(m, m, 180, 180), cols_mode=sb.COLS_PREFIX, band=0,
                    desc="diagnostic: 6.25M rows x 180 nnz, prefix columns")
    if name == "rows1000":     # diagnostic: the short-row block of g100000 alone (x12 rows)
        m = 1_050_000
        return dict(m=m, n=1_000_000, row_len=two_block(m, m, 1000, 1000), cols_mode=sb.COLS_PREFIX, band=0,
                    desc="diagnostic: 1.05M rows x 1,000 nnz, prefix columns")
    if name == "rows2":        # diagnostic: the short-row block of big50m alone
        m = 43_750_000
        return dict(m=m, n=50_000_000, row_len=two_block(m, m, 2, 2), cols_mode=sb.COLS_BANDRUN, band=1 << 20,
                    desc="diagnostic: 43.75M rows x 2 nnz, banded runs")
    if name == "circuit5m":    # BASELINE config 3
        m = 5_558_326

        def f():
            rng = np.random.default_rng(SEED)
            u = rng.random(m)
            ln = np.floor(2.0 * (1.0 - u) ** (-1.0 / 1.35)).astype(np.int64) + 2       # truncated power law, median ~5
            ln = np.minimum(ln, 200_000)
            hubs = rng.choice(m, size=12, replace=False)
            ln[hubs] = np.array([1_290_000, 620_000, 410_000, 300_000, 240_000, 200_000, 170_000, 150_000,
                                 130_000, 120_000, 110_000, 105_000])
            return ln
        return dict(m=m, n=m, row_len=f, cols_mode=sb.COLS_CIRCUIT, band=1 << 16,
                    desc="Circuit5M-shaped power law (5,558,326 rows, ~59.5M nnz, max row 1.29M, 80% banded / 20% uniform columns)")
    if name == "rail4284":     # BASELINE config 4
        m, n = 4284, 1_092_610

        def f():
            rng = np.random.default_rng(SEED + 1)
            ln = np.exp(rng.normal(7.45, 0.95, size=m)).astype(np.int64) + 1
            return np.clip(ln, 1, 56_000)
        return dict(m=m, n=n, row_len=f, cols_mode=sb.COLS_UNIFORM, band=0,
                    desc="rail4284-shaped short-wide (4,284 x 1,092,610, ~11.3M nnz, log-normal row lengths, uniform columns)")
    raise SystemExit("unknown workload " + name)


def host_sample(wl, frac_rows):
    """A bounded row sample of the workload on the HOST for the CPU legs: every block of the
    two-block shapes keeps its share of rows; values uniform(0,1).  Only PREFIX-column shapes
    (the g generator) are sampled structurally; other shapes take the first rows."""
    import sblas_b200 as sb
    lens = wl["row_len"]()
    m = len(lens)
    take = max(8, int(m * frac_rows))
    idx = np.unique(np.linspace(0, m - 1, take).astype(np.int64))
    sl = lens[idx]
    rp = np.zeros(len(sl) + 1, np.int64)
    rp[1:] = np.cumsum(sl)
    nnz = int(rp[-1])
    rng = np.random.default_rng(SEED)
    val = rng.random(nnz)
    if wl["cols_mode"] == sb.COLS_PREFIX:
        col = (np.arange(nnz, dtype=np.int64) - np.repeat(rp[:-1], sl)).astype(np.int32)
    else:
        col = rng.integers(0, wl["n"], size=nnz, dtype=np.int64).astype(np.int32)
    x = rng.random(wl["n"])
    y = rng.random(len(sl))
    return rp, col, val, x, y


def time_cpu(rp, col, val, x, y, reps, fn_name):
    import oracle
    L = oracle.lib()
    fn = getattr(L, fn_name)
    m = len(rp) - 1
    yy = y.copy()
    nt = fn(m, rp, col, val, x, ALPHA, BETA, yy)          # warm-up
    best = 1e30
    for _ in range(reps):
        yy[:] = y
        t0 = time.perf_counter()
        fn(m, rp, col, val, x, ALPHA, BETA, yy)
        best = min(best, time.perf_counter() - t0)
    return best, (nt if fn_name != "oracle_csr_spmv" else 1)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.p = None
        self.lines = []
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _pump(self):
        for line in self.p.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        for ts, ln in self.lines:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                mx.append(float(f[2]))
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[1]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                pass
        if not sm:      # timed window shorter than the sampling period: use every sample taken under load
            for ts, ln in self.lines:
                f = [c.strip() for c in ln.split(",")]
                try:
                    if ts >= t0 - 3.0:
                        sm.append(float(f[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """--impl reference: the reference has no CPU SpMV of its own and its GPU path does not
    compile against CUDA 12.9, so this arm times the oracle restatement
    of its csrmv semantics on all host cores, on a bounded row sample of the same workload."""
    if rank != 0:
        return
    wl = workload(args.workload)
    rp, col, val, x, y = host_sample(wl, args.cpu_sample)
    nnz = int(rp[-1])
    for _ in range(args.warmup):
        time_cpu(rp, col, val, x, y, 1, "oracle_csr_spmv_omp_balanced")
    t0 = time.perf_counter()
    best, cores = time_cpu(rp, col, val, x, y, args.steps, "oracle_csr_spmv_omp_balanced")
    wall = (time.perf_counter() - t0) / max(args.steps, 1)
    gf = 2.0 * nnz / best / 1e9
    sample = "%d of %d rows (%.0f%%, every block keeps its share), %d nnz" % (len(rp) - 1, wl["m"], 100 * args.cpu_sample, nnz)
    print(json.dumps({
        "impl": "reference", "metric": "double CSR SpMV GFLOP/s", "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": best * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "name": args.workload, "alpha": ALPHA, "beta": BETA},
        "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample,
                         "gbs": (12.0 * nnz + 20.0 * (len(rp) - 1)) / best / 1e9, "wall_ms_per_step": wall * 1e3},
        "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sblas")
    ap.add_argument("--workload", default=os.environ.get("SBLAS_BENCH_WORKLOAD", "g1m"))
    ap.add_argument("--kernel", type=int, default=1)
    ap.add_argument("--cpu-sample", type=float, default=0.1, help="fraction of rows in the CPU legs' sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--exchange", default="symm", choices=["symm", "nccl"],
                    help="split-row exchange for N>1: fused P2P over symmetric memory, or an NCCL all-gather")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import sblas_b200 as sb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl = workload(args.workload)
    m, n = wl["m"], wl["n"]
    lens = wl["row_len"]()
    rp = np.zeros(m + 1, np.int64)
    np.cumsum(lens, out=rp[1:])
    nnz = int(rp[-1])
    parts = sb.partition_v1(rp, world)
    s_idx, e_idx = int(parts["start_idx"][rank]), int(parts["end_idx"][rank])
    s_row, e_row = int(parts["start_row"][rank]), int(parts["end_row"][rank])
    dnnz = e_idx - s_idx + 1

    # ---- this rank's shard, generated on the GPU (synthetic, deterministic)
    d_val = torch.empty(dnnz, dtype=torch.float64, device="cuda")
    d_col = torch.empty(dnnz, dtype=torch.int32, device="cuda")
    d_rp = torch.from_numpy(rp[s_row:e_row + 2]).cuda()
    sb.synth_fill_csr(d_rp.data_ptr(), s_row, e_row - s_row + 1, s_idx, e_idx + 1, n, wl["cols_mode"], wl["band"],
                      SEED, d_val.data_ptr(), d_col.data_ptr())
    torch.cuda.synchronize()
    del d_rp
    plan = sb.Plan.create_rank(sb.V1, m, n, nnz, d_val.data_ptr(), rp, d_col.data_ptr(), world, rank, local,
                               kernel=args.kernel, flags=sb.SRC_DEVICE_SHARD, keep=(d_val, d_col))
    y_ptr, first_row, rows = plan.y_ptr()
    sb.synth_fill_uniform(plan.x_ptr(), n, SEED + 7, 0.0, 1.0)
    sb.synth_fill_uniform(y_ptr, rows, SEED + 9, 0.0, 1.0)
    sb.device_synchronize()
    stream = torch.cuda.ExternalStream(plan.stream())
    slots = plan.edge_slots
    edge = torch.zeros(max(slots, 1), dtype=torch.float64, device="cuda")
    table = torch.zeros(world * max(slots, 1), dtype=torch.float64, device="cuda")
    exchange = "none"
    if world > 1 and slots:
        exchange = args.exchange
        if exchange == "symm":
            try:
                import torch.distributed._symmetric_memory as symm
                tw = world * slots
                sbuf = symm.empty(2 * tw + 2 * world, dtype=torch.float64, device=torch.device("cuda", local))
                sbuf.zero_()
                hdl = symm.rendezvous(sbuf, dist.group.WORLD)
                torch.cuda.synchronize()
                dist.barrier()
                plan.bind_peer_tables(list(hdl.buffer_ptrs), tw)
            except Exception as ex:          # no peer mapping available: fall back to NCCL
                if rank == 0:
                    print("symmetric memory unavailable (%s); using NCCL all-gather" % ex, file=sys.stderr)
                exchange = "nccl"
        if exchange == "nccl":
            plan.bind_edge_table(edge.data_ptr())

    def step():
        plan.execute_device(ALPHA, BETA)
        if exchange == "symm":
            plan.exchange_merge(ALPHA, BETA)
        elif exchange == "nccl":
            dist.all_gather_into_tensor(table, edge)
            plan.merge_gathered(table.data_ptr(), ALPHA, BETA)

    # ---- parity check of a row sample at FULL size, before timing (oracle = checker only)
    check = None
    if not args.no_check:
        import oracle
        y0 = np.empty(rows)
        sb.memcpy(y0, y_ptr, 8 * rows, 2)
        xh = np.empty(n)
        sb.memcpy(xh, plan.x_ptr(), 8 * n, 2)
        with torch.cuda.stream(stream):
            step()
        torch.cuda.synchronize()
        y1 = np.empty(rows)
        sb.memcpy(y1, y_ptr, 8 * rows, 2)
        rng = np.random.default_rng(rank)
        skip = 1 if parts["start_flag"][rank] else 0
        cand = np.arange(s_row + skip, e_row + 1)
        pick = np.unique(np.concatenate([cand[:3], cand[-3:], rng.choice(cand, size=min(1500, len(cand)), replace=False)]))
        worst = 0.0
        for r in pick:
            b, e = int(rp[r]), int(rp[r + 1])
            if world == 1 or (b >= s_idx and e - 1 <= e_idx):
                vv = np.empty(e - b)
                cc = np.empty(e - b, np.int32)
                sb.memcpy(vv, d_val.data_ptr() + 8 * (b - s_idx), 8 * (e - b), 2)
                sb.memcpy(cc, d_col.data_ptr() + 4 * (b - s_idx), 4 * (e - b), 2)
                lrp = np.array([0, e - b], np.int64)
                yi = np.array([y0[r - first_row]])
                want = oracle.csr_spmv(lrp, cc, vv, xh, ALPHA, BETA, yi)[0]
                bound = oracle.csr_spmv_bound(lrp, cc, vv, xh, ALPHA, BETA, yi)[0]
                rel = abs(y1[r - first_row] - want) / bound
                worst = max(worst, rel)
        check = {"rows_checked": int(len(pick)), "max_err_over_bound": worst, "tolerance": 1e-12, "ok": bool(worst <= 1e-12)}
        if not check["ok"]:
            raise SystemExit("parity check failed on rank %d: %r" % (rank, check))
        sb.memcpy(y_ptr, y0, 8 * rows, 1)

    # ---- timed region
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None      # started early: nvidia-smi needs a moment
    if sampler is not None and sampler.p is not None and world > 1:
        t_w = time.time()                                      # the other ranks wait at the barrier below
        while not sampler.lines and time.time() - t_w < 3.0:
            time.sleep(0.05)
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
    barrier()
    if sampler is not None and sampler.p is not None and world == 1:
        t_w = time.time()                                      # keep the GPU busy until the first sample arrives
        while not sampler.lines and time.time() - t_w < 3.0:
            with torch.cuda.stream(stream):
                step()
            torch.cuda.synchronize()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.time()
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for k in range(args.steps):
            step()
            ev[k + 1].record(stream)
    ev[-1].synchronize()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    total_ms = ev[0].elapsed_time(ev[-1])
    per = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)])
    tt = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item()) / args.steps
    gflops = 2.0 * nnz / (ms_step * 1e-3) / 1e9

    # ---- algorithmic bytes: x_touched = distinct columns a shard reads
    if wl["cols_mode"] == sb.COLS_PREFIX:
        x_touched = int(lens.max())
    elif wl["band"] and wl["cols_mode"] in (sb.COLS_BANDRUN, sb.COLS_BANDED):
        x_touched = min(n, (e_row - s_row + 1) + 2 * wl["band"])       # this rank's rows plus the band
    else:
        x_touched = n
    alg = torch.tensor([plan.alg_bytes(True, min(x_touched, n))], dtype=torch.float64, device="cuda")
    alg_rank0 = float(alg.item())
    if world > 1:
        dist.all_reduce(alg, op=dist.ReduceOp.SUM)
    alg_total = float(alg.item())

    # ---- the dominant kernel alone (roofline): the plan's largest row panel, launched by itself
    KNAMES = {1: "spmv_vec_kernel", 2: "spmv_tile_kernel + spmv_tile_fixup", 3: "spmv_tma_kernel + spmv_tile_fixup",
              4: "spmv_vecp_kernel", 5: "spmv_short_kernel", 6: "spmv_rowtile_kernel"}
    units = plan.units()
    dom = max(units, key=lambda u: u["nz1"] - u["nz0"]) if units else None
    dom_ms = None
    if dom is not None:
        nrep = max(3, min(args.steps, 50))
        with torch.cuda.stream(stream):
            for _ in range(3):
                plan.execute_unit(dom["index"], ALPHA, BETA)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(nrep):
                plan.execute_unit(dom["index"], ALPHA, BETA)
            e1.record(stream)
        e1.synchronize()
        dom_ms = e0.elapsed_time(e1) / nrep
        u_nnz, u_rows = dom["nz1"] - dom["nz0"], dom["row_hi"] - dom["row_lo"] + 1
        if wl["cols_mode"] == sb.COLS_PREFIX:
            u_x = min(int(lens[dom["row_lo"]:dom["row_hi"] + 1].max()), n)
        elif wl["band"] and wl["cols_mode"] in (sb.COLS_BANDRUN, sb.COLS_BANDED):
            u_x = min(n, u_rows + 2 * wl["band"])
        else:
            u_x = n
        dom_alg = 12.0 * u_nnz + 4.0 * (u_rows + 1) + 8.0 * u_x + 16.0 * u_rows
    barrier()

    # ---- end to end: host x, y (pinned) in, host y out, every step
    xh_p = torch.empty(n, dtype=torch.float64).pin_memory()
    yh_p = torch.empty(m, dtype=torch.float64).pin_memory()
    xh_p.uniform_(0, 1)
    yh_p.uniform_(0, 1)
    xh_np, yh_np = xh_p.numpy(), yh_p.numpy()

    def e2e_step():
        if world == 1:                     # the plan's host API: one call, host x and y in, host y out
            plan.execute(ALPHA, xh_np, BETA, yh_np)
            return
        plan.upload(xh_np, yh_np)
        with torch.cuda.stream(stream):
            step()
        plan.download(yh_np)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    xw = plan.x_window()                  # every rank uploads the window of x its shard references
    xb = torch.tensor([8.0 * (xw[1] - xw[0] + 1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(xb, op=dist.ReduceOp.SUM)
    h2d = int(xb.item()) + 8 * m         # + the y slices, which add up to m (+ shared rows)
    d2h = 8 * m

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        # dominant kernel = spmv_tile_kernel on rank 0's shard; its launch (+ the few-us fix-up) is the N=1 step
        k_ms = float(np.median(per)) if world == 1 else None
        step_ach = alg_rank0 / (float(np.mean(per)) * 1e-3) / 1e9
        ach = dom_alg / (dom_ms * 1e-3) / 1e9 if dom_ms else step_ach
        dom_name = KNAMES.get(dom["kind"], "?") if dom else "?"
        traffic = None            # dram bytes read+written per launch from the committed ncu capture, if any
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr.get("%s:n%d" % (args.workload, world), {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        out = {
            "metric": "double CSR SpMV GFLOP/s", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "m": m, "n": n, "nnz": nnz,
                       "partition": "v1 nnz-balanced x%d" % world, "kernel": args.kernel, "exchange": exchange, "alpha": ALPHA, "beta": BETA,
                       "l2": "inputs (%.1f GB per step) far exceed the 126 MB L2; no flush" % (alg_total / 1e9)},
            "hbm_gbs": alg_total / (ms_step * 1e-3) / 1e9,
            "hbm_frac_of_8000": alg_total / (ms_step * 1e-3) / 1e9 / (8000.0 * world),
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "%s on rank 0's largest row panel (rows %d..%d, %d nnz), timed alone" % (
                             dom_name, dom["row_lo"], dom["row_hi"], dom["nz1"] - dom["nz0"]) if dom else None,
                         "alg_bytes_per_launch": dom_alg if dom else None, "ms_per_launch_mean": dom_ms,
                         "share_of_step": (dom_ms / float(np.mean(per))) if dom_ms else None,
                         "whole_step": {"achieved": step_ach, "frac": step_ach / peak, "alg_bytes": alg_rank0,
                                        "ms_mean": float(np.mean(per)), "ms_median": k_ms,
                                        "panels": [{"kernel": KNAMES.get(u["kind"], "?"), "rows": u["row_hi"] - u["row_lo"] + 1,
                                                    "nnz": u["nz1"] - u["nz0"]} for u in units]}},
            "e2e": {"value": 2.0 * nnz / e2e_s / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                    "api": ("sblas_spmv_plan_execute on a resident plan (host x, y in; host y out; y slices of the row panels "
                            "move while other panels compute)") if world == 1 else
                           "sblas_spmv_plan_upload + execute_device + fused split-row exchange + download on a resident plan"},
            "gpu_launches": args.steps * (plan.launches + (2 if exchange == "symm" else 1 if exchange == "nccl" else 0)),
            "clocks": clocks, "parity_check": check,
        }
        if not args.no_cpu and world == 1:
            rps, cols, vals, xs, ys = host_sample(wl, args.cpu_sample)
            snnz = int(rps[-1])
            t_mt, cores = time_cpu(rps, cols, vals, xs, ys, 5, "oracle_csr_spmv_omp_balanced")
            t_st, _ = time_cpu(rps, cols, vals, xs, ys, 2, "oracle_csr_spmv")
            out["cpu_baseline"] = {"value": 2.0 * snnz / t_mt / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "port",
                                   "sample": "%d of %d rows (every block keeps its share), %d nnz, best of 5" % (len(rps) - 1, m, snnz),
                                   "single_thread_gflops": 2.0 * snnz / t_st / 1e9}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    plan.destroy()


if __name__ == "__main__":
    main()
