#!/usr/bin/env python
"""Digest of an `ncu --set full --import-source on` report for profiles/: the key raw metrics of the first kernel
in the report, its warp stall reasons and the CUDA source lines with the most stall samples.
usage: python tools/ncu_digest.py <report.ncu-rep> <title> > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read.sum.per_second",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_sectors_mem_global_op_tma_ld.sum", "l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum",
        "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, title = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("# %s -- ncu --set full --clock-control none, one launch" % title)
    print("kernel:", vals[hdr.index("Kernel Name")])
    seen = set()
    for i, h in enumerate(hdr):
        base = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[0].isupper() else h
        if base in KEYS and base not in seen:
            seen.add(base)
            print("%-82s %18s %s" % (base, vals[i], units[i]))
    stall = {}
    for i, h in enumerate(hdr):
        if "smsp__average_warps_issue_stalled_" in h and h.endswith("_per_issue_active.ratio"):
            name = h.split("smsp__average_warps_issue_stalled_")[1].replace("_per_issue_active.ratio", "")
            try:
                stall[name] = float(vals[i])
            except ValueError:
                pass
    tot = sum(stall.values()) or 1.0
    print("\nwarp stall reasons (warps stalled per issue-active cycle, share): " +
          ", ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:8]))
    src = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"])
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) < 3:
        return
    hdr = rows[0]

    def col(name):
        for i, h in enumerate(hdr):
            if h.strip() == name:
                return i
        return -1
    ci_src, ci_samp, ci_inst = col("Source"), col("# Samples"), col("Instructions Executed")
    if ci_src < 0:
        ci_src = 1
    if ci_samp < 0:
        return
    per = defaultdict(lambda: [0.0, 0.0])
    for r in rows[1:]:
        if len(r) <= max(ci_src, ci_samp):
            continue
        try:
            per[r[ci_src].strip()[:150]][0] += float(r[ci_samp] or 0)
            if ci_inst >= 0:
                per[r[ci_src].strip()[:150]][1] += float(r[ci_inst] or 0)
        except ValueError:
            pass
    ts = sum(v[0] for v in per.values()) or 1.0
    ti = sum(v[1] for v in per.values()) or 1.0
    print("\ntop source lines by stall samples  (%samples  %instructions  source)")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:14]:
        print("%5.1f%%  %5.1f%%  %s" % (100 * v[0] / ts, 100 * v[1] / ti, k))


if __name__ == "__main__":
    main()
