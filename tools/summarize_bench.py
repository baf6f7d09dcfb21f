#!/usr/bin/env python
"""Markdown table rows from bench.py JSON lines (profiles/r02/bench_g1m_n<N>.json): one row per workload and N."""
import json
import sys


def rows(path):
    d = json.loads(open(path).read().strip().splitlines()[-1])
    n = d["n_gpus"]
    peak = d["roofline"]["peak"]
    out = [("g1m (config 2b, headline)", n, d["ms_per_step"], d["value"], d["hbm_gbs"], d["hbm_gbs"] / (peak * n),
            d["hbm_frac_of_8000"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity_check"]["ok"], d["clocks"])]
    for c in d.get("configs", []):
        if "error" in c:
            out.append((c["name"], n, None, None, None, None, None, None, None, c["error"], None))
            continue
        out.append((c["name"], n, c["ms_per_step"], c["gflops"], c["hbm_gbs"], c["whole_step_frac"], c["hbm_frac_of_8000"],
                    c["e2e"]["value"], c["e2e"]["ms_per_step"], c["parity_ok"], c["clocks"]))
    return out


def main():
    allr = []
    for p in sys.argv[1:]:
        allr += rows(p)
    base = {}
    for r in allr:
        if r[1] == 1 and r[2]:
            base[r[0].split("@")[0]] = r[2]
    print("| workload | N | ms/step | GFLOP/s | alg. GB/s (all GPUs) | of measured copy peak (per GPU) | of 8000 (per GPU) | speed-up vs N=1 | e2e GFLOP/s (ms) | parity | SM MHz (reasons, samples) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for r in sorted(allr, key=lambda r: (r[0].split("@")[0].replace("g1m", "a"), r[0], r[1])):
        if r[2] is None:
            print("| %s | %d | failed: %s |" % (r[0], r[1], r[9]))
            continue
        b = base.get(r[0].split("@")[0])
        ck = r[10] or {}
        print("| %s | %d | %.4f | %.0f | %.0f | %.2f | %.2f | %s | %.0f (%.3f) | %s | %s (%s, %s) |" % (
            r[0], r[1], r[2], r[3], r[4], r[5], r[6], ("%.2f×" % (b / r[2])) if b and r[1] > 1 else "—", r[7], r[8],
            "ok" if r[9] else "FAIL", ck.get("sm_mhz"), ",".join(ck.get("reasons") or []) or "none", ck.get("samples")))


if __name__ == "__main__":
    main()
