#!/usr/bin/env python
"""profiles/traffic.json from the ncu DRAM-traffic passes (tools/r2_prof.sh: dram__bytes_read.sum, dram__bytes_write.sum,
gpu__time_duration.sum of every spmv_* launch of a short bench run): per kernel the mean per launch, the dominant
kernel's bytes per launch (what bench.py's roofline.traffic quotes) and the sum over one product's launches.
usage: python tools/traffic_json.py <dir with traffic_<workload>.csv> > profiles/traffic.json"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

ALG = {"g1m": (1_000_000, 1_000_000, 1_212_500_000, 9000), "big50m": (50_000_000, 50_000_000, 1_212_500_000, 50_000_000),
       "circuit5m": (5_558_326, 5_558_326, 59_524_291, 5_558_326), "rail4284": (4284, 1_092_610, 11_279_909, 1_092_610)}


def main():
    d = sys.argv[1]
    out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --metrics pass over every spmv_* launch of a short "
                       "bench run (profiles/r02/traffic_<workload>.csv, tools/r2_prof.sh); dram_bytes_per_launch is the dominant "
                       "kernel's (what bench.py's roofline.traffic quotes), step_total the sum over one product's launches, "
                       "step_alg_bytes = 12*nnz + 4*(m+1) + 8*x_touched + 16*m"}
    for wl, (m, n, nnz, xt) in ALG.items():
        p = os.path.join(d, "traffic_%s.csv" % wl)
        if not os.path.exists(p):
            continue
        rows = [r for r in csv.reader(open(p)) if len(r) >= 15 and r[0].isdigit()]
        byid = defaultdict(dict)
        for r in rows:
            name = re.sub(r"\(.*", "", r[4]).replace("<unnamed>::", "").replace("void ", "").strip()
            byid[(r[0], name)][r[12]] = float(r[14])
        # whole-panel launches only: the end-to-end leg re-launches the row-aligned panels in pieces (shorter launches)
        longest = defaultdict(float)
        for (_, name), mm in byid.items():
            longest[name] = max(longest[name], mm.get("gpu__time_duration.sum", 0.0))
        per = defaultdict(lambda: defaultdict(list))
        for (_, name), mm in byid.items():
            if mm.get("gpu__time_duration.sum", 0.0) >= 0.8 * longest[name]:
                for k, v in mm.items():
                    per[name][k].append(v)
        kern = {}
        for name, mm in per.items():
            kern[name] = {"dram_read": sum(mm["dram__bytes_read.sum"]) / max(len(mm["dram__bytes_read.sum"]), 1),
                          "dram_write": sum(mm["dram__bytes_write.sum"]) / max(len(mm["dram__bytes_write.sum"]), 1),
                          "us": sum(mm["gpu__time_duration.sum"]) / max(len(mm["gpu__time_duration.sum"]), 1) / 1e3,
                          "launches_seen": len(mm["gpu__time_duration.sum"])}
        dom = max(kern, key=lambda k: kern[k]["us"])
        out["%s:n1" % wl] = {"kernel": dom, "dram_bytes_per_launch": kern[dom]["dram_read"] + kern[dom]["dram_write"],
                             "step_total": sum(k["dram_read"] + k["dram_write"] for k in kern.values()),
                             "step_alg_bytes": 12.0 * nnz + 4.0 * (m + 1) + 8.0 * xt + 16.0 * m, "kernels": kern}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
