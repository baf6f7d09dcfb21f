#!/bin/bash
# A/B of kernel families on the diagnostic workloads. usage: tools/diag_kinds.sh "<workloads>" "<kind:ipt ...>" [steps]
WLS=${1:-"rows180"}
KINDS=${2:-"tile:4 tile:8 tile:16 vecp:4 vecp:8"}
STEPS=${3:-20}
mkdir -p gpurun_out
for wl in $WLS; do
  for ki in $KINDS; do
    kind=${ki%%:*}; ipt=${ki##*:}
    SBLAS_KIND=$kind SBLAS_IPT=$ipt python bench.py --workload $wl --steps $STEPS --warmup 3 --no-cpu --e2e-steps 1 \
      > gpurun_out/kind_${wl}_${kind}${ipt}.json 2> gpurun_out/kind_${wl}_${kind}${ipt}.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/kind_${wl}_${kind}${ipt}.json").read().strip().splitlines()[-1])
    print("${wl} ${kind} ipt ${ipt}: %.3f ms  %.0f GB/s alg  frac %.3f  check %s" % (d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["parity_check"]["ok"]))
except Exception as e:
    print("${wl} ${kind} ${ipt}: FAILED", e)
PY
  done
done
