#!/bin/bash
# A/B of the TMA kernel's per-tile reduction paths on the diagnostic workloads (one row length each).
# usage: tools/diag_modes.sh "<workloads>" "<modes>" [steps]   -> gpurun_out/diag_<wl>_m<mode>.json
WLS=${1:-"rows180 rows100 rows2"}
MODES=${2:-"0 1 3"}
STEPS=${3:-30}
mkdir -p gpurun_out
for wl in $WLS; do
  for mode in $MODES; do
    SBLAS_TMA_MODE=$mode python bench.py --workload $wl --steps $STEPS --warmup 3 --no-cpu --e2e-steps 1 \
      > gpurun_out/diag_${wl}_m${mode}.json 2> gpurun_out/diag_${wl}_m${mode}.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/diag_${wl}_m${mode}.json").read().strip().splitlines()[-1])
    print("${wl} mode ${mode}: %.3f ms  %.0f GB/s alg  frac %.3f  check %s" % (d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["parity_check"]["ok"]))
except Exception as e:
    print("${wl} mode ${mode}: FAILED", e)
PY
  done
done
