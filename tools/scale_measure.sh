#!/bin/bash
# usage: tools/scale_measure.sh N "<workloads>" [steps]   (under gpurun --gpus N)
N=$1; WLS=${2:-"g1m big50m"}; STEPS=${3:-100}
O=gpurun_out/final; mkdir -p $O
for wl in $WLS; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps $STEPS --warmup 5 --workload $wl > $O/bench_${wl}_n$N.json 2> $O/bench_${wl}_n$N.err
  python - "$O/bench_${wl}_n$N.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "%.4f ms %.0f GFLOP/s total alg %.0f GB/s (%.3f of 8000/GPU) e2e %.0f launches %s check %s" % (d["ms_per_step"], d["value"], d["hbm_gbs"], d["hbm_frac_of_8000"], d["e2e"]["value"], d["gpu_launches"], d["parity_check"]["ok"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
