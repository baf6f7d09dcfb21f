#!/bin/bash
# Round 2, first GPU call (1 GPU): reference vectors, GPU tests, row-split A/B, ncu captures of the
# scattered-gather workloads.  Outputs under gpurun_out/r2a/.
O=gpurun_out/r2a; mkdir -p $O
python tests/golden/make_golden_y.py $O/ref_y.npz > $O/golden_y.log 2>&1; echo "golden rc=$?"
cp $O/ref_y.npz tests/golden/ref_y.npz 2>/dev/null
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
for wl in g100000 rows1000 circuit5m rail4284; do
  for med in 1 3; do
    SBLAS_MEDIUM=$med timeout 300 python bench.py --workload $wl --steps 40 --warmup 5 --no-cpu --e2e-steps 2 > $O/bench_${wl}_med$med.json 2> $O/bench_${wl}_med$med.err
    python - "$O/bench_${wl}_med$med.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "%.4f ms %.0f GFLOP/s alg %.0f GB/s panels %s" % (d["ms_per_step"], d["value"], d["hbm_gbs"], [(p["kernel"][:14],p["rows"],p["nnz"]) for p in d["roofline"]["whole_step"]["panels"]][:6]))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
  done
done
for wl in circuit5m rail4284; do
  CMD="python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-check --e2e-steps 1"
  $CMD > $O/plain_$wl.log 2>&1 && {
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_$wl.csv $CMD > $O/ncu_launches_$wl.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_tma_kernel -s 3 -c 1 -f -o $O/prof_${wl}_tma $CMD > $O/ncu_$wl.log 2>&1
  }
done
./tools/bin/hbm_probe > $O/hbm_probe.log 2>&1; tail -5 $O/hbm_probe.log
ls -la $O | head -40
