#!/bin/bash
# Round-end single-GPU measurements: GPU tests, bench lines for every workload, ncu launch list and
# full captures of the dominant kernels.  Outputs under gpurun_out/final/.
O=gpurun_out/final; mkdir -p $O
python bench.py --steps 200 --warmup 5 > $O/bench_g1m_n1.json 2> $O/bench_g1m_n1.err
for wl in g100000 big50m big50m_scatter circuit5m rail4284; do
  python bench.py --workload $wl --steps 60 --warmup 5 > $O/bench_${wl}_n1.json 2> $O/bench_${wl}_n1.err
done
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
for f in $O/bench_*_n1.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "%.3f ms %.0f GFLOP/s alg %.0f GB/s frac %.3f e2e %.0f GFLOP/s launches %s clocks %s" % (d["ms_per_step"], d["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"]))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
# one plain run per command line, then the ncu passes over the same command line
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-check --e2e-steps 1"
$CMD > $O/plain_g1m.log 2>&1 && {
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_g1m.csv $CMD > $O/ncu_launches_g1m.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:spmv_ -c 24 --csv --log-file $O/traffic_g1m.csv $CMD > $O/ncu_traffic_g1m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_tma_kernel -s 3 -c 1 -f -o $O/prof_g1m_tma $CMD > $O/ncu_g1m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_rowtile_kernel -s 3 -c 1 -f -o $O/prof_g1m_rowtile $CMD > $O/ncu_g1m_rt.log 2>&1
}
CMD2="python bench.py --workload big50m --steps 3 --warmup 3 --no-cpu --no-check --e2e-steps 1"
$CMD2 > $O/plain_big50m.log 2>&1 && {
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_big50m.csv $CMD2 > $O/ncu_launches_big50m.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:spmv_ -c 16 --csv --log-file $O/traffic_big50m.csv $CMD2 > $O/ncu_traffic_big50m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_rowtile_kernel -s 3 -c 1 -f -o $O/prof_big50m_rowtile $CMD2 > $O/ncu_big50m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_short_kernel -s 3 -c 1 -f -o $O/prof_big50m_short $CMD2 > $O/ncu_big50m_short.log 2>&1
}
ls $O | wc -l
