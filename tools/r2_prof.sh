#!/bin/bash
# Round 2 profiling pass (1 GPU): launch lists, DRAM traffic and `ncu --set full` captures of the dominant kernels.
# Every ncu pass runs only after the same command line has exited 0 without ncu.  Outputs under gpurun_out/r2p/.
O=gpurun_out/r2p; mkdir -p $O
B="--steps 3 --warmup 3 --no-cpu --no-check --no-extra --e2e-steps 1"
for wl in g1m big50m circuit5m rail4284; do
  CMD="python bench.py --workload $wl $B"
  $CMD > $O/plain_$wl.log 2>&1 || { echo "plain $wl failed"; continue; }
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$wl.csv $CMD > /dev/null 2>&1
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:spmv_ -c 40 --csv --log-file $O/traffic_$wl.csv $CMD > /dev/null 2>&1
done
cap() {   # name  kernel-regex  skip  command...
  local name=$1 k=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o $O/prof_$name "$@" > $O/ncu_$name.log 2>&1
  python tools/ncu_digest.py $O/prof_$name.ncu-rep "$name" > $O/digest_$name.txt 2>/dev/null
}
cap g1m_tma spmv_tma_kernel 3 python bench.py --workload g1m $B
cap g1m_rowtile spmv_rowtile_kernel 3 python bench.py --workload g1m $B
cap big50m_rowtile spmv_rowtile_kernel 3 python bench.py --workload big50m $B
cap big50m_short spmv_short_kernel 3 python bench.py --workload big50m $B
cap circuit5m_tma spmv_tma_kernel 3 python bench.py --workload circuit5m $B
cap rail4284_tma spmv_tma_kernel 3 python bench.py --workload rail4284 $B
python tools/spmm_time.py inproc 128 2 > $O/plain_spmm.log 2>&1 && cap spmm_rows spmm_rows_kernel 2 python tools/spmm_time.py inproc 128 2
ls -la $O | awk '{print $5, $9}' | tail -40
