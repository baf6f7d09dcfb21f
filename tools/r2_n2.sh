#!/bin/bash
# final 2-GPU pass: smoke(), the bench at N=2 (both arms), then two ncu captures on GPU 0
O=gpurun_out/r2m2; mkdir -p $O
python __graft_entry__.py smoke 2>&1 | tail -2
N=2
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err ) 2> $O/bench_n$N.time; echo "bench rc=$?"; tail -3 $O/bench_n$N.time
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 5 --warmup 2 > $O/bench_ref_n$N.json 2> $O/bench_ref_n$N.err; cut -c1-200 $O/bench_ref_n$N.json
python tools/summarize_bench.py $O/bench_n$N.json
B="--steps 3 --warmup 3 --no-cpu --no-check --no-extra --e2e-steps 1"
P=gpurun_out/r2p; mkdir -p $P
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_rowtile_kernel -s 3 -c 1 -f -o $P/prof_g1m_rowtile_lanegroups python bench.py --workload g1m $B > $P/ncu_g1m_rowtile_lanegroups.log 2>&1
python tools/ncu_digest.py $P/prof_g1m_rowtile_lanegroups.ncu-rep "g1m_rowtile (lane groups)" > $P/digest_g1m_rowtile_lanegroups.txt 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmm_segment_kernel -s 2 -c 1 -f -o $P/prof_spmm_segment python tools/spmm_time.py inproc 128 2 > $P/ncu_spmm_segment.log 2>&1
python tools/ncu_digest.py $P/prof_spmm_segment.ncu-rep "spmm_segment_kernel (inproc, n=128)" > $P/digest_spmm_segment.txt 2>/dev/null
grep -E "gpu__time|dram__bytes_read.sum  |issue_active|inst_executed.sum" $P/digest_g1m_rowtile_lanegroups.txt $P/digest_spmm_segment.txt | cut -c1-160
