#!/bin/bash
# usage: tools/r2_multi.sh N [steps]   (under gpurun --gpus N): the GPU tests that drive the reference entry points
# IN-PROCESS over 1..N GPUs (SBLAS_EXPECT_GPUS=N: fewer visible GPUs is an error), then the bench at N (both arms).
N=$1; STEPS=${2:-100}
O=gpurun_out/r2m$N; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
( time SBLAS_EXPECT_GPUS=$N timeout 1200 python -m pytest tests -m gpu -q -x --durations=6 \
    -k "qh768 or generator or versions_and_kernels or chained or byte_balanced or row_panels or row_tile or row_split or reference_code_live or spmm or sptrans or cli or spanning" \
    > $O/pytest_gpu_n$N.log 2>&1 ) 2> $O/pytest.time; echo "pytest rc=$?"; tail -12 $O/pytest_gpu_n$N.log; tail -3 $O/pytest.time
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps $STEPS --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err ) 2> $O/bench_n$N.time; echo "bench rc=$?"; tail -5 $O/bench_n$N.err; tail -3 $O/bench_n$N.time
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $N --steps 5 --warmup 2 > $O/bench_ref_n$N.json 2> $O/bench_ref_n$N.err ); echo "ref rc=$?"; cut -c1-400 $O/bench_ref_n$N.json | tail -1
python - $O/bench_n$N.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.0f ms %.4f e2e %.0f (%.4f ms) parity %s clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity_check"], d["clocks"]))
for c in d.get("configs", []):
    print(c.get("name"), c.get("error") or ("%.4f ms %.0f GF %.0f GB/s frac8000 %.3f parity %s samples %s e2e %.0f" % (c["ms_per_step"], c["gflops"], c["hbm_gbs"], c["hbm_frac_of_8000"], c["parity_ok"], c["clocks"]["samples"], c["e2e"]["value"])))
a=d.get("inprocess_api") or {}
print("inprocess ok", a.get("ok"), a.get("error"), [(c["matrix"][:8], c["entry"], round(c["ms_whole_call"],1), c["ok"]) for c in a.get("calls", [])], a.get("chain"))
print("refgpu", d.get("reference_gpu"))
for k in ("rank_chain", "spmm", "sptrans"):
    print(k, d.get(k))
print("wall", d.get("job_wall_s"))
PY
