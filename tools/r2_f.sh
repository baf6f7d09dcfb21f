#!/bin/bash
# 1 GPU: full GPU test-suite (per-test timeout from pytest.ini) + the bench line.
O=gpurun_out/r2f; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest.time; echo "pytest rc=$?"; tail -14 $O/pytest_gpu.log; cat $O/pytest.time | tail -3
( time timeout 900 python bench.py --steps 100 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err ) 2> $O/bench_n1.time; echo "bench rc=$?"; tail -3 $O/bench_n1.err; tail -3 $O/bench_n1.time
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f/bench_n1.json").read().strip().splitlines()[-1])
print("value %.0f ms %.4f e2e %.0f frac %.3f read_peak %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("read_peak")))
for c in d.get("configs", []):
    print(c.get("name"), c.get("error") or ("%.4f ms %.0f GF %.0f GB/s frac8000 %.3f parity %s e2e %.0f dom %s" % (c["ms_per_step"], c["gflops"], c["hbm_gbs"], c["hbm_frac_of_8000"], c["parity_ok"], c["e2e"]["value"], c.get("dominant"))))
for k in ("spmm", "sptrans"):
    print(k, d.get(k))
print("refgpu", d.get("reference_gpu"))
print("wall", d.get("job_wall_s"))
PY
