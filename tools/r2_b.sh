#!/bin/bash
# Round 2, GPU call B (1 GPU): reference vectors, GPU tests, the new bench line (both arms).
O=gpurun_out/r2b; mkdir -p $O
python -u tests/golden/make_golden_y.py $O/ref_y.npz > $O/golden_y.log 2>&1; echo "golden rc=$?"; tail -3 $O/golden_y.log
cp $O/ref_y.npz tests/golden/ref_y.npz 2>/dev/null
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
( time timeout 900 python bench.py --steps 100 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err ) 2> $O/bench_n1.time; echo "bench rc=$?"; tail -3 $O/bench_n1.err; cat $O/bench_n1.time
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err ) 2> $O/bench_ref.time; echo "ref rc=$?"; cat $O/bench_ref.json | cut -c1-600; cat $O/bench_ref.time
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b/bench_n1.json").read().strip().splitlines()[-1])
print("value %.0f ms %.4f e2e %.0f frac %.3f read_peak %s cpu %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("read_peak"), d.get("cpu_baseline")))
for c in d.get("configs", []):
    print(c.get("name"), c.get("error") or ("%.4f ms %.0f GF %.0f GB/s frac8000 %.3f parity %s clocks %s e2e %.0f" % (c["ms_per_step"], c["gflops"], c["hbm_gbs"], c["hbm_frac_of_8000"], c["parity_ok"], c["clocks"], c["e2e"]["value"])))
print("refgpu", d.get("reference_gpu"))
print("wall", d.get("job_wall_s"))
PY
