#!/bin/bash
# 1 GPU, short: full GPU test-suite, smoke(), then the one-shot call beside the reference's with the stage timers on.
O=gpurun_out/r2g; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest.time; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log; tail -3 $O/pytest.time
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
SBLAS_TIMING=1 timeout 200 python - > $O/one_shot.json 2> $O/one_shot.err <<'PY'
import json, bench
print(json.dumps(bench.reference_gpu(1)))
PY
echo "one-shot rc=$?"; cat $O/one_shot.json; grep "sblas" $O/one_shot.err | tail -24
