#!/bin/bash
# 2 GPUs, short: the SpMV GPU tests (in-process plans over 1 and 2 GPUs), then the one-shot call beside the reference's.
O=gpurun_out/r2g2; mkdir -p $O
( time timeout 600 python -m pytest tests/test_spmv_gpu.py tests/test_reference_y.py -m gpu -q -x > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest.time; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log; tail -3 $O/pytest.time
SBLAS_TIMING=1 timeout 300 python - > $O/one_shot.json 2> $O/one_shot.err <<'PY'
import json, bench
print(json.dumps(bench.reference_gpu(2)))
PY
echo "one-shot rc=$?"; cat $O/one_shot.json; grep "sblas" $O/one_shot.err | tail -12
