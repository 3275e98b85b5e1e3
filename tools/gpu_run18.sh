#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core or block6" > gpurun_out/pytest_tc.log 2>&1; tail -5 gpurun_out/pytest_tc.log
for MC in 30 50 70; do
  ROCQ_TC=1 ROCQ_TC_MIN_COST=$MC timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_tc$MC.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_tc$MC.log") if x.startswith("{")]
if l:
    j=json.loads(l[-1]); print("min_cost $MC value", j["value"], "ms", j["ms_per_step"], "dev_ms", j["device_ms_per_step"], "launches", j["gpu_launches"], "e2e", j["e2e"]["value"])
else:
    print(open("gpurun_out/bench_tc$MC.log").read()[-1500:])
PY
done
