#!/bin/bash
mkdir -p gpurun_out
ROCQ_HOST_PROFILE=1 ROCQ_TC=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_tcprof.log 2>&1
grep "host profile" gpurun_out/bench_tcprof.log | tail -4
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_tcprof.log") if x.startswith("{")]
j=json.loads(l[-1]); print("value", j["value"], "ms", j["ms_per_step"], "dev_ms", j["device_ms_per_step"], "launches", j["gpu_launches"], "e2e", j["e2e"]["value"])
PY
