#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; tail -4 gpurun_out/pytest.log
ROCQ_HOST_PROFILE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_tcprof.log 2>&1
grep "host profile" gpurun_out/bench_tcprof.log | tail -4
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_tcprof.log") if x.startswith("{")]
if l:
    j=json.loads(l[-1]); print("value", j["value"], "ms", j["ms_per_step"], "dev_ms", j["device_ms_per_step"], "launches", j["gpu_launches"], "e2e", j["e2e"]["value"]); print(json.dumps(j["roofline"])[:1500])
else: print(open("gpurun_out/bench_tcprof.log").read()[-2000:])
PY
nproc; lscpu | grep -E "Model name" | head -3
