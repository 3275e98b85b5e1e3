#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "exit $?" >> gpurun_out/tc_check.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
ROCQ_TC=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_tc.log 2>&1
ROCQ_TC=1 ROCQ_TC_MIN_COST=36 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_tc36.log 2>&1
ROCQ_TC=1 ROCQ_TC_MIN_COST=90 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_tc90.log 2>&1
tail -12 gpurun_out/tc_check.log; tail -3 gpurun_out/pytest.log; for f in gpurun_out/bench_tc*.log; do grep -o '"value": [0-9.]*' $f | head -1; done
