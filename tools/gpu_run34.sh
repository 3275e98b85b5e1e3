#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest.log 2>&1; grep -E "passed|failed|rror" gpurun_out/pytest.log | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-1200
ROCQ_PLAN_CACHE=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_n1_nocache.log 2>&1; tail -1 gpurun_out/bench_n1_nocache.log | cut -c1-700
