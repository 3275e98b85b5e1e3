#!/bin/bash
mkdir -p gpurun_out
free -g | head -2
( time timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1; tail -5 gpurun_out/bench_ref.log | cut -c1-900
timeout 600 python tools/config_bench.py > gpurun_out/config_bench2.log 2>&1; cat gpurun_out/config_bench2.log | cut -c1-330
