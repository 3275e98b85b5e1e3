#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.log 2>&1 || { tail -20 gpurun_out/bench_final.log; exit 1; }
tail -1 gpurun_out/bench_final.log | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench2.log 2>&1
tail -2 gpurun_out/ncu_bench2.log | cut -c1-300
