#!/bin/bash
mkdir -p gpurun_out
ROCQ_BENCH_QUBITS=28 ROCQ_BENCH_DEPTH=8 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain_fused.log 2>&1 && \
ROCQ_BENCH_QUBITS=28 ROCQ_BENCH_DEPTH=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_sweep -s 20 -c 3 -o gpurun_out/prof_phased python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_phased.log 2>&1
tail -2 gpurun_out/ncu_phased.log
