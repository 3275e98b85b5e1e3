#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest.log 2>&1; tail -6 gpurun_out/pytest.log
timeout 300 python tools/config_bench.py --only c1,c3 --reps 2 > gpurun_out/config_bench_c3.log 2>&1; cut -c1-400 gpurun_out/config_bench_c3.log
ROCQ_MERGE_DIAG=0 timeout 300 python tools/config_bench.py --only c3 --reps 1 --c3-qubits 30 > gpurun_out/config_bench_c3_30_nomerge.log 2>&1; cut -c1-400 gpurun_out/config_bench_c3_30_nomerge.log
timeout 300 python tools/config_bench.py --only c3 --reps 1 --c3-qubits 30 > gpurun_out/config_bench_c3_30.log 2>&1; cut -c1-400 gpurun_out/config_bench_c3_30.log
