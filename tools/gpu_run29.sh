#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest.log 2>&1; tail -4 gpurun_out/pytest.log
timeout 600 python tools/config_bench.py --only c1,c3,c5 --reps 2 > gpurun_out/config_bench.log 2>&1; cut -c1-330 gpurun_out/config_bench.log | head -3
timeout 300 python tools/config_bench.py --only c3 --reps 1 --c3-qubits 30 > gpurun_out/config_bench_c3_30.log 2>&1; cut -c1-330 gpurun_out/config_bench_c3_30.log
ROCQ_TC=0 timeout 300 python tools/config_bench.py --only c2 --reps 1 > gpurun_out/config_bench_c2_notc.log 2>&1; cut -c1-330 gpurun_out/config_bench_c2_notc.log
timeout 300 python tools/sweep_bench.py --n 30 --prec c64 --reps 5 > gpurun_out/sweep_c64.log 2>&1; cut -c1-120 gpurun_out/sweep_c64.log | tail -28
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_qft30_c128_d.csv python tools/config_bench.py --only c3 --reps 0 --c3-qubits 30 > gpurun_out/ncu_qft30.log 2>&1
grep -v "^==" gpurun_out/launches_qft30_c128_d.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-160 | tail -9
