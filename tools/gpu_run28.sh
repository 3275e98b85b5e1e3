#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest.log 2>&1; tail -6 gpurun_out/pytest.log
timeout 600 python tools/config_bench.py --only c1,c3,c5 --reps 2 > gpurun_out/config_bench.log 2>&1; cut -c1-330 gpurun_out/config_bench.log
timeout 300 python tools/config_bench.py --only c3 --reps 1 --c3-qubits 30 > gpurun_out/config_bench_c3_30.log 2>&1; cut -c1-330 gpurun_out/config_bench_c3_30.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_qft30_c128_c.csv python tools/config_bench.py --only c3 --reps 0 --c3-qubits 30 > gpurun_out/ncu_qft30.log 2>&1
grep -v "^==" gpurun_out/launches_qft30_c128_c.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-160 | tail -9
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_sweep -c 2 -o gpurun_out/qft28_c128_c -f python tools/config_bench.py --only c3 --reps 0 --c3-qubits 28 > gpurun_out/ncu_qft28.log 2>&1
