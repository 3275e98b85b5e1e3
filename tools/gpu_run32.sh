#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest.log 2>&1; grep -E "passed|failed|error" gpurun_out/pytest.log | tail -3
timeout 300 python tools/sweep_bench.py --n 30 --prec c64 --reps 10 > gpurun_out/sweep_bench_c64.log 2>&1; cut -c1-100 gpurun_out/sweep_bench_c64.log
timeout 300 python tools/sweep_bench.py --n 29 --prec c128 --reps 10 > gpurun_out/sweep_bench_c128.log 2>&1; cut -c1-100 gpurun_out/sweep_bench_c128.log | head -8
timeout 600 python tools/config_bench.py --reps 2 > gpurun_out/config_bench.log 2>&1; cut -c1-330 gpurun_out/config_bench.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_qft33_c128.csv python tools/config_bench.py --only c3 --reps 0 > gpurun_out/ncu_qft33.log 2>&1
grep -v "^==" gpurun_out/launches_qft33_c128.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-160 | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_sweep -c 3 -o gpurun_out/qft28_c128_final -f python tools/config_bench.py --only c3 --reps 0 --c3-qubits 28 > gpurun_out/ncu_qft28.log 2>&1
