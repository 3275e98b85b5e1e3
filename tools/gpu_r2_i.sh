#!/bin/bash
# Round 2, call I (1 GPU): QFT butterfly -- parity tests that run QFTs, then configs[2] timing + launch list.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "qft or ladders or c3 or c2_and or smoke" ) > gpurun_out/pytest_qft.log 2>&1; tail -4 gpurun_out/pytest_qft.log
timeout 600 python tools/config_bench.py --only c3 --reps 3 > gpurun_out/config_bench_c3.log 2>&1; cut -c1-420 gpurun_out/config_bench_c3.log
timeout 600 python tools/config_bench.py --only c3 --reps 3 --c3-qubits 30 2>&1 | cut -c1-420
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_qft33_c128.csv \
    python tools/config_bench.py --only c3 --reps 1 > gpurun_out/ncu_qft33.log 2>&1
grep tile_sweep gpurun_out/launches_qft33_c128.csv | tail -5 | awk -F'","' '{print $5, $NF}' | cut -c1-200
