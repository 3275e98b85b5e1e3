#!/bin/bash
# Round 2, call N (1 GPU): the failed test again, then ncu --set full of the QFT chain-phase sweeps (28 qubits c128) and
# launch list of QFT-33.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "eager_wide or apply_matrix" ) > gpurun_out/pytest_wide.log 2>&1; tail -3 gpurun_out/pytest_wide.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tile_sweep -c 2 -f -o gpurun_out/r02_qft28_c128_chain \
    python tools/config_bench.py --only c3 --c3-qubits 28 --reps 1 > gpurun_out/ncu_qft28.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_qft33_c128.csv \
    python tools/config_bench.py --only c3 --reps 1 > gpurun_out/ncu_qft33.log 2>&1
grep tile_sweep gpurun_out/launches_qft33_c128.csv | tail -5 | awk -F'","' '{print $5, $NF}' | cut -c1-200
