#!/bin/bash
mkdir -p gpurun_out
# per-launch durations of the QFT sweeps (c128, 30 qubits; c64, 30 qubits)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_qft30_c128.csv python tools/config_bench.py --only c3 --reps 0 --c3-qubits 30 > gpurun_out/ncu_qft30.log 2>&1
grep -v "^==" gpurun_out/launches_qft30_c128.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-160 | tail -20
# full capture of the QFT-28 c128 sweeps
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_sweep -c 7 -o gpurun_out/qft28_c128 -f python tools/config_bench.py --only c3 --reps 0 --c3-qubits 28 > gpurun_out/ncu_qft28.log 2>&1
ls -la gpurun_out/*.ncu-rep
