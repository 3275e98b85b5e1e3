#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_sweep -c 3 -o gpurun_out/qft28_c128_b -f python tools/config_bench.py --only c3 --reps 0 --c3-qubits 28 > gpurun_out/ncu_qft28.log 2>&1
ls -la gpurun_out/*.ncu-rep
