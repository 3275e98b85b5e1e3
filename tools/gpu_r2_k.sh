#!/bin/bash
# Round 2, call K (1 GPU): ncu --set full of the QFT window sweeps (28 qubits complex128, first two sweeps) and of the block
# sweep (28 qubits), with source correlation; timing first (never under the profiler).
mkdir -p gpurun_out
timeout 600 python tools/config_bench.py --only c3 --reps 2 2>&1 | cut -c1-330
timeout 600 python tools/config_bench.py --only c3 --reps 2 --c3-qubits 28 2>&1 | cut -c1-330
ncu --set full --import-source on --clock-control none -k regex:tile_sweep -c 2 -f -o gpurun_out/r02_qft28_c128 \
    python tools/config_bench.py --only c3 --c3-qubits 28 --reps 1 > gpurun_out/ncu_qft28.log 2>&1
ls -la gpurun_out/*.ncu-rep
