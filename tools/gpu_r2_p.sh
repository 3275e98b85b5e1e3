#!/bin/bash
# Round 2, call P (1 GPU): block tiles whose column bits are not the lowest seven (two-run blocks): parity + speed;
# full test suite; the distributed plan on 2 slices of one device (wall vs device, launch list).
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -6 gpurun_out/pytest.log
for qs in 10,11,12,13,14,15 11,12,13,27,28,29 8,9,26,27,28,29 24,25,26,27,28,29 4,5,6,27,28,29; do timeout 120 python tools/tc_time.py $qs 2>&1 | tail -1; done
ROCQ_TRACE_LAUNCHES=1 timeout 600 python tools/group_bench.py --ranks 2 --qubits 31 --parity-qubits 26 --steps 1 > gpurun_out/group_bench_1gpu.log 2>&1
grep -v "^\[launch\]" gpurun_out/group_bench_1gpu.log | cut -c1-700; grep "^\[launch\] rank 0" gpurun_out/group_bench_1gpu.log | tail -42 | awk '{print $4, $6, $8}' | tr '\n' ';'
