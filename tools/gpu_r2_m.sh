#!/bin/bash
# Round 2, call M (1 GPU): everything since call J on one device -- full GPU tests (butterfly chains, eager 5/6-qubit
# matrices through the block sweep, NVTX build, batch binding), QFT timing, and where the host time of a distributed step
# goes (2 slices on one device, host profile on).
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -6 gpurun_out/pytest.log
timeout 600 python tools/config_bench.py --only c3 --reps 3 2>&1 | cut -c1-330
timeout 600 python tools/config_bench.py --only c3 --reps 2 --c3-qubits 30 2>&1 | cut -c1-330
ROCQ_HOST_PROFILE=1 timeout 600 python tools/group_bench.py --ranks 2 --qubits 31 --parity-qubits 0 --steps 2 > gpurun_out/group_bench_1gpu.log 2>&1; tail -25 gpurun_out/group_bench_1gpu.log | cut -c1-600
