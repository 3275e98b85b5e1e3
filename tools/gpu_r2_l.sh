#!/bin/bash
# Round 2, call L (2 GPUs): bench --gpus 2 with its parity self-check; the single-process group on the same circuit
# (34 qubits over 2 devices) and its ncu launch list (one process, so ncu can follow it).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 900 $TR --master-port 29613 bench.py --gpus 2 --steps 2 --warmup 3 ) > gpurun_out/bench_n2.log 2>&1; tail -1 gpurun_out/bench_n2.log | cut -c1-1500
grep -o '"parity": {[^}]*}' gpurun_out/bench_n2.log; grep -o '"exchange": {[^}]*}' gpurun_out/bench_n2.log | cut -c1-300
timeout 900 python tools/group_bench.py > gpurun_out/group_bench_2.log 2>&1; cat gpurun_out/group_bench_2.log | cut -c1-800
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_group2.csv \
    python tools/group_bench.py --steps 1 --parity-qubits 0 > gpurun_out/ncu_group2.log 2>&1
python tools/launch_summary.py gpurun_out/launches_group2.csv > gpurun_out/launches_group2_summary.md 2>&1; head -16 gpurun_out/launches_group2_summary.md
