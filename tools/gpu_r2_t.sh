#!/bin/bash
# Round 2, call T (2 GPUs): bench --gpus 2 with the final build (block tiles with chosen columns on the slices), its launch
# list from rank 0 (ROCQ_TRACE_LAUNCHES), and the single-process group on the two devices.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 400 $TR --master-port 29613 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/bench_n2_final.log 2>&1; tail -1 gpurun_out/bench_n2_final.log | cut -c1-1300
grep -o '"parity": {[^}]*}' gpurun_out/bench_n2_final.log | cut -c1-250; grep -o '"exchange": {[^}]*}' gpurun_out/bench_n2_final.log | cut -c1-250
( ROCQ_TRACE_LAUNCHES=1 timeout 300 $TR --master-port 29615 bench.py --gpus 2 --steps 1 --warmup 3 --no-parity ) > gpurun_out/trace_n2.log 2>&1
grep "^\[launch\] rank 0" gpurun_out/trace_n2.log | tail -45 > gpurun_out/launches_n2_rank0.log; awk '{print $4, $6, $8}' gpurun_out/launches_n2_rank0.log | tr '\n' ';'
timeout 300 python tools/group_bench.py --parity-qubits 0 --steps 2 > gpurun_out/group_bench_2_final.log 2>&1; cat gpurun_out/group_bench_2_final.log | cut -c1-800
