#!/bin/bash
# Round 2, call W (8 GPUs): the N = 8 bench line (36 qubits) with the final build -- parity self-check, exchange figures --
# and rank 0's launch list of one step.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29613 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/bench_n8_final.log 2>&1; tail -1 gpurun_out/bench_n8_final.log | cut -c1-1300
grep -o '"parity": {[^}]*}' gpurun_out/bench_n8_final.log | cut -c1-250; grep -o '"exchange": {[^}]*}' gpurun_out/bench_n8_final.log | cut -c1-250
( ROCQ_TRACE_LAUNCHES=1 timeout 200 $TR --master-port 29615 bench.py --gpus 8 --steps 1 --warmup 3 --no-parity ) > gpurun_out/trace_n8.log 2>&1
grep "^\[launch\] rank 0" gpurun_out/trace_n8.log | tail -45 > gpurun_out/launches_n8_rank0_final.log; awk '{print $4, $6, $8}' gpurun_out/launches_n8_rank0_final.log | tail -42 | tr '\n' ';'
