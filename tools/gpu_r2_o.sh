#!/bin/bash
# Round 2, call O (8 GPUs): hardware parity of the distributed engine at 8 ranks (dist_check: eager, deferred, whole-circuit,
# sampling, measurement, index-bit swaps incl. global<->global, blocks on slices), the bench line with its parity self-check
# (36 qubits), and rank 0's launch list of one step (ROCQ_TRACE_LAUNCHES: CUDA events around every launch).
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( timeout 240 $TR --master-port 29611 tests/dist_check.py ) > gpurun_out/dist_check8.log 2>&1; tail -5 gpurun_out/dist_check8.log
( timeout 360 $TR --master-port 29613 bench.py --gpus 8 --steps 2 --warmup 3 ) > gpurun_out/bench_n8.log 2>&1; tail -1 gpurun_out/bench_n8.log | cut -c1-1400
grep -o '"parity": {[^}]*}' gpurun_out/bench_n8.log; grep -o '"exchange": {[^}]*}' gpurun_out/bench_n8.log | cut -c1-330
( ROCQ_TRACE_LAUNCHES=1 timeout 240 $TR --master-port 29615 bench.py --gpus 8 --steps 1 --warmup 3 --no-parity ) > gpurun_out/trace_n8.log 2>&1
grep "^\[launch\] rank 0" gpurun_out/trace_n8.log | tail -60 > gpurun_out/launches_n8_rank0.log; tail -3 gpurun_out/launches_n8_rank0.log
