#!/bin/bash
# Round 2, call AA (2 GPUs): the decoupled block sweep and the handle-owned memory pool on distributed slices -- 2-GPU tests
# (group + torchrun), the N = 2 bench line with its parity object, rank 0's launch list.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout -s KILL 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_dist.py -m gpu -x -q ) > gpurun_out/pytest_2gpu.log 2>&1; tail -3 gpurun_out/pytest_2gpu.log
( timeout -s KILL 300 $TR --master-port 29611 tests/dist_check.py ) > gpurun_out/dist_check2.log 2>&1; tail -3 gpurun_out/dist_check2.log | cut -c1-300
( timeout -s KILL 400 $TR --master-port 29613 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/bench_n2_final.log 2>&1; tail -1 gpurun_out/bench_n2_final.log | cut -c1-1300
grep -o '"parity": {[^}]*}' gpurun_out/bench_n2_final.log | cut -c1-250; grep -o '"exchange": {[^}]*}' gpurun_out/bench_n2_final.log | cut -c1-250
( ROCQ_TRACE_LAUNCHES=1 timeout -s KILL 300 $TR --master-port 29615 bench.py --gpus 2 --steps 1 --warmup 3 --no-parity ) > gpurun_out/trace_n2.log 2>&1
grep "^\[launch\] rank 0" gpurun_out/trace_n2.log | tail -45 > gpurun_out/launches_n2_rank0.log; awk '{print $4, $6, $8}' gpurun_out/launches_n2_rank0.log | tr '\n' ';'
