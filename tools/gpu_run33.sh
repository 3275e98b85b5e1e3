#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py ) > gpurun_out/dist_check2.log 2>&1; grep -E "world=|DIST CHECK|real" gpurun_out/dist_check2.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.log 2>&1; grep '^{' gpurun_out/bench_n2.log | cut -c1-2500
ROCQ_TC=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 1 --warmup 3 > gpurun_out/bench_n2_notc.log 2>&1; grep '^{' gpurun_out/bench_n2_notc.log | cut -c1-400
