#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.log 2>&1; grep '^{' gpurun_out/bench_n2.log | cut -c1-3000
