#!/bin/bash
# 8-GPU session: distributed parity (4 ranks) + the 36-qubit bench line
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 tests/dist_check.py > gpurun_out/dist_check4.log 2>&1; echo "exit $?" >> gpurun_out/dist_check4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tests/dist_check.py > gpurun_out/dist_check8.log 2>&1; echo "exit $?" >> gpurun_out/dist_check8.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_n8.log 2>&1; echo "exit $?" >> gpurun_out/bench_n8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/bench_n4.log 2>&1; echo "exit $?" >> gpurun_out/bench_n4.log
grep -h "DIST CHECK" gpurun_out/dist_check4.log gpurun_out/dist_check8.log | cut -c1-200; grep -o '"value": [0-9.]*' gpurun_out/bench_n8.log gpurun_out/bench_n4.log
