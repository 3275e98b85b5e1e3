#!/bin/bash
# 2-GPU session: distributed parity + sharded bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "exit $?" >> gpurun_out/dist_check.log
ROCQ_BENCH_QUBITS=31 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2_q31.log 2>&1; echo "exit $?" >> gpurun_out/bench_n2_q31.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "exit $?" >> gpurun_out/bench_n2.log
tail -5 gpurun_out/dist_check.log gpurun_out/bench_n2_q31.log gpurun_out/bench_n2.log
