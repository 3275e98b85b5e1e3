#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "exit $?" >> gpurun_out/dist_check.log
P=29520
for Q in 32 33 34; do
P=$((P+1))
ROCQ_BENCH_QUBITS=$Q ROCQ_BENCH_DEPTH=4 NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_n2_q$Q.log 2>&1; echo "exit $?" >> gpurun_out/bench_n2_q$Q.log
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "exit $?" >> gpurun_out/bench_n2.log
grep -h "DIST CHECK" gpurun_out/dist_check.log | cut -c1-300
