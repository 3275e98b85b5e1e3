#!/bin/bash
# Round 2, call B (2 GPUs): hardware parity of exactly what SCALE runs (deferring planner, blocks on), NCCL mover and
# the peer-memory mover (ROCQ_EXCHANGE=p2p), then the bench line with both.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo2.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 600 $TR --master-port 29611 tests/dist_check.py ) > gpurun_out/dist_check2_nccl.log 2>&1; tail -4 gpurun_out/dist_check2_nccl.log
( ROCQ_EXCHANGE=p2p timeout 600 $TR --master-port 29612 tests/dist_check.py ) > gpurun_out/dist_check2_p2p.log 2>&1; tail -4 gpurun_out/dist_check2_p2p.log
( timeout 900 $TR --master-port 29613 bench.py --gpus 2 --steps 2 --warmup 3 ) > gpurun_out/bench_n2_nccl.log 2>&1; tail -1 gpurun_out/bench_n2_nccl.log | cut -c1-1200
( ROCQ_EXCHANGE=p2p timeout 900 $TR --master-port 29614 bench.py --gpus 2 --steps 2 --warmup 3 ) > gpurun_out/bench_n2_p2p.log 2>&1; tail -1 gpurun_out/bench_n2_p2p.log | cut -c1-1200
( ROCQ_DIST_INORDER=1 ROCQ_EXCHANGE=p2p timeout 900 $TR --master-port 29615 bench.py --gpus 2 --steps 1 --warmup 3 ) > gpurun_out/bench_n2_inorder_p2p.log 2>&1
grep -o '"exchange": {[^}]*}' gpurun_out/bench_n2_*.log
