#!/bin/bash
# Round 2, first multi-GPU call (gpurun --gpus 2): the two things round 1 wrote after its GPU minutes were spent.
#  1. the deferring distributed planner (default) against the program-order one (ROCQ_DIST_INORDER=1);
#  2. the peer-memory exchange (ROCQ_EXCHANGE=p2p: IPC-mapped slices, in-place half-swap kernel) against NCCL send/recv.
# Parity first (tests/dist_check.py compares every slice with the oracle), then the bench lines.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 600 $TR --master-port 29611 tests/dist_check.py ) > gpurun_out/dist_check2_nccl.log 2>&1; tail -8 gpurun_out/dist_check2_nccl.log
( ROCQ_EXCHANGE=p2p timeout 600 $TR --master-port 29612 tests/dist_check.py ) > gpurun_out/dist_check2_p2p.log 2>&1; tail -8 gpurun_out/dist_check2_p2p.log
( timeout 900 $TR --master-port 29613 bench.py --gpus 2 --steps 2 --warmup 3 ) > gpurun_out/bench_n2_defer_nccl.log 2>&1; tail -1 gpurun_out/bench_n2_defer_nccl.log | cut -c1-600
( ROCQ_EXCHANGE=p2p timeout 900 $TR --master-port 29614 bench.py --gpus 2 --steps 2 --warmup 3 ) > gpurun_out/bench_n2_defer_p2p.log 2>&1; tail -1 gpurun_out/bench_n2_defer_p2p.log | cut -c1-600
# exchange-heavy comparison: the program-order planner issues 20 exchanges per step, so the mover dominates
( ROCQ_DIST_INORDER=1 ROCQ_EXCHANGE=p2p timeout 900 $TR --master-port 29615 bench.py --gpus 2 --steps 1 --warmup 3 ) > gpurun_out/bench_n2_inorder_p2p.log 2>&1
grep -o '"exchange": {[^}]*}' gpurun_out/bench_n2_*.log
