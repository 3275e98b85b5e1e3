#!/bin/bash
# Round 2, call V (2 GPUs): the N = 2 bench with the in-process NVML clock sampler.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29613 bench.py --gpus 2 --steps 4 --warmup 3 ) > gpurun_out/bench_n2_nvml.log 2>&1
tail -1 gpurun_out/bench_n2_nvml.log | grep -o '"ms_per_step": [0-9.]*\|"device_ms_per_step": [0-9.]*\|"value": [0-9.]*\|"clocks": {[^}]*}\|"parity": {[^}]*}' | cut -c1-300
tail -3 gpurun_out/bench_n2_nvml.log | cut -c1-300
