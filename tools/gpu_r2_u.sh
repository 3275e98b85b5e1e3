#!/bin/bash
# Round 2, call U (2 GPUs): where do the ~0.5 s between wall clock and device time of the N = 2 bench come from?
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mode in sampler nosampler; do
  if [ $mode = nosampler ]; then export ROCQ_BENCH_NO_CLOCKS=1; fi
  ( timeout 300 $TR --master-port 29613 bench.py --gpus 2 --steps 4 --warmup 3 --no-parity ) > gpurun_out/bench_n2_$mode.log 2>&1
  echo "$mode: $(tail -1 gpurun_out/bench_n2_$mode.log | grep -o '"ms_per_step": [0-9.]*\|"device_ms_per_step": [0-9.]*\|"value": [0-9.]*' | head -3 | tr '\n' ' ')"
done
