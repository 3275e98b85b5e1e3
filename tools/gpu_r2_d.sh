#!/bin/bash
# Round 2, call D (1 GPU): what does the per-column scale of the block sweep cost, and where?  Four builds of block_sweep.cu
# (RQ_BS_SCALE 0 = global scale as in round 1, 1 = per column / registers, 2 = per column / shared memory read twice,
# 3 = no exchange: wrong results, timing only), each timed alone (tc_time: 20 launches) and inside configs[1] (63 passes).
mkdir -p gpurun_out
for v in 0 1 2 3; do
  if [ $v = 1 ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/bs$v; fi
  echo "== variant $v"
  timeout 300 python tools/tc_time.py 2>&1 | tail -1
  timeout 300 python tools/config_bench.py --only c2 --reps 4 2>&1 | cut -c1-330
done > gpurun_out/bs_variants.log 2>&1
cat gpurun_out/bs_variants.log
