#!/bin/bash
# Round 2, call E (1 GPU): scout-warp column scales in the block sweep (speed vs the round-1 global scale), and the
# single-process multi-GPU group exercised with 2 and 4 slices on one device.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -15 gpurun_out/pytest.log
for v in 0 cur; do
  if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/bs$v; fi
  echo "== variant $v"
  timeout 300 python tools/tc_time.py 2>&1 | tail -1
  timeout 300 python tools/config_bench.py --only c2 --reps 4 2>&1 | cut -c1-330
done > gpurun_out/bs_variants2.log 2>&1
cat gpurun_out/bs_variants2.log
