#!/bin/bash
# Round 2, call F (1 GPU): guess-and-verify column scales in the block sweep: parity of the block tests, then speed against
# the round-1 global scale (lib_var/bs0), alone and inside configs[1].
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "block or c2 or peaked or cache or tensor" ) > gpurun_out/pytest_blocks.log 2>&1; tail -5 gpurun_out/pytest_blocks.log
for v in 0 cur; do
  if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/bs$v; fi
  echo "== variant $v"
  timeout 300 python tools/tc_time.py 2>&1 | tail -1
  timeout 300 python tools/config_bench.py --only c2 --reps 4 2>&1 | cut -c1-330
done > gpurun_out/bs_variants3.log 2>&1
cat gpurun_out/bs_variants3.log
