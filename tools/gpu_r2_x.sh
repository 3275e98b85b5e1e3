#!/bin/bash
# Round 2, call X (1 GPU): block sweep with decoupled loader / storer threads and separate input / output tile buffers.
# Quick accuracy check first (a hang stops the call early), block tests, then A/B against the previous build (lib_var/prev),
# alone and inside configs[1] under the power cap.
mkdir -p gpurun_out
timeout -s KILL 180 python tools/tc_check.py > gpurun_out/tc_check_x.log 2>&1 || { echo "tc_check failed or hung"; tail -5 gpurun_out/tc_check_x.log; exit 1; }
tail -3 gpurun_out/tc_check_x.log
( timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "block or c2 or peaked or cache or tensor or sparse or wide" ) > gpurun_out/pytest_blocks.log 2>&1; tail -3 gpurun_out/pytest_blocks.log
for v in prev cur prev cur; do
  if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/$v; fi
  echo "== variant $v"
  timeout -s KILL 300 python tools/tc_time.py 2>&1 | tail -1
  timeout -s KILL 300 python tools/config_bench.py --only c2 --reps 5 2>&1 | cut -c1-330
done > gpurun_out/bs_variants5.log 2>&1
cat gpurun_out/bs_variants5.log
