#!/bin/bash
# Round 2, call R (1 GPU): block sweep with the norm ratio measured on every 4th tile (A/B against the previous build, alone and
# inside configs[1] under the power cap), block tests, then ncu --set full of the block sweep for the traffic figure.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "block or c2 or peaked or cache or tensor or sparse or wide" ) > gpurun_out/pytest_blocks.log 2>&1; tail -3 gpurun_out/pytest_blocks.log
for v in prev cur prev cur; do
  if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/$v; fi
  echo "== variant $v"
  timeout 300 python tools/tc_time.py 2>&1 | tail -1
  timeout 300 python tools/config_bench.py --only c2 --reps 5 2>&1 | cut -c1-330
done > gpurun_out/bs_variants4.log 2>&1
cat gpurun_out/bs_variants4.log
unset ROCQ_LIB_DIR
timeout 300 ncu --set full --import-source on --clock-control none -k regex:block_sweep -s 3 -c 1 -f -o gpurun_out/r02_block_sweep \
    python tools/tc_time.py > gpurun_out/ncu_block.log 2>&1
ls -la gpurun_out/r02_block_sweep.ncu-rep
