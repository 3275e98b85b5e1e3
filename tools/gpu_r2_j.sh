#!/bin/bash
# Round 2, call J (1 GPU): thread-factor tables for the QFT window phases; 2 vs 3 resident CTAs per SM for the phased kernel.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "qft or ladders or c3 or c2_and or mixed" ) > gpurun_out/pytest_qft.log 2>&1; tail -4 gpurun_out/pytest_qft.log
for v in cur mb3; do
  if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/$v; fi
  echo "== variant $v"
  timeout 600 python tools/config_bench.py --only c3,c5 --reps 3 2>&1 | cut -c1-330
  ROCQ_TC=0 timeout 600 python tools/config_bench.py --only c2 --reps 2 2>&1 | cut -c1-330
done > gpurun_out/qft_variants.log 2>&1
cat gpurun_out/qft_variants.log
