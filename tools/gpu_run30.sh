#!/bin/bash
mkdir -p gpurun_out
for v in v0 v1 v2 v3; do
  if [ $v = v0 ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=/root/repo/lib_var/$v; fi
  echo "=== $v"
  timeout 200 python tools/sweep_bench.py --n 30 --prec c64 --reps 5 --only "q0" 2>&1 | cut -c1-90 | grep -v "^#"
  timeout 200 python tools/sweep_bench.py --n 30 --prec c64 --reps 5 --only "Rz" 2>&1 | cut -c1-90 | grep -v "^#"
  timeout 200 python tools/sweep_bench.py --n 29 --prec c128 --reps 5 --only "q0" 2>&1 | cut -c1-90 | grep -v "^#"
  timeout 300 python tools/config_bench.py --only c5 --reps 1 2>&1 | cut -c1-200 | head -1
  timeout 300 python tools/config_bench.py --only c3 --reps 1 --c3-qubits 30 2>&1 | cut -c1-200
  ROCQ_TC=0 timeout 300 python tools/config_bench.py --only c2 --reps 1 2>&1 | cut -c1-200
done > gpurun_out/ab_variants.log 2>&1
cat gpurun_out/ab_variants.log
