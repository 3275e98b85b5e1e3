#!/bin/bash
mkdir -p gpurun_out
for v in main old b1 b2 v6; do
  if [ $v = main ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=/root/repo/lib_var/$v; fi
  echo "=== $v"
  timeout 200 python tools/sweep_bench.py --n 30 --prec c64 --reps 5 --only "q0" 2>&1 | cut -c1-90 | grep -v "^#"
  timeout 200 python tools/sweep_bench.py --n 30 --prec c64 --reps 5 --only "Rz" 2>&1 | cut -c1-90 | grep -v "^#"
  timeout 200 python tools/sweep_bench.py --n 29 --prec c128 --reps 5 --only "q0" 2>&1 | cut -c1-90 | grep -v "^#"
  if [ $v = main ] || [ $v = v6 ]; then
    timeout 300 python tools/config_bench.py --only c5 --reps 2 2>&1 | cut -c1-200 | head -1
    timeout 300 python tools/config_bench.py --only c3 --reps 1 --c3-qubits 30 2>&1 | cut -c1-200
  fi
done > gpurun_out/ab_variants2.log 2>&1
cat gpurun_out/ab_variants2.log
