#!/bin/bash
# Round 2, call AB (1 GPU): batched expectation with the incremental parity word (tests, then configs[4] timing), wide matrices up to 10 qubits.
mkdir -p gpurun_out
( timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bindings.py -m gpu -x -q -k "expectation or apply_matrix_k or rocq_api or plugin" ) > gpurun_out/pytest_expect.log 2>&1; tail -3 gpurun_out/pytest_expect.log
for v in prev cur; do
  if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/$v; fi
  echo "== variant $v"
  timeout -s KILL 300 python tools/config_bench.py --only c5 --reps 3 2>&1 | cut -c1-900
done > gpurun_out/expect_variants.log 2>&1
cat gpurun_out/expect_variants.log
