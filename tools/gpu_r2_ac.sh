#!/bin/bash
# Round 2, call AC (1 GPU): default pool with the raised release threshold (e2e profile, cold 1M-shot sampling), batched
# expectation A/B (prev = POPC per term and amplitude), expectation / sampling tests.
mkdir -p gpurun_out
( timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "expectation or sampling or status or export" ) > gpurun_out/pytest_expect.log 2>&1; tail -3 gpurun_out/pytest_expect.log
( ROCQ_HOST_PROFILE=1 timeout -s KILL 300 python tools/e2e_profile.py ) 2>&1 | grep "^rep" | cut -c1-300 > gpurun_out/e2e_profile_defaultpool.log; cat gpurun_out/e2e_profile_defaultpool.log
for v in prev cur; do
  if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/$v; fi
  echo "== variant $v"
  timeout -s KILL 300 python tools/config_bench.py --only c5 --reps 2 2>&1 | grep -v ansatz | cut -c1-400
done > gpurun_out/expect_variants2.log 2>&1
cat gpurun_out/expect_variants2.log
