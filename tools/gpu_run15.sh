#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "exit $?" >> gpurun_out/tc_check.log
cat gpurun_out/tc_check.log
for D in 0 8 1 2 4 7; do ROCQ_BLOCK_DEBUG=$D timeout 120 python tools/tc_time.py; done > gpurun_out/tc_time.log 2>&1
ROCQ_BLOCK_DEBUG=0 timeout 120 python tools/tc_time.py 24,25,26,27,28,29 >> gpurun_out/tc_time.log 2>&1
ROCQ_BLOCK_DEBUG=7 timeout 120 python tools/tc_time.py 24,25,26,27,28,29 >> gpurun_out/tc_time.log 2>&1
cat gpurun_out/tc_time.log
