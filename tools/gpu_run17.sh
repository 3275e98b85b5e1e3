#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "exit $?" >> gpurun_out/tc_check.log
cat gpurun_out/tc_check.log
for D in 16 0; do ROCQ_BLOCK_DEBUG=$D timeout 120 python tools/tc_time.py; done > gpurun_out/tc_phase.log 2>&1
ROCQ_BLOCK_DEBUG=16 timeout 120 python tools/tc_time.py 0,1,2,3,4,5 >> gpurun_out/tc_phase.log 2>&1
cat gpurun_out/tc_phase.log
