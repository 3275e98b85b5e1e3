#!/bin/bash
mkdir -p gpurun_out
for D in 0 8 1 9 3 11 5 13 7 15; do ROCQ_BLOCK_DEBUG=$D timeout 120 python tools/tc_time.py; done > gpurun_out/tc_time.log 2>&1
cat gpurun_out/tc_time.log
