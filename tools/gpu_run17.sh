#!/bin/bash
mkdir -p gpurun_out
for D in 16 0 4; do ROCQ_BLOCK_DEBUG=$D timeout 120 python tools/tc_time.py; done > gpurun_out/tc_phase.log 2>&1
cat gpurun_out/tc_phase.log
