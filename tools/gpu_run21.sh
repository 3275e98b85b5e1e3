#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tc_inv.py 26 > gpurun_out/tc_inv.log 2>&1
timeout 600 python tools/tc_inv.py 30 2>&1 | grep -v window >> gpurun_out/tc_inv.log
cat gpurun_out/tc_inv.log
