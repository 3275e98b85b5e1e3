#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "exit $?" >> gpurun_out/tc_check.log
timeout 300 python tools/tc_check.py --quick > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_sweep -s 45 -c 2 -o gpurun_out/prof_block python tools/tc_check.py --quick > gpurun_out/ncu_block.log 2>&1
cat gpurun_out/tc_check.log | tail -20
