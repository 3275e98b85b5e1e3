#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/tc_time.py > gpurun_out/tc_time_plain.log 2>&1 || exit 1
cat gpurun_out/tc_time_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_sweep -s 12 -c 1 -f -o gpurun_out/prof_block5 python tools/tc_time.py > gpurun_out/ncu_block5.log 2>&1
tail -3 gpurun_out/ncu_block5.log
