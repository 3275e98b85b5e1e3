#!/bin/bash
# Round 2, call Z (1 GPU): where the uncached end-to-end step spends its wall time (tools/e2e_profile.py, ROCQ_HOST_PROFILE=1),
# then the default bench line.
mkdir -p gpurun_out
( ROCQ_HOST_PROFILE=1 timeout -s KILL 300 python tools/e2e_profile.py ) > gpurun_out/e2e_profile.log 2>&1; cat gpurun_out/e2e_profile.log | cut -c1-400
( timeout -s KILL 900 python bench.py ) > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-1500
