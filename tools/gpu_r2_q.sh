#!/bin/bash
# Round 2, call Q (1 GPU): compute-sanitizer racecheck + memcheck over the GPU tests that exercise both hot kernels
# (the reference has a real shared-memory race at swap_kernels.hip:108), then the default bench line.
mkdir -p gpurun_out
( timeout 500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "block6_tensor or mixed_bag or two_runs or swap_index" ) > gpurun_out/racecheck.log 2>&1
tail -12 gpurun_out/racecheck.log | cut -c1-300
( timeout 500 compute-sanitizer --tool memcheck --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -m gpu -x -q -k "block6_tensor or mixed_bag or two_runs or sampling_and or batched or reference_multi" ) > gpurun_out/memcheck.log 2>&1
tail -8 gpurun_out/memcheck.log | cut -c1-300
( time timeout 900 python bench.py ) > gpurun_out/bench_default.log 2>&1; tail -4 gpurun_out/bench_default.log | cut -c1-3000
