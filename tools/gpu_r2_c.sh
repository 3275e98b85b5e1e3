#!/bin/bash
# Round 2, call C (1 GPU): new tests (exact tolerances, full-depth parity, peaked state, batch expectation, device-side
# sampling, staged readback, C++ backend client), then the bench and the configs.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -15 gpurun_out/pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-qft > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-1800
timeout 600 python tools/config_bench.py > gpurun_out/config_bench.log 2>&1; cut -c1-400 gpurun_out/config_bench.log
