#!/bin/bash
# Round 2, call S (1 GPU): full GPU test suite, smoke(), the default bench line, one-gate / wide-matrix / reduction sweeps.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -5 gpurun_out/pytest.log
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
( time timeout 900 python bench.py ) > gpurun_out/bench_default.log 2>&1; tail -4 gpurun_out/bench_default.log | cut -c1-1200
timeout 600 python tools/sweep_bench.py > gpurun_out/sweep_bench_c64.log 2>&1; cat gpurun_out/sweep_bench_c64.log | cut -c1-200
timeout 600 python tools/sweep_bench.py --prec c128 --n 29 > gpurun_out/sweep_bench_c128.log 2>&1; tail -12 gpurun_out/sweep_bench_c128.log | cut -c1-200
