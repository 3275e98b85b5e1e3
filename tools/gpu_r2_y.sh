#!/bin/bash
# Round 2, call Y (1 GPU): full GPU test suite, smoke(), the default bench line with the decoupled block sweep, then ncu --set full of it.
mkdir -p gpurun_out
( time timeout -s KILL 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -5 gpurun_out/pytest.log
( timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
( time timeout -s KILL 900 python bench.py ) > gpurun_out/bench_default.log 2>&1; tail -4 gpurun_out/bench_default.log | cut -c1-1500
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:block_sweep -s 3 -c 1 -f -o gpurun_out/r02_block_sweep_decoupled \
    python tools/tc_time.py > gpurun_out/ncu_block.log 2>&1
ls -la gpurun_out/r02_block_sweep_decoupled.ncu-rep
