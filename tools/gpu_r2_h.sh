#!/bin/bash
# Round 2, call H (1 GPU): the QFT butterfly (Hadamard + its ladder as one window op): parity, then configs[2] timing and
# the launch list; full GPU test suite.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -6 gpurun_out/pytest.log
timeout 600 python tools/config_bench.py --only c3,c5 --reps 3 > gpurun_out/config_bench_c3.log 2>&1; cut -c1-420 gpurun_out/config_bench_c3.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_qft33_c128.csv \
    python tools/config_bench.py --only c3 --reps 1 > gpurun_out/ncu_qft33.log 2>&1
python tools/launch_summary.py gpurun_out/launches_qft33_c128.csv > gpurun_out/launches_qft33_c128_summary.md 2>&1; head -14 gpurun_out/launches_qft33_c128_summary.md
grep tile_sweep gpurun_out/launches_qft33_c128.csv | tail -12 | cut -c1-200
