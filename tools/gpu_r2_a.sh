#!/bin/bash
# Round 2, call A (1 GPU): state of HEAD before any round-2 change -- tests, bench line, configs, launch list.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -4 gpurun_out/pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-2500
timeout 600 python tools/config_bench.py > gpurun_out/config_bench.log 2>&1; cut -c1-400 gpurun_out/config_bench.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-qft > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches_bench.csv > gpurun_out/launches_bench_summary.md 2>&1; head -12 gpurun_out/launches_bench_summary.md
nvidia-smi topo -m > gpurun_out/topo.log 2>&1; nvidia-smi -L
