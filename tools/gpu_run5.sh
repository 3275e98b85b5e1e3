#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_phased.log 2>&1
for B in 60 120 240; do ROCQ_SWEEP_BUDGET=$B timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_budget_$B.log 2>&1; done
timeout 200 python tools/sweep_bench.py --n 30 --json gpurun_out/sweep_c64_T13.json > gpurun_out/sweep_c64_T13.log 2>&1
timeout 200 python tools/sweep_bench.py --n 29 --prec c128 --json gpurun_out/sweep_c128_T12.json > gpurun_out/sweep_c128_T12.log 2>&1
tail -4 gpurun_out/pytest.log
