#!/bin/bash
# First GPU session: parity, smoke, per-sweep roofline, tile-size sweep, first bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 200 python tools/sweep_bench.py --n 30 --json gpurun_out/sweep_c64_T13.json > gpurun_out/sweep_c64_T13.log 2>&1
ROCQ_TILE_BITS=12 timeout 200 python tools/sweep_bench.py --n 30 --json gpurun_out/sweep_c64_T12.json > gpurun_out/sweep_c64_T12.log 2>&1
ROCQ_TILE_BITS=11 timeout 200 python tools/sweep_bench.py --n 30 --json gpurun_out/sweep_c64_T11.json > gpurun_out/sweep_c64_T11.log 2>&1
timeout 200 python tools/sweep_bench.py --n 29 --prec c128 --json gpurun_out/sweep_c128_T12.json > gpurun_out/sweep_c128_T12.log 2>&1
ROCQ_TILE_BITS=11 timeout 200 python tools/sweep_bench.py --n 29 --prec c128 --json gpurun_out/sweep_c128_T11.json > gpurun_out/sweep_c128_T11.log 2>&1
for B in 40 80 160 320 1000000; do
  ROCQ_SWEEP_BUDGET=$B timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_budget_$B.log 2>&1
done
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -3 gpurun_out/smoke.log gpurun_out/pytest.log gpurun_out/bench.log
