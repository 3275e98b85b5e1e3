#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_packed.log 2>&1
ROCQ_BENCH_QUBITS=28 ROCQ_BENCH_DEPTH=8 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain_fused.log 2>&1 && \
ROCQ_BENCH_QUBITS=28 ROCQ_BENCH_DEPTH=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_sweep -s 20 -c 4 -o gpurun_out/prof_packed python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_packed.log 2>&1
tail -3 gpurun_out/pytest.log; grep -o '"value": [0-9.]*' gpurun_out/bench_packed.log | head -2
