#!/bin/bash
# single GPU: full gpu test suite, bench, ncu launch list + full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 300 python tools/sweep_bench.py --n 28 --reps 2 --only "H q12" > gpurun_out/plain_sweep.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_sweep -s 2 -c 2 -o gpurun_out/prof_single python tools/sweep_bench.py --n 28 --reps 2 --only "H q12" > gpurun_out/ncu_single.log 2>&1
ROCQ_BENCH_QUBITS=28 ROCQ_BENCH_DEPTH=8 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain_fused.log 2>&1 && \
ROCQ_BENCH_QUBITS=28 ROCQ_BENCH_DEPTH=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_sweep -s 20 -c 2 -o gpurun_out/prof_fused python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_fused.log 2>&1
ROCQ_TILE_BITS=12 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_T12.log 2>&1
tail -3 gpurun_out/pytest.log
