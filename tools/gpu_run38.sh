#!/bin/bash
# Round 2, first single-GPU call: confirm what round 1 changed on the host after its last GPU run
# (63-pass block plan, shifted-field shot mapping, compiled-reference leg), then refresh the evidence under profiles/.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -4 gpurun_out/pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-1500
# sampling: chunk size of the two-level scan (results are identical by construction; the e2e step uses 256 shots)
for cb in 10 13 16; do
  ROCQ_SAMPLE_CHUNK_BITS=$cb timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-qft > gpurun_out/bench_n1_cb$cb.log 2>&1
  echo "chunk bits $cb: $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench_n1_cb$cb.log)"
done
timeout 600 python tools/config_bench.py > gpurun_out/config_bench.log 2>&1; cut -c1-330 gpurun_out/config_bench.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-1500
# launch list of the same bench command (only after it exited 0 above); shares, not absolutes
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-qft > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches_bench.csv > gpurun_out/launches_bench_summary.md 2>&1; head -12 gpurun_out/launches_bench_summary.md
# QFT-33 complex128 with the two-direction sweep planner (5 sweeps expected): time first, then the launch list
timeout 600 python tools/config_bench.py --only c3 --reps 2 > gpurun_out/config_bench_c3.log 2>&1; cut -c1-400 gpurun_out/config_bench_c3.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_qft33_c128.csv \
    python tools/config_bench.py --only c3 --reps 1 > gpurun_out/ncu_qft33.log 2>&1
python tools/launch_summary.py gpurun_out/launches_qft33_c128.csv > gpurun_out/launches_qft33_c128_summary.md 2>&1; head -14 gpurun_out/launches_qft33_c128_summary.md
