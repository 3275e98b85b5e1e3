#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 900 python tools/config_bench.py --only c1,c3,c5 > gpurun_out/config_bench.log 2>&1; echo "exit $?" >> gpurun_out/config_bench.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1
tail -3 gpurun_out/pytest.log; cut -c1-300 gpurun_out/config_bench.log
