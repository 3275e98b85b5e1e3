#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; tail -4 gpurun_out/pytest.log
timeout 300 python tools/tc_check.py --quick 2>&1 | grep -E "after|n=30" 
