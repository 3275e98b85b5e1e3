#!/bin/bash
mkdir -p gpurun_out
for V in B C D; do
ROCQ_LIB_DIR=$PWD/rocquantum_b200/lib_var/$V timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_var_$V.log 2>&1
done
timeout 900 python tools/config_bench.py > gpurun_out/config_bench.log 2>&1; echo "exit $?" >> gpurun_out/config_bench.log
cat gpurun_out/config_bench.log | cut -c1-400
