#!/bin/bash
# tools/gpu_calls.sh -- the exact command lists of the round-2 GPU sessions, one function per call, kept as the record of how
# every file under profiles/r02_* was produced:   gpurun [--gpus N] -- 'bash tools/gpu_calls.sh <call>'
# (calls a..w: first session of the round; x..ad: second session.  lib_var/<name> = a variant library built with
#  ROCQ_LIB_DIR / ROCQ_OBJ_DIR / ROCQ_EXTRA_DEFS, see tools/README.md; "prev" = the library of the commit before the change under test.)
mkdir -p gpurun_out

# Round 2, call A (1 GPU): state of HEAD before any round-2 change -- tests, bench line, configs, launch list.
call_a() {
  ( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -4 gpurun_out/pytest.log
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-2500
  timeout 600 python tools/config_bench.py > gpurun_out/config_bench.log 2>&1; cut -c1-400 gpurun_out/config_bench.log
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-qft > gpurun_out/ncu_bench.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_bench.csv > gpurun_out/launches_bench_summary.md 2>&1; head -12 gpurun_out/launches_bench_summary.md
  nvidia-smi topo -m > gpurun_out/topo.log 2>&1; nvidia-smi -L
}

# Round 2, call B (2 GPUs): hardware parity of exactly what SCALE runs (deferring planner, blocks on), NCCL mover and
# the peer-memory mover (ROCQ_EXCHANGE=p2p), then the bench line with both.
call_b() {
  nvidia-smi topo -m > gpurun_out/topo2.log 2>&1
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
  ( timeout 600 $TR --master-port 29611 tests/dist_check.py ) > gpurun_out/dist_check2_nccl.log 2>&1; tail -4 gpurun_out/dist_check2_nccl.log
  ( ROCQ_EXCHANGE=p2p timeout 600 $TR --master-port 29612 tests/dist_check.py ) > gpurun_out/dist_check2_p2p.log 2>&1; tail -4 gpurun_out/dist_check2_p2p.log
  ( timeout 900 $TR --master-port 29613 bench.py --gpus 2 --steps 2 --warmup 3 ) > gpurun_out/bench_n2_nccl.log 2>&1; tail -1 gpurun_out/bench_n2_nccl.log | cut -c1-1200
  ( ROCQ_EXCHANGE=p2p timeout 900 $TR --master-port 29614 bench.py --gpus 2 --steps 2 --warmup 3 ) > gpurun_out/bench_n2_p2p.log 2>&1; tail -1 gpurun_out/bench_n2_p2p.log | cut -c1-1200
  ( ROCQ_DIST_INORDER=1 ROCQ_EXCHANGE=p2p timeout 900 $TR --master-port 29615 bench.py --gpus 2 --steps 1 --warmup 3 ) > gpurun_out/bench_n2_inorder_p2p.log 2>&1
  grep -o '"exchange": {[^}]*}' gpurun_out/bench_n2_*.log
}

# Round 2, call C (1 GPU): new tests (exact tolerances, full-depth parity, peaked state, batch expectation, device-side
# sampling, staged readback, C++ backend client), then the bench and the configs.
call_c() {
  ( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -15 gpurun_out/pytest.log
  timeout 900 python bench.py --steps 5 --warmup 3 --no-qft > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-1800
  timeout 600 python tools/config_bench.py > gpurun_out/config_bench.log 2>&1; cut -c1-400 gpurun_out/config_bench.log
}

# Round 2, call D (1 GPU): what does the per-column scale of the block sweep cost, and where?  Four builds of block_sweep.cu
# (RQ_BS_SCALE 0 = global scale as in round 1, 1 = per column / registers, 2 = per column / shared memory read twice,
# 3 = no exchange: wrong results, timing only), each timed alone (tc_time: 20 launches) and inside configs[1] (63 passes).
call_d() {
  for v in 0 1 2 3; do
    if [ $v = 1 ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/bs$v; fi
    echo "== variant $v"
    timeout 300 python tools/tc_time.py 2>&1 | tail -1
    timeout 300 python tools/config_bench.py --only c2 --reps 4 2>&1 | cut -c1-330
  done > gpurun_out/bs_variants.log 2>&1
  cat gpurun_out/bs_variants.log
}

# Round 2, call E (1 GPU): scout-warp column scales in the block sweep (speed vs the round-1 global scale), and the
# single-process multi-GPU group exercised with 2 and 4 slices on one device.
call_e() {
  ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -15 gpurun_out/pytest.log
  for v in 0 cur; do
    if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/bs$v; fi
    echo "== variant $v"
    timeout 300 python tools/tc_time.py 2>&1 | tail -1
    timeout 300 python tools/config_bench.py --only c2 --reps 4 2>&1 | cut -c1-330
  done > gpurun_out/bs_variants2.log 2>&1
  cat gpurun_out/bs_variants2.log
}

# Round 2, call F (1 GPU): guess-and-verify column scales in the block sweep: parity of the block tests, then speed against
# the round-1 global scale (lib_var/bs0), alone and inside configs[1].
call_f() {
  ( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "block or c2 or peaked or cache or tensor" ) > gpurun_out/pytest_blocks.log 2>&1; tail -5 gpurun_out/pytest_blocks.log
  for v in 0 cur; do
    if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/bs$v; fi
    echo "== variant $v"
    timeout 300 python tools/tc_time.py 2>&1 | tail -1
    timeout 300 python tools/config_bench.py --only c2 --reps 4 2>&1 | cut -c1-330
  done > gpurun_out/bs_variants3.log 2>&1
  cat gpurun_out/bs_variants3.log
}

# Round 2, call G (2 GPUs): (1) the GPU tests that need two devices (group mode on two real devices, torchrun dist_check),
# (2) bench --gpus 2 with its parity self-check, (3) the single-process group on the same circuit + its ncu launch list.
call_g() {
  ( timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_dist.py -m gpu -x -q ) > gpurun_out/pytest_2gpu.log 2>&1; tail -4 gpurun_out/pytest_2gpu.log
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
  ( timeout 900 $TR --master-port 29613 bench.py --gpus 2 --steps 2 --warmup 3 ) > gpurun_out/bench_n2.log 2>&1; tail -1 gpurun_out/bench_n2.log | cut -c1-2500
  grep -o '"parity": {[^}]*}' gpurun_out/bench_n2.log; grep -o '"exchange": {[^}]*}' gpurun_out/bench_n2.log | cut -c1-300
  timeout 900 python tools/group_bench.py > gpurun_out/group_bench_2.log 2>&1; cat gpurun_out/group_bench_2.log | cut -c1-700
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_group2.csv \
      python tools/group_bench.py --steps 1 --parity-qubits 0 > gpurun_out/ncu_group2.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_group2.csv > gpurun_out/launches_group2_summary.md 2>&1; head -16 gpurun_out/launches_group2_summary.md
}

# Round 2, call H (1 GPU): the QFT butterfly (Hadamard + its ladder as one window op): parity, then configs[2] timing and
# the launch list; full GPU test suite.
call_h() {
  ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -6 gpurun_out/pytest.log
  timeout 600 python tools/config_bench.py --only c3,c5 --reps 3 > gpurun_out/config_bench_c3.log 2>&1; cut -c1-420 gpurun_out/config_bench_c3.log
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_qft33_c128.csv \
      python tools/config_bench.py --only c3 --reps 1 > gpurun_out/ncu_qft33.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_qft33_c128.csv > gpurun_out/launches_qft33_c128_summary.md 2>&1; head -14 gpurun_out/launches_qft33_c128_summary.md
  grep tile_sweep gpurun_out/launches_qft33_c128.csv | tail -12 | cut -c1-200
}

# Round 2, call I (1 GPU): QFT butterfly -- parity tests that run QFTs, then configs[2] timing + launch list.
call_i() {
  ( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "qft or ladders or c3 or c2_and or smoke" ) > gpurun_out/pytest_qft.log 2>&1; tail -4 gpurun_out/pytest_qft.log
  timeout 600 python tools/config_bench.py --only c3 --reps 3 > gpurun_out/config_bench_c3.log 2>&1; cut -c1-420 gpurun_out/config_bench_c3.log
  timeout 600 python tools/config_bench.py --only c3 --reps 3 --c3-qubits 30 2>&1 | cut -c1-420
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_qft33_c128.csv \
      python tools/config_bench.py --only c3 --reps 1 > gpurun_out/ncu_qft33.log 2>&1
  grep tile_sweep gpurun_out/launches_qft33_c128.csv | tail -5 | awk -F'","' '{print $5, $NF}' | cut -c1-200
}

# Round 2, call J (1 GPU): thread-factor tables for the QFT window phases; 2 vs 3 resident CTAs per SM for the phased kernel.
call_j() {
  ( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "qft or ladders or c3 or c2_and or mixed" ) > gpurun_out/pytest_qft.log 2>&1; tail -4 gpurun_out/pytest_qft.log
  for v in cur mb3; do
    if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/$v; fi
    echo "== variant $v"
    timeout 600 python tools/config_bench.py --only c3,c5 --reps 3 2>&1 | cut -c1-330
    ROCQ_TC=0 timeout 600 python tools/config_bench.py --only c2 --reps 2 2>&1 | cut -c1-330
  done > gpurun_out/qft_variants.log 2>&1
  cat gpurun_out/qft_variants.log
}

# Round 2, call K (1 GPU): ncu --set full of the QFT window sweeps (28 qubits complex128, first two sweeps) and of the block
# sweep (28 qubits), with source correlation; timing first (never under the profiler).
call_k() {
  timeout 600 python tools/config_bench.py --only c3 --reps 2 2>&1 | cut -c1-330
  timeout 600 python tools/config_bench.py --only c3 --reps 2 --c3-qubits 28 2>&1 | cut -c1-330
  ncu --set full --import-source on --clock-control none -k regex:tile_sweep -c 2 -f -o gpurun_out/r02_qft28_c128 \
      python tools/config_bench.py --only c3 --c3-qubits 28 --reps 1 > gpurun_out/ncu_qft28.log 2>&1
  ls -la gpurun_out/*.ncu-rep
}

# Round 2, call L (2 GPUs): bench --gpus 2 with its parity self-check; the single-process group on the same circuit
# (34 qubits over 2 devices) and its ncu launch list (one process, so ncu can follow it).
call_l() {
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
  ( timeout 900 $TR --master-port 29613 bench.py --gpus 2 --steps 2 --warmup 3 ) > gpurun_out/bench_n2.log 2>&1; tail -1 gpurun_out/bench_n2.log | cut -c1-1500
  grep -o '"parity": {[^}]*}' gpurun_out/bench_n2.log; grep -o '"exchange": {[^}]*}' gpurun_out/bench_n2.log | cut -c1-300
  timeout 900 python tools/group_bench.py > gpurun_out/group_bench_2.log 2>&1; cat gpurun_out/group_bench_2.log | cut -c1-800
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_group2.csv \
      python tools/group_bench.py --steps 1 --parity-qubits 0 > gpurun_out/ncu_group2.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_group2.csv > gpurun_out/launches_group2_summary.md 2>&1; head -16 gpurun_out/launches_group2_summary.md
}

# Round 2, call M (1 GPU): everything since call J on one device -- full GPU tests (butterfly chains, eager 5/6-qubit
# matrices through the block sweep, NVTX build, batch binding), QFT timing, and where the host time of a distributed step
# goes (2 slices on one device, host profile on).
call_m() {
  ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -6 gpurun_out/pytest.log
  timeout 600 python tools/config_bench.py --only c3 --reps 3 2>&1 | cut -c1-330
  timeout 600 python tools/config_bench.py --only c3 --reps 2 --c3-qubits 30 2>&1 | cut -c1-330
  ROCQ_HOST_PROFILE=1 timeout 600 python tools/group_bench.py --ranks 2 --qubits 31 --parity-qubits 0 --steps 2 > gpurun_out/group_bench_1gpu.log 2>&1; tail -25 gpurun_out/group_bench_1gpu.log | cut -c1-600
}

# Round 2, call N (1 GPU): the failed test again, then ncu --set full of the QFT chain-phase sweeps (28 qubits c128) and
# launch list of QFT-33.
call_n() {
  ( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "eager_wide or apply_matrix" ) > gpurun_out/pytest_wide.log 2>&1; tail -3 gpurun_out/pytest_wide.log
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:tile_sweep -c 2 -f -o gpurun_out/r02_qft28_c128_chain \
      python tools/config_bench.py --only c3 --c3-qubits 28 --reps 1 > gpurun_out/ncu_qft28.log 2>&1
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_qft33_c128.csv \
      python tools/config_bench.py --only c3 --reps 1 > gpurun_out/ncu_qft33.log 2>&1
  grep tile_sweep gpurun_out/launches_qft33_c128.csv | tail -5 | awk -F'","' '{print $5, $NF}' | cut -c1-200
}

# Round 2, call O (8 GPUs): hardware parity of the distributed engine at 8 ranks (dist_check: eager, deferred, whole-circuit,
# sampling, measurement, index-bit swaps incl. global<->global, blocks on slices), the bench line with its parity self-check
# (36 qubits), and rank 0's launch list of one step (ROCQ_TRACE_LAUNCHES: CUDA events around every launch).
call_o() {
  nvidia-smi topo -m > gpurun_out/topo8.log 2>&1
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
  ( timeout 240 $TR --master-port 29611 tests/dist_check.py ) > gpurun_out/dist_check8.log 2>&1; tail -5 gpurun_out/dist_check8.log
  ( timeout 360 $TR --master-port 29613 bench.py --gpus 8 --steps 2 --warmup 3 ) > gpurun_out/bench_n8.log 2>&1; tail -1 gpurun_out/bench_n8.log | cut -c1-1400
  grep -o '"parity": {[^}]*}' gpurun_out/bench_n8.log; grep -o '"exchange": {[^}]*}' gpurun_out/bench_n8.log | cut -c1-330
  ( ROCQ_TRACE_LAUNCHES=1 timeout 240 $TR --master-port 29615 bench.py --gpus 8 --steps 1 --warmup 3 --no-parity ) > gpurun_out/trace_n8.log 2>&1
  grep "^\[launch\] rank 0" gpurun_out/trace_n8.log | tail -60 > gpurun_out/launches_n8_rank0.log; tail -3 gpurun_out/launches_n8_rank0.log
}

# Round 2, call P (1 GPU): block tiles whose column bits are not the lowest seven (two-run blocks): parity + speed;
# full test suite; the distributed plan on 2 slices of one device (wall vs device, launch list).
call_p() {
  ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -6 gpurun_out/pytest.log
  for qs in 10,11,12,13,14,15 11,12,13,27,28,29 8,9,26,27,28,29 24,25,26,27,28,29 4,5,6,27,28,29; do timeout 120 python tools/tc_time.py $qs 2>&1 | tail -1; done
  ROCQ_TRACE_LAUNCHES=1 timeout 600 python tools/group_bench.py --ranks 2 --qubits 31 --parity-qubits 26 --steps 1 > gpurun_out/group_bench_1gpu.log 2>&1
  grep -v "^\[launch\]" gpurun_out/group_bench_1gpu.log | cut -c1-700; grep "^\[launch\] rank 0" gpurun_out/group_bench_1gpu.log | tail -42 | awk '{print $4, $6, $8}' | tr '\n' ';'
}

# Round 2, call Q (1 GPU): compute-sanitizer racecheck + memcheck over the GPU tests that exercise both hot kernels
# (the reference has a real shared-memory race at swap_kernels.hip:108), then the default bench line.
call_q() {
  ( timeout 500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "block6_tensor or mixed_bag or two_runs or swap_index" ) > gpurun_out/racecheck.log 2>&1
  tail -12 gpurun_out/racecheck.log | cut -c1-300
  ( timeout 500 compute-sanitizer --tool memcheck --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -m gpu -x -q -k "block6_tensor or mixed_bag or two_runs or sampling_and or batched or reference_multi" ) > gpurun_out/memcheck.log 2>&1
  tail -8 gpurun_out/memcheck.log | cut -c1-300
  ( time timeout 900 python bench.py ) > gpurun_out/bench_default.log 2>&1; tail -4 gpurun_out/bench_default.log | cut -c1-3000
}

# Round 2, call R (1 GPU): block sweep with the norm ratio measured on every 4th tile (A/B against the previous build, alone and
# inside configs[1] under the power cap), block tests, then ncu --set full of the block sweep for the traffic figure.
call_r() {
  ( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "block or c2 or peaked or cache or tensor or sparse or wide" ) > gpurun_out/pytest_blocks.log 2>&1; tail -3 gpurun_out/pytest_blocks.log
  for v in prev cur prev cur; do
    if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/$v; fi
    echo "== variant $v"
    timeout 300 python tools/tc_time.py 2>&1 | tail -1
    timeout 300 python tools/config_bench.py --only c2 --reps 5 2>&1 | cut -c1-330
  done > gpurun_out/bs_variants4.log 2>&1
  cat gpurun_out/bs_variants4.log
  unset ROCQ_LIB_DIR
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:block_sweep -s 3 -c 1 -f -o gpurun_out/r02_block_sweep \
      python tools/tc_time.py > gpurun_out/ncu_block.log 2>&1
  ls -la gpurun_out/r02_block_sweep.ncu-rep
}

# Round 2, call S (1 GPU): full GPU test suite, smoke(), the default bench line, one-gate / wide-matrix / reduction sweeps.
call_s() {
  ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -5 gpurun_out/pytest.log
  ( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
  ( time timeout 900 python bench.py ) > gpurun_out/bench_default.log 2>&1; tail -4 gpurun_out/bench_default.log | cut -c1-1200
  timeout 600 python tools/sweep_bench.py > gpurun_out/sweep_bench_c64.log 2>&1; cat gpurun_out/sweep_bench_c64.log | cut -c1-200
  timeout 600 python tools/sweep_bench.py --prec c128 --n 29 > gpurun_out/sweep_bench_c128.log 2>&1; tail -12 gpurun_out/sweep_bench_c128.log | cut -c1-200
}

# Round 2, call T (2 GPUs): bench --gpus 2 with the final build (block tiles with chosen columns on the slices), its launch
# list from rank 0 (ROCQ_TRACE_LAUNCHES), and the single-process group on the two devices.
call_t() {
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
  ( timeout 400 $TR --master-port 29613 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/bench_n2_final.log 2>&1; tail -1 gpurun_out/bench_n2_final.log | cut -c1-1300
  grep -o '"parity": {[^}]*}' gpurun_out/bench_n2_final.log | cut -c1-250; grep -o '"exchange": {[^}]*}' gpurun_out/bench_n2_final.log | cut -c1-250
  ( ROCQ_TRACE_LAUNCHES=1 timeout 300 $TR --master-port 29615 bench.py --gpus 2 --steps 1 --warmup 3 --no-parity ) > gpurun_out/trace_n2.log 2>&1
  grep "^\[launch\] rank 0" gpurun_out/trace_n2.log | tail -45 > gpurun_out/launches_n2_rank0.log; awk '{print $4, $6, $8}' gpurun_out/launches_n2_rank0.log | tr '\n' ';'
  timeout 300 python tools/group_bench.py --parity-qubits 0 --steps 2 > gpurun_out/group_bench_2_final.log 2>&1; cat gpurun_out/group_bench_2_final.log | cut -c1-800
}

# Round 2, call U (2 GPUs): where do the ~0.5 s between wall clock and device time of the N = 2 bench come from?
call_u() {
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
  for mode in sampler nosampler; do
    if [ $mode = nosampler ]; then export ROCQ_BENCH_NO_CLOCKS=1; fi
    ( timeout 300 $TR --master-port 29613 bench.py --gpus 2 --steps 4 --warmup 3 --no-parity ) > gpurun_out/bench_n2_$mode.log 2>&1
    echo "$mode: $(tail -1 gpurun_out/bench_n2_$mode.log | grep -o '"ms_per_step": [0-9.]*\|"device_ms_per_step": [0-9.]*\|"value": [0-9.]*' | head -3 | tr '\n' ' ')"
  done
}

# Round 2, call V (2 GPUs): the N = 2 bench with the in-process NVML clock sampler.
call_v() {
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
  ( timeout 300 $TR --master-port 29613 bench.py --gpus 2 --steps 4 --warmup 3 ) > gpurun_out/bench_n2_nvml.log 2>&1
  tail -1 gpurun_out/bench_n2_nvml.log | grep -o '"ms_per_step": [0-9.]*\|"device_ms_per_step": [0-9.]*\|"value": [0-9.]*\|"clocks": {[^}]*}\|"parity": {[^}]*}' | cut -c1-300
  tail -3 gpurun_out/bench_n2_nvml.log | cut -c1-300
}

# Round 2, call W (8 GPUs): the N = 8 bench line (36 qubits) with the final build -- parity self-check, exchange figures --
# and rank 0's launch list of one step.
call_w() {
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
  ( timeout 300 $TR --master-port 29613 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/bench_n8_final.log 2>&1; tail -1 gpurun_out/bench_n8_final.log | cut -c1-1300
  grep -o '"parity": {[^}]*}' gpurun_out/bench_n8_final.log | cut -c1-250; grep -o '"exchange": {[^}]*}' gpurun_out/bench_n8_final.log | cut -c1-250
  ( ROCQ_TRACE_LAUNCHES=1 timeout 200 $TR --master-port 29615 bench.py --gpus 8 --steps 1 --warmup 3 --no-parity ) > gpurun_out/trace_n8.log 2>&1
  grep "^\[launch\] rank 0" gpurun_out/trace_n8.log | tail -45 > gpurun_out/launches_n8_rank0_final.log; awk '{print $4, $6, $8}' gpurun_out/launches_n8_rank0_final.log | tail -42 | tr '\n' ';'
}

# Round 2, call X (1 GPU): block sweep with decoupled loader / storer threads and separate input / output tile buffers.
# Quick accuracy check first (a hang stops the call early), block tests, then A/B against the previous build (lib_var/prev),
# alone and inside configs[1] under the power cap.
call_x() {
  timeout -s KILL 180 python tools/tc_check.py > gpurun_out/tc_check_x.log 2>&1 || { echo "tc_check failed or hung"; tail -5 gpurun_out/tc_check_x.log; exit 1; }
  tail -3 gpurun_out/tc_check_x.log
  ( timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "block or c2 or peaked or cache or tensor or sparse or wide" ) > gpurun_out/pytest_blocks.log 2>&1; tail -3 gpurun_out/pytest_blocks.log
  for v in prev cur prev cur; do
    if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/$v; fi
    echo "== variant $v"
    timeout -s KILL 300 python tools/tc_time.py 2>&1 | tail -1
    timeout -s KILL 300 python tools/config_bench.py --only c2 --reps 5 2>&1 | cut -c1-330
  done > gpurun_out/bs_variants5.log 2>&1
  cat gpurun_out/bs_variants5.log
}

# Round 2, call Y (1 GPU): full GPU test suite, smoke(), the default bench line with the decoupled block sweep, then ncu --set full of it.
call_y() {
  ( time timeout -s KILL 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -5 gpurun_out/pytest.log
  ( timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
  ( time timeout -s KILL 900 python bench.py ) > gpurun_out/bench_default.log 2>&1; tail -4 gpurun_out/bench_default.log | cut -c1-1500
  timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:block_sweep -s 3 -c 1 -f -o gpurun_out/r02_block_sweep_decoupled \
      python tools/tc_time.py > gpurun_out/ncu_block.log 2>&1
  ls -la gpurun_out/r02_block_sweep_decoupled.ncu-rep
}

# Round 2, call Z (1 GPU): where the uncached end-to-end step spends its wall time (tools/e2e_profile.py, ROCQ_HOST_PROFILE=1),
# then the default bench line.
call_z() {
  ( ROCQ_HOST_PROFILE=1 timeout -s KILL 300 python tools/e2e_profile.py ) > gpurun_out/e2e_profile.log 2>&1; cat gpurun_out/e2e_profile.log | cut -c1-400
  ( timeout -s KILL 900 python bench.py ) > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-1500
}

# Round 2, call AA (2 GPUs): the decoupled block sweep and the handle-owned memory pool on distributed slices -- 2-GPU tests
# (group + torchrun), the N = 2 bench line with its parity object, rank 0's launch list.
call_aa() {
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
  ( timeout -s KILL 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_dist.py -m gpu -x -q ) > gpurun_out/pytest_2gpu.log 2>&1; tail -3 gpurun_out/pytest_2gpu.log
  ( timeout -s KILL 300 $TR --master-port 29611 tests/dist_check.py ) > gpurun_out/dist_check2.log 2>&1; tail -3 gpurun_out/dist_check2.log | cut -c1-300
  ( timeout -s KILL 400 $TR --master-port 29613 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/bench_n2_final.log 2>&1; tail -1 gpurun_out/bench_n2_final.log | cut -c1-1300
  grep -o '"parity": {[^}]*}' gpurun_out/bench_n2_final.log | cut -c1-250; grep -o '"exchange": {[^}]*}' gpurun_out/bench_n2_final.log | cut -c1-250
  ( ROCQ_TRACE_LAUNCHES=1 timeout -s KILL 300 $TR --master-port 29615 bench.py --gpus 2 --steps 1 --warmup 3 --no-parity ) > gpurun_out/trace_n2.log 2>&1
  grep "^\[launch\] rank 0" gpurun_out/trace_n2.log | tail -45 > gpurun_out/launches_n2_rank0.log; awk '{print $4, $6, $8}' gpurun_out/launches_n2_rank0.log | tr '\n' ';'
}

# Round 2, call AB (1 GPU): batched expectation with the incremental parity word (tests, then configs[4] timing), wide matrices up to 10 qubits.
call_ab() {
  ( timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bindings.py -m gpu -x -q -k "expectation or apply_matrix_k or rocq_api or plugin" ) > gpurun_out/pytest_expect.log 2>&1; tail -3 gpurun_out/pytest_expect.log
  for v in prev cur; do
    if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/$v; fi
    echo "== variant $v"
    timeout -s KILL 300 python tools/config_bench.py --only c5 --reps 3 2>&1 | cut -c1-900
  done > gpurun_out/expect_variants.log 2>&1
  cat gpurun_out/expect_variants.log
}

# Round 2, call AC (1 GPU): default pool with the raised release threshold (e2e profile, cold 1M-shot sampling), batched
# expectation A/B (prev = POPC per term and amplitude), expectation / sampling tests.
call_ac() {
  ( timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "expectation or sampling or status or export" ) > gpurun_out/pytest_expect.log 2>&1; tail -3 gpurun_out/pytest_expect.log
  ( ROCQ_HOST_PROFILE=1 timeout -s KILL 300 python tools/e2e_profile.py ) 2>&1 | grep "^rep" | cut -c1-300 > gpurun_out/e2e_profile_defaultpool.log; cat gpurun_out/e2e_profile_defaultpool.log
  for v in prev cur; do
    if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/$v; fi
    echo "== variant $v"
    timeout -s KILL 300 python tools/config_bench.py --only c5 --reps 2 2>&1 | grep -v ansatz | cut -c1-400
  done > gpurun_out/expect_variants2.log 2>&1
  cat gpurun_out/expect_variants2.log
}

# Round 2, call AD (1 GPU): complex64 window-phase sweeps (the CUDA-core path: configs[1] with ROCQ_TC=0, the VQE ansatz) under
# build variants -- cur = 3 CTAs/SM at 80 registers, mb2 = 2 CTAs/SM at 128, mb2u = mb2 + warp-uniform trip counts, w3 = 3-bit windows.
call_ad() {
  for v in cur mb2 mb2u w3; do
    if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/$v; fi
    echo "== variant $v"
    ROCQ_TC=0 timeout -s KILL 300 python tools/config_bench.py --only c2 --reps 2 2>&1 | cut -c1-330
    timeout -s KILL 300 python tools/config_bench.py --only c5 --reps 2 2>&1 | grep ansatz | cut -c1-330
  done > gpurun_out/window_variants.log 2>&1
  cat gpurun_out/window_variants.log
}

# Round 2, call AE (1 GPU): trailing swaps folded into the store addresses (QFT sweeps lose their permutation passes): the
# tests that run QFTs / swaps / fused bags, then QFT-33 complex128 and its launch list.
call_ae() {
  ( timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "qft or ladders or mixed_bag or named_gate or c1_config or golden or fixtures or swap" ) > gpurun_out/pytest_qft.log 2>&1; tail -3 gpurun_out/pytest_qft.log
  timeout -s KILL 600 python tools/config_bench.py --only c3 --reps 2 2>&1 | cut -c1-400 | tee gpurun_out/config_bench_c3.log
  ( ROCQ_TRACE_LAUNCHES=1 timeout -s KILL 600 python tools/config_bench.py --only c3 --reps 1 ) 2>&1 | grep "^\[launch\]" | tail -6 | tee gpurun_out/launches_qft33_c128_trace.log
}

# Round 2, call AF (1 GPU): complex128 QFT sweeps -- qprev = before, qa = chain parameters fetched inside the group loop,
# cur = qa + the swizzle pass that enumerates only the pairs.  Tests first (QFTs, ladders, fused bags: both layouts), then QFT-33 per variant.
call_af() {
  ( timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "qft or ladders or mixed_bag or named_gate or c1_config or golden or c2_and" ) > gpurun_out/pytest_qft.log 2>&1; tail -3 gpurun_out/pytest_qft.log
  for v in qprev qa cur; do
    if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/$v; fi
    echo "== variant $v"
    ( ROCQ_TRACE_LAUNCHES=1 timeout -s KILL 600 python tools/config_bench.py --only c3 --reps 1 ) 2>&1 | grep "^\[launch\]\|device_ms" | tail -7 | cut -c1-200
  done > gpurun_out/qft_variants2.log 2>&1
  cat gpurun_out/qft_variants2.log
}

# Round 2, call AG (1 GPU): ncu --set full of two complex64 window-phase sweeps (configs[1] on the CUDA-core path, ROCQ_TC=0).
call_ag() {
  ROCQ_TC=0 timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:tile_sweep -s 8 -c 2 -f -o gpurun_out/r02_window_c64 \
      python tools/config_bench.py --only c2 --reps 1 > gpurun_out/ncu_window_c64.log 2>&1
  ls -la gpurun_out/r02_window_c64.ncu-rep
}

# Round 2, call AH (1 GPU): ncu --set full of the first two complex128 QFT-28 sweeps after the store map / chain-register / swizzle changes.
call_ah() {
  timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:tile_sweep -c 2 -f -o gpurun_out/r02_qft28_c128_final \
      python tools/config_bench.py --only c3 --c3-qubits 28 --reps 1 > gpurun_out/ncu_qft28.log 2>&1
  ls -la gpurun_out/r02_qft28_c128_final.ncu-rep
}

# Round 2, call AI (1 GPU): butterfly chains with per-tile thread-factor tables and the bitwise group base (qb = before).
call_ai() {
  ( timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "qft or ladders or mixed_bag or named_gate or c1_config or golden or c2_and" ) > gpurun_out/pytest_qft.log 2>&1; tail -3 gpurun_out/pytest_qft.log
  for v in qb cur; do
    if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/$v; fi
    echo "== variant $v"
    ( ROCQ_TRACE_LAUNCHES=1 timeout -s KILL 600 python tools/config_bench.py --only c3 --reps 1 ) 2>&1 | grep "^\[launch\]\|device_ms" | tail -7 | cut -c1-200
  done > gpurun_out/qft_variants3.log 2>&1
  cat gpurun_out/qft_variants3.log
}

# Round 2, call AJ (8 GPUs): the N = 8 bench line (36 qubits) with the final build of the round -- parity self-check, exchange
# figures; two timed steps to keep the 8-GPU charge small.
call_aj() {
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
  ( ROCQ_TRACE_LAUNCHES=1 timeout -s KILL 300 $TR --master-port 29613 bench.py --gpus 8 --steps 2 --warmup 3 ) > gpurun_out/bench_n8_final2.log 2>&1
  grep -v "^\[launch\]" gpurun_out/bench_n8_final2.log | tail -1 | cut -c1-1300
  grep -o '"parity": {[^}]*}' gpurun_out/bench_n8_final2.log | cut -c1-250; grep -o '"exchange": {[^}]*}' gpurun_out/bench_n8_final2.log | cut -c1-250
  grep "^\[launch\] rank 0" gpurun_out/bench_n8_final2.log | tail -45 > gpurun_out/launches_n8_rank0_final2.log; awk '{print $4, $6, $8}' gpurun_out/launches_n8_rank0_final2.log | tail -42 | tr '\n' ';'
}

# Round 2, call AK (1 GPU): batched expectation with the sign-word kernel for groups of >= 8 terms (qb = POPC per term and amplitude).
call_ak() {
  ( timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bindings.py tests/test_gpu_group.py -m gpu -x -q -k "expectation or rocq_api or plugin or group_matches" ) > gpurun_out/pytest_expect.log 2>&1; tail -3 gpurun_out/pytest_expect.log
  for v in qb cur; do
    if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/$v; fi
    echo "== variant $v"
    timeout -s KILL 300 python tools/config_bench.py --only c5 --reps 2 2>&1 | grep "batched" | cut -c1-300
  done > gpurun_out/expect_variants3.log 2>&1
  cat gpurun_out/expect_variants3.log
}

# Round 2, call AL (1 GPU): sign-word expectation kernel from 4 terms, groups capped at 16 terms (qb = POPC per term, groups of 32).
call_al() {
  ( timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bindings.py tests/test_gpu_group.py -m gpu -x -q -k "expectation or rocq_api or plugin or group_matches" ) > gpurun_out/pytest_expect.log 2>&1; tail -3 gpurun_out/pytest_expect.log
  for v in qb cur; do
    if [ $v = cur ]; then unset ROCQ_LIB_DIR; else export ROCQ_LIB_DIR=$PWD/lib_var/$v; fi
    echo "== variant $v"
    timeout -s KILL 300 python tools/expect_split.py 2>&1 | cut -c1-200
    timeout -s KILL 300 python tools/config_bench.py --only c5 --reps 2 2>&1 | grep "batched" | cut -c1-300
  done > gpurun_out/expect_split2.log 2>&1
  cat gpurun_out/expect_split2.log
}

if [ -z "$1" ] || ! declare -F "call_$1" > /dev/null; then echo "usage: bash tools/gpu_calls.sh {a|b|c|d|e|f|g|h|i|j|k|l|m|n|o|p|q|r|s|t|u|v|w|x|y|z|aa|ab|ac|ad|ae|af|ag|ah|ai|aj|ak|al}"; exit 2; fi
"call_$1"
