#!/usr/bin/env python
"""bench.py -- headline measurement of the state-vector gate-application path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N = 1   workload = BASELINE configs[1]: 30-qubit random-unitary circuit, depth 40, complex64, one B200.
        One "step" = |0..0> -> the whole 1800-gate circuit, submitted through rocsvxApplyCircuit.
N > 1   (torchrun, one rank per GPU) workload = configs[3]-class circuit on 33 + log2(N) qubits sharded by the
        top log2(N) index bits (36 qubits at N = 8), global<->local index-bit exchanges over NCCL.
Prints ONE JSON line (rank 0).  `value` is gates/s normalised to 30 qubits (gates * 2^(n-30) / s), which at
N = 1 is plain gates/s.  --impl reference times the CPU restatement of the reference's per-gate path
(oracle/, OpenMP over the host cores) on a bounded sample of the same circuit.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  Sampled in-process through NVML (nvidia_ml_py) from a small
    thread, 4 Hz: a polling `nvidia-smi -lms 200` child slowed the N = 2 run it was watching by 23 % (1326 vs 1079 ms per step,
    profiles/r02_bench_n2_sampler_ab.log) -- every iteration of it re-attaches to the driver.  `nvidia-smi` is only the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index=0):
        self.p = None
        self.gpu = gpu_index
        self.thread = None
        self.stop_flag = False
        self.sm, self.reasons, self.sm_max = [], set(), None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.gpu)

    def start(self):
        try:
            nv, h = self._nvml_handle()
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self.stop_flag:
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        mask = int(reasons_fn(h))
                        for bit, name in self.REASONS:
                            if mask & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    time.sleep(0.25)
            import threading
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "1000",
                                       "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(1.0)                              # it attaches to the driver before the timed region, not inside it
        except Exception:
            self.p = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if not self.sm:
                return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no samples"], "source": "nvml"}
            busy = [s for s in self.sm if s > 0.5 * max(self.sm)] or self.sm
            return {"sm_mhz": statistics.median(busy), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(self.sm),
                    "source": "nvml, in-process, 4 Hz"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 1000"}


# ------------------------------------------------------------------------------------------------------
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arms use every host core they may run on and
    report how many that is."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        gomp = C.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(n))
        return int(gomp.omp_get_max_threads())
    except OSError:
        return int(os.environ.get("OMP_NUM_THREADS", n))


def compiled_reference_sample(n_run, n, q1, q2, n1, n2, repeats=2):
    """The reference's OWN code (hipStateVec.cpp + *_kernels.hip compiled unmodified under the host HIP shim,
    oracle/_ref) on the host cores.  It defines no arbitrary-matrix entry point, so the workload's Haar gates
    cannot run on it; every one-qubit gate of the reference goes through the same generic 2x2 kernel
    (hipStateVec.cpp:100-140), so rocsvApplyRy stands for a Haar one-qubit gate at equal cost, and rocsvApplyCNOT --
    the cheapest two-qubit pass it has -- stands for a Haar two-qubit gate (a lower bound on its time)."""
    from oracle import sv_oracle as so
    if not so.ref_available("c64"):
        return None
    r = so.RefLib("c64")
    try:
        r.allocate(n_run)
        r.gate("ry", q1, 0.3); r.gate("cnot", q2[0], q2[1])           # warm-up: page in the state
        t1 = t2 = 0.0
        for _ in range(repeats):
            t = time.perf_counter(); assert r.gate("ry", q1, 0.7) == 0; t1 += time.perf_counter() - t
            t = time.perf_counter(); assert r.gate("cnot", q2[0], q2[1]) == 0; t2 += time.perf_counter() - t
    finally:
        r.close()
    scale = 2.0 ** (n - n_run) / repeats
    t1, t2 = t1 * scale, t2 * scale
    return {"value": (n1 + n2) / (n1 * t1 + n2 * t2) * 2.0 ** (n - 30), "unit": unit_for(n), "kind": "reference",
            "s_per_1q_gate": t1, "s_per_2q_gate": t2,
            "sample": (f"oracle/_ref (the reference's sources under the host HIP shim, OpenMP): rocsvApplyRy on qubit {q1} and "
                       f"rocsvApplyCNOT on qubits {list(q2)} at {n_run} qubits, {repeats} timed calls each"
                       + ("" if n_run == n else f", extrapolated x2^{n - n_run} to {n} qubits")
                       + "; stand-ins of equal (1q) / lower (2q) cost for the workload's Haar gates, which the reference cannot apply")}


def cpu_sample(n, gates, prec="c64", repeats=1):
    """Time the oracle port (one full pass per gate, like the reference's one-kernel-per-gate path) on a few
    gates of the workload.  Returns (seconds per 1q gate, seconds per 2q gate, cores)."""
    from oracle import sv_oracle as so
    from tests import util
    cores = use_all_host_threads()
    o = so.Oracle(n, prec)
    one = [g for g in gates if len(g[1]) == 1][:3]
    two = [g for g in gates if len(g[1]) == 2][:2]
    util.run_on_oracle(o, one[:1] + two[:1])          # warm-up: page in the state
    t1 = t2 = 0.0
    for _ in range(repeats):
        t = time.perf_counter(); util.run_on_oracle(o, one); t1 += (time.perf_counter() - t) / len(one)
        t = time.perf_counter(); util.run_on_oracle(o, two); t2 += (time.perf_counter() - t) / len(two)
    return t1 / repeats, t2 / repeats, cores


def host_ram_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def unit_for(n):
    return "gates/s" if n == 30 else "gates/s (30-qubit equivalent: gates*2^(n-30)/s)"


def workload_for(ngpus):
    from rocquantum_b200 import workloads
    if ngpus == 1:
        n = env_int("ROCQ_BENCH_QUBITS", 30)
        depth = env_int("ROCQ_BENCH_DEPTH", 40)
        return n, workloads.c2_random_unitary(n, depth, seed=30), f"C2: {n}-qubit random-unitary circuit, depth {depth}, fused sweeps, complex64"
    m = ngpus.bit_length() - 1
    n = env_int("ROCQ_BENCH_QUBITS", 33 + m)
    depth = env_int("ROCQ_BENCH_DEPTH", 20)
    return n, workloads.c4_global_layers(n, depth, seed=36), (f"C4: {n}-qubit random circuit, depth {depth}, sharded over {ngpus} GPUs by the top "
                                                             f"{m} qubits, global-qubit swaps via NCCL, complex64")


# ------------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the reference's algorithm (one pass over the state per gate, no fusion) on the host cores."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from oracle import sv_oracle as so
    so.build(ref=False)
    cores = use_all_host_threads()
    n, gates, wl = workload_for(args.gpus)
    n_run = n
    need_gb = (1 << n) * 8 / 1e9 * 1.3
    while n_run > 20 and need_gb > 0.6 * host_ram_gb():
        n_run -= 1
        need_gb /= 2
    n_run = min(n_run, env_int("ROCQ_REF_MAX_QUBITS", 30))
    from rocquantum_b200 import workloads
    sample_gates = workloads.c2_random_unitary(n_run, 1, seed=30) if args.gpus == 1 else workloads.c4_global_layers(n_run, 1, seed=36)
    from tests import util
    o = so.Oracle(n_run, "c64")
    one = [g for g in sample_gates if len(g[1]) == 1][:1]
    two = [g for g in sample_gates if len(g[1]) == 2][:1]
    n1 = sum(1 for g in gates if len(g[1]) == 1)
    n2 = len(gates) - n1
    times = []
    for step in range(args.warmup + args.steps):
        t = time.perf_counter(); util.run_on_oracle(o, one); a = time.perf_counter() - t
        t = time.perf_counter(); util.run_on_oracle(o, two); b = time.perf_counter() - t
        if step >= args.warmup:
            times.append((a, b))
    t1 = statistics.mean(x[0] for x in times) * 2.0 ** (n - n_run)
    t2 = statistics.mean(x[1] for x in times) * 2.0 ** (n - n_run)
    circuit_s = n1 * t1 + n2 * t2
    value = len(gates) / circuit_s * 2.0 ** (n - 30)
    del o
    try:
        compiled = compiled_reference_sample(n_run, n, one[0][1][0], tuple(two[0][1]), n1, n2)
        if compiled:
            compiled["cores"] = cores
    except Exception as ex:                          # the extra datum must not take the reference line down
        compiled = {"error": str(ex)[:200]}
    sample = (f"1 one-qubit + 1 two-qubit Haar gate of the workload per step at {n_run} qubits"
              + ("" if n_run == n else f", extrapolated x2^{n - n_run} to {n} qubits") + "; one full pass per gate (no fusion), OpenMP")
    line = {"impl": "reference", "metric": "gates_per_sec", "value": value, "unit": unit_for(n), "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": circuit_s * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c64", "data": "synthetic", "config": {"workload": wl},
            "cpu_baseline": {"value": value, "unit": unit_for(n), "cores": cores, "kind": "port", "sample": sample,
                             "compiled_reference": compiled},
            "e2e": {"value": value, "unit": unit_for(n), "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def run_engine(args):
    import torch
    import torch.distributed as dist
    from rocquantum_b200 import capi
    from rocquantum_b200.statevec import StateVector

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    ngpus = world
    n, gates, wl = workload_for(ngpus)
    ngates = len(gates)
    lib = capi.load("c64")
    arr, keep = capi.make_ops(gates)

    if ngpus == 1:
        sv = StateVector(n, "c64")
        n_local = n

        def step():
            sv.init()
            st = lib.rocsvxApplyCircuit(sv.h, sv.d, n, arr, ngates)
            assert st == 0, st
    else:
        from rocquantum_b200 import distributed
        sv = distributed.DistStateVector(n, "c64")
        n_local = sv.n_local

        def step():
            sv.init()
            sv.apply_ops(arr, ngates)

    def barrier():
        sv.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sv.stats(reset=True)
    clocks = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("ROCQ_BENCH_NO_CLOCKS"):       # (diagnosis only: is the sampler itself visible in the timing?)
        clocks.start()
    # device time of the K steps: CUDA events on the engine's own stream (rocsvxTimerStart/Stop), no sync inside the region
    sv.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dev_ms = sv.timer_stop()
    barrier()
    elapsed = time.perf_counter() - t0
    clk = clocks.stop() if rank == 0 else None
    st = sv.stats()
    if world > 1:
        t = torch.tensor([elapsed, dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, dev_ms = float(t[0]), float(t[1])
    norm = 2.0 ** (n - 30)
    value = ngates * args.steps / elapsed * norm
    sweeps_per_step = st.sweeps / args.steps

    # ---- the same kernel with ONE gate per sweep (what every rocsvApply* call costs): the HBM-bound regime ------
    one_gate = None
    if ngpus == 1:
        reps = 10
        for q in (0, n // 2, n - 1):
            sv.gate("h", q)
        sv.sync()
        sv.timer_start()
        for r in range(reps):
            sv.gate("h", (7 * r) % n)
        one_ms = sv.timer_stop() / reps
        one_gate = {"avg_launch_ms": one_ms, "achieved": 2.0 * (1 << n) * 8 / (one_ms * 1e-3) / 1e9, "gates": "H on 10 different qubits, one sweep each"}

    # ---- the tensor-core block sweep alone: ten 6-qubit Haar blocks back to back -------------------------------------
    block_alone = None
    if ngpus == 1 and st.blockSweeps > 0:
        from rocquantum_b200 import workloads as wl_
        U = wl_.haar_unitary(np.random.default_rng(5), 64)
        qs = list(range(n // 2 - 3, n // 2 + 3))
        sv.apply_block6(qs, U)
        sv.sync()
        sv.timer_start()
        for r in range(10):
            sv.apply_block6(qs, U)
        blk_ms = sv.timer_stop() / 10
        block_alone = {"avg_launch_ms": blk_ms, "achieved": 2.0 * (1 << n) * 8 / (blk_ms * 1e-3) / 1e9,
                       "what": "block_sweep_kernel, one Haar 64x64 block on qubits %s, ten launches back to back" % qs}

    # ---- end to end through the public call: host gate list in, host result out, every step ----------
    e2e_t, d2h = 0.0, 0
    # no kept plan here: every step converts, fuses and plans the host gate list again and uploads the block operands
    assert lib.rocsvxSetPlanCache(sv.h, 0) == 0
    step(); barrier()                                   # (first uncached step also pays one-time allocator growth)
    sv.stats(reset=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
        if ngpus == 1:
            z = sv.expect_z(0)                      # D2H read of the step's result
            s = sv.sample(list(range(min(n, 64))), 256)
            d2h = 8 + s.nbytes
        else:
            z = sv.expect_z(0)
            d2h = 8
    barrier()
    e2e_t = time.perf_counter() - t0
    st_e2e = sv.stats()
    if world > 1:
        t = torch.tensor([e2e_t], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = float(t[0])
    e2e_value = ngates * args.steps / e2e_t * norm

    # ---- N > 1: the result of exactly this engine configuration, checked.  The same circuit class on n_par qubits over the
    #      N ranks (deferring planner, tensor-core blocks, the exchange mover of this run) against a complex128 run of the
    #      same circuit on rank 0 alone, every amplitude ----------------------------------------------------------------
    parity = None
    if world > 1 and not args.no_parity:
        from rocquantum_b200 import distributed, workloads as wlp
        sv.close()
        m_ = ngpus.bit_length() - 1
        n_par = env_int("ROCQ_BENCH_PARITY_QUBITS", 26 + min(m_, 2))
        pg = wlp.c4_global_layers(n_par, 20, seed=36)
        d = distributed.DistStateVector(n_par, "c64")
        assert d.lib.rocsvxSetTensorCoreBlocks(d.h, 1) == 0          # the slices are below the 'auto' threshold: force the bench's path
        parr, pkeep = capi.make_ops(pg)
        d.apply_ops(parr, len(pg))
        pst = d.stats()
        sl = torch.from_numpy(d.local_slice().view(np.float32)).cuda()
        outs = [torch.empty_like(sl) for _ in range(world)] if rank == 0 else None
        dist.gather(sl, outs, dst=0)
        d.close()
        if rank == 0:
            got = np.concatenate([o.cpu().numpy() for o in outs]).view(np.complex64)
            del outs
            ref = StateVector(n_par, "c128")
            rarr, rkeep = capi.make_ops(pg)
            assert ref.lib.rocsvxApplyCircuit(ref.h, ref.d, n_par, rarr, len(pg)) == 0
            want = ref.state()
            ref.close()
            err = float(np.abs(got - want).max() / np.abs(want).max())
            parity = {"n": n_par, "gates": len(pg), "max_rel_err": err, "tolerance": 1e-5, "ok": bool(err < 1e-5),
                      "exchanges": int(pst.exchanges), "block_sweeps": int(pst.blockSweeps), "sweeps": int(pst.sweeps),
                      "what": "complex64 over %d ranks (blocks on, this run's exchange mover) vs complex128 on rank 0 alone, all 2^%d amplitudes, "
                              "error relative to the largest amplitude" % (world, n_par)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    sweep_bytes = 2.0 * (1 << n_local) * 8
    exchange = None
    if ngpus > 1:
        # the step's device time contains the NCCL exchanges; they are timed separately (CUDA events around each)
        exchange = {"per_step": st.exchanges / args.steps, "ms_per_step": st.exchangeMs / args.steps,
                    "sent_bytes_per_rank_per_step": st.exchangeBytes / args.steps,
                    "send_GBps_per_rank": (st.exchangeBytes / 1e9) / (st.exchangeMs * 1e-3) if st.exchangeMs > 0 else None,
                    "mover": "nccl" if os.environ.get("ROCQ_EXCHANGE", "")[:1] in ("n", "N") else "p2p",
                    "planner": "program order" if os.environ.get("ROCQ_DIST_INORDER", "0") not in ("", "0") else "deferring",
                    "what": "k rank bits <-> top-k local bits; mover nccl = ncclSend/ncclRecv per peer + staging->slice copy, "
                            "p2p = in-place half-swap kernel over IPC-mapped peer slices between two stream barriers; rank 0's figures"}
    sweep_ms = max(0.0, dev_ms - st.exchangeMs) if ngpus > 1 else dev_ms
    avg_sweep_ms = sweep_ms / max(1, st.sweeps)
    achieved = sweep_bytes / (avg_sweep_ms * 1e-3) / 1e9 if avg_sweep_ms > 0 else 0.0
    nblk = st.blockSweeps
    kernel = "block_sweep_kernel (tcgen05, %d of %d launches) + tile_sweep_kernel" % (nblk, st.sweeps) if nblk else "tile_sweep_kernel"
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "frac_of_8TBs_spec": achieved / 8000.0, "peak_source": peak_src,
                "traffic": None, "algorithmic_bytes_per_launch": sweep_bytes, "avg_launch_ms": avg_sweep_ms,
                "launches_per_step": sweeps_per_step, "gates_per_sweep": ngates / max(1.0, sweeps_per_step),
                "note": "every launch is one pass over the state (2 * 2^n * 8 B); avg_launch_ms = device time of the timed region "
                        "/ launches, idle gaps included.  With tensor-core blocks most launches carry one fused 64x64 unitary "
                        "(~8 two-qubit matrices) and run near the HBM roofline; block_sweep_alone / one_gate_sweep time the two "
                        "kernels alone"}
    if block_alone:
        block_alone["frac"] = block_alone["achieved"] / peaks["hbm_gbs"]
        roofline["block_sweep_alone"] = block_alone
    if one_gate:
        one_gate["frac"] = one_gate["achieved"] / peaks["hbm_gbs"]
        one_gate["frac_of_8TBs_spec"] = one_gate["achieved"] / 8000.0
        roofline["one_gate_sweep"] = one_gate
    roofline["tile_ops_per_sweep"] = st.opsExecuted / max(1, st.sweeps)     # after algebraic fusion
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            tj = json.load(open(prof))
            tj = tj["block_sweep_kernel" if nblk * 2 > st.sweeps else "tile_sweep_kernel"]      # the dominant kernel of this run
            # one ncu --set full capture (see the file); DRAM bytes scale with the state, so report per launch of THIS run
            roofline["traffic"] = tj["traffic_over_algorithmic"] * sweep_bytes
            roofline["traffic_source"] = tj["source"]
        except Exception:
            pass

    cpu = None
    if ngpus == 1 and not args.no_cpu:
        try:
            n_cpu = n
            while n_cpu > 20 and (1 << n_cpu) * 8 / 1e9 * 1.3 > 0.6 * host_ram_gb():
                n_cpu -= 1
            t1, t2, cores = cpu_sample(n_cpu, gates)
            scale = 2.0 ** (n - n_cpu)
            n1 = sum(1 for g in gates if len(g[1]) == 1)
            cpu_val = ngates / ((n1 * t1 + (ngates - n1) * t2) * scale)
            cpu = {"value": cpu_val, "unit": "gates/s", "cores": cores, "kind": "port",
                   "sample": f"3 one-qubit + 2 two-qubit Haar gates of the same circuit at {n_cpu} qubits"
                             + ("" if n_cpu == n else f" (extrapolated x2^{n - n_cpu})") + ", one full pass per gate, OpenMP over all cores; "
                             "'port' because the reference never defines rocsvApplyMatrix (SURVEY.md section 0.1)"}
            two = [g for g in gates if len(g[1]) == 2][0]
            cpu["compiled_reference"] = compiled_reference_sample(n_cpu, n, 0, tuple(two[1]), n1, ngates - n1, repeats=1)
            if cpu["compiled_reference"]:
                cpu["compiled_reference"]["cores"] = cores
        except Exception as ex:                      # the baseline must not take the bench down
            cpu = {"value": None, "unit": "gates/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    # ---- configs[0], the reference's own CPU-runnable case: the SAME 20-qubit circuit (named gates only) on the reference's
    #      compiled sources (oracle/_ref, all host cores) and on the engine, in this run, states compared -----------------
    c1 = None
    if ngpus == 1 and not args.no_cpu:
        try:
            from oracle import sv_oracle as so
            from rocquantum_b200 import workloads as wl1
            from tests import util
            if so.ref_available("c64"):
                cores1 = use_all_host_threads()
                g1 = wl1.c1_ghz_random_layers(20, 20, seed=20)
                r = so.RefLib("c64"); r.allocate(20); util.run_on_ref(r, g1[:60]); r.close()       # warm-up: threads, pages
                r = so.RefLib("c64"); r.allocate(20)
                t1 = time.perf_counter(); util.run_on_ref(r, g1); ref_s = time.perf_counter() - t1
                want = r.state(); r.close()
                e = StateVector(20, "c64")
                arr1, keep1 = capi.make_ops(g1)
                for _ in range(3):
                    e.init(); assert lib.rocsvxApplyCircuit(e.h, e.d, 20, arr1, len(g1)) == 0
                e.init(); e.sync(); e.timer_start()
                assert lib.rocsvxApplyCircuit(e.h, e.d, 20, arr1, len(g1)) == 0
                dev1 = e.timer_stop()
                err1 = util.rel_err(e.state(), want)
                e.close()
                c1 = {"workload": "C1: 20-qubit GHZ + 20 random 1q/2q layers, complex64, %d gates" % len(g1),
                      "reference_ms": ref_s * 1e3, "reference_gates_per_s": len(g1) / ref_s, "kind": "reference", "cores": cores1,
                      "engine_device_ms": dev1, "engine_gates_per_s": len(g1) / (dev1 * 1e-3),
                      "max_rel_err_vs_reference": err1, "tolerance": 1e-5,
                      "note": "reference = its hipStateVec.cpp + kernels compiled unmodified under the host HIP shim, one pass per gate; "
                              "engine = rocsvxApplyCircuit, state resident; the 8 MB state sits in L2 on the GPU"}
        except Exception as ex:                                  # never take the bench line down
            c1 = {"error": str(ex)[:200]}

    # ---- second headline of BASELINE.json's metric: HBM GB/s per sweep on the 33-qubit complex128 QFT (configs[2]) ------
    qft = None
    if ngpus == 1 and not args.no_qft:
        try:
            from rocquantum_b200 import workloads as wl_
            sv.close()                                           # 8.6 GB back before the 137 GB state
            nq = env_int("ROCQ_BENCH_QFT_QUBITS", 33)
            qg = wl_.c3_qft(nq, seed=33)
            qarr, qkeep = capi.make_ops(qg)
            q = StateVector(nq, "c128")
            best = None
            for rep in range(2):                                 # first pass: warm-up (plans, allocations)
                q.init(); q.sync(); q.stats(reset=True)
                assert q.lib.rocsvxApplyCircuit(q.h, q.d, nq, qarr, len(qg)) == 0
                q.sync()
                qs = q.stats()
                best = qs
            ms = best.lastSweepMs
            gbs = 2.0 * (1 << nq) * 16 * best.sweeps / (ms * 1e-3) / 1e9
            qft = {"workload": f"C3: {nq}-qubit QFT (benchmarks/run_benchmark.py:60-69) on a seeded basis state, complex128, {(1 << nq) * 16 / 1e9:.1f} GB state",
                   "gates": len(qg), "device_ms": ms, "gates_per_s": len(qg) / (ms * 1e-3), "sweeps": int(best.sweeps),
                   "GBps_per_sweep": gbs, "frac_of_measured_peak": gbs / peaks["hbm_gbs"], "frac_of_8TBs_spec": gbs / 8000.0,
                   "norm": q.norm2(), "note": "H + merged controlled-phase ladders in register windows are FP64/issue-bound "
                   "(DESIGN.md section 8); the swap sweeps of the same circuit run at the HBM roofline"}
            q.close()
        except Exception as ex:                                  # never take the bench line down
            qft = {"error": str(ex)[:200]}

    # ---- BASELINE configs[4]: VQE ansatz + batched Pauli-string expectation + 1M-shot sampling, 28 qubits ---------------
    vqe = None
    if ngpus == 1 and not args.no_vqe:
        try:
            from rocquantum_b200 import workloads as wv
            try:
                sv.close()
            except Exception:
                pass
            nv = env_int("ROCQ_BENCH_VQE_QUBITS", 28)
            layers = []
            for rep in range(4):
                layers += wv.c5_vqe_ansatz(nv, seed=5 + rep)[nv if rep else 0:]
            v = StateVector(nv, "c64")
            varr, vkeep = capi.make_ops(layers)
            for rep in range(3):
                v.init(); v.sync(); v.stats(reset=True)
                assert lib.rocsvxApplyCircuit(v.h, v.d, nv, varr, len(layers)) == 0
                v.sync()
                vs = v.stats()
            ans_ms, ans_sweeps = vs.lastSweepMs, int(vs.sweeps)
            rnd = wv.random_pauli_strings(nv, 64, 8, seed=5)
            ham = wv.hamiltonian_like_terms(nv, 64, seed=5)
            out = {}
            for name, terms in (("random64", rnd), ("hamiltonian64", ham)):
                v.expect_batch(terms)                            # warm-up
                v.stats(reset=True); v.sync(); v.timer_start()
                vals = v.expect_batch(terms)
                ms = v.timer_stop()
                groups = int(v.stats().expectationSweeps)
                out[name] = {"terms": len(terms), "read_sweeps": groups, "device_ms": ms, "ms_per_sweep": ms / max(1, groups),
                             "GBps_per_sweep": (1 << nv) * 8 / (ms / max(1, groups) * 1e-3) / 1e9, "sum": float(np.sum(vals))}
            v.sample(list(range(nv)), 1000)
            t0 = time.perf_counter()
            shots = v.sample(list(range(nv)), 1_000_000)
            samp_ms = (time.perf_counter() - t0) * 1e3
            vqe = {"workload": f"C5: {nv}-qubit hardware-efficient ansatz (examples/vqe_lih.py:74-95 pattern) x4, complex64",
                   "ansatz": {"gates": len(layers), "device_ms": ans_ms, "sweeps": ans_sweeps,
                              "GBps_per_sweep": 2.0 * (1 << nv) * 8 * ans_sweeps / (ans_ms * 1e-3) / 1e9},
                   "expectation_batch": out, "sampling_1M": {"wall_ms": samp_ms, "shots_per_s": 1e6 / (samp_ms * 1e-3),
                                                              "distinct_in_first_10k": int(len(set(shots[:10000].tolist())))},
                   "note": "expectation: terms grouped by x-mask, one read sweep (2^n * 8 B) per group, one D2H for all results; "
                           "random64 = 64 random strings of weight <= 8 (nearly all x-masks distinct), hamiltonian64 = Z / ZZ terms plus "
                           "XX, YY, XY, YX on shared pairs (a molecular Hamiltonian's structure)"}
            v.close()
        except Exception as ex:                                  # never take the bench line down
            vqe = {"error": str(ex)[:200]}

    line = {"metric": "gates_per_sec", "value": value, "unit": unit_for(n),
            "n_gpus": ngpus, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c64", "data": "synthetic",
            "config": {"workload": wl, "qubits": n, "gates": ngates, "l2": "state (8 * 2^n_local bytes) is far larger than the 126 MB L2",
                       "timing": "K steps between stream-sync + barrier, max over ranks; per-sweep time from CUDA events on the engine's stream"},
            "clocks": clk, "device_ms_per_step": dev_ms / args.steps,
            "e2e": {"value": e2e_value, "unit": unit_for(n),
                    "h2d_bytes_per_step": st_e2e.h2dBytes // args.steps, "d2h_bytes_per_step": d2h,
                    "what": "rocsvInitializeState + rocsvxApplyCircuit(host gate list, plan cache OFF: converted, fused, planned and its "
                            "block operands uploaded every step) + <Z0> and 256 sampled bitstrings read back, every step; h2d = block operand "
                            "terms + sweep programs travelling as kernel parameters"},
            "gpu_launches": int(st.kernelLaunches), "roofline": roofline, "cpu_baseline": cpu}
    if exchange:
        line["exchange"] = exchange
    if parity:
        line["parity"] = parity
    if vqe:
        line["vqe_c5"] = vqe
    if c1:
        line["c1_reference"] = c1
    if qft:
        line["qft33_c128"] = qft
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-qft", action="store_true", help="skip the 33-qubit complex128 QFT leg (137 GB)")
    ap.add_argument("--no-vqe", action="store_true", help="skip the configs[4] leg (28-qubit ansatz, expectation batch, 1M shots)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the self-check against the complex128 single-GPU run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
