#!/usr/bin/env python
"""tools/config_bench.py -- one timed pass over every single-GPU configuration of BASELINE.json (configs[0,1,2,4]).

Prints one JSON object per config: circuit wall time (device, CUDA events around the flush), gates/s, sweeps, achieved
GB/s per sweep, plus the read-sweep numbers of C5 (batched Pauli expectation, 1M-shot sampling).
    python tools/config_bench.py [--only c1,c2,c3,c5] [--reps 3]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocquantum_b200 import capi, workloads  # noqa: E402
from rocquantum_b200.statevec import StateVector  # noqa: E402


def run_circuit(name, n, prec, gates, reps):
    sv = StateVector(n, prec)
    arr, keep = capi.make_ops(gates)
    amp = 8 if prec == "c64" else 16
    best = None
    for r in range(reps + 1):
        sv.init()
        sv.sync()
        sv.stats(reset=True)
        t0 = time.perf_counter()
        st = sv.lib.rocsvxApplyCircuit(sv.h, sv.d, n, arr, len(gates))
        assert st == 0, st
        sv.sync()
        wall = time.perf_counter() - t0
        s = sv.stats()
        if r == 0:
            continue                      # warm-up
        if best is None or s.lastSweepMs < best["device_ms"]:
            best = dict(config=name, qubits=n, prec=prec, gates=len(gates), device_ms=s.lastSweepMs, wall_ms=wall * 1e3, sweeps=int(s.sweeps),
                        launches=int(s.kernelLaunches), gates_per_s=len(gates) / (s.lastSweepMs * 1e-3),
                        gbs_per_sweep=2.0 * (1 << n) * amp * s.sweeps / (s.lastSweepMs * 1e-3) / 1e9,
                        norm=sv.norm2())
    print(json.dumps(best), flush=True)
    return sv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c1,c2,c3,c5")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--c3-qubits", type=int, default=33)
    a = ap.parse_args()
    only = set(a.only.split(","))
    if "c1" in only:
        run_circuit("C1 GHZ + 20 random layers", 20, "c64", workloads.c1_ghz_random_layers(20, 20, seed=20), a.reps).close()
    if "c2" in only:
        run_circuit("C2 random-unitary depth 40", 30, "c64", workloads.c2_random_unitary(30, 40, seed=30), a.reps).close()
    if "c3" in only:
        n = a.c3_qubits
        run_circuit(f"C3 QFT-{n} complex128", n, "c128", workloads.c3_qft(n, seed=33), max(1, a.reps - 1)).close()
    if "c5" in only:
        n = 28
        layers = []
        for rep in range(4):
            layers += workloads.c5_vqe_ansatz(n, seed=5 + rep)[n if rep else 0:]
        sv = run_circuit("C5 VQE ansatz x4", n, "c64", layers, a.reps)
        terms = workloads.random_pauli_strings(n, 64, 8, seed=5)
        sv.sync()
        sv.timer_start()
        vals = [sv.expect_pauli(p, q) for p, q in terms]
        ms = sv.timer_stop()
        print(json.dumps(dict(config="C5 64 Pauli strings (weight<=8)", qubits=n, device_ms=ms, terms=len(terms),
                              gbs=len(terms) * (1 << n) * 8 / (ms * 1e-3) / 1e9, sum=sum(vals))), flush=True)
        for name, bt in (("random64", terms), ("hamiltonian64", workloads.hamiltonian_like_terms(n, 64, seed=5))):
            sv.expect_batch(bt)                                   # warm-up
            sv.stats(reset=True); sv.sync(); sv.timer_start()
            bv = sv.expect_batch(bt)
            ms = sv.timer_stop()
            groups = int(sv.stats().expectationSweeps)
            print(json.dumps(dict(config=f"C5 batched expectation, {name}", qubits=n, device_ms=ms, terms=len(bt), read_sweeps=groups,
                                  ms_per_sweep=ms / max(1, groups), sum=float(sum(bv)))), flush=True)
        sv.sample(list(range(n)), 1000)                           # warm-up (first call sizes the scratch)
        t0 = time.perf_counter()
        s = sv.sample(list(range(n)), 1_000_000)
        dt = time.perf_counter() - t0
        print(json.dumps(dict(config="C5 1M-shot sampling", qubits=n, wall_ms=dt * 1e3, shots_per_s=1e6 / dt, distinct=int(len(set(s[:10000].tolist()))))),
              flush=True)
        sv.close()


if __name__ == "__main__":
    main()
