#!/usr/bin/env python
"""tools/sweep_bench.py -- per-sweep HBM roofline of the tile-sweep kernel, one gate per sweep.

For each (gate, target placement) runs R eager rocsvApply* calls (R sweeps) between two CUDA events on the
handle's stream and prints achieved GB/s = 2 * 2^n * sizeof(amp) / t_sweep against the measured copy peak.
    python tools/sweep_bench.py [--n 30] [--prec c64] [--reps 10] [--json out.json]
"""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from rocquantum_b200 import workloads  # noqa: E402
from rocquantum_b200.statevec import StateVector  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=30)
    ap.add_argument("--prec", default="c64")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None, help="substring filter on case names")
    a = ap.parse_args()
    n = a.n
    peak = 6549.1
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    sv = StateVector(n, a.prec)
    sv.gate("h", 0)
    amp = 8 if a.prec == "c64" else 16
    sweep_bytes = 2.0 * (1 << n) * amp
    rng = np.random.default_rng(0)
    U2 = workloads.haar_unitary(rng, 4)
    U5, U6, U7 = workloads.haar_unitary(rng, 32), workloads.haar_unitary(rng, 64), workloads.haar_unitary(rng, 128)
    cases = []
    for t in sorted({0, 1, 3, 4, 5, 8, 12, 13, 20, n - 2, n - 1}):
        if t < n:
            cases.append((f"H q{t}", lambda t=t: sv.gate("h", t)))
    cases += [
        ("X q0", lambda: sv.gate("x", 0)), (f"X q{n-1}", lambda: sv.gate("x", n - 1)),
        ("Rz q7 (diag)", lambda: sv.gate("rz", 7, 0.3)), (f"Z q{n-1} (phase)", lambda: sv.gate("z", n - 1)),
        (f"CNOT 0->{n-1}", lambda: sv.gate("cnot", 0, n - 1)), (f"CNOT {n-1}->0", lambda: sv.gate("cnot", n - 1, 0)),
        (f"CZ 3,{n-2}", lambda: sv.gate("cz", 3, n - 2)), (f"SWAP 2,{n-1}", lambda: sv.gate("swap", 2, n - 1)),
        (f"CRY {n-1}->6", lambda: sv.gate("cry", n - 1, 6, 0.4)),
        ("U4 q3,q4 (circuit)", lambda: sv.apply_circuit([("matrix", [3, 4], [], 0.0, U2)])),
        (f"U4 q{n-2},q{n-1} (circuit)", lambda: sv.apply_circuit([("matrix", [n - 2, n - 1], [], 0.0, U2)])),
        (f"U4 q2,q{n-1} (circuit)", lambda: sv.apply_circuit([("matrix", [2, n - 1], [], 0.0, U2)])),
        # eager rocsvApplyMatrix with wide matrices: 5 / 6 qubits take the tensor-core block sweep (complex64), 7 the gather kernel
        ("U32 q8..12 (eager, device matrix)", lambda: sv.apply_matrix([8, 9, 10, 11, 12], U5)),
        ("U64 q8..13 (eager, device matrix)", lambda: sv.apply_matrix([8, 9, 10, 11, 12, 13], U6)),
        (f"U32 q{n-5}..{n-1} (eager, device matrix)", lambda: sv.apply_matrix(list(range(n - 5, n)), U5)),
        ("U128 q8..14 (eager, gather kernel)", lambda: sv.apply_matrix([8, 9, 10, 11, 12, 13, 14], U7)),
        ("chunk masses + scan + 256 shots (read sweep)", lambda: sv.sample(list(range(min(n, 64))), 256)),
        ("8 Z-strings in one batch (read sweep)", lambda: sv.expect_batch([("ZZ", [q, q + 1]) for q in range(8)])),
        ("norm (read sweep)", lambda: sv.norm2()),
        ("<X5 Y9 Z20> (read sweep)", lambda: sv.expect_pauli("XYZ", [5, 9, min(20, n - 1)])),
    ]
    rows = []
    print(f"# n={n} {a.prec}: sweep = {sweep_bytes/1e9:.2f} GB algorithmic, peak {peak:.0f} GB/s (measured copy), tile bits env={os.environ.get('ROCQ_TILE_BITS','default')}")
    for name, fn in cases:
        if a.only and a.only not in name:
            continue
        fn(); fn()
        sv.sync()
        sv.timer_start()
        for _ in range(a.reps):
            fn()
        ms = sv.timer_stop() / a.reps
        nbytes = sweep_bytes / 2 if "read sweep" in name else sweep_bytes
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append(dict(case=name, ms=ms, gbs=gbs, frac=gbs / peak))
        print(f"{name:32s} {ms:9.3f} ms  {gbs:8.1f} GB/s  {100*gbs/peak:5.1f}% of measured peak  {100*gbs/8000:5.1f}% of 8 TB/s")
    if a.json:
        json.dump(dict(n=n, prec=a.prec, peak=peak, tile_bits=os.environ.get("ROCQ_TILE_BITS"), rows=rows), open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
