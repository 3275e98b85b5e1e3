#!/usr/bin/env python
"""tools/expect_split.py -- where the batched expectation of Hamiltonian-like terms spends its time: the all-Z group alone
(one read sweep, up to 32 sign patterns), the XX/YY/XY/YX quadruples alone (one sweep per x-mask), and groups of 1..32 Z terms."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocquantum_b200 import workloads  # noqa: E402
from rocquantum_b200.statevec import StateVector  # noqa: E402

n = 28
sv = StateVector(n, "c64")
sv.apply_circuit(workloads.c5_vqe_ansatz(n, seed=5))
ham = workloads.hamiltonian_like_terms(n, 64, seed=5)
zs = [t for t in ham if set(t[0]) <= {"Z"}]
xs = [t for t in ham if not set(t[0]) <= {"Z"}]


def timed(terms, label):
    sv.expect_batch(terms)
    sv.stats(reset=True); sv.sync(); sv.timer_start()
    sv.expect_batch(terms)
    ms = sv.timer_stop()
    g = int(sv.stats().expectationSweeps)
    print(f"{label}: {len(terms)} terms, {g} read sweeps, {ms:.3f} ms device, {ms / max(1, g):.3f} ms per sweep", flush=True)


timed(zs, "all-Z terms")
timed(xs, "XX/YY/XY/YX quadruples")
for k in (1, 2, 4, 8, 16, 32):
    timed([("Z", [q % n]) for q in range(k)], f"{k} single-Z terms")
