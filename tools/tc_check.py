#!/usr/bin/env python
"""Validate / time the tensor-core block sweep against the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import sv_oracle as so
from rocquantum_b200 import workloads
from rocquantum_b200.statevec import StateVector
from tests import util

rng = np.random.default_rng(7)
for n, qs in ((13, [5, 6, 7, 8, 9, 10]), (16, [9, 5, 15, 7, 11, 12]), (18, [0, 3, 17, 8, 2, 12]), (14, [0, 1, 2, 3, 4, 5]), (15, [2, 3, 4, 5, 6, 7]), (20, [14, 15, 16, 17, 18, 19])):
    U = workloads.haar_unitary(rng, 64)
    v = util.random_state(n, seed=n)
    o = so.Oracle(n, "c64"); o.set_state(v); o.apply_matrix(qs, U)
    g = StateVector(n, "c64"); g.set_state(v); g.apply_block6(qs, U)
    got = g.state()
    err = util.rel_err(got, o.state)
    print(f"n={n} qubits={qs} rel err {err:.3e} norm {float(np.vdot(got, got).real):.7f}", flush=True)
# drift over many sequential blocks (every amplitude passes through all of them)
n = 20
v = util.random_state(n, seed=3)
o = so.Oracle(n, "c64"); o.set_state(v)
g = StateVector(n, "c64"); g.set_state(v)
for it in range(40):
    qs = sorted(int(x) for x in rng.choice(np.arange(0, n), size=6, replace=False))
    U = workloads.haar_unitary(rng, 64)
    o.apply_matrix(qs, U); g.apply_block6(qs, U)
    if it in (0, 9, 19, 39):
        got = g.state()
        print(f"after {it+1:2d} blocks: rel err {util.rel_err(got, o.state):.3e} norm-1 {float(np.vdot(got.astype(np.complex128), got.astype(np.complex128)).real)-1:+.3e}  (oracle norm-1 {o.norm2()-1:+.3e})", flush=True)
n = 30
g = StateVector(n, "c64")
g.gate("h", 0)
U = workloads.haar_unitary(rng, 64)
for qs in ([10, 11, 12, 13, 14, 15], [24, 25, 26, 27, 28, 29], [5, 9, 13, 17, 21, 25], [0, 1, 2, 3, 4, 5], [3, 4, 5, 6, 7, 8], [1, 2, 3, 4, 5, 6])[: (1 if "--quick" in sys.argv else 6)]:
    for _ in range(5):
        g.apply_block6(qs, U)
    g.sync()
    g.timer_start()
    for _ in range(20):
        g.apply_block6(qs, U)
    ms = g.timer_stop() / 20
    print(f"n=30 block {qs}: {ms:.3f} ms per sweep = {2*(1<<n)*8/(ms*1e-3)/1e9:.0f} GB/s; norm {g.norm2():.6f}", flush=True)
