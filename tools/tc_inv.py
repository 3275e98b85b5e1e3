"""Block sweep at larger sizes: U then U^dagger must return the state; circuit then inverse circuit must return |0..0>."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rocquantum_b200 import workloads
from rocquantum_b200.statevec import StateVector
from tests import util
rng = np.random.default_rng(3)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
g = StateVector(n, "c64")
v = util.random_state(n, seed=1).astype(np.complex64)
for a in range(0, n - 5):
    qs = list(range(a, a + 6))
    U = workloads.haar_unitary(rng, 64)
    g.set_state(v)
    g.apply_block6(qs, U); g.apply_block6(qs, U.conj().T)
    got = g.state()
    print(f"n={n} window {a}: |U^dag U psi - psi| = {util.rel_err(got, v):.3e}", flush=True)
for tc in (0, 1):
    for depth in (1, 2, 3):
        gates = workloads.c2_random_unitary(n, depth, seed=30)
        inv = [(name, t, c, th, np.asarray(M).conj().T) for name, t, c, th, M in reversed(gates)]
        h = StateVector(n, "c64"); h.set_tensor_core_blocks(bool(tc))
        h.apply_circuit(gates); h.apply_circuit(inv)
        print(f"n={n} tc={tc} depth={depth}: <Z0>={h.expect_zprod([0]):.6f} <Z{n-1}>={h.expect_zprod([n-1]):.6f} norm={h.norm2():.7f} blocks={h.stats().blockSweeps}", flush=True)
        del h
