import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rocquantum_b200 import workloads
from rocquantum_b200.statevec import StateVector
n = 30
g = StateVector(n, "c64"); g.gate("h", 0)
U = workloads.haar_unitary(np.random.default_rng(1), 64)
qs = [10, 11, 12, 13, 14, 15]
import sys
if len(sys.argv) > 1: qs = [int(x) for x in sys.argv[1].split(",")]
for _ in range(10): g.apply_block6(qs, U)
g.sync()
for rep in range(3):
    g.timer_start()
    for _ in range(20): g.apply_block6(qs, U)
    print(os.environ.get("ROCQ_BLOCK_DEBUG", "0"), qs, "ms per sweep", g.timer_stop() / 20, flush=True)
