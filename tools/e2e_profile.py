#!/usr/bin/env python
"""tools/e2e_profile.py -- where the end-to-end step of bench.py (plan cache off) spends its wall time: per call of the
loop (rocsvInitializeState, rocsvxApplyCircuit until it returns, <Z0>, 256 shots) and, with ROCQ_HOST_PROFILE=1, the
engine's own host phases (convert / fuse + plan / block build / launches)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocquantum_b200 import capi, workloads  # noqa: E402
from rocquantum_b200.statevec import StateVector  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
gates = workloads.c2_random_unitary(n, 40, seed=30)
arr, keep = capi.make_ops(gates)
sv = StateVector(n, "c64")
lib = sv.lib
assert lib.rocsvxSetPlanCache(sv.h, 0) == 0
for rep in range(4):
    sv.sync()
    t0 = time.perf_counter()
    sv.init()
    t1 = time.perf_counter()
    assert lib.rocsvxApplyCircuit(sv.h, sv.d, n, arr, len(gates)) == 0
    t2 = time.perf_counter()
    sv.sync()
    t3 = time.perf_counter()
    z = sv.expect_z(0)
    t4 = time.perf_counter()
    s = sv.sample(list(range(min(n, 64))), 256)
    t5 = time.perf_counter()
    st = sv.stats(reset=True)
    print(f"rep {rep}: init {1e3*(t1-t0):.2f} | ApplyCircuit returns after {1e3*(t2-t1):.2f} | drained after {1e3*(t3-t2):.2f} more | <Z0> {1e3*(t4-t3):.2f} | 256 shots {1e3*(t5-t4):.2f} "
          f"| total {1e3*(t5-t0):.2f} ms; device {st.lastSweepMs:.2f} ms, {st.sweeps} sweeps", flush=True)
