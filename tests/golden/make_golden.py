"""Generate tests/golden/*.npy from the REFERENCE ITSELF: its hipStateVec.cpp + kernels compiled unmodified
under the host HIP shim (make -C oracle ref; needs /root/reference).  Run from the repo root:
    python tests/golden/make_golden.py
The fixtures are small (2^10 amplitudes) and committed (c1 / mixed: the reference's rocsv* entry points; c2: its ApplyMatrix
spec kernel driven by oracle/hip_shim/spec_driver.cpp); the GPU box never needs /root/reference."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import sv_oracle as so  # noqa: E402
from rocquantum_b200 import workloads  # noqa: E402
from tests import util  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
so.build(ref=True)
for prec in ("c64", "c128"):
    r = so.RefLib(prec)
    r.allocate(10, 1)
    util.run_on_ref(r, workloads.c1_ghz_random_layers(10, 6, seed=20))
    np.save(os.path.join(HERE, f"c1_n10_{prec}.npy"), r.state())
    r.close()
    r = so.RefLib(prec)
    r.allocate(10, 1)
    util.run_on_ref(r, util.random_gates(10, 200, seed=4242, allow_matrix=False))
    np.save(os.path.join(HERE, f"mixed_n10_{prec}.npy"), r.state())
    r.close()
    # configs[1] in miniature: Haar one- and two-qubit gates.  The reference defines no entry point for them, but it ships
    # the kernel rocsvApplyMatrix was meant to launch (multi_qubit_kernels.hip:37-115); oracle/hip_shim/spec_driver.cpp
    # launches it, gate by gate.
    r = so.RefLib(prec)
    r.allocate(10, 1)
    for g in workloads.c2_random_unitary(10, 6, seed=30):
        r.spec_apply_matrix(list(g[1]), g[4])
    np.save(os.path.join(HERE, f"c2_n10_{prec}.npy"), r.state())
    r.close()
print("golden fixtures written to", HERE)
