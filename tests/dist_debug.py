"""Debug helper (torchrun, 2 GPUs): localise distributed mismatches by circuit family."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sv_oracle as so
from rocquantum_b200 import workloads
from rocquantum_b200.distributed import DistStateVector
from tests import util
from tests.dist_check import gather_full

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
prec = "c128"
for n in (14, 15, 17):
    fam = {
        "named": util.random_gates(n, 150, seed=n + world, allow_matrix=False),
        "bag_k1": util.random_gates(n, 150, seed=3 * n + world, maxk=1),
        "bag_k3": util.random_gates(n, 150, seed=3 * n + world, maxk=3),
        "c4": workloads.c4_global_layers(n, 8, seed=36, top=3),
        "c2": workloads.c2_random_unitary(n, 4, seed=30),
        "top1q": [("h", [n - 1], [], 0.0), ("ry", [3], [], 0.4), ("h", [n - 1], [], 0.0), ("rx", [n - 1], [], 0.3), ("cnot", [0], [n - 1], 0.0), ("ry", [n - 1], [], 1.1)],
    }
    for name, gates in fam.items():
        for budget in (None,):
            o = so.Oracle(n, prec); util.run_on_oracle(o, gates)
            d = DistStateVector(n, prec); d.apply_circuit(gates)
            st = d.stats()
            full = gather_full(d, d.local_slice())
            err = util.rel_err(full, o.state)
            nrm = float(np.vdot(full, full).real)
            if rank == 0:
                print(f"n={n} {name:8s} err={err:.3e} norm={nrm:.6f} sweeps={st.sweeps} launches={st.kernelLaunches}", flush=True)
            d.close()
dist.destroy_process_group()
