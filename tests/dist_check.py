"""Run under torchrun (one rank per GPU): distributed engine vs the oracle on the full state.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py
Exits non-zero on any mismatch.  Launched by tests/test_gpu_dist.py when >= 2 GPUs are visible."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sv_oracle as so  # noqa: E402
from rocquantum_b200 import workloads  # noqa: E402
from rocquantum_b200.distributed import DistStateVector  # noqa: E402
from tests import util  # noqa: E402


def gather_full(d, slice_):
    t = torch.from_numpy(slice_.view(np.float32 if d.prec == "c64" else np.float64)).cuda()
    outs = [torch.empty_like(t) for _ in range(d.world)]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy() for o in outs]).view(d.dtype)


def run_gates_per_gate(d, gates):
    for g in gates:
        name, targets, controls, theta = g[0], list(g[1]), list(g[2]), g[3]
        if name in ("h", "x", "y", "z", "s", "sdg", "t"):
            d.gate(name, targets[0])
        elif name in ("rx", "ry", "rz"):
            d.gate(name, targets[0], theta)
        elif name == "cnot":
            d.gate("cnot", controls[0], targets[0])
        elif name in ("cz", "swap"):
            d.gate(name, targets[0], targets[1])
        elif name in ("crx", "cry", "crz"):
            d.gate(name, controls[0], targets[0], theta)
        elif name == "mcx":
            d.gate("mcx", controls, targets[0])
        elif name == "cswap":
            d.gate("cswap", controls[0], targets[0], targets[1])
        else:
            raise ValueError(name)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
    fails = []
    for prec, tol in (("c64", 1e-5), ("c128", 1e-12)):
        for n in (10, 17):
            named = util.random_gates(n, 150, seed=n + world, allow_matrix=False)
            mixed = util.random_gates(n, 150, seed=3 * n + world, maxk=3) + workloads.c4_global_layers(n, 8, seed=36, top=3)
            # (a) one rocsvApply* call per gate, eager: every global target triggers an exchange
            o = so.Oracle(n, prec); util.run_on_oracle(o, named)
            d = DistStateVector(n, prec); run_gates_per_gate(d, named)
            err = util.rel_err(gather_full(d, d.local_slice()), o.state)
            if err > tol * 10: fails.append(("per-gate", prec, n, err))
            d.close()
            # (b) the same through the deferred queue
            d = DistStateVector(n, prec); d.set_fusion(True); run_gates_per_gate(d, named)
            err = util.rel_err(gather_full(d, d.local_slice()), o.state)
            if err > tol * 10: fails.append(("fusion", prec, n, err))
            d.close()
            # (c) whole circuit: Belady eviction, all rank bits traded at once
            o = so.Oracle(n, prec, seed=5); util.run_on_oracle(o, mixed)
            d = DistStateVector(n, prec, seed=5); d.apply_circuit(mixed)
            for ps, qs in [("Z", [n - 1]), ("X", [n - 1]), ("ZZ", [0, n - 1]), ("XY", [n - 2, n - 1]), ("YZX", [1, n - 1, 4])]:
                a, b = d.expect_pauli(ps, qs), o.expect_pauli(ps, qs)
                if abs(a - b) > tol * 10: fails.append(("expect", prec, n, ps, a, b))
            st = d.stats()
            full = gather_full(d, d.local_slice())
            err = util.rel_err(full, o.state)
            if err > tol * 10: fails.append(("circuit", prec, n, err))
            # (d) sampling / measurement on identical amplitudes: bit-exact
            nl = d.n_local
            d.init(); d.set_local_slice(o.state[rank << nl:(rank + 1) << nl])
            qs = [n - 1, 0, 3, n - 2]
            s1, s2 = d.sample(qs, 2000), o.sample(qs, 2000)
            if not np.array_equal(s1, s2): fails.append(("sample", prec, n, int((s1 != s2).sum())))
            for q in (n - 1, 2):
                m1, m2 = d.measure(q), o.measure(q)
                if m1 != m2: fails.append(("measure", prec, n, q, m1, m2))
            err = util.rel_err(gather_full(d, d.local_slice()), o.state)
            if err > tol: fails.append(("collapse", prec, n, err))
            # (e) rocsvSwapIndexBits incl. global<->global
            d.init(); d.set_local_slice(o.state[rank << nl:(rank + 1) << nl])
            pairs = [(0, n - 1), (n - 2, 3), (n - 1, n - 2)] if world >= 4 else [(0, n - 1), (n - 1, 5)]
            for a, b in pairs:
                d.swap_index_bits(a, b); o.swap_index_bits(a, b)
            if not np.array_equal(gather_full(d, d.local_slice()), o.state): fails.append(("swapbits", prec, n))
            if rank == 0:
                print(f"{prec} n={n} world={world}: circuit sweeps={st.sweeps} launches={st.kernelLaunches} ok so far, fails={len(fails)}", flush=True)
            d.close()
    # (f) tensor-core blocks on the local qubits of each slice (complex64), exchanges in between
    n = 22
    gates = workloads.c4_global_layers(n, 10, seed=36, top=3) + workloads.c2_random_unitary(n, 4, seed=30)
    o = so.Oracle(n, "c64"); util.run_on_oracle(o, gates)
    d = DistStateVector(n, "c64")
    assert d.lib.rocsvxSetTensorCoreBlocks(d.h, 1) == 0
    d.apply_circuit(gates)
    st = d.stats()
    err = util.rel_err(gather_full(d, d.local_slice()), o.state)
    if err > 1e-5 or st.blockSweeps == 0 or (world > 1 and st.exchanges == 0): fails.append(("blocks", n, err, int(st.blockSweeps), int(st.exchanges)))
    if rank == 0:
        print(f"c64 n={n} world={world}: blocks={st.blockSweeps} sweeps={st.sweeps} exchanges={st.exchanges} err={err:.2e}", flush=True)
    d.close()
    t = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("DIST CHECK", "PASS" if int(t) == 0 else f"FAIL {fails}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 0 else 1)


if __name__ == "__main__":
    main()
