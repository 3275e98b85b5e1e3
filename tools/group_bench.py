#!/usr/bin/env python
"""tools/group_bench.py -- the single-process multi-GPU mode (ONE handle, all visible devices) on the configs[3]-class
circuit: device time per step (slowest rank), exchanges, and a parity check against a complex128 single-GPU run at a
smaller size.  One process, so an ncu launch list of a distributed run can be taken from it:
    python tools/group_bench.py [--qubits 34] [--ranks 0] [--steps 2] [--parity-qubits 27]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocquantum_b200 import capi, workloads  # noqa: E402
from rocquantum_b200.statevec import StateVector  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=0)
    ap.add_argument("--ranks", type=int, default=0)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--parity-qubits", type=int, default=27)
    a = ap.parse_args()
    probe = StateVector(8, "c64", ranks=a.ranks)
    P = probe.dist_info()[1]
    probe.close()
    n = a.qubits or 33 + (P.bit_length() - 1)
    if a.parity_qubits:
        npq = a.parity_qubits
        gates = workloads.c4_global_layers(npq, 20, seed=36)
        g = StateVector(npq, "c64", ranks=a.ranks); g.set_tensor_core_blocks(True); g.apply_circuit(gates)
        st = g.stats()
        got = g.state(); g.close()
        r = StateVector(npq, "c128"); r.apply_circuit(gates); want = r.state(); r.close()
        err = float(np.abs(got - want).max() / np.abs(want).max())
        print(json.dumps(dict(what="single-process group vs complex128 single GPU", ranks=P, qubits=npq, gates=len(gates), max_rel_err=err,
                              ok=bool(err < 1e-5), exchanges=int(st.exchanges), block_sweeps=int(st.blockSweeps))), flush=True)
        del got, want
    gates = workloads.c4_global_layers(n, 20, seed=36)
    arr, keep = capi.make_ops(gates)
    g = StateVector(n, "c64", ranks=a.ranks)
    for _ in range(2):
        g.init(); assert g.lib.rocsvxApplyCircuit(g.h, None, n, arr, len(gates)) == 0
    g.sync(); g.stats(reset=True)
    g.timer_start(); t0 = time.perf_counter()
    host = []
    for _ in range(a.steps):
        ta = time.perf_counter(); g.init(); tb = time.perf_counter()
        assert g.lib.rocsvxApplyCircuit(g.h, None, n, arr, len(gates)) == 0
        host.append((round((tb - ta) * 1e3, 1), round((time.perf_counter() - tb) * 1e3, 1)))
    tc = time.perf_counter()
    ms = g.timer_stop() / a.steps
    wall = (time.perf_counter() - t0) / a.steps * 1e3
    print(json.dumps(dict(host_ms_per_call=host, what="(init, ApplyCircuit) host time of every step: the calls return when the launches are queued",
                          timer_stop_wait_ms=round((time.perf_counter() - tc) * 1e3, 1))), flush=True)
    st = g.stats()
    nl = g.dist_info()[2]
    sweeps = st.sweeps / a.steps
    ex_ms = st.exchangeMs / a.steps
    print(json.dumps(dict(what="single-process group, configs[3]-class circuit", ranks=P, qubits=n, local_qubits=nl, gates=len(gates),
                          device_ms_per_step=ms, wall_ms_per_step=wall, gates_per_s_30q_equiv=len(gates) / (ms * 1e-3) * 2.0 ** (n - 30),
                          sweeps_per_step=sweeps, block_sweeps_per_step=st.blockSweeps / a.steps, exchanges_per_step=st.exchanges / a.steps,
                          exchange_ms_per_step=ex_ms, exchange_GBps_per_rank=(st.exchangeBytes / 1e9) / (st.exchangeMs * 1e-3) if st.exchangeMs else None,
                          GBps_per_sweep=2.0 * (1 << nl) * 8 / ((ms - ex_ms) / max(1, sweeps) * 1e-3) / 1e9, norm=g.norm2())), flush=True)
    g.close()


if __name__ == "__main__":
    main()
