"""Test helpers: run gate-tuple circuits (rocquantum_b200.workloads format) on the oracle / the compiled
reference, and re-simulate a sweep plan dumped by rocsvxPlanCircuit.  Test infrastructure only."""
import ctypes as C

import numpy as np

from oracle import sv_oracle as so
from rocquantum_b200 import capi

TOL = {"c64": 1e-5, "c128": 1e-12}     # north_star: relative amplitude tolerance


def run_on_oracle(o: so.Oracle, gates):
    for g in gates:
        name, targets, controls, theta = g[0], list(g[1]), list(g[2]), g[3]
        if name == "matrix":
            if len(targets) == 1 and not controls:
                o.matrix1(targets[0], g[4])          # the reference's own 1q kernel arithmetic (single_qubit_kernels.hip:28-72)
            else:
                o.apply_matrix(targets, g[4], controls)
        elif name in ("h", "x", "y", "z", "s", "sdg", "t"):
            o.gate(name, targets[0])
        elif name in ("rx", "ry", "rz"):
            o.gate(name, targets[0], theta)
        elif name == "cnot":
            o.gate("cnot", controls[0], targets[0])
        elif name in ("cz", "swap"):
            o.gate(name, targets[0], targets[1])
        elif name in ("crx", "cry", "crz"):
            o.gate(name, controls[0], targets[0], theta)
        elif name == "mcx":
            o.gate("mcx", controls, targets[0])
        elif name == "cswap":
            o.gate("cswap", controls[0], targets[0], targets[1])
        else:
            raise ValueError(name)


def run_on_ref(r: so.RefLib, gates):
    """Only gates the reference defines (no 'matrix')."""
    for g in gates:
        name, targets, controls, theta = g[0], list(g[1]), list(g[2]), g[3]
        if name in ("h", "x", "y", "z", "s", "sdg", "t"):
            st = r.gate(name, targets[0])
        elif name in ("rx", "ry", "rz"):
            st = r.gate(name, targets[0], theta)
        elif name == "cnot":
            st = r.gate("cnot", controls[0], targets[0])
        elif name in ("cz", "swap"):
            st = r.gate(name, targets[0], targets[1])
        elif name in ("crx", "cry", "crz"):
            st = r.gate(name, controls[0], targets[0], theta)
        elif name == "mcx":
            st = r.gate("mcx", controls, targets[0])
        elif name == "cswap":
            st = r.gate("cswap", controls[0], targets[0], targets[1])
        else:
            raise ValueError(name)
        assert st == 0


def run_per_gate(sv, gates):
    """Drive the C ABI one rocsvApply* call per gate, like Circuit.flush() in the reference."""
    for g in gates:
        name, targets, controls, theta = g[0], list(g[1]), list(g[2]), g[3]
        if name == "matrix":
            sv.apply_matrix(targets, g[4], controls)
        elif name in ("h", "x", "y", "z", "s", "sdg", "t"):
            sv.gate(name, targets[0])
        elif name in ("rx", "ry", "rz"):
            sv.gate(name, targets[0], theta)
        elif name == "cnot":
            sv.gate("cnot", controls[0], targets[0])
        elif name in ("cz", "swap"):
            sv.gate(name, targets[0], targets[1])
        elif name in ("crx", "cry", "crz"):
            sv.gate(name, controls[0], targets[0], theta)
        elif name == "mcx":
            sv.gate("mcx", controls, targets[0])
        elif name == "cswap":
            sv.gate("cswap", controls[0], targets[0], targets[1])
        else:
            raise ValueError(name)


def plan_blocks(n, gates, min_cost=0.0):
    """-> (num_blocks, num_sweeps, steps) via rocsvxPlanCircuitBlocks (host only); a step is a dict with 'blk' (6 positions) for a
    tensor-core block or 'res' for an ordinary sweep, and its ops in execution order."""
    lib = capi.load("c64")
    arr, keep = capi.make_ops(gates)
    nb, ns = C.c_uint(), C.c_uint()
    size = 1 << 18
    while True:
        buf = C.create_string_buffer(size)
        st = lib.rocsvxPlanCircuitBlocks(n, arr, len(gates), C.c_double(min_cost), C.byref(nb), C.byref(ns), buf, size)
        assert st == 0, st
        txt = buf.value.decode()
        if len(txt) < size - 2:
            break
        size *= 4
    steps = []
    for line in txt.splitlines():
        tok = line.split()
        if tok[0] == "B":
            steps.append(dict(blk=[int(x) for x in tok[1:]], ops=[]))
        elif tok[0] == "S":
            steps.append(dict(T=int(tok[1]), rowbits=int(tok[2]), res=[int(x) for x in tok[4:]], ops=[]))
        else:
            kind = int(tok[1])
            cmask = int(tok[3], 16)
            ti, di = tok.index("targets"), tok.index("data")
            targets = [int(x) for x in tok[ti + 1:di]]
            vals = [float(x) for x in tok[di + 1:]]
            data = np.array(vals[0::2]) + 1j * np.array(vals[1::2])
            steps[-1]["ops"].append(dict(kind=kind, cmask=cmask, targets=targets, data=data))
    assert sum(1 for s in steps if "blk" in s) == nb.value and sum(1 for s in steps if "res" in s) == ns.value
    return nb.value, ns.value, steps


def plan(n, gates, tile_bits=0, prec="c64"):
    """-> (num_sweeps, [ {T,rowbits,res,ops:[{kind,cmask,targets,data}]} ])  via rocsvxPlanCircuit (host only)."""
    lib = capi.load(prec)
    arr, keep = capi.make_ops(gates)
    nsw = C.c_uint()
    size = 1 << 16
    while True:
        buf = C.create_string_buffer(size)
        st = lib.rocsvxPlanCircuit(n, tile_bits, arr, len(gates), C.byref(nsw), buf, size)
        assert st == 0, st
        txt = buf.value.decode()
        if len(txt) < size - 2:
            break
        size *= 4
    sweeps = []
    for line in txt.splitlines():
        tok = line.split()
        if tok[0] == "S":
            sweeps.append(dict(T=int(tok[1]), rowbits=int(tok[2]), res=[int(x) for x in tok[4:]], ops=[]))
        else:
            kind = int(tok[1])
            cmask = int(tok[3], 16)
            ti, di = tok.index("targets"), tok.index("data")
            targets = [int(x) for x in tok[ti + 1:di]]
            vals = [float(x) for x in tok[di + 1:]]
            data = np.array(vals[0::2]) + 1j * np.array(vals[1::2])
            sweeps[-1]["ops"].append(dict(kind=kind, cmask=cmask, targets=targets, data=data))
    assert len(sweeps) == nsw.value
    return nsw.value, sweeps


def simulate_plan(o: so.Oracle, sweeps):
    """Apply a dumped plan to an oracle state, checking each op only uses what its sweep makes resident."""
    n = o.n
    for sw in sweeps:
        res = set(sw["res"])
        assert len(res) == sw["T"] == min(n, sw["T"])
        assert sw["res"][:sw["rowbits"]] == list(range(sw["rowbits"]))
        for op in sw["ops"]:
            controls = [q for q in range(64) if (op["cmask"] >> q) & 1]
            t, k = op["targets"], len(op["targets"])
            if op["kind"] == 1:      # DENSE
                assert set(t) <= res, "dense target not resident"
                M = op["data"].reshape(1 << k, 1 << k).T       # column-major -> M[i][j]
                o.apply_matrix(t, M, controls)
            elif op["kind"] == 2:    # DIAG (targets/controls may be non-resident)
                if k == 0:
                    idx = np.arange(1 << n)
                    m = np.ones(1 << n, dtype=bool)
                    for c in controls:
                        m &= ((idx >> c) & 1).astype(bool)
                    st = o.state.reshape(o.batch, 1 << n)
                    st[:, m] = (st[:, m] * op["data"][0]).astype(o.dtype)
                else:
                    o.apply_matrix(t, np.diag(op["data"]), controls)
            elif op["kind"] == 5:    # DIAGP: data[0] * prod of data[1 + b] over the set bits targets[b], where all controls are 1
                idx = np.arange(1 << n)
                m = np.ones(1 << n, dtype=bool)
                for c in controls:
                    m &= ((idx >> c) & 1).astype(bool)
                f = np.full(1 << n, op["data"][0], dtype=np.complex128)
                for b, q in enumerate(t):
                    f = np.where((idx >> q) & 1, f * op["data"][1 + b], f)
                st = o.state.reshape(o.batch, 1 << n)
                st[:, m] = (st[:, m] * f[m]).astype(o.dtype)
            elif op["kind"] == 3:    # PERM_X
                assert set(t) <= res
                if controls:
                    o.mcx(controls, t[0])
                else:
                    o.gate("x", t[0])
            else:                    # PERM_SWAP
                assert set(t) <= res
                assert len(controls) <= 1
                if controls:
                    o.gate("cswap", controls[0], t[0], t[1])
                else:
                    o.gate("swap", t[0], t[1])


def rel_err(a, b):
    a = np.asarray(a, dtype=np.complex128)
    b = np.asarray(b, dtype=np.complex128)
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))


def random_state(n, batch=1, seed=0):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(batch << n) + 1j * rng.standard_normal(batch << n)
    v = v.reshape(batch, 1 << n)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.reshape(-1)


def random_gates(n, count, seed, allow_matrix=True, maxk=3):
    """A mixed bag over every gate family of the ABI."""
    import math
    from rocquantum_b200.workloads import haar_unitary
    rng = np.random.default_rng(seed)
    names = ["h", "x", "y", "z", "s", "sdg", "t", "rx", "ry", "rz"]
    if n >= 2:
        names += ["cnot", "cz", "swap", "crx", "cry", "crz"]
    if n >= 3:
        names += ["mcx", "cswap"]
    if allow_matrix:
        names += ["matrix", "matrix"]
    out = []
    for _ in range(count):
        g = names[int(rng.integers(len(names)))]
        q = [int(x) for x in rng.permutation(n)]
        th = float(rng.uniform(0, 2 * math.pi))
        if g in ("h", "x", "y", "z", "s", "sdg", "t"):
            out.append((g, [q[0]], [], 0.0))
        elif g in ("rx", "ry", "rz"):
            out.append((g, [q[0]], [], th))
        elif g == "cnot":
            out.append((g, [q[1]], [q[0]], 0.0))
        elif g in ("cz", "swap"):
            out.append((g, [q[0], q[1]], [], 0.0))
        elif g in ("crx", "cry", "crz"):
            out.append((g, [q[1]], [q[0]], th))
        elif g == "mcx":
            nc = int(rng.integers(1, min(n - 1, 4) + 1))
            out.append((g, [q[nc]], q[:nc], 0.0))
        elif g == "cswap":
            out.append((g, [q[1], q[2]], [q[0]], 0.0))
        else:
            k = int(rng.integers(1, min(maxk, n) + 1))
            nc = int(rng.integers(0, min(2, n - k) + 1))
            kind = int(rng.integers(3))
            if kind == 0:
                M = haar_unitary(rng, 1 << k)
            elif kind == 1:
                M = np.diag(np.exp(1j * rng.uniform(0, 2 * math.pi, size=1 << k)))
            else:
                M = np.diag(np.concatenate([np.ones((1 << k) - 1), [np.exp(1j * th)]]))
            out.append(("matrix", q[:k], q[k:k + nc], 0.0, M))
    return out


def dist_plan(n, nranks, gates, mode=0, canonicalize=True, prec="c64"):
    """-> (num_exchanges, steps, final_map): steps = [("R", [op dicts]) | ("X", [gpos])]  (rocsvxDistPlanCircuit, host only)"""
    lib = capi.load(prec)
    arr, keep = capi.make_ops(gates)
    nx = C.c_uint()
    size = 1 << 18
    while True:
        buf = C.create_string_buffer(size)
        st = lib.rocsvxDistPlanCircuit(n, nranks, arr, len(gates), mode, int(canonicalize), C.byref(nx), buf, size)
        assert st == 0, st
        txt = buf.value.decode()
        if len(txt) < size - 2:
            break
        size *= 4
    steps, fmap, block, nblocks = [], None, None, [0]
    for line in txt.splitlines():
        tok = line.split()
        if tok[0] == "R":
            block = None
            steps.append(("R", []))
        elif tok[0] == "S":
            assert all(int(x) < n - (nranks.bit_length() - 1) for x in tok[4:]), "rank bit made resident"
            block = None
        elif tok[0] == "B":                      # tensor-core block on six LOCAL positions; its ops follow
            block = [int(x) for x in tok[1:]]
            assert len(block) == 6 and all(x < n - (nranks.bit_length() - 1) for x in block), "block on a rank bit"
            nblocks[0] += 1
        elif tok[0] == "X":
            steps.append(("X", [int(x) for x in tok[1:]]))
        elif tok[0] == "M":
            fmap = [int(x) for x in tok[1:]]
        else:
            kind, cmask = int(tok[1]), int(tok[3], 16)
            ti, di = tok.index("targets"), tok.index("data")
            vals = [float(x) for x in tok[di + 1:]]
            op = dict(kind=kind, cmask=cmask, targets=[int(x) for x in tok[ti + 1:di]], data=np.array(vals[0::2]) + 1j * np.array(vals[1::2]))
            if block is not None:                # every op folded into a block lies inside it
                assert set(op["targets"]) | {q for q in range(64) if (cmask >> q) & 1} <= set(block)
            steps[-1][1].append(op)
    dist_plan.blocks = nblocks[0]
    return nx.value, steps, fmap


def simulate_dist_plan(o: so.Oracle, n_local, steps):
    """Re-simulate a distributed plan on ONE full state indexed by physical position (qubit = physical position)."""
    n = o.n
    for kind, payload in steps:
        if kind == "X":
            k = len(payload)
            for i, g in enumerate(payload):
                assert g >= n_local
                o.swap_index_bits(n_local - k + i, g)
        else:
            for op in payload:
                if op["kind"] not in (2, 5):
                    assert all(t < n_local for t in op["targets"]), "non-diagonal target on a rank bit"
            simulate_plan(o, [dict(T=n, rowbits=n, res=list(range(n)), ops=payload)])
