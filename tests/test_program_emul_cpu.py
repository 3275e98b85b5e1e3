"""Host pipeline down to the bits of the sweep program, without a GPU.

tests/hostemu/program_emul.cpp interprets the programs that host_ops.h emits for the tile-sweep kernel, loop for loop
like tile_sweep.cuh (group enumeration, register-window phases, swizzle, RQ_OP_DIAGP's per-thread / per-iteration split).
Gate list -> convert -> fuse -> merge controlled phases -> sweeps -> programs -> interpreted state must equal the oracle
running the gate list.  Test infrastructure only: the interpreter is compiled here from the product's headers."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import sv_oracle as so
from rocquantum_b200 import capi, workloads
from tests import util

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostemu", "program_emul.cpp")
OUT = os.path.join(HERE, "hostemu", "_build")
_LIBS = {}


def emu(prec):
    if prec not in _LIBS:
        os.makedirs(OUT, exist_ok=True)
        so_path = os.path.join(OUT, f"libprogram_emul_{prec}.so")
        deps = [SRC] + [os.path.join(HERE, "..", "rocquantum_b200", "csrc", f) for f in ("host_ops.h", "gate_convert.h", "sv_internal.h")]
        if not os.path.exists(so_path) or any(os.path.getmtime(d) > os.path.getmtime(so_path) for d in deps):
            cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so_path, SRC]
            if prec == "c128":
                cmd.insert(1, "-DROCQ_PRECISION_DOUBLE")
            subprocess.check_call(cmd)
        lib = C.CDLL(so_path)
        lib.hostemu_run_circuit.argtypes = [C.c_uint, C.c_uint, C.POINTER(capi.GateOp), C.c_size_t, C.c_void_p, C.c_uint, C.c_uint,
                                            C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        lib.hostemu_run_circuit.restype = C.c_int
        assert lib.hostemu_precision_bytes() == (4 if prec == "c64" else 8)
        _LIBS[prec] = lib
    return _LIBS[prec]


def run_emu(prec, n, gates, state, tile_bits=0, rank_bits=0, rank=0, flags=0):
    arr, keep = capi.make_ops(gates)
    v = np.ascontiguousarray(state, dtype=np.complex128).copy()
    nsw, nmerged = C.c_uint(), C.c_uint()
    st = emu(prec).hostemu_run_circuit(n, tile_bits, arr, len(gates), v.ctypes.data, rank_bits, rank, flags, C.byref(nsw), C.byref(nmerged))
    assert st == 0, st
    return v, nsw.value, nmerged.value


def oracle_run(n, gates, state):
    o = so.Oracle(n, "c128")
    o.set_state(state)
    util.run_on_oracle(o, gates)
    return o.state


def cp(j, i, theta):
    return ("matrix", [j], [i], 0.0, np.array([[1, 0], [0, np.exp(1j * theta)]]))


def diag_heavy_gates(n, count, seed, named_only=False):
    """Mostly controlled phases / CRZ / CZ / P fanned out from a few hub qubits, broken up by H, X, CNOT, SWAP and dense 2q."""
    rng = np.random.default_rng(seed)
    g = []
    while len(g) < count:
        r = rng.random()
        q = [int(x) for x in rng.permutation(n)]
        th = float(rng.uniform(0, 2 * math.pi))
        if r < 0.55:
            hub = q[0]
            for t in q[1:1 + int(rng.integers(1, n))]:
                kind = int(rng.integers(5))
                if kind == 0:
                    g.append(("cz", [hub, t], [], 0.0))
                elif kind == 1:
                    g.append(("crz", [t], [hub], float(rng.uniform(0, 2 * math.pi))))
                elif kind == 2:
                    g.append(("t", [hub], [], 0.0))
                elif named_only:
                    g.append(("crz", [hub], [t], float(rng.uniform(0, 2 * math.pi))))
                else:
                    g.append(cp(t, hub, float(rng.uniform(0, 2 * math.pi))) if rng.integers(2) else cp(hub, t, float(rng.uniform(0, 2 * math.pi))))
        elif r < 0.62:
            g.append(("x", [q[0]], [], 0.0))                   # carried forward through the diagonals (push_x_forward)
        elif r < 0.64:
            k = min(3, n)
            g.append(("matrix", q[:k], q[k:k + 1], 0.0, np.diag(np.exp(1j * rng.uniform(0, 2 * math.pi, size=1 << k)))))
        elif r < 0.7:
            g.append(("h", [q[0]], [], 0.0))
        elif r < 0.8:
            g.append(("cnot", [q[1]], [q[0]], 0.0))
        elif r < 0.85:
            g.append(("swap", [q[0], q[1]], [], 0.0))
        elif r < 0.9:
            g.append(("rz", [q[0]], [], th))
        elif named_only:
            g.append(("ry", [q[0]], [], th))
        else:
            g.append(("matrix", [q[0], q[1]], [], 0.0, workloads.haar_unitary(rng, 4)))
    return g[:count]


@pytest.mark.parametrize("prec,tol", [("c128", 1e-12), ("c64", 2e-6)])
@pytest.mark.parametrize("n,tile_bits", [(5, 0), (9, 6), (12, 8), (14, 0)])
def test_qft_programs(prec, tol, n, tile_bits):
    gates = workloads.c3_qft(n, seed=33)
    v = util.random_state(n, seed=n)
    out, nsw, nmerged = run_emu(prec, n, gates, v, tile_bits)
    assert util.rel_err(out, oracle_run(n, gates, v)) < tol
    assert nmerged >= n - 4                                   # one merged ladder per H(i) that keeps >= 3 controlled phases
    plain, nsw0, nm0 = run_emu(prec, n, gates, v, tile_bits, flags=1)
    assert nm0 == 0 and util.rel_err(plain, oracle_run(n, gates, v)) < tol


@pytest.mark.parametrize("prec,tol", [("c128", 1e-12), ("c64", 2e-6)])
@pytest.mark.parametrize("n,tile_bits,seed", [(4, 0, 1), (8, 5, 2), (11, 7, 3), (13, 0, 4), (14, 9, 5)])
def test_diagonal_heavy_programs(prec, tol, n, tile_bits, seed):
    gates = diag_heavy_gates(n, 160, seed)
    v = util.random_state(n, seed=seed)
    out, nsw, nmerged = run_emu(prec, n, gates, v, tile_bits)
    assert util.rel_err(out, oracle_run(n, gates, v)) < tol
    assert nmerged >= 1


@pytest.mark.parametrize("prec,tol", [("c128", 1e-12), ("c64", 2e-6)])
@pytest.mark.parametrize("n,tile_bits", [(3, 0), (7, 5), (10, 6), (13, 0), (14, 8)])
def test_mixed_bag_programs(prec, tol, n, tile_bits):
    """Every gate family of the ABI (k <= 3 matrices, controls, MCX, CSWAP): exercises window phases, swizzle, gcmask."""
    gates = util.random_gates(n, 220, seed=n, maxk=min(3, n))
    v = util.random_state(n, seed=n + 7)
    out, nsw, _ = run_emu(prec, n, gates, v, tile_bits)
    assert util.rel_err(out, oracle_run(n, gates, v)) < tol
    gates = workloads.c2_random_unitary(n, 6, seed=30) if n >= 2 else gates
    out, nsw, _ = run_emu(prec, n, gates, v, tile_bits)
    assert util.rel_err(out, oracle_run(n, gates, v)) < tol


def test_rank_slices_with_diagonals_on_rank_bits():
    """A distributed slice: controls and merged diagonal factors on rank bits come from hdr.high_base."""
    n, rb = 12, 2
    nl = n - rb
    rng = np.random.default_rng(9)
    gates = []
    for i in range(nl):
        gates.append(("h", [i], [], 0.0))
        for j in range(i + 1, n):                               # ladders that reach into the rank bits
            gates.append(cp(j, i, math.pi / (1 << (j - i))))
    for hub in range(nl, n):                                    # a rank bit as hub: the op exists on half of the ranks
        for t in range(0, nl, 2):
            gates.append(cp(t, hub, float(rng.uniform(0, 6))))
        gates.append(("cnot", [int(rng.integers(nl))], [hub], 0.0))
    v = util.random_state(n, seed=3)
    want = oracle_run(n, gates, v)
    for rank in range(1 << rb):
        sl = v[rank << nl:(rank + 1) << nl]
        out, nsw, nmerged = run_emu("c128", n, gates, sl, 7, rank_bits=rb, rank=rank)
        assert nmerged >= nl - 4
        assert util.rel_err(out, want[rank << nl:(rank + 1) << nl]) < 1e-12
        # the engine's path: rank bits resolved to constants before fusion (specialize_for_rank) -> purely local ops
        out, nsw2, _ = run_emu("c128", n, gates, sl, 7, rank_bits=rb, rank=rank, flags=2)
        assert util.rel_err(out, want[rank << nl:(rank + 1) << nl]) < 1e-12


@pytest.mark.parametrize("prec,tol", [("c128", 1e-12), ("c64", 2e-6)])
@pytest.mark.parametrize("n,rb,seed", [(9, 1, 1), (12, 2, 2), (13, 3, 3)])
def test_rank_specialised_mixed_bags(prec, tol, n, rb, seed):
    """Every gate family with controls / diagonal targets reaching into the rank bits, resolved per rank."""
    nl = n - rb
    rng = np.random.default_rng(seed)
    gates = []
    for g in util.random_gates(n, 260, seed=seed, maxk=3) + diag_heavy_gates(n, 120, seed):
        name, targets, controls = g[0], list(g[1]), list(g[2])
        diagonal = name in ("z", "s", "sdg", "t", "rz", "cz", "crz") or (name == "matrix" and np.count_nonzero(g[4] - np.diag(np.diagonal(g[4]))) == 0)
        if name == "cz":
            diagonal = True
        if not diagonal and any(t >= nl for t in targets):
            continue                                            # non-diagonal targets stay local (the planner exchanges first)
        gates.append(g)
    v = util.random_state(n, seed=seed)
    want = oracle_run(n, gates, v)
    for rank in range(1 << rb):
        sl = v[rank << nl:(rank + 1) << nl]
        out, _, _ = run_emu(prec, n, gates, sl, 6, rank_bits=rb, rank=rank, flags=2)
        assert util.rel_err(out, want[rank << nl:(rank + 1) << nl]) < tol
        out, _, _ = run_emu(prec, n, gates, sl, 6, rank_bits=rb, rank=rank)
        assert util.rel_err(out, want[rank << nl:(rank + 1) << nl]) < tol


@pytest.mark.parametrize("rb", [0, 1, 2])
def test_mixed_plan_with_blocks_on_rank_slices(rb):
    """The complex64 engine's plan on a slice: rank bits resolved, ops folded into 6-qubit blocks (applied here as plain
    64x64 products) and ordinary sweeps.  Blocks must hold local ops only and the result must equal the oracle."""
    n = 14 + rb
    nl = n - rb
    gates = [g for g in workloads.c4_global_layers(n, 5, seed=36, top=3) + workloads.c2_random_unitary(n, 3, seed=30)
             + diag_heavy_gates(n, 60, 7) if g[0] in ("cz", "crz", "t", "rz") or (g[0] == "matrix" and len(g[1]) == 1 and len(g[2]) == 1)
             or all(t < nl for t in g[1])]
    v = util.random_state(n, seed=rb)
    want = oracle_run(n, gates, v)
    for rank in range(1 << rb):
        sl = v[rank << nl:(rank + 1) << nl]
        out, launches, nblocks = run_emu("c64", n, gates, sl, 0, rank_bits=rb, rank=rank, flags=2 | 4)
        assert nblocks >= 3
        assert util.rel_err(out, want[rank << nl:(rank + 1) << nl]) < 2e-6


def test_shot_bit_gather_matches_the_per_bit_definition():
    """rocsvSample's index -> result-word mapping (rq::BitGather, host_ops.h): bit j = bit measured[j] of the index
    (hipStateVec.h:427-445), whatever runs the measured list happens to contain."""
    lib = emu("c64")
    lib.hostemu_gather_bits.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint), C.c_uint, C.c_void_p]
    lib.hostemu_gather_bits.restype = None
    rng = np.random.default_rng(7)
    idx = rng.integers(0, 1 << 63, size=4096, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=4096, dtype=np.uint64)
    idx[:4] = [0, 2**64 - 1, 1, 1 << 63]
    cases = [[], [0], [63], list(range(28)), list(range(64)), list(range(63, -1, -1)), [5, 6, 7, 2, 3, 40, 41, 42, 43, 0],
             [3, 3, 4, 4], [1, 0, 1, 2, 3], list(range(10, 40)) + list(range(0, 10)), [62, 63, 0, 1]]
    for _ in range(40):
        nm = int(rng.integers(1, 65))
        cases.append([int(q) for q in rng.integers(0, 64, size=nm)])                    # arbitrary, duplicates allowed
        start = int(rng.integers(0, 64 - nm + 1))
        perm = list(range(start, start + nm))
        cut = int(rng.integers(0, nm))
        cases.append(perm[cut:] + perm[:cut])                                            # two long runs
    for measured in cases:
        out = np.zeros(idx.size, dtype=np.uint64)
        lib.hostemu_gather_bits(idx.ctypes.data, idx.size, capi.uarr(measured) if measured else None, len(measured), out.ctypes.data)
        want = np.zeros(idx.size, dtype=np.uint64)
        for j, q in enumerate(measured):
            want |= ((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(j)
        assert np.array_equal(out, want), measured


@pytest.mark.parametrize("n,depth", [(13, 8), (15, 10)])
def test_mixed_plan_of_a_brick_circuit_through_the_interpreter(n, depth):
    """configs[1] in miniature under the two-policy block planner: blocks (as 64x64 products) and whatever ordinary sweeps
    remain, interpreted from the emitted programs, equal the oracle."""
    gates = workloads.c2_random_unitary(n, depth, seed=n)
    v = util.random_state(n, seed=n)
    want = oracle_run(n, gates, v)
    out, launches, nblocks = run_emu("c64", n, gates, v, 0, flags=4)
    assert nblocks >= 2
    assert util.rel_err(out, want) < 2e-6


@pytest.mark.parametrize("n,tile_bits,expect", [(14, 8, 3), (18, 9, 5), (17, 7, 6)])
def test_backward_planned_qft_through_the_interpreter(n, tile_bits, expect):
    """QFTs for which plan_sweeps keeps the backward-read plan (swaps, Hadamards and merged ladders in the same sweeps):
    the emitted complex128 programs, interpreted, equal the oracle."""
    gates = workloads.c3_qft(n, seed=33)
    v = util.random_state(n, seed=n)
    want = oracle_run(n, gates, v)
    out, nsw, nmerged = run_emu("c128", n, gates, v, tile_bits=tile_bits)
    assert nsw == expect and nmerged >= n - 4
    assert util.rel_err(out, want) < 1e-12


def test_random_circuits_through_both_planners():
    """A seeded mix of circuit families, sizes and tile widths through the interpreter: ordinary sweeps (forward or
    backward partition, whichever the planner kept) in both precisions, and the mixed block plan in complex64."""
    rng = np.random.default_rng(2024)
    for trial in range(16):
        n = int(rng.integers(8, 17)); tb = int(rng.integers(6, min(n, 10) + 1)); seed = int(rng.integers(1 << 30))
        g = [lambda: util.random_gates(n, 150, seed=seed, maxk=min(4, n)),
             lambda: diag_heavy_gates(n, 120, seed) + workloads.c3_qft(n, seed=seed),
             lambda: workloads.c3_qft(n, seed=seed) + util.random_gates(n, 60, seed=seed, maxk=2),
             lambda: workloads.c1_ghz_random_layers(n, 6, seed=seed) + workloads.c3_qft(n, seed=seed)[::-1]][trial % 4]()
        v = util.random_state(n, seed=seed % 1000)
        want = oracle_run(n, g, v)
        for prec, tol in (("c128", 1e-11), ("c64", 2e-5)):
            out, nsw, nm = run_emu(prec, n, g, v, tile_bits=tb)
            assert util.rel_err(out, want) < tol, (trial, n, tb, seed, prec)
    for trial in range(8):
        n = int(rng.integers(13, 18)); seed = int(rng.integers(1 << 30)); depth = int(rng.integers(2, 9))
        g = [lambda: workloads.c2_random_unitary(n, depth, seed=seed),
             lambda: workloads.c2_random_unitary(n, depth, seed=seed) + util.random_gates(n, 60, seed=seed, maxk=3),
             lambda: workloads.c4_global_layers(n, depth, seed=seed, top=3) + workloads.c2_random_unitary(n, 3, seed=seed),
             lambda: util.random_gates(n, 80, seed=seed, maxk=2) + workloads.c2_random_unitary(n, depth, seed=seed)[::-1]][trial % 4]()
        v = util.random_state(n, seed=seed % 1000)
        out, nsw, nb = run_emu("c64", n, g, v, 0, flags=4)
        assert util.rel_err(out, oracle_run(n, g, v)) < 2e-5, (trial, n, seed)


def take_counters(prec):
    a, b = C.c_uint(), C.c_uint()
    emu(prec).hostemu_take_counters(C.byref(a), C.byref(b))
    return a.value, b.value


@pytest.mark.parametrize("prec,tol", [("c128", 1e-12), ("c64", 2e-6)])
def test_trailing_swaps_ride_in_the_store_addresses(prec, tol):
    """Swaps of resident qubits that nothing later in the sweep depends on (a QFT's bit reversal) leave the program and become
    the store map sres[]; swaps below the minimal row, controlled swaps and swaps something still depends on stay passes."""
    rng = np.random.default_rng(8)
    # QFTs whose sweeps end in their swaps
    for n, tb in ((14, 0), (16, 10), (18, 9), (17, 12)):
        gates = workloads.c3_qft(n, seed=n)
        v = util.random_state(n, seed=n)
        take_counters(prec)
        out, nsw, nm = run_emu(prec, n, gates, v, tile_bits=tb)
        remaps, perms = take_counters(prec)
        assert util.rel_err(out, oracle_run(n, gates, v)) < tol, (n, tb)
        assert remaps >= 1, (n, tb)
        assert perms < n // 2, (n, tb, perms)                 # some of the n/2 swaps are gone from the programs
    # chains of swaps sharing qubits, swaps followed by gates on their qubits (must stay), a controlled swap, low-bit swaps
    for trial in range(12):
        n = int(rng.integers(10, 15))
        g = util.random_gates(n, 40, seed=100 + trial, maxk=2)
        q = [int(x) for x in rng.permutation(n)]
        g += [("swap", [q[0], q[1]], [], 0.0), ("swap", [q[1], q[2]], [], 0.0), ("swap", [q[3], q[0]], [], 0.0)]
        if trial % 3 == 0:
            g += [("h", [q[1]], [], 0.0)]                     # depends on the swaps before it
        if trial % 3 == 1:
            g += [("cswap", [q[4], q[5]], [q[6]], 0.0), ("swap", [q[7], q[8]], [], 0.0)]
        if trial % 4 == 2:
            g += [("swap", [0, 1], [], 0.0), ("swap", [2, q[2] if q[2] != 2 else 3], [], 0.0)]
        v = util.random_state(n, seed=trial)
        out, nsw, nm = run_emu(prec, n, g, v, tile_bits=int(rng.integers(8, 11)))
        assert util.rel_err(out, oracle_run(n, g, v)) < tol * 10, trial
