"""Parity tests proper (-m gpu): the CUDA engine, called through the C ABI, against the oracle on the same
seeded inputs, against the committed golden fixtures (outputs of the compiled reference), and -- at full size --
through size-independent properties.  Tolerances are north_star's: 1e-5 relative (complex64), 1e-12 (complex128);
sampled bitstrings and measurement outcomes must be bit-exact for the shared Philox stream."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from oracle import sv_oracle as so
from rocquantum_b200 import capi, workloads
from rocquantum_b200.statevec import StateVector
from tests import util

pytestmark = pytest.mark.gpu
TOL = util.TOL
PRECS = ["c64", "c128"]


def _pair(n, prec, batch=1, seed=1, fusion=False):
    v = util.random_state(n, batch, seed)
    o = so.Oracle(n, prec, batch=batch); o.set_state(v)
    g = StateVector(n, prec, batch=batch, fusion=fusion); g.set_state(v)
    return o, g


def test_native_library_is_loaded():
    lib = capi.load("c64")
    maps = open("/proc/self/maps").read()
    assert "libhipStateVec.so" in maps
    h = C.c_void_p()
    assert lib.rocsvCreate(C.byref(h)) == capi.SUCCESS and h.value
    assert lib.rocsvDestroy(h) == capi.SUCCESS


@pytest.mark.parametrize("prec", PRECS)
def test_golden_vectors_through_the_c_abi(prec):
    S, tol = 1 / math.sqrt(2), 1e-6
    g = StateVector(2, prec); g.gate("h", 0); g.gate("cnot", 0, 1)
    assert np.allclose(g.state(), [S, 0, 0, S], atol=tol)                      # Bell
    g = StateVector(1, prec); g.gate("rx", 0, math.pi / 2)
    assert np.allclose(g.state(), [math.cos(math.pi / 4), -1j * math.sin(math.pi / 4)], atol=tol)
    g = StateVector(1, prec); g.gate("ry", 0, math.pi / 2)
    assert np.allclose(g.state(), [math.cos(math.pi / 4), math.sin(math.pi / 4)], atol=tol)
    g = StateVector(1, prec); g.gate("h", 0); g.gate("rz", 0, math.pi / 2)
    assert np.allclose(g.state(), np.array([np.exp(-1j * math.pi / 4), np.exp(1j * math.pi / 4)]) * S, atol=tol)
    g = StateVector(3, prec); g.gate("x", 0)
    assert abs(g.state()[1] - 1) < tol
    g.gate("cnot", 0, 1)
    assert abs(g.state()[3] - 1) < tol
    g = StateVector(3, prec); g.apply_fused_1q(0, so.gate_matrix("h"))         # test_hipStateVec_multi_gpu.cpp:300-339
    st = g.state()
    assert np.allclose(st[:2], [S, S], atol=tol) and np.abs(st[2:]).max() < tol
    g = StateVector(3, prec); g.gate("x", 0); g.gate("x", 1); g.gate("mcx", [0, 1], 2)
    assert abs(g.state()[7] - 1) < tol
    g = StateVector(3, prec); g.gate("x", 0); g.gate("x", 1); g.gate("cswap", 0, 1, 2)
    assert abs(g.state()[5] - 1) < tol
    g = StateVector(3, prec); g.gate("h", 0); g.gate("cnot", 0, 1); g.gate("cnot", 1, 2)
    assert abs(g.expect_zprod([0, 1]) - 1) < 1e-6 and abs(g.expect_pauli("XY", [1, 2])) < 1e-6
    assert abs(g.expect_pauli("XYZ", [0, 1, 2])) < 1e-6


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name", ["c1_n10", "mixed_n10", "c2_n10"])
def test_compiled_reference_fixtures(prec, name):
    want = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}_{prec}.npy"))
    gates = {"c1_n10": lambda: workloads.c1_ghz_random_layers(10, 6, seed=20),
             "mixed_n10": lambda: util.random_gates(10, 200, seed=4242, allow_matrix=False),
             "c2_n10": lambda: workloads.c2_random_unitary(10, 6, seed=30)}[name]()      # c2: the reference's ApplyMatrix spec kernel
    for fusion in (False, True):
        g = StateVector(10, prec, fusion=fusion)
        util.run_per_gate(g, gates)
        assert util.rel_err(g.state(), want) < TOL[prec]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 13, 15])
def test_every_named_gate_on_every_target(prec, n):
    o, g = _pair(n, prec, seed=n)
    th = 0.3 + n
    for t in range(n):
        for name in ("h", "x", "y", "z", "s", "sdg", "t"):
            o.gate(name, t); g.gate(name, t)
        for name in ("rx", "ry", "rz"):
            o.gate(name, t, th); g.gate(name, t, th)
    if n >= 2:
        for a in range(n):
            b = (a * 5 + 1) % n
            if a == b:
                continue
            for name in ("cnot", "cz", "swap"):
                o.gate(name, a, b); g.gate(name, a, b)
            for name in ("crx", "cry", "crz"):
                o.gate(name, a, b, th); g.gate(name, a, b, th)
    if n >= 3:
        for a in range(n):
            b, c = (a + 1) % n, (a + n // 2 + 1) % n
            if len({a, b, c}) < 3:
                continue
            o.gate("mcx", [a, b], c); g.gate("mcx", [a, b], c)
            o.gate("cswap", a, b, c); g.gate("cswap", a, b, c)
    assert util.rel_err(g.state(), o.state) < TOL[prec]           # north_star's tolerance, no slack (~1e3 gates)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n,batch", [(4, 3), (9, 2), (14, 2), (16, 1), (18, 1)])
def test_mixed_bag_eager_and_fused(prec, n, batch):
    gates = util.random_gates(n, 120, seed=7 * n, maxk=min(4, n))
    v = util.random_state(n, batch, seed=n)
    o = so.Oracle(n, prec, batch=batch); o.set_state(v); util.run_on_oracle(o, gates)
    e = StateVector(n, prec, batch=batch); e.set_state(v); util.run_per_gate(e, gates)
    assert util.rel_err(e.state(), o.state) < TOL[prec]
    f = StateVector(n, prec, batch=batch); f.set_state(v); f.apply_circuit(gates)
    assert util.rel_err(f.state(), o.state) < TOL[prec]
    d = StateVector(n, prec, batch=batch, fusion=True); d.set_state(v); util.run_per_gate(d, gates)
    assert util.rel_err(d.state(), o.state) < TOL[prec]
    assert d.stats().sweeps < e.stats().sweeps


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10])            # 10 = the widest matrix the API accepts
def test_apply_matrix_k_qubits_with_controls(prec, k):
    n = 11 if k < 8 else 13
    rng = np.random.default_rng(k)
    o, g = _pair(n, prec, seed=k)
    for trial in range(3):
        q = [int(x) for x in rng.permutation(n)]
        nc = trial
        U = workloads.haar_unitary(rng, 1 << k)
        o.apply_matrix(q[:k], U, q[k:k + nc]); g.apply_matrix(q[:k], U, q[k:k + nc])
    assert util.rel_err(g.state(), o.state) < TOL[prec]


@pytest.mark.parametrize("prec", PRECS)
def test_c1_config_full_state(prec):
    """configs[0]: 20-qubit GHZ + 20 random layers, compared amplitude by amplitude."""
    n = 20
    gates = workloads.c1_ghz_random_layers(n, 20, seed=20)
    o = so.Oracle(n, prec); util.run_on_oracle(o, gates)
    g = StateVector(n, prec); g.apply_circuit(gates)
    st = g.state()
    assert util.rel_err(st, o.state) < TOL[prec]
    assert abs(g.norm2() - 1) < (1e-4 if prec == "c64" else 1e-11)
    assert g.stats().sweeps < len(gates) // 5


@pytest.mark.parametrize("prec", PRECS)
def test_c2_and_qft_configs_reduced(prec):
    n = 22
    for gates in (workloads.c2_random_unitary(n, 6, seed=30), workloads.c3_qft(n, seed=33)):
        o = so.Oracle(n, prec); util.run_on_oracle(o, gates)
        g = StateVector(n, prec); g.apply_circuit(gates)
        assert util.rel_err(g.state(), o.state) < TOL[prec]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n,batch,seed", [(4, 2, 1), (7, 1, 2), (9, 3, 3), (13, 1, 4), (14, 2, 5), (17, 1, 6), (20, 1, 7)])
def test_merged_controlled_phase_ladders(prec, n, batch, seed):
    """Runs of CP / CZ / CRZ / P sharing a control become one RQ_OP_DIAGP pass (host_ops.h: merge_diagonals); same result
    as the gate-by-gate oracle and as the engine with merging off, in fewer executed ops."""
    from tests.test_program_emul_cpu import diag_heavy_gates
    gates = diag_heavy_gates(n, 200, seed) + workloads.c3_qft(n, seed=seed)
    v = util.random_state(n, batch, seed=seed)
    o = so.Oracle(n, prec, batch=batch); o.set_state(v); util.run_on_oracle(o, gates)
    m = StateVector(n, prec, batch=batch); m.set_state(v); m.apply_circuit(gates)
    assert util.rel_err(m.state(), o.state) < TOL[prec]
    p = StateVector(n, prec, batch=batch); p.set_merge_diagonals(False); p.set_state(v); p.apply_circuit(gates)
    assert util.rel_err(p.state(), o.state) < TOL[prec]
    assert m.stats().opsExecuted < p.stats().opsExecuted
    # deferred rocsvApply* calls (named gates only: a device-matrix call flushes the queue)
    named = diag_heavy_gates(n, 200, seed, named_only=True)
    o = so.Oracle(n, prec, batch=batch); o.set_state(v); util.run_on_oracle(o, named)
    d = StateVector(n, prec, batch=batch, fusion=True); d.set_state(v); util.run_per_gate(d, named)
    q = StateVector(n, prec, batch=batch, fusion=True); q.set_merge_diagonals(False); q.set_state(v); util.run_per_gate(q, named)
    assert util.rel_err(d.state(), o.state) < TOL[prec] and util.rel_err(q.state(), o.state) < TOL[prec]
    assert d.stats().opsExecuted < q.stats().opsExecuted


def test_qft_analytic():
    """QFT of a basis state |x> has amplitudes exp(2 pi i x k / 2^n)/sqrt(2^n) (bit-reversed by the final swaps)."""
    n = 16
    rng = np.random.default_rng(33)
    xbits = [int(rng.integers(2)) for _ in range(n)]
    gates = workloads.c3_qft(n, seed=33)
    g = StateVector(n, "c128"); g.apply_circuit(gates)
    x = sum(b << q for q, b in enumerate(xbits))
    # run_benchmark.py's QFT treats qubit 0 as the most significant input bit
    xr = int(format(x, f"0{n}b")[::-1], 2)
    k = np.arange(1 << n)
    kr = np.zeros_like(k)
    for b in range(n):                       # output index is bit-reversed as well
        kr |= ((k >> b) & 1) << (n - 1 - b)
    want = np.exp(2j * math.pi * xr * kr / (1 << n)) / math.sqrt(1 << n)
    assert np.abs(g.state() - want).max() < 1e-10


@pytest.mark.parametrize("prec", PRECS)
def test_expectations(prec):
    n = 12
    o, g = _pair(n, prec, seed=12)
    gates = workloads.c5_vqe_ansatz(n, seed=5)
    util.run_on_oracle(o, gates); util.run_per_gate(g, gates)
    tol = 1e-5 if prec == "c64" else 1e-12
    for q in (0, 5, 11):
        assert abs(g.expect_z(q) - o.expect_pauli("Z", [q])) < tol
        assert abs(g.expect_x(q) - o.expect_pauli("X", [q])) < tol
        assert abs(g.expect_y(q) - o.expect_pauli("Y", [q])) < tol
    assert abs(g.expect_zprod([1, 4, 9]) - o.expect_pauli("ZZZ", [1, 4, 9])) < tol
    for ps, qs in workloads.random_pauli_strings(n, 40, 6, seed=5) + [("IXYZ", [0, 1, 2, 3]), ("YY", [10, 11]), ("YYY", [0, 6, 11])]:
        assert abs(g.expect_pauli(ps, qs) - o.expect_pauli(ps, qs)) < tol, (ps, qs)
    assert util.rel_err(g.state(), o.state) < TOL[prec]          # non-destructive
    # the literal vqe_lih.py ansatz (4 qubits) + batched call
    o4 = so.Oracle(4, prec); g4 = StateVector(4, prec)
    util.run_on_oracle(o4, workloads.c5_vqe_ansatz(4, 5)); g4.apply_circuit(workloads.c5_vqe_ansatz(4, 5))
    terms = [("ZZ", [0, 1]), ("XX", [2, 3]), ("Z", [0]), ("YY", [1, 2]), ("ZIZ", [0, 1, 3])]
    paulis = "".join(t[0] for t in terms).encode()
    qubits = capi.uarr([q for t in terms for q in t[1]])
    offs, acc = [0], 0
    for t in terms:
        acc += len(t[0]); offs.append(acc)
    res = (C.c_double * len(terms))()
    assert g4.lib.rocsvxGetExpectationPauliBatch(g4.h, g4.d, 4, paulis, qubits, capi.uarr(offs), len(terms), res) == 0
    for r, t in zip(res, terms):
        assert abs(r - o4.expect_pauli(*t)) < tol


@pytest.mark.parametrize("prec", PRECS)
def test_sampling_and_measure_bit_exact(prec):
    n = 14
    gates = workloads.c2_random_unitary(n, 4, seed=3)
    o = so.Oracle(n, prec, seed=99); util.run_on_oracle(o, gates)
    g = StateVector(n, prec, seed=99); g.set_state(o.state)      # identical amplitudes -> identical integer masses
    qs = [3, 0, 13, 7]
    assert np.array_equal(g.sample(qs, 5000), o.sample(qs, 5000))
    assert np.array_equal(g.sample(list(range(n)), 3000), o.sample(list(range(n)), 3000))   # second call: next counter
    for q in (2, 9, 13):
        og, pg = g.measure(q)
        oo, po = o.measure(q)
        assert og == oo and pg == po
        assert util.rel_err(g.state(), o.state) < TOL[prec]
    # statistics of the reference's own checks: tests/test_bindings.py:66-68 (+-10 %)
    b = StateVector(2, prec, seed=5); b.gate("h", 0); b.gate("cnot", 0, 1)
    s = b.sample([0, 1], 2000)
    assert set(np.unique(s)) <= {0, 3} and abs((s == 0).sum() - 1000) < 200


def test_swap_index_bits_and_batch_slice():
    n = 10
    o, g = _pair(n, "c64", batch=2, seed=4)
    for a, b in ((0, 9), (3, 4), (8, 2)):
        o.swap_index_bits(a, b); g.swap_index_bits(a, b)
    assert np.array_equal(g.state(), o.state)                       # pure data movement: bit-exact
    assert np.array_equal(g.state_slice(1), o.state[1 << n:])


def test_status_codes():
    lib = capi.load("c64")
    g = StateVector(3, "c64")
    assert g.gate_status("h", 3) == capi.INVALID_VALUE               # hipStateVec.cpp:108-110
    assert g.gate_status("cnot", 1, 1) == capi.INVALID_VALUE         # :439-443
    assert g.gate_status("cswap", 0, 1, 1) == capi.INVALID_VALUE     # :657-664
    assert lib.rocsvApplyMultiControlledX(g.h, g.d, 3, capi.uarr([0]), 0, 1) == capi.INVALID_VALUE   # :605-607
    assert lib.rocsvApplyMultiControlledX(g.h, g.d, 3, capi.uarr([1]), 1, 1) == capi.INVALID_VALUE   # :619-622
    assert lib.rocsvApplyMultiControlledX(g.h, g.d, 3, capi.uarr(list(range(64))), 64, 1) == capi.NOT_IMPLEMENTED  # :613-615
    out = np.empty(8, dtype=np.complex64)
    assert lib.rocsvGetStateVectorSlice(g.h, g.d, out.ctypes.data_as(C.c_void_p), 1) == capi.INVALID_VALUE   # :717-719
    assert lib.rocsvApplyH(g.h, None, 3, 0) == capi.SUCCESS          # NULL d_state = the handle's state (:80-85)
    assert abs(g.state()[1] - 1 / math.sqrt(2)) < 1e-6
    assert lib.rocsvEnsurePinnedBuffer(g.h, 1 << 20) == capi.SUCCESS and lib.rocsvGetPinnedBufferPointer(g.h)
    assert lib.rocsvFreePinnedBuffer(g.h) == capi.SUCCESS and not lib.rocsvGetPinnedBufferPointer(g.h)
    assert lib.rocsvFreeState(g.h) == capi.SUCCESS
    assert lib.rocsvApplyH(g.h, None, 3, 0) == capi.INVALID_VALUE    # no state any more (:282)


def test_full_size_properties_30_qubits():
    """BASELINE configs[1] size: properties that need no full-state oracle."""
    n = 30
    g = StateVector(n, "c64")
    gates = workloads.c2_random_unitary(n, 3, seed=30)
    g.apply_circuit(gates)
    assert abs(g.norm2() - 1) < 1e-4                                  # unitarity
    # running the inverse circuit returns to |0...0>
    inv = []
    for name, t, c, th, M in reversed(gates):
        inv.append((name, t, c, th, np.asarray(M).conj().T))
    g.apply_circuit(inv)
    assert abs(g.expect_zprod([0]) - 1) < 1e-4 and abs(g.expect_zprod([29]) - 1) < 1e-4
    s = g.sample(list(range(n)), 64)
    assert not s.any()
    # GHZ over all 30 qubits: only |0..0> and |1..1>, <Z0 Z29> = 1
    g.init(); g.apply_circuit(workloads.ghz(n))
    s = g.sample(list(range(n)), 256)
    assert set(np.unique(s)) <= {0, (1 << n) - 1} and 0 < (s == 0).sum() < 256
    assert abs(g.expect_zprod([0, 29]) - 1) < 1e-5 and abs(g.expect_pauli("X" * n, list(range(n))) - 1) < 1e-4


# ---- tensor-core 6-qubit blocks (complex64 only) ------------------------------------------------------------------
def test_block6_tensor_core_matches_oracle():
    rng = np.random.default_rng(11)
    for n, qs in ((13, [5, 6, 7, 8, 9, 10]), (15, [14, 2, 9, 0, 6, 11]), (19, [13, 14, 15, 16, 17, 18])):
        U = workloads.haar_unitary(rng, 64)
        o, g = _pair(n, "c64", seed=n)
        o.apply_matrix(qs, U); g.apply_block6(qs, U)
        assert util.rel_err(g.state(), o.state) < 2e-6            # bf16x3 split + fp32 accumulation in TMEM
    # a non-unitary matrix must not be renormalised
    o, g = _pair(14, "c64", seed=2)
    M = rng.standard_normal((64, 64)) + 1j * rng.standard_normal((64, 64))
    o.apply_matrix([5, 7, 8, 10, 12, 13], M); g.apply_block6([5, 7, 8, 10, 12, 13], M)
    assert util.rel_err(g.state(), o.state) < 5e-6
    assert g.lib.rocsvxApplyBlock6(g.h, g.d, 14, capi.uarr([5, 5, 8, 10, 12, 13]), (C.c_double * 8192)()) == capi.INVALID_VALUE
    d = StateVector(14, "c128")
    assert d.lib.rocsvxSetTensorCoreBlocks(d.h, 1) == capi.NOT_IMPLEMENTED


@pytest.mark.parametrize("n", [14, 22])
def test_circuit_with_tensor_core_blocks(n):
    gates = workloads.c2_random_unitary(n, 10, seed=30) + workloads.c1_ghz_random_layers(n, 4, seed=20)[n:]
    o = so.Oracle(n, "c64"); util.run_on_oracle(o, gates)
    g = StateVector(n, "c64"); g.set_tensor_core_blocks(True); g.apply_circuit(gates)
    assert util.rel_err(g.state(), o.state) < TOL["c64"]
    plain = StateVector(n, "c64"); plain.apply_circuit(gates)
    assert g.stats().sweeps != plain.stats().sweeps                  # blocks were actually formed
    assert abs(g.norm2() - 1) < 2e-5


def test_block6_batch_and_low_qubits():
    # blocks on index bits 0-4 (swizzled tile layout) and a batch of states (extra tensor-map dimension)
    rng = np.random.default_rng(12)
    for n, batch, qs in ((13, 3, [0, 1, 2, 3, 4, 5]), (15, 2, [1, 2, 3, 4, 5, 6]), (16, 2, [10, 11, 12, 13, 14, 15]), (17, 1, [3, 4, 5, 6, 7, 8])):
        U = workloads.haar_unitary(rng, 64)
        o, g = _pair(n, "c64", batch=batch, seed=n)
        o.apply_matrix(qs, U); g.apply_block6(qs, U)
        assert g.stats().blockSweeps == 1
        assert util.rel_err(g.state(), o.state) < 2e-6


def test_default_path_forms_blocks_from_24_qubits():
    n = 24
    gates = workloads.c2_random_unitary(n, 6, seed=30)
    o = so.Oracle(n, "c64"); util.run_on_oracle(o, gates)
    g = StateVector(n, "c64"); g.apply_circuit(gates)          # default = auto: tensor-core blocks on at >= 24 qubits
    assert g.stats().blockSweeps > 0
    assert util.rel_err(g.state(), o.state) < TOL["c64"]
    off = StateVector(n, "c64"); off.set_tensor_core_blocks(False); off.apply_circuit(gates)
    assert off.stats().blockSweeps == 0
    assert util.rel_err(off.state(), o.state) < TOL["c64"]


def test_blocks_on_sparse_states_circuit_and_inverse():
    # |0..0> -> circuit -> inverse circuit: most tiles are empty or hold rounding noise; the blocks' norm correction must
    # not pick its factor up from such tiles (regression: the norm fell to 0.90)
    n = 24
    for depth in (2, 3):
        gates = workloads.c2_random_unitary(n, depth, seed=30)
        inv = [(name, t, c, th, np.asarray(M).conj().T) for name, t, c, th, M in reversed(gates)]
        g = StateVector(n, "c64"); g.set_tensor_core_blocks(True)
        g.apply_circuit(gates); g.apply_circuit(inv)
        assert g.stats().blockSweeps > 0
        assert abs(g.norm2() - 1) < 2e-5 and abs(g.expect_zprod([0]) - 1) < 2e-5 and abs(g.expect_zprod([n - 1]) - 1) < 2e-5
        assert not g.sample(list(range(n)), 64).any()


def test_distributed_entry_points_on_one_rank_form_blocks():
    """rocsvAllocateDistributedState without rocsvxDistInit = one rank: the distributed code path (planner, RUN steps) on
    a single GPU, which now forms tensor-core blocks on the local qubits like the plain path."""
    n = 24
    lib = capi.load("c64")
    gates = workloads.c4_global_layers(n, 6, seed=36, top=3) + workloads.c2_random_unitary(n, 4, seed=30)
    o = so.Oracle(n, "c64"); util.run_on_oracle(o, gates)
    h = C.c_void_p()
    assert lib.rocsvCreate(C.byref(h)) == 0
    assert lib.rocsvAllocateDistributedState(h, n) == 0 and lib.rocsvInitializeDistributedState(h) == 0
    arr, keep = capi.make_ops(gates)
    assert lib.rocsvxApplyCircuit(h, None, n, arr, len(gates)) == 0
    out = np.empty(1 << n, dtype=np.complex64)
    assert lib.rocsvGetStateVectorFull(h, None, out.ctypes.data_as(C.c_void_p)) == 0
    st = capi.Stats()
    assert lib.rocsvxGetStats(h, C.byref(st), 0) == 0
    assert st.blockSweeps > 0
    assert util.rel_err(out, o.state) < TOL["c64"]
    assert lib.rocsvDestroy(h) == 0


@pytest.mark.parametrize("n", [14, 24])
def test_plan_cache_replays_identical_circuits_only(n):
    """rocsvxApplyCircuit keeps the launches of the last circuit: the identical gate list replays them (bit-identical state,
    planCacheHits counts), anything else -- other matrices, other settings -- is planned afresh."""
    gates = workloads.c2_random_unitary(n, 5, seed=30) + workloads.c1_ghz_random_layers(n, 3, seed=20)[n:]
    other = workloads.c2_random_unitary(n, 5, seed=31) + workloads.c1_ghz_random_layers(n, 3, seed=20)[n:]
    g = StateVector(n, "c64")
    g.apply_circuit(gates); first = g.state()
    assert g.stats().planCacheHits == 0
    g.init(); g.apply_circuit(gates)
    assert g.stats().planCacheHits == 1
    assert np.array_equal(g.state(), first)                          # same launches, same bits
    g.apply_circuit(gates)                                           # a third time on top: still the cached launches
    assert g.stats().planCacheHits == 2
    twice = g.state()
    g.init(); g.apply_circuit(other)                                 # same shape, other matrices: a miss
    assert g.stats().planCacheHits == 2
    o = so.Oracle(n, "c64"); util.run_on_oracle(o, other)
    assert util.rel_err(g.state(), o.state) < TOL["c64"]
    g.init(); g.apply_circuit(gates); g.apply_circuit(gates)         # back to the first circuit: miss, then hit
    assert g.stats().planCacheHits == 3
    assert np.array_equal(g.state(), twice)
    o = so.Oracle(n, "c64"); util.run_on_oracle(o, gates); util.run_on_oracle(o, gates)
    assert util.rel_err(twice, o.state) < TOL["c64"]
    g.set_tensor_core_blocks(n < 20)                                 # a setting changes the plan: miss
    g.init(); g.apply_circuit(gates)
    assert g.stats().planCacheHits == 3
    assert util.rel_err(g.state(), first) < TOL["c64"]
    d = StateVector(12, "c128"); q = workloads.c3_qft(12, seed=33)
    d.apply_circuit(q); a = d.state(); d.init(); d.apply_circuit(q)
    assert d.stats().planCacheHits == 1 and np.array_equal(d.state(), a)


# ---- round 2 -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", PRECS)
def test_apply_matrix_and_measure(prec):
    """rocsvApplyMatrixAndMeasure (hipStateVec.h:487-494) = ApplyMatrix then Measure on the shared Philox stream: outcome
    and post-measurement state against the oracle's two halves, for 1-, 2- and 3-qubit matrices, eager and deferred."""
    n = 12
    rng = np.random.default_rng(77)
    for fusion in (False, True):
        v = util.random_state(n, 1, seed=9)
        o = so.Oracle(n, prec, seed=31); o.set_state(v)
        g = StateVector(n, prec, fusion=fusion, seed=31); g.set_state(v)
        for k, q in ((1, 0), (2, 7), (3, 11), (2, 3)):
            t = [int(x) for x in rng.permutation(n)[:k]]
            U = workloads.haar_unitary(rng, 1 << k)
            o.apply_matrix(t, U)
            want, _ = o.measure(q)
            got = g.apply_matrix_and_measure(t, U, q)
            assert got == want
            assert util.rel_err(g.state(), o.state) < TOL[prec]
    g = StateVector(3, prec)
    out = C.c_int()
    assert g.lib.rocsvApplyMatrixAndMeasure(g.h, g.d, 3, capi.uarr([0]), 1, None, 0, C.byref(out)) == capi.INVALID_VALUE
    assert g.lib.rocsvApplyMatrixAndMeasure(g.h, g.d, 3, capi.uarr([0]), 1, C.c_void_p(g.d.value), 5, C.byref(out)) == capi.INVALID_VALUE


@pytest.mark.parametrize("prec", PRECS)
def test_batched_expectation_groups_by_x_mask(prec):
    """rocsvxGetExpectationPauliBatch: one read sweep per x-mask group, every result against oracle.expect_pauli."""
    n = 14
    o, g = _pair(n, prec, seed=21)
    gates = workloads.c5_vqe_ansatz(n, seed=5)
    util.run_on_oracle(o, gates); g.apply_circuit(gates)
    tol = 1e-5 if prec == "c64" else 1e-12
    zterms = [("Z" * len(q), q) for q in ([0], [3], [0, 3], [1, 2, 13], [5, 6, 7, 8], list(range(n)))]
    xy = [("XX", [2, 3]), ("YY", [2, 3]), ("XY", [2, 3]), ("YX", [3, 2]), ("XZX", [2, 9, 3]), ("YZZY", [3, 0, 13, 2])]     # one x-mask {2,3}
    rnd = workloads.random_pauli_strings(n, 40, 6, seed=8)
    terms = zterms + xy + rnd + [("", []), ("I", [4])]
    g.stats(reset=True)
    got = g.expect_batch(terms)
    for r, t in zip(got, terms):
        assert abs(r - o.expect_pauli(*t)) < tol, t
    masks = {frozenset(q for p, q in zip(*t) if p in "XY") for t in terms}
    assert g.stats().expectationSweeps == len(masks) < len(terms)
    assert util.rel_err(g.state(), o.state) < TOL[prec]              # non-destructive
    # more terms in one group than a sweep carries (16): 70 Z-strings -> 5 sweeps
    many = [("Z" * 3, [int(a), int(b), int(c)]) for a, b, c in (np.random.default_rng(i).permutation(n)[:3] for i in range(70))]
    g.stats(reset=True)
    got = g.expect_batch(many)
    assert g.stats().expectationSweeps == 5
    for r, t in zip(got, many):
        assert abs(r - o.expect_pauli(*t)) < tol
    # states of >= 2^15 loop indices take the sign-word kernel for groups of >= 4 terms: all-Z groups of 16 / 8 / 5 terms, groups
    # on one x-mask with mixed Y counts (4 and 7 terms), a pivot in the middle and on the top qubit
    n2 = 17
    o2, g2 = _pair(n2, prec, seed=23)
    gates2 = workloads.c5_vqe_ansatz(n2, seed=7) + workloads.c1_ghz_random_layers(n2, 2, seed=3)
    util.run_on_oracle(o2, gates2); g2.apply_circuit(gates2)
    rng = np.random.default_rng(99)
    big = [("Z" * k, sorted(int(x) for x in rng.permutation(n2)[:k])) for k in rng.integers(1, 7, size=29)]
    for a, b in ((4, 9), (0, 16)):
        mid = [q for q in range(a + 1, b)][:3]
        big += [(pa + "Z" * len(mid) + pb, [a] + mid + [b]) for pa, pb in (("X", "X"), ("Y", "Y"), ("X", "Y"), ("Y", "X"))]
    big += [("XZZY", [4, 1, 12, 9]), ("YZX", [4, 16, 9]), ("XX", [9, 4])]
    g2.stats(reset=True)
    got = g2.expect_batch(big)
    assert g2.stats().expectationSweeps == 2 + 1 + 1               # 29 Z-strings -> 16 + 13, then the two x-masks
    for r, t in zip(got, big):
        assert abs(r - o2.expect_pauli(*t)) < tol, t
    # parameter-shift batch: every state of a batch in the same sweeps
    b = 3
    v = util.random_state(9, b, seed=5)
    ob = [so.Oracle(9, prec) for _ in range(b)]
    for i, x in enumerate(ob):
        x.set_state(v[i << 9:(i + 1) << 9])
    gb = StateVector(9, prec, batch=b); gb.set_state(v)
    terms = [("ZZ", [0, 8]), ("XX", [1, 2]), ("YZ", [3, 4]), ("XY", [1, 2])]
    got = gb.expect_batch(terms, all_states=True)
    assert got.shape == (b, len(terms))
    for i in range(b):
        for r, t in zip(got[i], terms):
            assert abs(r - ob[i].expect_pauli(*t)) < tol
    bad = (C.c_double * 2)()
    assert g.lib.rocsvxGetExpectationPauliBatch(g.h, g.d, n, b"XQ", capi.uarr([0, 1]), capi.uarr([0, 1, 2]), 2, bad) == capi.INVALID_VALUE


def test_state_export_import_through_pinned_staging():
    """rocsvGetStateVectorFull / rocsvxSetStateVector of a state larger than the staging chunks (2 x 64 MB, double-buffered):
    bit-exact round trip, into pageable memory and into the handle's own pinned buffer."""
    n = 25                                                             # 256 MB of complex64: four chunks
    rng = np.random.default_rng(3)
    v = (rng.standard_normal(1 << n, dtype=np.float32) + 1j * rng.standard_normal(1 << n, dtype=np.float32)).astype(np.complex64)
    g = StateVector(n, "c64")
    g.set_state(v)
    assert np.array_equal(g.state(), v)
    assert g.lib.rocsvEnsurePinnedBuffer(g.h, v.nbytes) == capi.SUCCESS
    ptr = g.lib.rocsvGetPinnedBufferPointer(g.h)
    pinned = np.ctypeslib.as_array((C.c_float * (2 << n)).from_address(ptr)).view(np.complex64)
    assert g.lib.rocsvGetStateVectorFull(g.h, g.d, C.c_void_p(ptr)) == capi.SUCCESS
    assert np.array_equal(pinned, v)
    odd = StateVector(21, "c128", batch=3)                             # 3 x 32 MB: a chunk boundary inside batch member 1
    w = util.random_state(21, 3, seed=2)
    odd.set_state(w)
    assert np.array_equal(odd.state(), w.astype(np.complex128))
    assert np.array_equal(odd.state_slice(2), w[2 << 21:])


def test_sampling_is_chunking_independent_and_handles_degenerate_states():
    n = 16
    gates = workloads.c2_random_unitary(n, 3, seed=3)
    o = so.Oracle(n, "c64", seed=4); util.run_on_oracle(o, gates)
    g = StateVector(n, "c64", seed=4); g.set_state(o.state)
    want = o.sample([0, 15, 7], 4096)
    assert np.array_equal(g.sample([0, 15, 7], 4096), want)
    # basis state: every shot the same index; duplicates and an empty qubit list are legal requests
    b = StateVector(n, "c64", seed=1); b.gate("x", 3); b.gate("x", 15)
    assert set(np.unique(b.sample(list(range(n)), 100))) == {(1 << 3) | (1 << 15)}
    assert set(np.unique(b.sample([3, 3, 0, 15], 10))) == {0b1011}
    assert not b.sample([], 5).any()
    z = StateVector(6, "c64"); z.set_state(np.zeros(64, dtype=np.complex64))
    out = (C.c_uint64 * 4)()
    assert z.lib.rocsvSample(z.h, z.d, 6, capi.uarr([0]), 1, 4, out) == capi.FAILURE       # nothing to draw from


def test_block_path_keeps_small_amplitudes_of_a_peaked_state():
    """Per-amplitude RELATIVE accuracy on a peaked state through the tensor-core block path.  RX(small) on every qubit gives
    a product state whose amplitudes span 1 ... eps^n; a 6-qubit block acting on it mixes only amplitudes of one column, so
    a correct fp32-class kernel keeps every amplitude accurate relative to its COLUMN's magnitude."""
    n = 20
    eps = 2e-2
    rng = np.random.default_rng(5)
    pre = [("rx", [q], [], 2 * eps) for q in range(n)]
    blk = [8, 9, 10, 11, 12, 13]
    U = workloads.haar_unitary(rng, 64)
    o = so.Oracle(n, "c128"); util.run_on_oracle(o, pre)
    g = StateVector(n, "c64"); g.apply_circuit(pre)
    o.apply_matrix(blk, U); g.apply_block6(blk, U)
    got, want = g.state().astype(np.complex128), o.state
    # column = all amplitudes that differ only in the block bits
    idx = np.arange(1 << n)
    mask = sum(1 << q for q in blk)
    col = idx & ~mask
    colmax = np.zeros(1 << n)
    np.maximum.at(colmax, col, np.abs(want))
    rel = np.abs(got - want) / colmax[col]
    assert rel.max() < 1e-5, rel.max()


def test_c2_full_depth_parity_n26():
    """configs[1] in full depth (40 layers, 63-class plan of stacked fp16-split blocks) on 26 qubits, every amplitude
    against the oracle, plus 32 Pauli strings."""
    n = 26
    gates = workloads.c2_random_unitary(n, 40, seed=30)
    o = so.Oracle(n, "c64"); util.run_on_oracle(o, gates)
    g = StateVector(n, "c64"); g.apply_circuit(gates)
    assert g.stats().blockSweeps > 20
    assert util.rel_err(g.state(), o.state) < TOL["c64"]
    terms = workloads.random_pauli_strings(n, 32, 8, seed=6)
    got = g.expect_batch(terms)
    for r, t in zip(got, terms):
        assert abs(r - o.expect_pauli(*t)) < 1e-5


def test_c3_qft_parity_n24_c128():
    n = 24
    gates = workloads.c3_qft(n, seed=33)
    o = so.Oracle(n, "c128"); util.run_on_oracle(o, gates)
    g = StateVector(n, "c128"); g.apply_circuit(gates)
    assert util.rel_err(g.state(), o.state) < TOL["c128"]


def test_c2_at_30_qubits_against_the_c128_library():
    """BASELINE configs[1] at full size and depth: the complex64 engine (tensor-core blocks on) against the complex128
    library run of the same circuit on a strided subsample of 2^20 amplitudes and 64 random Pauli strings."""
    n = 30
    gates = workloads.c2_random_unitary(n, 40, seed=30)
    terms = workloads.random_pauli_strings(n, 64, 8, seed=7)
    stride = 1 << (n - 20)
    d = StateVector(n, "c128"); d.apply_circuit(gates)
    want_e = d.expect_batch(terms)
    want = _strided(d, n, stride, np.complex128)
    del d
    import gc; gc.collect()
    g = StateVector(n, "c64"); g.apply_circuit(gates)
    assert g.stats().blockSweeps >= 50
    got = _strided(g, n, stride, np.complex64)
    ref = np.abs(want).max()
    assert np.abs(got - want).max() / ref < TOL["c64"]
    got_e = g.expect_batch(terms)
    assert np.abs(got_e - want_e).max() < 1e-5


def _strided(sv, n, stride, dtype):
    """every stride-th amplitude, read through torch from the library's device pointer (test plumbing only)"""
    import torch
    N = 1 << n
    sv.sync()
    out = np.empty(N // stride, dtype=dtype)
    # a device gather through cudaMemcpy2D: rows of one amplitude, pitch = stride amplitudes
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    esz = out.itemsize
    rc = rt.cudaMemcpy2D(out.ctypes.data_as(C.c_void_p), C.c_size_t(esz), C.c_void_p(sv.d.value), C.c_size_t(stride * esz),
                         C.c_size_t(esz), C.c_size_t(N // stride), C.c_int(2))
    assert rc == 0, rc
    return out


def test_eager_wide_matrices_take_the_block_sweep():
    """rocsvApplyMatrix / rocsvApplyControlledMatrix with five or six qubits in all (complex64, n >= 13): one tensor-core block
    pass instead of the gather kernel; seven and more still gather.  Same results as the oracle either way."""
    n = 15
    rng = np.random.default_rng(44)
    o, g = _pair(n, "c64", seed=15)
    # (targets, controls): neighbouring and scattered sets; a set the block kernel's tile geometry cannot move (more than five
    # tensor-map dimensions) and everything wider than six qubits stays on the gather kernel
    cases = [([3, 4, 5, 6, 7], []), ([9, 8, 12, 10, 11, 13], []), ([0, 1, 2, 3, 4], [5]), ([14, 2, 7, 11, 5], []), ([1, 3, 5, 7, 9, 11], []),
             ([2, 3, 4, 5, 6, 7, 8], []), ([6, 7, 8, 9, 10, 11], [0])]
    blocks = 0
    for t, c in cases:
        U = workloads.haar_unitary(rng, 1 << len(t))
        before = g.stats().blockSweeps
        o.apply_matrix(t, U, c); g.apply_matrix(t, U, c)
        used = g.stats().blockSweeps - before
        assert used <= (1 if len(t) + len(c) <= 6 else 0)
        blocks += used
    assert blocks >= 3
    assert util.rel_err(g.state(), o.state) < TOL["c64"]
    # a non-unitary 5-qubit matrix keeps its norm change (no renormalisation)
    M = rng.standard_normal((32, 32)) + 1j * rng.standard_normal((32, 32))
    o.apply_matrix([3, 9, 1, 14, 6], M); g.apply_matrix([3, 9, 1, 14, 6], M)
    assert util.rel_err(g.state(), o.state) < TOL["c64"]


def test_block6_on_two_runs_of_index_bits():
    """A block made of two separate runs of index bits (what neighbouring logical qubits look like after an index-bit exchange
    in a distributed slice): the tile's column bits are then chosen next to a block run so that the tile still fits a
    five-dimensional tensor map -- same kernel, same results."""
    rng = np.random.default_rng(21)
    for n, batch, qs, blocks in ((20, 1, [11, 12, 13, 17, 18, 19], 1), (22, 1, [9, 10, 18, 19, 20, 21], 1), (19, 1, [5, 6, 7, 8, 17, 18], 1),
                                 (24, 1, [4, 5, 21, 22, 23, 6], 1), (21, 1, [10, 20, 19, 18, 17, 16], 1), (20, 2, [8, 9, 10, 11, 15, 16], 1),
                                 (19, 2, [5, 6, 7, 8, 17, 18], 0)):          # (last: the batch needs a sixth dimension -> generic dense path)
        U = workloads.haar_unitary(rng, 64)
        o, g = _pair(n, "c64", batch=batch, seed=n)
        o.apply_matrix(qs, U); g.apply_block6(qs, U)
        assert g.stats().blockSweeps == blocks, qs
        assert util.rel_err(g.state(), o.state) < 2e-6
