"""Single-process multi-GPU (-m gpu): ONE handle, rocsvAllocateDistributedState, all slices driven from this process --
the reference's contract (MULTI_GPU_GUIDE.md:11-17, test_hipStateVec_multi_gpu.cpp:109-339, python/rocq/api.py:53-57).
On a box with fewer devices than slices the slices are placed round-robin (rocsvxDistSetRanks), so the distributed engine --
planner, rank-specialised sweeps, tensor-core blocks on slices, the peer-memory exchange kernel, exact all-gathers -- runs
and is compared with the oracle on a single GPU as well."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import sv_oracle as so
from rocquantum_b200 import capi, workloads
from rocquantum_b200.statevec import StateVector
from tests import util

pytestmark = pytest.mark.gpu
TOL = util.TOL


@pytest.mark.parametrize("prec", ["c64", "c128"])
def test_reference_multi_gpu_vectors(prec):
    """test_hipStateVec_multi_gpu.cpp: Test 1 (:83-159 allocation + |0..0> on rank 0), Test 2 (:161-227 X on q0),
    Test 3 (:229-283 X, CNOT(0,1) -> index 3), Test 4 (:285-339 fused H on q0), 3 qubits, slices read raw per rank."""
    S = 1 / math.sqrt(2)
    g = StateVector(3, prec, ranks=2)
    rank, P, nl = g.dist_info()
    assert (P, nl) == (2, 2)                                      # numGlobalSliceQubits = 1, numLocalQubitsPerGpu = 2
    s0, s1 = g.rank_slice(0), g.rank_slice(1)
    assert s0[0] == 1 and not s0[1:].any() and not s1.any()
    assert g.lib.rocsvApplyX(g.h, None, 3, 0) == capi.SUCCESS
    s0, s1 = g.rank_slice(0), g.rank_slice(1)
    assert s0[1] == 1 and not s0[[0, 2, 3]].any() and not s1.any()
    assert g.lib.rocsvApplyCNOT(g.h, None, 3, 0, 1) == capi.SUCCESS
    s0, s1 = g.rank_slice(0), g.rank_slice(1)
    assert s0[3] == 1 and not s0[:3].any() and not s1.any()
    g = StateVector(3, prec, ranks=2)
    g.apply_fused_1q(0, so.gate_matrix("h"))
    s0, s1 = g.rank_slice(0), g.rank_slice(1)
    assert np.allclose(s0[:2], [S, S], atol=1e-6) and not s0[2:].any() and not s1.any()
    # a gate on the slice-selecting qubit (global): X(q2) moves the amplitude to rank 1
    g = StateVector(3, prec, ranks=2)
    g.gate("x", 2)
    full = g.state()
    assert full[4] == 1 and np.count_nonzero(full) == 1


@pytest.mark.parametrize("prec", ["c64", "c128"])
@pytest.mark.parametrize("ranks", [2, 4])
def test_group_matches_oracle(prec, ranks):
    for n in (8, 15):
        named = util.random_gates(n, 150, seed=n + ranks, allow_matrix=False)
        mixed = util.random_gates(n, 150, seed=3 * n + ranks, maxk=3) + workloads.c4_global_layers(n, 8, seed=36, top=3)
        # one rocsvApply* call per gate, eager: every global target triggers an exchange
        o = so.Oracle(n, prec); util.run_on_oracle(o, named)
        g = StateVector(n, prec, ranks=ranks); util.run_per_gate(g, named)
        assert g.dist_info()[1] == ranks
        assert util.rel_err(g.state(), o.state) < TOL[prec]
        # the same through the deferred queue
        g = StateVector(n, prec, ranks=ranks, fusion=True); util.run_per_gate(g, named)
        assert util.rel_err(g.state(), o.state) < TOL[prec]
        # whole circuit: deferring planner, all rank bits traded at once; device matrices through the C ABI
        o = so.Oracle(n, prec, seed=5); util.run_on_oracle(o, mixed)
        g = StateVector(n, prec, ranks=ranks, seed=5); g.apply_circuit(mixed)
        tol = 1e-5 if prec == "c64" else 1e-12
        for ps, qs in [("Z", [n - 1]), ("X", [n - 1]), ("ZZ", [0, n - 1]), ("XY", [n - 2, n - 1]), ("YZX", [1, n - 1, 4])]:
            assert abs(g.expect_pauli(ps, qs) - o.expect_pauli(ps, qs)) < tol
        terms = [("ZZ", [0, n - 1]), ("XX", [n - 1, 2]), ("Y", [n - 2])]
        assert np.abs(g.expect_batch(terms) - np.array([o.expect_pauli(*t) for t in terms])).max() < tol
        assert g.stats().exchanges > 0
        assert util.rel_err(g.state(), o.state) < TOL[prec]
        e = StateVector(n, prec, ranks=ranks); util.run_per_gate(e, mixed)          # rocsvApplyMatrix per gate on slices
        assert util.rel_err(e.state(), o.state) < TOL[prec]
        # sampling / measurement on identical amplitudes: bit-exact, whatever the number of slices
        g.set_state(o.state)
        qs = [n - 1, 0, 3, n - 2]
        assert np.array_equal(g.sample(qs, 2000), o.sample(qs, 2000))
        for q in (n - 1, 2):
            assert g.measure(q) == o.measure(q)
        assert util.rel_err(g.state(), o.state) < TOL[prec]
        # rocsvSwapIndexBits incl. local<->global and global<->global
        g.set_state(o.state)
        pairs = [(0, n - 1), (n - 2, 3), (n - 1, n - 2)] if ranks >= 4 else [(0, n - 1), (n - 1, 5)]
        for a, b in pairs:
            g.swap_index_bits(a, b); o.swap_index_bits(a, b)
        assert np.array_equal(g.state(), o.state)


@pytest.mark.parametrize("ranks", [2, 4])
def test_group_tensor_core_blocks_on_slices(ranks):
    n = 22
    gates = workloads.c4_global_layers(n, 10, seed=36, top=3) + workloads.c2_random_unitary(n, 4, seed=30)
    o = so.Oracle(n, "c64"); util.run_on_oracle(o, gates)
    g = StateVector(n, "c64", ranks=ranks)
    g.set_tensor_core_blocks(True)
    g.apply_circuit(gates)
    st = g.stats()
    assert st.blockSweeps > 0 and st.exchanges > 0
    assert util.rel_err(g.state(), o.state) < TOL["c64"]
    assert abs(g.norm2() - 1) < 2e-5


def test_group_lifecycle_and_fallbacks():
    lib = capi.load("c64")
    g = StateVector(10, "c64", ranks=4)
    assert g.dist_info()[1] == 4
    assert lib.rocsvAllocateDistributedState(g.h, 12) == capi.SUCCESS and lib.rocsvInitializeDistributedState(g.h) == capi.SUCCESS
    g.n = 12
    assert g.dist_info()[2] == 10 and g.state()[0] == 1
    assert lib.rocsvxDistInit(g.h, 0, 1, None) == capi.INVALID_VALUE            # a group front cannot become a multi-process rank
    assert lib.rocsvxDistGetRankSlice(g.h, 4, None, None) == capi.INVALID_VALUE
    # a plain state replaces the distributed one
    d = C.c_void_p()
    assert lib.rocsvAllocateState(g.h, 5, C.byref(d), 1) == capi.SUCCESS and lib.rocsvInitializeState(g.h, d, 5) == capi.SUCCESS
    assert lib.rocsvApplyX(g.h, d, 5, 4) == capi.SUCCESS
    out = np.empty(32, dtype=np.complex64)
    assert lib.rocsvGetStateVectorFull(g.h, d, out.ctypes.data_as(C.c_void_p)) == capi.SUCCESS and out[16] == 1
    # too few qubits for the requested slices: at least two local qubits per slice
    t = StateVector(3, "c64", ranks=4)
    assert t.dist_info()[1] == 2
    one = StateVector(6, "c64", ranks=1)
    assert one.dist_info()[1] == 1
    one.gate("h", 5)
    assert abs(one.state()[32] - 1 / math.sqrt(2)) < 1e-6
