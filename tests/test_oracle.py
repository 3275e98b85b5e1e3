"""The oracle against (a) the reference's own golden vectors, (b) the reference's own sources compiled
under the host HIP shim (oracle/_ref), (c) the Random123 Philox known-answer vectors.  CPU only."""
import math

import numpy as np
import pytest

from oracle import sv_oracle as so
from tests import util

S = 1 / math.sqrt(2)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert so.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert so.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert so.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_fix88_is_exact_floor():
    from fractions import Fraction
    rng = np.random.default_rng(0)
    for p in list(rng.uniform(0, 1, 50)) + [0.0, 1.0, 2.0 ** -60, 2.0 ** -88, 2.0 ** -89, 3e-27, 0.5]:
        assert so.fix88(float(p)) == int(Fraction(float(p)) * (1 << 88))


@pytest.mark.parametrize("prec", ["c64", "c128"])
def test_reference_golden_vectors(prec):
    tol = 1e-6   # the reference's own tolerance, test_hipStateVec_multi_gpu.cpp:28
    # Bell: test_bindings.py:56-70, tests/test_bindings.py:42-49
    o = so.Oracle(2, prec); o.gate("h", 0); o.gate("cnot", 0, 1)
    assert np.allclose(o.state, [S, 0, 0, S], atol=tol)
    # RX(pi/2)|0> = [cos pi/4, -i sin pi/4]: integrations/pennylane-rocq/tests/test_device.py:38-46
    o = so.Oracle(1, prec); o.gate("rx", 0, math.pi / 2)
    assert np.allclose(o.state, [math.cos(math.pi / 4), -1j * math.sin(math.pi / 4)], atol=tol)
    # RY(pi/2)|0> = [cos pi/4, sin pi/4]: integrations/cirq-rocm/cirq_rocm/tests/test_simulator.py:34-41
    o = so.Oracle(1, prec); o.gate("ry", 0, math.pi / 2)
    assert np.allclose(o.state, [math.cos(math.pi / 4), math.sin(math.pi / 4)], atol=tol)
    # H then RZ(pi/2) = [e^{-i pi/4}, e^{+i pi/4}]/sqrt2: integrations/qiskit-rocquantum-provider/tests/test_backend.py:36-44
    o = so.Oracle(1, prec); o.gate("h", 0); o.gate("rz", 0, math.pi / 2)
    assert np.allclose(o.state, np.array([np.exp(-1j * math.pi / 4), np.exp(1j * math.pi / 4)]) * S, atol=tol)
    # 3 qubits: X(q0) -> index 1; X(q0),CNOT(0,1) -> index 3; H(q0) -> idx 0,1 = 1/sqrt2
    # rocquantum/src/hipStateVec/test_hipStateVec_multi_gpu.cpp:185-207, 251-259, 300-339
    o = so.Oracle(3, prec); o.gate("x", 0)
    assert abs(o.state[1] - 1) < tol and np.abs(np.delete(o.state, 1)).max() < tol
    o.gate("cnot", 0, 1)
    assert abs(o.state[3] - 1) < tol and np.abs(np.delete(o.state, 3)).max() < tol
    o = so.Oracle(3, prec); o.apply_matrix([0], so.gate_matrix("h"))
    assert np.allclose(o.state[:2], [S, S], atol=tol) and np.abs(o.state[2:]).max() < tol
    # CCX |011> -> |111> (3 -> 7), CSWAP(0;1,2) 3 -> 5: tests/test_advanced_gates.py:58-63, 77-82
    o = so.Oracle(3, prec); o.gate("x", 0); o.gate("x", 1); o.gate("mcx", [0, 1], 2)
    assert abs(o.state[7] - 1) < tol
    o = so.Oracle(3, prec); o.gate("x", 0); o.gate("x", 1); o.gate("cswap", 0, 1, 2)
    assert abs(o.state[5] - 1) < tol
    # CRX(theta) on |10> (control=q0 set): tests/test_advanced_gates.py:31-42
    th = 0.7
    o = so.Oracle(2, prec); o.gate("x", 0); o.gate("crx", 0, 1, th)
    assert np.allclose(o.state, [0, math.cos(th / 2), 0, -1j * math.sin(th / 2)], atol=tol)
    # GHZ expectations: examples/expectation_example.py:55-57
    o = so.Oracle(3, prec); o.gate("h", 0); o.gate("cnot", 0, 1); o.gate("cnot", 1, 2)
    assert abs(o.expect_pauli("ZZ", [0, 1]) - 1) < 1e-6
    assert abs(o.expect_pauli("XY", [1, 2])) < 1e-6
    assert abs(o.expect_pauli("XYZ", [0, 1, 2])) < 1e-6
    # d/dtheta <Z> of RX(theta)|0> = -sin(theta): examples/gradient_example.py:56-61 (parameter shift)
    th = 0.37
    z = []
    for sgn in (+1, -1):
        o = so.Oracle(1, prec); o.gate("rx", 0, th + sgn * math.pi / 2); z.append(o.expect_pauli("Z", [0]))
    assert abs((z[0] - z[1]) / 2 + math.sin(th)) < 1e-6


@pytest.mark.parametrize("prec", ["c64", "c128"])
@pytest.mark.parametrize("n,batch", [(1, 1), (2, 3), (5, 1), (11, 2)])
def test_oracle_matches_compiled_reference(prec, n, batch):
    """Bit-exact agreement with the reference's own hipStateVec.cpp + kernels (25 defined entry points)."""
    if not so.ref_available(prec):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    gates = util.random_gates(n, 250, seed=100 + n, allow_matrix=False)
    v = util.random_state(n, batch, seed=n)
    o = so.Oracle(n, prec, batch=batch); o.set_state(v)
    r = so.RefLib(prec); r.allocate(n, batch); r.set_state(v)
    util.run_on_oracle(o, gates)
    util.run_on_ref(r, gates)
    assert np.array_equal(o.state, r.state())
    r.close()


@pytest.mark.parametrize("prec", ["c64", "c128"])
def test_golden_fixture_c1(prec):
    """tests/golden/c1_n10_*.npy was produced by the compiled reference (tests/golden/make_golden.py)."""
    import os
    from rocquantum_b200 import workloads
    path = os.path.join(os.path.dirname(__file__), "golden", f"c1_n10_{prec}.npy")
    want = np.load(path)
    o = so.Oracle(10, prec)
    util.run_on_oracle(o, workloads.c1_ghz_random_layers(10, 6, seed=20))
    assert np.array_equal(o.state, want)


@pytest.mark.parametrize("prec", ["c64", "c128"])
def test_golden_fixture_c2(prec):
    """tests/golden/c2_n10_*.npy: configs[1] in miniature, produced by the reference's own ApplyMatrix spec kernel gate by
    gate (tests/golden/make_golden.py).  The generic restatement reproduces it bit for bit; the route the GPU tests take
    (one-qubit gates through the reference's 2x2 kernel arithmetic) to rounding."""
    import os
    from rocquantum_b200 import workloads
    want = np.load(os.path.join(os.path.dirname(__file__), "golden", f"c2_n10_{prec}.npy"))
    gates = workloads.c2_random_unitary(10, 6, seed=30)
    o = so.Oracle(10, prec)
    for g in gates:
        o.apply_matrix(list(g[1]), g[4])
    assert np.array_equal(o.state, want)
    o2 = so.Oracle(10, prec)
    util.run_on_oracle(o2, gates)
    assert util.rel_err(o2.state, want) < (2e-6 if prec == "c64" else 1e-14)


def test_init_state_only_first_batch_member():
    # hipStateVec.cpp:260-268: (1,0) is written at absolute index 0 only
    o = so.Oracle(3, "c64", batch=2)
    assert o.state[0] == 1 and np.count_nonzero(o.state) == 1


def test_apply_matrix_bit_order_and_controls():
    # matrix index bit b <-> targets[b] (multi_qubit_kernels.hip:91-99); column-major is handled by the wrapper
    rng = np.random.default_rng(3)
    from rocquantum_b200.workloads import haar_unitary
    n = 5
    U = haar_unitary(rng, 4)
    v = util.random_state(n, seed=9)
    o = so.Oracle(n, "c128"); o.set_state(v); o.apply_matrix([3, 1], U, controls=[4])
    psi = v.reshape([2] * n)                      # axis 0 = qubit n-1 ... axis n-1 = qubit 0
    out = psi.copy()
    sub = psi[1]                                  # control qubit 4 = 1; remaining axes: q3,q2,q1,q0
    # U acts on (bit0 = q3, bit1 = q1): index = b_q3 + 2*b_q1
    U4 = U.reshape(2, 2, 2, 2)                    # [o1, o0, i1, i0] with bit1 first
    res = np.einsum("abcd,cxdy->axby", U4, sub.transpose(2, 1, 0, 3))   # in: [q1,q2,q3,q0] -> out [o_q1,q2,o_q3,q0]
    out[1] = res.transpose(2, 1, 0, 3)
    assert np.allclose(o.state, out.reshape(-1), atol=1e-12)


def test_sampling_spec_and_statistics():
    o = so.Oracle(3, "c128", seed=7); o.gate("h", 0); o.gate("cnot", 0, 1); o.gate("cnot", 1, 2)
    s = o.sample([0, 1, 2], 4000)
    assert set(np.unique(s)) <= {0, 7}
    assert abs((s == 0).mean() - 0.5) < 0.05          # examples/sampling_example.py:58-59
    # bit j of the result <-> measured[j]
    o = so.Oracle(3, "c128"); o.gate("x", 2)
    assert set(o.sample([2, 0], 10)) == {1} and set(o.sample([0, 2], 10)) == {2}
    # measure collapses and renormalises
    o = so.Oracle(2, "c128", seed=3); o.gate("h", 0); o.gate("cnot", 0, 1)
    out, p = o.measure(0)
    assert abs(p - 0.5) < 1e-12 and abs(o.norm2() - 1) < 1e-12
    assert abs(o.state[3 if out else 0]) == pytest.approx(1.0)


@pytest.mark.parametrize("prec", ["c64", "c128"])
def test_apply_matrix_matches_the_references_own_spec_kernel(prec):
    """rocsvApplyMatrix is declared but never defined by the reference; the kernel it was meant to
    launch, apply_multi_qubit_generic_matrix_kernel (multi_qubit_kernels.hip:37-115), is compiled unmodified into
    oracle/_ref and launched by oracle/hip_shim/spec_driver.cpp.  The oracle's restatement (and through it every GPU
    parity test of ApplyMatrix / the fused circuits of configs[1]) must agree with it BIT FOR BIT: same bit order of the
    matrix index, same column-major storage, same accumulation order in amplitude precision."""
    if not so.ref_available(prec):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    from rocquantum_b200.workloads import haar_unitary
    rng = np.random.default_rng(41)
    n = 9
    cases = [[0], [8], [3], [0, 1], [1, 0], [7, 2], [2, 7], [8, 0], [0, 4, 8], [5, 1, 3], [3, 2, 1], [0, 1, 2, 3], [8, 6, 1, 4], [2, 8, 0, 5]]
    for targets in cases:
        k = len(targets)
        U = haar_unitary(rng, 1 << k)
        v = util.random_state(n, seed=100 + k).astype(so.DT[prec])
        o = so.Oracle(n, prec); o.set_state(v); o.apply_matrix(targets, U)
        r = so.RefLib(prec); r.allocate(n); r.set_state(v); r.spec_apply_matrix(targets, U)
        got, want = o.state, r.state()
        r.close()
        assert np.array_equal(got, want), (prec, targets, float(np.abs(got - want).max()))
    # a non-unitary, non-symmetric matrix: transposition or conjugation mistakes cannot hide behind structure
    M = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    v = util.random_state(n, seed=7).astype(so.DT[prec])
    o = so.Oracle(n, prec); o.set_state(v); o.apply_matrix([6, 2], M)
    r = so.RefLib(prec); r.allocate(n); r.set_state(v); r.spec_apply_matrix([6, 2], M)
    assert np.array_equal(o.state, r.state())
    r.close()


def test_fused_circuit_matches_the_references_spec_kernel_gate_by_gate():
    """configs[1] in miniature (Haar one- and two-qubit gates, brick layers): the oracle run the GPU parity tests compare
    against equals the reference's own kernels applied gate by gate -- its ApplyMatrix spec kernel for every gate."""
    if not so.ref_available("c128"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    from rocquantum_b200 import workloads
    n = 10
    gates = workloads.c2_random_unitary(n, 4, seed=30)
    o = so.Oracle(n, "c128")
    r = so.RefLib("c128"); r.allocate(n)
    for g in gates:
        r.spec_apply_matrix(list(g[1]), g[4])
        o.apply_matrix(list(g[1]), g[4])            # the generic restatement for every gate, one-qubit ones included
    want = r.state()
    r.close()
    assert np.array_equal(o.state, want)
    o2 = so.Oracle(n, "c128")
    util.run_on_oracle(o2, gates)                   # the route the GPU tests use (one-qubit gates through matrix1)
    assert np.abs(o2.state - want).max() < 1e-14


@pytest.mark.parametrize("prec,tol", [("c64", 2e-6), ("c128", 1e-13)])
def test_measure_matches_the_references_measurement_kernels(prec, tol):
    """rocsvMeasure is declared only; its kernels exist (measurement_kernels.hip:12-99: prob0, collapse, sum of squares,
    renormalise) and are driven here for the outcome the oracle drew (the reference fixes no RNG: the stream is ours).
    Probability of the outcome and the post-measurement state must agree to summation-order accuracy -- the reference sums
    2^n terms sequentially in amplitude precision, the oracle exactly in fixed point."""
    if not so.ref_available(prec):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    n = 10
    seen = set()
    for q, seed in [(0, 1), (4, 2), (9, 3), (5, 4), (2, 5), (7, 6)]:
        v = util.random_state(n, seed=50 + seed).astype(so.DT[prec])
        o = so.Oracle(n, prec, seed=seed); o.set_state(v)
        outcome, prob = o.measure(q)
        seen.add(outcome)
        r = so.RefLib(prec); r.allocate(n); r.set_state(v)
        p0 = r.spec_measure_with_outcome(q, outcome)
        want = r.state()
        r.close()
        assert abs((p0 if outcome == 0 else 1.0 - p0) - prob) < 50 * tol
        assert util.rel_err(o.state, want) < tol
        bit = (np.arange(1 << n) >> q) & 1
        assert np.count_nonzero(o.state[bit != outcome]) == 0 and np.count_nonzero(want[bit != outcome]) == 0
    assert seen == {0, 1}                                           # both branches of the collapse were compared


@pytest.mark.parametrize("prec,tol", [("c64", 2e-5), ("c128", 1e-13)])
def test_z_expectations_match_the_references_outcome_probabilities(prec, tol):
    """<Z_q1 ... Z_qk> = sum over joint outcomes of (-1)^(number of ones) * probability (hipStateVec.h:382-400), with the
    probabilities taken from the reference's own calculate_multi_z_probabilities_kernel (measurement_kernels.hip:283-387)."""
    if not so.ref_available(prec):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    n = 9
    v = util.random_state(n, seed=77).astype(so.DT[prec])
    o = so.Oracle(n, prec); o.set_state(v)
    r = so.RefLib(prec); r.allocate(n); r.set_state(v)
    for qubits in ([0], [8], [3], [0, 1], [7, 2], [1, 4, 8], [8, 0, 5, 3], [0, 1, 2, 3, 4, 5, 6, 7]):
        probs = r.spec_multi_z_probabilities(qubits)
        assert abs(probs.sum() - 1.0) < 50 * tol
        signs = np.array([(-1) ** bin(b).count("1") for b in range(len(probs))], dtype=np.float64)
        want = float((signs * probs).sum())
        got = o.expect_pauli("Z" * len(qubits), qubits)
        assert abs(got - want) < 50 * tol, (qubits, got, want)
        # the bins themselves: bit j of the bin index is the value of qubits[j]
        idx = np.arange(1 << n)
        bins = np.zeros(1 << n, dtype=np.int64)
        for j, q in enumerate(qubits):
            bins |= ((idx >> q) & 1) << j
        mine = np.bincount(bins, weights=np.abs(v.astype(np.complex128)) ** 2, minlength=len(probs))
        assert np.abs(mine - probs).max() < 50 * tol
    r.close()


@pytest.mark.parametrize("prec", ["c64", "c128"])
def test_swap_index_bits_matches_the_references_local_permutation_kernel(prec):
    """rocsvSwapIndexBits (declared only), local<->local case: the reference's local_bit_swap_permutation_kernel
    (swap_kernels.hip:95-114), run as a real block of 2^n threads, moves every amplitude to the index with the two bits
    exchanged -- bit for bit what the oracle (and the engine's PERM sweep) does."""
    if not so.ref_available(prec):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    n = 7
    for a, b in [(0, 6), (6, 0), (2, 3), (1, 5), (4, 4)]:
        v = util.random_state(n, seed=60 + a).astype(so.DT[prec])
        o = so.Oracle(n, prec); o.set_state(v); o.swap_index_bits(a, b)
        r = so.RefLib(prec); r.allocate(n); r.set_state(v); r.spec_local_bit_swap(a, b)
        want = r.state()
        r.close()
        assert np.array_equal(o.state, want), (a, b)
        if a != b:
            assert not np.array_equal(want, v)


@pytest.mark.parametrize("prec,tol", [("c64", 2e-5), ("c128", 1e-12)])
def test_pauli_string_expectation_matches_the_references_basis_change_recipe(prec, tol):
    """rocsvGetExpectationPauliString (declared only) against reference code end to end: the reference's own recipe
    (rocquantum/utils/hamiltonian.py:35-59: Sdg then H on the Y qubits, H on the X qubits, then the Z product over all
    non-identity qubits), carried out with the reference's compiled gate kernels (rocsvApplySdg / rocsvApplyH) and its
    outcome-probability kernel (measurement_kernels.hip:283-387).  Also checks that the oracle leaves the state alone."""
    if not so.ref_available(prec):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    n = 9
    v = util.random_state(n, seed=88).astype(so.DT[prec])
    o = so.Oracle(n, prec); o.set_state(v)
    for paulis, qubits in [("X", [0]), ("Y", [8]), ("XY", [3, 5]), ("YX", [0, 8]), ("XYZ", [1, 2, 3]), ("ZYXI", [7, 0, 4, 2]),
                           ("YYYY", [8, 6, 1, 3]), ("XZXZXZ", [0, 1, 2, 3, 4, 5]), ("IXI", [2, 6, 7])]:
        r = so.RefLib(prec); r.allocate(n); r.set_state(v)
        for p, q in zip(paulis, qubits):
            if p == "Y":
                assert r.gate("sdg", q) == 0
        for p, q in zip(paulis, qubits):
            if p == "X":
                assert r.gate("h", q) == 0
        for p, q in zip(paulis, qubits):
            if p == "Y":
                assert r.gate("h", q) == 0
        zq = [q for p, q in zip(paulis, qubits) if p != "I"]
        probs = r.spec_multi_z_probabilities(zq)
        r.close()
        want = float(sum((-1) ** bin(b).count("1") * pb for b, pb in enumerate(probs)))
        got = o.expect_pauli(paulis, qubits)
        assert abs(got - want) < tol * 20, (paulis, qubits, got, want)
        assert np.array_equal(o.state, v)                           # non-destructive (hipStateVec.h:402-423)
