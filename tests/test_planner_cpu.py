"""Host logic without a GPU: algebraic fusion + sweep partition (rocsvxPlanCircuit) must be equivalent to the
gate-by-gate circuit.  The dumped plan is re-simulated on the oracle and compared with the oracle running
the original gate list."""
import numpy as np
import pytest

from oracle import sv_oracle as so
from rocquantum_b200 import workloads
from tests import util


def _check(n, gates, tile_bits=0, tol=1e-11):
    nsw, sweeps = util.plan(n, gates, tile_bits)
    v = util.random_state(n, seed=n + len(gates))
    a = so.Oracle(n, "c128"); a.set_state(v); util.run_on_oracle(a, gates)
    b = so.Oracle(n, "c128"); b.set_state(v); util.simulate_plan(b, sweeps)
    assert util.rel_err(b.state, a.state) < tol
    return nsw, sweeps


@pytest.mark.parametrize("n,tile_bits", [(3, 0), (6, 4), (9, 6), (12, 7), (14, 8)])
def test_mixed_gate_bag(n, tile_bits):
    _check(n, util.random_gates(n, 300, seed=n, maxk=min(3, n)), tile_bits)


@pytest.mark.parametrize("n,tile_bits", [(10, 6), (14, 8), (16, 13)])
def test_c2_random_unitary_plan(n, tile_bits):
    gates = workloads.c2_random_unitary(n, 8, seed=30)
    nsw, sweeps = _check(n, gates, tile_bits)
    nops = sum(len(s["ops"]) for s in sweeps)
    assert nops <= 8 * (n // 2 + 1)          # every 1q gate was absorbed into a neighbouring 2q matrix
    assert nsw < len(gates) // 4             # real fusion: far fewer sweeps than gates


def test_qft_plan_diagonals_do_not_force_residency():
    n, T = 14, 7
    gates = workloads.c3_qft(n, seed=33)
    nsw, sweeps = _check(n, gates, T)
    # controlled phases ride along in whatever sweep is open: about n/ (T - lowbits) sweeps for the H ladder + swaps
    assert nsw <= 2 * (n // 2) + 4


def test_c1_and_vqe_plans():
    _check(12, workloads.c1_ghz_random_layers(12, 6, seed=20), 8)
    _check(4, workloads.c5_vqe_ansatz(4, seed=5), 0)
    _check(10, workloads.c5_vqe_ansatz(10, seed=5), 6)


def test_single_gate_plans_are_one_sweep():
    for g in util.random_gates(9, 60, seed=5, maxk=3):
        nsw, sweeps = util.plan(9, [g], 6)
        assert nsw == 1 and len(sweeps[0]["ops"]) == 1


# ---- mixed plan: tensor-core blocks + ordinary sweeps (rocsvxPlanCircuitBlocks) ------------------------------------
def _check_blocks(n, gates, min_cost=0.0, tol=1e-11):
    nb, ns, steps = util.plan_blocks(n, gates, min_cost)
    for st in steps:
        if "blk" in st:                                   # a block: six positions, every op inside it
            assert len(st["blk"]) == 6 and sorted(st["blk"]) == st["blk"] and max(st["blk"]) < n
            for op in st["ops"]:
                qs = set(op["targets"]) | {q for q in range(64) if (op["cmask"] >> q) & 1}
                assert qs <= set(st["blk"])
            st.update(T=min(n, 13), rowbits=0, res=list(range(n)))      # simulate_plan: treat it as "everything resident"
            st["T"] = len(st["res"])
    v = util.random_state(n, seed=n + len(gates))
    a = so.Oracle(n, "c128"); a.set_state(v); util.run_on_oracle(a, gates)
    b = so.Oracle(n, "c128"); b.set_state(v); util.simulate_plan(b, steps)
    assert util.rel_err(b.state, a.state) < tol
    return nb, ns, steps


@pytest.mark.parametrize("n", [13, 16, 18])
def test_block_plan_is_equivalent_and_forms_blocks(n):
    gates = workloads.c2_random_unitary(n, 10, seed=30)
    nb, ns, steps = _check_blocks(n, gates)
    in_blocks = sum(len(st["ops"]) for st in steps if "blk" in st)
    total = sum(len(st["ops"]) for st in steps)
    assert nb >= 1 and in_blocks >= 0.8 * total            # a brick circuit goes (almost) entirely to blocks, also on qubits 0-4


def test_block_plan_mixed_bag_and_threshold():
    n = 15
    gates = util.random_gates(n, 250, seed=5, maxk=3)
    _check_blocks(n, gates)
    nb_hi, _, _ = _check_blocks(n, gates, min_cost=1e9)    # nothing reaches the threshold: ordinary sweeps only
    assert nb_hi == 0


def test_block_plan_large_brick_circuit_block_occupancy():
    # 30 qubits, depth 40 (BASELINE configs[1]): host-only planning, no simulation
    gates = workloads.c2_random_unitary(30, 40, seed=30)
    nb, ns, steps = util.plan_blocks(30, gates)
    in_blocks = sum(len(s["ops"]) for s in steps if "blk" in s)
    assert nb <= 90 and in_blocks / nb >= 6.5              # blocks stay full (a brick diamond holds 9 two-qubit matrices)


def test_block_plan_starts_a_brick_circuit_with_disjoint_blocks():
    """configs[1]: greedy absorption alone starts as a staircase of overlapping blocks (0-5, 4-9, 8-13, ...: 7 matrices
    each) and needs 70 passes; the planner also tries disjoint first blocks on untouched qubits and keeps the shorter plan:
    the second row already consists of full 9-matrix diamonds (plan_mixed, host_ops.h)."""
    gates = workloads.c2_random_unitary(30, 40, seed=30)
    nb, ns, steps = util.plan_blocks(30, gates)
    assert nb + ns <= 64                                   # one pass over the state per step; the diamond bound is 580 / 9 = 64.4 blocks
    first_row = [st["blk"] for st in steps[:5]]
    assert first_row == [list(range(6 * i, 6 * i + 6)) for i in range(5)]
    assert [len(st["ops"]) for st in steps[5:9]] == [9, 9, 9, 9]
    assert sum(len(st["ops"]) for st in steps) == 20 * 15 + 20 * 14      # every two-qubit matrix of the circuit, once


@pytest.mark.parametrize("n,depth", [(14, 12), (17, 9), (19, 7)])
def test_block_plan_with_spread_first_row_is_equivalent(n, depth):
    gates = workloads.c2_random_unitary(n, depth, seed=n)
    nb, ns, steps = _check_blocks(n, gates)
    assert nb >= 2


def test_qft33_plan_reads_the_circuit_backwards():
    """configs[2]: read forwards, the greedy partition has handed every resident slot to Hadamards before it meets the
    final swaps (6 sweeps, two of them pure permutations); read backwards the swaps seat both partners first and the
    Hadamards of those qubits join them (plan_sweeps, host_ops.h): 5 sweeps, none without arithmetic."""
    n = 33
    gates = workloads.c3_qft(n, seed=33)
    nsw, sweeps = util.plan(n, gates, 0, prec="c128")
    assert nsw == 5
    kinds = [sorted({op["kind"] for op in sw["ops"]}) for sw in sweeps]
    assert all(1 in k for k in kinds)                                   # every sweep carries Hadamards (DENSE), not only swaps
    hs = [sum(1 for op in sw["ops"] if op["kind"] == 1) for sw in sweeps]
    assert n - 2 <= sum(hs) <= n                                        # the Hadamards (algebraic fusion may pair the last ones)
    assert sum(len(sw["ops"]) for sw in sweeps) == 79                   # nothing lost, nothing doubled


@pytest.mark.parametrize("n,T,expect", [(14, 8, 3), (18, 9, 5), (17, 7, 6)])
def test_backward_qft_plans_are_equivalent(n, T, expect):
    """Small QFTs whose backward plan is one sweep shorter than the forward one (4 / 6 / 7 forwards, complex128 limits):
    the plan the engine would follow, replayed sweep by sweep on the oracle, is the circuit."""
    gates = workloads.c3_qft(n, seed=33)
    nsw, sweeps = util.plan(n, gates, T, prec="c128")
    assert nsw == expect
    v = util.random_state(n, seed=n)
    a = so.Oracle(n, "c128"); a.set_state(v); util.run_on_oracle(a, gates)
    b = so.Oracle(n, "c128"); b.set_state(v); util.simulate_plan(b, sweeps)
    assert util.rel_err(b.state, a.state) < 1e-11


def test_blockless_mixed_plan_uses_the_two_direction_sweep_planner():
    """complex64 from 24 qubits plans through plan_mixed; a circuit that forms no block (a QFT: one-qubit gates, wide
    diagonals, swaps) must still get the shorter of the forward / backward sweep partitions."""
    nb, ns, steps = util.plan_blocks(30, workloads.c3_qft(30, seed=33))
    assert nb == 0 and ns == 5
    _check_blocks(16, workloads.c3_qft(16, seed=33))
    _check_blocks(19, workloads.c3_qft(19, seed=33))
