"""Distributed host logic on the CPU: the step plan the engine would execute on P ranks (rocsvxDistPlanCircuit) is
re-simulated on one full state indexed by PHYSICAL position and must reproduce the logical circuit exactly."""
import numpy as np
import pytest

from oracle import sv_oracle as so
from rocquantum_b200 import workloads
from tests import util


def _check(n, nranks, gates, mode):
    m = nranks.bit_length() - 1
    nx, steps, fmap = util.dist_plan(n, nranks, gates, mode | 2, canonicalize=True)    # bit 1: fused + cut into sweeps
    assert fmap == list(range(n))                                   # readback layout is canonical
    v = util.random_state(n, seed=n * 7 + nranks)
    a = so.Oracle(n, "c128"); a.set_state(v); util.run_on_oracle(a, gates)
    b = so.Oracle(n, "c128"); b.set_state(v); util.simulate_dist_plan(b, n - m, steps)
    assert util.rel_err(b.state, a.state) < 1e-11
    return nx


@pytest.mark.parametrize("nranks", [2, 4, 8])
@pytest.mark.parametrize("mode", [0, 1, 8])
@pytest.mark.parametrize("n", [8, 11, 14])
def test_random_bags(n, nranks, mode):
    for seed in range(3):
        _check(n, nranks, util.random_gates(n, 120, seed=100 * seed + n, maxk=3), mode)
        _check(n, nranks, util.random_gates(n, 120, seed=100 * seed + n + 1, allow_matrix=False), mode)


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_c4_and_qft(nranks):
    n = 12
    nx = _check(n, nranks, workloads.c4_global_layers(n, 8, seed=36, top=3), 0)
    nx1 = _check(n, nranks, workloads.c4_global_layers(n, 8, seed=36, top=3), 1)
    assert nx <= nx1                                                # look-ahead never needs more exchanges than per-gate
    nx8 = _check(n, nranks, workloads.c4_global_layers(n, 8, seed=36, top=3), 8)    # strictly in program order
    assert nx <= nx8
    _check(n, nranks, workloads.c3_qft(n, seed=33), 0)
    _check(n, nranks, workloads.c2_random_unitary(n, 6, seed=30), 0)


def test_diagonal_and_controlled_gates_on_rank_bits_need_no_exchange():
    n, nranks = 10, 4
    gates = [("h", [0], [], 0.0), ("cz", [1, 9], [], 0.0), ("rz", [8], [], 0.3), ("cnot", [2], [9], 0.0), ("z", [9], [], 0.0),
             ("crz", [8], [3], 0.5), ("mcx", [4], [8, 9], 0.0), ("t", [8], [], 0.0)]
    nx, steps, fmap = util.dist_plan(n, nranks, gates, 0, canonicalize=True)
    assert nx == 0
    _check(n, nranks, gates, 0)


@pytest.mark.parametrize("nranks", [2, 8])
def test_distributed_slices_form_blocks_on_local_qubits_only(nranks):
    """mode bit 2: every RUN step planned with tensor-core blocks (what the complex64 engine does on a slice): blocks lie on
    local positions, hold only ops inside them, and the plan still equals the circuit."""
    n = 17
    m = nranks.bit_length() - 1
    for gates in (workloads.c4_global_layers(n, 8, seed=36, top=3), workloads.c2_random_unitary(n, 6, seed=30),
                  workloads.c3_qft(n, seed=33) + util.random_gates(n, 80, seed=3, maxk=3)):
        nx, steps, fmap = util.dist_plan(n, nranks, gates, 2 | 4, canonicalize=True)
        assert fmap == list(range(n))
        v = util.random_state(n, seed=nranks)
        a = so.Oracle(n, "c128"); a.set_state(v); util.run_on_oracle(a, gates)
        b = so.Oracle(n, "c128"); b.set_state(v); util.simulate_dist_plan(b, n - m, steps)
        assert util.rel_err(b.state, a.state) < 1e-11
    assert util.dist_plan.blocks > 0


@pytest.mark.parametrize("nranks", [2, 8])
def test_deferral_cuts_the_exchanges_of_a_brick_circuit(nranks):
    """Every layer of configs[3] puts a dense gate on every qubit: in program order the parked qubits come back once per
    layer; with deferral the far side of the register runs ahead and a handful of exchanges is enough."""
    n, depth = 16, 12
    gates = workloads.c4_global_layers(n, depth, seed=36, top=3)
    nx_deferred = _check(n, nranks, gates, 0)
    nx_inorder = _check(n, nranks, gates, 8)
    assert nx_inorder >= depth - 1
    assert nx_deferred <= 4 and nx_deferred * 3 <= nx_inorder
    # canonicalize=False: the exchanges of the circuit proper
    nx_raw, _, _ = util.dist_plan(n, nranks, gates, 0, canonicalize=False)
    assert nx_raw <= 3
