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
