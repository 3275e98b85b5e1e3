"""Synthetic circuits of BASELINE.json's configs, as lists of gate tuples
(name, targets, controls, theta[, matrix]) accepted by capi.make_ops / StateVector.apply_circuit.

All generators are seeded (numpy default_rng) so the GPU engine, the oracle and the compiled reference
run the identical circuit."""
from __future__ import annotations

import math

import numpy as np


def haar_unitary(rng: np.random.Generator, dim: int) -> np.ndarray:
    z = (rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))) / math.sqrt(2.0)
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return q * (d / np.abs(d))


def ghz(n: int):
    return [("h", [0], [], 0.0)] + [("cnot", [i + 1], [i], 0.0) for i in range(n - 1)]


def c1_ghz_random_layers(n: int = 20, depth: int = 20, seed: int = 20):
    """configs[0]: GHZ, then `depth` layers of random 1q rotations on every qubit + CZ/CNOT bricks.
    Uses only gates the reference defines, so the compiled reference (oracle/_ref) can run it."""
    rng = np.random.default_rng(seed)
    g = ghz(n)
    for layer in range(depth):
        for q in range(n):
            g.append((("rx", "ry", "rz")[int(rng.integers(3))], [q], [], float(rng.uniform(0, 2 * math.pi))))
        for q in range(layer % 2, n - 1, 2):
            if rng.integers(2):
                g.append(("cz", [q, q + 1], [], 0.0))
            else:
                g.append(("cnot", [q + 1], [q], 0.0))
    return g


def c2_random_unitary(n: int = 30, depth: int = 40, seed: int = 30):
    """configs[1]: per layer a Haar-random 1q unitary on every qubit, then Haar-random 2q unitaries on
    brick pairs (q, q+1), offset alternating with the layer."""
    rng = np.random.default_rng(seed)
    g = []
    for layer in range(depth):
        for q in range(n):
            g.append(("matrix", [q], [], 0.0, haar_unitary(rng, 2)))
        for q in range(layer % 2, n - 1, 2):
            g.append(("matrix", [q, q + 1], [], 0.0, haar_unitary(rng, 4)))
    return g


def c3_qft(n: int = 33, seed: int = 33):
    """configs[2]: X on a seeded subset, then the QFT exactly as benchmarks/run_benchmark.py:60-69 writes it:
    for i: H(i); for j > i: CP(pi/2^(j-i)) on (j, i); then floor(n/2) SWAPs."""
    rng = np.random.default_rng(seed)
    g = [("x", [q], [], 0.0) for q in range(n) if rng.integers(2)]
    for i in range(n):
        g.append(("h", [i], [], 0.0))
        for j in range(i + 1, n):
            ph = np.exp(1j * math.pi / (1 << (j - i)))
            g.append(("matrix", [j], [i], 0.0, np.array([[1, 0], [0, ph]])))      # controlled phase
    for i in range(n // 2):
        g.append(("swap", [i, n - 1 - i], [], 0.0))
    return g


def c4_global_layers(n: int = 36, depth: int = 20, seed: int = 36, top: int = 3):
    """configs[3]: brick layers of Haar 1q + CZ; every 4th layer also hits the top `top` qubits with
    Haar-random 2q unitaries so that a sharded run must exchange index bits."""
    rng = np.random.default_rng(seed)
    g = []
    for layer in range(depth):
        for q in range(n):
            g.append(("matrix", [q], [], 0.0, haar_unitary(rng, 2)))
        for q in range(layer % 2, n - 1, 2):
            g.append(("cz", [q, q + 1], [], 0.0))
        if layer % 4 == 3:
            for q in range(n - top, n - 1):
                g.append(("matrix", [q, q + 1], [], 0.0, haar_unitary(rng, 4)))
    return g


def c5_vqe_ansatz(n: int = 4, seed: int = 5):
    """configs[4]: the hardware-efficient ansatz of examples/vqe_lih.py:74-95 -- H on all, RY layer, CNOT ring,
    RY layer; 2n parameters uniform in [0, 2pi)."""
    rng = np.random.default_rng(seed)
    th = rng.uniform(0, 2 * math.pi, size=2 * n)
    g = [("h", [q], [], 0.0) for q in range(n)]
    g += [("ry", [q], [], float(th[q])) for q in range(n)]
    g += [("cnot", [(q + 1) % n], [q], 0.0) for q in range(n)]
    g += [("ry", [q], [], float(th[n + q])) for q in range(n)]
    return g


def random_pauli_strings(n: int, count: int = 64, max_weight: int = 8, seed: int = 5):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        w = int(rng.integers(1, min(max_weight, n) + 1))
        qs = sorted(int(q) for q in rng.choice(n, size=w, replace=False))
        ps = "".join("XYZ"[int(rng.integers(3))] for _ in range(w))
        out.append((ps, qs))
    return out


def hamiltonian_like_terms(n: int, count: int = 64, seed: int = 5):
    """Pauli terms with the structure of a molecular Hamiltonian (examples/vqe_lih.py:66-70 scaled up): single-qubit Z,
    nearest-neighbour ZZ, and for a few qubit pairs the four strings XX, YY, XY, YX (with Z-strings in between) that share one
    x-mask.  -> [(paulis, qubits)]"""
    rng = np.random.default_rng(seed)
    terms = [("Z", [q]) for q in range(min(n, count // 4))]
    terms += [("ZZ", [q, q + 1]) for q in range(min(n - 1, count // 4))]
    while len(terms) < count:
        a, b = sorted(int(x) for x in rng.permutation(n)[:2])
        mid = [q for q in range(a + 1, b)][:2]
        for pa, pb in (("X", "X"), ("Y", "Y"), ("X", "Y"), ("Y", "X")):
            if len(terms) < count:
                terms.append((pa + "Z" * len(mid) + pb, [a] + mid + [b]))
    return terms


def count_gates(gates) -> int:
    return len(gates)
