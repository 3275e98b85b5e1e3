"""The two pybind11 modules of the reference's Python boundary build, import and expose the reference's names (no GPU)."""
import os
import sys

import pytest

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rocquantum_b200", "lib")


@pytest.fixture(scope="module", autouse=True)
def _bindings():
    from rocquantum_b200 import build
    build.build_bindings()
    if LIB not in sys.path:
        sys.path.insert(0, LIB)


def test_rocquantum_bind_surface():
    import rocquantum_bind as rb                                     # bindings.cpp:14-106
    assert rb.QSim is rb.QuantumSimulator
    for name in ("reset", "apply_gate", "apply_matrix", "get_statevector", "measure", "num_qubits", "ApplyGate", "Execute", "GetStateVector"):
        assert hasattr(rb.QuantumSimulator, name), name
    assert hasattr(rb, "MLIRCompiler")
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            rb.QuantumSimulator(2)                                   # fails loudly: no CPU fallback
    with pytest.raises(ValueError):
        rb.QuantumSimulator(0)                                       # std::invalid_argument (simulator.cpp:64-66)


def test_rocq_hip_backend_surface():
    import _rocq_hip_backend as b                                    # python/rocq/bindings.cpp:142-494
    names = ["rocqStatus", "DeviceBuffer", "RocsvHandle", "allocate_state_internal", "initialize_state", "allocate_distributed_state",
             "initialize_distributed_state", "apply_x", "apply_y", "apply_z", "apply_h", "apply_s", "apply_t", "apply_sdg", "apply_rx",
             "apply_ry", "apply_rz", "apply_cnot", "apply_cz", "apply_swap", "apply_crx", "apply_cry", "apply_crz", "apply_mcx",
             "apply_cswap", "apply_matrix", "apply_controlled_matrix", "measure", "get_expectation_value_z", "get_expectation_value_x",
             "get_expectation_value_y", "get_expectation_value_pauli_product_z", "get_expectation_pauli_string", "sample",
             "get_state_vector_full", "get_state_vector_slice", "create_device_matrix_from_numpy", "GateOp", "GateFusion"]
    for n in names:
        assert hasattr(b, n), n
    assert int(b.rocqStatus.SUCCESS) == 0 and int(b.rocqStatus.NOT_IMPLEMENTED) == 5 and b.SUCCESS == b.rocqStatus.SUCCESS
    op = b.GateOp(); op.name = "H"; op.targets = [0]; op.controls = []; op.params = []
    assert op.name == "H"
