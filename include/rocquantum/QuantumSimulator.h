// include/rocquantum/QuantumSimulator.h -- same class surface as the reference's
// include/rocquantum/QuantumSimulator.h:11-40 (complex128 simulator behind the `rocquantum_bind` pybind module),
// re-hosted on the B200 engine: every method is a handful of rocsv* calls on libhipStateVec_f64.so.
#pragma once

#include <complex>
#include <cstdint>
#include <string>
#include <vector>

struct rocsvInternalHandle;

namespace rocquantum {

using Qubits = std::vector<unsigned>;                      // spelling only: the reference's parameter types
using Amplitudes = std::vector<std::complex<double>>;

class QuantumSimulator {
public:
    explicit QuantumSimulator(unsigned num_qubits);
    ~QuantumSimulator();
    QuantumSimulator(const QuantumSimulator&) = delete;
    QuantumSimulator& operator=(const QuantumSimulator&) = delete;

    void reset();                                          // back to |0...0>
    // Names, case-insensitive: H/Hadamard, X/PauliX, Y/PauliY, Z/PauliZ, I/Identity, S, Sdg, T, RX, RY, RZ (params[0]),
    // CNOT/CX, CZ, SWAP (targets = {control, target} / {a, b}).  The reference's table (simulator.cpp:41-48) only ever
    // matches "H", the rotations and CNOT/CX because of its upper-casing; the plugins send the rest.
    void apply_gate(const std::string& gate_name, const Qubits& targets, const std::vector<double>& params = {});
    // matrix: 2^k x 2^k, ROW-major like the reference binding (bindings.cpp:48-55); k = targets.size() (reference: k = 1 only)
    void apply_matrix(const Amplitudes& matrix, const Qubits& targets);
    Amplitudes get_statevector() const;
    // `shots` full-register basis indices, whatever `qubits` says -- the reference's behaviour (simulator.cpp:153-184),
    // which its callers rely on (tests/test_bindings.py:54-68).  Reproducible: Philox stream seeded by set_seed.
    std::vector<long long> measure(const Qubits& qubits, int shots);
    unsigned num_qubits() const noexcept;
    void set_seed(std::uint64_t seed);                     // extension: the reference seeds from random_device

    // Legacy spellings kept by the reference for its older QSim bindings (QuantumSimulator.h:28-33)
    void ApplyGate(const std::string& gate_name, int target_qubit);
    void ApplyGate(const std::string& gate_name, int control_qubit, int target_qubit);
    void ApplyGate(const Amplitudes& gate_matrix, int target_qubit);
    void Execute();                                        // waits for the handle's stream
    Amplitudes GetStateVector() const;

private:
    void ensure_valid_qubit(unsigned qubit) const;
    unsigned num_qubits_;
    rocsvInternalHandle* handle_;                          // engine handle (libhipStateVec_f64.so)
    void* device_state_;                                   // owned by the handle
};

using QSim = QuantumSimulator;

}  // namespace rocquantum
