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

class QuantumSimulator {
public:
    explicit QuantumSimulator(unsigned num_qubits);
    ~QuantumSimulator();
    QuantumSimulator(const QuantumSimulator&) = delete;
    QuantumSimulator& operator=(const QuantumSimulator&) = delete;

    void reset();
    // Names, case-insensitive: H/Hadamard, X/PauliX, Y/PauliY, Z/PauliZ, I/Identity, S, Sdg, T, RX, RY, RZ (params[0]),
    // CNOT/CX, CZ, SWAP (targets = {control, target} / {a, b}).  The reference's table (simulator.cpp:41-48) only ever
    // matches "H", the rotations and CNOT/CX because of its upper-casing; the plugins send the rest.
    void apply_gate(const std::string& gate_name, const std::vector<unsigned>& targets, const std::vector<double>& params = {});
    // matrix: 2^k x 2^k, ROW-major like the reference binding (bindings.cpp:48-55); k = targets.size() (reference: k = 1 only)
    void apply_matrix(const std::vector<std::complex<double>>& matrix, const std::vector<unsigned>& targets);
    std::vector<std::complex<double>> get_statevector() const;
    // `shots` full-register basis indices, whatever `qubits` says -- the reference's behaviour (simulator.cpp:153-184),
    // which its callers rely on (tests/test_bindings.py:54-68).  Reproducible: Philox stream seeded by set_seed.
    std::vector<long long> measure(const std::vector<unsigned>& qubits, int shots);
    unsigned num_qubits() const noexcept;
    void set_seed(std::uint64_t seed);

    // Legacy API (QuantumSimulator.h:28-33)
    void ApplyGate(const std::string& gate_name, int target_qubit);
    void ApplyGate(const std::string& gate_name, int control_qubit, int target_qubit);
    void ApplyGate(const std::vector<std::complex<double>>& gate_matrix, int target_qubit);
    void Execute();
    std::vector<std::complex<double>> GetStateVector() const;

private:
    void ensure_valid_qubit(unsigned qubit) const;
    unsigned num_qubits_;
    rocsvInternalHandle* handle_;
    void* device_state_;
};

using QSim = QuantumSimulator;

}  // namespace rocquantum
