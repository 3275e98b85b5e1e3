// include/rocqCompiler/QuantumBackend.h -- interface of the reference's rocqCompiler/QuantumBackend.h:12-33.
#ifndef QUANTUM_BACKEND_H
#define QUANTUM_BACKEND_H

#include <complex>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace rocq {

class QuantumBackend {
public:
    virtual ~QuantumBackend() = default;
    virtual void initialize(unsigned num_qubits) = 0;
    virtual void apply_gate(const std::string& gate_name, const std::vector<unsigned>& targets) = 0;
    virtual void apply_parametrized_gate(const std::string& gate_name, double parameter, const std::vector<unsigned>& targets) = 0;
    virtual std::vector<std::complex<double>> get_state_vector() = 0;
    virtual void destroy() = 0;
};

// "hip_statevec" is the only backend (HipStateVecBackend.cpp:248-253); anything else throws std::invalid_argument.
std::unique_ptr<QuantumBackend> create_backend(const std::string& backend_name);

}  // namespace rocq

#endif
