// include/rocqCompiler/QuantumBackend.h -- the abstract backend interface the reference's compiler layer programs against
// (reference: rocqCompiler/QuantumBackend.h:12-33).  Only the class surface is shared with the reference -- same names,
// argument order and virtual-function order, so reference-side callers recompile against this header unchanged; the type
// aliases below are spelling only (they name the very types of the reference's signatures).
#pragma once
#ifndef ROCQ_B200_QUANTUM_BACKEND_H
#define ROCQ_B200_QUANTUM_BACKEND_H

#include <complex>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace rocq {

using QubitList = std::vector<unsigned>;                  // qubit indices, little-endian (index bit q <-> qubit q)
using AmplitudeVector = std::vector<std::complex<double>>;

class QuantumBackend {
public:
    virtual ~QuantumBackend() = default;

    /* allocate |0...0> on `num_qubits` qubits; throws std::runtime_error when the device cannot hold it */
    virtual void initialize(unsigned num_qubits) = 0;
    /* named gate without angle: h x/paulix y/pauliy z/pauliz s sdg/sdag t cx/cnot cz swap mcx/ccx/toffoli cswap/fredkin */
    virtual void apply_gate(const std::string& gate_name, const QubitList& targets) = 0;
    /* named gate with one angle: rx ry rz crx cry crz */
    virtual void apply_parametrized_gate(const std::string& gate_name, double parameter, const QubitList& targets) = 0;
    /* 2^n amplitudes, always complex128 on the host side */
    virtual AmplitudeVector get_state_vector() = 0;
    /* release device memory; the object may be initialize()d again */
    virtual void destroy() = 0;
};

// "hip_statevec" is the only backend (reference HipStateVecBackend.cpp:248-253); anything else throws std::invalid_argument.
std::unique_ptr<QuantumBackend> create_backend(const std::string& backend_name);

}  // namespace rocq

#endif  // ROCQ_B200_QUANTUM_BACKEND_H
