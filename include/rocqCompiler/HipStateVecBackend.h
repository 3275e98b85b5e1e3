// include/rocqCompiler/HipStateVecBackend.h -- the one concrete QuantumBackend: a name -> rocsv* dispatcher over the C ABI
// (class surface of the reference's rocqCompiler/HipStateVecBackend.h:12-53; the implementation is
// rocquantum_b200/csrc/facade/HipStateVecBackend.cpp and talks to libhipStateVec(_f64).so only through include/hipStateVec.h).
#pragma once
#ifndef ROCQ_B200_HIP_STATE_VEC_BACKEND_H
#define ROCQ_B200_HIP_STATE_VEC_BACKEND_H

#include <string>
#include <vector>

#include "QuantumBackend.h"
#include "rocquantum/hipStateVec.h"

namespace rocq {

class HipStateVecBackend : public QuantumBackend {
public:
    HipStateVecBackend();
    ~HipStateVecBackend() override;

    // QuantumBackend, in its declaration order
    void initialize(unsigned num_qubits) override;
    void apply_gate(const std::string& gate_name, const QubitList& targets) override;
    void apply_parametrized_gate(const std::string& gate_name, double parameter, const QubitList& targets) override;
    AmplitudeVector get_state_vector() override;
    void destroy() override;

private:
    rocsvHandle_t sim_handle;          // engine handle (one CUDA stream)
    unsigned num_qubits;
    rocComplex* device_state;          // owned by the handle; kept to pass back as d_state
    bool is_initialized;
};

}  // namespace rocq

#endif  // ROCQ_B200_HIP_STATE_VEC_BACKEND_H
