// include/rocqCompiler/HipStateVecBackend.h -- interface of the reference's rocqCompiler/HipStateVecBackend.h:12-53:
// a name -> rocsv* dispatcher over the C ABI.
#ifndef HIP_STATE_VEC_BACKEND_H
#define HIP_STATE_VEC_BACKEND_H

#include <string>
#include <vector>

#include "QuantumBackend.h"
#include "rocquantum/hipStateVec.h"

namespace rocq {

class HipStateVecBackend : public QuantumBackend {
public:
    HipStateVecBackend();
    ~HipStateVecBackend() override;

    void initialize(unsigned num_qubits) override;
    void apply_gate(const std::string& gate_name, const std::vector<unsigned>& targets) override;
    void apply_parametrized_gate(const std::string& gate_name, double parameter, const std::vector<unsigned>& targets) override;
    std::vector<std::complex<double>> get_state_vector() override;
    void destroy() override;

private:
    rocsvHandle_t sim_handle;
    unsigned num_qubits;
    rocComplex* device_state;
    bool is_initialized;
};

}  // namespace rocq

#endif
