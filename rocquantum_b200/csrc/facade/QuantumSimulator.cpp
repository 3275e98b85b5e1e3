// rocquantum::QuantumSimulator on the B200 engine (complex128 build).  Interface: include/rocquantum/QuantumSimulator.h,
// which mirrors the reference's include/rocquantum/QuantumSimulator.h:11-40 / rocquantum/src/simulator.cpp:60-218.
// Differences from the reference are deliberate and documented in SURVEY.md section 2.2 / appendix A:
//   * RZ(theta) = diag(e^{-i theta/2}, e^{+i theta/2}) -- the reference's own Qiskit test (test_backend.py:36-44) and
//     rocsvApplyRz; simulator.cpp:35-37 has the conjugate;
//   * CNOT is the textbook gate for any control/target order (kernels.hip.cpp:41-46 is wrong for control < target);
//   * gate names match case-insensitively, including the S/T/CZ/SWAP names the plugins send;
//   * no per-gate device malloc / synchronise (simulator.cpp:126-144): gates are enqueued on the handle's stream.
#define ROCQ_PRECISION_DOUBLE 1
#include "rocquantum/QuantumSimulator.h"

#include <algorithm>
#include <cctype>
#include <stdexcept>

#include "hipStateVec.h"

namespace rocquantum {

namespace {
std::string upper(const std::string& s) {
    std::string u(s.size(), '\0');
    std::transform(s.begin(), s.end(), u.begin(), [](unsigned char c) { return (char)std::toupper(c); });
    return u;
}
void check(rocqStatus_t st, const char* what) {
    if (st == ROCQ_STATUS_INVALID_VALUE) throw std::out_of_range(std::string("QuantumSimulator: invalid argument in ") + what);
    if (st != ROCQ_STATUS_SUCCESS) throw std::runtime_error(std::string("QuantumSimulator: ") + what + " failed (status " + std::to_string((int)st) + ")");
}
}  // namespace

QuantumSimulator::QuantumSimulator(unsigned num_qubits) : num_qubits_(num_qubits), handle_(nullptr), device_state_(nullptr) {
    if (num_qubits_ == 0) throw std::invalid_argument("QuantumSimulator requires at least one qubit.");      // simulator.cpp:64-66
    if (rocsvCreate(&handle_) != ROCQ_STATUS_SUCCESS) throw std::runtime_error("QuantumSimulator: no usable CUDA device (no CPU fallback)");
    rocComplex* d = nullptr;
    const rocqStatus_t st = rocsvAllocateState(handle_, num_qubits_, &d, 1);
    if (st != ROCQ_STATUS_SUCCESS) { rocsvDestroy(handle_); handle_ = nullptr; throw std::runtime_error("QuantumSimulator: state allocation failed"); }
    device_state_ = d;
    reset();
}

QuantumSimulator::~QuantumSimulator() {
    if (handle_) rocsvDestroy(handle_);
}

void QuantumSimulator::reset() { check(rocsvInitializeState(handle_, (rocComplex*)device_state_, num_qubits_), "reset"); }

void QuantumSimulator::apply_gate(const std::string& gate_name, const std::vector<unsigned>& targets, const std::vector<double>& params) {
    if (targets.empty()) throw std::invalid_argument("apply_gate requires at least one target qubit.");
    const std::string g = upper(gate_name);
    for (unsigned t : targets) ensure_valid_qubit(t);
    rocComplex* d = (rocComplex*)device_state_;
    const unsigned n = num_qubits_, t0 = targets[0];
    auto two = [&](const char* nm) { if (targets.size() != 2) throw std::runtime_error(std::string(nm) + " requires 2 target qubits."); };
    auto angle = [&]() { if (params.empty()) throw std::runtime_error("Rotation gate requires an angle parameter."); return params[0]; };
    if (g == "CNOT" || g == "CX") { two("CNOT"); check(rocsvApplyCNOT(handle_, d, n, targets[0], targets[1]), "CNOT"); }
    else if (g == "CZ") { two("CZ"); check(rocsvApplyCZ(handle_, d, n, targets[0], targets[1]), "CZ"); }
    else if (g == "SWAP") { two("SWAP"); check(rocsvApplySWAP(handle_, d, n, targets[0], targets[1]), "SWAP"); }
    else if (targets.size() != 1) throw std::runtime_error("Only single-qubit matrices are supported.");
    else if (g == "H" || g == "HADAMARD") check(rocsvApplyH(handle_, d, n, t0), "H");
    else if (g == "X" || g == "PAULIX") check(rocsvApplyX(handle_, d, n, t0), "X");
    else if (g == "Y" || g == "PAULIY") check(rocsvApplyY(handle_, d, n, t0), "Y");
    else if (g == "Z" || g == "PAULIZ") check(rocsvApplyZ(handle_, d, n, t0), "Z");
    else if (g == "I" || g == "IDENTITY") { /* nothing to do */ }
    else if (g == "S") check(rocsvApplyS(handle_, d, n, t0), "S");
    else if (g == "SDG" || g == "SDAG") check(rocsvApplySdg(handle_, d, n, t0), "Sdg");
    else if (g == "T") check(rocsvApplyT(handle_, d, n, t0), "T");
    else if (g == "RX") check(rocsvApplyRx(handle_, d, n, t0, angle()), "RX");
    else if (g == "RY") check(rocsvApplyRy(handle_, d, n, t0, angle()), "RY");
    else if (g == "RZ") check(rocsvApplyRz(handle_, d, n, t0, angle()), "RZ");
    else throw std::runtime_error("Gate '" + gate_name + "' is not supported.");
}

void QuantumSimulator::apply_matrix(const std::vector<std::complex<double>>& matrix, const std::vector<unsigned>& targets) {
    const size_t k = targets.size();
    if (k == 0 || k > 8) throw std::runtime_error("apply_matrix supports 1 to 8 target qubits.");
    const size_t D = (size_t)1 << k;
    if (matrix.size() != D * D) throw std::runtime_error(k == 1 ? "Matrix must have 4 elements." : "Matrix size does not match the targets.");
    for (unsigned t : targets) ensure_valid_qubit(t);
    // row-major in (bindings.cpp:48-55: u00,u01,u10,u11), column-major interleaved doubles out
    std::vector<double> cm(2 * D * D);
    for (size_t i = 0; i < D; ++i)
        for (size_t j = 0; j < D; ++j) { cm[2 * (i + j * D)] = matrix[i * D + j].real(); cm[2 * (i + j * D) + 1] = matrix[i * D + j].imag(); }
    rocsvxGateOp op{};
    op.kind = ROCSVX_MATRIX;
    op.numTargets = (uint32_t)k;
    for (size_t b = 0; b < k; ++b) op.targets[b] = targets[b];
    op.matrix = cm.data();
    check(rocsvxApplyCircuit(handle_, (rocComplex*)device_state_, num_qubits_, &op, 1), "apply_matrix");
}

std::vector<std::complex<double>> QuantumSimulator::get_statevector() const {
    std::vector<std::complex<double>> out((size_t)1 << num_qubits_);
    check(rocsvGetStateVectorFull(handle_, (rocComplex*)device_state_, reinterpret_cast<rocComplex*>(out.data())), "get_statevector");
    return out;
}

std::vector<long long> QuantumSimulator::measure(const std::vector<unsigned>& qubits, int shots) {
    for (unsigned q : qubits) ensure_valid_qubit(q);
    if (shots <= 0) return {};
    if (num_qubits_ > 63) throw std::runtime_error("measure supports at most 63 qubits.");
    std::vector<unsigned> all(num_qubits_);
    for (unsigned q = 0; q < num_qubits_; ++q) all[q] = q;
    std::vector<uint64_t> raw((size_t)shots);
    check(rocsvSample(handle_, (rocComplex*)device_state_, num_qubits_, all.data(), num_qubits_, (unsigned)shots, raw.data()), "measure");
    return std::vector<long long>(raw.begin(), raw.end());
}

unsigned QuantumSimulator::num_qubits() const noexcept { return num_qubits_; }
void QuantumSimulator::set_seed(std::uint64_t seed) { check(rocsvxSetSeed(handle_, seed), "set_seed"); }

void QuantumSimulator::ApplyGate(const std::string& gate_name, int target_qubit) { apply_gate(gate_name, {(unsigned)target_qubit}, {}); }
void QuantumSimulator::ApplyGate(const std::string& gate_name, int control_qubit, int target_qubit) {
    apply_gate(gate_name, {(unsigned)control_qubit, (unsigned)target_qubit}, {});
}
void QuantumSimulator::ApplyGate(const std::vector<std::complex<double>>& gate_matrix, int target_qubit) { apply_matrix(gate_matrix, {(unsigned)target_qubit}); }
void QuantumSimulator::Execute() { check(rocsvxSynchronize(handle_), "Execute"); }
std::vector<std::complex<double>> QuantumSimulator::GetStateVector() const { return get_statevector(); }

void QuantumSimulator::ensure_valid_qubit(unsigned qubit) const {
    if (qubit >= num_qubits_) throw std::out_of_range("Qubit index out of bounds for simulator instance.");   // simulator.cpp:208-212
}

}  // namespace rocquantum
