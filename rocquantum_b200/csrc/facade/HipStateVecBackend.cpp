// rocq::HipStateVecBackend on the B200 engine.  Interface and gate-name tables: the reference's
// rocqCompiler/HipStateVecBackend.cpp:65-149 (names, aliases, arities), :153-253 (lifecycle, errors, factory),
// :259-349 (dispatch).  A pure client of the C ABI, precision = that of the library it is linked with.
#include "rocqCompiler/HipStateVecBackend.h"

#include <algorithm>
#include <cctype>
#include <functional>
#include <unordered_map>

namespace rocq {

namespace {

struct Spec {
    size_t min_targets, max_targets;                 // max 0 = unbounded
    std::function<rocqStatus_t(rocsvHandle_t, rocComplex*, unsigned, const std::vector<unsigned>&, double)> call;
};

std::string lower(const std::string& s) {
    std::string o(s.size(), '\0');
    std::transform(s.begin(), s.end(), o.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    return o;
}

#define RQ_G1(fn) [](rocsvHandle_t h, rocComplex* d, unsigned n, const std::vector<unsigned>& t, double) { return fn(h, d, n, t[0]); }
#define RQ_G2(fn) [](rocsvHandle_t h, rocComplex* d, unsigned n, const std::vector<unsigned>& t, double) { return fn(h, d, n, t[0], t[1]); }
#define RQ_R1(fn) [](rocsvHandle_t h, rocComplex* d, unsigned n, const std::vector<unsigned>& t, double a) { return fn(h, d, n, t[0], a); }
#define RQ_R2(fn) [](rocsvHandle_t h, rocComplex* d, unsigned n, const std::vector<unsigned>& t, double a) { return fn(h, d, n, t[0], t[1], a); }

const std::unordered_map<std::string, Spec>& plain_table() {
    static const std::unordered_map<std::string, Spec> t = [] {
        std::unordered_map<std::string, Spec> m;
        const Spec h{1, 1, RQ_G1(rocsvApplyH)}, x{1, 1, RQ_G1(rocsvApplyX)}, y{1, 1, RQ_G1(rocsvApplyY)}, z{1, 1, RQ_G1(rocsvApplyZ)};
        const Spec s{1, 1, RQ_G1(rocsvApplyS)}, sdg{1, 1, RQ_G1(rocsvApplySdg)}, tg{1, 1, RQ_G1(rocsvApplyT)};
        const Spec cx{2, 2, RQ_G2(rocsvApplyCNOT)}, cz{2, 2, RQ_G2(rocsvApplyCZ)}, sw{2, 2, RQ_G2(rocsvApplySWAP)};
        // mcx/ccx/toffoli: the LAST entry is the target, the rest are controls (HipStateVecBackend.cpp:299-312)
        const Spec mcx{2, 0, [](rocsvHandle_t hd, rocComplex* d, unsigned n, const std::vector<unsigned>& q, double) {
                           return rocsvApplyMultiControlledX(hd, d, n, q.data(), (unsigned)q.size() - 1, q.back());
                       }};
        const Spec csw{3, 3, [](rocsvHandle_t hd, rocComplex* d, unsigned n, const std::vector<unsigned>& q, double) {
                           return rocsvApplyCSWAP(hd, d, n, q[0], q[1], q[2]);
                       }};
        m = {{"h", h}, {"x", x}, {"paulix", x}, {"y", y}, {"pauliy", y}, {"z", z}, {"pauliz", z}, {"s", s}, {"sdg", sdg}, {"sdag", sdg},
             {"t", tg}, {"cx", cx}, {"cnot", cx}, {"cz", cz}, {"swap", sw}, {"mcx", mcx}, {"ccx", mcx}, {"toffoli", mcx},
             {"cswap", csw}, {"fredkin", csw}};
        return m;
    }();
    return t;
}
const std::unordered_map<std::string, Spec>& param_table() {
    static const std::unordered_map<std::string, Spec> t = {
        {"rx", {1, 1, RQ_R1(rocsvApplyRx)}}, {"ry", {1, 1, RQ_R1(rocsvApplyRy)}}, {"rz", {1, 1, RQ_R1(rocsvApplyRz)}},
        {"crx", {2, 2, RQ_R2(rocsvApplyCRX)}}, {"cry", {2, 2, RQ_R2(rocsvApplyCRY)}}, {"crz", {2, 2, RQ_R2(rocsvApplyCRZ)}}};
    return t;
}

void check_status(rocqStatus_t st, const std::string& what) {        // HipStateVecBackend.cpp:24-29
    if (st != ROCQ_STATUS_SUCCESS)
        throw std::runtime_error("hipStateVec error during " + what + " (status " + std::to_string((int)st) + ")");
}
void check_arity(const std::vector<unsigned>& t, const Spec& s, const std::string& name) {
    if (t.size() < s.min_targets)
        throw std::invalid_argument("Gate '" + name + "' expected at least " + std::to_string(s.min_targets) + " target qubits but received " +
                                    std::to_string(t.size()) + ".");
    if (s.max_targets != 0 && t.size() > s.max_targets)
        throw std::invalid_argument("Gate '" + name + "' expected at most " + std::to_string(s.max_targets) + " target qubits but received " +
                                    std::to_string(t.size()) + ".");
}

}  // namespace

HipStateVecBackend::HipStateVecBackend() : sim_handle(nullptr), num_qubits(0), device_state(nullptr), is_initialized(false) {
    if (rocsvCreate(&sim_handle) != ROCQ_STATUS_SUCCESS) throw std::runtime_error("Failed to create hipStateVec handle.");
}
HipStateVecBackend::~HipStateVecBackend() {
    if (sim_handle) { destroy(); rocsvDestroy(sim_handle); }
}
void HipStateVecBackend::initialize(unsigned n_qubits) {
    if (n_qubits == 0) throw std::invalid_argument("hipStateVec backend requires at least one qubit.");
    num_qubits = n_qubits;
    rocComplex* buffer = nullptr;
    check_status(rocsvAllocateState(sim_handle, num_qubits, &buffer, 1), "state allocation");
    device_state = buffer;
    check_status(rocsvInitializeState(sim_handle, device_state, num_qubits), "state initialisation");
    is_initialized = true;
}
void HipStateVecBackend::apply_gate(const std::string& gate_name, const std::vector<unsigned>& targets) {
    if (!is_initialized) throw std::runtime_error("Backend not initialized.");
    const auto it = plain_table().find(lower(gate_name));
    if (it == plain_table().end()) throw std::runtime_error("Unknown gate: " + gate_name);
    check_arity(targets, it->second, gate_name);
    check_status(it->second.call(sim_handle, device_state, num_qubits, targets, 0.0), "apply " + gate_name);
}
void HipStateVecBackend::apply_parametrized_gate(const std::string& gate_name, double parameter, const std::vector<unsigned>& targets) {
    if (!is_initialized) throw std::runtime_error("Backend not initialized.");
    const auto it = param_table().find(lower(gate_name));
    if (it == param_table().end()) throw std::runtime_error("Unknown parametrised gate: " + gate_name);
    check_arity(targets, it->second, gate_name);
    check_status(it->second.call(sim_handle, device_state, num_qubits, targets, parameter), "apply " + gate_name);
}
std::vector<std::complex<double>> HipStateVecBackend::get_state_vector() {
    if (!is_initialized) throw std::runtime_error("Backend not initialized.");
    const size_t N = (size_t)1 << num_qubits;
    std::vector<rocComplex> raw(N);
    check_status(rocsvGetStateVectorFull(sim_handle, device_state, raw.data()), "fetch state vector");
    std::vector<std::complex<double>> out(N);
    for (size_t i = 0; i < N; ++i) out[i] = {(double)raw[i].x, (double)raw[i].y};
    return out;
}
void HipStateVecBackend::destroy() {
    if (sim_handle && device_state) { rocsvFreeState(sim_handle); device_state = nullptr; }
    is_initialized = false;
    num_qubits = 0;
}

std::unique_ptr<QuantumBackend> create_backend(const std::string& backend_name) {
    if (backend_name == "hip_statevec") return std::make_unique<HipStateVecBackend>();
    throw std::invalid_argument("Unknown backend: " + backend_name);
}

}  // namespace rocq
