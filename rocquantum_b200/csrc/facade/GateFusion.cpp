// rocquantum::GateFusion on the B200 engine (interface: include/rocquantum/GateFusion.h; reference:
// rocquantum/src/hipStateVec/GateFusion.cpp:89-156).  The queue is translated to rocsvxGateOp records and submitted
// in ONE call; the engine performs the fusion.  Unknown gate names are an error, not a silent drop (:152-153).
#include "rocquantum/GateFusion.h"

#include <algorithm>
#include <cctype>

namespace rocquantum {

GateFusion::GateFusion(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits) : handle_(handle), d_state_(d_state), numQubits_(numQubits) {}

rocqStatus_t GateFusion::processQueue(const std::vector<GateOp>& queue) {
    std::vector<rocsvxGateOp> ops;
    ops.reserve(queue.size());
    for (const GateOp& g : queue) {
        std::string nm(g.name.size(), '\0');
        std::transform(g.name.begin(), g.name.end(), nm.begin(), [](unsigned char c) { return (char)std::toupper(c); });
        rocsvxGateOp o{};
        uint64_t cm = 0;
        for (unsigned c : g.controls) { if (c >= 64) return ROCQ_STATUS_INVALID_VALUE; cm |= 1ull << c; }
        std::vector<unsigned> t = g.targets;
        auto kind = [&](int k, size_t nt, size_t nc, bool angle) -> bool {
            if (t.size() != nt || (nc != (size_t)-1 && g.controls.size() != nc) || (angle && g.params.empty())) return false;
            o.kind = k; o.numTargets = (uint32_t)nt;
            for (size_t i = 0; i < nt; ++i) o.targets[i] = t[i];
            o.controlMask = cm;
            o.theta = angle ? g.params[0] : 0.0;
            return true;
        };
        bool ok = false;
        if (nm == "H") ok = kind(ROCSVX_H, 1, 0, false);
        else if (nm == "X") ok = kind(ROCSVX_X, 1, 0, false);
        else if (nm == "Y") ok = kind(ROCSVX_Y, 1, 0, false);
        else if (nm == "Z") ok = kind(ROCSVX_Z, 1, 0, false);
        else if (nm == "S") ok = kind(ROCSVX_S, 1, 0, false);
        else if (nm == "SDG") ok = kind(ROCSVX_SDG, 1, 0, false);
        else if (nm == "T") ok = kind(ROCSVX_T, 1, 0, false);
        else if (nm == "RX") ok = kind(ROCSVX_RX, 1, 0, true);
        else if (nm == "RY") ok = kind(ROCSVX_RY, 1, 0, true);
        else if (nm == "RZ") ok = kind(ROCSVX_RZ, 1, 0, true);
        else if (nm == "CNOT" || nm == "CX") ok = kind(ROCSVX_CNOT, 1, 1, false);
        else if (nm == "CZ") {                          // accepted as (control, target) or as two targets
            if (t.size() == 1 && g.controls.size() == 1) { t.insert(t.begin(), g.controls[0]); cm = 0; }
            ok = kind(ROCSVX_CZ, 2, (size_t)-1, false);
            o.controlMask = 0;
        }
        else if (nm == "SWAP") ok = kind(ROCSVX_SWAP, 2, 0, false);
        else if (nm == "CRX") ok = kind(ROCSVX_CRX, 1, 1, true);
        else if (nm == "CRY") ok = kind(ROCSVX_CRY, 1, 1, true);
        else if (nm == "CRZ") ok = kind(ROCSVX_CRZ, 1, 1, true);
        else if (nm == "MCX" || nm == "CCX" || nm == "TOFFOLI") ok = kind(ROCSVX_MCX, 1, (size_t)-1, false) && cm != 0;
        else if (nm == "CSWAP" || nm == "FREDKIN") ok = kind(ROCSVX_CSWAP, 2, 1, false);
        if (!ok) return ROCQ_STATUS_INVALID_VALUE;
        ops.push_back(o);
    }
    return rocsvxApplyCircuit(handle_, d_state_, numQubits_, ops.data(), ops.size());
}

}  // namespace rocquantum
