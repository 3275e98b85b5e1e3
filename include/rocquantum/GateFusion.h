// include/rocquantum/GateFusion.h -- the gate-queue entry of the reference (rocquantum/include/rocquantum/GateFusion.h:12-33),
// same class surface.  processQueue hands the WHOLE queue to the engine's circuit-level entry point (rocsvxApplyCircuit),
// which fuses algebraically and cuts the queue into HBM sweeps; the reference fuses one gate around a CNOT and silently
// drops every other gate (reference GateFusion.cpp:152-153).
#pragma once
#ifndef ROCQ_B200_GATEFUSION_H
#define ROCQ_B200_GATEFUSION_H

#include <string>
#include <vector>

#include "hipStateVec.h"

namespace rocquantum {

// One queued gate.  name (case-insensitive): X Y Z H S SDG T RX RY RZ CNOT/CX CZ SWAP CRX CRY CRZ MCX/CCX CSWAP;
// params[0] is the angle of a rotation; control qubits go in `controls` (CZ also accepts its two qubits as targets).
struct GateOp {
    std::string name;
    std::vector<unsigned> targets, controls;
    std::vector<double> params;
};

class GateFusion {
public:
    GateFusion(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits);
    // INVALID_VALUE for an unknown name or a bad qubit list: nothing of the queue has been applied then
    rocqStatus_t processQueue(const std::vector<GateOp>& queue);

private:
    rocsvHandle_t handle_;             // not owned
    rocComplex* d_state_;              // not owned; NULL = the handle's own state
    unsigned numQubits_;
};

}  // namespace rocquantum

#endif  // ROCQ_B200_GATEFUSION_H
