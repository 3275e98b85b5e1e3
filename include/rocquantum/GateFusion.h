// include/rocquantum/GateFusion.h -- interface of the reference's rocquantum/include/rocquantum/GateFusion.h:12-33.
// processQueue hands the WHOLE queue to the engine's circuit-level entry point (rocsvxApplyCircuit), which fuses
// algebraically and cuts the queue into HBM sweeps; the reference fuses one gate around a CNOT and silently drops
// every other gate (GateFusion.cpp:152-153).
#ifndef GATEFUSION_H
#define GATEFUSION_H

#include <string>
#include <vector>

#include "hipStateVec.h"

namespace rocquantum {

struct GateOp {
    std::string name;                 // X Y Z H S SDG T RX RY RZ CNOT/CX CZ SWAP CRX CRY CRZ MCX/CCX CSWAP (case-insensitive)
    std::vector<unsigned> targets;
    std::vector<unsigned> controls;
    std::vector<double> params;
};

class GateFusion {
public:
    GateFusion(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits);
    rocqStatus_t processQueue(const std::vector<GateOp>& queue);

private:
    rocsvHandle_t handle_;
    rocComplex* d_state_;
    unsigned numQubits_;
};

}  // namespace rocquantum

#endif
