// rocquantum_b200/csrc/gate_convert.h -- rocsvxGateOp (include/hipStateVec.h) -> HostOp.  Pure C++, shared by the engine
// and the host-only test interpreter (tests/hostemu): the named gates' matrices follow hipStateVec.cpp:276-427, 544-595.
#pragma once
#include <cmath>
#include <vector>

#include "../../include/hipStateVec.h"
#include "host_ops.h"

namespace rq {

inline rocqStatus_t convert_ops(unsigned n, const rocsvxGateOp* ops, size_t numOps, std::vector<HostOp>& out) {
    out.reserve(numOps);
    for (size_t i = 0; i < numOps; ++i) {
        const rocsvxGateOp& g = ops[i];
        const unsigned t0 = g.targets[0], t1 = g.targets[1];
        const uint64_t cm = g.controlMask;
        auto okq = [&](unsigned q) { return q < n; };
        if (n < 64 && (cm >> n)) return ROCQ_STATUS_INVALID_VALUE;
        const double c = std::cos(g.theta / 2.0), s = std::sin(g.theta / 2.0), r = 1.0 / std::sqrt(2.0);
        const bool one_ctrl = __builtin_popcountll(cm) == 1;
        switch (g.kind) {
            case ROCSVX_H: case ROCSVX_X: case ROCSVX_Y: case ROCSVX_Z: case ROCSVX_S: case ROCSVX_SDG: case ROCSVX_T:
            case ROCSVX_RX: case ROCSVX_RY: case ROCSVX_RZ:
                if (!okq(t0) || cm) return ROCQ_STATUS_INVALID_VALUE;
                break;
            case ROCSVX_CNOT: case ROCSVX_CRX: case ROCSVX_CRY: case ROCSVX_CRZ:
                if (!okq(t0) || !one_ctrl || ((cm >> t0) & 1ull)) return ROCQ_STATUS_INVALID_VALUE;
                break;
            case ROCSVX_MCX:
                if (!okq(t0) || cm == 0 || ((cm >> t0) & 1ull)) return ROCQ_STATUS_INVALID_VALUE;
                break;
            case ROCSVX_CZ: case ROCSVX_SWAP:
                if (!okq(t0) || !okq(t1) || t0 == t1 || cm) return ROCQ_STATUS_INVALID_VALUE;
                break;
            case ROCSVX_CSWAP:
                if (!okq(t0) || !okq(t1) || t0 == t1 || !one_ctrl || ((cm >> t0) & 1ull) || ((cm >> t1) & 1ull)) return ROCQ_STATUS_INVALID_VALUE;
                break;
            case ROCSVX_MATRIX: break;
            default: return ROCQ_STATUS_INVALID_VALUE;
        }
        switch (g.kind) {
            case ROCSVX_H: out.push_back(rq::make_dense1(t0, r, r, r, -r)); break;
            case ROCSVX_X: out.push_back(rq::make_x(t0)); break;
            case ROCSVX_Y: out.push_back(rq::make_dense1(t0, 0.0, -cd(0.0, 1.0), cd(0.0, 1.0), 0.0)); break;
            case ROCSVX_Z: out.push_back(rq::make_phase(1ull << t0, -1.0)); break;
            case ROCSVX_S: out.push_back(rq::make_phase(1ull << t0, cd(0.0, 1.0))); break;
            case ROCSVX_SDG: out.push_back(rq::make_phase(1ull << t0, -cd(0.0, 1.0))); break;
            case ROCSVX_T: { const double ph = 3.14159265358979323846 / 4.0; out.push_back(rq::make_phase(1ull << t0, cd(std::cos(ph), std::sin(ph)))); break; }
            case ROCSVX_RX: out.push_back(rq::make_dense1(t0, c, cd(0, -s), cd(0, -s), c)); break;
            case ROCSVX_RY: out.push_back(rq::make_dense1(t0, c, -s, s, c)); break;
            case ROCSVX_RZ: out.push_back(rq::make_diag1(t0, cd(c, -s), cd(c, s))); break;
            case ROCSVX_CNOT: case ROCSVX_MCX: out.push_back(rq::make_x(t0, cm)); break;
            case ROCSVX_CZ: out.push_back(rq::make_phase((1ull << t0) | (1ull << t1), -1.0)); break;
            case ROCSVX_SWAP: out.push_back(rq::make_swap(t0, t1)); break;
            case ROCSVX_CRX: out.push_back(rq::make_dense1(t0, c, cd(0, -s), cd(0, -s), c, cm)); break;
            case ROCSVX_CRY: out.push_back(rq::make_dense1(t0, c, -s, s, c, cm)); break;
            case ROCSVX_CRZ: out.push_back(rq::make_diag1(t0, cd(c, -s), cd(c, s), cm)); break;
            case ROCSVX_CSWAP: out.push_back(rq::make_swap(t0, t1, cm)); break;
            case ROCSVX_MATRIX: {
                const unsigned k = g.numTargets;
                if (k == 0 || k > 8 || !g.matrix) return ROCQ_STATUS_INVALID_VALUE;
                uint64_t seen = cm;
                std::vector<unsigned> ts(g.targets, g.targets + k);
                for (unsigned q : ts) { if (!okq(q) || ((seen >> q) & 1ull)) return ROCQ_STATUS_INVALID_VALUE; seen |= 1ull << q; }
                const size_t D = (size_t)1 << k;
                std::vector<cd> m(D * D);
                for (size_t e = 0; e < D * D; ++e) m[e] = cd(g.matrix[2 * e], g.matrix[2 * e + 1]);
                out.push_back(rq::make_matrix(ts, cm, m));
                break;
            }
        }
    }
    return ROCQ_STATUS_SUCCESS;
}

}  // namespace rq
