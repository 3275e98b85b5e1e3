// rocquantum_b200/csrc/dist.h -- distributed (one process per GPU) state: top log2(P) index bits select
// the rank, global = (rank << n_local) | local (reference layout: swap_kernels.hip:10-22,
// MULTI_GPU_GUIDE.md:19-24).  The engine keeps a logical->physical qubit map; gates on physical
// positions >= n_local are made local by an index-bit exchange over NCCL (dist.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <utility>
#include <vector>

#include "../../include/hipStateVec.h"
#include "dist_plan.h"
#include "host_ops.h"

struct rocsvInternalHandle;
struct rocsvGroup;

namespace rq {

struct Dist {
    bool inited = false;             // rocsvxDistInit done
    int rank = 0, nranks = 1;
    unsigned n_total = 0, n_local = 0, n_global = 0;
    DistPlanner plan;                // logical->physical map + step planner (dist_plan.h)
    rocsvGroup* group = nullptr;     // single-process mode (group.h): the ranks are threads of this process and meet through the group, not NCCL
    void* comm = nullptr;            // ncclComm_t (multi-process mode)
    void* nccl = nullptr;            // dlopen handle
    rq_cplx* staging = nullptr;      // exchange staging (device)
    size_t staging_amps = 0;
    uint64_t* d_gather = nullptr;    // small device buffer for all-gathers
    uint64_t exchanges = 0;          // statistics
    uint64_t exchanged_amps = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;   // event pairs around exchanges not yet folded into stats
    std::vector<void*> peer_state;   // ROCQ_EXCHANGE=p2p: every rank's slice mapped through CUDA IPC ([rank] = own pointer); else empty

    bool active() const { return inited && n_total > 0; }
    unsigned num_local() const { return n_local; }
    unsigned num_total() const { return n_total; }
    uint64_t global_mask() const { return n_global ? (((1ull << n_global) - 1ull) << n_local) : 0ull; }
    void reset_layout() { plan.reset(n_total, n_local); }     // identity logical->physical map (fresh data in canonical order)
    uint64_t high_base() const { return active() ? ((uint64_t)rank << n_local) : 0ull; }

    rocqStatus_t init(rocsvInternalHandle* h, int rank_, int nranks_, const void* id128);
    rocqStatus_t init_in_group(rocsvInternalHandle* h, int rank_, rocsvGroup* g);
    rocqStatus_t allocate(rocsvInternalHandle* h, unsigned total_qubits);
    rocqStatus_t initialize(rocsvInternalHandle* h);
    void shutdown(rocsvInternalHandle* h);

    // rewrite op from logical to physical positions, exchanging slices first if a non-diagonal target is global
    rocqStatus_t localize(rocsvInternalHandle* h, HostOp& op);
    rocqStatus_t run_circuit(rocsvInternalHandle* h, std::vector<HostOp>& ops);
    rocqStatus_t swap_index_bits(rocsvInternalHandle* h, unsigned q1, unsigned q2);
    rocqStatus_t canonicalize(rocsvInternalHandle* h);
    rocqStatus_t allreduce_sum(rocsvInternalHandle* h, double* v, unsigned count);
    rocqStatus_t pauli_expect(rocsvInternalHandle* h, uint64_t xm, uint64_t zm, unsigned ny, double* result);
    rocqStatus_t measure(rocsvInternalHandle* h, unsigned q, int* outcome, double* probability);
    rocqStatus_t sample(rocsvInternalHandle* h, const unsigned* measured, unsigned nm, unsigned shots, uint64_t* out);
};

// pure host: segments of a global<->local index-bit exchange (see include/hipStateVec.h)
size_t plan_exchange(unsigned n_local, int nranks, int rank, const unsigned* local_bits, const unsigned* global_bits,
                     unsigned npairs, rocsvxExchangeSeg* segs, size_t maxsegs);
// the same exchange as in-place half-swaps against peer memory (recvOffset = offset in the peer's slice)
size_t plan_peer_swap(unsigned n_local, int nranks, int rank, const unsigned* local_bits, const unsigned* global_bits,
                      unsigned npairs, rocsvxExchangeSeg* segs, size_t maxsegs);

}  // namespace rq
