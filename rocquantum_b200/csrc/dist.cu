// rocquantum_b200/csrc/dist.cu -- distributed state over NCCL (one process per GPU).
// See dist.h.  The reference documents this path (MULTI_GPU_GUIDE.md, hipStateVec.h:84-137) but defines
// none of it; its packing kernels (swap_kernels.hip:46-89) use atomic cursors whose arrival order is
// nondeterministic.  Here a k-bit global<->local exchange moves, per peer, block-contiguous runs that land
// directly in their final position -- no counts, no packing pass.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

#include "dist.h"

namespace rq {

// Swapping global bits G[i] (absolute positions >= n_local) with local bits L[i]: this rank keeps the
// amplitudes whose local pattern on L equals its own rank pattern on G; for every other pattern b it sends
// {x : x_L = b} to the peer whose G-pattern is b and receives that peer's {x : x_L = c} (c = own pattern)
// into the same positions.  Runs are contiguous over the local bits below min(L).
size_t plan_exchange(unsigned n_local, int nranks, int rank, const unsigned* local_bits, const unsigned* global_bits,
                     unsigned npairs, rocsvxExchangeSeg* segs, size_t maxsegs) {
    (void)nranks;
    uint64_t lmask = 0;
    unsigned minl = n_local;
    for (unsigned i = 0; i < npairs; ++i) { lmask |= 1ull << local_bits[i]; if (local_bits[i] < minl) minl = local_bits[i]; }
    unsigned own = 0;                                   // own pattern: bit i = rank's value on global_bits[i]
    for (unsigned i = 0; i < npairs; ++i) own |= (unsigned)(((uint64_t)rank >> (global_bits[i] - n_local)) & 1ull) << i;
    const uint64_t run = 1ull << minl;
    // free local bits above minl that are not exchanged enumerate the runs of one pattern
    std::vector<unsigned> freebits;
    for (unsigned p = minl; p < n_local; ++p) if (!((lmask >> p) & 1ull)) freebits.push_back(p);
    const uint64_t nruns = 1ull << freebits.size();
    size_t count = 0;
    for (unsigned b = 0; b < (1u << npairs); ++b) {
        if (b == own) continue;
        int peer = rank;
        uint64_t pat = 0;
        for (unsigned i = 0; i < npairs; ++i) {
            const unsigned bit = (b >> i) & 1u, g = global_bits[i] - n_local;
            peer = (peer & ~(1 << g)) | ((int)bit << g);
            pat |= (uint64_t)bit << local_bits[i];
        }
        for (uint64_t r = 0; r < nruns; ++r) {
            uint64_t off = pat;
            for (size_t f = 0; f < freebits.size(); ++f) off |= ((r >> f) & 1ull) << freebits[f];
            if (segs && count < maxsegs) segs[count] = rocsvxExchangeSeg{peer, off, off, run};
            ++count;
        }
    }
    return count;
}

rocqStatus_t Dist::init(rocsvInternalHandle*, int, int, const void*) { return ROCQ_STATUS_NOT_IMPLEMENTED; }
rocqStatus_t Dist::allocate(rocsvInternalHandle*, unsigned) { return ROCQ_STATUS_NOT_IMPLEMENTED; }
rocqStatus_t Dist::initialize(rocsvInternalHandle*) { return ROCQ_STATUS_NOT_IMPLEMENTED; }
void Dist::shutdown() {}
rocqStatus_t Dist::localize(rocsvInternalHandle*, HostOp&) { return ROCQ_STATUS_NOT_IMPLEMENTED; }
rocqStatus_t Dist::run_circuit(rocsvInternalHandle*, std::vector<HostOp>&) { return ROCQ_STATUS_NOT_IMPLEMENTED; }
rocqStatus_t Dist::swap_index_bits(rocsvInternalHandle*, unsigned, unsigned) { return ROCQ_STATUS_NOT_IMPLEMENTED; }
rocqStatus_t Dist::canonicalize(rocsvInternalHandle*) { return ROCQ_STATUS_NOT_IMPLEMENTED; }
rocqStatus_t Dist::allreduce_sum(rocsvInternalHandle*, double*, unsigned) { return ROCQ_STATUS_NOT_IMPLEMENTED; }
rocqStatus_t Dist::pauli_expect(rocsvInternalHandle*, uint64_t, uint64_t, unsigned, double*) { return ROCQ_STATUS_NOT_IMPLEMENTED; }
rocqStatus_t Dist::measure(rocsvInternalHandle*, unsigned, int*, double*) { return ROCQ_STATUS_NOT_IMPLEMENTED; }
rocqStatus_t Dist::sample(rocsvInternalHandle*, const unsigned*, unsigned, unsigned, uint64_t*) { return ROCQ_STATUS_NOT_IMPLEMENTED; }

}  // namespace rq

rq::Dist& rq_engine_dist(rocsvInternalHandle* h);

extern "C" {

rocqStatus_t rocsvAllocateDistributedState(rocsvHandle_t h, unsigned totalNumQubits) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    return rq_engine_dist(h).allocate(h, totalNumQubits);
}
rocqStatus_t rocsvInitializeDistributedState(rocsvHandle_t h) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    return rq_engine_dist(h).initialize(h);
}
rocqStatus_t rocsvxDistGetUniqueId(void*) { return ROCQ_STATUS_NOT_IMPLEMENTED; }
rocqStatus_t rocsvxDistInit(rocsvHandle_t h, int rank, int numRanks, const void* id128) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    return rq_engine_dist(h).init(h, rank, numRanks, id128);
}
rocqStatus_t rocsvxDistGetInfo(rocsvHandle_t h, int* rank, int* numRanks, unsigned* numLocalQubits, rocComplex** d_localSlice) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    rq::Dist& d = rq_engine_dist(h);
    if (rank) *rank = d.rank;
    if (numRanks) *numRanks = d.nranks;
    if (numLocalQubits) *numLocalQubits = d.n_local;
    (void)d_localSlice;
    return ROCQ_STATUS_SUCCESS;
}
rocqStatus_t rocsvxDistPlanExchange(unsigned numLocalQubits, int numRanks, int rank, const unsigned* localBits,
                                    const unsigned* globalBits, unsigned numPairs, rocsvxExchangeSeg* segs, size_t maxSegs,
                                    size_t* numSegs) {
    if (!localBits || !globalBits || numPairs == 0 || numRanks < 1 || (numRanks & (numRanks - 1)) || rank < 0 || rank >= numRanks)
        return ROCQ_STATUS_INVALID_VALUE;
    unsigned M = 0;
    while ((1 << M) < numRanks) ++M;
    uint64_t seenL = 0, seenG = 0;
    for (unsigned i = 0; i < numPairs; ++i) {
        if (localBits[i] >= numLocalQubits || globalBits[i] < numLocalQubits || globalBits[i] >= numLocalQubits + M) return ROCQ_STATUS_INVALID_VALUE;
        if (((seenL >> localBits[i]) & 1ull) || ((seenG >> globalBits[i]) & 1ull)) return ROCQ_STATUS_INVALID_VALUE;
        seenL |= 1ull << localBits[i];
        seenG |= 1ull << globalBits[i];
    }
    const size_t c = rq::plan_exchange(numLocalQubits, numRanks, rank, localBits, globalBits, numPairs, segs, maxSegs);
    if (numSegs) *numSegs = c;
    return ROCQ_STATUS_SUCCESS;
}

}  // extern "C"
