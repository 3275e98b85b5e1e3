// rocquantum_b200/csrc/engine.h -- the handle behind rocsvHandle_t and the hooks dist.cu uses.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "../../include/hipStateVec.h"
#include "dist.h"
#include "group.h"
#include "host_ops.h"
#include "sv_internal.h"

#include <nvtx3/nvToolsExt.h>      // header-only (NVTX 3): no link dependency; a no-op unless a profiler injects itself

namespace rq {
// NVTX range around one launch group of the engine: a sweep, an exchange, a reduction ("tracing").
// Visible in Nsight Systems / ncu --nvtx as rocq/<what>; costs two empty calls when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char* what) { nvtxRangePushA(what); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
// Stream-ordered allocation from the pool the handle uses: the device's default pool with its release threshold raised to
// "keep everything" at rocsvCreate.  With the default threshold (0) the pool hands its unused memory back to the OS at every
// synchronisation, so a loop of circuit + read-back calls re-mapped its scratch each time -- measured as stalls of
// 30-100 ms per step on a B200 (profiles/r02_e2e_profile_before.log).
inline cudaError_t pool_alloc(void** p, size_t bytes, cudaMemPool_t pool, cudaStream_t s) {
    return pool ? cudaMallocFromPoolAsync(p, bytes, pool, s) : cudaMallocAsync(p, bytes, s);
}
// stream-ordered scratch that is released on every path out of its scope (early error returns included)
struct StreamBuf {
    void* p = nullptr;
    cudaStream_t s;
    cudaMemPool_t pool;
    StreamBuf(cudaStream_t st, cudaMemPool_t pl) : s(st), pool(pl) {}
    StreamBuf(const StreamBuf&) = delete;
    StreamBuf& operator=(const StreamBuf&) = delete;
    cudaError_t alloc(size_t bytes) { return pool_alloc(&p, bytes, pool, s); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
    ~StreamBuf() { if (p) cudaFreeAsync(p, s); }
};
}  // namespace rq

// One launch of a planned circuit, kept so that resubmitting the identical circuit skips fusion, planning and the host-side
// matrix products (rocsvxApplyCircuit's plan cache): a block sweep (parameters, tensor map, its operand terms in a device
// buffer the entry owns) or a tile sweep (the program that travels in the kernel parameters).
struct rocsvCachedStep {
    bool block = false, large = false;
    rq_block_params bp{};
    alignas(64) unsigned char tmap[128];
    void* d_terms = nullptr;
    std::vector<unsigned char> prog;
    unsigned ops = 0;
};
struct rocsvPlanCache {
    bool valid = false;
    uint64_t key[2] = {0, 0};
    size_t bytes = 0;
    std::vector<rocsvCachedStep> steps;
};

// Reference handle: hipStateVec.cpp:62-68 {stream, batchSize, numQubits, d_state, ownsState}.
struct rocsvInternalHandle {
    cudaStream_t stream = nullptr;
    cudaMemPool_t pool = nullptr;       // the device's default pool, kept from releasing at every sync (rq::pool_alloc)
    size_t batchSize = 1;
    unsigned numQubits = 0;
    rq_cplx* d_state = nullptr;
    bool ownsState = false;
    // deferred gate queue (fusion mode)
    bool fusion = false;
    std::vector<rq::HostOp> queue;
    rq_cplx* queue_state = nullptr;
    unsigned queue_n = 0;
    // RNG
    uint64_t seed = 0, draws = 0;       // seed: std::random_device at rocsvCreate unless ROCQ_SEED / rocsvxSetSeed
    bool seedExplicit = false;
    // scratch
    double* d_partials = nullptr;       // RBLOCKS doubles + 8 results
    uint64_t* d_upartials = nullptr;    // 4*RBLOCKS + 4
    void* h_scratch = nullptr;          // pinned, 4 KB
    void* pinned = nullptr;             // user-visible pinned buffer (rocsvEnsurePinnedBuffer)
    size_t pinnedSize = 0;
    void* stage[2] = {nullptr, nullptr};   // page-locked staging of state import / export (two halves in flight)
    cudaEvent_t stageEv[2] = {nullptr, nullptr};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, tm0 = nullptr, tm1 = nullptr;
    // tuning
    int tcBlocks = -1;                  // 6-qubit tensor-core blocks in rocsvxApplyCircuit / fused flushes: 0 off, 1 on, -1 auto (on from RQ_BLOCK_AUTO_QUBITS qubits)
    double blockMinCost = 54.0;         // fold >= this much HostOp::cost() (3 dense 2q gates) or stay on the CUDA cores
    bool mergeDiagonals = true;         // runs of controlled phases on one hub qubit become one RQ_OP_DIAGP (ROCQ_MERGE_DIAG=0: off)
    unsigned tileBits = RQ_MAX_TILE_BITS;
    double budget = 1e30;
    rocsvxStats stats{};
    // plan cache of rocsvxApplyCircuit (ROCQ_PLAN_CACHE=0: off)
    bool planCache = true;
    rocsvPlanCache cache;
    std::vector<rocsvCachedStep>* recording = nullptr;   // non-null while a circuit is being planned for the cache
    bool recordingValid = true;
    rq::Dist dist;
    // ROCQ_TRACE_LAUNCHES=1: a CUDA event pair around every sweep / exchange launch, dumped (label, ms) to stderr by
    // rocsvxGetStats -- a launch list that also works where ncu cannot follow (one rank of a multi-process run)
    bool traceLaunches = false;
    struct Traced { const char* what; unsigned ops; cudaEvent_t e0, e1; };
    std::vector<Traced> traced;
    // single-process multi-GPU (group.h): this handle is the front of `group`; its own stream / state stay unused
    rocsvGroup* group = nullptr;
    int wantRanks = 0;                  // rocsvxDistSetRanks: slices of the next rocsvAllocateDistributedState (0: one per visible device)
};
void rq_group_destroy(rocsvInternalHandle* h);
rocqStatus_t rq_group_create(rocsvInternalHandle* h, int ranks);

// trace scope: records the event pair around the launches issued while it lives (no-op unless h->traceLaunches)
struct rq_trace_scope {
    rocsvInternalHandle* h;
    size_t idx = (size_t)-1;
    rq_trace_scope(rocsvInternalHandle* h_, const char* what, unsigned ops = 0) : h(h_) {
        if (!h->traceLaunches || h->traced.size() >= 4096) return;
        rocsvInternalHandle::Traced t{what, ops, nullptr, nullptr};
        if (cudaEventCreate(&t.e0) != cudaSuccess || cudaEventCreate(&t.e1) != cudaSuccess) return;
        cudaEventRecord(t.e0, h->stream);
        idx = h->traced.size();
        h->traced.push_back(t);
    }
    ~rq_trace_scope() { if (idx != (size_t)-1) cudaEventRecord(h->traced[idx].e1, h->stream); }
};
rocqStatus_t rq_engine_flush(rocsvInternalHandle* h);
rocqStatus_t rq_engine_run(rocsvInternalHandle* h, rq_cplx* state, unsigned n, const std::vector<rq::HostOp>& ops, bool fused);
rocqStatus_t rq_engine_fetch(rocsvInternalHandle* h, const void* dsrc, void* hdst, size_t bytes);
uint64_t rq_uniform53(uint64_t seed, uint64_t call, uint64_t shot);
