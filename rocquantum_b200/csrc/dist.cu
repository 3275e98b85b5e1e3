// rocquantum_b200/csrc/dist.cu -- distributed state over NCCL, one process per GPU (see dist.h).
//
// The reference documents this path (MULTI_GPU_GUIDE.md, hipStateVec.h:84-137) but defines none of it; its
// packing kernels (swap_kernels.hip:46-89) use atomic cursors whose arrival order is nondeterministic, so a
// receiver could not place the data.  Here a k-bit global<->local exchange moves, per peer, block-contiguous
// runs that land directly in their final position (no counts, no packing pass):
//   * the engine keeps a logical->physical qubit map; physical positions >= n_local are the rank bits;
//   * a gate with a non-diagonal target on a rank bit triggers an exchange that swaps k rank bits with the
//     top k local bits (the evicted logical qubits are first moved to those top slots by PERM ops that ride in
//     the preceding fused sweep), chosen by farthest-next-use when a whole circuit is known;
//   * controls and diagonal gates on rank bits need no communication (the tile base carries the rank);
//   * scalars (probability masses, expectation values) are exact all-gathers / fp64 all-reduces.
// Two data movers for the exchange: NCCL send/recv into a staging buffer + copy (default), and -- behind
// ROCQ_EXCHANGE=p2p -- an in-place swap kernel over IPC-mapped peer slices (peer_swap_kernel below).
#include <cuda_runtime.h>

#include <cstdlib>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "dist_plan.h"
#include "engine.h"
#include "group.h"

namespace rq {

// ---- NCCL through dlopen: the library has no link-time dependency on it --------------------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { fprintf(stderr, "hipStateVec(B200): cannot load libnccl: %s\n", dlerror()); return false; }
#define RQ_SYM(f) f = reinterpret_cast<decltype(f)>(dlsym(lib, "nccl" #f)); if (!f) { fprintf(stderr, "hipStateVec(B200): nccl" #f " missing\n"); return false; }
        RQ_SYM(GetUniqueId) RQ_SYM(CommInitRank) RQ_SYM(CommDestroy) RQ_SYM(Send) RQ_SYM(Recv) RQ_SYM(GroupStart) RQ_SYM(GroupEnd)
        RQ_SYM(AllReduce) RQ_SYM(AllGather) RQ_SYM(GetErrorString)
#undef RQ_SYM
        return true;
    }
};
static NcclApi g_nccl;

#define RQ_NCCL(call)                                                                             \
    do {                                                                                          \
        const ncclResult_t _r = (call);                                                           \
        if (_r != ncclSuccess) {                                                                  \
            fprintf(stderr, "hipStateVec(B200): %s failed: %s\n", #call, g_nccl.GetErrorString(_r)); \
            return ROCQ_STATUS_RCCL_ERROR;                                                        \
        }                                                                                         \
    } while (0)
#define RQ_CU(call)                                                                               \
    do {                                                                                          \
        const cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess) {                                                                  \
            fprintf(stderr, "hipStateVec(B200): %s failed: %s\n", #call, cudaGetErrorString(_e)); \
            return ROCQ_STATUS_HIP_ERROR;                                                         \
        }                                                                                         \
    } while (0)
#define RQ_OK(call) do { const rocqStatus_t _s = (call); if (_s != ROCQ_STATUS_SUCCESS) return _s; } while (0)

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
    std::vector<unsigned> freebits;                     // non-exchanged local bits above minl enumerate the runs
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

// Peer-memory form of the same exchange.  Every segment of plan_exchange is one half of an in-place swap between this
// rank's run and the peer's matching run.  The two ranks split each swap -- the lower rank moves the first half of the run,
// the higher rank the rest -- so every element is read and written by exactly one of them and nothing is staged.  Output
// segment: {peer, sendOffset = offset in OWN slice, recvOffset = offset in the PEER's slice, count}; empty halves are
// dropped.  Pure host code (tests/test_dist_cpu.py re-plays it on numpy shards).
size_t plan_peer_swap(unsigned n_local, int nranks, int rank, const unsigned* local_bits, const unsigned* global_bits,
                      unsigned npairs, rocsvxExchangeSeg* out, size_t maxsegs) {
    const size_t ns = plan_exchange(n_local, nranks, rank, local_bits, global_bits, npairs, nullptr, 0);
    std::vector<rocsvxExchangeSeg> mine(ns), theirs(ns);
    plan_exchange(n_local, nranks, rank, local_bits, global_bits, npairs, mine.data(), ns);
    std::vector<char> done(ns, 0);
    size_t count = 0;
    for (size_t i = 0; i < ns; ++i) {
        if (done[i]) continue;
        const int peer = mine[i].peer;
        plan_exchange(n_local, nranks, peer, local_bits, global_bits, npairs, theirs.data(), ns);   // same segment count on every rank
        size_t j = 0;
        for (size_t m = i; m < ns; ++m) {                      // the m-th run towards `peer` pairs with the peer's m-th run towards us
            if (mine[m].peer != peer) continue;
            while (j < ns && theirs[j].peer != rank) ++j;
            if (j == ns) return 0;                              // plans disagree: cannot happen for valid bit lists
            done[m] = 1;
            const uint64_t half = mine[m].count / 2;
            const bool low = rank < peer;
            const uint64_t skip = low ? 0 : half, len = low ? half : mine[m].count - half;
            if (len) {
                if (out && count < maxsegs) out[count] = rocsvxExchangeSeg{peer, mine[m].sendOffset + skip, theirs[j].sendOffset + skip, len};
                ++count;
            }
            ++j;
        }
    }
    return count;
}

// swap local[i] <-> peer[i]: 16-byte (or one-amplitude) elements, U independent loads of each side in flight per thread
#define RQ_PEER_MAXSEG 15
struct PeerSwapArgs {
    void* local[RQ_PEER_MAXSEG];
    void* peer[RQ_PEER_MAXSEG];
    uint64_t count[RQ_PEER_MAXSEG];          // elements of type V
};
template <typename V>
__global__ void __launch_bounds__(256) peer_swap_kernel(const PeerSwapArgs a) {
    constexpr int U = 4;
    V* __restrict__ L = static_cast<V*>(a.local[blockIdx.y]);
    V* __restrict__ P = static_cast<V*>(a.peer[blockIdx.y]);
    const uint64_t n = a.count[blockIdx.y];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        V p[U], l[U];
#pragma unroll
        for (int u = 0; u < U; ++u) p[u] = P[i + u * stride];          // remote reads first: the long-latency side
#pragma unroll
        for (int u = 0; u < U; ++u) l[u] = L[i + u * stride];
#pragma unroll
        for (int u = 0; u < U; ++u) { P[i + u * stride] = l[u]; L[i + u * stride] = p[u]; }
    }
    for (; i < n; i += stride) {
        const V p = P[i], l = L[i];
        P[i] = l;
        L[i] = p;
    }
}

// ---- lifecycle -------------------------------------------------------------------------------------------
rocqStatus_t Dist::init(rocsvInternalHandle* h, int rank_, int nranks_, const void* id128) {
    (void)h;
    if (nranks_ < 1 || (nranks_ & (nranks_ - 1)) || rank_ < 0 || rank_ >= nranks_) return ROCQ_STATUS_INVALID_VALUE;   // hipStateVec.h:86
    if (inited) return ROCQ_STATUS_INVALID_VALUE;
    rank = rank_;
    nranks = nranks_;
    if (nranks > 1) {
        if (!id128) return ROCQ_STATUS_INVALID_VALUE;
        if (!g_nccl.load()) return ROCQ_STATUS_RCCL_ERROR;
        ncclUniqueId id;
        static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id size");
        memcpy(&id, id128, 128);
        ncclComm_t c = nullptr;
        RQ_NCCL(g_nccl.CommInitRank(&c, nranks, id, rank));
        comm = c;
        RQ_CU(cudaMalloc(&d_gather, (size_t)(64 + 8 * nranks) * sizeof(uint64_t)));
    }
    inited = true;
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t Dist::init_in_group(rocsvInternalHandle* h, int rank_, rocsvGroup* g) {
    (void)h;
    if (inited || !g || rank_ < 0 || rank_ >= g->P) return ROCQ_STATUS_INVALID_VALUE;
    rank = rank_;
    nranks = g->P;
    group = g;
    RQ_CU(cudaMalloc(&d_gather, (size_t)(64 + 8 * nranks) * sizeof(uint64_t)));
    RQ_CU(cudaEventCreateWithFlags(&g->ev[rank], cudaEventDisableTiming));
    inited = true;
    return ROCQ_STATUS_SUCCESS;
}

// ---- peer slices through CUDA IPC (ROCQ_EXCHANGE=p2p) ------------------------------------------------------------------
// default mover: the peer-memory swap kernel (measured on 2 B200s: 688 GB/s per direction against 303 GB/s for
// ncclSend/ncclRecv + staging copy, profiles/r02_bench_n2_*.log); ROCQ_EXCHANGE=nccl keeps the NCCL mover
static bool want_p2p() {
    const char* e = getenv("ROCQ_EXCHANGE");
    return !(e && (e[0] == 'n' || e[0] == 'N'));
}
// stream-ordered barrier over all ranks: nobody's stream passes it before everybody's stream has reached it
static rocqStatus_t stream_barrier(rocsvInternalHandle* h, Dist& d) {
    if (d.group) {
        // one process: every rank records an event, and every stream waits for all the others' events.  The host barriers
        // only order the host calls (record before wait, wait before the next record); the streams never block the host.
        rocsvGroup* g = d.group;
        RQ_CU(cudaEventRecord(g->ev[d.rank], h->stream));
        g->barrier();
        for (int r = 0; r < d.nranks; ++r) if (r != d.rank) RQ_CU(cudaStreamWaitEvent(h->stream, g->ev[r], 0));
        g->barrier();
        return ROCQ_STATUS_SUCCESS;
    }
    int* flag = reinterpret_cast<int*>(d.d_gather + 40);               // a word neither the all-gathers nor allreduce_sum use
    RQ_NCCL(g_nccl.AllReduce(flag, flag, 1, ncclInt, ncclMax, (ncclComm_t)d.comm, h->stream));
    return ROCQ_STATUS_SUCCESS;
}
static void close_peers(rocsvInternalHandle* h, Dist& d) {
    if (d.group) {                                                     // plain pointers of one address space: nothing to unmap,
        if (d.peer_state.empty()) return;                              // but nobody may free its slice while a peer's kernel uses it
        cudaStreamSynchronize(h->stream);
        d.peer_state.clear();
        d.group->barrier();
        return;
    }
    if (d.peer_state.empty()) return;
    cudaStreamSynchronize(h->stream);
    for (int r = 0; r < d.nranks; ++r)
        if (r != d.rank && d.peer_state[r]) cudaIpcCloseMemHandle(d.peer_state[r]);
    d.peer_state.clear();
    if (d.comm && stream_barrier(h, d) == ROCQ_STATUS_SUCCESS) cudaStreamSynchronize(h->stream);   // owners free only after every importer closed
}
// map every rank's slice; all ranks agree on the outcome (one rank without peer access sends everybody back to NCCL)
static rocqStatus_t open_peers(rocsvInternalHandle* h, Dist& d) {
    if (d.group) {
        rocsvGroup* g = d.group;
        g->slice[d.rank] = h->d_state;
        g->barrier();
        int ok = 1;
        for (int r = 0; r < d.nranks && ok; ++r) {
            if (g->device[r] == g->device[d.rank]) continue;           // same device: directly addressable
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, g->device[d.rank], g->device[r]) != cudaSuccess || !can) { cudaGetLastError(); ok = 0; break; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(g->device[r], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
            cudaGetLastError();
        }
        std::vector<void*> ptrs(g->slice);
        double fail = ok ? 0.0 : 1.0;
        RQ_OK(d.allreduce_sum(h, &fail, 1));                           // (ends with a barrier: g->slice may be rewritten afterwards)
        if (fail != 0.0) {
            if (d.rank == 0) fprintf(stderr, "hipStateVec(B200): single-process multi-GPU needs peer access between the devices (%d rank(s) without)\n", (int)fail);
            return ROCQ_STATUS_NOT_IMPLEMENTED;
        }
        d.peer_state.swap(ptrs);
        return ROCQ_STATUS_SUCCESS;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    if (d.nranks > 64) return ROCQ_STATUS_SUCCESS;                     // the gathered handles travel through the 4 KB pinned scratch
    cudaIpcMemHandle_t mine;
    int ok = cudaIpcGetMemHandle(&mine, h->d_state) == cudaSuccess;
    if (!ok) { cudaGetLastError(); memset(&mine, 0, sizeof mine); }
    uint64_t* send = d.d_gather;
    uint64_t* recv = d.d_gather + 64;                                  // 8 * nranks words = 64 bytes per rank
    RQ_CU(cudaMemcpyAsync(send, &mine, sizeof mine, cudaMemcpyHostToDevice, h->stream));
    RQ_NCCL(g_nccl.AllGather(send, recv, sizeof mine, ncclChar, (ncclComm_t)d.comm, h->stream));
    std::vector<cudaIpcMemHandle_t> all((size_t)d.nranks);
    RQ_OK(rq_engine_fetch(h, recv, all.data(), all.size() * sizeof mine));
    std::vector<void*> ptrs((size_t)d.nranks, nullptr);
    ptrs[d.rank] = h->d_state;
    for (int r = 0; r < d.nranks && ok; ++r) {
        if (r == d.rank) continue;
        if (cudaIpcOpenMemHandle(&ptrs[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ptrs[r] = nullptr; ok = 0; }
    }
    double agree = ok ? 0.0 : 1.0;                                      // number of ranks that failed
    RQ_OK(d.allreduce_sum(h, &agree, 1));
    if (agree != 0.0) {
        for (int r = 0; r < d.nranks; ++r) if (r != d.rank && ptrs[r]) cudaIpcCloseMemHandle(ptrs[r]);
        if (d.rank == 0) fprintf(stderr, "hipStateVec(B200): ROCQ_EXCHANGE=p2p: peer slices cannot be mapped on %d rank(s); using NCCL send/recv\n", (int)agree);
        return ROCQ_STATUS_SUCCESS;
    }
    d.peer_state.swap(ptrs);
    return ROCQ_STATUS_SUCCESS;
}

static rocqStatus_t allgather_u64(rocsvInternalHandle* h, Dist& d, const uint64_t* mine, unsigned count, std::vector<uint64_t>& out);

void Dist::shutdown(rocsvInternalHandle* h) {
    close_peers(h, *this);
    for (auto& ev : timed) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    timed.clear();
    if (staging) { cudaFree(staging); staging = nullptr; }
    if (d_gather) { cudaFree(d_gather); d_gather = nullptr; }
    if (comm) { g_nccl.CommDestroy((ncclComm_t)comm); comm = nullptr; }
    if (group && group->ev[rank]) { cudaEventDestroy(group->ev[rank]); group->ev[rank] = nullptr; }
    group = nullptr;
    inited = false;
    n_total = 0;
}

rocqStatus_t Dist::allocate(rocsvInternalHandle* h, unsigned total_qubits) {
    if (!inited) { rank = 0; nranks = 1; inited = true; }            // single process, single GPU: a plain state
    unsigned M = 0;
    while ((1 << M) < nranks) ++M;
    if (total_qubits < M || total_qubits > 60) return ROCQ_STATUS_INVALID_VALUE;
    n_total = 0;                                                      // inactive while (re)allocating
    close_peers(h, *this);                                            // importers unmap before any owner frees its slice
    const rocqStatus_t s = rocsvAllocateState(h, total_qubits - M, nullptr, 1);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    n_global = M;
    n_local = total_qubits - M;
    n_total = total_qubits;
    plan.reset(n_total, n_local);
    if (staging) { cudaFree(staging); staging = nullptr; }
    if (nranks > 1) {
        if (group) {
            RQ_OK(open_peers(h, *this));                  // one process: the exchange always goes through peer memory
        } else {
            const uint64_t slice = 1ull << n_local;
            staging_amps = std::min<uint64_t>(slice, 1ull << 25) * (uint64_t)(nranks - 1);     // <= 256 MiB (c64) per peer
            if (cudaMalloc(&staging, staging_amps * sizeof(rq_cplx)) != cudaSuccess) { cudaGetLastError(); return ROCQ_STATUS_ALLOCATION_FAILED; }
            if (want_p2p()) RQ_OK(open_peers(h, *this));
        }
        // every rank evaluates the same Philox stream (measure / sample): all ranks continue rank 0's
        std::vector<uint64_t> all;
        const uint64_t mine[2] = {h->seed, h->draws};
        RQ_OK(allgather_u64(h, *this, mine, 2, all));
        h->seed = all[0];
        h->draws = all[1];
    }
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t Dist::initialize(rocsvInternalHandle* h) {
    if (!active() || !h->d_state) return ROCQ_STATUS_INVALID_VALUE;
    h->queue.clear();
    plan.reset(n_total, n_local);
    const int e = rq_launch_init_state(h->d_state, (size_t)1 << n_local, rank == 0, h->stream);     // hipStateVec.h:97-99
    h->stats.kernelLaunches++;
    h->numQubits = n_local;
    return e == 0 ? ROCQ_STATUS_SUCCESS : ROCQ_STATUS_HIP_ERROR;
}

// ---- the exchange -------------------------------------------------------------------------------------------
// Event pairs wait for rocsvxGetStats; a caller that never asks for stats must not accumulate them without bound.
static void note_exchange(rocsvInternalHandle* h, Dist& d, cudaEvent_t e0, cudaEvent_t e1) {
    if (d.timed.size() >= 256) {                       // the oldest pair finished long ago: fold it in now
        const auto ev = d.timed.front();
        float ms = 0.f;
        cudaEventSynchronize(ev.second);
        if (cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) h->stats.exchangeMs += (double)ms;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
        d.timed.erase(d.timed.begin());
    }
    d.timed.push_back({e0, e1});                       // resolved (and destroyed) by rocsvxGetStats
}
// move the data of one EXCHANGE step: the k rank bits gpos[] trade places with the top-k local bits
static rocqStatus_t exchange_data(rocsvInternalHandle* h, Dist& d, const std::vector<unsigned>& gpos) {
    const NvtxRange nvtx("rocq/exchange_nccl");
    const rq_trace_scope trace(h, "exchange_nccl");
    const unsigned k = (unsigned)gpos.size();
    if (k == 0) return ROCQ_STATUS_SUCCESS;
    if (!d.comm) return ROCQ_STATUS_NOT_IMPLEMENTED;                   // (single-process groups always have peer slices)
    std::vector<unsigned> lpos(k);
    for (unsigned i = 0; i < k; ++i) lpos[i] = d.n_local - k + i;
    std::vector<rocsvxExchangeSeg> segs((size_t)(1u << k));
    const size_t ns = plan_exchange(d.n_local, d.nranks, d.rank, lpos.data(), gpos.data(), k, segs.data(), segs.size());
    const uint64_t run = 1ull << (d.n_local - k);
    const uint64_t chunk = std::min<uint64_t>(run, d.staging_amps / (uint64_t)std::max<size_t>(ns, 1));
    ncclComm_t comm = (ncclComm_t)d.comm;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    RQ_CU(cudaEventCreate(&e0));
    RQ_CU(cudaEventCreate(&e1));
    RQ_CU(cudaEventRecord(e0, h->stream));
    for (uint64_t c = 0; c < run; c += chunk) {
        const uint64_t len = std::min<uint64_t>(chunk, run - c);
        RQ_NCCL(g_nccl.GroupStart());
        for (size_t i = 0; i < ns; ++i) {
            RQ_NCCL(g_nccl.Send(h->d_state + segs[i].sendOffset + c, len * sizeof(rq_cplx), ncclChar, segs[i].peer, comm, h->stream));
            RQ_NCCL(g_nccl.Recv(d.staging + i * chunk, len * sizeof(rq_cplx), ncclChar, segs[i].peer, comm, h->stream));
        }
        RQ_NCCL(g_nccl.GroupEnd());
        for (size_t i = 0; i < ns; ++i)
            RQ_CU(cudaMemcpyAsync(h->d_state + segs[i].recvOffset + c, d.staging + i * chunk, len * sizeof(rq_cplx),
                                  cudaMemcpyDeviceToDevice, h->stream));
    }
    RQ_CU(cudaEventRecord(e1, h->stream));
    note_exchange(h, d, e0, e1);
    d.exchanges++;
    d.exchanged_amps += run * ns;
    h->stats.exchanges++;
    h->stats.exchangeBytes += run * ns * sizeof(rq_cplx);
    return ROCQ_STATUS_SUCCESS;
}

// The same EXCHANGE step over peer memory: after a stream-ordered barrier (every rank has finished the sweeps before the
// exchange) each rank swaps its halves of the runs in place against the peers' slices -- reads and writes cross NVLink in
// both directions at once, no staging pass -- and a second barrier keeps the next sweep behind the peers' writes.
static rocqStatus_t exchange_data_p2p(rocsvInternalHandle* h, Dist& d, const std::vector<unsigned>& gpos) {
    const NvtxRange nvtx("rocq/exchange_peer");
    const rq_trace_scope trace(h, "exchange_peer");
    const unsigned k = (unsigned)gpos.size();
    std::vector<unsigned> lpos(k);
    for (unsigned i = 0; i < k; ++i) lpos[i] = d.n_local - k + i;
    std::vector<rocsvxExchangeSeg> segs((size_t)(1u << k));
    const size_t ns = plan_peer_swap(d.n_local, d.nranks, d.rank, lpos.data(), gpos.data(), k, segs.data(), segs.size());
    const uint64_t run = 1ull << (d.n_local - k);
    PeerSwapArgs a{};
    const uint64_t per16 = 16 / sizeof(rq_cplx);                       // amplitudes per 16-byte element
    const bool vec = (run / 2) % per16 == 0 && (run - run / 2) % per16 == 0;   // both halves start and end on 16-byte boundaries
    const uint64_t per = vec ? per16 : 1;
    uint64_t longest = 0;
    for (size_t i = 0; i < ns; ++i) {
        a.local[i] = h->d_state + segs[i].sendOffset;
        a.peer[i] = static_cast<rq_cplx*>(d.peer_state[segs[i].peer]) + segs[i].recvOffset;
        a.count[i] = segs[i].count / per;
        longest = std::max(longest, a.count[i]);
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    RQ_CU(cudaEventCreate(&e0));
    RQ_CU(cudaEventCreate(&e1));
    RQ_CU(cudaEventRecord(e0, h->stream));
    RQ_OK(stream_barrier(h, d));
    if (ns) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const uint64_t want = (longest + 256 * 4 - 1) / (256 * 4);      // one pass of the unrolled loop per thread at least
        const unsigned gx = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)(4 * sms + ns - 1) / ns));
        const dim3 grid(gx, (unsigned)ns);
        if (vec) peer_swap_kernel<uint4><<<grid, 256, 0, h->stream>>>(a);
        else peer_swap_kernel<rq_cplx><<<grid, 256, 0, h->stream>>>(a);
        RQ_CU(cudaGetLastError());
        h->stats.kernelLaunches++;
    }
    RQ_OK(stream_barrier(h, d));
    RQ_CU(cudaEventRecord(e1, h->stream));
    note_exchange(h, d, e0, e1);
    d.exchanges++;
    d.exchanged_amps += run * ((1ull << k) - 1);
    h->stats.exchanges++;
    h->stats.exchangeBytes += run * ((1ull << k) - 1) * sizeof(rq_cplx);     // amplitudes that leave this slice, as in the NCCL path
    return ROCQ_STATUS_SUCCESS;
}

// execute and clear the planner's steps
static rocqStatus_t execute_steps(rocsvInternalHandle* h, Dist& d) {
    for (DistStep& st : d.plan.steps) {
        if (st.kind == DistStep::RUN) RQ_OK(rq_engine_run(h, h->d_state, d.n_local, st.ops, true));
        else if (!d.peer_state.empty() && st.gpos.size() <= 4 && !st.gpos.empty()) RQ_OK(exchange_data_p2p(h, d, st.gpos));
        else RQ_OK(exchange_data(h, d, st.gpos));
    }
    d.plan.steps.clear();
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t Dist::localize(rocsvInternalHandle* h, HostOp& op) {
    plan.steps.clear();
    plan.pending.clear();
    if (!plan.add_op(op)) return ROCQ_STATUS_NOT_IMPLEMENTED;          // more global targets than free local qubits
    HostOp phys = std::move(plan.pending.back());
    plan.pending.clear();
    if (!plan.steps.empty()) {
        if (nranks == 1) return ROCQ_STATUS_FAILURE;
        RQ_OK(rq_engine_flush(h));                                        // queued ops use the previous layout
        RQ_OK(execute_steps(h, *this));
    }
    op = std::move(phys);
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t Dist::run_circuit(rocsvInternalHandle* h, std::vector<HostOp>& ops) {
    plan.steps.clear();
    plan.pending.clear();
    static const bool inorder = getenv("ROCQ_DIST_INORDER") != nullptr && atoi(getenv("ROCQ_DIST_INORDER")) != 0;
    const std::vector<unsigned> saved = plan.map;      // nothing has been executed yet: a failed plan must not move the map
    bool planned = !inorder && plan.add_circuit(ops);
    if (!planned) {                                   // program order (also the fallback if the deferring planner gives up)
        plan.map = saved;
        plan.steps.clear();
        plan.pending.clear();
        planned = plan.add_circuit_inorder(ops);
        if (!planned) { plan.map = saved; plan.steps.clear(); plan.pending.clear(); return ROCQ_STATUS_FAILURE; }
    }
    plan.flush_pending();
    cudaEventRecord(h->ev0, h->stream);
    RQ_OK(execute_steps(h, *this));
    cudaEventRecord(h->ev1, h->stream);
    h->stats.lastSweepMs = -1.0;
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t Dist::swap_index_bits(rocsvInternalHandle* h, unsigned q1, unsigned q2) {
    // Physically relabelling two index bits equals a SWAP gate on the two logical qubits (hipStateVec.h:123-137).
    // local<->local: a PERM sweep; local<->global: one exchange; global<->global: through local slots
    // (the reference leaves this case NOT_IMPLEMENTED, MULTI_GPU_GUIDE.md:50).
    RQ_OK(rq_engine_flush(h));                        // gates still queued (fusion mode) come before the relabelling
    HostOp op = make_swap(q1, q2);
    RQ_OK(localize(h, op));
    std::vector<HostOp> one{op};
    return rq_engine_run(h, h->d_state, n_local, one, false);
}

// restore the identity layout (needed before the slice is handed to the caller)
rocqStatus_t Dist::canonicalize(rocsvInternalHandle* h) {
    RQ_OK(rq_engine_flush(h));
    plan.steps.clear();
    plan.pending.clear();
    plan.canonicalize();
    return execute_steps(h, *this);
}

// ---- scalars ---------------------------------------------------------------------------------------------------
rocqStatus_t Dist::allreduce_sum(rocsvInternalHandle* h, double* v, unsigned count) {
    if (nranks == 1) return ROCQ_STATUS_SUCCESS;
    if (count > 32) return ROCQ_STATUS_INVALID_VALUE;
    if (group) {                                                       // threads of one process: through host memory, rank order
        for (unsigned i = 0; i < count; ++i) group->dvals[(size_t)rank * 32 + i] = v[i];
        group->barrier();
        for (unsigned i = 0; i < count; ++i) {
            double s = 0.0;
            for (int r = 0; r < nranks; ++r) s += group->dvals[(size_t)r * 32 + i];
            v[i] = s;
        }
        group->barrier();
        return ROCQ_STATUS_SUCCESS;
    }
    double* dbuf = reinterpret_cast<double*>(d_gather);
    RQ_CU(cudaMemcpyAsync(dbuf, v, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    RQ_NCCL(g_nccl.AllReduce(dbuf, dbuf, count, ncclDouble, ncclSum, (ncclComm_t)comm, h->stream));
    return rq_engine_fetch(h, dbuf, v, count * sizeof(double));
}

// gather `count` u64 from every rank: out[r*count + i]
static rocqStatus_t allgather_u64(rocsvInternalHandle* h, Dist& d, const uint64_t* mine, unsigned count, std::vector<uint64_t>& out) {
    out.assign((size_t)count * d.nranks, 0);
    if (d.nranks == 1) { for (unsigned i = 0; i < count; ++i) out[i] = mine[i]; return ROCQ_STATUS_SUCCESS; }
    if (count > 8) return ROCQ_STATUS_INVALID_VALUE;
    if (d.group) {
        for (unsigned i = 0; i < count; ++i) d.group->uvals[(size_t)d.rank * 8 + i] = mine[i];
        d.group->barrier();
        for (int r = 0; r < d.nranks; ++r)
            for (unsigned i = 0; i < count; ++i) out[(size_t)r * count + i] = d.group->uvals[(size_t)r * 8 + i];
        d.group->barrier();
        return ROCQ_STATUS_SUCCESS;
    }
    uint64_t* send = d.d_gather;
    uint64_t* recv = d.d_gather + 64;
    RQ_CU(cudaMemcpyAsync(send, mine, count * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
    RQ_NCCL(g_nccl.AllGather(send, recv, count, ncclUint64, (ncclComm_t)d.comm, h->stream));
    return rq_engine_fetch(h, recv, out.data(), out.size() * sizeof(uint64_t));
}

rocqStatus_t Dist::pauli_expect(rocsvInternalHandle* h, uint64_t xm, uint64_t zm, unsigned ny, double* result) {
    // X/Y factors pair amplitudes across the flipped bits: those qubits must be local
    plan.steps.clear();
    plan.pending.clear();
    if (!plan.make_local(xm)) return ROCQ_STATUS_NOT_IMPLEMENTED;
    RQ_OK(execute_steps(h, *this));
    const std::vector<unsigned>& map = plan.map;
    uint64_t pxm = 0, pzm = 0;
    for (unsigned q = 0; q < n_total; ++q) {
        if ((xm >> q) & 1ull) pxm |= 1ull << map[q];
        if ((zm >> q) & 1ull) pzm |= 1ull << map[q];
    }
    const uint64_t lmask = (1ull << n_local) - 1ull;
    const unsigned nb = rq_reduce_blocks();
    if (rq_launch_pauli_expect(h->d_state, n_local, pxm & lmask, pzm & lmask, ny, h->d_partials, nb, h->d_partials + nb, h->stream) != 0)
        return ROCQ_STATUS_HIP_ERROR;
    h->stats.kernelLaunches += 2;
    double v = 0.0;
    RQ_OK(rq_engine_fetch(h, h->d_partials + nb, &v, sizeof v));
    if (__builtin_popcountll(((uint64_t)rank << n_local) & pzm) & 1) v = -v;       // Z factors on rank bits: a sign per rank
    RQ_OK(allreduce_sum(h, &v, 1));
    *result = v;
    return ROCQ_STATUS_SUCCESS;
}

typedef unsigned __int128 u128;
static inline u128 mul_u53(u128 S, uint64_t U) {
    const u128 A = (u128)(uint64_t)S * U, B = (u128)(uint64_t)(S >> 64) * U;
    return (B << 11) + (A >> 53);
}
static inline double u128_to_double(u128 v) { return (double)(uint64_t)(v >> 64) * 0x1p64 + (double)(uint64_t)v; }

rocqStatus_t Dist::measure(rocsvInternalHandle* h, unsigned q, int* outcome, double* probability) {
    const unsigned p = plan.map[q];
    const bool local = p < n_local;
    const unsigned nb = rq_reduce_blocks();
    if (rq_launch_fixed_masses(h->d_state, n_local, local ? p : 63u, h->d_upartials, nb, h->d_upartials + 4 * nb, h->stream) != 0)
        return ROCQ_STATUS_HIP_ERROR;
    h->stats.kernelLaunches += 2;
    uint64_t m[4];
    RQ_OK(rq_engine_fetch(h, h->d_upartials + 4 * nb, m, sizeof m));
    std::vector<uint64_t> all;
    RQ_OK(allgather_u64(h, *this, m, 4, all));
    u128 S0 = 0, S1 = 0;
    for (int r = 0; r < nranks; ++r) {
        const u128 a0 = ((u128)all[4 * r] << 64) | all[4 * r + 1], a1 = ((u128)all[4 * r + 2] << 64) | all[4 * r + 3];
        if (local) { S0 += a0; S1 += a1; }
        else if ((r >> (p - n_local)) & 1) S1 += a0 + a1;
        else S0 += a0 + a1;
    }
    const uint64_t U = rq_uniform53(h->seed, h->draws++, 0);                  // same stream on every rank
    const int out = mul_u53(S0 + S1, U) < S0 ? 0 : 1;
    const double tot = u128_to_double(S0 + S1), mass = u128_to_double(out ? S1 : S0);
    if (!(mass > 0.0)) return ROCQ_STATUS_FAILURE;
    *outcome = out;
    if (probability) *probability = mass / tot;
    const double scale = 1.0 / std::sqrt(mass * 0x1p-88);
    int e;
    if (local) e = rq_launch_collapse(h->d_state, n_local, p, out, scale, h->stream);
    else if ((int)((rank >> (p - n_local)) & 1) == out) e = rq_launch_collapse(h->d_state, n_local, 63u, 0, scale, h->stream);   // keep: scale all
    else e = rq_launch_init_state(h->d_state, (size_t)1 << n_local, 0, h->stream);                                              // discard: zero all
    h->stats.kernelLaunches++;
    return e == 0 ? ROCQ_STATUS_SUCCESS : ROCQ_STATUS_HIP_ERROR;
}

rocqStatus_t Dist::sample(rocsvInternalHandle* h, const unsigned* measured, unsigned nm, unsigned shots, uint64_t* out) {
    // Everything stays on the devices: chunk masses -> exact scan -> (all-gather of the slice totals) -> one warp per shot.
    // Exactly one rank owns each shot and writes its result word; the others leave 0, so a sum over ranks gathers.
    const unsigned n = n_local;
    const unsigned cb = n < 10 ? n : (n > 30 ? n - 20 : 10);
    const uint64_t nchunks = 1ull << (n - cb);
    StreamBuf scratch(h->stream, h->pool);
    RQ_CU(scratch.alloc((2 * nchunks + 2 * RQ_SCAN_MAXSEG + 4 + (size_t)shots) * sizeof(uint64_t)));
    uint64_t* d_hi = scratch.as<uint64_t>();
    uint64_t *d_lo = d_hi + nchunks, *d_btot = d_lo + nchunks, *d_tot = d_btot + 2 * RQ_SCAN_MAXSEG, *d_idx = d_tot + 4;
    if (rq_launch_chunk_masses(h->d_state, n, cb, d_hi, d_lo, h->stream) != 0) return ROCQ_STATUS_HIP_ERROR;
    if (rq_launch_scan_masses(d_hi, d_lo, nchunks, d_btot, d_tot, h->stream) != 0) return ROCQ_STATUS_HIP_ERROR;
    std::vector<uint64_t> all((size_t)2 * nranks, 0);
    if (nranks == 1 || group) {
        uint64_t mine[2];
        RQ_OK(rq_engine_fetch(h, d_tot, mine, sizeof mine));
        RQ_OK(allgather_u64(h, *this, mine, 2, all));
    } else {
        uint64_t* recv = d_gather + 64;
        RQ_NCCL(g_nccl.AllGather(d_tot, recv, 2, ncclUint64, (ncclComm_t)comm, h->stream));
        RQ_OK(rq_engine_fetch(h, recv, all.data(), all.size() * sizeof(uint64_t)));
    }
    u128 total = 0, win = 0;
    for (int r = 0; r < nranks; ++r) {
        const u128 m = ((u128)all[2 * r] << 64) | all[2 * r + 1];
        if (r < rank) win += m;
        total += m;
    }
    if (total == 0) return ROCQ_STATUS_FAILURE;
    const uint64_t tot4[4] = {(uint64_t)(total >> 64), (uint64_t)total, (uint64_t)(win >> 64), (uint64_t)win};
    RQ_CU(cudaMemcpyAsync(d_tot, tot4, sizeof tot4, cudaMemcpyHostToDevice, h->stream));     // pageable source: staged before the call returns
    rq_shot_map map{};
    map.high_base = (uint64_t)rank << n_local;
    map.miss = 0;
    map.nm = nm;
    for (unsigned j = 0; j < nm; ++j) map.pos[j] = (uint8_t)plan.map[measured[j]];           // physical position of every measured qubit
    if (rq_launch_sample(h->d_state, n, cb, d_hi, d_lo, nchunks, d_tot, h->seed, h->draws, shots, 0, &map, d_idx, h->stream) != 0)
        return ROCQ_STATUS_HIP_ERROR;
    h->draws++;
    h->stats.kernelLaunches += nchunks > 256 ? 5 : 4;
    if (group) {                                                       // one process: the ranks' words meet in host memory
        std::vector<uint64_t>& mine = group->shots[rank];
        mine.resize(shots);
        RQ_CU(cudaMemcpyAsync(mine.data(), d_idx, (size_t)shots * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
        RQ_CU(cudaStreamSynchronize(h->stream));
        group->barrier();
        for (unsigned s = 0; s < shots; ++s) {
            uint64_t v = 0;
            for (int r = 0; r < nranks; ++r) v |= group->shots[r][s];
            out[s] = v;
        }
        group->barrier();
        return ROCQ_STATUS_SUCCESS;
    }
    if (nranks > 1) RQ_NCCL(g_nccl.AllReduce(d_idx, d_idx, shots, ncclUint64, ncclSum, (ncclComm_t)comm, h->stream));
    RQ_CU(cudaMemcpyAsync(out, d_idx, (size_t)shots * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    RQ_CU(cudaStreamSynchronize(h->stream));
    return ROCQ_STATUS_SUCCESS;
}

}  // namespace rq

extern "C" {

// hipStateVec.h:84-104.  A handle that is a rank of a multi-process job (rocsvxDistInit) allocates its slice.  Any other
// handle follows the reference's model -- one process, one handle, all GPUs (MULTI_GPU_GUIDE.md:11-17): the state is sharded
// over the visible devices (the largest power of two of them; rocsvxDistSetRanks / ROCQ_NUM_GPUS choose otherwise) and the
// handle becomes the front of a group of per-device ranks (group.h).  With one device it is a plain single-slice state.
rocqStatus_t rocsvAllocateDistributedState(rocsvHandle_t h, unsigned totalNumQubits) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    if (totalNumQubits > 60) return ROCQ_STATUS_INVALID_VALUE;
    if (!h->dist.inited || h->group) {
        int want = h->wantRanks;
        if (want <= 0) if (const char* e = getenv("ROCQ_NUM_GPUS")) want = atoi(e);
        if (want <= 0) { if (cudaGetDeviceCount(&want) != cudaSuccess) { cudaGetLastError(); want = 1; } }
        int P = 1;
        while (2 * P <= want) P *= 2;
        while (P > 1 && (1ull << (totalNumQubits > 2 ? totalNumQubits - 2 : 0)) < (uint64_t)P) P /= 2;    // keep >= 2 local qubits per slice
        if (h->group && h->group->P != P) rq_group_destroy(h);
        if (P > 1 && !h->group) {
            rocsvFreeState(h);                                 // the front handle holds no state of its own
            const rocqStatus_t s = rq_group_create(h, P);
            if (s != ROCQ_STATUS_SUCCESS) return s;
        }
        if (h->group) return h->group->run([&](rocsvInternalHandle* c, int) { return c->dist.allocate(c, totalNumQubits); });
    }
    return h->dist.allocate(h, totalNumQubits);
}
rocqStatus_t rocsvInitializeDistributedState(rocsvHandle_t h) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    if (h->group) return h->group->run([&](rocsvInternalHandle* c, int) { return c->dist.initialize(c); });
    return h->dist.initialize(h);
}
rocqStatus_t rocsvxDistSetRanks(rocsvHandle_t h, int numRanks) {
    if (!h || numRanks < 0 || (numRanks & (numRanks - 1))) return ROCQ_STATUS_INVALID_VALUE;
    h->wantRanks = numRanks;
    return ROCQ_STATUS_SUCCESS;
}
rocqStatus_t rocsvxDistGetRankSlice(rocsvHandle_t h, int rank, int* device, rocComplex** d_slice) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    if (!h->group) {
        if (rank != 0) return ROCQ_STATUS_INVALID_VALUE;
        if (device) cudaGetDevice(device);
        if (d_slice) *d_slice = reinterpret_cast<rocComplex*>(h->d_state);
        return ROCQ_STATUS_SUCCESS;
    }
    if (rank < 0 || rank >= h->group->P || !h->group->child[(size_t)rank]) return ROCQ_STATUS_INVALID_VALUE;
    if (device) *device = h->group->device[(size_t)rank];
    if (d_slice) *d_slice = reinterpret_cast<rocComplex*>(h->group->child[(size_t)rank]->d_state);
    return ROCQ_STATUS_SUCCESS;
}
rocqStatus_t rocsvxDistGetUniqueId(void* id128) {
    if (!id128) return ROCQ_STATUS_INVALID_VALUE;
    if (!rq::g_nccl.load()) return ROCQ_STATUS_RCCL_ERROR;
    ncclUniqueId id;
    if (rq::g_nccl.GetUniqueId(&id) != ncclSuccess) return ROCQ_STATUS_RCCL_ERROR;
    memcpy(id128, &id, 128);
    return ROCQ_STATUS_SUCCESS;
}
rocqStatus_t rocsvxDistInit(rocsvHandle_t h, int rank, int numRanks, const void* id128) {
    if (!h || h->group) return ROCQ_STATUS_INVALID_VALUE;
    return h->dist.init(h, rank, numRanks, id128);
}
rocqStatus_t rocsvxDistGetInfo(rocsvHandle_t h, int* rank, int* numRanks, unsigned* numLocalQubits, rocComplex** d_localSlice) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    if (h->group && h->group->child[0]) {                          // the front of a single-process group answers for rank 0
        const rocqStatus_t s = rocsvxDistGetInfo(h->group->child[0], rank, numRanks, numLocalQubits, d_localSlice);
        if (rank) *rank = 0;
        return s;
    }
    if (rank) *rank = h->dist.rank;
    if (numRanks) *numRanks = h->dist.nranks;
    if (numLocalQubits) *numLocalQubits = h->dist.n_local;
    if (d_localSlice) *d_localSlice = reinterpret_cast<rocComplex*>(h->d_state);
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
rocqStatus_t rocsvxDistPlanPeerSwap(unsigned numLocalQubits, int numRanks, int rank, const unsigned* localBits,
                                    const unsigned* globalBits, unsigned numPairs, rocsvxExchangeSeg* segs, size_t maxSegs,
                                    size_t* numSegs) {
    size_t c = 0;
    const rocqStatus_t s = rocsvxDistPlanExchange(numLocalQubits, numRanks, rank, localBits, globalBits, numPairs, nullptr, 0, &c);   // argument checks
    if (s != ROCQ_STATUS_SUCCESS) return s;
    c = rq::plan_peer_swap(numLocalQubits, numRanks, rank, localBits, globalBits, numPairs, segs, maxSegs);
    if (numSegs) *numSegs = c;
    return ROCQ_STATUS_SUCCESS;
}

}  // extern "C"
