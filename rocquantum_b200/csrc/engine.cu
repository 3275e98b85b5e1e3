// rocquantum_b200/csrc/engine.cu -- host side of the drop-in C ABI (include/hipStateVec.h).
//
// Mirrors, entry point by entry point, /root/reference/rocquantum/src/hipStateVec/hipStateVec.cpp
// (handle :62-68, validation and status codes :100-186, lifecycle :190-272, gates :276-687, readback
// :691-730) and implements the 17 entry points the reference only declares.  Every gate becomes a
// HostOp (host_ops.h); eager mode runs it as a one-op sweep, fusion mode queues it and cuts the queue
// into fused sweeps at the next flush.  There is no CPU fallback: if CUDA is unusable, rocsvCreate
// fails with ROCQ_STATUS_HIP_ERROR.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hipStateVec.h"
#include <cuda.h>
#include <cuda_fp16.h>

#include "engine.h"
#include "gate_convert.h"

using rq::cd;
using rq::HostOp;

static_assert(sizeof(rocComplex) == sizeof(rq_cplx), "rocComplex layout");
static_assert(sizeof(rq_program_large) <= 32764, "kernel parameter limit");


namespace {

typedef rocsvInternalHandle H;

// Single-process multi-GPU (group.h): a call on the front handle of a group runs on every rank's own handle at once.
// `c` = the rank's handle, `r` = its rank.  d_state arguments are dropped: a distributed state is always the handle's own
// (MULTI_GPU_GUIDE.md: "multi-GPU always NULL").
#define RQ_FWD(h, expr)                                                                                        \
    do {                                                                                                       \
        if ((h) && (h)->group)                                                                                 \
            return (h)->group->run([&](rocsvInternalHandle* c, int r) -> rocqStatus_t { (void)r; (void)c; return (expr); }); \
    } while (0)
// an output every rank computes identically: rank 0 writes the caller's memory, the others a scratch of their own
template <typename T>
struct PerRank {
    T* user;
    std::vector<std::vector<T>> scratch;
    PerRank(T* u, int ranks, size_t count = 1) : user(u), scratch((size_t)ranks, std::vector<T>(u ? count : 0)) {}
    T* at(int r) { return (r == 0 || !user) ? user : scratch[(size_t)r].data(); }
};

inline rocqStatus_t cuda_status(int e, const char* what) {
    if (e == (int)cudaSuccess) return ROCQ_STATUS_SUCCESS;
    fprintf(stderr, "hipStateVec(B200): %s failed: %s\n", what, cudaGetErrorString((cudaError_t)e));
    return ROCQ_STATUS_HIP_ERROR;
}
#define RQ_CUDA(call, what)                                            \
    do {                                                               \
        const rocqStatus_t _s = cuda_status((int)(call), what);        \
        if (_s != ROCQ_STATUS_SUCCESS) return _s;                      \
    } while (0)

inline rq_cplx* resolve(H* h, rocComplex* ext) {          // hipStateVec.cpp:80-85
    if (ext) return reinterpret_cast<rq_cplx*>(ext);
    return h ? h->d_state : nullptr;
}

rq::PlanLimits limits_for(const H* h, bool large) {
    rq::PlanLimits L;
    L.tile_bits = h->tileBits;
    L.budget = h->budget;
    if (large) { L.max_ops = sizeof(rq_program_large::ops) / sizeof(rq_tile_op); L.pool_cplx = sizeof(rq_program_large::pool) / sizeof(rq_cplx); }
    else { L.max_ops = sizeof(rq_program_small::ops) / sizeof(rq_tile_op); L.pool_cplx = sizeof(rq_program_small::pool) / sizeof(rq_cplx); }
    if (h->dist.active()) L.never_resident = h->dist.global_mask();
    return L;
}

rocqStatus_t run_block(H* h, rq_cplx* state, unsigned n, const std::vector<unsigned>& blk, const std::vector<cd>& U, int unitary = -1);
rocqStatus_t launch_plan(H* h, rq_cplx* state, unsigned n, const rq::SweepPlan& sp, const std::vector<HostOp>& seg, bool large);

rocqStatus_t launch_plan(H* h, rq_cplx* state, unsigned n, const rq::SweepPlan& sp, const std::vector<HostOp>& seg, bool large) {
    int e;
    const void* prog = nullptr;
    size_t prog_bytes = 0;
    const rq::NvtxRange nvtx("rocq/tile_sweep");
    const rq_trace_scope trace(h, "tile_sweep", (unsigned)sp.ops.size());
    if (large) {
        static thread_local rq_program_large P;
        if (!rq::build_program(P, sp, seg, n, h->batchSize, h->dist.high_base())) return ROCQ_STATUS_FAILURE;
        e = rq_launch_sweep_large(state, &P, h->stream);
        prog = &P; prog_bytes = sizeof(P);
    } else {
        static thread_local rq_program_small P;
        if (!rq::build_program(P, sp, seg, n, h->batchSize, h->dist.high_base())) return ROCQ_STATUS_FAILURE;
        e = rq_launch_sweep_small(state, &P, h->stream);
        prog = &P; prog_bytes = sizeof(P);
    }
    RQ_CUDA(e, "tile sweep launch");
    h->stats.h2dBytes += prog_bytes;
    h->stats.kernelLaunches++;
    h->stats.sweeps++;
    h->stats.opsExecuted += sp.ops.size();
    if (h->recording) {                                  // plan cache: keep the program (it has no device pointers unless ext)
        const bool ext = large ? static_cast<const rq_program_large*>(prog)->hdr.ext_matrix != nullptr
                               : static_cast<const rq_program_small*>(prog)->hdr.ext_matrix != nullptr;
        if (ext) h->recordingValid = false;
        rocsvCachedStep st;
        st.large = large;
        st.ops = (unsigned)sp.ops.size();
        st.prog.assign(static_cast<const unsigned char*>(prog), static_cast<const unsigned char*>(prog) + prog_bytes);
        h->recording->push_back(std::move(st));
    }
    return ROCQ_STATUS_SUCCESS;
}

bool op_is_unitary(const HostOp& o) {
    const unsigned k = (unsigned)o.targets.size(), D = 1u << k;
    if (o.kind == HostOp::DIAG || o.kind == HostOp::DIAGP) {
        for (const cd& d : o.data) if (std::abs(std::abs(d) - 1.0) > 1e-9) return false;
        return true;
    }
    if (o.kind != HostOp::DENSE) return true;
    if (o.ext || o.data.size() != (size_t)D * D) return false;
    for (unsigned a = 0; a < D; ++a)
        for (unsigned b = a; b < D; ++b) {
            cd s(0.0, 0.0);
            for (unsigned r = 0; r < D; ++r) s += std::conj(o.data[r + (size_t)D * a]) * o.data[r + (size_t)D * b];
            if (std::abs(s - (a == b ? cd(1.0, 0.0) : cd(0.0, 0.0))) > 1e-9) return false;
        }
    return true;
}

// Mixed execution of one fused op list (rq::plan_mixed): wherever enough arithmetic can be folded into six qubits
// (positions >= 5) it becomes ONE 64x64 unitary applied by the tensor-core block sweep; everything else goes through
// ordinary tile sweeps.
rocqStatus_t run_ops_with_blocks(H* h, rq_cplx* state, unsigned n, const std::vector<HostOp>& seg) {
    rq::BlockLimits BL;
    BL.min_cost = h->blockMinCost;
    BL.batch = h->batchSize;
    static const bool prof = getenv("ROCQ_HOST_PROFILE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    double t_sweep = 0, t_build = 0, t_launch = 0;
    const auto tp0 = now();
    const std::vector<rq::MixedStep> steps = rq::plan_mixed(seg, n, limits_for(h, true), BL);
    const auto tp1 = now();
    struct Report { bool on; double& a; double& b; double& c; double plan; ~Report() { if (on) fprintf(stderr, "[host profile] plan %.2f ms, sweeps %.2f ms, block build %.2f ms, block launch %.2f ms\n", plan, a, b, c); } } report{prof, t_sweep, t_build, t_launch, ms(tp0, tp1)};
    for (const rq::MixedStep& st : steps) {
        const auto t0 = now();
        if (!st.block) {
            const rocqStatus_t s = launch_plan(h, state, n, st.sweep, seg, true);
            t_sweep += ms(t0, now());
            if (s != ROCQ_STATUS_SUCCESS) return s;
            continue;
        }
        std::vector<cd> U(64 * 64, cd(0.0, 0.0));
        for (unsigned c = 0; c < 64; ++c) U[c + 64u * c] = cd(1.0, 0.0);
        bool unitary = true;
        for (int idx : st.ops) {                               // U <- op * U
            unitary = unitary && op_is_unitary(seg[idx]);
            rq::apply_small_columns(seg[idx], st.blk, U.data(), 64);
        }
        const auto t1 = now();
        const rocqStatus_t s = run_block(h, state, n, st.blk, U, unitary ? 1 : 0);
        t_build += ms(t0, t1);
        t_launch += ms(t1, now());
        if (s != ROCQ_STATUS_SUCCESS) return s;
        h->stats.opsExecuted += st.ops.size();
        if (h->recording && !h->recording->empty() && h->recording->back().block) h->recording->back().ops = (unsigned)st.ops.size();
    }
    return ROCQ_STATUS_SUCCESS;
}

// run ops (already validated) on `state`: dense/diag ops wider than the tile kernel handles go through
// the gather kernel, everything else through planned sweeps.
rocqStatus_t run_ops(H* h, rq_cplx* state, unsigned n, const std::vector<HostOp>& ops_in, bool fused) {
    // a slice of a distributed state: controls and diagonal factors on rank bits are constants of this rank
    std::vector<HostOp> resolved;
    if (h->dist.active() && h->dist.global_mask()) resolved = rq::specialize_for_rank(ops_in, n, h->dist.high_base());
    const std::vector<HostOp>& ops = (h->dist.active() && h->dist.global_mask()) ? resolved : ops_in;
    size_t i = 0;
    while (i < ops.size()) {
        // segment of tile-capable ops
        size_t j = i;
        while (j < ops.size() && ops[j].targets.size() <= 4) ++j;
        if (j > i) {
            std::vector<HostOp> seg(ops.begin() + i, ops.begin() + j);
            if (fused && seg.size() > 1) {
                // ladders of controlled phases are merged BEFORE algebraic fusion, which would otherwise promote their first
                // member to a dense 4x4 to absorb the preceding one-qubit gate
                if (h->mergeDiagonals) seg = rq::merge_diagonals(rq::push_x_forward(seg));
                seg = rq::fuse_algebraic(seg, n, h->dist.active() ? h->dist.global_mask() : 0ull);
            }
            const bool tc = h->tcBlocks > 0 || (h->tcBlocks < 0 && n >= RQ_BLOCK_AUTO_QUBITS);
            // (a distributed slice forms blocks on its local qubits; ops that touch rank bits stay in the ordinary sweeps)
            if (fused && tc && sizeof(rq_real) == 4 && n >= 13 && seg.size() > 1) {
                const rocqStatus_t s = run_ops_with_blocks(h, state, n, seg);
                if (s != ROCQ_STATUS_SUCCESS) return s;
                i = j;
                continue;
            }
            // a lone op normally travels in the small program; a 4-qubit host matrix needs the large pool
            const bool large = seg.size() > 1 || rq::pool_need(seg[0]) > sizeof(rq_program_small::pool) / sizeof(rq_cplx);
            const rq::PlanLimits L = limits_for(h, large);
            const std::vector<rq::SweepPlan> plans = rq::plan_sweeps(seg, n, L);
            for (const rq::SweepPlan& sp : plans) {
                const rocqStatus_t s = launch_plan(h, state, n, sp, seg, large);
                if (s != ROCQ_STATUS_SUCCESS) return s;
            }
            i = j;
        }
        if (i < ops.size()) {                 // one wide op through the gather kernel
            const HostOp& o = ops[i];
            const unsigned k = (unsigned)o.targets.size();
            if (k > 10) return ROCQ_STATUS_NOT_IMPLEMENTED;
            // complex64, at most six qubits in all (targets + controls): ONE tensor-core block sweep, the matrix embedded in a
            // 64x64 block (a 5-qubit matrix is padded with an idle qubit, controls become block qubits).  HBM-bound like any
            // block pass instead of the gather kernel's strided scalar loads.
            if (sizeof(rq_real) == 4 && n >= 13 && o.kind == HostOp::DENSE && h->tcBlocks != 0) {
                const uint64_t Q = o.qubits();
                const unsigned nq = (unsigned)__builtin_popcountll(Q);
                if (nq <= RQ_BLOCK_QUBITS && (n >= 64 || !(Q >> n))) {
                    uint64_t bm = 0;
                    if (nq == RQ_BLOCK_QUBITS) {
                        if (rq::block_supported(Q, n, h->batchSize)) bm = Q;
                    } else {                                                   // five qubits: any idle sixth one the kernel's tile geometry accepts
                        for (unsigned q = 0; q < n && !bm; ++q)
                            if (!((Q >> q) & 1ull) && rq::block_supported(Q | (1ull << q), n, h->batchSize)) bm = Q | (1ull << q);
                    }
                    if (bm) {
                        HostOp host = o;
                        if (o.ext) {                                          // device matrix of an eager rocsvApplyMatrix call
                            const size_t D = (size_t)1 << k;
                            std::vector<rq_cplx> hm(D * D);
                            RQ_CUDA(cudaMemcpyAsync(hm.data(), o.ext, D * D * sizeof(rq_cplx), cudaMemcpyDeviceToHost, h->stream), "matrix D2H");
                            RQ_CUDA(cudaStreamSynchronize(h->stream), "sync");
                            host.ext = nullptr;
                            host.data.resize(D * D);
                            for (size_t e = 0; e < D * D; ++e) host.data[e] = cd(hm[e].x, hm[e].y);
                        }
                        std::vector<unsigned> blk;
                        for (unsigned q = 0; q < n; ++q) if ((bm >> q) & 1ull) blk.push_back(q);
                        const std::vector<cd> U = rq::to_matrix(host, blk);
                        const rocqStatus_t s = run_block(h, state, n, blk, U, -1);
                        if (s != ROCQ_STATUS_SUCCESS) return s;
                        h->stats.opsExecuted++;
                        ++i;
                        continue;
                    }
                }
            }
            h->recordingValid = false;                   // the gather kernel's matrix upload is not replayable
            const rq_cplx* dm = reinterpret_cast<const rq_cplx*>(o.ext);
            rq::StreamBuf tmp(h->stream, h->pool);     // freed (stream-ordered) on every path out of this scope
            if (!dm) {                        // host matrix (or diagonal): upload, stream-ordered
                const size_t D = (size_t)1 << k;
                std::vector<rq_cplx> hm(D * D, rq_cplx{0, 0});
                if (o.kind == HostOp::DIAG) for (size_t d = 0; d < D; ++d) { hm[d + d * D].x = (rq_real)o.data[d].real(); hm[d + d * D].y = (rq_real)o.data[d].imag(); }
                else for (size_t e = 0; e < D * D; ++e) { hm[e].x = (rq_real)o.data[e].real(); hm[e].y = (rq_real)o.data[e].imag(); }
                RQ_CUDA(tmp.alloc(D * D * sizeof(rq_cplx)), "cudaMallocAsync");
                RQ_CUDA(cudaMemcpyAsync(tmp.p, hm.data(), D * D * sizeof(rq_cplx), cudaMemcpyHostToDevice, h->stream), "matrix upload");
                RQ_CUDA(cudaStreamSynchronize(h->stream), "sync");     // hm goes out of scope
                h->stats.h2dBytes += D * D * sizeof(rq_cplx);
                dm = tmp.as<rq_cplx>();
            }
            // controls on rank bits of a distributed state select whole slices: resolve them here
            const uint64_t lmask = n >= 64 ? ~0ull : ((1ull << n) - 1ull), gctl = o.cmask & ~lmask;
            if ((h->dist.high_base() & gctl) == gctl) {
                RQ_CUDA(rq_launch_gather(state, n, h->batchSize, o.targets.data(), k, o.cmask & lmask, dm, h->stream), "gather launch");
                h->stats.kernelLaunches++;
            }
            h->stats.opsExecuted++;
            ++i;
        }
    }
    return ROCQ_STATUS_SUCCESS;
}

// ---- tensor-core 6-qubit blocks (block_sweep.cu) -------------------------------------------------------------------
// U: 64x64 complex column-major, index bit b <-> b-th smallest block position.  Writes Re U and Im U, each split in two fp16
// terms (hi, lo), as 64 x 64 K-major core-matrix operands in the order the kernel's descriptors expect:
// Re hi | Re lo | Im hi | Im lo, 8 KB each.
void build_block_terms(const std::vector<cd>& U, std::vector<uint16_t>& out) {
    out.assign(RQ_BLOCK_UBYTES / 2, 0);
    for (unsigned o = 0; o < 64; ++o)
        for (unsigned k = 0; k < 64; ++k) {
            const cd u = U[o + 64u * k];
            const size_t off = ((o & 7u) * 16u + (k >> 3) * 128u + (o >> 3) * 1024u + (k & 7u) * 2u) / 2u;
            for (int part = 0; part < 2; ++part) {
                const float v = (float)(part ? u.imag() : u.real());
                const __half hi = __float2half_rn(v);
                const __half lo = __float2half_rn(v - __half2float(hi));
                memcpy(&out[(size_t)(2 * part) * 4096 + off], &hi, 2);
                memcpy(&out[(size_t)(2 * part + 1) * 4096 + off], &lo, 2);
            }
        }
}

// Tensor map over the state (elements = 8-byte amplitudes) whose box is exactly one tile of the block sweep, in the
// shared-memory order and with the 128-byte swizzle rq::block_layout describes.
bool build_block_tensor_map(const rq_cplx* state, const rq::BlockLayout& L, CUtensorMap* tm) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const encode_fn encode = []() -> encode_fn {       // looked up once; initialisation of a local static is thread-safe
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            return reinterpret_cast<encode_fn>(fn);
        return nullptr;
    }();
    if (!encode) return false;
    cuuint64_t dims[5], strides[5];
    cuuint32_t box[5], estr[5];
    for (unsigned d = 0; d < L.rank; ++d) { dims[d] = L.dims[d]; strides[d] = L.strides[d]; box[d] = L.box[d]; estr[d] = 1; }
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, L.rank, const_cast<rq_cplx*>(state), dims, strides + 1, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// launch one block sweep: block positions (ascending), matrix U over them
// unitary: 1 / 0 when the caller knows, -1 = check U here
rocqStatus_t run_block(H* h, rq_cplx* state, unsigned n, const std::vector<unsigned>& blk, const std::vector<cd>& U, int unitary) {
    if (sizeof(rq_real) != 4) return ROCQ_STATUS_NOT_IMPLEMENTED;
    if (n < 13 || blk.size() != RQ_BLOCK_QUBITS) return ROCQ_STATUS_INVALID_VALUE;
    const rq::NvtxRange nvtx("rocq/block_sweep");
    rq_block_params P{};
    P.n = n; P.T = 13; P.ntiles = (uint64_t)h->batchSize << (n - 13);
    uint64_t bm = 0;
    for (unsigned b = 0; b < 6; ++b) { P.blk[b] = (uint8_t)blk[b]; bm |= 1ull << blk[b]; }
    rq::BlockLayout L;
    CUtensorMap tm;
    if (!rq::block_layout(bm, n, h->batchSize, L) || !build_block_tensor_map(state, L, &tm)) return ROCQ_STATUS_NOT_IMPLEMENTED;
    memcpy(P.col, L.col, 7); memcpy(P.res, L.res, 13); memcpy(P.tbits, L.tbits, 5); memcpy(P.lp_blk, L.lp_blk, 6); memcpy(P.lp_col, L.lp_col, 7);
    P.trank = L.rank;
    // unitary?  (then every tile column keeps its norm, which the kernel restores exactly)
    double dev = unitary == 1 ? 0.0 : unitary == 0 ? 1.0 : 0.0;
    for (unsigned a = 0; a < 64 && unitary < 0; ++a)
        for (unsigned b = a; b < 64; ++b) {
            cd s(0.0, 0.0);
            for (unsigned r = 0; r < 64; ++r) s += std::conj(U[r + 64u * a]) * U[r + 64u * b];
            dev = std::max(dev, std::abs(s - (a == b ? cd(1.0, 0.0) : cd(0.0, 0.0))));
        }
    P.renorm = dev < 1e-9 ? 1u : 0u;
    if (const char* dbgenv = getenv("ROCQ_BLOCK_DEBUG")) { P.pad = (uint32_t)atoi(dbgenv); if (P.pad & 8u) P.renorm = 0; }
    // fp16 inputs: |amplitude| <= 1 for a normalised state, typical magnitude 2^(-n/2); 2^14 keeps a factor 4 of headroom
    P.scale = ldexpf(1.0f, (int)std::min(14u, n / 2));
    std::vector<uint16_t> terms;
    build_block_terms(U, terms);
    void* d_terms = nullptr;
    const bool keep = h->recording != nullptr && P.pad == 0;                                     // plan cache: the entry owns the buffer
    if (keep) RQ_CUDA(cudaMalloc(&d_terms, RQ_BLOCK_UBYTES), "block terms alloc (cached)");
    else RQ_CUDA(rq::pool_alloc(&d_terms, RQ_BLOCK_UBYTES, h->pool, h->stream), "block terms alloc");
    // pageable source: cudaMemcpyAsync stages it before returning, so `terms` may go out of scope
    struct TermsGuard {                                     // an early return must not leak the operand buffer
        void*& p; bool cached; cudaStream_t s; bool armed = true;
        ~TermsGuard() { if (armed && p) { if (cached) cudaFree(p); else cudaFreeAsync(p, s); p = nullptr; } }
    } guard{d_terms, keep, h->stream};
    RQ_CUDA(cudaMemcpyAsync(d_terms, terms.data(), RQ_BLOCK_UBYTES, cudaMemcpyHostToDevice, h->stream), "block terms upload");
    {
        const rq_trace_scope trace(h, "block_sweep");
        RQ_CUDA(rq_launch_block_sweep(state, &P, d_terms, &tm, h->stream), "block sweep launch");
    }
    guard.armed = false;
    if (keep) {
        rocsvCachedStep st;
        st.block = true;
        st.bp = P;
        memcpy(st.tmap, &tm, sizeof st.tmap);
        st.d_terms = d_terms;
        h->recording->push_back(std::move(st));
    } else {
        if (h->recording) h->recordingValid = false;
        RQ_CUDA(cudaFreeAsync(d_terms, h->stream), "block terms free");
    }
    h->stats.kernelLaunches++;
    h->stats.sweeps++;
    h->stats.blockSweeps++;
    h->stats.h2dBytes += RQ_BLOCK_UBYTES;
    return ROCQ_STATUS_SUCCESS;
}

// ---- plan cache of rocsvxApplyCircuit -------------------------------------------------------------------------------
void free_steps(H* h, std::vector<rocsvCachedStep>& steps) {
    bool any = false;
    for (rocsvCachedStep& st : steps) any = any || st.d_terms;
    if (any && h->stream) cudaStreamSynchronize(h->stream);       // a launch may still read the operand terms
    for (rocsvCachedStep& st : steps) if (st.d_terms) { cudaFree(st.d_terms); st.d_terms = nullptr; }
    steps.clear();
}
void drop_cache(H* h) {
    free_steps(h, h->cache.steps);
    h->cache.valid = false;
}
// two independent 64-bit word hashes over everything that determines the launches
struct Hash2 {
    uint64_t a = 0xcbf29ce484222325ull, b = 0x9e3779b97f4a7c15ull;
    size_t bytes = 0;
    void word(uint64_t w) {
        a = (a ^ w) * 0x100000001b3ull; a ^= a >> 29;
        b = (b + w) * 0xff51afd7ed558ccdull; b ^= b >> 32;
    }
    void mem(const void* p, size_t n) {
        const unsigned char* c = static_cast<const unsigned char*>(p);
        bytes += n;
        for (; n >= 8; n -= 8, c += 8) { uint64_t w; memcpy(&w, c, 8); word(w); }
        if (n) { uint64_t w = 0; memcpy(&w, c, n); word(w ^ ((uint64_t)n << 56)); }
    }
};
Hash2 circuit_key(const H* h, const rq_cplx* state, unsigned n, const rocsvxGateOp* ops, size_t numOps) {
    Hash2 k;
    k.word((uint64_t)(uintptr_t)state); k.word(n); k.word(h->batchSize); k.word((uint64_t)(int64_t)h->tcBlocks); k.word(h->mergeDiagonals);
    k.word(h->tileBits); k.mem(&h->blockMinCost, sizeof(double)); k.mem(&h->budget, sizeof(double)); k.word(numOps);
    for (size_t i = 0; i < numOps; ++i) {
        const rocsvxGateOp& g = ops[i];
        k.word(((uint64_t)(uint32_t)g.kind << 32) | g.numTargets);
        k.mem(g.targets, sizeof g.targets);
        k.word(g.controlMask);
        k.mem(&g.theta, sizeof(double));
        if (g.kind == ROCSVX_MATRIX && g.matrix && g.numTargets <= 8) k.mem(g.matrix, sizeof(double) * 2 * ((size_t)1 << (2 * g.numTargets)));
    }
    return k;
}
rocqStatus_t replay_cache(H* h, rq_cplx* state) {
    const rq::NvtxRange nvtx("rocq/replay_cached_plan");
    for (const rocsvCachedStep& st : h->cache.steps) {
        if (st.block) {
            RQ_CUDA(rq_launch_block_sweep(state, &st.bp, st.d_terms, st.tmap, h->stream), "block sweep launch (cached)");
            h->stats.blockSweeps++;
            h->stats.opsExecuted += st.ops;
        } else {
            const int e = st.large ? rq_launch_sweep_large(state, reinterpret_cast<const rq_program_large*>(st.prog.data()), h->stream)
                                   : rq_launch_sweep_small(state, reinterpret_cast<const rq_program_small*>(st.prog.data()), h->stream);
            RQ_CUDA(e, "tile sweep launch (cached)");
            h->stats.h2dBytes += st.prog.size();
            h->stats.opsExecuted += st.ops;
        }
        h->stats.kernelLaunches++;
        h->stats.sweeps++;
    }
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t flush(H* h) {
    if (h->queue.empty()) return ROCQ_STATUS_SUCCESS;
    std::vector<HostOp> q;
    q.swap(h->queue);
    cudaEventRecord(h->ev0, h->stream);
    const rocqStatus_t s = run_ops(h, h->queue_state, h->queue_n, q, true);
    cudaEventRecord(h->ev1, h->stream);
    h->stats.lastSweepMs = -1.0;         // resolved lazily in rocsvxGetStats
    return s;
}

#define RQ_OK(call) do { const rocqStatus_t _s = (call); if (_s != ROCQ_STATUS_SUCCESS) return _s; } while (0)

// a call that overwrites `state`: gates queued for THIS buffer are dead and dropped, gates queued for another buffer run first
rocqStatus_t discard_or_flush_queue(H* h, const rq_cplx* state) {
    if (h->queue.empty()) return ROCQ_STATUS_SUCCESS;
    if (h->queue_state == state) { h->queue.clear(); return ROCQ_STATUS_SUCCESS; }
    return flush(h);
}

// ---- state import / export through page-locked staging (the reference: one synchronous hipMemcpy into
//      pageable memory, hipStateVec.cpp:691-706) ------------------------------------------------------------------------
// A pageable destination cannot be DMA'd into, so a plain cudaMemcpy stages it through small driver buffers on one host
// thread.  Here the transfer is cut into STAGE_BYTES chunks that alternate between two page-locked buffers of the handle: the
// DMA of chunk c runs while the host (several threads) drains chunk c-1 into the caller's memory.  A destination that is
// already page-locked (cudaHostAlloc / cudaHostRegister, e.g. the handle's own rocsvEnsurePinnedBuffer) is one DMA.
constexpr size_t STAGE_BYTES = (size_t)64 << 20;
void parallel_copy(void* dst, const void* src, size_t bytes) {
    static const unsigned T = std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
    if (T == 1 || bytes < ((size_t)4 << 20)) { memcpy(dst, src, bytes); return; }
    const size_t per = ((bytes / T) + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; ++t) {
        const size_t b = std::min(bytes, per * t), e = std::min(bytes, per * (t + 1));
        if (e > b) th.emplace_back([=] { memcpy((char*)dst + b, (const char*)src + b, e - b); });
    }
    memcpy(dst, src, std::min(bytes, per));
    for (std::thread& x : th) x.join();
}
bool is_page_locked(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}
rocqStatus_t ensure_stage(H* h) {
    for (int b = 0; b < 2; ++b) {
        if (!h->stage[b]) RQ_CUDA(cudaHostAlloc(&h->stage[b], STAGE_BYTES, cudaHostAllocDefault), "staging buffer");
        if (!h->stageEv[b]) RQ_CUDA(cudaEventCreateWithFlags(&h->stageEv[b], cudaEventDisableTiming), "staging event");
    }
    return ROCQ_STATUS_SUCCESS;
}
rocqStatus_t staged_d2h(H* h, void* dst, const void* src, size_t bytes) {
    if (bytes <= ((size_t)8 << 20) || is_page_locked(dst)) {
        RQ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream), "state D2H");
        RQ_CUDA(cudaStreamSynchronize(h->stream), "stream sync");
        return ROCQ_STATUS_SUCCESS;
    }
    RQ_OK(ensure_stage(h));
    const size_t nchunks = (bytes + STAGE_BYTES - 1) / STAGE_BYTES;
    auto len = [&](size_t c) { return std::min(STAGE_BYTES, bytes - c * STAGE_BYTES); };
    for (size_t c = 0; c <= nchunks; ++c) {
        if (c < nchunks) {                                   // buffer c&1 was drained two iterations ago
            RQ_CUDA(cudaMemcpyAsync(h->stage[c & 1], (const char*)src + c * STAGE_BYTES, len(c), cudaMemcpyDeviceToHost, h->stream), "state D2H chunk");
            RQ_CUDA(cudaEventRecord(h->stageEv[c & 1], h->stream), "staging event");
        }
        if (c >= 1) {
            RQ_CUDA(cudaEventSynchronize(h->stageEv[(c - 1) & 1]), "staging wait");
            parallel_copy((char*)dst + (c - 1) * STAGE_BYTES, h->stage[(c - 1) & 1], len(c - 1));
        }
    }
    return ROCQ_STATUS_SUCCESS;
}
rocqStatus_t staged_h2d(H* h, void* dst, const void* src, size_t bytes) {
    if (bytes <= ((size_t)8 << 20) || is_page_locked(src)) {
        RQ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream), "state H2D");
        RQ_CUDA(cudaStreamSynchronize(h->stream), "stream sync");
        return ROCQ_STATUS_SUCCESS;
    }
    RQ_OK(ensure_stage(h));
    const size_t nchunks = (bytes + STAGE_BYTES - 1) / STAGE_BYTES;
    for (size_t c = 0; c < nchunks; ++c) {
        const size_t len = std::min(STAGE_BYTES, bytes - c * STAGE_BYTES);
        if (c >= 2) RQ_CUDA(cudaEventSynchronize(h->stageEv[c & 1]), "staging wait");     // the DMA that last read this buffer
        parallel_copy(h->stage[c & 1], (const char*)src + c * STAGE_BYTES, len);
        RQ_CUDA(cudaMemcpyAsync((char*)dst + c * STAGE_BYTES, h->stage[c & 1], len, cudaMemcpyHostToDevice, h->stream), "state H2D chunk");
        RQ_CUDA(cudaEventRecord(h->stageEv[c & 1], h->stream), "staging event");
    }
    RQ_CUDA(cudaStreamSynchronize(h->stream), "stream sync");
    return ROCQ_STATUS_SUCCESS;
}

// validate-and-dispatch for one gate: reference order of checks (hipStateVec.cpp:280-282, 108-110)
rocqStatus_t submit(H* h, rocComplex* d_state, unsigned n, HostOp&& op) {
    rq_cplx* state = resolve(h, d_state);
    if (!state) return ROCQ_STATUS_INVALID_VALUE;
    const uint64_t Q = op.qubits();
    if (n < 64 && (Q >> n) != 0) return ROCQ_STATUS_INVALID_VALUE;
    if (h->dist.active()) {
        const rocqStatus_t s = h->dist.localize(h, op);     // may exchange slices; rewrites op to physical positions
        if (s != ROCQ_STATUS_SUCCESS) return s;
        n = h->dist.num_local();
    }
    h->stats.gatesSubmitted++;
    if (h->fusion) {
        if (!h->queue.empty() && (h->queue_state != state || h->queue_n != n)) {
            const rocqStatus_t s = flush(h);
            if (s != ROCQ_STATUS_SUCCESS) return s;
        }
        h->queue_state = state;
        h->queue_n = n;
        h->queue.push_back(std::move(op));
        return ROCQ_STATUS_SUCCESS;
    }
    std::vector<HostOp> one;
    one.push_back(std::move(op));
    return run_ops(h, state, n, one, false);
}

inline bool valid_q(unsigned q, unsigned n) { return q < n; }     // hipStateVec.cpp:96-98

rocqStatus_t single(H* h, rocComplex* d, unsigned n, unsigned t, HostOp&& op) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    if (!resolve(h, d)) return ROCQ_STATUS_INVALID_VALUE;
    if (!valid_q(t, n)) return ROCQ_STATUS_INVALID_VALUE;
    return submit(h, d, n, std::move(op));
}
rocqStatus_t two(H* h, rocComplex* d, unsigned n, unsigned a, unsigned b, HostOp&& op) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    if (!resolve(h, d)) return ROCQ_STATUS_INVALID_VALUE;
    if (!valid_q(a, n) || !valid_q(b, n) || a == b) return ROCQ_STATUS_INVALID_VALUE;
    return submit(h, d, n, std::move(op));
}

const cd I1(0.0, 1.0);

// read back `count` doubles / u64 from device scratch (synchronises the stream)
rocqStatus_t fetch(H* h, const void* dsrc, void* hdst, size_t bytes) {
    RQ_CUDA(cudaMemcpyAsync(h->h_scratch, dsrc, bytes, cudaMemcpyDeviceToHost, h->stream), "scratch D2H");
    RQ_CUDA(cudaStreamSynchronize(h->stream), "stream sync");
    memcpy(hdst, h->h_scratch, bytes);
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t pauli_expect(H* h, rq_cplx* state, unsigned n, uint64_t xm, uint64_t zm, unsigned ny, double* result) {
    const rq::NvtxRange nvtx("rocq/pauli_expectation");
    const unsigned nb = rq_reduce_blocks();
    RQ_CUDA(rq_launch_pauli_expect(state, n, xm, zm, ny, h->d_partials, nb, h->d_partials + nb, h->stream), "expectation launch");
    h->stats.kernelLaunches += 2;
    rocqStatus_t s = fetch(h, h->d_partials + nb, result, sizeof(double));
    if (s == ROCQ_STATUS_SUCCESS && h->dist.active()) s = h->dist.allreduce_sum(h, result, 1);
    return s;
}

typedef unsigned __int128 u128;
inline double u128_to_double(u128 v) { return (double)(uint64_t)(v >> 64) * 0x1p64 + (double)(uint64_t)v; }
inline u128 mul_u53(u128 S, uint64_t U) {
    const u128 A = (u128)(uint64_t)S * U, B = (u128)(uint64_t)(S >> 64) * U;
    return (B << 11) + (A >> 53);
}
void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
uint64_t uniform53(uint64_t seed, uint64_t call, uint64_t shot) {
    const uint32_t ctr[4] = {(uint32_t)shot, (uint32_t)(shot >> 32), (uint32_t)call, (uint32_t)(call >> 32)};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t x[4];
    philox4x32_10(ctr, key, x);
    return (((uint64_t)x[0] << 32) | x[1]) >> 11;
}

}  // namespace

// ---- single-process multi-GPU group (group.h) ----------------------------------------------------------------------
void rq_group_destroy(rocsvInternalHandle* h) {
    if (!h || !h->group) return;
    rocsvGroup* g = h->group;
    h->group = nullptr;
    g->run([&](H* c, int r) { const rocqStatus_t s = c ? rocsvDestroy(c) : ROCQ_STATUS_SUCCESS; g->child[(size_t)r] = nullptr; return s; });
    delete g;
}
// one worker thread + one ordinary handle per slice; the ranks inherit the front handle's settings
rocqStatus_t rq_group_create(rocsvInternalHandle* h, int ranks) {
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1) { cudaGetLastError(); return ROCQ_STATUS_HIP_ERROR; }
    rocsvGroup* g = new rocsvGroup(ranks, visible);
    const rocqStatus_t s = g->run([&](H*, int r) -> rocqStatus_t {
        rocsvHandle_t c = nullptr;
        const rocqStatus_t cs = rocsvCreate(&c);
        if (cs != ROCQ_STATUS_SUCCESS) return cs;
        g->child[(size_t)r] = c;
        c->fusion = h->fusion; c->tcBlocks = h->tcBlocks; c->mergeDiagonals = h->mergeDiagonals; c->planCache = h->planCache;
        c->blockMinCost = h->blockMinCost; c->budget = h->budget; c->tileBits = h->tileBits;
        c->seed = h->seed; c->draws = h->draws; c->seedExplicit = true;      // one Philox stream for all ranks
        return c->dist.init_in_group(c, r, g);
    });
    h->group = g;
    if (s != ROCQ_STATUS_SUCCESS) rq_group_destroy(h);
    return s;
}

// hooks for dist.cu
rocqStatus_t rq_engine_flush(rocsvInternalHandle* h) { return flush(h); }
rocqStatus_t rq_engine_run(rocsvInternalHandle* h, rq_cplx* state, unsigned n, const std::vector<HostOp>& ops, bool fused) {
    return run_ops(h, state, n, ops, fused);
}
rocqStatus_t rq_engine_fetch(rocsvInternalHandle* h, const void* dsrc, void* hdst, size_t bytes) { return fetch(h, dsrc, hdst, bytes); }
uint64_t rq_uniform53(uint64_t seed, uint64_t call, uint64_t shot) { return uniform53(seed, call, shot); }

extern "C" {

// ---- lifecycle (hipStateVec.cpp:190-251) ---------------------------------------------------------------
rocqStatus_t rocsvCreate(rocsvHandle_t* handle) {
    if (!handle) return ROCQ_STATUS_INVALID_VALUE;
    *handle = nullptr;
    H* h = new H();
    if (cudaStreamCreate(&h->stream) != cudaSuccess) {
        fprintf(stderr, "hipStateVec(B200): no usable CUDA device (%s); this engine has no CPU fallback\n",
                cudaGetErrorString(cudaGetLastError()));
        delete h;
        return ROCQ_STATUS_HIP_ERROR;
    }
    {   // stream-ordered scratch comes from the device's default pool, told to keep its memory across synchronisations
        // (engine.h, rq::pool_alloc).  A pool of the handle's own did the same for the steady state but paid ~90 ms for its
        // first large allocation (1M-shot sampling, profiles/r02_expect_variants.log).
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetDefaultMemPool(&h->pool, dev) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(h->pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            cudaGetLastError();
            h->pool = nullptr;
        }
    }
    const unsigned nb = rq_reduce_blocks();
    if (rq_sweep_configure() != 0 || rq_block_configure() != 0 || cudaMalloc(&h->d_partials, (nb + 8) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_upartials, (4 * nb + 4) * sizeof(uint64_t)) != cudaSuccess ||
        cudaHostAlloc(&h->h_scratch, 4096, cudaHostAllocDefault) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
        cudaEventCreate(&h->tm0) != cudaSuccess || cudaEventCreate(&h->tm1) != cudaSuccess) {
        fprintf(stderr, "hipStateVec(B200): handle setup failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        rocsvDestroy(h);
        return ROCQ_STATUS_HIP_ERROR;
    }
    // like the reference's QuantumSimulator (simulator.cpp:174: std::random_device): every handle draws its own measurement /
    // sampling stream unless a seed is set (rocsvxSetSeed, or ROCQ_SEED for whole runs)
    if (const char* e = getenv("ROCQ_SEED")) h->seed = strtoull(e, nullptr, 0);
    else { std::random_device rd; h->seed = ((uint64_t)rd() << 32) | (uint64_t)rd(); }
    if (const char* e = getenv("ROCQ_TRACE_LAUNCHES")) h->traceLaunches = atoi(e) != 0;
    if (const char* e = getenv("ROCQ_FUSION")) h->fusion = atoi(e) != 0;
    if (const char* e = getenv("ROCQ_TILE_BITS")) { const int t = atoi(e); if (t >= 6 && t <= RQ_MAX_TILE_BITS) h->tileBits = (unsigned)t; }
    if (const char* e = getenv("ROCQ_TC")) h->tcBlocks = (e[0] == 'a' || e[0] == '-') ? -1 : (atoi(e) != 0 && sizeof(rq_real) == 4) ? 1 : 0;
    if (const char* e = getenv("ROCQ_MERGE_DIAG")) h->mergeDiagonals = atoi(e) != 0;
    if (const char* e = getenv("ROCQ_PLAN_CACHE")) h->planCache = atoi(e) != 0;
    if (const char* e = getenv("ROCQ_TC_MIN_COST")) { const double b = atof(e); if (b > 0) h->blockMinCost = b; }
    if (const char* e = getenv("ROCQ_SWEEP_BUDGET")) { const double b = atof(e); if (b > 0) h->budget = b; }
    *handle = h;
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t rocsvDestroy(rocsvHandle_t h) {
    if (!h) return ROCQ_STATUS_SUCCESS;                   // hipStateVec.cpp:203-210
    rq_group_destroy(h);
    if (h->stream) { flush(h); cudaStreamSynchronize(h->stream); }
    h->dist.shutdown(h);
    rocsvFreeState(h);
    if (h->d_partials) cudaFree(h->d_partials);
    if (h->d_upartials) cudaFree(h->d_upartials);
    if (h->h_scratch) cudaFreeHost(h->h_scratch);
    if (h->pinned) cudaFreeHost(h->pinned);
    for (int b = 0; b < 2; ++b) { if (h->stage[b]) cudaFreeHost(h->stage[b]); if (h->stageEv[b]) cudaEventDestroy(h->stageEv[b]); }
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->tm0) cudaEventDestroy(h->tm0);
    if (h->tm1) cudaEventDestroy(h->tm1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t rocsvAllocateState(rocsvHandle_t h, unsigned numQubits, rocComplex** d_state, size_t batchSize) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    rq_group_destroy(h);                                  // a plain state replaces a distributed one
    if (numQubits > 40) return ROCQ_STATUS_ALLOCATION_FAILED;
    flush(h);
    cudaStreamSynchronize(h->stream);
    drop_cache(h);
    h->batchSize = batchSize > 0 ? batchSize : 1;
    h->numQubits = numQubits;
    if (h->d_state && h->ownsState) { cudaFree(h->d_state); h->d_state = nullptr; h->ownsState = false; }
    size_t total = h->batchSize * ((size_t)1 << numQubits);
    if (total * sizeof(rq_cplx) < 16) total = 16 / sizeof(rq_cplx);       // keep bulk copies 16-byte sized
    rq_cplx* p = nullptr;
    if (cudaMalloc(&p, total * sizeof(rq_cplx)) != cudaSuccess) { cudaGetLastError(); return ROCQ_STATUS_ALLOCATION_FAILED; }
    if (d_state) *d_state = reinterpret_cast<rocComplex*>(p);
    h->d_state = p;
    h->ownsState = true;
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t rocsvFreeState(rocsvHandle_t h) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    rq_group_destroy(h);
    if (h->stream) { flush(h); cudaStreamSynchronize(h->stream); }
    drop_cache(h);
    if (h->d_state && h->ownsState) cudaFree(h->d_state);
    h->d_state = nullptr;
    h->ownsState = false;
    h->numQubits = 0;
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t rocsvInitializeState(rocsvHandle_t h, rocComplex* d_state, unsigned numQubits) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    RQ_FWD(h, rocsvInitializeDistributedState(c));
    rq_cplx* state = resolve(h, d_state);
    if (!state) return ROCQ_STATUS_INVALID_VALUE;
    // a distributed handle has one state, its slices: the reset is the distributed one (it also resets the qubit map)
    if (h->dist.active() && state == h->d_state) return h->dist.initialize(h);
    RQ_OK(discard_or_flush_queue(h, state));
    const size_t total = h->batchSize * ((size_t)1 << numQubits);
    RQ_CUDA(rq_launch_init_state(state, total, 1, h->stream), "initialize state");
    h->stats.kernelLaunches++;
    h->numQubits = numQubits;
    return ROCQ_STATUS_SUCCESS;
}

// ---- named gates (hipStateVec.cpp:276-427) ---------------------------------------------------------------
rocqStatus_t rocsvApplyH(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned t) {
    RQ_FWD(h, rocsvApplyH(c, nullptr, n, t));
    const double s = 1.0 / std::sqrt(2.0);
    return single(h, d, n, t, rq::make_dense1(t, s, s, s, -s));
}
rocqStatus_t rocsvApplyX(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned t) {
    RQ_FWD(h, rocsvApplyX(c, nullptr, n, t));
    return single(h, d, n, t, rq::make_x(t)); }
rocqStatus_t rocsvApplyY(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned t) {
    RQ_FWD(h, rocsvApplyY(c, nullptr, n, t));
    return single(h, d, n, t, rq::make_dense1(t, 0.0, -I1, I1, 0.0));
}
rocqStatus_t rocsvApplyZ(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned t) {
    RQ_FWD(h, rocsvApplyZ(c, nullptr, n, t));
    return single(h, d, n, t, rq::make_phase(1ull << t, -1.0)); }
rocqStatus_t rocsvApplyS(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned t) {
    RQ_FWD(h, rocsvApplyS(c, nullptr, n, t));
    return single(h, d, n, t, rq::make_phase(1ull << t, I1)); }
rocqStatus_t rocsvApplySdg(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned t) {
    RQ_FWD(h, rocsvApplySdg(c, nullptr, n, t));
    return single(h, d, n, t, rq::make_phase(1ull << t, -I1)); }
rocqStatus_t rocsvApplyT(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned t) {
    RQ_FWD(h, rocsvApplyT(c, nullptr, n, t));
    const double ph = 3.14159265358979323846 / 4.0;
    return single(h, d, n, t, rq::make_phase(1ull << t, cd(std::cos(ph), std::sin(ph))));
}
rocqStatus_t rocsvApplyRx(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned t, double theta) {
    RQ_FWD(h, rocsvApplyRx(c, nullptr, n, t, theta));
    const double c = std::cos(theta / 2.0), s = std::sin(theta / 2.0);
    return single(h, d, n, t, rq::make_dense1(t, c, cd(0.0, -s), cd(0.0, -s), c));
}
rocqStatus_t rocsvApplyRy(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned t, double theta) {
    RQ_FWD(h, rocsvApplyRy(c, nullptr, n, t, theta));
    const double c = std::cos(theta / 2.0), s = std::sin(theta / 2.0);
    return single(h, d, n, t, rq::make_dense1(t, c, -s, s, c));
}
rocqStatus_t rocsvApplyRz(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned t, double theta) {
    RQ_FWD(h, rocsvApplyRz(c, nullptr, n, t, theta));
    const double c = std::cos(theta / 2.0), s = std::sin(theta / 2.0);
    return single(h, d, n, t, rq::make_diag1(t, cd(c, -s), cd(c, s)));
}

// ---- two-qubit and controlled gates (hipStateVec.cpp:431-687) -----------------------------------------------
rocqStatus_t rocsvApplyCNOT(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned ctl_, unsigned t) {
    RQ_FWD(h, rocsvApplyCNOT(c, nullptr, n, ctl_, t));
    const unsigned c = ctl_;
    return two(h, d, n, c, t, rq::make_x(t, 1ull << (c & 63))); }
rocqStatus_t rocsvApplyCZ(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned a, unsigned b) {
    RQ_FWD(h, rocsvApplyCZ(c, nullptr, n, a, b));
    return two(h, d, n, a, b, rq::make_phase((1ull << (a & 63)) | (1ull << (b & 63)), -1.0)); }
rocqStatus_t rocsvApplySWAP(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned a, unsigned b) {
    RQ_FWD(h, rocsvApplySWAP(c, nullptr, n, a, b));
    return two(h, d, n, a, b, rq::make_swap(a, b)); }
rocqStatus_t rocsvApplyCRX(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned ctl_, unsigned t, double theta) {
    RQ_FWD(h, rocsvApplyCRX(c, nullptr, n, ctl_, t, theta));
    const unsigned c = ctl_;
    const double co = std::cos(theta / 2.0), s = std::sin(theta / 2.0);
    return two(h, d, n, c, t, rq::make_dense1(t, co, cd(0.0, -s), cd(0.0, -s), co, 1ull << (c & 63)));
}
rocqStatus_t rocsvApplyCRY(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned ctl_, unsigned t, double theta) {
    RQ_FWD(h, rocsvApplyCRY(c, nullptr, n, ctl_, t, theta));
    const unsigned c = ctl_;
    const double co = std::cos(theta / 2.0), s = std::sin(theta / 2.0);
    return two(h, d, n, c, t, rq::make_dense1(t, co, -s, s, co, 1ull << (c & 63)));
}
rocqStatus_t rocsvApplyCRZ(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned ctl_, unsigned t, double theta) {
    RQ_FWD(h, rocsvApplyCRZ(c, nullptr, n, ctl_, t, theta));
    const unsigned c = ctl_;
    const double co = std::cos(theta / 2.0), s = std::sin(theta / 2.0);
    return two(h, d, n, c, t, rq::make_diag1(t, cd(co, -s), cd(co, s), 1ull << (c & 63)));
}

rocqStatus_t rocsvApplyMultiControlledX(rocsvHandle_t h, rocComplex* d, unsigned n, const unsigned* controls, unsigned nc, unsigned t) {
    RQ_FWD(h, rocsvApplyMultiControlledX(c, nullptr, n, controls, nc, t));
    if (!h || !controls || nc == 0) return ROCQ_STATUS_INVALID_VALUE;          // hipStateVec.cpp:605-607
    if (!resolve(h, d)) return ROCQ_STATUS_INVALID_VALUE;
    if (!valid_q(t, n)) return ROCQ_STATUS_INVALID_VALUE;
    if (nc > 63) return ROCQ_STATUS_NOT_IMPLEMENTED;                           // :613-615
    uint64_t mask = 0;
    for (unsigned i = 0; i < nc; ++i) {
        if (!valid_q(controls[i], n) || controls[i] == t) return ROCQ_STATUS_INVALID_VALUE;
        mask |= 1ull << controls[i];
    }
    return submit(h, d, n, rq::make_x(t, mask));
}

rocqStatus_t rocsvApplyCSWAP(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned ctl_, unsigned a, unsigned b) {
    RQ_FWD(h, rocsvApplyCSWAP(c, nullptr, n, ctl_, a, b));
    const unsigned c = ctl_;
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    if (!resolve(h, d)) return ROCQ_STATUS_INVALID_VALUE;
    if (!valid_q(c, n) || !valid_q(a, n) || !valid_q(b, n) || c == a || c == b || a == b) return ROCQ_STATUS_INVALID_VALUE;
    return submit(h, d, n, rq::make_swap(a, b, 1ull << c));
}

// ---- arbitrary matrices (hipStateVec.h:118-120, 151-157, 461-468) -----------------------------------------------
static rocqStatus_t apply_device_matrix(H* h, rocComplex* d, unsigned n, const unsigned* controls, unsigned nc,
                                        const unsigned* targets, unsigned k, const rocComplex* d_matrix) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    if (!resolve(h, d)) return ROCQ_STATUS_INVALID_VALUE;
    if (!targets || k == 0 || !d_matrix || (nc > 0 && !controls)) return ROCQ_STATUS_INVALID_VALUE;
    if (k > 10) return ROCQ_STATUS_NOT_IMPLEMENTED;
    if (nc > 63) return ROCQ_STATUS_NOT_IMPLEMENTED;
    uint64_t seen = 0, cmask = 0;
    for (unsigned i = 0; i < k; ++i) {
        if (!valid_q(targets[i], n) || ((seen >> targets[i]) & 1ull)) return ROCQ_STATUS_INVALID_VALUE;
        seen |= 1ull << targets[i];
    }
    for (unsigned i = 0; i < nc; ++i) {
        if (!valid_q(controls[i], n) || ((seen >> controls[i]) & 1ull)) return ROCQ_STATUS_INVALID_VALUE;
        seen |= 1ull << controls[i];
        cmask |= 1ull << controls[i];
    }
    HostOp o;
    o.targets.assign(targets, targets + k);
    o.cmask = cmask;
    if (h->fusion || h->dist.active()) {
        // deferred: the caller may free/overwrite the matrix before the flush -> take a host copy now
        const size_t D = (size_t)1 << k;
        std::vector<rq_cplx> hm(D * D);
        RQ_CUDA(cudaMemcpy(hm.data(), d_matrix, D * D * sizeof(rq_cplx), cudaMemcpyDeviceToHost), "matrix D2H");
        std::vector<cd> m(D * D);
        for (size_t e = 0; e < D * D; ++e) m[e] = cd(hm[e].x, hm[e].y);
        o = rq::make_matrix(o.targets, cmask, m);
    } else {
        o.kind = HostOp::DENSE;
        o.ext = d_matrix;
    }
    return submit(h, d, n, std::move(o));
}

rocqStatus_t rocsvApplyMatrix(rocsvHandle_t h, rocComplex* d, unsigned n, const unsigned* qubitIndices, unsigned numTargetQubits,
                              const rocComplex* matrixDevice, unsigned matrixDim) {
    RQ_FWD(h, rocsvApplyMatrix(c, nullptr, n, qubitIndices, numTargetQubits, matrixDevice, matrixDim));
    if (numTargetQubits < 32 && matrixDim != (1u << numTargetQubits)) return ROCQ_STATUS_INVALID_VALUE;
    return apply_device_matrix(h, d, n, nullptr, 0, qubitIndices, numTargetQubits, matrixDevice);
}
rocqStatus_t rocsvApplyControlledMatrix(rocsvHandle_t h, rocComplex* d, unsigned n, const unsigned* controls, unsigned nc,
                                        const unsigned* targets, unsigned nt, const rocComplex* d_matrix) {
    RQ_FWD(h, rocsvApplyControlledMatrix(c, nullptr, n, controls, nc, targets, nt, d_matrix));
    return apply_device_matrix(h, d, n, controls, nc, targets, nt, d_matrix);
}
rocqStatus_t rocsvApplyFusedSingleQubitMatrix(rocsvHandle_t h, unsigned targetQubit, const rocComplex* d_fusedMatrix) {
    RQ_FWD(h, rocsvApplyFusedSingleQubitMatrix(c, targetQubit, d_fusedMatrix));
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    const unsigned n = h->dist.active() ? h->dist.num_total() : h->numQubits;
    return apply_device_matrix(h, nullptr, n, nullptr, 0, &targetQubit, 1, d_fusedMatrix);
}

rocqStatus_t rocsvSwapIndexBits(rocsvHandle_t h, unsigned q1, unsigned q2) {
    RQ_FWD(h, rocsvSwapIndexBits(c, q1, q2));
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    const unsigned n = h->dist.active() ? h->dist.num_total() : h->numQubits;
    if (!h->d_state) return ROCQ_STATUS_INVALID_VALUE;
    if (!valid_q(q1, n) || !valid_q(q2, n)) return ROCQ_STATUS_INVALID_VALUE;
    if (q1 == q2) return ROCQ_STATUS_SUCCESS;
    if (h->dist.active()) return h->dist.swap_index_bits(h, q1, q2);
    return submit(h, nullptr, n, rq::make_swap(q1, q2));
}

// ---- readback (hipStateVec.cpp:691-730) ---------------------------------------------------------------
rocqStatus_t rocsvGetStateVectorFull(rocsvHandle_t h, rocComplex* d_state, rocComplex* h_state) {
    if (!h || !h_state) return ROCQ_STATUS_INVALID_VALUE;
    // one process, all slices: the caller's buffer holds the WHOLE state, slice r at r * 2^n_local (canonical order)
    RQ_FWD(h, rocsvGetStateVectorFull(c, nullptr, h_state + ((size_t)r << c->dist.num_local())));
    rq_cplx* state = resolve(h, d_state);
    if (!state) return ROCQ_STATUS_INVALID_VALUE;
    rocqStatus_t s = flush(h);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    if (h->dist.active()) { s = h->dist.canonicalize(h); if (s != ROCQ_STATUS_SUCCESS) return s; }
    const size_t total = h->batchSize * ((size_t)1 << h->numQubits);
    return staged_d2h(h, h_state, state, total * sizeof(rq_cplx));
}
rocqStatus_t rocsvGetStateVectorSlice(rocsvHandle_t h, rocComplex* d_state, rocComplex* h_state, unsigned batch_index) {
    if (!h || !h_state) return ROCQ_STATUS_INVALID_VALUE;
    if (h->group) return batch_index == 0 ? rocsvGetStateVectorFull(h, nullptr, h_state) : ROCQ_STATUS_INVALID_VALUE;   // distributed states are not batched
    rq_cplx* state = resolve(h, d_state);
    if (!state) return ROCQ_STATUS_INVALID_VALUE;
    if (batch_index >= h->batchSize) return ROCQ_STATUS_INVALID_VALUE;
    rocqStatus_t s = flush(h);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    if (h->dist.active()) { s = h->dist.canonicalize(h); if (s != ROCQ_STATUS_SUCCESS) return s; }
    const size_t N = (size_t)1 << h->numQubits;
    return staged_d2h(h, h_state, state + batch_index * N, N * sizeof(rq_cplx));
}

// ---- pinned buffer (hipStateVec.h:307-324) ---------------------------------------------------------------
rocqStatus_t rocsvEnsurePinnedBuffer(rocsvHandle_t h, size_t minSizeBytes) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    if (h->pinned && h->pinnedSize >= minSizeBytes) return ROCQ_STATUS_SUCCESS;
    if (h->pinned) { cudaFreeHost(h->pinned); h->pinned = nullptr; h->pinnedSize = 0; }
    if (minSizeBytes == 0) return ROCQ_STATUS_SUCCESS;
    if (cudaHostAlloc(&h->pinned, minSizeBytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); h->pinned = nullptr; return ROCQ_STATUS_ALLOCATION_FAILED; }
    h->pinnedSize = minSizeBytes;
    return ROCQ_STATUS_SUCCESS;
}
void* rocsvGetPinnedBufferPointer(rocsvHandle_t h) { return h ? h->pinned : nullptr; }
rocqStatus_t rocsvFreePinnedBuffer(rocsvHandle_t h) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    if (h->pinned) cudaFreeHost(h->pinned);
    h->pinned = nullptr;
    h->pinnedSize = 0;
    return ROCQ_STATUS_SUCCESS;
}

// ---- expectation values (hipStateVec.h:340-423) ---------------------------------------------------------------
static rocqStatus_t parse_pauli(unsigned n, const char* paulis, const unsigned* qubits, unsigned k, uint64_t& xm, uint64_t& zm, unsigned& ny) {
    if (k > 0 && (!paulis || !qubits)) return ROCQ_STATUS_INVALID_VALUE;
    xm = zm = 0;
    ny = 0;
    uint64_t seen = 0;
    for (unsigned j = 0; j < k; ++j) {
        if (!valid_q(qubits[j], n) || ((seen >> qubits[j]) & 1ull)) return ROCQ_STATUS_INVALID_VALUE;
        seen |= 1ull << qubits[j];
        const uint64_t bit = 1ull << qubits[j];
        switch (paulis[j]) {
            case 'I': case 'i': break;
            case 'X': case 'x': xm |= bit; break;
            case 'Y': case 'y': xm |= bit; zm |= bit; ++ny; break;
            case 'Z': case 'z': zm |= bit; break;
            default: return ROCQ_STATUS_INVALID_VALUE;
        }
    }
    return ROCQ_STATUS_SUCCESS;
}
static rocqStatus_t expect_string(H* h, rocComplex* d, unsigned n, const char* paulis, const unsigned* qubits, unsigned k, double* result) {
    if (!h || !result) return ROCQ_STATUS_INVALID_VALUE;
    if (h->group) {
        PerRank<double> out(result, h->group->P);
        return h->group->run([&](H* c, int r) { return expect_string(c, nullptr, n, paulis, qubits, k, out.at(r)); });
    }
    rq_cplx* state = resolve(h, d);
    if (!state) return ROCQ_STATUS_INVALID_VALUE;
    uint64_t xm = 0, zm = 0;
    unsigned ny = 0;
    rocqStatus_t s = parse_pauli(n, paulis, qubits, k, xm, zm, ny);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    s = flush(h);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    if (h->dist.active()) return h->dist.pauli_expect(h, xm, zm, ny, result);
    return pauli_expect(h, state, n, xm, zm, ny, result);
}
rocqStatus_t rocsvGetExpectationValueSinglePauliZ(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned t, double* r) {
    return expect_string(h, d, n, "Z", &t, 1, r);
}
rocqStatus_t rocsvGetExpectationValueSinglePauliX(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned t, double* r) { return expect_string(h, d, n, "X", &t, 1, r); }
rocqStatus_t rocsvGetExpectationValueSinglePauliY(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned t, double* r) { return expect_string(h, d, n, "Y", &t, 1, r); }
rocqStatus_t rocsvGetExpectationValuePauliProductZ(rocsvHandle_t h, rocComplex* d, unsigned n, const unsigned* qs, unsigned k, double* r) {
    if (k > 64) return ROCQ_STATUS_INVALID_VALUE;
    char buf[65];
    memset(buf, 'Z', sizeof buf);
    buf[k] = 0;
    return expect_string(h, d, n, buf, qs, k, r);
}
rocqStatus_t rocsvGetExpectationPauliString(rocsvHandle_t h, rocComplex* d, unsigned n, const char* paulis, const unsigned* qs, unsigned k, double* r) {
    if (paulis && strlen(paulis) != k) return ROCQ_STATUS_INVALID_VALUE;       // hipStateVec.h:413
    return expect_string(h, d, n, paulis, qs, k, r);
}
// Batched form (callers python/rocq/api.py:520-643 get_expval / grad, solvers/vqe_solver.py:120-136): terms
// are grouped by x-mask, every group is ONE read sweep that accumulates all its sign patterns (all-Z terms: one sweep in
// total), and the results come back in one copy.  allStates: evaluate every state of the batch (parameter-shift batches
// through batchSize), results[state * numTerms + term]; otherwise state 0 only, like the single-term calls.
static rocqStatus_t expect_batch(H* h, rocComplex* d, unsigned n, const char* paulis, const unsigned* qubits, const unsigned* offsets,
                                 unsigned numTerms, double* results, bool allStates) {
    if (!h || !results || !offsets) return ROCQ_STATUS_INVALID_VALUE;
    if (h->group) {
        PerRank<double> out(results, h->group->P, numTerms);
        return h->group->run([&](H* c, int r) { return expect_batch(c, nullptr, n, paulis, qubits, offsets, numTerms, out.at(r), allStates); });
    }
    rq_cplx* state = resolve(h, d);
    if (!state) return ROCQ_STATUS_INVALID_VALUE;
    if (numTerms == 0) return ROCQ_STATUS_SUCCESS;
    struct Term { uint64_t xm, zm; unsigned ny; };
    std::vector<Term> terms(numTerms);
    for (unsigned t = 0; t < numTerms; ++t) {
        const unsigned b = offsets[t], e = offsets[t + 1];
        if (e < b) return ROCQ_STATUS_INVALID_VALUE;
        const rocqStatus_t s = parse_pauli(n, paulis + b, qubits + b, e - b, terms[t].xm, terms[t].zm, terms[t].ny);
        if (s != ROCQ_STATUS_SUCCESS) return s;
    }
    RQ_OK(flush(h));
    if (h->dist.active()) {                                // slices: one term at a time (X/Y qubits may need an exchange first)
        for (unsigned t = 0; t < numTerms; ++t) RQ_OK(h->dist.pauli_expect(h, terms[t].xm, terms[t].zm, terms[t].ny, results + t));
        return ROCQ_STATUS_SUCCESS;
    }
    const rq::NvtxRange nvtx("rocq/pauli_expectation_batch");
    const unsigned nstates = allStates ? (unsigned)h->batchSize : 1u;
    std::vector<rq_pauli_group> groups;                    // first-appearance order of the x-masks, <= RQ_PAULI_GROUP_MAX terms each
    for (unsigned t = 0; t < numTerms; ++t) {
        rq_pauli_group* g = nullptr;
        // (16 terms per sweep, not the 32 a group can hold: 32 accumulators take 150 registers -- one resident block per SM --
        //  and run 2.3 ms at 28 qubits against 2 x 0.73 ms for two sweeps of 16, profiles/r02_expect_split.log)
        for (rq_pauli_group& c : groups) if (c.xmask == terms[t].xm && c.nterms < RQ_PAULI_GROUP_TERMS) { g = &c; break; }
        if (!g) { groups.emplace_back(); g = &groups.back(); memset(g, 0, sizeof *g); g->xmask = terms[t].xm; }
        g->zmask[g->nterms] = terms[t].zm;
        g->ny[g->nterms] = (uint8_t)(terms[t].ny & 0xffu);          // only ny mod 4 matters
        g->index[g->nterms] = t;
        g->nterms++;
    }
    rq::StreamBuf buf(h->stream, h->pool);
    const size_t npart = (size_t)rq_reduce_blocks() * RQ_PAULI_GROUP_MAX * nstates, nres = (size_t)numTerms * nstates;
    RQ_CUDA(buf.alloc((npart + nres) * sizeof(double)), "expectation scratch");
    double* d_part = buf.as<double>();
    double* d_res = d_part + npart;
    for (const rq_pauli_group& g : groups) {
        RQ_CUDA(rq_launch_pauli_group(state, n, nstates, &g, numTerms, d_part, d_res, h->stream), "expectation group launch");
        h->stats.kernelLaunches += 2;
    }
    h->stats.expectationSweeps += groups.size();
    RQ_CUDA(cudaMemcpyAsync(results, d_res, nres * sizeof(double), cudaMemcpyDeviceToHost, h->stream), "expectation D2H");
    RQ_CUDA(cudaStreamSynchronize(h->stream), "stream sync");
    return ROCQ_STATUS_SUCCESS;
}
rocqStatus_t rocsvxGetExpectationPauliBatch(rocsvHandle_t h, rocComplex* d, unsigned n, const char* paulis, const unsigned* qubits,
                                            const unsigned* offsets, unsigned numTerms, double* results) {
    return expect_batch(h, d, n, paulis, qubits, offsets, numTerms, results, false);
}
rocqStatus_t rocsvxGetExpectationPauliBatchAllStates(rocsvHandle_t h, rocComplex* d, unsigned n, const char* paulis, const unsigned* qubits,
                                                     const unsigned* offsets, unsigned numTerms, double* results) {
    return expect_batch(h, d, n, paulis, qubits, offsets, numTerms, results, true);
}
rocqStatus_t rocsvxSetTensorCoreBlocks(rocsvHandle_t h, int enabled) {
    if (h && h->group && (enabled <= 0 || sizeof(rq_real) == 4)) h->tcBlocks = enabled > 0 ? 1 : enabled < 0 ? -1 : 0;
    RQ_FWD(h, rocsvxSetTensorCoreBlocks(c, enabled));
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    if (enabled > 0 && sizeof(rq_real) != 4) return ROCQ_STATUS_NOT_IMPLEMENTED;
    h->tcBlocks = enabled > 0 ? 1 : enabled < 0 ? -1 : 0;
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t rocsvxSetPlanCache(rocsvHandle_t h, int enabled) {
    if (h && h->group) h->planCache = enabled != 0;
    RQ_FWD(h, rocsvxSetPlanCache(c, enabled));
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    h->planCache = enabled != 0;
    if (!enabled) drop_cache(h);
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t rocsvxSetMergeDiagonals(rocsvHandle_t h, int enabled) {
    if (h && h->group) h->mergeDiagonals = enabled != 0;
    RQ_FWD(h, rocsvxSetMergeDiagonals(c, enabled));
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    const rocqStatus_t s = flush(h);                     // queued gates keep the setting they were submitted under
    h->mergeDiagonals = enabled != 0;
    return s;
}

rocqStatus_t rocsvxApplyBlock6(rocsvHandle_t h, rocComplex* d, unsigned n, const unsigned* qubits, const double* matrix) {
    RQ_FWD(h, rocsvxApplyBlock6(c, nullptr, n, qubits, matrix));
    if (!h || !qubits || !matrix) return ROCQ_STATUS_INVALID_VALUE;
    rq_cplx* state = resolve(h, d);
    if (!state) return ROCQ_STATUS_INVALID_VALUE;
    if (sizeof(rq_real) != 4 || h->dist.active()) return ROCQ_STATUS_NOT_IMPLEMENTED;
    if (n < 13) return ROCQ_STATUS_INVALID_VALUE;
    uint64_t seen = 0;
    for (unsigned b = 0; b < 6; ++b) {
        if (qubits[b] >= n || ((seen >> qubits[b]) & 1ull)) return ROCQ_STATUS_INVALID_VALUE;
        seen |= 1ull << qubits[b];
    }
    const rocqStatus_t s = flush(h);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    // reorder the matrix so that index bit b <-> b-th smallest qubit
    std::vector<unsigned> blk(qubits, qubits + 6);
    std::sort(blk.begin(), blk.end());
    unsigned bitmap[6];                                    // API bit b -> sorted bit
    for (unsigned b = 0; b < 6; ++b) bitmap[b] = (unsigned)(std::find(blk.begin(), blk.end(), qubits[b]) - blk.begin());
    auto remap = [&](unsigned i) { unsigned o = 0; for (unsigned b = 0; b < 6; ++b) if ((i >> b) & 1u) o |= 1u << bitmap[b]; return o; };
    std::vector<cd> U(64 * 64);
    for (unsigned c = 0; c < 64; ++c)
        for (unsigned r = 0; r < 64; ++r) U[remap(r) + 64u * remap(c)] = cd(matrix[2 * (r + 64u * c)], matrix[2 * (r + 64u * c) + 1]);
    h->stats.gatesSubmitted++;
    uint64_t bm = 0;
    for (unsigned q : blk) bm |= 1ull << q;
    if (!rq::block_supported(bm, n, h->batchSize)) {
        // the tile of a very scattered block does not fit one five-dimensional tensor map: generic dense path
        std::vector<HostOp> one{rq::make_matrix(blk, 0ull, U)};
        return run_ops(h, state, n, one, false);
    }
    return run_block(h, state, n, blk, U);
}

rocqStatus_t rocsvxGetNorm(rocsvHandle_t h, rocComplex* d, unsigned n, double* r) { return expect_string(h, d, n, "", nullptr, 0, r); }

// ---- measurement (hipStateVec.h:172-177; algorithm measurement_kernels.hip:37-77, MULTI_GPU_GUIDE.md:61-78) ----
rocqStatus_t rocsvMeasure(rocsvHandle_t h, rocComplex* d, unsigned n, unsigned q, int* outcome, double* probability) {
    if (!h || !outcome) return ROCQ_STATUS_INVALID_VALUE;
    if (h->group) {
        PerRank<int> o(outcome, h->group->P);
        PerRank<double> p(probability, h->group->P);
        return h->group->run([&](H* c, int r) { return rocsvMeasure(c, nullptr, n, q, o.at(r), p.at(r)); });
    }
    rq_cplx* state = resolve(h, d);
    if (!state) return ROCQ_STATUS_INVALID_VALUE;
    if (!valid_q(q, n)) return ROCQ_STATUS_INVALID_VALUE;
    rocqStatus_t s = flush(h);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    if (h->dist.active()) return h->dist.measure(h, q, outcome, probability);
    const unsigned nb = rq_reduce_blocks();
    RQ_CUDA(rq_launch_fixed_masses(state, n, q, h->d_upartials, nb, h->d_upartials + 4 * nb, h->stream), "mass reduction");
    h->stats.kernelLaunches += 2;
    uint64_t m[4];
    s = fetch(h, h->d_upartials + 4 * nb, m, sizeof m);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    const u128 S0 = ((u128)m[0] << 64) | m[1], S1 = ((u128)m[2] << 64) | m[3];
    const uint64_t U = uniform53(h->seed, h->draws++, 0);
    const int out = mul_u53(S0 + S1, U) < S0 ? 0 : 1;
    const double tot = u128_to_double(S0 + S1), mass = u128_to_double(out ? S1 : S0);
    if (!(mass > 0.0)) return ROCQ_STATUS_FAILURE;
    *outcome = out;
    if (probability) *probability = mass / tot;
    RQ_CUDA(rq_launch_collapse(state, n, q, out, 1.0 / std::sqrt(mass * 0x1p-88), h->stream), "collapse");
    h->stats.kernelLaunches++;
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t rocsvApplyMatrixAndMeasure(rocsvHandle_t h, rocComplex* d, unsigned n, const unsigned* targets, unsigned nt,
                                        const rocComplex* d_matrix, unsigned qubitToMeasure, int* outcome) {
    if (h && h->group && outcome) {
        PerRank<int> o(outcome, h->group->P);
        return h->group->run([&](H* c, int r) { return rocsvApplyMatrixAndMeasure(c, nullptr, n, targets, nt, d_matrix, qubitToMeasure, o.at(r)); });
    }
    const rocqStatus_t s = apply_device_matrix(h, d, n, nullptr, 0, targets, nt, d_matrix);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    double p = 0.0;
    return rocsvMeasure(h, d, n, qubitToMeasure, outcome, &p);
}

// ---- sampling (hipStateVec.h:439-445) ---------------------------------------------------------------
rocqStatus_t rocsvSample(rocsvHandle_t h, rocComplex* d, unsigned n, const unsigned* measured, unsigned nm, unsigned numShots,
                         uint64_t* h_results) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    if (h->group) {
        PerRank<uint64_t> out(h_results, h->group->P, numShots);
        return h->group->run([&](H* c, int r) { return rocsvSample(c, nullptr, n, measured, nm, numShots, out.at(r)); });
    }
    rq_cplx* state = resolve(h, d);
    if (!state) return ROCQ_STATUS_INVALID_VALUE;
    if (nm > 64 || (nm > 0 && !measured)) return ROCQ_STATUS_INVALID_VALUE;
    for (unsigned j = 0; j < nm; ++j) if (!valid_q(measured[j], n)) return ROCQ_STATUS_INVALID_VALUE;
    if (numShots == 0) return ROCQ_STATUS_SUCCESS;
    if (!h_results) return ROCQ_STATUS_INVALID_VALUE;
    rocqStatus_t s = flush(h);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    if (h->dist.active()) return h->dist.sample(h, measured, nm, numShots, h_results);

    unsigned cb = n < 10 ? n : (n > 30 ? n - 20 : 10);                    // chunk of 2^cb amplitudes, <= 2^20 chunks
    // Masses are exact integers, so the sampled indices do not depend on the chunking: ROCQ_SAMPLE_CHUNK_BITS trades the
    // scan (2^(n-cb) chunk masses) against the in-chunk walk per shot (2^cb amplitudes) -- tuning only.
    static const int cb_env = getenv("ROCQ_SAMPLE_CHUNK_BITS") ? atoi(getenv("ROCQ_SAMPLE_CHUNK_BITS")) : -1;
    if (cb_env >= 0) cb = std::min<unsigned>(n, std::max<unsigned>((unsigned)cb_env, n > 30 ? n - 20 : 0u));
    const uint64_t nchunks = 1ull << (n - cb);
    // everything stays on the device: chunk masses -> exact scan -> one warp per shot -> result words (bit j = measured
    // qubit j); the host sees one copy of the results and the total mass
    const rq::NvtxRange nvtx("rocq/sample");
    rq::StreamBuf scratch(h->stream, h->pool);
    RQ_CUDA(scratch.alloc((2 * nchunks + 2 * RQ_SCAN_MAXSEG + 4 + (size_t)numShots) * sizeof(uint64_t)), "sampling scratch");
    uint64_t* d_hi = scratch.as<uint64_t>();
    uint64_t *d_lo = d_hi + nchunks, *d_btot = d_lo + nchunks, *d_tot = d_btot + 2 * RQ_SCAN_MAXSEG, *d_idx = d_tot + 4;
    RQ_CUDA(rq_launch_chunk_masses(state, n, cb, d_hi, d_lo, h->stream), "chunk masses");
    RQ_CUDA(cudaMemsetAsync(d_tot, 0, 4 * sizeof(uint64_t), h->stream), "totals clear");      // win = 0: one rank holds everything
    RQ_CUDA(rq_launch_scan_masses(d_hi, d_lo, nchunks, d_btot, d_tot, h->stream), "mass scan");
    rq_shot_map map{};
    map.nm = nm;
    for (unsigned j = 0; j < nm; ++j) map.pos[j] = (uint8_t)measured[j];
    RQ_CUDA(rq_launch_sample(state, n, cb, d_hi, d_lo, nchunks, d_tot, h->seed, h->draws++, numShots, 0, &map, d_idx, h->stream), "sample");
    h->stats.kernelLaunches += nchunks > 256 ? 5 : 4;
    RQ_CUDA(cudaMemcpyAsync(h_results, d_idx, (size_t)numShots * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream), "shots D2H");
    uint64_t tot[2];
    s = fetch(h, d_tot, tot, sizeof tot);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    if (tot[0] == 0 && tot[1] == 0) return ROCQ_STATUS_FAILURE;           // an all-zero state has no distribution to draw from
    return ROCQ_STATUS_SUCCESS;
}

// ---- extensions ---------------------------------------------------------------
unsigned rocsvxGetPrecisionBytes(void) { return (unsigned)sizeof(rq_real); }
rocqStatus_t rocsvxSetSeed(rocsvHandle_t h, uint64_t seed) {
    if (h && h->group) { h->seed = seed; h->draws = 0; h->seedExplicit = true; }
    RQ_FWD(h, rocsvxSetSeed(c, seed));
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    h->seed = seed;
    h->draws = 0;
    h->seedExplicit = true;
    return ROCQ_STATUS_SUCCESS;
}
rocqStatus_t rocsvxSetStateVector(rocsvHandle_t h, rocComplex* d_state, const rocComplex* h_state) {
    if (!h || !h_state) return ROCQ_STATUS_INVALID_VALUE;
    RQ_FWD(h, rocsvxSetStateVector(c, nullptr, h_state + ((size_t)r << c->dist.num_local())));
    rq_cplx* state = resolve(h, d_state);
    if (!state) return ROCQ_STATUS_INVALID_VALUE;
    RQ_OK(discard_or_flush_queue(h, state));
    // the host slice of a distributed state is in canonical layout: the loaded data starts from the identity qubit map
    if (h->dist.active() && state == h->d_state) h->dist.reset_layout();
    const size_t total = h->batchSize * ((size_t)1 << h->numQubits);
    return staged_h2d(h, state, h_state, total * sizeof(rq_cplx));
}
rocqStatus_t rocsvxSynchronize(rocsvHandle_t h) {
    RQ_FWD(h, rocsvxSynchronize(c));
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    const rocqStatus_t s = flush(h);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    RQ_CUDA(cudaStreamSynchronize(h->stream), "stream sync");
    return ROCQ_STATUS_SUCCESS;
}
rocqStatus_t rocsvxSetFusion(rocsvHandle_t h, int enabled) {
    if (h && h->group) h->fusion = enabled != 0;
    RQ_FWD(h, rocsvxSetFusion(c, enabled));
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    if (!enabled) { const rocqStatus_t s = flush(h); if (s != ROCQ_STATUS_SUCCESS) return s; }
    h->fusion = enabled != 0;
    return ROCQ_STATUS_SUCCESS;
}
rocqStatus_t rocsvxFlush(rocsvHandle_t h) {
    RQ_FWD(h, rocsvxFlush(c));
    return h ? flush(h) : ROCQ_STATUS_INVALID_VALUE;
}

using rq::convert_ops;      // gate_convert.h

rocqStatus_t rocsvxApplyCircuit(rocsvHandle_t h, rocComplex* d, unsigned n, const rocsvxGateOp* ops, size_t numOps) {
    RQ_FWD(h, rocsvxApplyCircuit(c, nullptr, n, ops, numOps));
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    rq_cplx* state = resolve(h, d);
    if (!state) return ROCQ_STATUS_INVALID_VALUE;
    if (numOps == 0) return ROCQ_STATUS_SUCCESS;
    if (!ops) return ROCQ_STATUS_INVALID_VALUE;
    static const bool prof = getenv("ROCQ_HOST_PROFILE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<HostOp> hops;
    rocqStatus_t s = convert_ops(n, ops, numOps, hops);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    s = flush(h);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    h->stats.gatesSubmitted += numOps;
    if (h->dist.active()) return h->dist.run_circuit(h, hops);
    const auto t1 = std::chrono::steady_clock::now();
    // Plan cache: the identical circuit on the identical state buffer and settings replays the recorded launches (all the
    // device work is done again; what is skipped is fusion, planning and the host-side products of the block matrices).
    const bool cacheable = h->planCache && !getenv("ROCQ_BLOCK_DEBUG");
    Hash2 key;
    if (cacheable) {
        key = circuit_key(h, state, n, ops, numOps);
        if (h->cache.valid && h->cache.key[0] == key.a && h->cache.key[1] == key.b && h->cache.bytes == key.bytes) {
            cudaEventRecord(h->ev0, h->stream);
            s = replay_cache(h, state);
            cudaEventRecord(h->ev1, h->stream);
            h->stats.lastSweepMs = -1.0;
            h->stats.planCacheHits++;
            return s;
        }
    }
    std::vector<rocsvCachedStep> rec;
    if (cacheable) { h->recording = &rec; h->recordingValid = true; }
    cudaEventRecord(h->ev0, h->stream);
    s = run_ops(h, state, n, hops, true);
    cudaEventRecord(h->ev1, h->stream);
    h->recording = nullptr;
    if (cacheable) {
        if (s == ROCQ_STATUS_SUCCESS && h->recordingValid) {
            drop_cache(h);
            h->cache.steps = std::move(rec);
            h->cache.key[0] = key.a; h->cache.key[1] = key.b; h->cache.bytes = key.bytes;
            h->cache.valid = true;
        } else {
            free_steps(h, rec);
        }
    }
    if (prof) fprintf(stderr, "[host profile] ApplyCircuit: convert %.2f ms, run_ops (fuse + plan + launches) %.2f ms\n",
                      std::chrono::duration<double, std::milli>(t1 - t0).count(),
                      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    h->stats.lastSweepMs = -1.0;
    return s;
}

rocqStatus_t rocsvxTimerStart(rocsvHandle_t h) {
    RQ_FWD(h, rocsvxTimerStart(c));
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    RQ_CUDA(cudaEventRecord(h->tm0, h->stream), "timer start");
    return ROCQ_STATUS_SUCCESS;
}
rocqStatus_t rocsvxTimerStop(rocsvHandle_t h, double* ms) {
    if (!h || !ms) return ROCQ_STATUS_INVALID_VALUE;
    if (h->group) {                                        // device time of the slowest rank
        std::vector<double> each((size_t)h->group->P, 0.0);
        const rocqStatus_t s = h->group->run([&](H* c, int r) { return rocsvxTimerStop(c, &each[(size_t)r]); });
        *ms = *std::max_element(each.begin(), each.end());
        return s;
    }
    RQ_CUDA(cudaEventRecord(h->tm1, h->stream), "timer stop");
    RQ_CUDA(cudaEventSynchronize(h->tm1), "timer sync");
    float f = 0.f;
    RQ_CUDA(cudaEventElapsedTime(&f, h->tm0, h->tm1), "timer elapsed");
    *ms = (double)f;
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t rocsvxGetStats(rocsvHandle_t h, rocsvxStats* stats, int reset) {
    if (!h) return ROCQ_STATUS_INVALID_VALUE;
    if (h->group) {                                        // rank 0's counters (every rank launches the same plan)
        PerRank<rocsvxStats> out(stats, h->group->P);
        return h->group->run([&](H* c, int r) { return rocsvxGetStats(c, out.at(r), reset); });
    }
    if (h->stats.lastSweepMs < 0.0) {
        float ms = 0.f;
        cudaEventSynchronize(h->ev1);
        h->stats.lastSweepMs = cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess ? (double)ms : 0.0;
    }
    for (auto& ev : h->dist.timed) {
        float ms = 0.f;
        cudaEventSynchronize(ev.second);
        if (cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) h->stats.exchangeMs += (double)ms;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    h->dist.timed.clear();
    if (!h->traced.empty()) {                              // ROCQ_TRACE_LAUNCHES: the launch list since the last call
        cudaStreamSynchronize(h->stream);
        double total = 0.0;
        for (auto& t : h->traced) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) {
                fprintf(stderr, "[launch] rank %d %s ops %u ms %.3f\n", h->dist.rank, t.what, t.ops, (double)ms);
                total += ms;
            }
            cudaEventDestroy(t.e0);
            cudaEventDestroy(t.e1);
        }
        fprintf(stderr, "[launch] rank %d total of %zu launches: %.3f ms\n", h->dist.rank, h->traced.size(), total);
        h->traced.clear();
    }
    if (stats) *stats = h->stats;
    if (reset) h->stats = rocsvxStats{};
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t rocsvxPlanCircuitBlocks(unsigned n, const rocsvxGateOp* ops, size_t numOps, double minCost, unsigned* numBlocks,
                                     unsigned* numSweeps, char* buf, size_t bufSize) {
    if (!ops && numOps) return ROCQ_STATUS_INVALID_VALUE;
    std::vector<HostOp> hops;
    const rocqStatus_t s = convert_ops(n, ops, numOps, hops);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    for (const HostOp& o : hops) if (o.targets.size() > 4) return ROCQ_STATUS_NOT_IMPLEMENTED;
    std::vector<HostOp> fused = hops.size() > 1 ? rq::fuse_algebraic(rq::merge_diagonals(rq::push_x_forward(hops)), n) : hops;
    rq::PlanLimits L;
    L.max_ops = sizeof(rq_program_large::ops) / sizeof(rq_tile_op);
    L.pool_cplx = sizeof(rq_program_large::pool) / sizeof(rq_cplx);
    if (const char* e = getenv("ROCQ_SWEEP_BUDGET")) { const double b = atof(e); if (b > 0) L.budget = b; }
    rq::BlockLimits BL;
    if (minCost > 0.0) BL.min_cost = minCost;
    const std::vector<rq::MixedStep> steps = rq::plan_mixed(fused, n, L, BL);
    unsigned nb = 0, ns = 0;
    std::string text;
    char line[64];
    for (const rq::MixedStep& st : steps) {
        if (st.block) {
            ++nb;
            text += "B";
            for (unsigned p : st.blk) { snprintf(line, sizeof line, " %u", p); text += line; }
            text += "\n";
            rq::SweepPlan sp;                       // ops in program order, dumped like a sweep's
            sp.ops = st.ops;
            text += rq::dump_ops(sp.ops, fused);
        } else {
            ++ns;
            text += rq::dump_plan(std::vector<rq::SweepPlan>{st.sweep}, fused);
        }
    }
    if (numBlocks) *numBlocks = nb;
    if (numSweeps) *numSweeps = ns;
    if (buf && bufSize) {
        const size_t c = std::min(bufSize - 1, text.size());
        memcpy(buf, text.data(), c);
        buf[c] = 0;
    }
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t rocsvxPlanCircuit(unsigned n, unsigned tileBits, const rocsvxGateOp* ops, size_t numOps, unsigned* numSweeps, char* buf,
                               size_t bufSize) {
    if (!ops && numOps) return ROCQ_STATUS_INVALID_VALUE;
    std::vector<HostOp> hops;
    const rocqStatus_t s = convert_ops(n, ops, numOps, hops);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    for (const HostOp& o : hops) if (o.targets.size() > 4) return ROCQ_STATUS_NOT_IMPLEMENTED;
    std::vector<HostOp> fused = hops.size() > 1 ? rq::fuse_algebraic(rq::merge_diagonals(rq::push_x_forward(hops)), n) : hops;
    rq::PlanLimits L;
    if (tileBits >= 1 && tileBits <= RQ_MAX_TILE_BITS) L.tile_bits = tileBits;
    L.max_ops = sizeof(rq_program_large::ops) / sizeof(rq_tile_op);
    L.pool_cplx = sizeof(rq_program_large::pool) / sizeof(rq_cplx);
    if (const char* e = getenv("ROCQ_SWEEP_BUDGET")) { const double b = atof(e); if (b > 0) L.budget = b; }
    const std::vector<rq::SweepPlan> plans = rq::plan_sweeps(fused, n, L);
    // every plan must be emittable
    static thread_local rq_program_large P;
    for (const rq::SweepPlan& sp : plans)
        if (!rq::build_program(P, sp, fused, n, 1, 0)) return ROCQ_STATUS_FAILURE;
    if (numSweeps) *numSweeps = (unsigned)plans.size();
    if (buf && bufSize) {
        const std::string txt = rq::dump_plan(plans, fused);
        const size_t m = txt.size() < bufSize - 1 ? txt.size() : bufSize - 1;
        memcpy(buf, txt.data(), m);
        buf[m] = 0;
    }
    return ROCQ_STATUS_SUCCESS;
}

rocqStatus_t rocsvxDistPlanCircuit(unsigned n, int numRanks, const rocsvxGateOp* ops, size_t numOps, int mode, int canonicalize,
                                   unsigned* numExchanges, char* buf, size_t bufSize) {
    if ((!ops && numOps) || numRanks < 1 || (numRanks & (numRanks - 1))) return ROCQ_STATUS_INVALID_VALUE;
    unsigned M = 0;
    while ((1 << M) < numRanks) ++M;
    if (n < M) return ROCQ_STATUS_INVALID_VALUE;
    std::vector<HostOp> hops;
    const rocqStatus_t s = convert_ops(n, ops, numOps, hops);
    if (s != ROCQ_STATUS_SUCCESS) return s;
    rq::DistPlanner P;
    P.reset(n, n - M);
    if ((mode & 1) == 0) { if (!((mode & 8) ? P.add_circuit_inorder(hops) : P.add_circuit(hops))) return ROCQ_STATUS_FAILURE; }
    else for (const HostOp& o : hops) if (!P.add_op(o)) return ROCQ_STATUS_NOT_IMPLEMENTED;
    P.flush_pending();
    if (canonicalize) P.canonicalize();
    unsigned nx = 0;
    for (const rq::DistStep& st : P.steps) nx += st.kind == rq::DistStep::EXCHANGE;
    if (numExchanges) *numExchanges = nx;
    if (buf && bufSize) {
        std::string txt;
        if (mode & 2) {
            // the full pipeline: every RUN step fused and cut into sweeps exactly as the engine does on a slice
            rq::PlanLimits L;
            L.max_ops = sizeof(rq_program_large::ops) / sizeof(rq_tile_op);
            L.pool_cplx = sizeof(rq_program_large::pool) / sizeof(rq_cplx);
            L.never_resident = P.global_mask();
            if (const char* e = getenv("ROCQ_TILE_BITS")) { const int t = atoi(e); if (t >= 1 && t <= RQ_MAX_TILE_BITS) L.tile_bits = (unsigned)t; }
            static thread_local rq_program_large prog;
            for (const rq::DistStep& st : P.steps) {
                if (st.kind == rq::DistStep::EXCHANGE) {
                    txt += "X";
                    for (unsigned g : st.gpos) txt += " " + std::to_string(g);
                    txt += "\n";
                    continue;
                }
                txt += "R\n";
                for (const HostOp& o : st.ops) if (o.targets.size() > 4) return ROCQ_STATUS_NOT_IMPLEMENTED;
                // mode bit 4 (16): what ONE rank really executes -- rank = mode >> 8: controls / diagonal factors on rank bits are
                // resolved first (specialize_for_rank), exactly as the engine does on a slice; mode bit 5 (32): the engine's own
                // block threshold (ROCQ_TC_MIN_COST or the default) instead of "a block whenever one is possible"
                const std::vector<HostOp> spec = (mode & 16) ? rq::specialize_for_rank(st.ops, n - M, (uint64_t)((unsigned)mode >> 8) << (n - M)) : st.ops;
                std::vector<HostOp> fused = spec.size() > 1 ? rq::fuse_algebraic(rq::merge_diagonals(rq::push_x_forward(spec)), n - M, P.global_mask()) : spec;
                if (mode & 4) {                              // with tensor-core blocks on the local qubits (complex64 engine)
                    rq::BlockLimits BL;
                    BL.min_cost = 0.0;
                    if (mode & 32) { BL.min_cost = rocsvInternalHandle().blockMinCost; if (const char* e = getenv("ROCQ_TC_MIN_COST")) { const double b = atof(e); if (b > 0) BL.min_cost = b; } }
                    for (const rq::MixedStep& ms : rq::plan_mixed(fused, n - M, L, BL)) {
                        if (ms.block) {
                            txt += "B";
                            for (unsigned p : ms.blk) txt += " " + std::to_string(p);
                            txt += "\n" + rq::dump_ops(ms.ops, fused);
                        } else {
                            if (!rq::build_program(prog, ms.sweep, fused, n - M, 1, 0)) return ROCQ_STATUS_FAILURE;
                            txt += rq::dump_plan(std::vector<rq::SweepPlan>{ms.sweep}, fused);
                        }
                    }
                    continue;
                }
                const std::vector<rq::SweepPlan> plans = rq::plan_sweeps(fused, n - M, L);
                for (const rq::SweepPlan& sp : plans)
                    if (!rq::build_program(prog, sp, fused, n - M, 1, 0)) return ROCQ_STATUS_FAILURE;
                txt += rq::dump_plan(plans, fused);
            }
            txt += "M";
            for (unsigned q = 0; q < n; ++q) txt += " " + std::to_string(P.map[q]);
            txt += "\n";
        } else {
            txt = P.dump();
        }
        const size_t m = txt.size() < bufSize - 1 ? txt.size() : bufSize - 1;
        memcpy(buf, txt.data(), m);
        buf[m] = 0;
    }
    return ROCQ_STATUS_SUCCESS;
}

}  // extern "C"
