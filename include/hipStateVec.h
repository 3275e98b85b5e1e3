/* include/hipStateVec.h -- the drop-in C ABI of the B200-native state-vector engine.
 *
 * Source-compatible twin of the reference header
 *   /root/reference/rocquantum/include/rocquantum/hipStateVec.h   (cited below as REF:<line>)
 * with the <hip/hip_runtime.h> dependency removed: rocComplex is a plain interleaved (re,im)
 * pair, layout-identical to hipFloatComplex / hipDoubleComplex (and float2 / double2).
 * Every rocsv* prototype below has exactly the reference's signature; the 17 entry points the
 * reference declares but never defines are implemented here from their
 * documented contract.  Entry points prefixed rocsvx are extensions of this engine.
 *
 * Precision is fixed at library build time, like the reference (REF:6-15):
 *   libhipStateVec.so       rocComplex = complex64   (default)
 *   libhipStateVec_f64.so   rocComplex = complex128  (built with -DROCQ_PRECISION_DOUBLE)
 *
 * Pointers: d_state / matrixDevice / d_matrix / d_fusedMatrix are DEVICE pointers; qubit index
 * arrays, result pointers and h_* buffers are HOST pointers (REF:147, 435).  d_state == NULL means
 * "the state owned by the handle" (reference hipStateVec.cpp:80-85).
 * No function throws; all return rocqStatus_t (REF:22-31).
 */
#ifndef HIPSTATEVEC_H
#define HIPSTATEVEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef ROCQ_PRECISION_DOUBLE                                   /* REF:7-10 */
typedef struct { double x, y; } rocComplex;
typedef double real_t;
static const real_t REAL_EPSILON = 1e-12;
#else                                                          /* REF:11-15 */
typedef struct { float x, y; } rocComplex;
typedef float real_t;
static const real_t REAL_EPSILON = 1e-6f;
#endif

struct rocsvInternalHandle;                                    /* REF:18-19 */
typedef struct rocsvInternalHandle* rocsvHandle_t;

typedef enum {                                                 /* REF:22-31 */
    ROCQ_STATUS_SUCCESS = 0,
    ROCQ_STATUS_FAILURE = 1,
    ROCQ_STATUS_INVALID_VALUE = 2,
    ROCQ_STATUS_ALLOCATION_FAILED = 3,
    ROCQ_STATUS_HIP_ERROR = 4,          /* any CUDA runtime / launch failure */
    ROCQ_STATUS_NOT_IMPLEMENTED = 5,
    ROCQ_STATUS_RCCL_ERROR = 6          /* any NCCL failure */
} rocqStatus_t;

#ifdef __cplusplus
extern "C" {
#endif

/* ---- lifecycle and state (REF:43, 51, 61, 69, 79) ---- */
rocqStatus_t rocsvCreate(rocsvHandle_t* handle);
rocqStatus_t rocsvDestroy(rocsvHandle_t handle);
/* numQubits > 40 (8.8 TB of complex64) is refused with ROCQ_STATUS_ALLOCATION_FAILED before any allocation is tried. */
rocqStatus_t rocsvAllocateState(rocsvHandle_t handle, unsigned numQubits, rocComplex** d_state, size_t batchSize);
rocqStatus_t rocsvFreeState(rocsvHandle_t handle);
rocqStatus_t rocsvInitializeState(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits);

/* ---- distributed state (REF:92, 104).  One process per GPU: call rocsvxDistInit first; each rank
 *      then owns the slice global = (rank << numLocalQubits) | local (swap_kernels.hip:10-22). ---- */
rocqStatus_t rocsvAllocateDistributedState(rocsvHandle_t handle, unsigned totalNumQubits);
rocqStatus_t rocsvInitializeDistributedState(rocsvHandle_t handle);

/* ---- fused 1-qubit matrix on the handle's state; 2x2 DEVICE, column-major (REF:118-120) ---- */
rocqStatus_t rocsvApplyFusedSingleQubitMatrix(rocsvHandle_t handle, unsigned targetQubit, const rocComplex* d_fusedMatrix);

/* ---- relabel two qubit positions by physically permuting amplitudes (REF:135-137) ---- */
rocqStatus_t rocsvSwapIndexBits(rocsvHandle_t handle, unsigned qubit_idx1, unsigned qubit_idx2);

/* ---- arbitrary k-qubit matrix: DEVICE, column-major M[i + j*dim], index bit b <-> qubitIndices[b]
 *      (REF:151-157; layout spec multi_qubit_kernels.hip:22-30, 91-99) ---- */
rocqStatus_t rocsvApplyMatrix(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits,
                              const unsigned* qubitIndices, unsigned numTargetQubits,
                              const rocComplex* matrixDevice, unsigned matrixDim);

/* ---- measure one qubit, collapse and renormalise (REF:172-177) ---- */
rocqStatus_t rocsvMeasure(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits,
                          unsigned qubitToMeasure, int* outcome, double* probability);

/* ---- named gates (REF:184-232) ---- */
rocqStatus_t rocsvApplyX(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned targetQubit);
rocqStatus_t rocsvApplyY(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned targetQubit);
rocqStatus_t rocsvApplyZ(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned targetQubit);
rocqStatus_t rocsvApplyH(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned targetQubit);
rocqStatus_t rocsvApplyS(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned targetQubit);
rocqStatus_t rocsvApplyT(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned targetQubit);
rocqStatus_t rocsvApplySdg(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned targetQubit);
rocqStatus_t rocsvApplyRx(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned targetQubit, double theta);
rocqStatus_t rocsvApplyRy(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned targetQubit, double theta);
rocqStatus_t rocsvApplyRz(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned targetQubit, double theta);

/* ---- two-qubit and controlled gates (REF:241-281) ---- */
rocqStatus_t rocsvApplyCNOT(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned controlQubit, unsigned targetQubit);
rocqStatus_t rocsvApplyCZ(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned qubit1, unsigned qubit2);
rocqStatus_t rocsvApplySWAP(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned qubit1, unsigned qubit2);
rocqStatus_t rocsvApplyCRX(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned controlQubit, unsigned targetQubit, double theta);
rocqStatus_t rocsvApplyCRY(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned controlQubit, unsigned targetQubit, double theta);
rocqStatus_t rocsvApplyCRZ(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned controlQubit, unsigned targetQubit, double theta);
rocqStatus_t rocsvApplyMultiControlledX(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, const unsigned* controlQubits, unsigned numControlQubits, unsigned targetQubit);
rocqStatus_t rocsvApplyCSWAP(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned controlQubit, unsigned targetQubit1, unsigned targetQubit2);

/* ---- readback (REF:286, 291) ---- */
rocqStatus_t rocsvGetStateVectorFull(rocsvHandle_t handle, rocComplex* d_state, rocComplex* h_state);
rocqStatus_t rocsvGetStateVectorSlice(rocsvHandle_t handle, rocComplex* d_state, rocComplex* h_state, unsigned batch_index);

/* ---- handle-owned pinned host staging buffer (REF:307, 316, 324) ---- */
rocqStatus_t rocsvEnsurePinnedBuffer(rocsvHandle_t handle, size_t minSizeBytes);
void* rocsvGetPinnedBufferPointer(rocsvHandle_t handle);
rocqStatus_t rocsvFreePinnedBuffer(rocsvHandle_t handle);

/* ---- expectation values (REF:340-423).  All are NON-destructive here (a superset of the
 *      reference wording for X/Y, REF:349, 367). ---- */
rocqStatus_t rocsvGetExpectationValueSinglePauliZ(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned targetQubit, double* result);
rocqStatus_t rocsvGetExpectationValueSinglePauliX(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned targetQubit, double* result);
rocqStatus_t rocsvGetExpectationValueSinglePauliY(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, unsigned targetQubit, double* result);
rocqStatus_t rocsvGetExpectationValuePauliProductZ(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, const unsigned* targetQubits, unsigned numTargetPaulis, double* result);
rocqStatus_t rocsvGetExpectationPauliString(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, const char* pauliString, const unsigned* targetQubits, unsigned numTargetPaulis, double* result);

/* ---- sampling: uint64 per shot, bit j = outcome of measuredQubits[j] (REF:439-445) ---- */
rocqStatus_t rocsvSample(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, const unsigned* measuredQubits, unsigned numMeasuredQubits, unsigned numShots, uint64_t* h_results);

/* ---- controlled k-qubit matrix (REF:461-468) and matrix + measure (REF:487-494) ---- */
rocqStatus_t rocsvApplyControlledMatrix(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, const unsigned* controlQubits, unsigned numControls, const unsigned* targetQubits, unsigned numTargets, const rocComplex* d_matrix);
rocqStatus_t rocsvApplyMatrixAndMeasure(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, const unsigned* targetQubits, unsigned numTargetQubits, const rocComplex* d_matrix, unsigned qubitToMeasure, int* outcome);

/* ======================================================================================
 * Extensions (not in the reference header).  They exist because the reference leaves the RNG,
 * state import, circuit-level submission and the multi-process layout unspecified.
 * ====================================================================================== */

/* sizeof(real_t) of this build: 4 (complex64) or 8 (complex128). */
unsigned rocsvxGetPrecisionBytes(void);

/* Seed of the Philox4x32-10 stream behind rocsvMeasure / rocsvSample (default 0).  Draw `shot` of the
 * c-th measuring call on a handle uses counter (shot, c); see DESIGN.md "sampling spec". */
rocqStatus_t rocsvxSetSeed(rocsvHandle_t handle, uint64_t seed);

/* Host -> device import of batchSize*2^n amplitudes (the reference has no import at all). */
rocqStatus_t rocsvxSetStateVector(rocsvHandle_t handle, rocComplex* d_state, const rocComplex* h_state);

/* Block until everything enqueued on the handle's stream has finished (flushes deferred gates). */
rocqStatus_t rocsvxSynchronize(rocsvHandle_t handle);

/* Deferred execution.  With fusion enabled, rocsvApply* calls validate, enqueue and return; the
 * queue is partitioned into fused HBM sweeps and launched by rocsvxFlush or by any entry point that
 * reads the state (readback, measure, sample, expectation, free, destroy).  Default: disabled, or
 * the value of the ROCQ_FUSION environment variable at rocsvCreate. */
rocqStatus_t rocsvxSetFusion(rocsvHandle_t handle, int enabled);
rocqStatus_t rocsvxFlush(rocsvHandle_t handle);

/* Circuit-level submission: the whole gate queue in one call (what Circuit.flush() /
 * GateFusion::processQueue feed one gate at a time in the reference, python/rocq/api.py:74-89,
 * GateFusion.cpp:89-156).  Always fused, regardless of rocsvxSetFusion. */
typedef enum {
    ROCSVX_H = 0, ROCSVX_X, ROCSVX_Y, ROCSVX_Z, ROCSVX_S, ROCSVX_SDG, ROCSVX_T,
    ROCSVX_RX, ROCSVX_RY, ROCSVX_RZ,                 /* targets[0], theta */
    ROCSVX_CNOT, ROCSVX_CZ, ROCSVX_SWAP,             /* CNOT: controlMask bit + targets[0]; CZ/SWAP: targets[0..1] */
    ROCSVX_CRX, ROCSVX_CRY, ROCSVX_CRZ,              /* controlMask (1 bit), targets[0], theta */
    ROCSVX_MCX,                                      /* controlMask, targets[0] */
    ROCSVX_CSWAP,                                    /* controlMask (1 bit), targets[0..1] */
    ROCSVX_MATRIX                                    /* numTargets, targets[], controlMask, matrix (HOST) */
} rocsvxGateKind;

typedef struct {
    int32_t kind;              /* rocsvxGateKind */
    uint32_t numTargets;
    uint32_t targets[8];
    uint64_t controlMask;      /* bit q set => qubit q is a control (acts where all controls are 1) */
    double theta;
    const double* matrix;      /* ROCSVX_MATRIX only: HOST, column-major 2^k x 2^k, interleaved (re,im) doubles */
} rocsvxGateOp;

/* rocsvxApplyCircuit keeps the launches it planned for the most recent circuit: resubmitting the identical gate list (same
 * gates, matrices, state buffer and settings; compared by a 128-bit hash of the whole list) replays them and skips fusion,
 * planning and the host-side matrix products.  Every kernel runs again; only host work is saved.  ROCQ_PLAN_CACHE=0 turns
 * it off. */
rocqStatus_t rocsvxApplyCircuit(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits,
                                const rocsvxGateOp* ops, size_t numOps);
rocqStatus_t rocsvxSetPlanCache(rocsvHandle_t handle, int enabled);     /* 0: every rocsvxApplyCircuit call plans afresh (and drops the kept plan) */

/* Tensor-core path (complex64 library only): apply one dense 6-qubit matrix (HOST, column-major 64x64, interleaved
 * (re,im) doubles, index bit b <-> qubits[b], any six distinct qubits) to the whole state in one HBM pass with tcgen05
 * MMAs (two-term fp16 split of both operands, fp32 accumulation in tensor memory; ~2e-7 relative error per block).
 * Needs numQubits >= 13.  rocsvxApplyCircuit and fused flushes use the same kernel for the 6-qubit blocks their planner
 * forms: rocsvxSetTensorCoreBlocks(h, 1) always, (h, 0) never, (h, -1) = default: from 24 qubits on.  Environment:
 * ROCQ_TC=0|1|auto, ROCQ_TC_MIN_COST. */
rocqStatus_t rocsvxApplyBlock6(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, const unsigned* qubits,
                               const double* matrix);
rocqStatus_t rocsvxSetTensorCoreBlocks(rocsvHandle_t handle, int enabled);

/* Fused execution merges runs of controlled phases / controlled one-qubit diagonals that share a control qubit (a QFT's
 * CP ladder after each H) into ONE diagonal pass per run instead of one per gate.  On by default; 0 switches it off
 * (environment: ROCQ_MERGE_DIAG=0). */
rocqStatus_t rocsvxSetMergeDiagonals(rocsvHandle_t handle, int enabled);

/* ||psi||^2 of batch member 0 (one read sweep). */
rocqStatus_t rocsvxGetNorm(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits, double* result);

/* Batched Pauli-string expectation: numTerms strings in one call (what python/rocq/api.py:520-643 get_expval / grad and
 * rocquantum/solvers/vqe_solver.py:120-136 evaluate one term and one full-state pass at a time).  paulis = concatenated
 * characters, qubits = concatenated qubit indices, offsets[t]..offsets[t+1] delimit term t (offsets has numTerms+1
 * entries).  results[t] = <psi|P_t|psi> of batch member 0.  Terms are grouped by the set of qubits carrying X or Y: every
 * group is ONE read sweep over the state (all terms made of I and Z only: one sweep in total), and all results return in
 * one device->host copy.  ...AllStates evaluates every state of the handle's batch (parameter-shift batches prepared
 * through batchSize): results[state * numTerms + t]. */
rocqStatus_t rocsvxGetExpectationPauliBatch(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits,
                                            const char* paulis, const unsigned* qubits, const unsigned* offsets,
                                            unsigned numTerms, double* results);
rocqStatus_t rocsvxGetExpectationPauliBatchAllStates(rocsvHandle_t handle, rocComplex* d_state, unsigned numQubits,
                                                     const char* paulis, const unsigned* qubits, const unsigned* offsets,
                                                     unsigned numTerms, double* results);

/* Counters since rocsvCreate / the last reset: what the engine actually launched. */
typedef struct {
    uint64_t kernelLaunches;   /* every kernel of this library */
    uint64_t sweeps;           /* launches of the fused tile-sweep kernel */
    uint64_t gatesSubmitted;   /* gates received through rocsvApply* / rocsvxApplyCircuit */
    uint64_t opsExecuted;      /* tile ops after algebraic fusion */
    uint64_t h2dBytes;         /* host->device bytes the engine itself moved: sweep programs (kernel parameters), matrices */
    double   lastSweepMs;      /* device time of the most recent flush (CUDA events), 0 if none */
    uint64_t exchanges;        /* multi-process: global<->local index-bit exchanges executed */
    uint64_t exchangeBytes;    /* multi-process: bytes this rank sent to peers in those exchanges */
    double   exchangeMs;       /* multi-process: device time of those exchanges (CUDA events around each) */
    uint64_t blockSweeps;      /* launches of the tensor-core block-sweep kernel (counted in `sweeps` too) */
    uint64_t planCacheHits;    /* rocsvxApplyCircuit calls that replayed the recorded launches of an identical earlier call */
    uint64_t expectationSweeps; /* read sweeps launched by the batched expectation calls (one per x-mask group) */
} rocsvxStats;
rocqStatus_t rocsvxGetStats(rocsvHandle_t handle, rocsvxStats* stats, int reset);

/* Device timer on the handle's stream (CUDA events): Start records, Stop records, flushes nothing and
 * returns the elapsed milliseconds after synchronising on the stop event. */
rocqStatus_t rocsvxTimerStart(rocsvHandle_t handle);
rocqStatus_t rocsvxTimerStop(rocsvHandle_t handle, double* milliseconds);

/* Host-only planning (no GPU needed): partition `ops` into sweeps exactly as the engine would and
 * describe the plan as text into buf (NUL-terminated, truncated to bufSize).  Returns the number of
 * sweeps through *numSweeps.  tileBits = 0 selects the engine default. */
rocqStatus_t rocsvxPlanCircuit(unsigned numQubits, unsigned tileBits, const rocsvxGateOp* ops, size_t numOps,
                               unsigned* numSweeps, char* buf, size_t bufSize);

/* Host-only: the mixed plan the engine follows when tensor-core blocks are enabled (complex64).  Text: one line
 * "B p0..p5" per 6-qubit block or "S ..." per ordinary sweep, each followed by its "O ..." op lines in execution order.
 * minCost <= 0 selects the engine default. */
rocqStatus_t rocsvxPlanCircuitBlocks(unsigned numQubits, const rocsvxGateOp* ops, size_t numOps, double minCost,
                                     unsigned* numBlocks, unsigned* numSweeps, char* buf, size_t bufSize);

/* ---- multi-process distribution (one process per GPU).  The 128-byte id is NCCL's unique id: rank 0
 *      obtains it and the host broadcasts it by any means (torch.distributed in bench.py). ---- */
rocqStatus_t rocsvxDistGetUniqueId(void* id128);
rocqStatus_t rocsvxDistInit(rocsvHandle_t handle, int rank, int numRanks, const void* id128);
rocqStatus_t rocsvxDistGetInfo(rocsvHandle_t handle, int* rank, int* numRanks, unsigned* numLocalQubits, rocComplex** d_localSlice);

/* ---- single-process distribution: the reference's model (MULTI_GPU_GUIDE.md:11-17, test_hipStateVec_multi_gpu.cpp:109-339,
 *      python/rocq/api.py:53-57).  rocsvAllocateDistributedState on a handle that is NOT a rank of a multi-process job shards
 *      the state over the visible devices from this one process: the largest power of two of them, at least two local
 *      qubits per slice (ROCQ_NUM_GPUS caps it).  Every rocsv* call on the handle then acts on all slices (pass d_state =
 *      NULL), scalar results are returned once, rocsvGetStateVectorFull / rocsvxSetStateVector move the WHOLE 2^n state
 *      (slice r at r * 2^numLocalQubits).  The devices must have peer access to each other.
 *      rocsvxDistSetRanks fixes the number of slices of the next allocation (a power of two; 0 = default).  More slices than
 *      devices are placed round-robin -- useful only to exercise the distributed engine on one GPU.
 *      rocsvxDistGetRankSlice: device ordinal and raw device pointer of one slice (what the reference's test reads through
 *      its test-visible handle, test_hipStateVec_multi_gpu.cpp:43-81). ---- */
rocqStatus_t rocsvxDistSetRanks(rocsvHandle_t handle, int numRanks);
rocqStatus_t rocsvxDistGetRankSlice(rocsvHandle_t handle, int rank, int* device, rocComplex** d_slice);

/* Host-only plan of a global<->local index-bit exchange: swapping the k = numPairs global bits
 * globalBits[] with the local bits localBits[] on `rank` of `numRanks`.  Writes up to maxSegs
 * segments {peer, sendOffset, recvOffset, count} in amplitudes; returns how many through *numSegs. */
typedef struct { int32_t peer; uint64_t sendOffset; uint64_t recvOffset; uint64_t count; } rocsvxExchangeSeg;
rocqStatus_t rocsvxDistPlanExchange(unsigned numLocalQubits, int numRanks, int rank,
                                    const unsigned* localBits, const unsigned* globalBits, unsigned numPairs,
                                    rocsvxExchangeSeg* segs, size_t maxSegs, size_t* numSegs);
/* The same exchange as the engine moves it with ROCQ_EXCHANGE=p2p (peer slices mapped through CUDA IPC): in-place swaps
 * own[sendOffset + i] <-> peer's slice[recvOffset + i], i < count.  The two ranks of a pair split every run -- the lower
 * rank swaps its first half, the higher rank the rest -- so each amplitude is moved by exactly one rank, nothing is staged. */
rocqStatus_t rocsvxDistPlanPeerSwap(unsigned numLocalQubits, int numRanks, int rank,
                                    const unsigned* localBits, const unsigned* globalBits, unsigned numPairs,
                                    rocsvxExchangeSeg* segs, size_t maxSegs, size_t* numSegs);

/* Host-only plan of a whole distributed circuit (no GPU, no NCCL): the steps the engine would execute on numRanks
 * ranks -- "R" blocks of ops in PHYSICAL qubit positions (positions >= numLocalQubits are rank bits) and "X g.."
 * lines (exchange the listed rank bits with the top local bits) -- followed by the final logical->physical map "M ..".
 * mode bit 0: 0 = rocsvxApplyCircuit (ops that need a rank bit are deferred together with what depends on them while the
 * rest of the circuit runs; exchanges evict the qubits whose next use is farthest), 1 = one rocsvApply* call per gate.
 * mode bit 1: also run algebraic fusion + sweep partition on every "R" block, as the engine does per slice ("S T
 * rowbits res: .." lines precede the ops of each sweep); with bit 2 the partition includes tensor-core blocks ("B p0..p5"
 * lines).  mode bit 3: rocsvxApplyCircuit strictly in program order (the engine with ROCQ_DIST_INORDER=1).
 * canonicalize != 0 appends the steps that restore the identity layout. */
rocqStatus_t rocsvxDistPlanCircuit(unsigned totalNumQubits, int numRanks, const rocsvxGateOp* ops, size_t numOps, int mode,
                                   int canonicalize, unsigned* numExchanges, char* buf, size_t bufSize);

#ifdef __cplusplus
}
#endif

#endif /* HIPSTATEVEC_H */
