// rocquantum_b200/csrc/sv_internal.h
// Shared host/device definitions of the sweep program that the tile-sweep kernel executes, and the
// thin internal C ABI (rq_*) between the C++ host engine and the .cu launchers.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef ROCQ_PRECISION_DOUBLE
typedef double rq_real;
#else
typedef float rq_real;
#endif
struct rq_cplx { rq_real x, y; };

// ---- sweep program ---------------------------------------------------------------------------------
// One launch of the tile-sweep kernel = one pass over HBM.  The state is cut into tiles of 2^T
// amplitudes: T "resident" qubit positions res[0..T) (ascending; local bit j of a tile <-> global
// position res[j]) and one tile per assignment of the other n-T bits.  A tile is staged in shared
// memory with 1-D bulk async copies (one per contiguous row of 2^rowbits amplitudes), every op of the
// program is applied to it in order, and it is written back with bulk async stores.
#define RQ_MAX_DIAGP 32
enum : uint8_t { RQ_OP_DENSE = 1, RQ_OP_DIAG = 2, RQ_OP_PERM = 3, RQ_OP_DIAGP = 4 };

struct rq_tile_op {                 // 64 bytes
    uint8_t kind;                   // RQ_OP_*
    uint8_t k;                      // DENSE: number of targets (1..4).  DIAG: number of table bits (0..3).  DIAGP: non-resident factors
    uint8_t nfix;                   // entries of fix[]: local positions held fixed while enumerating
    uint8_t ext;                    // DENSE: 1 => matrix is read from hdr.ext_matrix (device pointer)
    uint8_t t[4];                   // DENSE: local position of matrix bit b.  DIAG: local position of table bit b, 0xFF = non-resident
                                    // DIAGP: t[0] = per-thread factors (group-index bits 0..t[0]-1), t[1] = log2 of the table over the bits above,
                                    //        t[2] = slot of its per-tile factor (hdr.diagp_op[]), t[3] = index among the DIAGP ops of
                                    //        its register-window phase (0xFF: it has a pass of its own)
    uint8_t gq[4];                  // DIAG: global position of table bit b when non-resident
    uint8_t fix[16];                // ascending local positions (targets of DENSE/PERM and local controls)
    uint32_t setmask;               // OR-ed into the enumerated local index (controls = 1; PERM select value)
    uint32_t xm;                    // PERM: partner = idx ^ xm.  DIAGP: offset (from moff) of the non-resident factors
    uint32_t moff;                  // offset into pool[], in complex elements
    uint32_t cm_out;                // register phases: local controls outside the phase window (checked on the group base)
    uint64_t gcmask;                // controls on non-resident positions: op is skipped for tiles whose base lacks a bit
    uint8_t wt[4];                  // register phases: window-bit index (0..V-1) of target b
    uint8_t cm_in;                  // register phases: controls inside the window, in window-bit coordinates
    uint8_t fuse;                   // register phases, peephole of build_phases: RQ_FUSE_SKIP = this Hadamard is carried out by the
                                    // next op; RQ_FUSE_BUTTERFLY = this DIAGP first applies the Hadamard on its hub (window bit cm_in):
                                    // a0' = r (a0 + a1), a1' = r * phase * (a0 - a1) -- the radix-2 butterfly of a QFT
    uint8_t pad[2];
};
// RQ_FUSE_BUTTERFLY_UP: a butterfly whose ladder has no factor on the window bits BELOW its hub (every QFT ladder): the
// window part of the phase then depends only on the window bits above the hub, 2^(V-1-W) distinct values instead of 2^(V-1).
enum : uint8_t { RQ_FUSE_NONE = 0, RQ_FUSE_SKIP = 1, RQ_FUSE_BUTTERFLY = 2, RQ_FUSE_BUTTERFLY_UP = 3 };

// A phase is what happens between two shared-memory round trips of the tile.
//   kind 0: one op, applied in place in shared memory (any op).
//   kind 1: register window -- every thread loads the 2^v amplitudes that differ in the v window bits w[] (local
//           positions >= 4, so the loads are bank-conflict free), applies ops first..first+count-1 whose non-diagonal
//           targets all lie in the window entirely in registers, and stores them back once.
//   kind 2: a register window whose ops are nothing but (Hadamard-like, ladder) pairs fused into butterflies: executed
//           by a routine without the per-op interpreter (run_chain_phase).
struct rq_phase {
    uint8_t kind, v, first, count;
    uint8_t w[4];                   // ascending local positions
};

struct rq_sweep_hdr {
    uint32_t n;                     // qubits per state vector (local qubits in distributed mode)
    uint32_t T;                     // tile bits
    uint32_t nops;
    uint32_t rowbits;               // log2(amplitudes per contiguous row)
    uint32_t nphases;
    uint32_t max_phase_ops;         // largest rq_phase::count (selects the kernel variant)
    uint32_t swz;                   // 1: the tile lives XOR-swizzled in shared memory (ops on the lowest local bits)
    uint32_t pad;
    uint64_t ntiles;                // batch * 2^(n-T)
    uint64_t high_base;             // OR-ed into every tile's base index for predicates (rank << n_local)
    const void* ext_matrix;         // device matrix of an op with ext = 1 (column-major, rq_cplx)
    uint8_t res[16];                // ascending resident global positions
    uint8_t sres[16];               // where local bit j is STORED: res[j] unless trailing swaps of resident qubits were folded into
                                    // the store addressing (host_ops.h, build_program); identity on the row bits
    uint8_t ndiagp;                 // RQ_OP_DIAGP ops of the program; their per-tile factors are computed while the tile loads
    uint8_t diagp_op[RQ_MAX_DIAGP]; // op index of slot s
    uint8_t pad2[7];
};

template <int MAXOPS, int POOL_CPLX>
struct rq_program {
    rq_sweep_hdr hdr;
    rq_tile_op ops[MAXOPS];
    rq_phase phases[MAXOPS];
    alignas(16) rq_cplx pool[POOL_CPLX];
};
// Kernel parameters may be up to 32764 bytes (CUDA >= 12.1); the program travels as a
// __grid_constant__ parameter so that op headers and gate matrices are read through the constant
// cache with warp-uniform addresses.
#ifdef ROCQ_PRECISION_DOUBLE
typedef rq_program<8, 288> rq_program_small;        //  ~5.3 KB
typedef rq_program<160, 1248> rq_program_large;     // ~31.6 KB
#else
typedef rq_program<8, 288> rq_program_small;        //  ~3.0 KB
typedef rq_program<160, 2496> rq_program_large;     // ~31.6 KB
#endif

// pool slots per element of a dense matrix: complex64 stores (re,im) and the precomputed pair (-im,+im) for FFMA2
#ifdef ROCQ_PRECISION_DOUBLE
#define RQ_MSLOTS 1
#else
#define RQ_MSLOTS 2
#endif
#define RQ_TILE_THREADS 256
// Shared-memory bank geometry: an LDS.64 (complex64) wavefront is 16 lanes x 8 B, an LDS.128 (complex128) wavefront
// 8 lanes x 16 B, so the lowest RQ_SWZ_BITS index bits select the bank group.  Ops whose targets include one of those
// bits would serialise (all lanes of a wavefront agree on them), so such sweeps keep the tile XOR-swizzled:
//   phys(idx) = idx ^ ((idx >> RQ_SWZ_BITS) & (2^RQ_SWZ_BITS - 1))
#ifdef ROCQ_PRECISION_DOUBLE
#define RQ_SWZ_BITS 3
#else
#define RQ_SWZ_BITS 4
#endif
#define RQ_WINDOW_MIN_POS RQ_SWZ_BITS                // without the swizzle, register-window bits sit above the bank bits
#ifndef RQ_WINDOW_BITS
#ifdef ROCQ_PRECISION_DOUBLE
#define RQ_WINDOW_BITS 3                             // 8 amplitudes = 32 registers per thread
#else
#define RQ_WINDOW_BITS 4                             // 16 amplitudes = 32 registers per thread
#endif
#endif
#define RQ_PHASE_MAX_DIAGP RQ_WINDOW_BITS             // RQ_OP_DIAGP ops per register-window phase (their thread factors stay in registers)
#ifndef RQ_PHASED_MIN_BLOCKS
#define RQ_PHASED_MIN_BLOCKS 3                       // resident CTAs per SM the phased variant is compiled for (3 x 64 KB tiles; measured: QFT-33 c128 740 -> 644 ms vs 2)
#endif
#ifdef ROCQ_PRECISION_DOUBLE
#define RQ_MAX_TILE_BITS 12                          // 2^12 * 16 B = 64 KB
#define RQ_MIN_ROW_BITS 4                            // rows >= 256 B
#else
#define RQ_MAX_TILE_BITS 13                          // 2^13 * 8 B = 64 KB
#define RQ_MIN_ROW_BITS 5                            // rows >= 256 B
#endif

// ---- tensor-core block sweep (block_sweep.cu, complex64 only) ------------------------------------------------------
// One 6-qubit dense block applied to the whole state: tile = 64 block values x 128 columns (13 resident positions).
struct rq_block_params {
    uint32_t n;                     // qubits per state vector
    uint32_t T;                     // 13
    uint32_t renorm;                // 1: the block is unitary -> restore every column's norm in the epilogue
    uint32_t pad;                   // debug switches (ROCQ_BLOCK_DEBUG), 0 in production
    float scale;                    // power of two that brings amplitudes into the fp16 normal range
    uint32_t pad2;
    uint64_t ntiles;                // batch * 2^(n-13)
    uint8_t res[16];                // ascending resident positions (block + column bits)
    uint8_t blk[8];                 // 6 block positions, ascending: bit b of the block value <-> blk[b]
    uint8_t col[8];                 // 7 column positions, ascending
    uint32_t trank;                 // rank of the tensor map whose box is one tile
    uint8_t tbits[5];               // per tensor-map dimension: tile-index bits it consumes (0: resident, 255: all that remain)
    uint8_t lp_blk[6];              // bit of the tile-local amplitude index (shared-memory layout) that block bit b lands on
    uint8_t lp_col[7];              // ... and column bit b
    uint8_t pad3[2];
};
#define RQ_BLOCK_QUBITS 6
#define RQ_BLOCK_COLBITS 7
#define RQ_BLOCK_AUTO_QUBITS 24        // 'auto': tensor-core blocks from this many qubits (state beyond L2; below, launch overheads dominate)
#define RQ_BLOCK_UBYTES 32768        // Re U and Im U of the 64x64 block, two fp16 terms each, in UMMA K-major core-matrix order

// rocsvSample: result bit j of a shot = bit pos[j] of (high_base | sampled local index); nm = RQ_SHOT_RAW keeps the index itself.
// miss = what a shot owned by another rank of a distributed state leaves behind (0, so that a sum over ranks gathers).
#define RQ_SHOT_RAW 255u
#define RQ_SCAN_MAXSEG 1024
struct rq_shot_map {
    uint64_t high_base;
    uint64_t miss;
    uint32_t nm;
    uint8_t pos[64];
};

// One group of Pauli terms that share the x-mask (rocsvxGetExpectationPauliBatch): evaluated by one read sweep.
#define RQ_PAULI_GROUP_MAX 32
#define RQ_PAULI_GROUP_TERMS 16       // terms the engine puts in one group (one read sweep)
struct rq_pauli_group {
    uint64_t xmask;
    uint32_t nterms;
    uint32_t pad;
    uint64_t zmask[RQ_PAULI_GROUP_MAX];
    uint32_t index[RQ_PAULI_GROUP_MAX];   // position of the term in the caller's list
    uint8_t ny[RQ_PAULI_GROUP_MAX];       // number of Y factors
};

// ---- thin C ABI to the launchers (all return a cudaError_t as int; stream is a cudaStream_t) --------
extern "C" {
int rq_launch_sweep_small(rq_cplx* state, const rq_program_small* prog, void* stream);
int rq_launch_sweep_large(rq_cplx* state, const rq_program_large* prog, void* stream);
int rq_sweep_configure(void);    // opt in to > 48 KB dynamic shared memory; call once per device

// generic k-qubit dense matrix with controls, matrix in device memory (column-major); any k <= 10
int rq_launch_gather(rq_cplx* state, unsigned n, size_t batch, const unsigned* h_targets, unsigned k,
                     uint64_t cmask, const rq_cplx* d_matrix, void* stream);

int rq_launch_init_state(rq_cplx* state, size_t total_amps, int write_one, void* stream);

// reductions: results land in d_out (device doubles / uint64), caller copies them back
int rq_launch_pauli_expect(const rq_cplx* state, unsigned n, uint64_t xmask, uint64_t zmask, unsigned ny,
                           double* d_partials, unsigned nblocks, double* d_out, void* stream);
// d_partials: RBLOCKS * RQ_PAULI_GROUP_MAX * nstates doubles; results[state * num_terms_total + G->index[t]]
int rq_launch_pauli_group(const rq_cplx* state, unsigned n, unsigned nstates, const rq_pauli_group* G, unsigned num_terms_total,
                          double* d_partials, double* d_results, void* stream);
int rq_launch_fixed_masses(const rq_cplx* state, unsigned n, unsigned q, uint64_t* d_partials, unsigned nblocks,
                           uint64_t* d_out4, void* stream);
int rq_launch_collapse(rq_cplx* state, unsigned n, unsigned q, int outcome, double scale, void* stream);
int rq_launch_chunk_masses(const rq_cplx* state, unsigned n, unsigned chunk_bits, uint64_t* d_chunk_hi,
                           uint64_t* d_chunk_lo, void* stream);
// exact inclusive scan of count 128-bit masses (hi[], lo[]) in place; d_btot: 2 * RQ_SCAN_MAXSEG words of scratch; the grand
// total lands in d_total2[0..1]
int rq_launch_scan_masses(uint64_t* d_hi, uint64_t* d_lo, uint64_t count, uint64_t* d_btot, uint64_t* d_total2, void* stream);
// d_totals4 = {total.hi, total.lo, win.hi, win.lo} in device memory; map: how a sampled basis index becomes a result word
int rq_launch_sample(const rq_cplx* state, unsigned n, unsigned chunk_bits, const uint64_t* d_incl_hi,
                     const uint64_t* d_incl_lo, uint64_t nchunks, const uint64_t* d_totals4, uint64_t seed, uint64_t call,
                     unsigned shots, uint64_t shot_offset, const rq_shot_map* map, uint64_t* d_indices, void* stream);
unsigned rq_reduce_blocks(void);
int rq_block_configure(void);
// d_uterms: RQ_BLOCK_UBYTES device bytes (+ 256 bytes of debug counters)
// tensor_map: host pointer to a 128-byte CUtensorMap (P->trank dims), or NULL
int rq_launch_block_sweep(rq_cplx* state, const rq_block_params* P, const void* d_uterms, const void* tensor_map, void* stream);
}
