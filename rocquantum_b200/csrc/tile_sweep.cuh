#pragma once
// rocquantum_b200/csrc/tile_sweep.cuh  --  the fused gate sweep, hand-written for sm_100a.
// (kernel templates; instantiated per program type / layout in tile_sweep_{small,large}{,_swz}.cu to build in parallel)
//
// One launch = one pass over HBM: 2 * 2^n * sizeof(amp) algorithmic bytes, however many gates the
// program carries.  It replaces, for every gate of the reference's path, the one-kernel-per-gate
// grid-stride loops of /root/reference/rocquantum/src/hipStateVec/{single,two,multi}_qubit_kernels.hip
// (launched from hipStateVec.cpp:100-186, 431-687): same arithmetic per amplitude pair
// (single_qubit_kernels.hip:64-67), but
//   * a tile of 2^T amplitudes (T resident qubit positions) is staged in shared memory by the TMA
//     engine: one cp.async.bulk (SASS UBLKCP) per contiguous row, completion on an mbarrier, so no
//     thread spends registers or issue slots on the copy and several tiles per SM are in flight;
//   * every op of the program is applied to the resident tile (dense 1..4-qubit matrices with
//     controls, diagonal phases, pair permutations), with gate matrices read warp-uniformly from the
//     constant bank of the __grid_constant__ program;
//   * controls and diagonal factors on NON-resident qubits are resolved per tile from the tile's base
//     index, so they never force a qubit to be resident;
//   * the tile goes back with bulk async stores (smem -> global).
#include <cuda_runtime.h>
#include <stdint.h>

#include "sv_internal.h"

namespace {

constexpr int NT = RQ_TILE_THREADS;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool SWZ>
__device__ __forceinline__ uint32_t sidx(uint32_t idx) {
    return SWZ ? (idx ^ ((idx >> RQ_SWZ_BITS) & ((1u << RQ_SWZ_BITS) - 1u))) : idx;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ rq_cplx ldg_cplx(const rq_cplx* p) {
#ifdef ROCQ_PRECISION_DOUBLE
    const double2 v = __ldg(reinterpret_cast<const double2*>(p));
#else
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
#endif
    rq_cplx r;
    r.x = v.x;
    r.y = v.y;
    return r;
}
// ---- complex multiply-accumulate -----------------------------------------------------------------------
// complex64: Blackwell's packed fp32 FMA (PTX fma.rn.f32x2, SASS FFMA2) does one complex MAC in TWO instructions:
//   acc(re,im) += m.re * (v.re, v.im);   acc(re,im) += m.im * (-v.im, v.re)
// (scalar-broadcast and half-swap are operand modifiers of FFMA2), instead of the 6 (FMUL+FFMA+FADD per component)
// that `acc += m*v` compiles to without reassociation.  complex128: explicit 4-FMA chains (DFMA).
#ifdef ROCQ_PRECISION_DOUBLE
struct cin { double x, y; };
struct cacc { double x, y; };
__device__ __forceinline__ cin cprep(const rq_cplx a) { return cin{a.x, a.y}; }
__device__ __forceinline__ cacc czero() { return cacc{0.0, 0.0}; }
__device__ __forceinline__ void cmac(cacc& acc, const rq_cplx m, const cin v) {
    acc.x = fma(m.x, v.x, acc.x);
    acc.x = fma(-m.y, v.y, acc.x);
    acc.y = fma(m.x, v.y, acc.y);
    acc.y = fma(m.y, v.x, acc.y);
}
__device__ __forceinline__ rq_cplx cget(const cacc a) { return rq_cplx{a.x, a.y}; }
#else
struct cin { uint64_t p, q; };            // (re, im) and (-im, re)
typedef uint64_t cacc;
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ cin cprep(const rq_cplx a) { return cin{pack2(a.x, a.y), pack2(-a.y, a.x)}; }
__device__ __forceinline__ cacc czero() { return 0ull; }
__device__ __forceinline__ void cmac(cacc& acc, const rq_cplx m, const cin v) {
    acc = fma2(pack2(m.x, m.x), v.p, acc);
    acc = fma2(pack2(m.y, m.y), v.q, acc);
}
__device__ __forceinline__ rq_cplx cget(const cacc a) {
    rq_cplx r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(a));
    return r;
}
#endif
__device__ __forceinline__ rq_cplx cmul(rq_cplx a, rq_cplx b) {
    cacc acc = czero();
    cmac(acc, a, cprep(b));
    return cget(acc);
}

// ---- register amplitudes of the window phases -------------------------------------------------------------------
// complex64: an amplitude stays a packed 64-bit (re,im) register pair from LDS.64 to STS.64, and a complex MAC is
//   acc += (m.re,m.re) * (v.re,v.im);  acc += (-m.im,+m.im) * (v.im,v.re)
// i.e. two FFMA2 with no per-amplitude fix-up: the scalar broadcast and the half swap are FFMA2 operand modifiers and
// the pair (-im,+im) is precomputed by the host in the second pool slot of every dense matrix element.
#ifdef ROCQ_PRECISION_DOUBLE
typedef rq_cplx ramp;
struct mel { double re, im; };
__device__ __forceinline__ ramp ramp_load(const rq_cplx* sm, uint32_t i) { return sm[i]; }
__device__ __forceinline__ void ramp_store(rq_cplx* sm, uint32_t i, const ramp v) { sm[i] = v; }
__device__ __forceinline__ mel mload(const rq_cplx* M, int e) { const rq_cplx m = M[e]; return mel{m.x, m.y}; }
__device__ __forceinline__ ramp rzero() { return rq_cplx{0.0, 0.0}; }
__device__ __forceinline__ void rmac(ramp& acc, const mel m, const ramp v) {
    acc.x = fma(m.re, v.x, acc.x);
    acc.x = fma(-m.im, v.y, acc.x);
    acc.y = fma(m.re, v.y, acc.y);
    acc.y = fma(m.im, v.x, acc.y);
}
__device__ __forceinline__ ramp rmul(const rq_cplx d, const ramp v) { return cmul(d, v); }
__device__ __forceinline__ ramp rsel(bool on, const ramp r, const ramp a) { return rq_cplx{on ? r.x : a.x, on ? r.y : a.y}; }
__device__ __forceinline__ ramp radd(const ramp a, const ramp b) { return rq_cplx{a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ ramp rsub(const ramp a, const ramp b) { return rq_cplx{a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ ramp rscale(rq_real r, const ramp a) { return rq_cplx{r * a.x, r * a.y}; }
#else
typedef uint64_t ramp;
struct mel { float re; uint64_t im2; };
__device__ __forceinline__ ramp ramp_load(const rq_cplx* sm, uint32_t i) { return *reinterpret_cast<const uint64_t*>(sm + i); }
__device__ __forceinline__ void ramp_store(rq_cplx* sm, uint32_t i, const ramp v) { *reinterpret_cast<uint64_t*>(sm + i) = v; }
__device__ __forceinline__ mel mload(const rq_cplx* M, int e) {
    const float4 q = *reinterpret_cast<const float4*>(M + 2 * e);      // (re, im, -im, +im): one 128-bit constant load
    return mel{q.x, pack2(q.z, q.w)};
}
__device__ __forceinline__ ramp rzero() { return 0ull; }
__device__ __forceinline__ uint64_t swap2(uint64_t v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return pack2(hi, lo);
}
__device__ __forceinline__ void rmac(ramp& acc, const mel m, const ramp v) {
    acc = fma2(pack2(m.re, m.re), v, acc);
    acc = fma2(m.im2, swap2(v), acc);
}
__device__ __forceinline__ ramp rmul(const rq_cplx d, const ramp v) {
    ramp acc = fma2(pack2(d.x, d.x), v, 0ull);
    return fma2(pack2(-d.y, d.y), swap2(v), acc);
}
__device__ __forceinline__ ramp rsel(bool on, const ramp r, const ramp a) { return on ? r : a; }
__device__ __forceinline__ ramp radd(const ramp a, const ramp b) { return fma2(pack2(1.f, 1.f), b, a); }
__device__ __forceinline__ ramp rsub(const ramp a, const ramp b) { return fma2(pack2(-1.f, -1.f), b, a); }
__device__ __forceinline__ ramp rscale(rq_real r, const ramp a) { return fma2(pack2(r, r), a, 0ull); }
#endif

// deposit the bits of g around the fixed positions fix[0..nfix) (ascending), leaving zeros there.
// Opening a zero at position p is  g + (g & ~(2^p - 1)); the masks of the first four positions are loaded once per op
// (0xffffffff = "no position": a no-op), so that the per-group cost is two ALU instructions per position and no loop.
struct fixmasks { uint32_t m[4]; uint32_t nfix; };
__device__ __forceinline__ fixmasks load_fix(const rq_tile_op& o) {
    fixmasks F;
    F.nfix = o.nfix;
#pragma unroll
    for (uint32_t f = 0; f < 4; ++f) F.m[f] = f < F.nfix ? ((1u << o.fix[f]) - 1u) : 0xffffffffu;
    return F;
}
__device__ __forceinline__ uint32_t spread(uint32_t g, const rq_tile_op& o, const fixmasks& F) {
#pragma unroll
    for (uint32_t f = 0; f < 4; ++f) g += g & ~F.m[f];
    for (uint32_t f = 4; f < F.nfix; ++f) {
        const uint32_t p = o.fix[f];
        g = ((g >> p) << (p + 1)) | (g & ((1u << p) - 1u));
    }
    return g;
}

// f(g, it) for every group g = tid + 2^8 * it of the thread.
// Build switches, measured on B200 (profiles/r01_loop_variants.log).  Window phases SELECT results on controls outside the
// window instead of branching (-DRQ_BRANCH_PRED restores the branch): without the divergent region the phase's matrices
// and tables are fetched through the uniform datapath (LDCU) instead of vector-indexed LDC; VQE ansatz 24.3 -> 22.7 ms,
// QFT-30 complex128 108 -> 103 ms.  -DRQ_UNIFORM_LOOPS additionally gives tiles of >= 2^8 groups a warp-uniform trip
// count: 8 % faster on heavily fused complex64 sweeps (678 -> 623 ms on configs[1] without tensor-core blocks) but the
// longer uniform-datapath instruction stream costs 60 % on CNOT-heavy phases, so the default keeps thread-indexed loops.
template <typename F>
__device__ __forceinline__ void for_groups(uint32_t ngroups, uint32_t tid, F&& f) {
    static_assert(NT == 256, "group index = tid + 2^8 * it");
#ifndef RQ_UNIFORM_LOOPS
    for (uint32_t g = tid, it = 0; g < ngroups; g += NT, ++it) f(g, it);
    return;
#endif
    if (ngroups >= NT) {
        for (uint32_t it = 0; it < (ngroups >> 8); ++it) f(tid + (it << 8), it);
    } else if (tid < ngroups) {
        f(tid, 0u);
    }
}

template <int K, bool EXT, bool SWZ>
__device__ __forceinline__ void op_dense(rq_cplx* sm, const rq_tile_op& o, const rq_cplx* pool, const rq_cplx* ext,
                                         uint32_t T, uint32_t tid) {
    constexpr int D = 1 << K;
    uint32_t off[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < K; ++b)
            if ((j >> b) & 1) v |= 1u << o.t[b];
        off[j] = v;
    }
    const uint32_t ngroups = 1u << (T - o.nfix);
    const rq_cplx* M = EXT ? ext : (pool + o.moff);
    const fixmasks F = load_fix(o);
    const uint32_t setmask = o.setmask;
    for_groups(ngroups, tid, [&](uint32_t g, uint32_t) {
        const uint32_t base = spread(g, o, F) | setmask;
        cin a[D];
#pragma unroll
        for (int j = 0; j < D; ++j) a[j] = cprep(sm[sidx<SWZ>(base | off[j])]);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            cacc acc = czero();
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const rq_cplx m = EXT ? ldg_cplx(M + i + j * D) : M[(i + j * D) * RQ_MSLOTS];   // column-major, as the API
                cmac(acc, m, a[j]);
            }
            sm[sidx<SWZ>(base | off[i])] = cget(acc);
        }
    });
}

// diagonal: amp[idx] *= d[sel], sel bit b taken from the local index or, for a non-resident qubit,
// from the tile base.  Qubits whose "0" entries are all 1 were turned into controls by the host.
template <bool SWZ>
__device__ __forceinline__ void op_diag(rq_cplx* sm, const rq_tile_op& o, const rq_cplx* pool, uint32_t T, uint32_t tid,
                                        uint64_t gbase) {
    uint32_t selbase = 0;
    for (uint32_t b = 0; b < o.k; ++b)
        if (o.t[b] == 0xFF) selbase |= (uint32_t)((gbase >> o.gq[b]) & 1ull) << b;
    const uint32_t ngroups = 1u << (T - o.nfix);
    const rq_cplx* D = pool + o.moff;
    const fixmasks F = load_fix(o);
    const uint32_t setmask = o.setmask;
    for_groups(ngroups, tid, [&](uint32_t g, uint32_t) {
        const uint32_t idx = spread(g, o, F) | setmask;
        uint32_t sel = selbase;
        for (uint32_t b = 0; b < o.k; ++b)
            if (o.t[b] != 0xFF) sel |= ((idx >> o.t[b]) & 1u) << b;
        const uint32_t pi = sidx<SWZ>(idx);
        sm[pi] = cmul(D[sel], sm[pi]);
    });
}

// product of one-qubit diagonals under common controls (host: merge_diagonals, a QFT's ladder of controlled phases):
//   amp[idx] *= C * prod_{bit set} f_bit.
// The factors of the NON-resident qubits (bits of the tile's `outer` word) and C are the same for the whole tile:
// diagp_tile_factors computes them, one warp per op, while the bulk loads of the tile are in flight.  With 2^8 threads the
// low 8 bits of the group index are the thread's own, so their factors fold into one per-thread constant; the group-index
// bits above are warp-uniform and read a small host-built table: two complex multiplies per amplitude.
// (MODE 2) The same warp goes on to build the op's two 16-entry thread-factor tables -- low four group-index bits (times the
// tile factor), high four -- for the first RQ_TFAC_SLOTS ops of the sweep: a butterfly-chain phase then needs neither a
// scan over its ops, a table build nor a barrier of its own before it starts (they were 13 % of the stall samples and most
// of the barrier stalls of a QFT sweep, profiles/r02_ncu_qft28_c128_final_lines.txt).
constexpr uint32_t RQ_TFAC_SLOTS = sizeof(rq_real) == 8 ? 12u : 16u;     // 6 KB / 4 KB of shared memory
template <bool TABLES, typename Prog>
__device__ __forceinline__ void diagp_tile_factors(const Prog& prog, rq_cplx* gfac, rq_cplx (*tall)[32], uint32_t tid, uint64_t outer) {
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    for (uint32_t s = warp; s < prog.hdr.ndiagp; s += NT / 32) {
        const rq_tile_op& o = prog.ops[prog.hdr.diagp_op[s]];
        const rq_cplx* P = prog.pool + o.moff;
        const uint32_t ng = o.k;
        const rq_cplx* G = P + o.xm;
        const uint8_t* gbit = reinterpret_cast<const uint8_t*>(G + ng);
        rq_cplx f = lane == 0 ? P[0] : rq_cplx{(rq_real)1, (rq_real)0};
        for (uint32_t j = lane; j < ng; j += 32)
            if ((outer >> gbit[j]) & 1ull) f = cmul(G[j], f);
#pragma unroll
        for (uint32_t d = 16; d > 0; d >>= 1) {
            rq_cplx g;
            g.x = __shfl_xor_sync(0xffffffffu, f.x, d);
            g.y = __shfl_xor_sync(0xffffffffu, f.y, d);
            f = cmul(g, f);
        }
        if (lane == 0) gfac[s] = f;                          // (after the xor reduction every lane holds the product)
        if (TABLES && s < RQ_TFAC_SLOTS && o.t[3] != 0xFF) {
            const rq_cplx* A = P + 1;
            const uint32_t na = o.t[0], first_bit = (lane >> 4) * 4u;
            rq_cplx t = lane < 16u ? f : rq_cplx{(rq_real)1, (rq_real)0};
#pragma unroll
            for (uint32_t b = 0; b < 4; ++b) {
                const uint32_t i = first_bit + b;
                if (i < na && ((lane >> b) & 1u)) t = cmul(A[i], t);
            }
            tall[s][lane] = t;
        }
    }
}

// tile factor x the factors of the thread's own group-index bits
__device__ __forceinline__ rq_cplx diagp_thread_factor(const rq_tile_op& o, const rq_cplx* pool, uint32_t tid, const rq_cplx* gfac) {
    static_assert(NT == 256, "the per-thread factor covers group-index bits 0-7");
    const rq_cplx* A = pool + o.moff + 1;
    const uint32_t na = o.t[0];
    rq_cplx f = gfac[o.t[2]];
    for (uint32_t i = 0; i < na; ++i) {                      // select, not branch: A[i] stays a uniform load
        const rq_cplx fm = cmul(A[i], f);
        const bool bit = (tid >> i) & 1u;
        f.x = bit ? fm.x : f.x;
        f.y = bit ? fm.y : f.y;
    }
    return f;
}

template <bool SWZ>
__device__ __forceinline__ void op_diagp(rq_cplx* sm, const rq_tile_op& o, const rq_cplx* pool, uint32_t T, uint32_t tid,
                                         const rq_cplx* gfac) {
    const rq_cplx* B = pool + o.moff + 1 + o.t[0];
    const rq_cplx f = diagp_thread_factor(o, pool, tid, gfac);
    const uint32_t ngroups = 1u << (T - o.nfix);
    const fixmasks F = load_fix(o);
    const uint32_t setmask = o.setmask;
    for_groups(ngroups, tid, [&](uint32_t g, uint32_t m) {
        const uint32_t pi = sidx<SWZ>(spread(g, o, F) | setmask);
        sm[pi] = cmul(cmul(B[m], f), sm[pi]);
    });
}

// pair permutation: swap(idx, idx ^ xm) over the idx whose fixed bits equal setmask
template <bool SWZ>
__device__ __forceinline__ void op_perm(rq_cplx* sm, const rq_tile_op& o, uint32_t T, uint32_t tid) {
    const uint32_t ngroups = 1u << (T - o.nfix);
    const fixmasks F = load_fix(o);
    const uint32_t setmask = o.setmask, xm = o.xm;
    for_groups(ngroups, tid, [&](uint32_t g, uint32_t) {
        const uint32_t l0 = spread(g, o, F) | setmask, i0 = sidx<SWZ>(l0), i1 = sidx<SWZ>(l0 ^ xm);
        const rq_cplx a = sm[i0], b = sm[i1];
        sm[i0] = b;
        sm[i1] = a;
    });
}

// ---- register-window phases ---------------------------------------------------------------------------
// Every thread owns the D = 2^V amplitudes that differ in the V window bits; all ops of the phase act on them in
// registers, so the tile makes ONE shared-memory round trip per phase instead of one per op.  Window bits sit at
// local positions >= 4: for a fixed register slot the lanes of a warp read consecutive amplitudes (no bank conflicts).
// `on`: the thread's group satisfies the op's controls OUTSIDE the window.  Results are selected, not branched around: a
// phase without divergent control flow lets every op header / matrix / table load go through the uniform datapath
// (a divergent region turns them into vector-indexed LDC, which the address-divergence unit serialises).
template <int V, int W, bool CTRL>
__device__ __forceinline__ void win_dense1(ramp (&a)[1 << V], const rq_cplx* M, uint32_t cm_in, bool on = true) {
    const mel m00 = mload(M, 0), m10 = mload(M, 1), m01 = mload(M, 2), m11 = mload(M, 3);          // column-major
#pragma unroll
    for (int j = 0; j < (1 << V); ++j) {
        if (j & (1 << W)) continue;
        if (CTRL && (j & cm_in) != cm_in) continue;
        const ramp a0 = a[j], a1 = a[j | (1 << W)];
        ramp r0 = rzero(), r1 = rzero();
        rmac(r0, m00, a0); rmac(r0, m01, a1);
        rmac(r1, m10, a0); rmac(r1, m11, a1);
        a[j] = CTRL ? rsel(on, r0, a0) : r0;
        a[j | (1 << W)] = CTRL ? rsel(on, r1, a1) : r1;
    }
}
// matrix bit 0 <-> window bit W0, matrix bit 1 <-> window bit W1 (the host orders the targets ascending)
template <int V, int W0, int W1, bool CTRL>
__device__ __forceinline__ void win_dense2(ramp (&a)[1 << V], const rq_cplx* M, uint32_t cm_in, bool on = true) {
    mel m[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) m[e] = mload(M, e);
#pragma unroll
    for (int j = 0; j < (1 << V); ++j) {
        if (j & ((1 << W0) | (1 << W1))) continue;
        if (CTRL && (j & cm_in) != cm_in) continue;
        const ramp x0 = a[j], x1 = a[j | (1 << W0)], x2 = a[j | (1 << W1)], x3 = a[j | (1 << W0) | (1 << W1)];
        ramp r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ramp acc = rzero();
            rmac(acc, m[i], x0); rmac(acc, m[i + 4], x1); rmac(acc, m[i + 8], x2); rmac(acc, m[i + 12], x3);
            r[i] = acc;
        }
        if (CTRL) { r[0] = rsel(on, r[0], x0); r[1] = rsel(on, r[1], x1); r[2] = rsel(on, r[2], x2); r[3] = rsel(on, r[3], x3); }
        a[j] = r[0]; a[j | (1 << W0)] = r[1]; a[j | (1 << W1)] = r[2]; a[j | (1 << W0) | (1 << W1)] = r[3];
    }
}
template <int V, int W>
__device__ __forceinline__ void win_x(ramp (&a)[1 << V], uint32_t cm_in, bool on) {
#pragma unroll
    for (int j = 0; j < (1 << V); ++j) {
        if (j & (1 << W)) continue;
        if ((j & cm_in) != cm_in) continue;
        const ramp t = a[j], u = a[j | (1 << W)];
        a[j] = rsel(on, u, t);
        a[j | (1 << W)] = rsel(on, t, u);
    }
}
template <int V, int W0, int W1>
__device__ __forceinline__ void win_swap(ramp (&a)[1 << V], uint32_t cm_in, bool on) {
#pragma unroll
    for (int j = 0; j < (1 << V); ++j) {
        if (j & ((1 << W0) | (1 << W1))) continue;
        if ((j & cm_in) != cm_in) continue;
        const ramp t = a[j | (1 << W0)], u = a[j | (1 << W1)];
        a[j | (1 << W0)] = rsel(on, u, t);
        a[j | (1 << W1)] = rsel(on, t, u);
    }
}

// Hadamard-like matrix on window bit W (real entries +-v, v = 1/sqrt 2) followed by the phase ladder hanging on that qubit
// (RQ_FUSE_BUTTERFLY): the radix-2 butterfly of a QFT,
//   a0' = c0 (a0 +- a1),   a1' = (c1 * phase(j)) (a0 -+ a1),   c0 = m00, c1 = m10,   SWP: a0' takes the difference
// 14 FP operations per pair instead of 16 (dense 2x2) + 8 (phase) + selects.  fc = c1 * (tile, thread and table factors).
template <int V, int W, bool SWP, bool UP>
__device__ __forceinline__ void win_butterfly(ramp (&a)[1 << V], const rq_cplx* Wt, const rq_cplx fc, rq_real c0) {
    // UP: the ladder has no factor on the window bits below its hub, so the phase of a pair depends only on the window
    // bits above W: 2^(V-1-W) complex products instead of 2^(V-1)
    constexpr int NU = UP ? (1 << (V - 1 - W)) : (1 << (V - 1));
    rq_cplx ph[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const int j = UP ? ((u << (W + 1)) | (1 << W)) : ((((u >> W) << (W + 1)) | (u & ((1 << W) - 1))) | (1 << W));
        ph[u] = cmul(Wt[j], fc);
    }
#pragma unroll
    for (int j = 0; j < (1 << V); ++j) {
        if (j & (1 << W)) continue;
        const ramp a0 = a[j], a1 = a[j | (1 << W)];
        const ramp s = radd(a0, a1), d = rsub(a0, a1);
        a[j] = rscale(c0, SWP ? d : s);
        const int u = UP ? (j >> (W + 1)) : (((j >> (W + 1)) << W) | (j & ((1 << W) - 1)));
        a[j | (1 << W)] = rmul(ph[u], SWP ? s : d);
    }
}
template <int V, int W>
__device__ __forceinline__ void win_butterfly_sel(ramp (&a)[1 << V], const rq_cplx* Wt, const rq_cplx fc, rq_real c0, bool swp, bool up) {
    if (up) {
        if (swp) win_butterfly<V, W, true, true>(a, Wt, fc, c0);
        else win_butterfly<V, W, false, true>(a, Wt, fc, c0);
    } else {
        if (swp) win_butterfly<V, W, true, false>(a, Wt, fc, c0);
        else win_butterfly<V, W, false, false>(a, Wt, fc, c0);
    }
}
template <int V>
__device__ __forceinline__ void win_butterfly_any(ramp (&a)[1 << V], uint32_t ci, const rq_cplx* Wt, const rq_cplx fc, rq_real c0, bool swp, bool up) {
    if (ci == 1u) win_butterfly_sel<V, 0>(a, Wt, fc, c0, swp, up);
    else if (ci == 2u) win_butterfly_sel<V, 1>(a, Wt, fc, c0, swp, up);
    else if (ci == 4u) win_butterfly_sel<V, 2>(a, Wt, fc, c0, swp, up);
    else if (V > 3) win_butterfly_sel<V, (V > 3 ? 3 : 0)>(a, Wt, fc, c0, swp, up);
}

template <int V>
__device__ __forceinline__ void win_dispatch1(ramp (&a)[1 << V], const rq_tile_op& o, const rq_cplx* M, bool dense, bool on) {
    const uint32_t w = o.wt[0], ci = o.cm_in;
    if (dense && ci == 0 && o.cm_out == 0) {                                  // uncontrolled: straight-line code, matrix stays in registers
        if (w == 0) win_dense1<V, 0, false>(a, M, 0);
        else if (w == 1) win_dense1<V, 1, false>(a, M, 0);
        else if (w == 2) win_dense1<V, 2, false>(a, M, 0);
        else if (V > 3) win_dense1<V, (V > 3 ? 3 : 0), false>(a, M, 0);
    } else if (dense) {
        if (w == 0) win_dense1<V, 0, true>(a, M, ci, on);
        else if (w == 1) win_dense1<V, 1, true>(a, M, ci, on);
        else if (w == 2) win_dense1<V, 2, true>(a, M, ci, on);
        else if (V > 3) win_dense1<V, (V > 3 ? 3 : 0), true>(a, M, ci, on);
    } else {
        if (w == 0) win_x<V, 0>(a, ci, on);
        else if (w == 1) win_x<V, 1>(a, ci, on);
        else if (w == 2) win_x<V, 2>(a, ci, on);
        else if (V > 3) win_x<V, (V > 3 ? 3 : 0)>(a, ci, on);
    }
}
template <int V>
__device__ __forceinline__ void win_dispatch2(ramp (&a)[1 << V], const rq_tile_op& o, const rq_cplx* M, bool dense, bool on) {
    const uint32_t pair = o.wt[0] * 4u + o.wt[1], ci = o.cm_in;      // wt[0] < wt[1]
    const bool plain = ci == 0 && o.cm_out == 0;
#define RQ_PAIR(A, B)                                            \
    case (A) * 4 + (B):                                          \
        if (dense && plain) win_dense2<V, A, B, false>(a, M, 0); \
        else if (dense) win_dense2<V, A, B, true>(a, M, ci, on); \
        else win_swap<V, A, B>(a, ci, on);                       \
        break;
    switch (pair) {
        RQ_PAIR(0, 1) RQ_PAIR(0, 2) RQ_PAIR(1, 2)
        default:
            if (V > 3) {
                switch (pair) {
                    RQ_PAIR(0, (V > 3 ? 3 : 1)) RQ_PAIR(1, (V > 3 ? 3 : 2)) RQ_PAIR(2, (V > 3 ? 3 : 2) + (V > 3 ? 0 : 1))
                    default: break;
                }
            }
            break;
    }
#undef RQ_PAIR
}

// The factor of the thread's own eight group-index bits is the same product for every tile and costs eight dependent
// complex multiply + select steps per op when every thread forms it alone (a third of a QFT phase).  Instead one warp
// per op builds two 16-entry tables in shared memory -- low four bits (times the tile factor), high four bits -- and
// every thread multiplies its two entries.
template <typename Prog>
__device__ __forceinline__ void phase_thread_factors(const Prog& prog, const rq_phase& ph, uint32_t tid, const rq_cplx* gfac,
                                                     rq_cplx (*tfac)[32], rq_cplx (&fA)[RQ_PHASE_MAX_DIAGP]) {
#pragma unroll
    for (int d = 0; d < RQ_PHASE_MAX_DIAGP; ++d) fA[d] = rq_cplx{(rq_real)1, (rq_real)0};
    if (!prog.hdr.ndiagp) return;
    bool any = false;
    for (uint32_t oi = ph.first; oi < (uint32_t)ph.first + ph.count; ++oi) {
        const rq_tile_op& o = prog.ops[oi];
        if (o.kind != RQ_OP_DIAGP) continue;
        any = true;
        if ((tid >> 5) != o.t[3]) continue;                  // warp d serves the phase's d-th DIAGP op
        const rq_cplx* A = prog.pool + o.moff + 1;
        const uint32_t na = o.t[0], e = tid & 31u, first_bit = (e >> 4) * 4u;
        rq_cplx f = e < 16u ? gfac[o.t[2]] : rq_cplx{(rq_real)1, (rq_real)0};
#pragma unroll
        for (uint32_t b = 0; b < 4; ++b) {
            const uint32_t i = first_bit + b;
            if (i < na && ((e >> b) & 1u)) f = cmul(A[i], f);
        }
        tfac[o.t[3]][e] = f;
    }
    if (any) {                                               // (uniform: the phase's op list is the same for all threads)
        __syncthreads();
#pragma unroll
        for (int d = 0; d < RQ_PHASE_MAX_DIAGP; ++d) fA[d] = cmul(tfac[d][tid & 15u], tfac[d][16u + (tid >> 4)]);
    }
}

// ---- butterfly chains (rq_phase kind 2) -----------------------------------------------------------------------------
// A phase made of nothing but butterflies -- a radix-2^V pass of a QFT over the window bits -- without the per-op
// interpreter: the parameters of the <= V butterflies are fetched once, the group loop is load, butterflies, store.
template <int V, bool SWZ, typename Prog>
__device__ __forceinline__ void run_chain_phase(rq_cplx* sm, const Prog& prog, const rq_phase& ph, uint32_t T, uint32_t tid,
                                                const rq_cplx* gfac, rq_cplx (*tfac)[32], rq_cplx (*tall)[32]) {
    constexpr int D = 1 << V;
    static_assert(RQ_PHASE_MAX_DIAGP >= V, "a chain holds up to V butterflies");
    rq_cplx fA[RQ_PHASE_MAX_DIAGP];
    const uint32_t nb = ph.count >> 1;
    bool tabled = true;                                      // (uniform) every butterfly's ladder has a per-tile table
#pragma unroll
    for (int k = 0; k < V; ++k)
        if ((uint32_t)k < nb) tabled = tabled && prog.ops[ph.first + 2u * (uint32_t)k + 1u].t[2] < RQ_TFAC_SLOTS;
    if (tabled) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
            fA[k] = rq_cplx{(rq_real)1, (rq_real)0};
            if ((uint32_t)k < nb) {
                const uint32_t s = prog.ops[ph.first + 2u * (uint32_t)k + 1u].t[2];
                fA[k] = cmul(tall[s][tid & 15u], tall[s][16u + (tid >> 4)]);
            }
        }
    } else {
        phase_thread_factors(prog, ph, tid, gfac, tfac, fA);
    }
    // The parameters of butterfly k (table pointers, the Hadamard's two real entries, hub, variant) are warp-uniform and are
    // fetched where they are used, inside the group loop: held across the loop they cost ~33 vector registers, which at
    // 80 registers per thread (three 64 KB tiles per SM) went to local memory; the loop runs twice per phase, so
    // re-fetching them through the uniform datapath is cheaper than the spills.
    uint32_t stride[V];
#pragma unroll
    for (int b = 0; b < V; ++b) stride[b] = 1u << ph.w[b];
    const uint32_t ngroups = 1u << (T - V);
    // opening zeros at the window positions is bitwise: deposit(tid + 2^8 it) = deposit(tid) | deposit(2^8 it)
    auto deposit = [&](uint32_t x) {
#pragma unroll
        for (int b = 0; b < V; ++b) {
            const uint32_t p = ph.w[b];
            x = ((x >> p) << (p + 1)) | (x & ((1u << p) - 1u));
        }
        return x;
    };
    const uint32_t base_tid = deposit(tid);
    for (uint32_t g = tid, it = 0; g < ngroups; g += NT, ++it) {
        const uint32_t base = base_tid | deposit(it << 8);      // (the second term is warp-uniform)
        ramp a[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            uint32_t idx = base;
#pragma unroll
            for (int b = 0; b < V; ++b) if (j & (1 << b)) idx |= stride[b];
            a[j] = ramp_load(sm, sidx<SWZ>(idx));
        }
#pragma unroll
        for (int k = 0; k < V; ++k) {
            if ((uint32_t)k >= nb) break;
            const rq_tile_op& h = prog.ops[ph.first + 2u * (uint32_t)k];
            const rq_tile_op& o = prog.ops[ph.first + 2u * (uint32_t)k + 1u];
            const rq_cplx* H = prog.pool + h.moff;
            const rq_cplx* Bt = prog.pool + o.moff + 1 + o.t[0];
            const rq_cplx* Wt = Bt + (1u << o.t[1]);
            const rq_real c0 = H[0].x, c1 = H[1 * RQ_MSLOTS].x;
            const bool swp = (c0 < 0) != (H[2 * RQ_MSLOTS].x < 0);
            const rq_cplx ft = cmul(Bt[it], fA[k]);             // the phase's k-th DIAGP op has slot k (build_phases numbers them in order)
            const rq_cplx fc = rq_cplx{ft.x * c1, ft.y * c1};
            win_butterfly_any<V>(a, o.cm_in, Wt, fc, c0, swp, o.fuse == RQ_FUSE_BUTTERFLY_UP);
        }
#pragma unroll
        for (int j = 0; j < D; ++j) {
            uint32_t idx = base;
#pragma unroll
            for (int b = 0; b < V; ++b) if (j & (1 << b)) idx |= stride[b];
            ramp_store(sm, sidx<SWZ>(idx), a[j]);
        }
    }
}

template <int V, bool SWZ, typename Prog>
__device__ __forceinline__ void run_window_phase(rq_cplx* sm, const Prog& prog, const rq_phase& ph, uint32_t T, uint32_t tid,
                                                 uint64_t gbase, const rq_cplx* gfac, rq_cplx (*tfac)[32]) {
    constexpr int D = 1 << V;
    uint32_t stride[V];
#pragma unroll
    for (int b = 0; b < V; ++b) stride[b] = 1u << ph.w[b];
    const uint32_t ngroups = 1u << (T - V);
    // RQ_OP_DIAGP ops of the phase: tile factor x thread factor, once per phase (the group loop only adds the table
    // over the group-index bits above the thread's and the table over the window bits)
    rq_cplx fA[RQ_PHASE_MAX_DIAGP];
    phase_thread_factors(prog, ph, tid, gfac, tfac, fA);
    // the host only builds window phases for tiles of >= 2^(V+8) amplitudes: a warp-uniform trip count, so that op
    // headers, matrices and tables are fetched through the uniform datapath
#ifndef RQ_UNIFORM_LOOPS
    for (uint32_t g = tid, it = 0; g < ngroups; g += NT, ++it) {
#else
    for (uint32_t it = 0; it < (ngroups >> 8); ++it) {
        const uint32_t g = tid + (it << 8);
#endif
        uint32_t base = g;
#pragma unroll
        for (int b = 0; b < V; ++b) {
            const uint32_t p = ph.w[b];
            base = ((base >> p) << (p + 1)) | (base & ((1u << p) - 1u));
        }
        ramp a[D];
        uint32_t lidx[D];                                   // logical local index of every register slot
#pragma unroll
        for (int j = 0; j < D; ++j) {
            uint32_t idx = base;
#pragma unroll
            for (int b = 0; b < V; ++b) if (j & (1 << b)) idx |= stride[b];
            lidx[j] = idx;
            a[j] = ramp_load(sm, sidx<SWZ>(idx));
        }
        for (uint32_t oi = ph.first; oi < (uint32_t)ph.first + ph.count; ++oi) {
            const rq_tile_op& o = prog.ops[oi];
            if ((gbase & o.gcmask) != o.gcmask) continue;       // uniform per tile
#ifdef RQ_BRANCH_PRED
            if ((base & o.cm_out) != o.cm_out) continue;
            const bool on = true;
#else
            const bool on = (base & o.cm_out) == o.cm_out;       // per thread: selected, never branched on
#endif
            if (o.fuse == RQ_FUSE_SKIP) continue;               // a Hadamard the next op (a butterfly) carries out
            const rq_cplx* M = prog.pool + o.moff;
            if (o.kind == RQ_OP_DIAG) {
                const uint32_t lc = o.setmask, k = o.k;
                if (k == 0) {                                   // pure phase on "all controls set" (Z, S, T, CZ, CP, ...)
                    const rq_cplx ph = M[0];
#pragma unroll
                    for (int j = 0; j < D; ++j)
                        if ((lidx[j] & lc) == lc) a[j] = rmul(ph, a[j]);
                } else {
                    // table bit b comes from local position tb[b], or (non-resident) is the tile constant in selbase
                    uint32_t selbase = 0, tb[4] = {32, 32, 32, 32};
#pragma unroll
                    for (uint32_t b = 0; b < 4; ++b) {
                        if (b >= k) continue;
                        const uint32_t p = o.t[b];
                        if (p == 0xFF) selbase |= (uint32_t)((gbase >> o.gq[b]) & 1ull) << b; else tb[b] = p;
                    }
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        const uint32_t idx = lidx[j];
                        if ((idx & lc) != lc) continue;
                        uint32_t sel = selbase;
#pragma unroll
                        for (uint32_t b = 0; b < 4; ++b) sel |= (tb[b] < 32 ? ((idx >> tb[b]) & 1u) : 0u) << b;
                        a[j] = rmul(M[sel], a[j]);
                    }
                }
            } else if (o.kind == RQ_OP_DIAGP) {
                const uint32_t ci = o.cm_in;
                const rq_cplx* B = M + 1 + o.t[0];
                const rq_cplx* Wt = B + (1u << o.t[1]);
                rq_cplx fs = fA[0];
#pragma unroll
                for (int d = 1; d < RQ_PHASE_MAX_DIAGP; ++d) if (o.t[3] == d) fs = fA[d];
                const rq_cplx ft = cmul(B[it], fs);              // NT = 2^8: `it` = the group-index bits above the thread's
                if (o.fuse >= RQ_FUSE_BUTTERFLY) {               // (uniform) Hadamard-like op on the hub, then the ladder: one butterfly
                    const rq_cplx* H = prog.pool + prog.ops[oi - 1].moff;    // the skipped op's matrix, column-major: m00, m10, m01, m11
                    const rq_real c0 = H[0].x, c1 = H[1 * RQ_MSLOTS].x;
                    const bool swp = (c0 < 0) != (H[2 * RQ_MSLOTS].x < 0);
                    const rq_cplx fc = rq_cplx{ft.x * c1, ft.y * c1};
                    win_butterfly_any<V>(a, ci, Wt, fc, c0, swp, o.fuse == RQ_FUSE_BUTTERFLY_UP);
                    continue;
                }
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    if ((j & ci) != ci) continue;                // controls inside the window: uniform
                    a[j] = rsel(on, rmul(cmul(Wt[j], ft), a[j]), a[j]);
                }
            } else if (o.kind == RQ_OP_DENSE) {
                if (o.k == 1) win_dispatch1<V>(a, o, M, true, on);
                else win_dispatch2<V>(a, o, M, true, on);
            } else {
                if (o.k == 1) win_dispatch1<V>(a, o, M, false, on);
                else win_dispatch2<V>(a, o, M, false, on);
            }
        }
#pragma unroll
        for (int j = 0; j < D; ++j) ramp_store(sm, sidx<SWZ>(lidx[j]), a[j]);
    }
}

// MODE 0 (narrow): one op per shared-memory pass, dense ops of 1-2 qubits only, <= 64 registers so that 3+ tiles per SM
//                  are in flight -- the variant for one-gate and lightly fused sweeps (HBM-bound).
// MODE 1 (wide)  : also 3- and 4-qubit dense ops (16 amplitudes per thread in registers).
// MODE 2 (phased): register-window phases for heavily fused sweeps (compute-bound): several ops per smem round trip.
// SWZ: the tile is kept XOR-swizzled in shared memory (sv_internal.h) between a swizzle pass after the TMA load and
//      an unswizzle pass before the TMA store; chosen for sweeps with ops on the lowest local bits.
template <bool SWZ>
__device__ __forceinline__ void swizzle_pass(rq_cplx* sm, uint32_t T, uint32_t tid) {
    if (!SWZ) return;
    // phys(idx) = idx ^ ((idx >> B) & (2^B - 1)) is an involution inside every aligned group of 2^(2B) amplitudes: with
    // h = the group's upper B bits and l its lower B bits, (h, l) trades places with (h, l ^ h).  Only the pairs are
    // enumerated: h = 1 .. 2^B - 1, and of each pair the member whose bit `top set bit of h` is clear -- 2^(B-1) per h.
    constexpr uint32_t B = RQ_SWZ_BITS, HM = (1u << B) - 1u, PER = HM << (B - 1);     // swaps per group of 2^(2B)
    if (T < 2 * B) return;
    const uint32_t total = PER << (T - 2 * B);
    for (uint32_t k = tid; k < total; k += NT) {
        const uint32_t grp = k / PER, r = k - grp * PER;
        const uint32_t h = 1u + (r >> (B - 1)), m = r & ((1u << (B - 1)) - 1u);
        const uint32_t tb = 31u - (uint32_t)__clz(h);                                    // l has a 0 there: insert it into m
        const uint32_t l = ((m >> tb) << (tb + 1)) | (m & ((1u << tb) - 1u));
        const uint32_t i0 = (grp << (2 * B)) | (h << B) | l, i1 = i0 ^ h;
        const rq_cplx a = sm[i0], b = sm[i1];
        sm[i0] = b;
        sm[i1] = a;
    }
    __syncthreads();
}

template <typename Prog, int MODE, bool SWZ>
__global__ void __launch_bounds__(NT, MODE == 0 ? 4 : (MODE == 2 ? RQ_PHASED_MIN_BLOCKS : 1)) tile_sweep_kernel(rq_cplx* __restrict__ state, const __grid_constant__ Prog prog) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    rq_cplx* sm = reinterpret_cast<rq_cplx*>(smem_raw);
    __shared__ __align__(8) uint64_t bar_storage;
    __shared__ __align__(16) rq_cplx gfac[RQ_MAX_DIAGP];        // per-tile factors of the RQ_OP_DIAGP ops
    __shared__ __align__(16) rq_cplx tfac[MODE == 2 ? RQ_PHASE_MAX_DIAGP : 1][32];   // window phases: thread-factor tables of the phase's DIAGP ops
    __shared__ __align__(16) rq_cplx tall[MODE == 2 ? RQ_TFAC_SLOTS : 1][32];        // ... built once per tile for the sweep's first ops (butterfly chains)

    const uint32_t tid = threadIdx.x;
    const uint32_t T = prog.hdr.T, n = prog.hdr.n, rowbits = prog.hdr.rowbits;
    const uint32_t bar = smem_u32(&bar_storage);

    // tile -> (batch member, base index with zeros at the resident positions)
    const uint64_t tile = blockIdx.x;
    const uint64_t member = tile >> (n - T);
    uint64_t base = tile & ((1ull << (n - T)) - 1ull);
    for (uint32_t j = 0; j < T; ++j) {
        const uint32_t p = prog.hdr.res[j];
        base = ((base >> p) << (p + 1)) | (base & ((1ull << p) - 1ull));
    }
    rq_cplx* gtile = state + (member << n) + base;
    const uint64_t gbase = base | prog.hdr.high_base;
    // bit j = value of the j-th non-resident position, then the rank bits of a distributed state (RQ_OP_DIAGP)
    const uint64_t outer = (tile & ((1ull << (n - T)) - 1ull)) | ((prog.hdr.high_base >> n) << (n - T));

    const uint32_t nrows = 1u << (T - rowbits);
    const uint32_t rowbytes = (uint32_t)sizeof(rq_cplx) << rowbits;

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, rowbytes * nrows);
    }
    __syncthreads();
    for (uint32_t r = tid; r < nrows; r += NT) {
        uint64_t goff = 0;
        for (uint32_t i = 0; i < T - rowbits; ++i) goff |= (uint64_t)((r >> i) & 1u) << prog.hdr.res[rowbits + i];
        bulk_g2s(smem_u32(sm) + r * rowbytes, gtile + goff, rowbytes, bar);
    }
    if (prog.hdr.ndiagp) {                                       // while the tile is in flight
        diagp_tile_factors<MODE == 2>(prog, gfac, tall, tid, outer);
        __syncthreads();
    }
    mbar_wait(bar, 0);
    swizzle_pass<SWZ>(sm, T, tid);

    const rq_cplx* ext = reinterpret_cast<const rq_cplx*>(prog.hdr.ext_matrix);
    constexpr bool WIDE = MODE != 0;
    const uint32_t nsteps = MODE == 2 ? prog.hdr.nphases : prog.hdr.nops;
    for (uint32_t step = 0; step < nsteps; ++step) {
        uint32_t i = step;
        if (MODE == 2) {
            const rq_phase& ph = prog.phases[step];
            if (ph.kind == 2) {
                run_chain_phase<RQ_WINDOW_BITS, SWZ>(sm, prog, ph, T, tid, gfac, tfac, tall);
                __syncthreads();
                continue;
            }
            if (ph.kind == 1) {
                run_window_phase<RQ_WINDOW_BITS, SWZ>(sm, prog, ph, T, tid, gbase, gfac, tfac);
                __syncthreads();
                continue;
            }
            i = ph.first;
        }
        const rq_tile_op& o = prog.ops[i];
        if ((gbase & o.gcmask) != o.gcmask) continue;          // uniform per tile: no divergent barrier
        switch (o.kind) {
            case RQ_OP_DENSE:
                if (o.ext) {
                    switch (o.k) {
                        case 1: op_dense<1, true, SWZ>(sm, o, prog.pool, ext, T, tid); break;
                        case 2: op_dense<2, true, SWZ>(sm, o, prog.pool, ext, T, tid); break;
                        case 3: if (WIDE) op_dense<3, true, SWZ>(sm, o, prog.pool, ext, T, tid); break;
                        default: if (WIDE) op_dense<4, true, SWZ>(sm, o, prog.pool, ext, T, tid); break;
                    }
                } else {
                    switch (o.k) {
                        case 1: op_dense<1, false, SWZ>(sm, o, prog.pool, ext, T, tid); break;
                        case 2: op_dense<2, false, SWZ>(sm, o, prog.pool, ext, T, tid); break;
                        case 3: if (WIDE) op_dense<3, false, SWZ>(sm, o, prog.pool, ext, T, tid); break;
                        default: if (WIDE) op_dense<4, false, SWZ>(sm, o, prog.pool, ext, T, tid); break;
                    }
                }
                break;
            case RQ_OP_DIAG: op_diag<SWZ>(sm, o, prog.pool, T, tid, gbase); break;
            case RQ_OP_DIAGP: op_diagp<SWZ>(sm, o, prog.pool, T, tid, gfac); break;
            default: op_perm<SWZ>(sm, o, T, tid); break;
        }
        __syncthreads();
    }

    swizzle_pass<SWZ>(sm, T, tid);      // back to the linear layout the bulk stores expect
    fence_async_smem();       // generic-proxy writes to smem -> visible to the async (TMA) proxy
    __syncthreads();
    for (uint32_t r = tid; r < nrows; r += NT) {          // (sres: trailing swaps of resident qubits ride in the store addresses)
        uint64_t goff = 0;
        for (uint32_t i = 0; i < T - rowbits; ++i) goff |= (uint64_t)((r >> i) & 1u) << prog.hdr.sres[rowbits + i];
        bulk_s2g(gtile + goff, smem_u32(sm) + r * rowbytes, rowbytes);
    }
    bulk_commit_wait_read();  // smem must stay valid until the TMA engine has read it
}

template <typename Prog, bool SWZ>
int launch(rq_cplx* state, const Prog* prog, void* stream) {
    const size_t smem = sizeof(rq_cplx) << prog->hdr.T;
    const unsigned grid = (unsigned)prog->hdr.ntiles;
    bool wide = false;
    for (uint32_t i = 0; i < prog->hdr.nops; ++i) wide |= (prog->ops[i].kind == RQ_OP_DENSE && prog->ops[i].k > 2);
    const int mode = (prog->hdr.nphases > 0 && prog->hdr.max_phase_ops >= 2) ? 2 : (wide ? 1 : 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 2) tile_sweep_kernel<Prog, 2, SWZ><<<grid, NT, smem, st>>>(state, *prog);
    else if (mode == 1) tile_sweep_kernel<Prog, 1, SWZ><<<grid, NT, smem, st>>>(state, *prog);
    else tile_sweep_kernel<Prog, 0, SWZ><<<grid, NT, smem, st>>>(state, *prog);
    return (int)cudaGetLastError();
}

template <typename Prog, bool SWZ>
int configure() {
    const int bytes = (int)(sizeof(rq_cplx) << RQ_MAX_TILE_BITS);
    cudaError_t e = cudaFuncSetAttribute(tile_sweep_kernel<Prog, 0, SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tile_sweep_kernel<Prog, 1, SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tile_sweep_kernel<Prog, 2, SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return (int)e;
}

}  // namespace

