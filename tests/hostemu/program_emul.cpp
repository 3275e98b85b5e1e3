// tests/hostemu/program_emul.cpp -- TEST INFRASTRUCTURE ONLY (built and loaded by tests/test_program_emul_cpu.py).
//
// A host interpreter of the sweep programs that rocquantum_b200/csrc/host_ops.h emits for the tile-sweep kernel.  It
// lets the CPU suite check the whole host pipeline -- gate conversion, algebraic fusion, merging of controlled phases,
// sweep partition, phase building and the bit-level program encoding (fix / setmask / gcmask / window fields / pool
// layout) -- without a GPU, by following tile_sweep.cuh loop for loop (same group enumeration, same per-thread /
// per-iteration decomposition of RQ_OP_DIAGP).  It is not part of the product: nothing under rocquantum_b200/ links or
// loads it, and it is compiled from the product's headers so that it can never drift from what the engine launches.
//
// Arithmetic is done in double on the (rq_real-rounded) pool values.
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../../rocquantum_b200/csrc/gate_convert.h"
#include "../../rocquantum_b200/csrc/host_ops.h"

namespace {

typedef std::complex<double> cd;
constexpr uint32_t NT = RQ_TILE_THREADS;

inline cd pc(const rq_cplx& c) { return cd((double)c.x, (double)c.y); }
inline uint32_t sidx(bool swz, uint32_t idx) { return swz ? (idx ^ ((idx >> RQ_SWZ_BITS) & ((1u << RQ_SWZ_BITS) - 1u))) : idx; }

inline uint32_t spread(uint32_t g, const rq_tile_op& o) {
    for (uint32_t f = 0; f < o.nfix; ++f) {
        const uint32_t p = o.fix[f];
        g = ((g >> p) << (p + 1)) | (g & ((1u << p) - 1u));
    }
    return g;
}

// what run_chain_phase relies on: (skipped Hadamard-like op, butterfly) pairs only, at most V of them, the k-th butterfly's
// thread-factor slot is k, one window bit as hub, no predicate of any kind, and -- for the UP variant -- a window table
// that really does not depend on the window bits below the hub
template <typename Prog>
bool chain_is_wellformed(const Prog& prog, const rq_phase& ph) {
    if ((ph.count & 1u) || ph.count > 2u * ph.v) return false;
    for (uint32_t k = 0; k < ph.count / 2u; ++k) {
        const rq_tile_op& h = prog.ops[ph.first + 2u * k];
        const rq_tile_op& o = prog.ops[ph.first + 2u * k + 1u];
        if (h.kind != RQ_OP_DENSE || h.fuse != RQ_FUSE_SKIP || h.k != 1) return false;
        if (o.kind != RQ_OP_DIAGP || o.fuse < RQ_FUSE_BUTTERFLY || o.t[3] != k || o.cm_out || o.gcmask) return false;
        if (o.cm_in == 0 || (o.cm_in & (o.cm_in - 1u)) || o.cm_in != (1u << h.wt[0])) return false;
        if (o.fuse == RQ_FUSE_BUTTERFLY_UP) {
            const rq_cplx* Wt = prog.pool + o.moff + 1 + o.t[0] + (1u << o.t[1]);
            for (uint32_t j = 0; j < (1u << ph.v); ++j) {
                const uint32_t hi = j & ~(2u * o.cm_in - 1u);              // keep only the window bits above the hub
                if (pc(Wt[j | o.cm_in]) != pc(Wt[hi | o.cm_in])) return false;
            }
        }
    }
    return true;
}

template <typename Prog>
bool emulate_window_phase(std::vector<cd>& sm, const Prog& prog, const rq_phase& ph, uint32_t T, uint64_t gbase, const cd* gfac, bool swz) {
    const uint32_t V = ph.v, D = 1u << V;
    const uint32_t ngroups = 1u << (T - V);
    std::vector<cd> a(D);
    std::vector<uint32_t> lidx(D);
    for (uint32_t tid = 0; tid < NT; ++tid)
    for (uint32_t g = tid; g < ngroups; g += NT) {
        cd fA[RQ_PHASE_MAX_DIAGP];                                  // per thread, once per phase in the kernel
        for (uint32_t oi = ph.first; oi < (uint32_t)ph.first + ph.count; ++oi) {
            const rq_tile_op& o = prog.ops[oi];
            if (o.kind != RQ_OP_DIAGP) continue;
            if (o.t[3] >= RQ_PHASE_MAX_DIAGP || o.t[2] >= prog.hdr.ndiagp) return false;
            const rq_cplx* A = prog.pool + o.moff + 1;
            cd f = gfac[o.t[2]];
            for (uint32_t i = 0; i < o.t[0]; ++i) if ((tid >> i) & 1u) f *= pc(A[i]);
            fA[o.t[3]] = f;
        }
        uint32_t base = g;
        for (uint32_t b = 0; b < V; ++b) {
            const uint32_t p = ph.w[b];
            base = ((base >> p) << (p + 1)) | (base & ((1u << p) - 1u));
        }
        for (uint32_t j = 0; j < D; ++j) {
            uint32_t idx = base;
            for (uint32_t b = 0; b < V; ++b) if (j & (1u << b)) idx |= 1u << ph.w[b];
            lidx[j] = idx;
            a[j] = sm[sidx(swz, idx)];
        }
        for (uint32_t oi = ph.first; oi < (uint32_t)ph.first + ph.count; ++oi) {
            const rq_tile_op& o = prog.ops[oi];
            if ((gbase & o.gcmask) != o.gcmask) continue;
            if ((base & o.cm_out) != o.cm_out) continue;
            if (o.fuse == RQ_FUSE_SKIP) {                             // a Hadamard the next op (a butterfly) carries out
                if (o.kind != RQ_OP_DENSE || oi + 1 >= (uint32_t)ph.first + ph.count || prog.ops[oi + 1].fuse < RQ_FUSE_BUTTERFLY) return false;
                continue;
            }
            const rq_cplx* M = prog.pool + o.moff;
            const uint32_t ci = o.cm_in;
            if (o.kind == RQ_OP_DIAG) {
                const uint32_t lc = o.setmask, k = o.k;
                for (uint32_t j = 0; j < D; ++j) {
                    if ((lidx[j] & lc) != lc) continue;
                    uint32_t sel = 0;
                    for (uint32_t b = 0; b < k; ++b) {
                        const uint32_t p = o.t[b];
                        sel |= (p == 0xFF ? (uint32_t)((gbase >> o.gq[b]) & 1ull) : ((lidx[j] >> p) & 1u)) << b;
                    }
                    a[j] *= pc(M[sel]);
                }
            } else if (o.kind == RQ_OP_DIAGP) {
                const rq_cplx* B = M + 1 + o.t[0];
                const rq_cplx* Wt = B + (1u << o.t[1]);
                if ((g >> 8) >= (1u << o.t[1])) return false;
                const cd ft = pc(B[g >> 8]) * fA[o.t[3]];
                if (o.fuse >= RQ_FUSE_BUTTERFLY) {                    // as win_butterfly: Hadamard on the hub (window bit ci), then the ladder
                    if (ci == 0 || (ci & (ci - 1)) || oi == ph.first || prog.ops[oi - 1].fuse != RQ_FUSE_SKIP) return false;
                    const rq_cplx* H = prog.pool + prog.ops[oi - 1].moff;
                    const double c0 = H[0].x, c1 = H[1 * RQ_MSLOTS].x;
                    const bool swp = (c0 < 0) != (H[2 * RQ_MSLOTS].x < 0);
                    for (uint32_t j = 0; j < D; ++j) {
                        if (j & ci) continue;
                        const cd a0 = a[j], a1 = a[j | ci];
                        const cd sm_ = a0 + a1, df = a0 - a1;
                        a[j] = c0 * (swp ? df : sm_);
                        a[j | ci] = (pc(Wt[j | ci]) * (ft * c1)) * (swp ? sm_ : df);
                    }
                    continue;
                }
                for (uint32_t j = 0; j < D; ++j)
                    if ((j & ci) == ci) a[j] *= pc(Wt[j]) * ft;
            } else if (o.kind == RQ_OP_DENSE) {
                if (o.ext || o.k > 2) return false;
                if (o.k == 1) {
                    const uint32_t W = o.wt[0];
                    const cd m00 = pc(M[0 * RQ_MSLOTS]), m10 = pc(M[1 * RQ_MSLOTS]), m01 = pc(M[2 * RQ_MSLOTS]), m11 = pc(M[3 * RQ_MSLOTS]);
                    for (uint32_t j = 0; j < D; ++j) {
                        if (j & (1u << W)) continue;
                        if ((j & ci) != ci) continue;
                        const cd a0 = a[j], a1 = a[j | (1u << W)];
                        a[j] = m00 * a0 + m01 * a1;
                        a[j | (1u << W)] = m10 * a0 + m11 * a1;
                    }
                } else {
                    const uint32_t W0 = o.wt[0], W1 = o.wt[1];
                    if (!(W0 < W1)) return false;
                    for (uint32_t j = 0; j < D; ++j) {
                        if (j & ((1u << W0) | (1u << W1))) continue;
                        if ((j & ci) != ci) continue;
                        const uint32_t s[4] = {j, j | (1u << W0), j | (1u << W1), j | (1u << W0) | (1u << W1)};
                        const cd x[4] = {a[s[0]], a[s[1]], a[s[2]], a[s[3]]};
                        for (uint32_t i = 0; i < 4; ++i) {
                            cd acc(0.0, 0.0);
                            for (uint32_t c = 0; c < 4; ++c) acc += pc(M[(i + 4 * c) * RQ_MSLOTS]) * x[c];
                            a[s[i]] = acc;
                        }
                    }
                }
            } else if (o.kind == RQ_OP_PERM) {
                if (o.k == 1) {
                    const uint32_t W = o.wt[0];
                    for (uint32_t j = 0; j < D; ++j) {
                        if (j & (1u << W)) continue;
                        if ((j & ci) != ci) continue;
                        std::swap(a[j], a[j | (1u << W)]);
                    }
                } else {
                    const uint32_t W0 = o.wt[0], W1 = o.wt[1];
                    for (uint32_t j = 0; j < D; ++j) {
                        if (j & ((1u << W0) | (1u << W1))) continue;
                        if ((j & ci) != ci) continue;
                        std::swap(a[j | (1u << W0)], a[j | (1u << W1)]);
                    }
                }
            } else {
                return false;
            }
        }
        for (uint32_t j = 0; j < D; ++j) sm[sidx(swz, lidx[j])] = a[j];
    }
    return true;
}

template <typename Prog>
bool emulate_op(std::vector<cd>& sm, const Prog& prog, const rq_tile_op& o, uint32_t T, uint64_t gbase, const cd* gfac, bool swz) {
    const uint32_t ngroups = 1u << (T - o.nfix);
    if (o.kind == RQ_OP_DENSE) {
        if (o.ext) return false;
        const uint32_t K = o.k, D = 1u << K;
        std::vector<uint32_t> off(D);
        for (uint32_t j = 0; j < D; ++j) {
            uint32_t v = 0;
            for (uint32_t b = 0; b < K; ++b) if ((j >> b) & 1u) v |= 1u << o.t[b];
            off[j] = v;
        }
        const rq_cplx* M = prog.pool + o.moff;
        std::vector<cd> a(D);
        for (uint32_t g = 0; g < ngroups; ++g) {
            const uint32_t base = spread(g, o) | o.setmask;
            for (uint32_t j = 0; j < D; ++j) a[j] = sm[sidx(swz, base | off[j])];
            for (uint32_t i = 0; i < D; ++i) {
                cd acc(0.0, 0.0);
                for (uint32_t j = 0; j < D; ++j) acc += pc(M[(i + j * D) * RQ_MSLOTS]) * a[j];
                sm[sidx(swz, base | off[i])] = acc;
            }
        }
    } else if (o.kind == RQ_OP_DIAG) {
        uint32_t selbase = 0;
        for (uint32_t b = 0; b < o.k; ++b)
            if (o.t[b] == 0xFF) selbase |= (uint32_t)((gbase >> o.gq[b]) & 1ull) << b;
        const rq_cplx* Dg = prog.pool + o.moff;
        for (uint32_t g = 0; g < ngroups; ++g) {
            const uint32_t idx = spread(g, o) | o.setmask;
            uint32_t sel = selbase;
            for (uint32_t b = 0; b < o.k; ++b)
                if (o.t[b] != 0xFF) sel |= ((idx >> o.t[b]) & 1u) << b;
            sm[sidx(swz, idx)] *= pc(Dg[sel]);
        }
    } else if (o.kind == RQ_OP_DIAGP) {
        const rq_cplx* P = prog.pool + o.moff;
        const uint32_t na = o.t[0], nb = o.t[1];
        const rq_cplx* A = P + 1;
        const rq_cplx* B = A + na;
        if (o.t[2] >= prog.hdr.ndiagp || &prog.ops[prog.hdr.diagp_op[o.t[2]]] != &o || o.t[3] != 0xFF) return false;
        for (uint32_t tid = 0; tid < NT; ++tid) {                   // one "thread" at a time, as the kernel decomposes it
            cd f = gfac[o.t[2]];
            for (uint32_t i = 0; i < na; ++i)
                if ((tid >> i) & 1u) f *= pc(A[i]);
            for (uint32_t g = tid, m = 0; g < ngroups; g += NT, ++m) {
                if (m >= (1u << nb)) return false;                  // table overrun
                const uint32_t pi = sidx(swz, spread(g, o) | o.setmask);
                sm[pi] *= pc(B[m]) * f;
            }
        }
    } else if (o.kind == RQ_OP_PERM) {
        for (uint32_t g = 0; g < ngroups; ++g) {
            const uint32_t l0 = spread(g, o) | o.setmask;
            std::swap(sm[sidx(swz, l0)], sm[sidx(swz, l0 ^ o.xm)]);
        }
    } else {
        return false;
    }
    return true;
}

unsigned g_store_remaps = 0, g_perm_ops = 0;     // programs whose store map differs from the load map / permutation ops executed as passes

template <typename Prog>
bool emulate_program(const Prog& prog, cd* state) {
    const uint32_t T = prog.hdr.T, n = prog.hdr.n;
    for (uint32_t j = 0; j < T; ++j) if (prog.hdr.sres[j] != prog.hdr.res[j]) { ++g_store_remaps; break; }
    for (uint32_t i = 0; i < prog.hdr.nops; ++i) g_perm_ops += prog.ops[i].kind == RQ_OP_PERM;
    const bool swz = prog.hdr.swz != 0;
    // the launcher's choice of kernel variant (tile_sweep.cuh: launch)
    const bool phased = prog.hdr.nphases > 0 && prog.hdr.max_phase_ops >= 2;
    std::vector<cd> sm((size_t)1 << T);
    for (uint64_t tile = 0; tile < prog.hdr.ntiles; ++tile) {
        const uint64_t member = tile >> (n - T);
        uint64_t base = tile & ((1ull << (n - T)) - 1ull);
        for (uint32_t j = 0; j < T; ++j) {
            const uint32_t p = prog.hdr.res[j];
            base = ((base >> p) << (p + 1)) | (base & ((1ull << p) - 1ull));
        }
        cd* gtile = state + (member << n) + base;
        const uint64_t gbase = base | prog.hdr.high_base;
        const uint64_t outer = (tile & ((1ull << (n - T)) - 1ull)) | ((prog.hdr.high_base >> n) << (n - T));
        auto goff = [&](uint32_t l) {
            uint64_t o = 0;
            for (uint32_t j = 0; j < T; ++j) o |= (uint64_t)((l >> j) & 1u) << prog.hdr.res[j];
            return o;
        };
        for (uint32_t j = 0; j < prog.hdr.rowbits; ++j) if (prog.hdr.res[j] != j) return false;     // rows must be contiguous
        for (uint32_t l = 0; l < (1u << T); ++l) sm[sidx(swz, l)] = gtile[goff(l)];                 // bulk load + swizzle pass
        cd gfac[RQ_MAX_DIAGP];                                       // diagp_tile_factors
        for (uint32_t s = 0; s < prog.hdr.ndiagp; ++s) {
            const rq_tile_op& o = prog.ops[prog.hdr.diagp_op[s]];
            if (o.kind != RQ_OP_DIAGP) return false;
            const rq_cplx* P = prog.pool + o.moff;
            const uint32_t ng = o.k;
            const rq_cplx* G = P + o.xm;
            const uint8_t* gbit = reinterpret_cast<const uint8_t*>(G + ng);
            cd f = pc(P[0]);
            for (uint32_t j = 0; j < ng; ++j)
                if ((outer >> gbit[j]) & 1ull) f *= pc(G[j]);
            gfac[s] = f;
        }
        const uint32_t nsteps = phased ? prog.hdr.nphases : prog.hdr.nops;
        for (uint32_t step = 0; step < nsteps; ++step) {
            uint32_t i = step;
            if (phased) {
                const rq_phase& ph = prog.phases[step];
                if (ph.kind == 1 || ph.kind == 2) {                    // kind 2 (butterfly chain): the same arithmetic without the op interpreter
                    if (ph.kind == 2 && !chain_is_wellformed(prog, ph)) return false;
                    if (!emulate_window_phase(sm, prog, ph, T, gbase, gfac, swz)) return false;
                    continue;
                }
                i = ph.first;
            }
            const rq_tile_op& o = prog.ops[i];
            if ((gbase & o.gcmask) != o.gcmask) continue;
            if (!emulate_op(sm, prog, o, T, gbase, gfac, swz)) return false;
        }
        auto soff = [&](uint32_t l) {                                // the stores' own map (trailing swaps folded into the addresses)
            uint64_t o = 0;
            for (uint32_t j = 0; j < T; ++j) o |= (uint64_t)((l >> j) & 1u) << prog.hdr.sres[j];
            return o;
        };
        for (uint32_t j = 0; j < prog.hdr.rowbits; ++j) if (prog.hdr.sres[j] != j) return false;    // rows stay contiguous
        {   // sres must be a permutation of res
            uint64_t a = 0, b = 0;
            for (uint32_t j = 0; j < T; ++j) { a |= 1ull << prog.hdr.res[j]; b |= 1ull << prog.hdr.sres[j]; }
            if (a != b) return false;
        }
        for (uint32_t l = 0; l < (1u << T); ++l) gtile[soff(l)] = sm[sidx(swz, l)];
    }
    return true;
}

}  // namespace

extern "C" {

unsigned hostemu_precision_bytes(void) { return (unsigned)sizeof(rq_real); }
// since the last call: programs with swaps folded into their store addresses, permutation ops left in the programs
void hostemu_take_counters(unsigned* storeRemaps, unsigned* permOps) {
    if (storeRemaps) *storeRemaps = g_store_remaps;
    if (permOps) *permOps = g_perm_ops;
    g_store_remaps = 0;
    g_perm_ops = 0;
}

// state: 2 * 2^n doubles (re, im interleaved), updated in place.  rankBits/rank: emulate the slice of one rank of a
// distributed state (state then holds 2^(n - rankBits) amplitudes and ops may control / act diagonally on rank bits).
// flags bit 0: skip merge_diagonals / push_x_forward, bit 1: resolve rank bits first (specialize_for_rank),
// bit 2: the mixed plan with 6-qubit blocks (*numMerged then returns the number of blocks).  Returns 0, or a negative code naming the stage that failed.
int hostemu_run_circuit(unsigned n, unsigned tileBits, const rocsvxGateOp* ops, size_t numOps, double* state, unsigned rankBits,
                        unsigned rank, unsigned flags, unsigned* numSweeps, unsigned* numMerged) {
    std::vector<rq::HostOp> hops;
    if (rq::convert_ops(n, ops, numOps, hops) != ROCQ_STATUS_SUCCESS) return -1;
    for (const rq::HostOp& o : hops) if (o.targets.size() > 4) return -2;
    const unsigned nl = n - rankBits;
    const uint64_t gmask = rankBits ? (((1ull << rankBits) - 1ull) << nl) : 0ull;
    for (const rq::HostOp& o : hops) if (o.nondiag() & gmask) return -3;
    if ((flags & 2u) && rankBits) hops = rq::specialize_for_rank(hops, nl, (uint64_t)rank << nl);      // as the engine does on a slice
    if (!(flags & 1u) && hops.size() > 1) hops = rq::merge_diagonals(rq::push_x_forward(hops));
    std::vector<rq::HostOp> fused = hops.size() > 1 ? rq::fuse_algebraic(hops, nl, gmask) : hops;
    unsigned merged = 0;
    for (const rq::HostOp& o : fused) merged += o.kind == rq::HostOp::DIAGP;
    if (numMerged) *numMerged = merged;
    rq::PlanLimits L;
    if (tileBits >= 1 && tileBits <= RQ_MAX_TILE_BITS) L.tile_bits = tileBits;
    L.max_ops = sizeof(rq_program_large::ops) / sizeof(rq_tile_op);
    L.pool_cplx = sizeof(rq_program_large::pool) / sizeof(rq_cplx);
    L.never_resident = gmask;
    static thread_local rq_program_large P;
    cd* st = reinterpret_cast<cd*>(state);
    if (flags & 4u) {
        // the mixed plan of the complex64 engine: 6-qubit blocks (their ops folded into ONE 64x64 matrix exactly as
        // engine.cu does, applied here as a plain matrix product -- the tensor-core kernel's arithmetic is not modelled)
        // interleaved with ordinary sweeps
        rq::BlockLimits BL;
        BL.min_cost = 0.0;
        unsigned launches = 0, nblocks = 0;
        for (const rq::MixedStep& ms : rq::plan_mixed(fused, nl, L, BL)) {
            ++launches;
            if (!ms.block) {
                if (!rq::build_program(P, ms.sweep, fused, nl, 1, (uint64_t)rank << nl)) return -4;
                if (!emulate_program(P, st)) return -5;
                continue;
            }
            ++nblocks;
            std::vector<cd> U(64 * 64, cd(0.0, 0.0));
            for (unsigned c = 0; c < 64; ++c) U[c + 64u * c] = cd(1.0, 0.0);
            for (int idx : ms.ops) {
                if ((fused[idx].qubits() >> nl) != 0) return -6;      // a block may only hold local ops
                rq::apply_small_columns(fused[idx], ms.blk, U.data(), 64);
            }
            uint64_t bm = 0;
            for (unsigned p : ms.blk) bm |= 1ull << p;
            std::vector<cd> in(64), out(64);
            for (uint64_t base = 0; base < (1ull << nl); ++base) {
                if (base & bm) continue;
                for (unsigned j = 0; j < 64; ++j) {
                    uint64_t off = 0;
                    for (unsigned b = 0; b < 6; ++b) if ((j >> b) & 1u) off |= 1ull << ms.blk[b];
                    in[j] = st[base | off];
                }
                for (unsigned r = 0; r < 64; ++r) {
                    cd acc(0.0, 0.0);
                    for (unsigned j = 0; j < 64; ++j) acc += U[r + 64u * j] * in[j];
                    out[r] = acc;
                }
                for (unsigned j = 0; j < 64; ++j) {
                    uint64_t off = 0;
                    for (unsigned b = 0; b < 6; ++b) if ((j >> b) & 1u) off |= 1ull << ms.blk[b];
                    st[base | off] = out[j];
                }
            }
        }
        if (numSweeps) *numSweeps = launches;
        if (numMerged) *numMerged = nblocks;
        return 0;
    }
    const std::vector<rq::SweepPlan> plans = rq::plan_sweeps(fused, nl, L);
    if (numSweeps) *numSweeps = (unsigned)plans.size();
    for (const rq::SweepPlan& sp : plans) {
        if (!rq::build_program(P, sp, fused, nl, 1, (uint64_t)rank << nl)) return -4;
        if (!emulate_program(P, st)) return -5;
    }
    return 0;
}

// rq::BitGather (host_ops.h) as rocsvSample uses it: out[s] bit j = bit measured[j] of idx[s]
void hostemu_gather_bits(const uint64_t* idx, size_t shots, const unsigned* measured, unsigned nm, uint64_t* out) {
    const rq::BitGather gather(measured, nm);
    for (size_t s = 0; s < shots; ++s) out[s] = gather(idx[s]);
}

}  // extern "C"
