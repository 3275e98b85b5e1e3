// rocquantum_b200/csrc/host_ops.h -- host-side gate representation, algebraic fusion and the sweep
// planner.  Pure C++ (no CUDA): the same code plans on a machine without a GPU (rocsvxPlanCircuit).
//
// Role in the reference: this is what sits behind Circuit.flush() (python/rocq/api.py:74-89, "placeholder
// ... without fusion") and GateFusion::processQueue (rocquantum/src/hipStateVec/GateFusion.cpp:89-156,
// which only fuses one 1q gate before/after a CNOT).  Here every queue is (1) fused algebraically --
// runs of 1q gates and neighbouring 2q gates collapse into single 2x2 / 4x4 matrices -- and (2) cut into
// sweeps: maximal in-order groups of ops whose non-diagonal targets fit in one resident set of T qubits.
#pragma once
#include <algorithm>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "sv_internal.h"

namespace rq {

typedef std::complex<double> cd;

struct HostOp {
    enum Kind { DENSE = 1, DIAG = 2, PERM_X = 3, PERM_SWAP = 4, DIAGP = 5 };
    int kind = DENSE;
    std::vector<unsigned> targets;   // DENSE: matrix bit b <-> targets[b].  DIAG: table bit b.  PERM_X: {t}.  PERM_SWAP: {a,b}
                                     // DIAGP: any number of qubits, each contributing its own factor
    uint64_t cmask = 0;              // control qubits: op acts where all are 1
    std::vector<cd> data;            // DENSE: 2^k x 2^k column-major.  DIAG: 2^k entries
                                     // DIAGP (product of one-qubit diagonals under common controls, merge_diagonals):
                                     //   amp *= data[0] * prod over b with bit targets[b] set of data[1 + b]
    const void* ext = nullptr;       // DENSE only: matrix stays in device memory (eager rocsvApplyMatrix)
    bool dead = false;               // fuse_algebraic: absorbed into another op
    bool defer = false;              // plan_sweeps: leave this op to a later step (plan_mixed keeps dense ops for the tensor-core blocks)

    uint64_t tmask() const { uint64_t m = 0; for (unsigned t : targets) m |= 1ull << t; return m; }
    uint64_t qubits() const { return tmask() | cmask; }
    // qubits that must be resident in the tile: everything the op does not act on diagonally
    uint64_t nondiag() const { return (kind == DIAG || kind == DIAGP) ? 0ull : tmask(); }
    bool host_dense_uncontrolled() const { return kind == DENSE && cmask == 0 && ext == nullptr; }
    double cost() const {            // rough FMA-equivalents per amplitude, for the sweep budget
        const double frac = 1.0 / (double)(1ull << std::min(8, __builtin_popcountll(cmask)));
        if (kind == DENSE) return 4.0 * (double)(1u << targets.size()) * frac + 2.0;
        if (kind == DIAG) return 6.0 * frac + 1.0;
        if (kind == DIAGP) return 10.0 * frac + 2.0;
        return 2.0 * frac + 1.0;
    }
};

// ---- constructors for the reference's named gates (matrices as hipStateVec.cpp:276-427, 544-595) ----
inline HostOp make_dense1(unsigned t, cd m00, cd m01, cd m10, cd m11, uint64_t cmask = 0) {
    HostOp o; o.kind = HostOp::DENSE; o.targets = {t}; o.cmask = cmask; o.data = {m00, m10, m01, m11}; return o;   // column-major
}
inline HostOp make_phase(uint64_t qmask, cd ph) {                      // multiply by ph where all qubits of qmask are 1
    HostOp o; o.kind = HostOp::DIAG; o.cmask = qmask; o.data = {ph}; return o;
}
inline HostOp make_diag1(unsigned t, cd d0, cd d1, uint64_t cmask = 0) {
    HostOp o; o.kind = HostOp::DIAG; o.targets = {t}; o.cmask = cmask; o.data = {d0, d1}; return o;
}
inline HostOp make_x(unsigned t, uint64_t cmask = 0) { HostOp o; o.kind = HostOp::PERM_X; o.targets = {t}; o.cmask = cmask; return o; }
inline HostOp make_swap(unsigned a, unsigned b, uint64_t cmask = 0) { HostOp o; o.kind = HostOp::PERM_SWAP; o.targets = {a, b}; o.cmask = cmask; return o; }

// A diagonal table bit whose "0" half is all ones is a control: move it to cmask (halves the table and
// lets the kernel enumerate only the amplitudes that change).
inline void canonicalize_diag(HostOp& o) {
    if (o.kind != HostOp::DIAG) return;
    bool changed = true;
    while (changed && !o.targets.empty()) {
        changed = false;
        const unsigned k = (unsigned)o.targets.size();
        for (unsigned b = 0; b < k; ++b) {
            bool ctl = true;
            for (unsigned s = 0; s < (1u << k); ++s)
                if (!((s >> b) & 1u) && o.data[s] != cd(1.0, 0.0)) { ctl = false; break; }
            if (!ctl) continue;
            std::vector<cd> nd;
            for (unsigned s = 0; s < (1u << k); ++s)
                if ((s >> b) & 1u) nd.push_back(o.data[s]);
            o.cmask |= 1ull << o.targets[b];
            o.targets.erase(o.targets.begin() + b);
            o.data.swap(nd);
            changed = true;
            break;
        }
    }
}

// Turn a host matrix into the cheapest op kind: diagonal -> DIAG (+controls), else DENSE.
inline HostOp make_matrix(const std::vector<unsigned>& targets, uint64_t cmask, const std::vector<cd>& colmajor) {
    const unsigned k = (unsigned)targets.size(), D = 1u << k;
    bool diag = k <= 4;
    for (unsigned i = 0; i < D && diag; ++i)
        for (unsigned j = 0; j < D; ++j)
            if (i != j && colmajor[i + (size_t)j * D] != cd(0.0, 0.0)) { diag = false; break; }
    HostOp o; o.targets = targets; o.cmask = cmask;
    if (diag) {
        o.kind = HostOp::DIAG;
        o.data.resize(D);
        for (unsigned i = 0; i < D; ++i) o.data[i] = colmajor[i + (size_t)i * D];
        canonicalize_diag(o);
    } else {
        o.kind = HostOp::DENSE;
        o.data = colmajor;
    }
    return o;
}

// ---- tiny host simulator over an explicit qubit list, used only to build fused matrices ------------
// v has 2^m entries; bit i of its index <-> qs[i].  Applies op (whose qubits must all be in qs).
inline void apply_small(const HostOp& o, const std::vector<unsigned>& qs, std::vector<cd>& v) {
    const unsigned m = (unsigned)qs.size();
    auto pos = [&](unsigned q) { for (unsigned i = 0; i < m; ++i) if (qs[i] == q) return i; return 0u; };
    uint32_t cm = 0;
    for (unsigned q = 0; q < 64; ++q) if ((o.cmask >> q) & 1ull) cm |= 1u << pos(q);
    const unsigned k = (unsigned)o.targets.size();
    std::vector<unsigned> tp(k);
    uint32_t tm = 0;
    for (unsigned b = 0; b < k; ++b) { tp[b] = pos(o.targets[b]); tm |= 1u << tp[b]; }
    const uint32_t N = 1u << m;
    if (o.kind == HostOp::DIAG) {
        for (uint32_t i = 0; i < N; ++i) {
            if ((i & cm) != cm) continue;
            unsigned s = 0;
            for (unsigned b = 0; b < k; ++b) s |= ((i >> tp[b]) & 1u) << b;
            v[i] *= o.data[s];
        }
    } else if (o.kind == HostOp::DIAGP) {
        for (uint32_t i = 0; i < N; ++i) {
            if ((i & cm) != cm) continue;
            cd f = o.data[0];
            for (unsigned b = 0; b < k; ++b) if ((i >> tp[b]) & 1u) f *= o.data[1 + b];
            v[i] *= f;
        }
    } else if (o.kind == HostOp::PERM_X) {
        for (uint32_t i = 0; i < N; ++i)
            if ((i & cm) == cm && !(i & tm)) std::swap(v[i], v[i | tm]);
    } else if (o.kind == HostOp::PERM_SWAP) {
        const uint32_t a = 1u << tp[0], b = 1u << tp[1];
        for (uint32_t i = 0; i < N; ++i)
            if ((i & cm) == cm && (i & a) && !(i & b)) std::swap(v[i], v[(i ^ a) | b]);
    } else {
        const unsigned D = 1u << k;
        std::vector<cd> in(D);
        for (uint32_t i = 0; i < N; ++i) {
            if ((i & cm) != cm || (i & tm)) continue;
            std::vector<uint32_t> off(D);
            for (unsigned j = 0; j < D; ++j) { uint32_t f = 0; for (unsigned b = 0; b < k; ++b) if ((j >> b) & 1u) f |= 1u << tp[b]; off[j] = f; }
            for (unsigned j = 0; j < D; ++j) in[j] = v[i | off[j]];
            for (unsigned r = 0; r < D; ++r) {
                cd acc(0.0, 0.0);
                for (unsigned j = 0; j < D; ++j) acc += o.data[r + (size_t)j * D] * in[j];
                v[i | off[r]] = acc;
            }
        }
    }
}
// The same for every column of a column-major 2^m x ncols matrix at once (M <- op * M), without per-group allocations
// and with the complex products written out in real arithmetic: this is on the launch path of the tensor-core blocks.
inline void apply_small_columns(const HostOp& o, const std::vector<unsigned>& qs, cd* M, unsigned ncols) {
    const unsigned m = (unsigned)qs.size(), k = (unsigned)o.targets.size();
    if (o.kind != HostOp::DENSE || k > 4) {
        std::vector<cd> col(1u << m);
        for (unsigned c = 0; c < ncols; ++c) {
            std::copy(M + ((size_t)c << m), M + ((size_t)(c + 1) << m), col.begin());
            apply_small(o, qs, col);
            std::copy(col.begin(), col.end(), M + ((size_t)c << m));
        }
        return;
    }
    auto pos = [&](unsigned q) { for (unsigned i = 0; i < m; ++i) if (qs[i] == q) return i; return 0u; };
    uint32_t cm = 0, tm = 0, off[16];
    for (unsigned q = 0; q < 64; ++q) if ((o.cmask >> q) & 1ull) cm |= 1u << pos(q);
    unsigned tp[4];
    for (unsigned b = 0; b < k; ++b) { tp[b] = pos(o.targets[b]); tm |= 1u << tp[b]; }
    const unsigned D = 1u << k, N = 1u << m;
    for (unsigned j = 0; j < D; ++j) { uint32_t f = 0; for (unsigned b = 0; b < k; ++b) if ((j >> b) & 1u) f |= 1u << tp[b]; off[j] = f; }
    double ur[256], ui[256], xr[16], xi[16];
    for (unsigned e = 0; e < D * D; ++e) { ur[e] = o.data[e].real(); ui[e] = o.data[e].imag(); }
    for (unsigned c = 0; c < ncols; ++c) {
        cd* v = M + ((size_t)c << m);
        for (uint32_t i = 0; i < N; ++i) {
            if ((i & cm) != cm || (i & tm)) continue;
            for (unsigned j = 0; j < D; ++j) { xr[j] = v[i | off[j]].real(); xi[j] = v[i | off[j]].imag(); }
            for (unsigned r = 0; r < D; ++r) {
                double ar = 0.0, ai = 0.0;
                for (unsigned j = 0; j < D; ++j) {
                    const double a = ur[r + j * D], b = ui[r + j * D];
                    ar += a * xr[j] - b * xi[j];
                    ai += a * xi[j] + b * xr[j];
                }
                v[i | off[r]] = cd(ar, ai);
            }
        }
    }
}

// column-major matrix of op over the ordered qubit list qs
inline std::vector<cd> to_matrix(const HostOp& o, const std::vector<unsigned>& qs) {
    const unsigned D = 1u << qs.size();
    std::vector<cd> M((size_t)D * D);
    for (unsigned j = 0; j < D; ++j) {
        std::vector<cd> v(D, cd(0.0, 0.0));
        v[j] = cd(1.0, 0.0);
        apply_small(o, qs, v);
        for (unsigned i = 0; i < D; ++i) M[i + (size_t)j * D] = v[i];
    }
    return M;
}
inline std::vector<cd> matmul(const std::vector<cd>& A, const std::vector<cd>& B, unsigned D) {   // A*B, column-major
    std::vector<cd> C((size_t)D * D, cd(0.0, 0.0));
    for (unsigned j = 0; j < D; ++j)
        for (unsigned l = 0; l < D; ++l) {
            const cd b = B[l + (size_t)j * D];
            if (b == cd(0.0, 0.0)) continue;
            for (unsigned i = 0; i < D; ++i) C[i + (size_t)j * D] += A[i + (size_t)l * D] * b;
        }
    return C;
}



// ---- a distributed slice sees rank bits as constants ---------------------------------------------------------------
// On the rank whose index bits >= n_local read `high_base`, a control on a rank bit is either always satisfied (drop the
// control) or never (drop the op), and a diagonal factor on a rank bit is a constant.  Resolving them per rank BEFORE
// fusion and planning leaves purely local ops: CZ / CP across the slice boundary become one-qubit diagonals that fuse
// into their neighbours and into tensor-core blocks instead of forcing an ordinary sweep of their own.  (Non-diagonal
// targets never sit on rank bits: the distributed planner exchanges index bits first.)
inline std::vector<HostOp> specialize_for_rank(const std::vector<HostOp>& in, unsigned n_local, uint64_t high_base) {
    const uint64_t lmask = n_local >= 64 ? ~0ull : ((1ull << n_local) - 1ull);
    bool touched = false;
    for (const HostOp& o : in) if (o.qubits() & ~lmask) { touched = true; break; }
    if (!touched) return in;
    std::vector<HostOp> out;
    out.reserve(in.size());
    for (const HostOp& o : in) {
        if (!(o.qubits() & ~lmask)) { out.push_back(o); continue; }
        const uint64_t gc = o.cmask & ~lmask;
        if ((high_base & gc) != gc) continue;                       // a control on a rank bit that is 0 here: identity
        HostOp r = o;
        r.cmask &= lmask;
        if (o.kind == HostOp::DIAG) {
            std::vector<unsigned> keep;                              // table bits that stay
            uint32_t fixed_mask = 0, fixed_val = 0;
            for (unsigned b = 0; b < o.targets.size(); ++b) {
                if (o.targets[b] < n_local) { keep.push_back(b); continue; }
                fixed_mask |= 1u << b;
                if ((high_base >> o.targets[b]) & 1ull) fixed_val |= 1u << b;
            }
            if (fixed_mask) {
                r.targets.clear();
                for (unsigned b : keep) r.targets.push_back(o.targets[b]);
                r.data.assign((size_t)1 << keep.size(), cd(1.0, 0.0));
                for (unsigned s2 = 0; s2 < (1u << keep.size()); ++s2) {
                    uint32_t full = fixed_val;
                    for (unsigned i = 0; i < keep.size(); ++i) if ((s2 >> i) & 1u) full |= 1u << keep[i];
                    r.data[s2] = o.data[full];
                }
                canonicalize_diag(r);
            }
            if (r.targets.empty() && r.cmask == 0) {                 // a phase on the whole slice: keep it as a 1q diagonal
                if (r.data[0] == cd(1.0, 0.0)) continue;
                r.targets = {0u};
                r.data = {r.data[0], r.data[0]};
            }
        } else if (o.kind == HostOp::DIAGP) {
            r.targets.clear();
            r.data.assign(1, o.data[0]);
            for (unsigned b = 0; b < o.targets.size(); ++b) {
                if (o.targets[b] < n_local) { r.targets.push_back(o.targets[b]); r.data.push_back(o.data[1 + b]); }
                else if ((high_base >> o.targets[b]) & 1ull) r.data[0] *= o.data[1 + b];
            }
            if (r.cmask == 0 && r.targets.empty()) { r.kind = HostOp::DIAG; r.targets = {0u}; r.data = {r.data[0], r.data[0]}; }
        } else if (o.tmask() & ~lmask) {
            out.push_back(o);                                        // not expected; leave it to the planner to reject
            continue;
        }
        out.push_back(std::move(r));
    }
    return out;
}

// ---- pushing X gates forward through diagonal ops --------------------------------------------------------
// An uncontrolled X(q) followed by ops that act diagonally on q (controlled phases, Rz, controls of diagonal ops) and then
// by a dense gate on q forces q to be resident twice (and blocks every merged diagonal run that touches q until it has
// run).  Since  D X_q = X_q (X_q D X_q)  and X_q D X_q is again diagonal (the table with bit q flipped), the X can be
// carried forward and folded into the dense gate:  M X_q = M with its columns permuted.  Exact algebra, fewer ops.
inline void conjugate_diag_by_x(HostOp& o, unsigned q) {
    unsigned b = 0;
    const unsigned k = (unsigned)o.targets.size();
    for (; b < k; ++b) if (o.targets[b] == q) break;
    if (b == k) {                                         // q is a control: make it a table bit first ("0" half = 1)
        o.cmask &= ~(1ull << q);
        o.targets.push_back(q);
        std::vector<cd> nd((size_t)2 << k, cd(1.0, 0.0));
        for (unsigned s = 0; s < (1u << k); ++s) nd[s | (1u << k)] = o.data[s];
        o.data.swap(nd);
    }
    const unsigned K = (unsigned)o.targets.size();
    std::vector<cd> fl((size_t)1 << K);
    for (unsigned s = 0; s < (1u << K); ++s) fl[s] = o.data[s ^ (1u << b)];
    o.data.swap(fl);
    canonicalize_diag(o);
}
inline std::vector<HostOp> push_x_forward(const std::vector<HostOp>& in) {
    std::vector<HostOp> ops = in;
    bool changed = false;
    for (size_t p = 0; p < ops.size(); ++p) {
        if (ops[p].dead || ops[p].kind != HostOp::PERM_X || ops[p].cmask != 0) continue;
        const unsigned q = ops[p].targets[0];
        const uint64_t qb = 1ull << q;
        size_t stop = ops.size();
        bool ok = true, crossed = false;
        for (size_t j = p + 1; j < ops.size(); ++j) {
            const HostOp& o = ops[j];
            if (o.dead || !(o.qubits() & qb)) continue;
            if (o.kind == HostOp::DIAG && !o.ext && (((o.tmask() & qb) != 0) || o.targets.size() < 4)) { crossed = true; continue; }
            stop = j;
            break;
        }
        if (stop == ops.size() || !crossed) continue;
        HostOp& M = ops[stop];
        if (!(M.kind == HostOp::DENSE && M.cmask == 0 && !M.ext && (M.tmask() & qb) && M.targets.size() <= 4)) ok = false;
        if (!ok) continue;
        for (size_t j = p + 1; j < stop; ++j)
            if (!ops[j].dead && (ops[j].qubits() & qb)) conjugate_diag_by_x(ops[j], q);
        unsigned b = 0;
        while (M.targets[b] != q) ++b;
        const unsigned D = 1u << M.targets.size();
        std::vector<cd> nm((size_t)D * D);
        for (unsigned c = 0; c < D; ++c)
            for (unsigned r = 0; r < D; ++r) nm[r + (size_t)c * D] = M.data[r + (size_t)(c ^ (1u << b)) * D];      // M * X_q
        M.data.swap(nm);
        ops[p].dead = true;
        changed = true;
    }
    if (!changed) return in;
    std::vector<HostOp> out;
    out.reserve(ops.size());
    for (HostOp& o : ops) if (!o.dead) out.push_back(std::move(o));
    return out;
}

// ---- algebraic fusion -------------------------------------------------------------------------------
// Rule 1: an uncontrolled op on <= 2 qubits folds (left-multiplies) into the last op touching all of
//         its qubits when that op is an uncontrolled host DENSE whose targets contain them.
// Rule 2: an uncontrolled 2-qubit op absorbs (right-multiplies) the pending uncontrolled 1-qubit DENSE
//         ops sitting directly before it on its qubits; permutation / diagonal 2q gates (CNOT, CZ, SWAP)
//         are promoted to a dense 4x4 only when that lets them absorb such a neighbour
//         (the reference's GateFusion.cpp:95-147 pattern: post * CNOT * pre as one 4x4).
// `forbid`: positions that may never become targets of a DENSE op (rank bits of a distributed state); promotion of a
//         diagonal / controlled 2q gate to a dense 4x4 is skipped when it touches one of them.
inline std::vector<HostOp> fuse_algebraic(const std::vector<HostOp>& in, unsigned n, uint64_t forbid = 0) {
    std::vector<HostOp> out;
    out.reserve(in.size());
    std::vector<int> last(n > 64 ? n : 64, -1);
    for (const HostOp& op : in) {
        const uint64_t Q = op.qubits();
        const int nq = __builtin_popcountll(Q);
        std::vector<unsigned> qs;
        for (unsigned q = 0; q < 64; ++q) if ((Q >> q) & 1ull) qs.push_back(q);
        const bool small_host = op.ext == nullptr && nq >= 1 && nq <= 2 && op.targets.size() <= 2 &&
                                (op.kind != HostOp::DENSE || op.cmask == 0 || nq <= 2);
        bool absorbed = false;
        if (small_host) {
            // Rule 1
            const int j = last[qs[0]];
            bool same = j >= 0;
            for (unsigned q : qs) same = same && last[q] == j;
            if (same && out[j].host_dense_uncontrolled() && (out[j].tmask() & Q) == Q) {
                HostOp& G = out[j];
                const unsigned D = 1u << G.targets.size();
                // op expressed over G's target order (identity on G's other targets)
                std::vector<cd> U = to_matrix(op, G.targets);
                G.data = matmul(U, G.data, D);
                absorbed = true;
            }
        }
        if (!absorbed && small_host && nq == 1 && op.kind == HostOp::DENSE && op.cmask == 0) {
            // Rule 1b: the previous op on this qubit is a lone uncontrolled 1q diagonal / X: merge into one 2x2
            const int j = last[qs[0]];
            if (j >= 0 && !out[j].dead && out[j].ext == nullptr && out[j].qubits() == Q && out[j].kind != HostOp::DENSE) {
                HostOp G;
                G.kind = HostOp::DENSE;
                G.targets = qs;
                G.data = matmul(to_matrix(op, qs), to_matrix(out[j], qs), 2);
                out[j] = G;
                absorbed = true;
            }
        }
        if (!absorbed && small_host && nq == 2) {
            // Rule 2
            std::vector<int> pend;
            for (unsigned q : qs) {
                const int j = last[q];
                if (j >= 0 && !out[j].dead && out[j].host_dense_uncontrolled() && out[j].targets.size() == 1 &&
                    out[j].targets[0] == q)
                    pend.push_back(j);
            }
            const bool genuinely_dense = op.kind == HostOp::DENSE && op.cmask == 0;
            if ((genuinely_dense || !pend.empty()) && !(Q & forbid)) {
                HostOp G;
                G.kind = HostOp::DENSE;
                G.targets = (op.kind == HostOp::DENSE && op.cmask == 0) ? op.targets : qs;
                G.data = to_matrix(op, G.targets);
                for (int j : pend) {
                    G.data = matmul(G.data, to_matrix(out[j], G.targets), 4);
                    out[j].dead = true;
                }
                out.push_back(G);
                for (unsigned q : qs) last[q] = (int)out.size() - 1;
                absorbed = true;
            }
        }
        if (!absorbed) {
            out.push_back(op);
            for (unsigned q : qs) last[q] = (int)out.size() - 1;
        }
    }
    std::vector<HostOp> live;
    live.reserve(out.size());
    for (HostOp& o : out)
        if (!o.dead) live.push_back(std::move(o));
    return live;
}


// ---- merging runs of controlled phases -----------------------------------------------------------------
// A QFT-class circuit applies, after every H(i), the controlled phases CP(j, i) for all j > i: n(n-1)/2 diagonal ops, each
// of which would touch a quarter of the tile.  Diagonal ops sharing one control qubit h ("hub") multiply to
//     amp *= C * prod_{k : bit k set} f_k        wherever bit h is set,
// ONE pass over half of the tile (HostOp::DIAGP, device op RQ_OP_DIAGP).  Members: a pure phase on {h} or {h, k}
// (P, CZ, CP) and a one-qubit diagonal on k controlled by h (CRZ).  A later member may be moved back to the group's
// first member when nothing in between acts non-diagonally on its qubits (`nd_since`).  Groups of fewer than
// `min_members` stay as they were.
inline std::vector<HostOp> merge_diagonals(const std::vector<HostOp>& in, unsigned min_members = 3) {
    struct Group { uint64_t cand; uint64_t nd_since = 0; std::vector<int> members; bool open = true; };
    std::vector<Group> groups;
    std::vector<int> active;
    auto candidates = [](const HostOp& o) -> uint64_t {
        if (o.kind != HostOp::DIAG || o.ext) return 0ull;
        for (const cd& d : o.data) if (!(std::abs(d) > 0.5 && std::abs(d) < 2.0)) return 0ull;
        const int nc = __builtin_popcountll(o.cmask);
        if (o.targets.empty() && (nc == 1 || nc == 2)) return o.cmask;
        if (o.targets.size() == 1 && nc == 1) return o.cmask;
        return 0ull;
    };
    for (size_t i = 0; i < in.size(); ++i) {
        const HostOp& o = in[i];
        const uint64_t cand = candidates(o), Q = o.qubits();
        bool joined = false;
        if (cand) {
            for (int gi : active) {
                Group& g = groups[gi];
                if (!(Q & g.nd_since) && (cand & g.cand)) { g.cand &= cand; g.members.push_back((int)i); joined = true; break; }
            }
            if (!joined) {
                Group g;
                g.cand = cand;
                g.members.push_back((int)i);
                groups.push_back(std::move(g));
                active.push_back((int)groups.size() - 1);
                if (active.size() > 96) active.erase(active.begin());
            }
        }
        const uint64_t nd = o.nondiag();
        if (nd) {
            size_t w = 0;
            for (size_t a = 0; a < active.size(); ++a) {
                Group& g = groups[active[a]];
                g.nd_since |= nd;
                if ((g.cand & ~g.nd_since) != 0) active[w++] = active[a];      // a hub is still reachable
            }
            active.resize(w);
        }
    }
    std::vector<int> role(in.size(), -1);            // -1: keep, -2: dropped (merged), g >= 0: first member of group g
    bool any = false;
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        const Group& g = groups[gi];
        if (g.members.size() < min_members) continue;
        any = true;
        role[g.members[0]] = (int)gi;
        for (size_t m = 1; m < g.members.size(); ++m) role[g.members[m]] = -2;
    }
    if (!any) return in;
    std::vector<HostOp> out;
    out.reserve(in.size());
    for (size_t i = 0; i < in.size(); ++i) {
        if (role[i] == -1) { out.push_back(in[i]); continue; }
        if (role[i] == -2) continue;
        const Group& g = groups[role[i]];
        const unsigned hub = (unsigned)__builtin_ctzll(g.cand);
        HostOp P;
        P.kind = HostOp::DIAGP;
        P.cmask = 1ull << hub;
        P.data.push_back(cd(1.0, 0.0));
        auto factor = [&](unsigned q) -> cd& {
            for (size_t b = 0; b < P.targets.size(); ++b) if (P.targets[b] == q) return P.data[1 + b];
            P.targets.push_back(q);
            P.data.push_back(cd(1.0, 0.0));
            return P.data.back();
        };
        for (int mi : g.members) {
            const HostOp& m = in[mi];
            if (m.targets.empty()) {
                const uint64_t other = m.cmask & ~(1ull << hub);
                if (!other) P.data[0] *= m.data[0];
                else factor((unsigned)__builtin_ctzll(other)) *= m.data[0];
            } else {                                   // diag(d0, d1) on targets[0], controlled by the hub
                P.data[0] *= m.data[0];
                factor(m.targets[0]) *= m.data[1] / m.data[0];
            }
        }
        out.push_back(std::move(P));
    }
    return out;
}

// ---- sweep planner ----------------------------------------------------------------------------------
struct SweepPlan {
    std::vector<unsigned> res;       // ascending resident positions, size T
    std::vector<int> ops;            // indices into the op list, program order
    unsigned rowbits = 0;
};

struct PlanLimits {
    unsigned tile_bits = RQ_MAX_TILE_BITS;
    unsigned min_row_bits = RQ_MIN_ROW_BITS;
    unsigned max_ops = 160;
    unsigned pool_cplx = 1408;
    double budget = 1e30;            // max summed HostOp::cost() per sweep (first op always accepted)
    uint64_t never_resident = 0;     // positions that may not be resident (global qubits of a distributed state)
};

inline unsigned pool_need(const HostOp& o, unsigned T = RQ_MAX_TILE_BITS) {
    if (o.ext) return 0;
    if (o.kind == HostOp::DENSE) return RQ_MSLOTS * (1u << (2 * o.targets.size())) + 1u;     // +1: slot alignment
    if (o.kind == HostOp::DIAG) return 1u << o.targets.size();
    if (o.kind == HostOp::DIAGP) {                  // upper bound of build_program's layouts (pass of its own without resident control / window phase)
        const unsigned k = (unsigned)o.targets.size();
        return 1u + std::min(T, 8u) + (1u << (T > 8 ? T - 8 : 0)) + (1u << RQ_WINDOW_BITS) + k + (k + (unsigned)sizeof(rq_cplx) - 1u) / (unsigned)sizeof(rq_cplx);
    }
    return 0;
}

// Greedy in-order partition.  An op may join the current sweep when (a) no earlier deferred op conflicts
// with it on a qubit (diagonal/control action on a shared qubit commutes, so only non-diagonal overlap
// blocks) and (b) its non-diagonal targets fit in the resident set.  Diagonal factors and controls never
// need residency: the kernel resolves them from the tile base.
inline std::vector<SweepPlan> plan_sweeps_forward(const std::vector<HostOp>& ops, unsigned n, const PlanLimits& L) {
    std::vector<SweepPlan> plans;
    const unsigned T = std::min(n, L.tile_bits);
    std::vector<char> done(ops.size(), 0);
    size_t first = 0, remaining = ops.size();
    while (remaining > 0) {
        while (first < ops.size() && done[first]) ++first;
        uint64_t R = 0;
        const unsigned lowbits = std::min(std::min(L.min_row_bits, T), n);
        for (unsigned p = 0; p < lowbits; ++p) R |= 1ull << p;
        uint64_t blockedAny = 0, blockedND = 0;
        unsigned nops = 0, pool = 0, ndiagp = 0;
        double cost = 0.0;
        SweepPlan sp;
        size_t scanned = 0;
        for (size_t i = first; i < ops.size() && scanned < 8192; ++i) {
            if (done[i]) continue;
            ++scanned;
            const HostOp& o = ops[i];
            const uint64_t nd = o.nondiag(), dg = o.qubits() & ~nd;
            bool ok = !o.defer && !((nd & (blockedAny | blockedND)) || (dg & blockedAny));
            if (ok) {
                const uint64_t newR = R | nd;
                ok = __builtin_popcountll(newR) <= (int)T && !(nd & L.never_resident) && nops < L.max_ops &&
                     pool + pool_need(o, T) <= L.pool_cplx && (nops == 0 || cost + o.cost() <= L.budget) &&
                     (o.kind != HostOp::DENSE || o.targets.size() <= 4) && (o.kind != HostOp::DIAG || o.targets.size() <= 4) &&
                     !(o.ext && nops > 0) && !(o.kind == HostOp::DIAGP && ndiagp >= RQ_MAX_DIAGP);
                if (ok) {
                    R = newR;
                    ndiagp += o.kind == HostOp::DIAGP;
                    sp.ops.push_back((int)i);
                    done[i] = 1;
                    --remaining;
                    ++nops;
                    pool += pool_need(o, T);
                    cost += o.cost();
                    if (o.ext) break;            // a device-matrix op travels alone (one ext pointer per program)
                    continue;
                }
            }
            blockedAny |= nd;
            blockedND |= dg;
        }
        if (sp.ops.empty()) {                    // cannot happen for valid ops; avoid an endless loop
            sp.ops.push_back((int)first);
            done[first] = 1;
            --remaining;
            R |= ops[first].nondiag();
        }
        for (unsigned p = 0; p < n && __builtin_popcountll(R) < (int)T; ++p)
            if (!((L.never_resident >> p) & 1ull)) R |= 1ull << p;
        for (unsigned p = 0; p < 64; ++p) if ((R >> p) & 1ull) sp.res.push_back(p);
        sp.rowbits = 0;
        while (sp.rowbits < sp.res.size() && sp.res[sp.rowbits] == sp.rowbits) ++sp.rowbits;
        plans.push_back(std::move(sp));
    }
    return plans;
}

// The greedy partition depends on the direction it reads the circuit in: residents are granted to the ops that come first.
// A QFT ends in swaps (i, n-1-i); read forwards, every sweep has spent its resident slots on Hadamards long before the
// swaps are seen, and they need sweeps of their own; read BACKWARDS the swaps come first, each sweep seats both partners
// of a few swaps and then the Hadamards of exactly those qubits -- QFT-33: 5 sweeps instead of 6, and no sweep that only
// permutes.  Reversal is sound because the rule that lets an op overtake a deferred one (they commute) is symmetric: a
// valid sweep schedule of the reversed list, executed back to front, is a valid schedule of the list.  Both directions
// are planned; the backward plan is used only when it is strictly shorter.
// Within a run of consecutive ops on pairwise DISJOINT qubits any order is the same circuit.  The greedy planner seats ops
// in list order, so such a run is put in the order in which the ops that follow it need its qubits (stable): the ops a
// sweep seats first are then the ones whose successors can join them in the same sweep.
inline void order_disjoint_runs_by_next_use(std::vector<HostOp>& ops, std::vector<int>& orig) {
    const size_t N = ops.size();
    for (size_t i = 0; i < N;) {
        uint64_t mask = ops[i].qubits();
        size_t j = i + 1;
        while (j < N && !(ops[j].qubits() & mask)) { mask |= ops[j].qubits(); ++j; }
        if (j - i >= 2) {
            std::vector<std::pair<size_t, size_t>> key;                 // (next use, position in the run)
            for (size_t r = i; r < j; ++r) {
                size_t next = N;
                const uint64_t Q = ops[r].qubits();
                for (size_t k = j; k < N && k < j + 4096; ++k) if (ops[k].qubits() & Q) { next = k; break; }
                key.push_back({next, r});
            }
            std::stable_sort(key.begin(), key.end(), [](const std::pair<size_t, size_t>& a, const std::pair<size_t, size_t>& b) { return a.first < b.first; });
            std::vector<HostOp> run;
            std::vector<int> run_orig;
            for (const auto& kv : key) { run.push_back(std::move(ops[kv.second])); run_orig.push_back(orig[kv.second]); }
            for (size_t r = i; r < j; ++r) { ops[r] = std::move(run[r - i]); orig[r] = run_orig[r - i]; }
        }
        i = j;
    }
}

// Diagonal ops need no residency, so the diagonal ops a sweep BEGINS with can just as well end the sweep before it.  That
// matters for a ladder whose Hadamard closed the previous sweep: moved back, it meets its hub resident (one butterfly in a
// register window) instead of gating whole tiles of the next sweep.
inline void pull_leading_diagonals_back(std::vector<SweepPlan>& plans, const std::vector<HostOp>& ops, unsigned n, const PlanLimits& L) {
    const unsigned T = std::min(n, L.tile_bits);
    for (size_t k = 0; k + 1 < plans.size(); ++k) {
        SweepPlan& a = plans[k];
        SweepPlan& b = plans[k + 1];
        unsigned nops = (unsigned)a.ops.size(), pool = 0, ndiagp = 0;
        for (int i : a.ops) { pool += pool_need(ops[i], T); ndiagp += ops[i].kind == HostOp::DIAGP; if (ops[i].ext) nops = L.max_ops; }
        size_t moved = 0;
        while (moved < b.ops.size() && b.ops.size() - moved > 1) {                // never empty a sweep
            const HostOp& o = ops[b.ops[moved]];
            if (o.nondiag() || o.ext || o.defer || (o.kind != HostOp::DIAG && o.kind != HostOp::DIAGP)) break;
            if (o.kind == HostOp::DIAG && o.targets.size() > 4) break;
            if (nops + 1 > L.max_ops || pool + pool_need(o, T) > L.pool_cplx || (o.kind == HostOp::DIAGP && ndiagp >= RQ_MAX_DIAGP)) break;
            ++nops; pool += pool_need(o, T); ndiagp += o.kind == HostOp::DIAGP;
            a.ops.push_back(b.ops[moved]);
            ++moved;
        }
        b.ops.erase(b.ops.begin(), b.ops.begin() + (long)moved);
    }
}

inline std::vector<SweepPlan> plan_sweeps_unfixed(const std::vector<HostOp>& ops, unsigned n, const PlanLimits& L) {
    std::vector<SweepPlan> fwd = plan_sweeps_forward(ops, n, L);
    if (fwd.size() < 3) return fwd;
    for (const HostOp& o : ops) if (o.ext || o.defer) return fwd;
    std::vector<HostOp> rev(ops.rbegin(), ops.rend());
    std::vector<int> orig(ops.size());
    for (size_t k = 0; k < ops.size(); ++k) orig[k] = (int)(ops.size() - 1 - k);
    order_disjoint_runs_by_next_use(rev, orig);
    std::vector<SweepPlan> bwd = plan_sweeps_forward(rev, n, L);
    if (bwd.size() >= fwd.size()) return fwd;
    std::reverse(bwd.begin(), bwd.end());
    for (SweepPlan& sp : bwd) {
        for (int& k : sp.ops) k = orig[k];
        std::sort(sp.ops.begin(), sp.ops.end());                        // program order inside a sweep
    }
    return bwd;
}
inline std::vector<SweepPlan> plan_sweeps(const std::vector<HostOp>& ops, unsigned n, const PlanLimits& L) {
    std::vector<SweepPlan> plans = plan_sweeps_unfixed(ops, n, L);
    if (L.budget >= 1e29) pull_leading_diagonals_back(plans, ops, n, L);     // (a cost budget per sweep is the caller's to keep)
    return plans;
}

// ---- tile geometry of the tensor-core block sweep (block_sweep.cu) -------------------------------------------------------
// A tile = the 6 block bits + the 7 lowest other index bits ("columns").  It is moved by ONE tensor-map copy whose box lists,
// in shared-memory order: index bits 0-3 (a 128-byte row), the remaining column bits, the remaining block bits; index bits
// that are consecutive in that list AND in the state share a dimension.  Bits outside the tile form one box-1 dimension per
// run (addressed by tile-index bits), the last of which also carries the batch member.  At most five dimensions.
struct BlockLayout {
    unsigned rank = 0;
    uint64_t dims[5], strides[5];    // elements (8-byte amplitudes), bytes
    uint32_t box[5];
    uint8_t tbits[5];                // tile-index bits a dimension consumes (0: in the tile, 255: all that remain)
    uint8_t col[7], lp_col[7], lp_blk[6];
    uint8_t res[13];                 // tile bits, ascending
};
// layout for a given choice of the seven column bits (colmask: 7 non-block positions containing every non-block bit of 0-3)
inline bool block_layout_with_columns(uint64_t blockmask, uint64_t colmask, unsigned n, size_t batch, BlockLayout& L) {
    unsigned blk[6], nb = 0, nc = 0;
    for (unsigned p = 0; p < n; ++p) {
        if ((blockmask >> p) & 1ull) blk[nb++] = p;
        else if ((colmask >> p) & 1ull) L.col[nc++] = (uint8_t)p;
    }
    if (nb != 6 || nc != 7) return false;
    const uint64_t tile = blockmask | colmask;
    if ((tile & 0xFull) != 0xFull) return false;                            // a 128-byte row (index bits 0-3) is always inside the tile
    for (unsigned p = 0, r = 0; p < n; ++p) if ((tile >> p) & 1ull) L.res[r++] = (uint8_t)p;
    // shared-memory order of the tile bits
    unsigned order[13], no = 0;
    for (unsigned p = 0; p < 4; ++p) order[no++] = p;                       // bits 0-3 are always in the tile
    for (unsigned c = 0; c < 7; ++c) if (L.col[c] >= 4) order[no++] = L.col[c];
    // block bits: the run that continues the last column bit first (it then shares that column run's dimension), the rest
    // ascending -- the kernel addresses block bits through lp_blk[], so their order in shared memory is free
    uint64_t placed = 0;
    if (no > 4)
        for (unsigned p = order[no - 1] + 1; p < n && ((blockmask >> p) & 1ull); ++p) { order[no++] = p; placed |= 1ull << p; }
    for (unsigned b = 0; b < 6; ++b) if (blk[b] >= 4 && !((placed >> blk[b]) & 1ull)) order[no++] = blk[b];
    for (unsigned l = 0; l < 13; ++l) {
        const unsigned p = order[l];
        bool is_blk = false;
        for (unsigned b = 0; b < 6; ++b) if (blk[b] == p) { L.lp_blk[b] = (uint8_t)l; is_blk = true; }
        if (!is_blk) for (unsigned c = 0; c < 7; ++c) if (L.col[c] == p) L.lp_col[c] = (uint8_t)l;
    }
    unsigned rank = 0;
    auto add = [&](unsigned start, unsigned len, bool in_tile) {
        if (rank >= 5) { rank = 99; return; }
        L.dims[rank] = 1ull << len;
        L.strides[rank] = 8ull << start;
        L.box[rank] = in_tile ? (1u << len) : 1u;
        L.tbits[rank] = in_tile ? 0 : (uint8_t)len;
        ++rank;
    };
    add(0, 4, true);
    for (unsigned l = 4; l < 13 && rank <= 5;) {                            // runs of the ordered list, at most 8 bits each (box <= 256)
        unsigned m = l + 1;
        while (m < 13 && order[m] == order[m - 1] + 1 && m - l < 8) ++m;
        add(order[l], m - l, true);
        l = m;
    }
    int last_free = -1;
    for (unsigned p = 4; p < n && rank <= 5;) {                             // runs of bits outside the tile, ascending
        if ((tile >> p) & 1ull) { ++p; continue; }
        unsigned q = p;
        while (q < n && !((tile >> q) & 1ull)) ++q;
        add(p, q - p, false);
        last_free = (int)rank - 1;
        p = q;
    }
    if (rank > 5) return false;
    if (last_free >= 0 && L.strides[last_free] * L.dims[last_free] == (8ull << n)) L.dims[last_free] *= batch;   // top run: absorbs the batch
    else if (batch > 1 || last_free < 0) {
        if (rank >= 5) return false;
        L.dims[rank] = batch;
        L.strides[rank] = 8ull << n;
        L.box[rank] = 1;
        last_free = (int)rank++;
    }
    L.tbits[last_free] = 255;
    for (unsigned d = 0; d < rank; ++d) if (L.dims[d] > 0xffffffffull || L.box[d] > 256) return false;
    L.rank = rank;
    return true;
}
// The column bits are free to choose (any seven positions outside the block, as long as index bits 0-3 are in the tile).
// The lowest seven give the longest contiguous pieces; when the block is two separate runs of index bits (neighbouring
// logical qubits that an index-bit exchange left far apart in a distributed slice) that choice needs a sixth tensor-map
// dimension, and columns that extend a block run downwards or upwards instead keep the tile within five.
inline bool block_layout(uint64_t blockmask, unsigned n, size_t batch, BlockLayout& L) {
    if (n < 13 || n > 40 || __builtin_popcountll(blockmask) != 6 || (blockmask >> n)) return false;
    const uint64_t must = 0xFull & ~blockmask;                              // non-block bits of the 128-byte row
    const unsigned extra = 7u - (unsigned)__builtin_popcountll(must);
    auto take = [&](unsigned from, int dir, unsigned count, uint64_t have) {   // `count` free positions >= 4 walking from `from` in direction dir
        uint64_t m = 0;
        for (int p = (int)from; p >= 4 && p < (int)n && (unsigned)__builtin_popcountll(m) < count; p += dir)
            if (!((blockmask >> p) & 1ull) && !((have >> p) & 1ull)) m |= 1ull << p;
        return m;
    };
    uint64_t cands[16];
    unsigned nc = 0;
    cands[nc++] = must | take(4, +1, extra, 0);                             // the lowest free positions (what round 1 always took)
    for (unsigned p = 4; p < n; ++p) {                                      // around every run of block bits above the row
        if (!((blockmask >> p) & 1ull) || (p > 4 && ((blockmask >> (p - 1)) & 1ull))) continue;
        unsigned q = p;
        while (q + 1 < n && ((blockmask >> (q + 1)) & 1ull)) ++q;            // run [p, q]
        if (nc + 3 > 16) break;
        const uint64_t below = take(p - 1, -1, extra, 0);
        cands[nc++] = must | below | take(q + 1, +1, extra - (unsigned)__builtin_popcountll(below), below);
        const uint64_t above = take(q + 1, +1, extra, 0);
        cands[nc++] = must | above | take(p - 1, -1, extra - (unsigned)__builtin_popcountll(above), above);
        const uint64_t half = take(p - 1, -1, extra / 2, 0);
        cands[nc++] = must | half | take(q + 1, +1, extra - (unsigned)__builtin_popcountll(half), half);
    }
    for (unsigned c = 0; c < nc; ++c) {
        if (__builtin_popcountll(cands[c]) != 7) continue;
        if (block_layout_with_columns(blockmask, cands[c], n, batch, L)) return true;
    }
    return false;
}
inline bool block_supported(uint64_t blockmask, unsigned n, size_t batch) {
    BlockLayout L;
    return block_layout(blockmask, n, batch, L);
}

// ---- mixed plan: tensor-core blocks + ordinary sweeps ----------------------------------------------------------------
// A step is either one 6-qubit block (every op of `ops` folded into one 64x64 unitary, applied by block_sweep.cu in one HBM
// pass) or one ordinary tile sweep.  Oldest first, which keeps the frontier of the circuit flat so that blocks stay full:
// take the first op not yet executed; if it can live in a block, try several 6-qubit sets around it (grown along the
// interactions of the ops that follow, biased to lower / higher qubits), fold into each every op that is free to run
// inside it, and keep the set that absorbs the most arithmetic.  If that is below `min_cost` (FMA-equivalents per
// amplitude, HostOp::cost), or the op cannot live in a block, run one ordinary sweep; that sweep leaves the dense ops a
// block could take to the blocks.
struct BlockLimits {
    unsigned qubits = 6;
    unsigned min_pos = 0;            // block positions must be >= min_pos
    double min_cost = 50.0;          // ~ three dense two-qubit matrices: below that the CUDA-core sweep is cheaper
    size_t batch = 1;
    bool (*supported)(uint64_t blockmask, unsigned n, size_t batch) = block_supported;   // can the kernel move this tile?
};
struct MixedStep {
    bool block = false;
    std::vector<unsigned> blk;       // block: six ascending positions
    std::vector<int> ops;            // block: indices folded into it, program order
    SweepPlan sweep;                 // otherwise
};


// policy 0: the candidate set that absorbs the most arithmetic, always.
// policy 1: the same, except while the oldest op sits on qubits NO earlier step has touched (the start of a circuit, or of
//           an untouched part of the register): there the candidate that overlaps the already-touched qubits least wins.
//           Greedy absorption starts a brick circuit as a staircase of overlapping blocks (0-5, 4-9, 8-13, ... with 7 gates
//           each) that takes ~25 blocks to flatten out; disjoint first blocks (0-5, 6-11, ...: 6 gates each) put the very
//           next row on full 9-gate diamonds.  configs[1]: 63 instead of 70 passes.
inline std::vector<MixedStep> plan_mixed_policy(const std::vector<HostOp>& ops, unsigned n, const PlanLimits& L, const BlockLimits& BL, int policy) {
    std::vector<MixedStep> steps;
    std::vector<char> done(ops.size(), 0);
    size_t remaining = ops.size(), first = 0;
    uint64_t touched = 0;                                      // qubits of every op executed so far
    const uint64_t low = BL.min_pos >= 64 ? ~0ull : ((1ull << BL.min_pos) - 1ull);
    const uint64_t all = n >= 64 ? ~0ull : ((1ull << n) - 1ull);
    auto eligible = [&](const HostOp& o) {
        const uint64_t Q = o.qubits();
        return !o.ext && !(Q & low) && (unsigned)__builtin_popcountll(Q) <= BL.qubits && !(Q & L.never_resident) && !(Q & ~all);
    };
    const bool possible = n >= 13 && n >= BL.min_pos + BL.qubits;
    // every op that is free to run (nothing deferred conflicts with it) and lies inside B, in program order
    auto fold = [&](uint64_t B, std::vector<int>& pick) {
        double cost = 0.0;
        uint64_t bAny = 0, bND = 0;
        size_t scanned = 0;
        for (size_t i = first; i < ops.size() && scanned < 4096; ++i) {
            if (done[i]) continue;
            ++scanned;
            const HostOp& o = ops[i];
            const uint64_t Q = o.qubits(), nd = o.nondiag(), dg = Q & ~nd;
            if (!(Q & ~B) && !o.ext && !((nd & (bAny | bND)) || (dg & bAny))) {
                pick.push_back((int)i);
                cost += o.cost();
            } else {
                bAny |= nd;
                bND |= dg;
            }
        }
        return cost;
    };
    // grow a qubit set from the seed along the ops that follow; bias 0: in program order, 1: lowest neighbour first,
    // 2: highest first, 3 / 4: alternating starting low / high
    auto grow = [&](size_t seed, int bias) {
        uint64_t B = ops[seed].qubits();
        for (unsigned round = 0; (unsigned)__builtin_popcountll(B) < BL.qubits && round < BL.qubits; ++round) {
            uint64_t nb = 0;                                   // qubits that an upcoming eligible op connects to B
            size_t scanned = 0;
            for (size_t i = seed; i < ops.size() && scanned < 512; ++i) {
                if (done[i]) continue;
                ++scanned;
                const uint64_t Q = ops[i].qubits();
                if ((Q & B) && (Q & ~B) && eligible(ops[i])) {
                    nb |= Q & ~B;
                    if (bias == 0) break;
                }
            }
            if (!nb) break;
            const bool take_low = bias == 1 || (bias == 3 && !(round & 1)) || (bias == 4 && (round & 1));
            const unsigned q = (bias == 0 || take_low) ? (unsigned)__builtin_ctzll(nb) : 63u - (unsigned)__builtin_clzll(nb);
            B |= 1ull << q;
        }
        // pad with neighbouring positions (keeps the tile a short list of contiguous index-bit runs)
        while ((unsigned)__builtin_popcountll(B) < BL.qubits) {
            const unsigned hi = 63u - (unsigned)__builtin_clzll(B), lo = (unsigned)__builtin_ctzll(B);
            bool added = false;
            auto ok = [&](unsigned p) { return p < n && p >= BL.min_pos && !((B >> p) & 1ull) && !((L.never_resident >> p) & 1ull); };
            for (unsigned p = lo; p <= hi && !added; ++p) if (ok(p)) { B |= 1ull << p; added = true; }
            if (!added && ok(hi + 1)) { B |= 1ull << (hi + 1); added = true; }
            if (!added && lo > 0 && ok(lo - 1)) { B |= 1ull << (lo - 1); added = true; }
            if (!added) break;
        }
        return B;
    };
    while (remaining > 0) {
        while (first < ops.size() && done[first]) ++first;
        std::vector<int> best;
        uint64_t bestB = 0;
        double best_cost = -1.0;
        const bool spread = policy == 1 && !(ops[first].qubits() & touched);
        unsigned best_overlap = ~0u;
        if (possible && eligible(ops[first])) {
            uint64_t tried[5];
            unsigned ntried = 0;
            for (int bias = 0; bias < 5; ++bias) {
                const uint64_t B = grow(first, bias);
                if ((unsigned)__builtin_popcountll(B) != BL.qubits) continue;
                bool dup = false;
                for (unsigned t = 0; t < ntried; ++t) dup |= tried[t] == B;
                if (dup) continue;
                tried[ntried++] = B;
                if (BL.supported && !BL.supported(B, n, BL.batch)) continue;
                std::vector<int> pick;
                const double cost = fold(B, pick);
                if (spread) {                                  // least overlap first, among the sets a block is worth launching for
                    if (cost < BL.min_cost) continue;
                    const unsigned ov = (unsigned)__builtin_popcountll(B & touched);
                    if (ov < best_overlap || (ov == best_overlap && cost > best_cost)) { best_overlap = ov; best_cost = cost; best.swap(pick); bestB = B; }
                } else if (cost > best_cost) { best_cost = cost; best.swap(pick); bestB = B; }
            }
        }
        // The oldest op's neighbourhood is not worth a block (the sparse tip of a dependency cone, the tail of a distributed
        // RUN step): before an ordinary sweep swallows everything that is free -- dozens of dense gates at ~0.27 block passes
        // each -- look for a worthwhile block around the other free ops of the frontier.
        if (possible && best_cost < BL.min_cost) {
            uint64_t bAny = 0, bND = 0;
            unsigned seeds = 0;
            size_t scanned = 0;
            for (size_t i = first; i < ops.size() && scanned < 1024 && seeds < 48; ++i) {
                if (done[i]) continue;
                ++scanned;
                const HostOp& o = ops[i];
                const uint64_t nd = o.nondiag(), dg = o.qubits() & ~nd;
                const bool free_ = !((nd & (bAny | bND)) || (dg & bAny));
                bAny |= nd;                                    // (whatever is not taken now blocks what follows it)
                bND |= dg;
                if (!free_ || i == first || !eligible(o) || o.kind != HostOp::DENSE) continue;
                ++seeds;
                for (int bias = 0; bias < 3; ++bias) {
                    const uint64_t B = grow(i, bias);
                    if ((unsigned)__builtin_popcountll(B) != BL.qubits) continue;
                    if (BL.supported && !BL.supported(B, n, BL.batch)) continue;
                    std::vector<int> pick;
                    const double cost = fold(B, pick);
                    if (cost >= BL.min_cost && cost > best_cost) { best_cost = cost; best.swap(pick); bestB = B; }
                }
            }
        }
        if (best_cost >= BL.min_cost && !best.empty()) {
            MixedStep st;
            st.block = true;
            for (unsigned p = 0; p < n; ++p) if ((bestB >> p) & 1ull) st.blk.push_back(p);
            st.ops = best;
            for (int i : best) { done[i] = 1; --remaining; touched |= ops[i].qubits(); }
            steps.push_back(std::move(st));
            continue;
        }
        if (getenv("ROCQ_PLAN_DEBUG")) {
            unsigned unsupported = 0, tried_n = 0;
            if (possible && eligible(ops[first]))
                for (int bias = 0; bias < 5; ++bias) {
                    const uint64_t B = grow(first, bias);
                    ++tried_n;
                    if ((unsigned)__builtin_popcountll(B) == BL.qubits && BL.supported && !BL.supported(B, n, BL.batch)) ++unsupported;
                }
            fprintf(stderr, "[plan] sweep fallback: oldest op kind %d qubits %llx eligible %d best_cost %.0f candidates unsupported %u/%u\n", ops[first].kind,
                    (unsigned long long)ops[first].qubits(), (int)(possible && eligible(ops[first])), best_cost, unsupported, tried_n);
        }
        // One ordinary sweep over what is left.  When blocks are possible it leaves the dense ops a block could take alone
        // (marked `defer` in the copy), unless the oldest op is such an op itself.
        std::vector<HostOp> rest;
        std::vector<int> back;
        const bool keep_for_blocks = possible && !eligible(ops[first]);
        for (size_t i = first; i < ops.size(); ++i)
            if (!done[i]) {
                rest.push_back(ops[i]);
                rest.back().defer = keep_for_blocks && ops[i].kind == HostOp::DENSE && eligible(ops[i]);
                back.push_back((int)i);
            }
        std::vector<SweepPlan> plans = plan_sweeps_forward(rest, n, L);   // only the first sweep is used: forward order
        MixedStep st;
        st.sweep = std::move(plans[0]);
        for (int& k : st.sweep.ops) { k = back[k]; done[k] = 1; --remaining; touched |= ops[k].qubits(); }
        steps.push_back(std::move(st));
    }
    return steps;
}

// Estimated duration of a plan in block passes.  A block pass is HBM-bound whatever it carries (1.0); an ordinary sweep is
// HBM-bound (0.75 of a block pass) until its arithmetic takes over: ~0.9 ms per dense two-qubit matrix at 30 qubits against
// 3.3 ms per block pass, i.e. 0.015 per unit of HostOp::cost (profiles/r01_ncu_tile_sweep_fused_after.md).
inline double plan_time(const std::vector<MixedStep>& steps, const std::vector<HostOp>& ops) {
    double t = 0.0;
    for (const MixedStep& st : steps) {
        if (st.block) { t += 1.0; continue; }
        double c = 0.0;
        for (int i : st.sweep.ops) c += ops[i].cost();
        t += std::max(0.75, 0.015 * c);
    }
    return t;
}

// Plan under both policies and keep the plan with the shorter estimated duration.
inline std::vector<MixedStep> plan_mixed(const std::vector<HostOp>& ops, unsigned n, const PlanLimits& L, const BlockLimits& BL) {
    std::vector<MixedStep> greedy = plan_mixed_policy(ops, n, L, BL, 0);
    bool any_block = false;
    for (const MixedStep& st : greedy) any_block = any_block || st.block;
    if (!any_block) {                                          // no block formed: the plan is a plain sweep partition, so the
        std::vector<SweepPlan> plans = plan_sweeps(ops, n, L);  // two-direction sweep planner may know a shorter one
        if (plans.size() >= greedy.size()) return greedy;
        std::vector<MixedStep> steps(plans.size());
        for (size_t i = 0; i < plans.size(); ++i) steps[i].sweep = std::move(plans[i]);
        return steps;
    }
    std::vector<MixedStep> spread = plan_mixed_policy(ops, n, L, BL, 1);
    return plan_time(spread, ops) < plan_time(greedy, ops) ? spread : greedy;
}

// ---- order of the ops inside one sweep ---------------------------------------------------------------------------
// Ops that commute (disjoint qubits, or acting diagonally on the shared ones) may be reordered inside a sweep.  This
// pulls together the ops that fit in one register window of V qubits (same greedy/blocked-set rule as plan_sweeps, one
// level down), so that build_phases finds long phases: for a brick circuit {(a,b),(c,d),(b,c)} instead of {(a,b),(c,d)}.
inline std::vector<int> phase_friendly_order(const SweepPlan& sp, const std::vector<HostOp>& ops, unsigned V) {
    std::vector<int> rest = sp.ops, out;
    out.reserve(rest.size());
    while (!rest.empty()) {
        uint64_t W = 0, blockedAny = 0, blockedND = 0;
        std::vector<int> keep;
        bool took = false;
        for (int idx : rest) {
            const HostOp& o = ops[idx];
            const uint64_t nd = o.nondiag(), dg = o.qubits() & ~nd;
            const bool eligible = o.kind == HostOp::DIAG || o.kind == HostOp::DIAGP || o.kind == HostOp::PERM_X || o.kind == HostOp::PERM_SWAP ||
                                  (o.kind == HostOp::DENSE && o.targets.size() <= 2 && !o.ext);
            const bool free_ = !((nd & (blockedAny | blockedND)) || (dg & blockedAny));
            if (free_ && eligible && (unsigned)__builtin_popcountll(W | nd) <= V) {
                W |= nd;
                out.push_back(idx);
                took = true;
            } else if (free_ && !took && !eligible) {          // a wide / device-matrix op at the head: runs alone
                out.push_back(idx);
                took = true;
                blockedAny |= nd;
                blockedND |= dg;
                W = ~0ull;                                       // nothing joins it
            } else {
                keep.push_back(idx);
                blockedAny |= nd;
                blockedND |= dg;
            }
        }
        if (!took) { out.push_back(keep.front()); keep.erase(keep.begin()); }
        rest.swap(keep);
    }
    return out;
}

// ---- phases: which ops share one shared-memory round trip of the tile (see rq_phase) -------------------------------
template <typename Prog>
inline void build_phases(Prog& P, unsigned T) {
    const unsigned V = RQ_WINDOW_BITS, MINP = P.hdr.swz ? 0u : (unsigned)RQ_WINDOW_MIN_POS;
    P.hdr.nphases = 0;
    P.hdr.max_phase_ops = 0;
    const unsigned nops = P.hdr.nops;
    auto legacy = [&](unsigned i) {
        rq_phase& ph = P.phases[P.hdr.nphases++];
        ph = rq_phase{};
        ph.kind = 0; ph.first = (uint8_t)i; ph.count = 1;
        if (P.hdr.max_phase_ops < 1) P.hdr.max_phase_ops = 1;
        if (P.ops[i].kind == RQ_OP_DIAGP) P.ops[i].t[3] = 0xFF;
    };
    // window-eligible: DIAG (needs nothing); DENSE k<=2 from the pool and PERM with every target at a position >= MINP
    auto need_of = [&](const rq_tile_op& o, uint32_t& need) -> bool {
        need = 0;
        if (o.kind == RQ_OP_DIAG || o.kind == RQ_OP_DIAGP) return true;
        if (o.kind == RQ_OP_DENSE) {
            if (o.ext || o.k > 2) return false;
            for (unsigned b = 0; b < o.k; ++b) { if (o.t[b] < MINP) return false; need |= 1u << o.t[b]; }
            return true;
        }
        // PERM: xm holds the target positions
        for (unsigned j = 0; j < T; ++j) if ((o.xm >> j) & 1u) { if (j < MINP) return false; need |= 1u << j; }
        return true;
    };
    // the kernel's window loop has a warp-uniform trip count: 2^(T-V) groups must fill the 2^8 threads
    if (T < MINP + V || T < V + 8) { for (unsigned i = 0; i < nops; ++i) legacy(i); return; }
    unsigned i = 0;
    while (i < nops) {
        uint32_t need = 0;
        if (!need_of(P.ops[i], need)) { legacy(i++); continue; }
        uint32_t W = 0;
        const unsigned first = i;
        unsigned ndp = 0;                                  // RQ_OP_DIAGP ops of the phase: their thread factors live in registers
        while (i < nops) {
            uint32_t nd = 0;
            if (!need_of(P.ops[i], nd)) break;
            if ((unsigned)__builtin_popcount(W | nd) > V) break;
            if (P.ops[i].kind == RQ_OP_DIAGP && ndp >= RQ_PHASE_MAX_DIAGP) break;
            // a one-qubit dense op directly followed by the ladder hanging on its qubit becomes one butterfly (below) if both
            // land in the same phase: do not end the phase between them
            if (P.ops[i].kind == RQ_OP_DENSE && P.ops[i].k == 1 && !P.ops[i].setmask && !P.ops[i].gcmask && i + 1 < nops &&
                P.ops[i + 1].kind == RQ_OP_DIAGP && !P.ops[i + 1].gcmask && P.ops[i + 1].setmask == (1u << P.ops[i].t[0]) &&
                ndp >= RQ_PHASE_MAX_DIAGP && i > first) break;
            ndp += P.ops[i].kind == RQ_OP_DIAGP;
            W |= nd;
            ++i;
        }
        if (W == 0 || i - first < 2) {                     // diagonal-only run, or a single op: nothing to share
            for (unsigned j = first; j < i; ++j) legacy(j);
            continue;
        }
        for (unsigned p = T; p-- > MINP && (unsigned)__builtin_popcount(W) < V;) if (!((W >> p) & 1u)) W |= 1u << p;
        rq_phase& ph = P.phases[P.hdr.nphases++];
        ph = rq_phase{};
        ph.kind = 1; ph.v = (uint8_t)V; ph.first = (uint8_t)first; ph.count = (uint8_t)(i - first);
        int widx[32];
        unsigned nb = 0;
        for (unsigned p = 0; p < T; ++p) { widx[p] = -1; if ((W >> p) & 1u) { ph.w[nb] = (uint8_t)p; widx[p] = (int)nb++; } }
        if (ph.count > P.hdr.max_phase_ops) P.hdr.max_phase_ops = ph.count;
        ndp = 0;
        for (unsigned j = first; j < i; ++j) {
            rq_tile_op& o = P.ops[j];
            if (o.kind == RQ_OP_DIAG) { o.cm_in = 0; o.cm_out = 0; continue; }      // controls are evaluated per amplitude
            uint32_t lc = o.setmask;                       // local controls (+ for SWAP the select bit, removed below)
            if (o.kind == RQ_OP_DIAGP) {
                o.t[3] = (uint8_t)ndp++;
            } else if (o.kind == RQ_OP_DENSE) {
                for (unsigned b = 0; b < o.k; ++b) o.wt[b] = (uint8_t)widx[o.t[b]];
            } else {
                unsigned b = 0;
                for (unsigned p = 0; p < T; ++p) if ((o.xm >> p) & 1u) o.wt[b++] = (uint8_t)widx[p];
                o.k = (uint8_t)b;                          // 1 = X-type, 2 = SWAP-type
                lc &= ~o.xm;
            }
            o.cm_in = 0; o.cm_out = 0;
            for (unsigned p = 0; p < T; ++p) {
                if (!((lc >> p) & 1u)) continue;
                if (widx[p] >= 0) o.cm_in |= (uint8_t)(1u << widx[p]); else o.cm_out |= 1u << p;
            }
        }
        // Peephole: an uncontrolled Hadamard directly followed by the merged ladder it feeds (RQ_OP_DIAGP whose only
        // control is the Hadamard's qubit: every step "H(i); CP(j, i) for j > i" of a QFT) is one radix-2 butterfly --
        // add, subtract, one complex multiply -- instead of a dense 2x2 (16 FMA per pair) plus a phase pass.
        // Hadamard-like: real, all four entries +-v with v ~ 1/sqrt(2) and an odd number of minus signs (H, and H with an X
        // folded into it by push_x_forward).  a0' = m00 (a0 +- a1), a1' = m10 (a0 -+ a1) * phase.
        for (unsigned j = first; j + 1 < i; ++j) {
            rq_tile_op& h = P.ops[j];
            rq_tile_op& d = P.ops[j + 1];
            if (h.kind != RQ_OP_DENSE || h.k != 1 || h.ext || h.cm_in || h.cm_out || h.gcmask || h.setmask) continue;
            if (d.kind != RQ_OP_DIAGP || d.cm_out || d.gcmask || d.cm_in != (uint8_t)(1u << h.wt[0])) continue;
            const rq_cplx* M = P.pool + h.moff;
            const rq_cplx m00 = M[0], m10 = M[1 * RQ_MSLOTS], m01 = M[2 * RQ_MSLOTS], m11 = M[3 * RQ_MSLOTS];
            const rq_real v = m00.x < 0 ? -m00.x : m00.x;
            auto absq = [](rq_real x) { return x < 0 ? -x : x; };
            if (m00.y != 0 || m10.y != 0 || m01.y != 0 || m11.y != 0) continue;
            if (absq(m10.x) != v || absq(m01.x) != v || absq(m11.x) != v || absq(v * v - (rq_real)0.5) > (rq_real)1e-6) continue;
            if (!((m00.x * m01.x > 0) != (m10.x * m11.x > 0))) continue;       // rows must be (+,+)/(+,-) in some order and sign
            h.fuse = RQ_FUSE_SKIP;
            d.fuse = RQ_FUSE_BUTTERFLY;
        }
        bool chain = ((i - first) & 1u) == 0;                      // nothing but butterflies: the interpreter-free routine
        for (unsigned j = first; j < i && chain; j += 2) chain = P.ops[j].fuse == RQ_FUSE_SKIP && P.ops[j + 1].fuse == RQ_FUSE_BUTTERFLY;
        if (chain) ph.kind = 2;
    }
}

// ---- program emission ---------------------------------------------------------------------------------
template <typename Prog>
inline bool build_program(Prog& P, const SweepPlan& sp, const std::vector<HostOp>& ops, unsigned n, size_t batch,
                          uint64_t high_base) {
    const unsigned T = (unsigned)sp.res.size();
    if (T > n) return false;
    for (unsigned r : sp.res) if (r >= n) return false;          // e.g. a rank bit of a distributed state: never resident
    P.hdr.n = n; P.hdr.T = T; P.hdr.nops = 0; P.hdr.rowbits = sp.rowbits;
    P.hdr.ntiles = (uint64_t)batch << (n - T);
    P.hdr.high_base = high_base;
    P.hdr.ext_matrix = nullptr;
    P.hdr.ndiagp = 0;
    int local[64];
    uint64_t R = 0;
    for (int q = 0; q < 64; ++q) local[q] = -1;
    for (unsigned j = 0; j < T; ++j) { P.hdr.res[j] = (uint8_t)sp.res[j]; local[sp.res[j]] = (int)j; R |= 1ull << sp.res[j]; }
    unsigned pool = 0;
    struct PendingDiagp { unsigned op; cd C; std::vector<cd> floc, G; std::vector<uint8_t> gbit; };
    std::vector<PendingDiagp> pending;
    const unsigned maxops = (unsigned)(sizeof(P.ops) / sizeof(P.ops[0])), maxpool = (unsigned)(sizeof(P.pool) / sizeof(P.pool[0]));
    std::vector<int> order = phase_friendly_order(sp, ops, RQ_WINDOW_BITS);
    // Swaps that nothing after them in the sweep depends on -- the bit reversal at the end of a QFT -- need no pass over the
    // tile: exchanging two resident bits is exchanging the positions their values are STORED to.  Such swaps (uncontrolled,
    // both qubits resident above the minimal row) leave the program and become the store map sres[]; rows shrink to the
    // lowest swapped position if need be.  Every phase saved is one shared-memory round trip of the tile (complex128:
    // 2 x 64 KB at 128 B/clk per SM, the bound of the swap-heavy QFT sweeps).
    {
        unsigned where[16];                                   // local bit j's value is stored to position res[where[j]]
        for (unsigned j = 0; j < T; ++j) where[j] = j;
        uint64_t later = 0;                                   // qubits of the ops that stay, from the end backwards
        std::vector<int> keep, absorbed;
        unsigned rowbits = P.hdr.rowbits;
        for (size_t i = order.size(); i-- > 0;) {
            const HostOp& o = ops[order[i]];
            bool take = o.kind == HostOp::PERM_SWAP && o.cmask == 0 && o.targets.size() == 2 && !(o.qubits() & later);
            if (take) {
                const int la = local[o.targets[0]], lb = local[o.targets[1]];
                take = la >= 0 && lb >= 0 && (unsigned)std::min(la, lb) >= std::min<unsigned>(RQ_MIN_ROW_BITS, T);
            }
            if (take) absorbed.push_back(order[i]);
            else { keep.push_back(order[i]); later |= o.qubits(); }
        }
        if (!absorbed.empty() && !keep.empty()) {             // (a sweep of nothing but swaps keeps its permutation passes: rare, and it is HBM-bound anyway)
            std::reverse(keep.begin(), keep.end());
            std::reverse(absorbed.begin(), absorbed.end());   // program order
            for (int idx : absorbed) {
                const unsigned la = (unsigned)local[ops[idx].targets[0]], lb = (unsigned)local[ops[idx].targets[1]];
                for (unsigned j = 0; j < T; ++j) { if (where[j] == la) where[j] = lb; else if (where[j] == lb) where[j] = la; }
                rowbits = std::min(rowbits, std::min(la, lb));
            }
            order.swap(keep);
            P.hdr.rowbits = rowbits;
        }
        for (unsigned j = 0; j < 16; ++j) P.hdr.sres[j] = j < T ? P.hdr.res[where[j]] : 0;
    }
    for (int idx : order) {
        const HostOp& o = ops[idx];
        if (P.hdr.nops >= maxops) return false;
        rq_tile_op& t = P.ops[P.hdr.nops++];
        t = rq_tile_op{};
        uint32_t fixmask = 0;
        for (unsigned q = 0; q < 64; ++q) {
            if (!((o.cmask >> q) & 1ull)) continue;
            if (local[q] >= 0) { t.setmask |= 1u << local[q]; fixmask |= 1u << local[q]; }
            else t.gcmask |= 1ull << q;
        }
        const unsigned k = (unsigned)o.targets.size();
        if (o.kind == HostOp::DENSE) {
            t.kind = RQ_OP_DENSE; t.k = (uint8_t)k;
            for (unsigned b = 0; b < k; ++b) { if (local[o.targets[b]] < 0) return false; t.t[b] = (uint8_t)local[o.targets[b]]; fixmask |= 1u << t.t[b]; }
            if (o.ext) { t.ext = 1; P.hdr.ext_matrix = o.ext; }
            else {
                const unsigned need = 1u << (2 * k);
                if (RQ_MSLOTS == 2 && (pool & 1u)) ++pool;         // 16-byte alignment of the two-slot elements
                if (pool + RQ_MSLOTS * need > maxpool) return false;
                t.moff = pool;
                const bool flip = k == 2 && t.t[0] > t.t[1];          // keep 2q targets ascending: matrix bit 0 <-> lower position
                if (flip) std::swap(t.t[0], t.t[1]);
                for (unsigned e = 0; e < need; ++e) {
                    unsigned src = e;
                    if (flip) {                                       // swap the two index bits of row and column
                        const unsigned r = e & 3u, c = e >> 2;
                        const unsigned rs = ((r & 1u) << 1) | (r >> 1), cs = ((c & 1u) << 1) | (c >> 1);
                        src = rs + 4u * cs;
                    }
                    const rq_real re = (rq_real)o.data[src].real(), im = (rq_real)o.data[src].imag();
                    P.pool[pool + RQ_MSLOTS * e].x = re;
                    P.pool[pool + RQ_MSLOTS * e].y = im;
                    if (RQ_MSLOTS == 2) { P.pool[pool + 2 * e + 1].x = -im; P.pool[pool + 2 * e + 1].y = im; }
                }
                pool += RQ_MSLOTS * need;
            }
        } else if (o.kind == HostOp::DIAG) {
            t.kind = RQ_OP_DIAG; t.k = (uint8_t)k;
            for (unsigned b = 0; b < k; ++b) {
                if (local[o.targets[b]] >= 0) t.t[b] = (uint8_t)local[o.targets[b]];
                else { t.t[b] = 0xFF; t.gq[b] = (uint8_t)o.targets[b]; }
            }
            const unsigned need = 1u << k;
            if (pool + need > maxpool) return false;
            t.moff = pool;
            for (unsigned e = 0; e < need; ++e) { P.pool[pool + e].x = (rq_real)o.data[e].real(); P.pool[pool + e].y = (rq_real)o.data[e].imag(); }
            pool += need;
        } else if (o.kind == HostOp::DIAGP) {
            // the pool layout depends on the phase the op lands in: emitted by emit_diagp below, after build_phases
            t.kind = RQ_OP_DIAGP;
            if (P.hdr.ndiagp >= RQ_MAX_DIAGP) return false;
            t.t[2] = P.hdr.ndiagp;
            P.hdr.diagp_op[P.hdr.ndiagp++] = (uint8_t)(P.hdr.nops - 1);
            PendingDiagp pd;
            pd.op = P.hdr.nops - 1;
            pd.C = o.data[0];
            pd.floc.assign(T, cd(1.0, 0.0));
            for (unsigned b = 0; b < k; ++b) {
                const unsigned q = o.targets[b];
                if (local[q] >= 0) { pd.floc[local[q]] *= o.data[1 + b]; continue; }
                unsigned bit;
                if (q < n) { bit = q; for (unsigned j = 0; j < T; ++j) if (sp.res[j] < q) --bit; }     // rank among the non-resident positions
                else bit = (n - T) + (q - n);                                                        // rank bit of a distributed state
                if (bit >= 64) return false;
                pd.G.push_back(o.data[1 + b]);
                pd.gbit.push_back((uint8_t)bit);
            }
            pending.push_back(std::move(pd));
        } else if (o.kind == HostOp::PERM_X) {
            t.kind = RQ_OP_PERM;
            if (local[o.targets[0]] < 0) return false;
            t.xm = 1u << local[o.targets[0]];
            fixmask |= t.xm;
        } else {
            t.kind = RQ_OP_PERM;
            if (local[o.targets[0]] < 0 || local[o.targets[1]] < 0) return false;
            const uint32_t a = 1u << local[o.targets[0]], b = 1u << local[o.targets[1]];
            t.xm = a | b;
            t.setmask |= a;                 // enumerate a=1,b=0; partner has a=0,b=1
            fixmask |= a | b;
        }
        for (unsigned j = 0; j < T; ++j) if ((fixmask >> j) & 1u) t.fix[t.nfix++] = (uint8_t)j;
    }
    // Register windows over the bank-selecting bits would serialise 2^RQ_SWZ_BITS-fold on shared-memory banks: sweeps with
    // several ops on those bits keep the tile XOR-swizzled (needs 2*RQ_SWZ_BITS local bits).  The swizzle costs two extra
    // passes over the tile, more than the two-way conflicts of ONE op applied in place, so a lone low-bit op (every eager
    // rocsvApply* call on qubits 0-3) stays on the linear layout and remains HBM-bound.
    P.hdr.swz = 0;
    if (T >= 2 * RQ_SWZ_BITS) {
        unsigned low_ops = 0;
        for (unsigned i = 0; i < P.hdr.nops; ++i) {
            const rq_tile_op& t = P.ops[i];
            uint32_t tm = 0;
            if (t.kind == RQ_OP_DENSE) for (unsigned b = 0; b < t.k; ++b) tm |= 1u << t.t[b];
            else if (t.kind == RQ_OP_PERM) tm = t.xm;
            if (tm & ((1u << RQ_SWZ_BITS) - 1u)) ++low_ops;
        }
        P.hdr.swz = low_ops >= 2 ? 1 : 0;
    }
    build_phases(P, T);
    // RQ_OP_DIAGP tables.  pool: C | A[na]: factor of group-index bit i < na | B[2^nb]: product over the group-index bits
    // above na | (register-window phase only) W[2^V]: product over the window bits | G[ng]: factors of non-resident qubits |
    // ng bytes: which bit of the tile's "outer" word (tile_sweep.cuh) each of them reads.
    // Group-index bit i <-> i-th "free" local position: in a pass of its own every position that is not a control (controls
    // are fixed by the enumeration); in a window phase every non-window position (controls are checked per amplitude).
    for (const PendingDiagp& pd : pending) {
        rq_tile_op& t = P.ops[pd.op];
        const rq_phase* ph = nullptr;
        for (unsigned i = 0; i < P.hdr.nphases; ++i)
            if (P.phases[i].first <= pd.op && pd.op < (unsigned)P.phases[i].first + P.phases[i].count) ph = &P.phases[i];
        if (!ph) return false;
        const bool win = ph->kind == 1 || ph->kind == 2;
        if (win && t.fuse == RQ_FUSE_BUTTERFLY) {                  // no factor on the window bits below the hub?
            bool up = true;
            for (unsigned b = 0; b < ph->v && !((t.cm_in >> b) & 1u); ++b) up = up && pd.floc[ph->w[b]] == cd(1.0, 0.0);
            if (up) t.fuse = RQ_FUSE_BUTTERFLY_UP;
        }
        uint32_t skip = 0;
        if (win) for (unsigned b = 0; b < ph->v; ++b) skip |= 1u << ph->w[b];
        else for (unsigned f = 0; f < t.nfix; ++f) skip |= 1u << t.fix[f];
        std::vector<unsigned> nf;
        for (unsigned j = 0; j < T; ++j) if (!((skip >> j) & 1u)) nf.push_back(j);
        const unsigned nfree = (unsigned)nf.size(), na = std::min(nfree, 8u), nb = nfree - na, ng = (unsigned)pd.G.size();
        const unsigned nw = win ? (1u << ph->v) : 0u, nbytes_cplx = (ng + (unsigned)sizeof(rq_cplx) - 1u) / (unsigned)sizeof(rq_cplx);
        const unsigned need = 1u + na + (1u << nb) + nw + ng + nbytes_cplx;
        if (pool + need > maxpool || ng > 255) return false;
        t.moff = pool; t.k = (uint8_t)ng; t.t[0] = (uint8_t)na; t.t[1] = (uint8_t)nb;
        t.xm = 1u + na + (1u << nb) + nw;
        auto put = [&](const cd& c) { P.pool[pool].x = (rq_real)c.real(); P.pool[pool].y = (rq_real)c.imag(); ++pool; };
        put(pd.C);
        for (unsigned i = 0; i < na; ++i) put(pd.floc[nf[i]]);
        for (unsigned m = 0; m < (1u << nb); ++m) {
            cd f(1.0, 0.0);
            for (unsigned i = 0; i < nb; ++i) if ((m >> i) & 1u) f *= pd.floc[nf[na + i]];
            put(f);
        }
        for (unsigned jw = 0; jw < nw; ++jw) {
            cd f(1.0, 0.0);
            for (unsigned b = 0; b < ph->v; ++b) if ((jw >> b) & 1u) f *= pd.floc[ph->w[b]];
            put(f);
        }
        for (const cd& g : pd.G) put(g);
        if (nbytes_cplx) {
            memset(&P.pool[pool], 0, nbytes_cplx * sizeof(rq_cplx));
            memcpy(&P.pool[pool], pd.gbit.data(), ng);
            pool += nbytes_cplx;
        }
    }
    return true;
}

// ---- text dump of a plan (rocsvxPlanCircuit): enough to re-simulate it independently --------------------
inline std::string dump_ops(const std::vector<int>& order, const std::vector<HostOp>& ops) {
    std::string s;
    char buf[128];
    for (int i : order) {
        const HostOp& o = ops[i];
        snprintf(buf, sizeof buf, "O %d cmask %llx targets", o.kind, (unsigned long long)o.cmask);
        s += buf;
        for (unsigned t : o.targets) { snprintf(buf, sizeof buf, " %u", t); s += buf; }
        s += " data";
        for (const cd& c : o.data) { snprintf(buf, sizeof buf, " %.17g %.17g", c.real(), c.imag()); s += buf; }
        s += "\n";
    }
    return s;
}

inline std::string dump_plan(const std::vector<SweepPlan>& plans, const std::vector<HostOp>& ops) {
    std::string s;
    char buf[128];
    for (const SweepPlan& sp : plans) {
        snprintf(buf, sizeof buf, "S %u %u res:", (unsigned)sp.res.size(), sp.rowbits);
        s += buf;
        for (unsigned r : sp.res) { snprintf(buf, sizeof buf, " %u", r); s += buf; }
        s += "\n";
        s += dump_ops(phase_friendly_order(sp, ops, RQ_WINDOW_BITS), ops);      // the order the kernel executes
    }
    return s;
}

// ---- sampled basis indices -> result words ---------------------------------------------------------------------------
// Bit j of a result word is bit measured[j] of the sampled basis index (hipStateVec.h:427-445).  Maximal runs
// measured[j + 1] == measured[j] + 1 move as ONE shifted field, so the usual request "every qubit, in order" costs a single
// mask per shot instead of one step per bit (a million 28-bit shots: 28 M steps on the host otherwise).
struct BitGather {
    struct Field { unsigned src, dst; uint64_t mask; };
    std::vector<Field> fields;
    BitGather(const unsigned* measured, unsigned nm) {
        for (unsigned j = 0; j < nm;) {
            unsigned len = 1;
            while (j + len < nm && measured[j + len] == measured[j] + len) ++len;
            fields.push_back(Field{measured[j], j, len >= 64 ? ~0ull : ((1ull << len) - 1ull)});
            j += len;
        }
    }
    uint64_t operator()(uint64_t idx) const {
        uint64_t bits = 0;
        for (const Field& f : fields) bits |= ((idx >> f.src) & f.mask) << f.dst;
        return bits;
    }
};

}  // namespace rq
