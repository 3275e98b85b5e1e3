// rocquantum_b200/csrc/dist_plan.h -- pure-host planner of a distributed circuit: turns logical gates into
// steps "run these ops on every slice" / "exchange rank bits with the top local bits", tracking the
// logical->physical qubit map.  No CUDA, no NCCL: the same code is driven by dist.cu (executor) and by
// rocsvxDistPlanCircuit (tests re-simulate the plan on the CPU).
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "host_ops.h"

namespace rq {

struct DistStep {
    enum Kind { RUN = 0, EXCHANGE = 1 };
    int kind = RUN;
    std::vector<HostOp> ops;          // RUN: ops in PHYSICAL positions (targets < n_local are slice-local)
    std::vector<unsigned> gpos;       // EXCHANGE: rank-bit positions traded with the top gpos.size() local positions, in order
};

struct DistPlanner {
    unsigned n_total = 0, n_local = 0, n_global = 0;
    std::vector<unsigned> map;        // logical -> physical
    std::vector<DistStep> steps;
    std::vector<HostOp> pending;

    void reset(unsigned total, unsigned local) {
        n_total = total; n_local = local; n_global = total - local;
        map.resize(total);
        for (unsigned q = 0; q < total; ++q) map[q] = q;
        steps.clear(); pending.clear();
    }
    uint64_t global_mask() const { return n_global ? (((1ull << n_global) - 1ull) << n_local) : 0ull; }
    unsigned logical_at(unsigned phys) const {
        for (unsigned q = 0; q < n_total; ++q) if (map[q] == phys) return q;
        return ~0u;
    }
    void to_physical(const HostOp& in, HostOp& out) const {
        out = in;
        for (unsigned& t : out.targets) t = map[t];
        uint64_t cm = 0;
        for (unsigned q = 0; q < n_total; ++q) if ((in.cmask >> q) & 1ull) cm |= 1ull << map[q];
        out.cmask = cm;
    }
    void flush_pending() {
        if (pending.empty()) return;
        DistStep s; s.kind = DistStep::RUN; s.ops.swap(pending);
        steps.push_back(std::move(s));
    }

    // Make the logical qubits `bring` (on rank bits) local by trading them with the local logical qubits `evict`.
    // Evictees are first moved to the top local slots by physical SWAP ops appended to the pending run.
    void trade(const std::vector<unsigned>& bring, const std::vector<unsigned>& evict) {
        const unsigned k = (unsigned)bring.size();
        if (k == 0) return;
        auto is_evictee = [&](unsigned logical) { return std::find(evict.begin(), evict.end(), logical) != evict.end(); };
        for (unsigned i = 0; i < k; ++i) {
            const unsigned p = map[evict[i]];
            if (p >= n_local - k) continue;                               // already in a top slot
            for (unsigned s = n_local - k; s < n_local; ++s) {
                const unsigned occupant = logical_at(s);
                if (is_evictee(occupant)) continue;
                pending.push_back(make_swap(p, s));
                map[occupant] = p;
                map[evict[i]] = s;
                break;
            }
        }
        flush_pending();
        DistStep x; x.kind = DistStep::EXCHANGE;
        for (unsigned i = 0; i < k; ++i) x.gpos.push_back(map[bring[i]]);
        for (unsigned i = 0; i < k; ++i) {                                // relabel: slot i <-> gpos[i]
            const unsigned l = n_local - k + i, g = x.gpos[i];
            const unsigned ql = logical_at(l), qg = logical_at(g);
            map[ql] = g;
            map[qg] = l;
        }
        steps.push_back(std::move(x));
    }

    // one gate: evict the highest local qubits the gate does not touch
    bool add_op(const HostOp& op) {
        HostOp phys;
        to_physical(op, phys);
        const uint64_t gm = global_mask();
        if (phys.nondiag() & gm) {
            std::vector<unsigned> bring, evict;
            for (unsigned t : op.targets) if ((1ull << map[t]) & gm & phys.nondiag()) bring.push_back(t);
            const uint64_t used = op.qubits();
            for (unsigned p = n_local; p-- > 0 && evict.size() < bring.size();) {
                const unsigned l = logical_at(p);
                if (!((used >> l) & 1ull)) evict.push_back(l);
            }
            if (evict.size() < bring.size()) return false;
            trade(bring, evict);
            to_physical(op, phys);
        }
        pending.push_back(std::move(phys));
        return true;
    }

    // Whole circuit with deferral ("cache blocking" over the rank bits).  An op whose non-diagonal targets sit on a rank bit
    // is DEFERRED instead of forcing an exchange at once, and so is every later op that does not commute past a deferred
    // one (same rule as plan_sweeps: a non-diagonal target may not meet any deferred qubit, a diagonal/control qubit may
    // not meet a deferred non-diagonal target).  Everything else in the REST OF THE CIRCUIT runs in the current layout: in
    // a brick circuit the dependency cone of a parked qubit widens by one qubit per layer, so the far side of the register
    // advances many layers per residency.  Only when nothing more can run are rank bits traded -- for the local qubits
    // whose next non-diagonal use among the deferred ops is farthest (Belady) -- and the deferred list is rescanned.
    // A depth-20 brick circuit on 34 qubits needs 2 exchanges this way instead of 20.
    bool add_circuit(const std::vector<HostOp>& ops) {
        const uint64_t gm = global_mask();
        std::vector<size_t> remaining(ops.size());
        for (size_t i = 0; i < ops.size(); ++i) remaining[i] = i;
        while (!remaining.empty()) {
            uint64_t blockedAny = 0, blockedND = 0;
            std::vector<size_t> deferred;
            for (size_t idx : remaining) {
                const HostOp& o = ops[idx];
                const uint64_t nd = o.nondiag(), dg = o.qubits() & ~nd;
                const bool free_ = !((nd & (blockedAny | blockedND)) || (dg & blockedAny));
                HostOp phys;
                to_physical(o, phys);
                if (free_ && !(phys.nondiag() & gm)) { pending.push_back(std::move(phys)); continue; }
                deferred.push_back(idx);
                blockedAny |= nd;
                blockedND |= dg;
            }
            remaining.swap(deferred);
            if (remaining.empty()) break;
            // next non-diagonal use of every logical qubit among the deferred ops
            const size_t never = remaining.size() + n_total;
            std::vector<size_t> next(n_total, never);
            for (size_t j = 0; j < remaining.size(); ++j) {
                const uint64_t nd = ops[remaining[j]].nondiag();
                for (unsigned q = 0; q < n_total; ++q)
                    if (((nd >> q) & 1ull) && next[q] == never) next[q] = j;
            }
            for (unsigned q = 0; q < n_total; ++q) if (next[q] == never) next[q] = remaining.size() + q;
            std::vector<unsigned> order(n_total);
            for (unsigned q = 0; q < n_total; ++q) order[q] = q;
            std::stable_sort(order.begin(), order.end(), [&](unsigned a, unsigned b) { return next[a] > next[b]; });
            std::vector<char> want_global(n_total, 0);
            for (unsigned r = 0; r < n_global; ++r) want_global[order[r]] = 1;
            std::vector<unsigned> bring, evict;
            for (unsigned q = 0; q < n_total; ++q) {
                const bool is_global = map[q] >= n_local;
                if (is_global && !want_global[q]) bring.push_back(q);
                if (!is_global && want_global[q]) evict.push_back(q);
            }
            if (bring.size() != evict.size() || bring.empty()) return false;
            trade(bring, evict);
            HostOp phys;
            to_physical(ops[remaining[0]], phys);                     // progress: the oldest deferred op can run now
            if (phys.nondiag() & gm) return false;
        }
        return true;
    }

    // whole circuit, strictly in program order: on a global non-diagonal target trade ALL rank bits for the local qubits whose next
    // non-diagonal use is farthest (Belady), so one exchange is amortised over as many gates as possible
    bool add_circuit_inorder(const std::vector<HostOp>& ops) {
        const uint64_t gm = global_mask();
        for (size_t i = 0; i < ops.size(); ++i) {
            HostOp phys;
            to_physical(ops[i], phys);
            if (phys.nondiag() & gm) {
                std::vector<size_t> next(n_total, ops.size() + n_total);
                unsigned found = 0;
                for (size_t j = i; j < ops.size() && found < n_total; ++j) {
                    const uint64_t nd = ops[j].nondiag();
                    for (unsigned q = 0; q < n_total; ++q)
                        if (((nd >> q) & 1ull) && next[q] >= ops.size()) { next[q] = j; ++found; }
                }
                for (unsigned q = 0; q < n_total; ++q) if (next[q] >= ops.size()) next[q] = ops.size() + q;
                std::vector<unsigned> order(n_total);
                for (unsigned q = 0; q < n_total; ++q) order[q] = q;
                std::stable_sort(order.begin(), order.end(), [&](unsigned a, unsigned b) { return next[a] > next[b]; });
                std::vector<char> want_global(n_total, 0);
                for (unsigned r = 0; r < n_global; ++r) want_global[order[r]] = 1;
                std::vector<unsigned> bring, evict;
                for (unsigned q = 0; q < n_total; ++q) {
                    const bool is_global = map[q] >= n_local;
                    if (is_global && !want_global[q]) bring.push_back(q);
                    if (!is_global && want_global[q]) evict.push_back(q);
                }
                if (bring.size() != evict.size() || bring.empty()) return false;
                trade(bring, evict);
                to_physical(ops[i], phys);
                if (phys.nondiag() & gm) return false;
            }
            pending.push_back(std::move(phys));
        }
        return true;
    }

    // bring a set of logical qubits local (expectation values with X/Y factors on rank bits)
    bool make_local(uint64_t logical_mask) {
        const uint64_t gm = global_mask();
        std::vector<unsigned> bring, evict;
        for (unsigned q = 0; q < n_total; ++q) if (((logical_mask >> q) & 1ull) && ((1ull << map[q]) & gm)) bring.push_back(q);
        if (bring.empty()) return true;
        for (unsigned p = n_local; p-- > 0 && evict.size() < bring.size();) {
            const unsigned l = logical_at(p);
            if (!((logical_mask >> l) & 1ull)) evict.push_back(l);
        }
        if (evict.size() < bring.size()) return false;
        trade(bring, evict);
        return true;
    }

    // restore the identity layout
    void canonicalize() {
        bool ident = true;
        for (unsigned q = 0; q < n_total; ++q) ident = ident && map[q] == q;
        if (ident) return;
        if (n_global) {
            std::vector<unsigned> bring, evict;
            for (unsigned q = 0; q < n_total; ++q) {
                const bool is_global = map[q] >= n_local, should = q >= n_local;
                if (is_global && !should) bring.push_back(q);
                if (!is_global && should) evict.push_back(q);
            }
            trade(bring, evict);
            for (unsigned g = n_local; g < n_total; ++g) {                // rank bits holding the wrong global qubit
                if (map[g] == g) continue;
                const unsigned occupant = logical_at(g), spare = logical_at(n_local - 1);
                trade({occupant}, {spare});                               // occupant -> local, spare -> position g
                trade({g}, {occupant});                                   // logical g -> local, occupant -> g's old position
                trade({spare}, {g});                                      // logical g -> position g, spare -> local
            }
        }
        for (unsigned q = 0; q < n_local; ++q) {                          // local positions: swaps
            if (map[q] == q) continue;
            const unsigned p = map[q], other = logical_at(q);
            pending.push_back(make_swap(p, q));
            map[other] = p;
            map[q] = q;
        }
        flush_pending();
    }

    std::string dump() const {
        std::string s;
        char buf[128];
        for (const DistStep& st : steps) {
            if (st.kind == DistStep::EXCHANGE) {
                s += "X";
                for (unsigned g : st.gpos) { snprintf(buf, sizeof buf, " %u", g); s += buf; }
                s += "\n";
                continue;
            }
            s += "R\n";
            for (const HostOp& o : st.ops) {
                snprintf(buf, sizeof buf, "O %d cmask %llx targets", o.kind, (unsigned long long)o.cmask);
                s += buf;
                for (unsigned t : o.targets) { snprintf(buf, sizeof buf, " %u", t); s += buf; }
                s += " data";
                for (const cd& c : o.data) { snprintf(buf, sizeof buf, " %.17g %.17g", c.real(), c.imag()); s += buf; }
                s += "\n";
            }
        }
        s += "M";
        for (unsigned q = 0; q < n_total; ++q) { snprintf(buf, sizeof buf, " %u", map[q]); s += buf; }
        s += "\n";
        return s;
    }
};

}  // namespace rq
