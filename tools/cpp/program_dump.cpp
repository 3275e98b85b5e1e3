// tools/cpp/program_dump.cpp -- host-only: what the tile-sweep kernel will execute for the 33-qubit complex128 QFT
// (or another QFT size): sweeps, resident sets, phases, ops per phase and the peephole fusions of build_phases.
//   g++ -O1 -std=c++17 -DROCQ_PRECISION_DOUBLE -I include -I /usr/local/cuda/include tools/cpp/program_dump.cpp -o /tmp/program_dump && /tmp/program_dump 33
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../rocquantum_b200/csrc/gate_convert.h"
#include "../../rocquantum_b200/csrc/host_ops.h"

int main(int argc, char** argv) {
    const unsigned n = argc > 1 ? (unsigned)atoi(argv[1]) : 33;
    std::vector<rq::HostOp> ops;
    const double s = 1.0 / std::sqrt(2.0);
    for (unsigned q = 0; q < n; q += 3) ops.push_back(rq::make_x(q));                     // some input basis state
    for (unsigned i = 0; i < n; ++i) {
        ops.push_back(rq::make_dense1(i, s, s, s, -s));
        for (unsigned j = i + 1; j < n; ++j) {
            const double th = M_PI / std::ldexp(1.0, (int)(j - i));
            ops.push_back(rq::make_phase((1ull << i) | (1ull << j), rq::cd(std::cos(th), std::sin(th))));
        }
    }
    for (unsigned i = 0; i < n / 2; ++i) ops.push_back(rq::make_swap(i, n - 1 - i));
    std::vector<rq::HostOp> fused = rq::fuse_algebraic(rq::merge_diagonals(rq::push_x_forward(ops)), n);
    rq::PlanLimits L;
    L.max_ops = sizeof(rq_program_large::ops) / sizeof(rq_tile_op);
    L.pool_cplx = sizeof(rq_program_large::pool) / sizeof(rq_cplx);
    const std::vector<rq::SweepPlan> plans = rq::plan_sweeps(fused, n, L);
    static rq_program_large P;
    printf("%zu ops after fusion, %zu sweeps\n", fused.size(), plans.size());
    for (const rq::SweepPlan& sp : plans) {
        if (!rq::build_program(P, sp, fused, n, 1, 0)) { printf("build_program failed\n"); return 1; }
        printf("sweep: T=%u rowbits=%u swz=%u ops=%u phases=%u res:", P.hdr.T, P.hdr.rowbits, P.hdr.swz, P.hdr.nops, P.hdr.nphases);
        for (unsigned j = 0; j < P.hdr.T; ++j) printf(" %u", P.hdr.res[j]);
        printf("\n");
        for (unsigned p = 0; p < P.hdr.nphases; ++p) {
            const rq_phase& ph = P.phases[p];
            printf("  phase %u kind %u window", p, ph.kind);
            if (ph.kind == 1) for (unsigned b = 0; b < ph.v; ++b) printf(" %u", ph.w[b]);
            printf(" :");
            for (unsigned i = ph.first; i < (unsigned)ph.first + ph.count; ++i) {
                const rq_tile_op& o = P.ops[i];
                const char* k = o.kind == RQ_OP_DENSE ? "D" : o.kind == RQ_OP_DIAG ? "G" : o.kind == RQ_OP_DIAGP ? "P" : "X";
                printf(" %s%u", k, o.k);
                if (o.kind == RQ_OP_DENSE) printf("(t%u)", o.t[0]);
                if (o.fuse) printf("[f%u]", o.fuse);
                if (o.cm_in || o.cm_out || o.gcmask) printf("{ci%x co%x g%llx}", o.cm_in, o.cm_out, (unsigned long long)o.gcmask);
            }
            printf("\n");
        }
    }
    return 0;
}
