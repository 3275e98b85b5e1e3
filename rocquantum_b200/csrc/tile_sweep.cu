// rocquantum_b200/csrc/tile_sweep.cu  --  the fused gate sweep, hand-written for sm_100a.
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
__device__ __forceinline__ rq_cplx cmul(rq_cplx a, rq_cplx b) {
    rq_cplx r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
__device__ __forceinline__ void cfma(rq_cplx& acc, rq_cplx m, rq_cplx v) {
    acc.x += m.x * v.x - m.y * v.y;
    acc.y += m.x * v.y + m.y * v.x;
}

// deposit the bits of g around the fixed positions fix[0..nfix) (ascending), leaving zeros there
__device__ __forceinline__ uint32_t spread(uint32_t g, const rq_tile_op& o) {
    for (uint32_t f = 0; f < o.nfix; ++f) {
        const uint32_t p = o.fix[f];
        g = ((g >> p) << (p + 1)) | (g & ((1u << p) - 1u));
    }
    return g;
}

template <int K, bool EXT>
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
    for (uint32_t g = tid; g < ngroups; g += NT) {
        const uint32_t base = spread(g, o) | o.setmask;
        rq_cplx a[D];
#pragma unroll
        for (int j = 0; j < D; ++j) a[j] = sm[base | off[j]];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            rq_cplx acc = {0, 0};
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const rq_cplx m = EXT ? ldg_cplx(M + i + j * D) : M[i + j * D];   // column-major, as the API
                cfma(acc, m, a[j]);
            }
            sm[base | off[i]] = acc;
        }
    }
}

// diagonal: amp[idx] *= d[sel], sel bit b taken from the local index or, for a non-resident qubit,
// from the tile base.  Qubits whose "0" entries are all 1 were turned into controls by the host.
__device__ __forceinline__ void op_diag(rq_cplx* sm, const rq_tile_op& o, const rq_cplx* pool, uint32_t T, uint32_t tid,
                                        uint64_t gbase) {
    uint32_t selbase = 0;
    for (uint32_t b = 0; b < o.k; ++b)
        if (o.t[b] == 0xFF) selbase |= (uint32_t)((gbase >> o.gq[b]) & 1ull) << b;
    const uint32_t ngroups = 1u << (T - o.nfix);
    const rq_cplx* D = pool + o.moff;
    for (uint32_t g = tid; g < ngroups; g += NT) {
        const uint32_t idx = spread(g, o) | o.setmask;
        uint32_t sel = selbase;
        for (uint32_t b = 0; b < o.k; ++b)
            if (o.t[b] != 0xFF) sel |= ((idx >> o.t[b]) & 1u) << b;
        sm[idx] = cmul(D[sel], sm[idx]);
    }
}

// pair permutation: swap(idx, idx ^ xm) over the idx whose fixed bits equal setmask
__device__ __forceinline__ void op_perm(rq_cplx* sm, const rq_tile_op& o, uint32_t T, uint32_t tid) {
    const uint32_t ngroups = 1u << (T - o.nfix);
    for (uint32_t g = tid; g < ngroups; g += NT) {
        const uint32_t i0 = spread(g, o) | o.setmask, i1 = i0 ^ o.xm;
        const rq_cplx a = sm[i0], b = sm[i1];
        sm[i0] = b;
        sm[i1] = a;
    }
}

// WIDE = false: dense ops of 1-2 qubits only, <= 64 registers so that 4+ tiles per SM are in flight.
// WIDE = true : also 3- and 4-qubit dense ops (16 amplitudes per thread in registers), fewer tiles per SM.
template <typename Prog, bool WIDE>
__global__ void __launch_bounds__(NT, WIDE ? 1 : 4) tile_sweep_kernel(rq_cplx* __restrict__ state, const __grid_constant__ Prog prog) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    rq_cplx* sm = reinterpret_cast<rq_cplx*>(smem_raw);
    __shared__ __align__(8) uint64_t bar_storage;

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
    mbar_wait(bar, 0);

    const rq_cplx* ext = reinterpret_cast<const rq_cplx*>(prog.hdr.ext_matrix);
    for (uint32_t i = 0; i < prog.hdr.nops; ++i) {
        const rq_tile_op& o = prog.ops[i];
        if ((gbase & o.gcmask) != o.gcmask) continue;          // uniform per tile: no divergent barrier
        switch (o.kind) {
            case RQ_OP_DENSE:
                if (o.ext) {
                    switch (o.k) {
                        case 1: op_dense<1, true>(sm, o, prog.pool, ext, T, tid); break;
                        case 2: op_dense<2, true>(sm, o, prog.pool, ext, T, tid); break;
                        case 3: if (WIDE) op_dense<3, true>(sm, o, prog.pool, ext, T, tid); break;
                        default: if (WIDE) op_dense<4, true>(sm, o, prog.pool, ext, T, tid); break;
                    }
                } else {
                    switch (o.k) {
                        case 1: op_dense<1, false>(sm, o, prog.pool, ext, T, tid); break;
                        case 2: op_dense<2, false>(sm, o, prog.pool, ext, T, tid); break;
                        case 3: if (WIDE) op_dense<3, false>(sm, o, prog.pool, ext, T, tid); break;
                        default: if (WIDE) op_dense<4, false>(sm, o, prog.pool, ext, T, tid); break;
                    }
                }
                break;
            case RQ_OP_DIAG: op_diag(sm, o, prog.pool, T, tid, gbase); break;
            default: op_perm(sm, o, T, tid); break;
        }
        __syncthreads();
    }

    fence_async_smem();       // generic-proxy writes to smem -> visible to the async (TMA) proxy
    __syncthreads();
    for (uint32_t r = tid; r < nrows; r += NT) {
        uint64_t goff = 0;
        for (uint32_t i = 0; i < T - rowbits; ++i) goff |= (uint64_t)((r >> i) & 1u) << prog.hdr.res[rowbits + i];
        bulk_s2g(gtile + goff, smem_u32(sm) + r * rowbytes, rowbytes);
    }
    bulk_commit_wait_read();  // smem must stay valid until the TMA engine has read it
}

template <typename Prog>
int launch(rq_cplx* state, const Prog* prog, void* stream) {
    const size_t smem = sizeof(rq_cplx) << prog->hdr.T;
    bool wide = false;
    for (uint32_t i = 0; i < prog->hdr.nops; ++i) wide |= (prog->ops[i].kind == RQ_OP_DENSE && prog->ops[i].k > 2);
    if (wide)
        tile_sweep_kernel<Prog, true><<<(unsigned)prog->hdr.ntiles, NT, smem, (cudaStream_t)stream>>>(state, *prog);
    else
        tile_sweep_kernel<Prog, false><<<(unsigned)prog->hdr.ntiles, NT, smem, (cudaStream_t)stream>>>(state, *prog);
    return (int)cudaGetLastError();
}

}  // namespace

extern "C" int rq_sweep_configure(void) {
    const int bytes = (int)(sizeof(rq_cplx) << RQ_MAX_TILE_BITS);
    cudaError_t e = cudaFuncSetAttribute(tile_sweep_kernel<rq_program_small, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tile_sweep_kernel<rq_program_small, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tile_sweep_kernel<rq_program_large, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tile_sweep_kernel<rq_program_large, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return (int)e;
}
extern "C" int rq_launch_sweep_small(rq_cplx* state, const rq_program_small* prog, void* stream) { return launch(state, prog, stream); }
extern "C" int rq_launch_sweep_large(rq_cplx* state, const rq_program_large* prog, void* stream) { return launch(state, prog, stream); }
