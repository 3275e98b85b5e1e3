// rocquantum_b200/csrc/block_sweep.cu -- tensor-core sweep: ONE fused 6-qubit dense block (a 64x64 complex unitary
// holding every gate the planner could fold into six qubits) applied to the whole state in one HBM pass.
//
// Why: ncu shows the fused CUDA-core sweeps of the depth-40 random-unitary circuit issue/FMA bound (profiles/), at
// ~0.9 ms per dense 2q gate on 30 qubits against 2.5 ms for the HBM pass.  A 6-qubit block is a real 128x128 matrix
// (complex -> [[Re,-Im],[Im,Re]]) times a 128 x (#columns) panel: GEMM-shaped, so it goes to the 5th-gen tensor cores.
//
// How (sm_100a):
//   * tile = 64 block values x 128 columns.  D[column][out] = sum_k X'[column][k] * A'[out][k], k over (re|im, block
//     value): tcgen05.mma.cta_group::1.kind::f16, M = 128 (tile columns), N = 128, K = 16 per instruction, operands in
//     shared memory (K-major, no swizzle, 8x16B core matrices), accumulator in TMEM (128 lanes x 128 fp32 columns).
//   * fp32 accuracy from bf16 tensor cores by splitting every operand in three bf16 terms (hi + mid + lo = 24 mantissa
//     bits) and accumulating the six products of order <= 2 in the fp32 accumulator: 48 MMAs per tile.
//   * columns map to TMEM lanes, so that the 32 lanes of a warp read 32 consecutive amplitudes from global memory
//     (one 256-byte segment per LDG.64) and the epilogue (tcgen05.ld -> float2 STG) writes them back the same way:
//     the tile never needs an fp32 staging buffer, shared memory holds only the bf16 operand terms (96 KB + 96 KB).
//   * persistent CTA per SM; the next tile's amplitudes are prefetched into registers while the tensor core works.
// complex64 only: there is no fp64 tensor path for this (complex128 stays on the CUDA-core sweep).
#ifndef ROCQ_PRECISION_DOUBLE
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sv_internal.h"

namespace {

constexpr int BT = 512;                        // threads per CTA: 4 warps per TMEM lane quarter, 16 block values per thread
constexpr uint32_t TERM_BYTES = 32768;         // one bf16 term of a 128 x 128 operand
constexpr uint32_t SMEM_U = 0, SMEM_X = 3 * TERM_BYTES, SMEM_TAB = 6 * TERM_BYTES;   // + 64 u64 block offsets
constexpr uint32_t SMEM_NORM = SMEM_TAB + 64 * 8;                 // 2 x 4 x 128 floats: column norms in / out
constexpr uint32_t SMEM_BYTES = SMEM_NORM + 8 * 128 * 4 + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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

// shared-memory matrix descriptor: K-major, no swizzle.  Core matrix = 8 rows x 16 bytes, stored contiguously (128 B);
// LBO = byte distance between the two K-halves of one K=16 instruction, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor (kind::f16): D = F32, A = B = BF16, both K-major, N = 128, M = 128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}" ::"r"(tmem_d), "l"(a), "l"(b), "r"(IDESC), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// x = hi + mid + lo with three bf16 values (round to nearest each time; the residuals are exact in fp32)
__device__ __forceinline__ void split3(float x, float& hi, float& mid, float& lo) {
    hi = __bfloat162float(__float2bfloat16_rn(x));
    const float r = x - hi;
    mid = __bfloat162float(__float2bfloat16_rn(r));
    lo = (r - mid);
}
__device__ __forceinline__ uint32_t bf2(float a, float b) {           // a -> low half, b -> high half
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}

__device__ __forceinline__ uint64_t tile_base(uint64_t tile, const rq_block_params& P, uint64_t& member) {
    member = tile >> (P.n - P.T);
    uint64_t base = tile & ((1ull << (P.n - P.T)) - 1ull);
    for (uint32_t j = 0; j < P.T; ++j) {
        const uint32_t p = P.res[j];
        base = ((base >> p) << (p + 1)) | (base & ((1ull << p) - 1ull));
    }
    return base;
}

__global__ void __launch_bounds__(BT, 1) block_sweep_kernel(float2* __restrict__ state, const unsigned char* __restrict__ uterms,
                                                             const __grid_constant__ rq_block_params P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar_u, bar_mma;
    __shared__ uint32_t tmem_slot;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* boff = reinterpret_cast<uint64_t*>(smem + SMEM_TAB);
    float* cnorm = reinterpret_cast<float*>(smem + SMEM_NORM);          // [in|out][half][column]

    if (tid == 0) {
        mbar_init(smem_u32(&bar_u), 1);
        mbar_init(smem_u32(&bar_mma), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < 64) {                                          // global offset of block value t
        uint64_t o = 0;
        for (uint32_t b = 0; b < 6; ++b) o |= (uint64_t)((tid >> b) & 1u) << P.blk[b];
        boff[tid] = o;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (tid == 0) {                                          // the three bf16 terms of A' stay resident for all tiles
        mbar_expect_tx(smem_u32(&bar_u), 3 * TERM_BYTES);
        for (uint32_t j = 0; j < 3; ++j) bulk_g2s(smem_u32(smem + SMEM_U + j * TERM_BYTES), uterms + (size_t)j * TERM_BYTES, TERM_BYTES, smem_u32(&bar_u));
    }

    // this thread: tile column n (= TMEM lane), block values [16*qt, 16*qt+16)
    const uint32_t ncol = tid & 127, qt = tid >> 7;
    uint64_t coff = 0;
    for (uint32_t b = 0; b < 7; ++b) coff |= (uint64_t)((ncol >> b) & 1u) << P.col[b];
    const uint32_t xrow = (ncol & 7u) * 16u + (ncol >> 3) * 2048u;          // byte offset of row `ncol` in an operand term

    float2 raw[16];
    uint64_t tile = blockIdx.x, member;
    if (tile < P.ntiles) {
        const uint64_t base = tile_base(tile, P, member);
        const float2* g = state + (member << P.n) + base + coff;
#pragma unroll
        for (int j = 0; j < 16; ++j) raw[j] = g[boff[16 * qt + j]];
    }
    mbar_wait(smem_u32(&bar_u), 0);

    uint32_t phase = 0;
    for (; tile < P.ntiles; tile += gridDim.x) {
        const uint64_t base = tile_base(tile, P, member);
        float2* gt = state + (member << P.n) + base + coff;

        // ---- convert: 8 block values at a time -> one 16-byte row chunk per (term, re|im) ----
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (P.pad & 2u) break;
            float h[16], m[16], l[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                split3(raw[8 * c + j].x, h[j], m[j], l[j]);
                split3(raw[8 * c + j].y, h[8 + j], m[8 + j], l[8 + j]);
            }
            const uint32_t kg_re = 2 * qt + c, kg_im = 8 + 2 * qt + c;          // K group (8 values) of the re / im block
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                const uint32_t o = xrow + (part ? kg_im : kg_re) * 128u;
                const int s = 8 * part;
                *reinterpret_cast<uint4*>(smem + SMEM_X + 0 * TERM_BYTES + o) = make_uint4(bf2(h[s], h[s + 1]), bf2(h[s + 2], h[s + 3]), bf2(h[s + 4], h[s + 5]), bf2(h[s + 6], h[s + 7]));
                *reinterpret_cast<uint4*>(smem + SMEM_X + 1 * TERM_BYTES + o) = make_uint4(bf2(m[s], m[s + 1]), bf2(m[s + 2], m[s + 3]), bf2(m[s + 4], m[s + 5]), bf2(m[s + 6], m[s + 7]));
                *reinterpret_cast<uint4*>(smem + SMEM_X + 2 * TERM_BYTES + o) = make_uint4(bf2(l[s], l[s + 1]), bf2(l[s + 2], l[s + 3]), bf2(l[s + 4], l[s + 5]), bf2(l[s + 6], l[s + 7]));
            }
        }
        if (P.renorm) {                                                    // |column|^2 going in (this thread's 16 block values)
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) s = fmaf(raw[j].x, raw[j].x, fmaf(raw[j].y, raw[j].y, s));
            cnorm[qt * 128 + ncol] = s;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> visible to the tensor core
        tc_fence_before();
        __syncthreads();

        // ---- 48 MMAs: six term products x eight K steps, one issuing thread ----
        const uint32_t dbg = P.pad;        // timing experiments only (ROCQ_BLOCK_DEBUG): 1 = no MMA, 2 = no convert, 4 = no stores
        if (tid == 0 && !(dbg & 1u)) {
            tc_fence_after();
            const uint32_t xa = smem_u32(smem + SMEM_X), ua = smem_u32(smem + SMEM_U);
            // The tensor core truncates when it adds a K=16 partial sum to the fp32 accumulator, a bias that grows with the
            // number of accumulation steps.  So the dominant product hi*hi (8 steps) gets its own accumulator D0 and the five
            // correction products (2^-8 .. 2^-16 smaller) go to D1; the epilogue adds D0 + D1 in fp32.
            // (X term, U term): hi*hi | hi*mid, mid*hi, mid*mid, hi*lo, lo*hi
            const uint32_t tx[6] = {0, 0, 1, 1, 0, 2}, tu[6] = {0, 1, 0, 1, 2, 0};
#pragma unroll
            for (int p = 0; p < 6; ++p)
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma(tmem_d + (p == 0 ? 0u : 128u), umma_desc(xa + tx[p] * TERM_BYTES + ks * 256u, 128u, 2048u),
                         umma_desc(ua + tu[p] * TERM_BYTES + ks * 256u, 128u, 2048u), (p == 0 || p == 1) ? (ks != 0) : 1u);
            umma_commit(smem_u32(&bar_mma));
        }

        // ---- prefetch the next tile while the tensor core runs ----
        const uint64_t next = tile + gridDim.x;
        if (next < P.ntiles) {
            uint64_t nm;
            const uint64_t nb = tile_base(next, P, nm);
            const float2* g = state + (nm << P.n) + nb + coff;
#pragma unroll
            for (int j = 0; j < 16; ++j) raw[j] = g[boff[16 * qt + j]];
        }

        if (!(dbg & 1u)) {
            mbar_wait(smem_u32(&bar_mma), phase);
            phase ^= 1;
        }
        tc_fence_after();

        // ---- epilogue: TMEM lane = column; this thread's 16 block values: re at column t, im at column 64 + t ----
        float2 out[16];
        {
            uint32_t re0[16], im0[16], re1[16], im1[16];
            const uint32_t taddr = tmem_d + (((warp & 3u) * 32u) << 16) + 16u * qt;
            tmem_ld16(taddr, re0);
            tmem_ld16(taddr + 64u, im0);
            tmem_ld16(taddr + 128u, re1);
            tmem_ld16(taddr + 192u, im1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j)
                out[j] = make_float2(__uint_as_float(re0[j]) + __uint_as_float(re1[j]), __uint_as_float(im0[j]) + __uint_as_float(im1[j]));
        }
        float scale = 1.f;
        if (P.renorm) {
            // A unitary block preserves the norm of every tile column (it only mixes the 64 block values of a column).
            // Restoring it removes the systematic shrink of the tensor core's truncating accumulation, which would
            // otherwise grow linearly with the number of sweeps.
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) s = fmaf(out[j].x, out[j].x, fmaf(out[j].y, out[j].y, s));
            cnorm[512 + qt * 128 + ncol] = s;
            __syncthreads();
            const float sin = (cnorm[ncol] + cnorm[128 + ncol]) + (cnorm[256 + ncol] + cnorm[384 + ncol]);
            const float sout = (cnorm[512 + ncol] + cnorm[640 + ncol]) + (cnorm[768 + ncol] + cnorm[896 + ncol]);
            if (sout > 0.f && sin > 0.f) scale = sqrtf(sin / sout);
        }
        if (!(dbg & 4u)) {
#pragma unroll
            for (int j = 0; j < 16; ++j) gt[boff[16 * qt + j]] = make_float2(out[j].x * scale, out[j].y * scale);
        }
        tc_fence_before();
        __syncthreads();            // D and the X terms are free again
    }

    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256u) : "memory");
}

}  // namespace

extern "C" int rq_block_configure(void) {
    return (int)cudaFuncSetAttribute(block_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
}

extern "C" int rq_launch_block_sweep(rq_cplx* state, const rq_block_params* P, const void* d_uterms, void* stream) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)(P->ntiles < (uint64_t)sms ? P->ntiles : (uint64_t)sms);
    block_sweep_kernel<<<grid, BT, SMEM_BYTES, (cudaStream_t)stream>>>(reinterpret_cast<float2*>(state),
                                                                       reinterpret_cast<const unsigned char*>(d_uterms), *P);
    return (int)cudaGetLastError();
}
#else
#include "sv_internal.h"
extern "C" int rq_block_configure(void) { return 0; }
extern "C" int rq_launch_block_sweep(rq_cplx*, const rq_block_params*, const void*, void*) { return 801; /* cudaErrorNotSupported */ }
#endif
