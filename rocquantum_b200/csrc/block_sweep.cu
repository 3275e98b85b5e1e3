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
//   * fp32-class accuracy from fp16 tensor-core inputs: amplitudes are scaled by a power of two (|amp| <= 1 -> the fp16
//     normal range), then both operands are split in two fp16 terms (hi + lo = 22 mantissa bits) and the three products
//     of order <= 1 are accumulated in fp32: 24 MMAs per tile (the lo*lo product is 2^-22 relative and dropped).
//   * the tensor core truncates when it adds a K = 16 partial sum into the accumulator; the correction products are
//     therefore accumulated first and the dominant hi*hi product last, and for a unitary block the epilogue restores the
//     norm of every tile column (which the block preserves exactly).
//   * columns map to TMEM lanes, so that the 32 lanes of a warp read 32 consecutive amplitudes from global memory
//     (one 256-byte segment per LDG.64) and the epilogue (tcgen05.ld -> float2 STG) writes them back the same way:
//     the tile never needs an fp32 staging buffer; shared memory holds the fp16 operand terms only.
//   * software pipeline in a persistent CTA per SM: operand buffers and accumulators are double-buffered, so that the
//     MMAs of tile i run while the CUDA cores do the epilogue of tile i-1 and the split of tile i+1; global loads run
//     two tiles ahead (128 KB in flight per SM).
// complex64 only: there is no fp64 tensor path for this (complex128 stays on the CUDA-core sweep).
#ifndef ROCQ_PRECISION_DOUBLE
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "sv_internal.h"

namespace {

constexpr int BT = 512;                        // threads per CTA: 4 warps per TMEM lane quarter, 16 block values per thread
constexpr uint32_t TERM_BYTES = 32768;         // one fp16 term of a 128 x 128 operand
constexpr uint32_t TILE_BYTES = 65536;         // 2^13 complex64 amplitudes
constexpr uint32_t SMEM_U = 0;                                    // 2 terms of A' (the B operand of the MMA)
constexpr uint32_t SMEM_S = 2 * TERM_BYTES;                       // 2 staging tiles (fp32 amplitudes, TMA destination)
constexpr uint32_t SMEM_TAB = SMEM_S + 2 * TILE_BYTES;            // 64 u64 global block offsets
constexpr uint32_t SMEM_ROW = SMEM_TAB + 64 * 8;                  // <= 256 u64 global row offsets
constexpr uint32_t SMEM_NIN = SMEM_ROW + 256 * 8;                 // column norms going in: [buffer][quarter][column]
constexpr uint32_t SMEM_NOUT = SMEM_NIN + 2 * 4 * 128 * 4;        // column norms coming out: [quarter][column]
constexpr uint32_t SMEM_BYTES = SMEM_NOUT + 4 * 128 * 4 + 64;
// TMEM columns, per pipeline buffer b (at 256 * b): [0,128) accumulator D, [128,192) hi term of X', [192,256) lo term of X'
constexpr uint32_t TM_D = 0, TM_XH = 128, TM_XL = 192, TM_BUF = 256;

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
// instruction descriptor (kind::f16): D = F32, A = B = F16, both K-major, N = 128, M = 128
constexpr uint32_t IDESC = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}" ::"r"(tmem_d), "l"(a), "l"(b), "r"(IDESC), "r"(accumulate), "r"(0u) : "memory");
}
// A operand (X') from tensor memory, B operand (A' = block matrix) from shared memory
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(b), "r"(IDESC), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// (a, b) scaled -> hi and lo fp16 pairs: a*s = hi + lo up to 2^-22 relative (the residual is exact in fp32)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
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
    __shared__ __align__(8) uint64_t bar_u, bar_mma[2], bar_full[2];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t tbase[8];      // ring: amplitude offset of the CTA's tile i at [i & 7], computed by one thread per tile
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* boff = reinterpret_cast<uint64_t*>(smem + SMEM_TAB);
    uint64_t* rowoff = reinterpret_cast<uint64_t*>(smem + SMEM_ROW);
    float* nin = reinterpret_cast<float*>(smem + SMEM_NIN);
    float* nout = reinterpret_cast<float*>(smem + SMEM_NOUT);
    const uint32_t rowbits = P.rowbits, nrows = 1u << (13u - rowbits), rowbytes = 8u << rowbits;

    if (tid == 0) {
        mbar_init(smem_u32(&bar_u), 1);
        mbar_init(smem_u32(&bar_mma[0]), 1);
        mbar_init(smem_u32(&bar_mma[1]), 1);
        mbar_init(smem_u32(&bar_full[0]), 1);
        mbar_init(smem_u32(&bar_full[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                                         // all 512 columns: two buffers x (D | X' hi | X' lo)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < 64) {                                          // global offset of block value t
        uint64_t o = 0;
        for (uint32_t b = 0; b < 6; ++b) o |= (uint64_t)((tid >> b) & 1u) << P.blk[b];
        boff[tid] = o;
    }
    if (tid >= 64 && tid < 64 + nrows) {                     // global offset of staging row r: its bits go to the resident positions above the row
        const uint32_t r = tid - 64;
        uint64_t o = 0;
        for (uint32_t j = rowbits; j < 13; ++j) o |= (uint64_t)((r >> (j - rowbits)) & 1u) << P.res[j];
        rowoff[r] = o;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (tid == 0) {                                          // the two fp16 terms of A' stay resident for all tiles
        mbar_expect_tx(smem_u32(&bar_u), 2 * TERM_BYTES);
        for (uint32_t j = 0; j < 2; ++j) bulk_g2s(smem_u32(smem + SMEM_U + j * TERM_BYTES), uterms + (size_t)j * TERM_BYTES, TERM_BYTES, smem_u32(&bar_u));
    }

    // this thread: tile column ncol (= TMEM lane), block values [16*qt, 16*qt+16)
    const uint32_t ncol = tid & 127, qt = tid >> 7;
    uint64_t coff = 0;
    for (uint32_t b = 0; b < 7; ++b) coff |= (uint64_t)((ncol >> b) & 1u) << P.col[b];
    // the same thing inside a staging tile: local index = resident bits compacted in ascending order
    uint32_t lpos_blk[6], lcol = 0, lqt = 0;
    {
        uint32_t lpos_col[7];
        for (uint32_t j = 0, ib = 0, ic = 0; j < 13; ++j) {
            if (ib < 6 && P.res[j] == P.blk[ib]) lpos_blk[ib++] = j;
            else lpos_col[ic++] = j;
        }
        for (uint32_t b = 0; b < 7; ++b) lcol |= ((ncol >> b) & 1u) << lpos_col[b];
        lqt = ((qt & 1u) << lpos_blk[4]) | ((qt >> 1) << lpos_blk[5]);
    }
    const uint32_t s0 = 8u << lpos_blk[0], s1 = 8u << lpos_blk[1], s2 = 8u << lpos_blk[2], s3 = 8u << lpos_blk[3];   // byte strides of value bits 0..3
    const uint32_t sbase = (lcol | lqt) * 8u;
    const uint32_t dbg = P.pad;        // timing experiments only (ROCQ_BLOCK_DEBUG): 1 = no MMA, 2 = no split, 4 = no stores, 8 = no renorm
    const float scale = P.scale, inv_scale = 1.f / P.scale;
    const uint64_t first = blockIdx.x, stride = gridDim.x;
    const uint64_t cnt = P.ntiles > first ? (P.ntiles - first + stride - 1) / stride : 0;     // tiles of this CTA
    const uint64_t* off = boff + 16 * qt;                                   // block-value offsets of this thread (shared memory)
    const uint32_t tlane = tmem_d + (((warp & 3u) * 32u) << 16);            // this warp's TMEM lane quarter

    // phase timers (ROCQ_BLOCK_DEBUG & 16): threads 0 and 64 of CTA 0 accumulate clock deltas between marks
    long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = 0;
    const bool timed = (dbg & 16u) && blockIdx.x == 0 && (tid == 0 || tid == 64);
    auto mark = [&](int k) {
        if (timed) {
            const long long t = clock64();
            tacc[k] += t - tlast;
            tlast = t;
        }
    };
    if (timed) tlast = clock64();

    // the index deposit costs ~130 instructions: one thread does it per tile, ahead of time, and publishes the result
    auto publish_tile = [&](uint64_t i) {
        if (i < cnt) {
            uint64_t member;
            const uint64_t base = tile_base(first + i * stride, P, member);
            tbase[i & 7u] = (member << P.n) + base;
        }
    };
    // bulk-copy the rows of tile i into staging buffer i & 1 (tbase[i & 7] must be visible).  A bulk copy is issued from
    // the uniform datapath, i.e. one at a time per warp (~60 clocks each): every warp issues its share of the rows.
    auto load_tile = [&](uint64_t i) {
        if (i < cnt && lane == 0) {
            const uint32_t bar = smem_u32(&bar_full[i & 1u]);
            if (warp == 0) mbar_expect_tx(bar, TILE_BYTES);
            const float2* g = state + tbase[i & 7u];
            const uint32_t dst = smem_u32(smem + SMEM_S + (uint32_t)(i & 1u) * TILE_BYTES);
            for (uint32_t r = warp; r < nrows; r += BT / 32) bulk_g2s(dst + r * rowbytes, g + rowoff[r], rowbytes, bar);
        }
    };

    // ---- epilogue of tile i from pipeline buffer B: TMEM lane = column; re at column t, im at column 64 + t ----
    auto epilogue = [&](auto BC, uint64_t i) {
        constexpr uint32_t B = decltype(BC)::value;
        mark(4);
        if (!(dbg & 1u)) mbar_wait(smem_u32(&bar_mma[B]), (uint32_t)(i >> 1) & 1u);
        tc_fence_after();
        mark(5);
        // (TMEM reads run at 64 B/clk per SM, so the accumulator is read exactly once)
        const uint32_t taddr = tlane + B * TM_BUF + TM_D + 16u * qt;
        uint32_t re[16], im[16];
        tmem_ld16(taddr, re);
        tmem_ld16(taddr + 64u, im);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        mark(6);
        float f = inv_scale;
        if (P.renorm) {
            // A unitary block preserves the norm of every tile column (it only mixes the 64 block values of a column).
            // Restoring it removes the systematic shrink of the tensor core's truncating accumulation, which would
            // otherwise grow linearly with the number of sweeps.
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) s = fmaf(__uint_as_float(re[j]), __uint_as_float(re[j]), fmaf(__uint_as_float(im[j]), __uint_as_float(im[j]), s));
            nout[qt * 128 + ncol] = s;
            const float* ni = nin + B * 512;
            const float sin = (ni[ncol] + ni[128 + ncol]) + (ni[256 + ncol] + ni[384 + ncol]);
            __syncthreads();
            const float sout = (nout[ncol] + nout[128 + ncol]) + (nout[256 + ncol] + nout[384 + ncol]);   // of the scaled outputs
            if (sout > 0.f && sin > 0.f) f = sqrtf(sin / sout);
        }
        mark(7);
        if (!(dbg & 4u)) {
            float2* gt = state + tbase[i & 7u] + coff;
#pragma unroll
            for (int j = 0; j < 16; ++j) gt[off[j]] = make_float2(__uint_as_float(re[j]) * f, __uint_as_float(im[j]) * f);
        }
        tc_fence_before();
        mark(8);
    };

    // ---- one pipeline step: split tile i (staging buffer B) into TMEM, start its MMAs, refill the staging buffer with tile i+2,
    //      finish tile i-1 ----
    auto step = [&](auto BC, uint64_t i) {
        constexpr uint32_t B = decltype(BC)::value;
        mbar_wait(smem_u32(&bar_full[B]), (uint32_t)(i >> 1) & 1u);
        mark(0);
        const unsigned char* S = smem + SMEM_S + B * TILE_BYTES + sbase;
        float sin = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {                                      // 8 block values -> 4 packed words per (term, re|im)
            uint32_t hr[4], lr[4], hi[4], li[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int v0 = 8 * c + 2 * j, v1 = v0 + 1;
                const float2 a0 = *reinterpret_cast<const float2*>(S + ((v0 & 1) ? s0 : 0u) + ((v0 & 2) ? s1 : 0u) + ((v0 & 4) ? s2 : 0u) + ((v0 & 8) ? s3 : 0u));
                const float2 a1 = *reinterpret_cast<const float2*>(S + ((v1 & 1) ? s0 : 0u) + ((v1 & 2) ? s1 : 0u) + ((v1 & 4) ? s2 : 0u) + ((v1 & 8) ? s3 : 0u));
                sin = fmaf(a0.x, a0.x, fmaf(a0.y, a0.y, fmaf(a1.x, a1.x, fmaf(a1.y, a1.y, sin))));
                split2(a0.x * scale, a1.x * scale, hr[j], lr[j]);
                split2(a0.y * scale, a1.y * scale, hi[j], li[j]);
            }
            // packed words [8*qt + 4*c, +4) of the re half and [32 + 8*qt + 4*c, +4) of the im half of this thread's TMEM lane
            const uint32_t wre = 8u * qt + 4u * c, wim = 32u + 8u * qt + 4u * c;
            if (!(dbg & 2u)) {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tlane + B * TM_BUF + TM_XH + wre), "r"(hr[0]), "r"(hr[1]), "r"(hr[2]), "r"(hr[3]) : "memory");
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tlane + B * TM_BUF + TM_XL + wre), "r"(lr[0]), "r"(lr[1]), "r"(lr[2]), "r"(lr[3]) : "memory");
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tlane + B * TM_BUF + TM_XH + wim), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tlane + B * TM_BUF + TM_XL + wim), "r"(li[0]), "r"(li[1]), "r"(li[2]), "r"(li[3]) : "memory");
            }
        }
        if (P.renorm) nin[B * 512 + qt * 128 + ncol] = sin;                // |column|^2 going in (this thread's 16 block values)
        if (tid == 32) publish_tile(i + 3);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        mark(1);
        tc_fence_before();
        __syncthreads();                                                   // staging buffer B is consumed, X' of tile i is in TMEM
        mark(2);

        if (warp == 0) {
            if (lane == 0 && !(dbg & 1u)) {                                // 24 MMAs, one issuing thread
                tc_fence_after();
                const uint32_t uh = smem_u32(smem + SMEM_U), ul = uh + TERM_BYTES;
                const uint32_t d = tmem_d + B * TM_BUF + TM_D, xh = tmem_d + B * TM_BUF + TM_XH, xl = tmem_d + B * TM_BUF + TM_XL;
                // The tensor core truncates when it adds a K = 16 partial sum into the fp32 accumulator.  The two correction
                // products (2^-11 smaller) therefore go first, while the accumulator is small, and the dominant hi*hi product
                // last: only its eight additions truncate at full magnitude.
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) umma_ts(d, xh + 8u * ks, umma_desc(ul + ks * 256u, 128u, 2048u), ks != 0);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) umma_ts(d, xl + 8u * ks, umma_desc(uh + ks * 256u, 128u, 2048u), 1u);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) umma_ts(d, xh + 8u * ks, umma_desc(uh + ks * 256u, 128u, 2048u), 1u);
                umma_commit(smem_u32(&bar_mma[B]));
            }
        }
        load_tile(i + 2);                                                  // two tiles ahead, into the buffer just consumed
        mark(3);
        if (i >= 1) epilogue(std::integral_constant<uint32_t, B ^ 1u>{}, i - 1);
    };

    if (tid == 32) { publish_tile(0); publish_tile(1); publish_tile(2); }
    __syncthreads();
    load_tile(0);
    load_tile(1);
    mbar_wait(smem_u32(&bar_u), 0);

    for (uint64_t i = 0; i < cnt; i += 2) {
        step(std::integral_constant<uint32_t, 0u>{}, i);
        if (i + 1 < cnt) step(std::integral_constant<uint32_t, 1u>{}, i + 1);
    }
    if (cnt > 0) {
        if ((cnt - 1) & 1u) epilogue(std::integral_constant<uint32_t, 1u>{}, cnt - 1);
        else epilogue(std::integral_constant<uint32_t, 0u>{}, cnt - 1);
    }

    if (timed) {
        long long* out = reinterpret_cast<long long*>(const_cast<unsigned char*>(uterms) + 2 * TERM_BYTES) + (tid ? 16 : 0);
        for (int k = 0; k < 10; ++k) out[k] = tacc[k];
        out[10] = (long long)cnt;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
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
