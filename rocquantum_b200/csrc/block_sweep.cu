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
//   * fp32-class accuracy from fp16 tensor-core inputs: every tile COLUMN (the 64 amplitudes one block matrix mixes) is
//     scaled by its own power of two so that its largest component lands in [2^14, 2^15) -- whatever the magnitude of the
//     state there, peaked or unnormalised -- then both operands are split in two fp16 terms (hi + lo = 22 mantissa
//     bits) and the three products of order <= 1 are accumulated in fp32: 48 MMAs per tile (lo*lo is 2^-22 relative and
//     dropped).  Components more than 2^-39 below their column's maximum are lost, which is beyond fp32 resolution anyway.
//   * the tensor core truncates when it adds a K = 16 partial sum into the accumulator; the correction products are
//     therefore accumulated first and the dominant hi*hi product last, and for a unitary block the epilogue restores the
//     norm of every tile column (which the block preserves exactly).
//   * columns map to TMEM lanes, so that the 32 lanes of a warp read 32 consecutive amplitudes from global memory
//     (one 256-byte segment per LDG.64) and the epilogue (tcgen05.ld -> float2 STG) writes them back the same way:
//     the tile never needs an fp32 staging buffer; shared memory holds the fp16 operand terms only.
//   * software pipeline in a persistent CTA per SM: operand buffers and accumulators are double-buffered, so that the
//     MMAs of tile i run while the CUDA cores do the epilogue of tile i-1 and the split of tile i+1.
//   * shared memory = two INPUT tiles + one OUTPUT tile.  The fp32 amplitudes of a tile are dead as soon as their fp16
//     terms sit in tensor memory, so an input buffer is handed back to the loader right after the split (not after the
//     epilogue two steps later): one or two tile loads (64-128 KB per SM) are in flight at any time.  A loader thread,
//     a storer thread and the MMA thread each run their own loop in a warp of their own; the 16 worker warps meet them
//     -- and each other -- only through mbarriers (full / empty per input buffer, X' / D per pipeline buffer, written /
//     drained for the output buffer), never through a CTA-wide barrier, so the warps drift apart and their LDS, TMEM and
//     ALU phases overlap.
// complex64 only: there is no fp64 tensor path for this (complex128 stays on the CUDA-core sweep).
#ifndef ROCQ_PRECISION_DOUBLE
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "sv_internal.h"

namespace {

constexpr int NW = 512;                        // worker threads: 4 warps per TMEM lane quarter, 16 block values per thread
constexpr int BT = NW + 96;                    // + three single-thread roles in warps of their own: MMA issue, tile loads, tile stores
constexpr uint32_t NORM_EVERY = 4;             // unitary blocks: the norm ratio is measured on every 4th tile of a CTA
constexpr uint32_t MAT_BYTES = 8192;           // one fp16 term of a 64 x 64 operand (Re U or Im U)
constexpr uint32_t TILE_BYTES = 65536;         // 2^13 complex64 amplitudes
constexpr uint32_t SMEM_U = 0;                                    // Re U hi | Re U lo | Im U hi | Im U lo  (B operands, N = 64)
constexpr uint32_t SMEM_IN = 4 * MAT_BYTES;                       // 2 input tiles (fp32 amplitudes), loaded by tensor-map bulk copies
constexpr uint32_t SMEM_OUT = SMEM_IN + 2 * TILE_BYTES;           // 1 output tile in the same layout, stored by a tensor-map bulk copy
constexpr uint32_t SMEM_BYTES = SMEM_OUT + TILE_BYTES;
// TMEM columns, per pipeline buffer b (at 256 * b): [0,64) Re D, [64,128) Im D, then the packed fp16 pairs of X':
// [128,160) Re hi, [160,192) Im hi, [192,224) Re lo, [224,256) Im lo
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
// instruction descriptor (kind::f16): D = F32, A = B = F16, both K-major, N = 64, M = 128; bit 14 negates B
constexpr uint32_t IDESC = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t IDESC_NEGB = IDESC | (1u << 14);

// A operand (X') from tensor memory, B operand (A' = block matrix) from shared memory
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// whole-tile copies through a tensor map (rank 2..5; coordinates in elements of 8 bytes, innermost first)
__device__ __forceinline__ void tensor_g2s(uint32_t dst, const CUtensorMap* tm, const int32_t (&c)[5], uint32_t rank, uint32_t bar) {
    const uint64_t t = reinterpret_cast<uint64_t>(tm);
    if (rank == 2) asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                ::"r"(dst), "l"(t), "r"(bar), "r"(c[0]), "r"(c[1]) : "memory");
    else if (rank == 3) asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                     ::"r"(dst), "l"(t), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory");
    else if (rank == 4) asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                     ::"r"(dst), "l"(t), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory");
    else asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                      ::"r"(dst), "l"(t), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
}
__device__ __forceinline__ void tensor_s2g(const CUtensorMap* tm, const int32_t (&c)[5], uint32_t rank, uint32_t src) {
    const uint64_t t = reinterpret_cast<uint64_t>(tm);
    if (rank == 2) asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(t), "r"(src), "r"(c[0]), "r"(c[1]) : "memory");
    else if (rank == 3) asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(t), "r"(src), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory");
    else if (rank == 4) asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(t), "r"(src), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory");
    else asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(t), "r"(src), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

__global__ void __launch_bounds__(BT, 1) block_sweep_kernel(float2* __restrict__ state, const unsigned char* __restrict__ uterms,
                                                             const __grid_constant__ rq_block_params P,
                                                             const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // bar_full / bar_empty: input tile buffer loaded / split by all worker warps; bar_x / bar_mma: X' of a pipeline buffer in
    // tensor memory / its products accumulated; bar_owritten / bar_odrained: output tile written by all worker warps / read by its store
    __shared__ __align__(8) uint64_t bar_u, bar_mma[2], bar_x[2], bar_full[2], bar_empty[2], bar_owritten, bar_odrained, bar_scale[2][4];
    __shared__ uint32_t tmem_slot;
    __shared__ float2 red[2][16];      // per-warp (|in|^2, |out|^2) of a tile, double-buffered
    __shared__ uint8_t cexp[2][4][128]; // biased exponent of max |component| per (tile parity, quarter of the block values, column)
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(smem_u32(&bar_u), 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bar_mma[b]), 1);
            mbar_init(smem_u32(&bar_x[b]), NW / 32);
            mbar_init(smem_u32(&bar_full[b]), 1);
            mbar_init(smem_u32(&bar_empty[b]), NW / 32);
        }
        mbar_init(smem_u32(&bar_owritten), NW / 32);
        mbar_init(smem_u32(&bar_odrained), 1);
        for (int b = 0; b < 8; ++b) mbar_init(smem_u32(&bar_scale[b >> 2][b & 3]), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                                         // all 512 columns: two buffers x (D | X' hi | X' lo)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < 32) reinterpret_cast<float2*>(red)[tid] = make_float2(0.f, 0.f);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t dbg = P.pad;        // timing experiments only (ROCQ_BLOCK_DEBUG): 1 = no MMA, 2 = no split, 4 = no stores, 8 = no renorm
    const uint64_t first = blockIdx.x, stride = gridDim.x;
    const uint64_t cnt = P.ntiles > first ? (P.ntiles - first + stride - 1) / stride : 0;     // tiles of this CTA

    // A tile is one box of a tensor map over the state: one bulk-tensor instruction per tile and direction.
    const uint32_t trank = P.trank;
    auto tile_coords = [&](uint64_t i, int32_t (&c)[5]) {
        uint64_t t = first + i * stride;                                   // tile index: non-resident bits, compacted, member on top
#pragma unroll
        for (uint32_t d = 0; d < 5; ++d) {
            const uint32_t len = P.tbits[d];
            if (len == 0) c[d] = 0;
            else if (len == 255u) c[d] = (int32_t)t;
            else { c[d] = (int32_t)(t & ((1ull << len) - 1ull)); t >>= len; }
        }
    };

    if (warp == NW / 32) {
        // ================================ the MMA thread ================================
        if (lane == 0) {
            mbar_expect_tx(smem_u32(&bar_u), 4 * MAT_BYTES);          // the block matrix stays resident for all tiles
            bulk_g2s(smem_u32(smem + SMEM_U), uterms, 4 * MAT_BYTES, smem_u32(&bar_u));
            mbar_wait(smem_u32(&bar_u), 0);
            const uint32_t urh = smem_u32(smem + SMEM_U), url = urh + MAT_BYTES, uih = urh + 2 * MAT_BYTES, uil = urh + 3 * MAT_BYTES;
            for (uint64_t i = 0; i < cnt && !(dbg & 1u); ++i) {
                const uint32_t b = (uint32_t)i & 1u;
                mbar_wait(smem_u32(&bar_x[b]), (uint32_t)(i >> 1) & 1u);       // X' of tile i is in tensor memory
                tc_fence_after();
                const uint32_t dre = tmem_d + b * TM_BUF + TM_D, dim = dre + 64u;
                const uint32_t xrh = tmem_d + b * TM_BUF + TM_XH, xih = xrh + 32u, xrl = tmem_d + b * TM_BUF + TM_XL, xil = xrl + 32u;
                // Re D = Xr Ur^T - Xi Ui^T, Im D = Xr Ui^T + Xi Ur^T, each as three fp16 products (hi*lo, lo*hi, hi*hi).
                // The tensor core truncates when it adds a K = 16 partial sum into the fp32 accumulator, so the two correction
                // products (2^-11 smaller) go first, while the accumulator is small, and the dominant hi*hi product last.
                auto product = [&](uint32_t xr, uint32_t xi, uint32_t ur, uint32_t ui, bool first_product) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t dur = umma_desc(ur + ks * 256u, 128u, 1024u), dui = umma_desc(ui + ks * 256u, 128u, 1024u);
                        umma_ts(dre, xr + 8u * ks, dur, IDESC, !(first_product && ks == 0));
                        umma_ts(dim, xr + 8u * ks, dui, IDESC, !(first_product && ks == 0));
                        umma_ts(dre, xi + 8u * ks, dui, IDESC_NEGB, 1u);
                        umma_ts(dim, xi + 8u * ks, dur, IDESC, 1u);
                    }
                };
                product(xrh, xih, url, uil, true);
                product(xrl, xil, urh, uih, false);
                product(xrh, xih, urh, uih, false);
                umma_commit(smem_u32(&bar_mma[b]));
            }
        }
    } else if (warp == NW / 32 + 1) {
        // ================================ the loader thread ================================
        // tile i goes to input buffer i & 1 as soon as every worker warp has split tile i - 2 out of it
        if (lane == 0) {
            for (uint64_t i = 0; i < cnt; ++i) {
                const uint32_t b = (uint32_t)i & 1u;
                if (i >= 2) mbar_wait(smem_u32(&bar_empty[b]), (uint32_t)((i >> 1) - 1u) & 1u);
                int32_t c[5];
                tile_coords(i, c);
                mbar_expect_tx(smem_u32(&bar_full[b]), TILE_BYTES);
                tensor_g2s(smem_u32(smem + SMEM_IN + b * TILE_BYTES), &tmap, c, trank, smem_u32(&bar_full[b]));
            }
        }
    } else if (warp == NW / 32 + 2) {
        // ================================ the storer thread ================================
        // tile j leaves the output buffer once every worker warp has written its part; the buffer is handed back when the
        // bulk store has READ it (the global writes complete behind)
        if (lane == 0) {
            for (uint64_t j = 0; j < cnt; ++j) {
                mbar_wait(smem_u32(&bar_owritten), (uint32_t)j & 1u);
                if (!(dbg & 4u)) {
                    int32_t c[5];
                    tile_coords(j, c);
                    tensor_s2g(&tmap, c, trank, smem_u32(smem + SMEM_OUT));
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                mbar_arrive(smem_u32(&bar_odrained));
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");              // the last stores must be complete before the CTA exits
        }
    } else {
        // ================================ 16 worker warps ================================
        // this thread: tile column ncol (= TMEM lane), block values [16*qt, 16*qt+16)
        const uint32_t ncol = tid & 127, qt = tid >> 7;
        // Position inside a staged tile.  The tensor map lays a tile out as 512 rows of 16 amplitudes (index bits 0-3 of the
        // state), then the other column bits, then the other block bits (P.lp_col / P.lp_blk: bit of the tile-local amplitude
        // index that each column / block bit lands on), with the 128-byte swizzle: 16-byte chunk c of row r is stored at chunk
        // c ^ (r & 7).  Local bits 4-6 are always column bits, so the swizzle is one XOR constant per thread, and lanes that
        // differ in the lowest column bits hit different banks wherever the block sits -- also on index bits 0-4.
        uint32_t lcol = 0;
        for (uint32_t b = 0; b < 7; ++b) lcol |= ((ncol >> b) & 1u) << P.lp_col[b];
        const uint32_t lbase = lcol | ((qt & 1u) << P.lp_blk[4]) | ((qt >> 1) << P.lp_blk[5]);
        const uint32_t sbase = (lbase ^ (((lbase >> 4) & 7u) << 1)) * 8u;                      // bytes, swizzled
        const uint32_t s0 = 8u << P.lp_blk[0], s1 = 8u << P.lp_blk[1], s2 = 8u << P.lp_blk[2], s3 = 8u << P.lp_blk[3];   // byte strides of value bits 0..3
        const uint32_t tlane = tmem_d + (((warp & 3u) * 32u) << 16);            // this warp's TMEM lane quarter

        // shared-memory address of this thread's block value v of a tile
        // (value bits and the thread's own bits are disjoint, and the swizzle is an XOR: offsets combine with XOR)
        auto vaddr = [&](int v) -> uint32_t { return ((v & 1) ? s0 : 0u) ^ ((v & 2) ? s1 : 0u) ^ ((v & 4) ? s2 : 0u) ^ ((v & 8) ? s3 : 0u); };

        // ---- epilogue of tile i from pipeline buffer B: TMEM lane = column; re at column t, im at column 64 + t.  The thread
        //      writes, into the output buffer, the positions it read from the input buffer. ----
        float my_in = 0.f;                                                     // |.|^2 of this thread's inputs of the tile in flight
        float my_inv = 1.f;                                                    // 1 / (scale of this thread's column) of the tile in flight
        // Column scales by guess and verify.  Every tile column (the 64 amplitudes one block matrix mixes, held by the four
        // threads with this lane in warps w, w+4, w+8, w+12) is multiplied by its own power of two before the fp16 split, so
        // that peaked or unnormalised states keep fp32-class accuracy.  The four threads must agree on it, and agreeing
        // BEFORE the split costs 14 % of the pass (a barrier between the load and the split puts the SM in lock step:
        // profiles/r02_block_scale_variants.md).  So the split runs at once with a GUESS -- the scale this column position
        // wanted in the previous tile, identical in the four threads by induction -- while the exponents of the four threads'
        // maxima travel through shared memory behind an mbarrier that is only waited for after the split.  If the guess puts
        // the column's true maximum inside [2^6, 2^15) the tensor memory already holds valid terms (>= 30 bits below the
        // column maximum are kept); otherwise (first tile of a CTA, magnitude jumps of > 2^3 up or 2^6 down) the warp splits
        // its tile again with the exact scale.
        int guess_exp = (int)(__float_as_uint(P.scale) >> 23) - 127;           // log2 of the guess
        float fcorr = 1.f;                                                     // norm correction (unitary blocks)
        float acc_in = 0.f, acc_out = 0.f;                                     // |in|^2, |scaled out|^2 over the CTA's finished tiles
        auto epilogue = [&](auto BC, uint64_t i, float in2, float inv_scale) {
            constexpr uint32_t B = decltype(BC)::value;
            if (!(dbg & 1u)) mbar_wait(smem_u32(&bar_mma[B]), (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            // (TMEM reads run at 64 B/clk per SM, so the accumulator is read exactly once)
            const uint32_t taddr = tlane + B * TM_BUF + TM_D + 16u * qt;
            uint32_t re[16], im[16];
            tmem_ld16(taddr, re);
            tmem_ld16(taddr + 64u, im);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            // the output buffer is free once the store of tile i-1 has read it.  That store was issued after EVERY worker warp
            // had finished its epilogue of tile i-1, so what they left in red[] is visible from here on as well.
            if (i > 0) mbar_wait(smem_u32(&bar_odrained), (uint32_t)(i - 1) & 1u);
            // A unitary block preserves the norm of the tile.  The tensor core's truncating accumulation shrinks it
            // systematically (~1e-7 per sweep); the ratio measured on the tiles this CTA has finished removes that bias.
            // Only every NORM_EVERY-th tile is measured: the bias is the same everywhere, and the two sums of squares plus
            // their reductions were ~12 % of the workers' instructions.
            if (P.renorm && i > 0 && ((i - 1) % NORM_EVERY) == 0) {
                float2 v = red[(i - 1) & 1u][lane & 15u];
#pragma unroll
                for (int m = 8; m >= 1; m >>= 1) {
                    v.x += __shfl_xor_sync(0xffffffffu, v.x, m);
                    v.y += __shfl_xor_sync(0xffffffffu, v.y, m);
                }
                // mass-weighted over all previous tiles of this CTA: an empty or nearly empty tile (whose ratio is rounding
                // noise) must not set the factor of a tile that carries the state's weight
                acc_in += v.x;
                acc_out += v.y;
                if (acc_in > 0.f && acc_out > 0.f) {
                    const float r = sqrtf(acc_in / acc_out);
                    if (fabsf(r - 1.f) < 1e-4f) fcorr = r;
                }
            }
            const float f = fcorr * inv_scale;
            unsigned char* S = smem + SMEM_OUT;
            float out2 = 0.f;
            const bool tracked = P.renorm && (i % NORM_EVERY) == 0;            // (uniform)
            if (tracked) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float a = __uint_as_float(re[j]), b = __uint_as_float(im[j]);
                    out2 = fmaf(a, a, fmaf(b, b, out2));
                }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float a = __uint_as_float(re[j]), b = __uint_as_float(im[j]);
                *reinterpret_cast<float2*>(S + (sbase ^ vaddr(j))) = make_float2(a * f, b * f);
            }
            if (tracked) {
                out2 *= inv_scale * inv_scale;                                 // back to the state's own units (columns differ in scale)
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) {
                    in2 += __shfl_xor_sync(0xffffffffu, in2, m);
                    out2 += __shfl_xor_sync(0xffffffffu, out2, m);
                }
                if (lane == 0) red[i & 1u][warp] = make_float2(in2, out2);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> visible to the bulk copy engine
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_owritten));               // one arrival per worker warp
        };

        // ---- one pipeline step: split tile i into tensor memory (buffer B), hand it to the MMA thread and its input buffer back to
        //      the loader, finish tile i-1 ----
        auto step = [&](auto BC, uint64_t i) {
            constexpr uint32_t B = decltype(BC)::value;                        // = i & 1: tensor-memory buffer and input tile buffer
            mbar_wait(smem_u32(&bar_full[B]), (uint32_t)(i >> 1) & 1u);
            const unsigned char* S = smem + SMEM_IN + B * TILE_BYTES;
            float in2 = 0.f;
            float mx = 0.f;
            auto split_tile = [&](int sexp, bool measure, bool first_pass) {
                const float scale = __uint_as_float((uint32_t)(sexp + 127) << 23);
#pragma unroll
                for (int c = 0; c < 2; ++c) {                                  // 8 block values -> 4 packed words per (term, re|im)
                    uint32_t hr[4], lr[4], hi[4], li[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 a0 = *reinterpret_cast<const float2*>(S + (sbase ^ vaddr(8 * c + 2 * j)));
                        const float2 a1 = *reinterpret_cast<const float2*>(S + (sbase ^ vaddr(8 * c + 2 * j + 1)));
                        if (measure) in2 = fmaf(a0.x, a0.x, fmaf(a0.y, a0.y, fmaf(a1.x, a1.x, fmaf(a1.y, a1.y, in2))));
                        if (first_pass) mx = fmaxf(fmaxf(mx, fmaxf(fabsf(a0.x), fabsf(a0.y))), fmaxf(fabsf(a1.x), fabsf(a1.y)));
                        split2(a0.x * scale, a1.x * scale, hr[j], lr[j]);
                        split2(a0.y * scale, a1.y * scale, hi[j], li[j]);
                    }
                    // packed words [8*qt + 4*c, +4) of the re part and of the im part of this thread's TMEM lane
                    const uint32_t w = 8u * qt + 4u * c, xh = tlane + B * TM_BUF + TM_XH, xl = tlane + B * TM_BUF + TM_XL;
                    if (!(dbg & 2u)) {
                        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(xh + w), "r"(hr[0]), "r"(hr[1]), "r"(hr[2]), "r"(hr[3]) : "memory");
                        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(xl + w), "r"(lr[0]), "r"(lr[1]), "r"(lr[2]), "r"(lr[3]) : "memory");
                        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(xh + 32u + w), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
                        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(xl + 32u + w), "r"(li[0]), "r"(li[1]), "r"(li[2]), "r"(li[3]) : "memory");
                    }
                }
            };
            split_tile(guess_exp, P.renorm && (i % NORM_EVERY) == 0, true);
            // publish this thread's exponent, then see what the column's four threads found
            const uint32_t sbar = smem_u32(&bar_scale[B][warp & 3u]);
            cexp[B][qt][ncol] = (uint8_t)(__float_as_uint(mx) >> 23);
            __syncwarp();
            if (lane == 0) mbar_arrive(sbar);
            mbar_wait(sbar, (uint32_t)(i >> 1) & 1u);
            const int ecol = (int)max(max((uint32_t)cexp[B][0][ncol], (uint32_t)cexp[B][1][ncol]),
                                      max((uint32_t)cexp[B][2][ncol], (uint32_t)cexp[B][3][ncol])) - 127;   // floor(log2 max)
            // want the maximum in [2^12, 2^13); keep the guess while it leaves it in [2^6, 2^16).  Columns of zeros or
            // denormals (ecol = -127) take any scale; inf / nan columns stay what they are.
            const int want_exp = min(max(12 - ecol, -100), 100);
            int use_exp = guess_exp;
            if (ecol > -127 && (ecol + guess_exp < 6 || ecol + guess_exp > 14)) use_exp = want_exp;     // (2^16 itself is past fp16)
            if (__any_sync(0xffffffffu, use_exp != guess_exp)) split_tile(use_exp, false, false);
            guess_exp = ecol > -127 ? want_exp : guess_exp;
            const float cur_inv = __uint_as_float((uint32_t)(127 - use_exp) << 23);
            // every tcgen05.st of this warp has taken its registers -- and with them the shared-memory loads that fed them --
            // before the two hand-overs: X' to the MMA thread, the input buffer back to the loader
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {                                                   // one arrival per worker warp on each
                mbar_arrive(smem_u32(&bar_x[B]));
                mbar_arrive(smem_u32(&bar_empty[B]));
            }
            const float prev_in = my_in, prev_inv = my_inv;
            my_in = in2;
            my_inv = cur_inv;
            if (i >= 1) epilogue(std::integral_constant<uint32_t, B ^ 1u>{}, i - 1, prev_in, prev_inv);
        };

        for (uint64_t i = 0; i < cnt; i += 2) {
            step(std::integral_constant<uint32_t, 0u>{}, i);
            if (i + 1 < cnt) step(std::integral_constant<uint32_t, 1u>{}, i + 1);
        }
        if (cnt > 0) {
            if ((cnt - 1) & 1u) epilogue(std::integral_constant<uint32_t, 1u>{}, cnt - 1, my_in, my_inv);
            else epilogue(std::integral_constant<uint32_t, 0u>{}, cnt - 1, my_in, my_inv);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
}

}  // namespace

extern "C" int rq_block_configure(void) {
    return (int)cudaFuncSetAttribute(block_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
}

extern "C" int rq_launch_block_sweep(rq_cplx* state, const rq_block_params* P, const void* d_uterms, const void* tensor_map, void* stream) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)(P->ntiles < (uint64_t)sms ? P->ntiles : (uint64_t)sms);
    CUtensorMap tm{};
    if (tensor_map) tm = *reinterpret_cast<const CUtensorMap*>(tensor_map);
    block_sweep_kernel<<<grid, BT, SMEM_BYTES, (cudaStream_t)stream>>>(reinterpret_cast<float2*>(state),
                                                                       reinterpret_cast<const unsigned char*>(d_uterms), *P, tm);
    return (int)cudaGetLastError();
}
#else
#include "sv_internal.h"
extern "C" int rq_block_configure(void) { return 0; }
extern "C" int rq_launch_block_sweep(rq_cplx*, const rq_block_params*, const void*, const void*, void*) { return 801; /* cudaErrorNotSupported */ }
#endif
