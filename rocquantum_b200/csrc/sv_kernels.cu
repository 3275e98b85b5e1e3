// rocquantum_b200/csrc/sv_kernels.cu -- everything on the path that is not the fused tile sweep:
//   * generic k-qubit dense matrix with controls (any k <= 10), the spec of which is the reference's
//     never-launched apply_multi_qubit_generic_matrix_kernel (multi_qubit_kernels.hip:37-115)
//   * state initialisation (hipStateVec.cpp:253-272)
//   * Pauli-string expectation, probability masses, collapse+renormalise, sampling -- the contracts of
//     hipStateVec.h:172-177, 340-445 whose only reference kernels are the unlaunched placeholders in
//     measurement_kernels.hip.  Reductions use warp shuffles + one partial per block, fp64 (or exact
//     128-bit fixed point where a decision depends on the sum) instead of the reference's real_t
//     shared-memory trees (measurement_kernels.hip:103-157).
#include <cuda_runtime.h>
#include <stdint.h>

#include "sv_internal.h"

namespace {

constexpr int RT = 256;                 // threads per block for reductions / elementwise
constexpr unsigned RBLOCKS = 148 * 8;   // one wave of 8 blocks per SM on a 148-SM B200

struct u128 { uint64_t hi, lo; };
__device__ __forceinline__ void add128(u128& a, const u128 b) {
    a.lo += b.lo;
    a.hi += b.hi + (a.lo < b.lo ? 1ull : 0ull);
}
__device__ __forceinline__ bool lt128(const u128 a, const u128 b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
__device__ __forceinline__ u128 sub128(const u128 a, const u128 b) {
    u128 r;
    r.lo = a.lo - b.lo;
    r.hi = a.hi - b.hi - (a.lo < b.lo ? 1ull : 0ull);
    return r;
}

// floor(p * 2^88) as an exact 128-bit integer -- same definition as oracle/sv_oracle.c:orc_fix88.
__device__ __forceinline__ u128 fix88(double p) {
    const uint64_t bits = (uint64_t)__double_as_longlong(p);
    u128 r = {0, 0};
    if (bits >> 63) return r;
    const unsigned e = (unsigned)((bits >> 52) & 0x7ff);
    if (e == 0 || e == 0x7ff) return r;
    const uint64_t m = (bits & 0xFFFFFFFFFFFFFull) | (1ull << 52);
    int shift = (int)e - 987;
    if (shift >= 0) {
        if (shift > 72) shift = 72;
        if (shift == 0) { r.lo = m; }
        else if (shift < 64) { r.lo = m << shift; r.hi = m >> (64 - shift); }
        else { r.hi = m << (shift - 64); }
        return r;
    }
    if (shift <= -64) return r;
    r.lo = m >> (-shift);
    return r;
}
__device__ __forceinline__ double prob(const rq_cplx a) { return fma((double)a.x, (double)a.x, (double)a.y * (double)a.y); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ u128 warp_sum128(u128 v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        u128 o;
        o.hi = __shfl_xor_sync(0xffffffffu, v.hi, d);
        o.lo = __shfl_xor_sync(0xffffffffu, v.lo, d);
        add128(v, o);
    }
    return v;
}

// ---------------------------------------------------------------------------------------------------
// generic dense matrix on k targets with a control mask.  A block stages G groups of 2^k amplitudes in
// shared memory, then each thread produces (group,row) outputs.  Matrix: device memory, column-major.
// ---------------------------------------------------------------------------------------------------
struct gather_params {
    unsigned n, k, nfix;
    unsigned tpos[10];
    unsigned char fix[80];
    uint64_t cmask;
    uint64_t ngroups_per_state;
    uint64_t ngroups_total;
};

__global__ void __launch_bounds__(RT) gather_dense_kernel(rq_cplx* __restrict__ state, const rq_cplx* __restrict__ M,
                                                          const __grid_constant__ gather_params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    rq_cplx* in = reinterpret_cast<rq_cplx*>(smem_raw);
    const unsigned D = 1u << P.k;
    const unsigned G = D >= RT ? 1u : RT / D;           // groups per block iteration
    const unsigned items = G * D;
    for (uint64_t g0 = (uint64_t)blockIdx.x * G; g0 < P.ngroups_total; g0 += (uint64_t)gridDim.x * G) {
        for (unsigned it = threadIdx.x; it < items; it += RT) {
            const unsigned gl = it / D, j = it % D;
            const uint64_t g = g0 + gl;
            if (g < P.ngroups_total) {
                uint64_t base = g % P.ngroups_per_state;
                const uint64_t member = g / P.ngroups_per_state;
                for (unsigned f = 0; f < P.nfix; ++f) {
                    const unsigned p = P.fix[f];
                    base = ((base >> p) << (p + 1)) | (base & ((1ull << p) - 1ull));
                }
                base |= P.cmask;
                uint64_t off = 0;
                for (unsigned b = 0; b < P.k; ++b) off |= (uint64_t)((j >> b) & 1u) << P.tpos[b];
                in[it] = state[(member << P.n) + base + off];
            }
        }
        __syncthreads();
        for (unsigned it = threadIdx.x; it < items; it += RT) {
            const unsigned gl = it / D, r = it % D;
            const uint64_t g = g0 + gl;
            if (g < P.ngroups_total) {
                rq_cplx acc = {0, 0};
                const rq_cplx* a = in + gl * D;
                for (unsigned j = 0; j < D; ++j) {
                    const rq_cplx m = M[r + (size_t)j * D], v = a[j];
                    acc.x += m.x * v.x - m.y * v.y;
                    acc.y += m.x * v.y + m.y * v.x;
                }
                uint64_t base = g % P.ngroups_per_state;
                const uint64_t member = g / P.ngroups_per_state;
                for (unsigned f = 0; f < P.nfix; ++f) {
                    const unsigned p = P.fix[f];
                    base = ((base >> p) << (p + 1)) | (base & ((1ull << p) - 1ull));
                }
                base |= P.cmask;
                uint64_t off = 0;
                for (unsigned b = 0; b < P.k; ++b) off |= (uint64_t)((r >> b) & 1u) << P.tpos[b];
                state[(member << P.n) + base + off] = acc;
            }
        }
        __syncthreads();
    }
}

__global__ void set_one_kernel(rq_cplx* state) {
    state[0].x = (rq_real)1;
    state[0].y = (rq_real)0;
}

// ---------------------------------------------------------------------------------------------------
// <psi|P|psi>, P = i^ny (-1)^{popcount(i & zmask)} |i^xmask><i|.  Each amplitude is read once: for
// xmask != 0 the pair (i, i^xmask) is visited from its member whose pivot bit (highest bit of xmask) is 0.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RT) pauli_expect_kernel(const rq_cplx* __restrict__ state, unsigned n, uint64_t xmask,
                                                          uint64_t zmask, unsigned ny, double* __restrict__ partials) {
    const uint64_t N = 1ull << n;
    const uint64_t stride = (uint64_t)gridDim.x * RT;
    double acc = 0.0;
    if (xmask == 0) {
        for (uint64_t i = (uint64_t)blockIdx.x * RT + threadIdx.x; i < N; i += stride) {
            const double p = prob(state[i]);
            acc += (__popcll(i & zmask) & 1) ? -p : p;
        }
    } else {
        const unsigned pv = 63u - (unsigned)__clzll((long long)xmask);
        const uint64_t low = (1ull << pv) - 1ull;
        // ny even: term = s * i^ny * 2 Re(conj(a_j) a_i);  ny odd: term = s * i^(ny+1) * 2 Im(conj(a_j) a_i)
        const double c = ((ny & 3u) == 0u || (ny & 3u) == 3u) ? 2.0 : -2.0;
        for (uint64_t h = (uint64_t)blockIdx.x * RT + threadIdx.x; h < (N >> 1); h += stride) {
            const uint64_t i = ((h & ~low) << 1) | (h & low), j = i ^ xmask;
            const rq_cplx a = state[i], b = state[j];
            const double tr = (double)b.x * a.x + (double)b.y * a.y;
            const double ti = (double)b.x * a.y - (double)b.y * a.x;
            const double v = (ny & 1u) ? ti : tr;
            acc += (__popcll(i & zmask) & 1) ? -v : v;
        }
        acc *= c;
    }
    __shared__ double wsum[RT / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < RT / 32; ++w) s += wsum[w];
        partials[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(RT) sum_partials_kernel(const double* __restrict__ partials, unsigned nblocks,
                                                          double* __restrict__ out) {
    __shared__ double wsum[RT / 32];
    double acc = 0.0;
    for (unsigned i = threadIdx.x; i < nblocks; i += RT) acc += partials[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < RT / 32; ++w) s += wsum[w];
        out[0] = s;
    }
}

// ---------------------------------------------------------------------------------------------------
// Batched expectation: every term of a group shares the x-mask, so ONE read sweep serves them all -- the pair product
// conj(a_j) a_i is formed once and each term only adds its sign (-1)^{popcount(i & zmask_t)} and picks Re or Im by the
// parity of its Y count.  All-Z Hamiltonian terms are one group (xmask = 0): one sweep however many there are.
// grid = (blocks, states of the batch); partials[(state * blocks + block) * TT + t].
// ---------------------------------------------------------------------------------------------------
template <int TT>
__global__ void __launch_bounds__(RT) pauli_group_kernel(const rq_cplx* __restrict__ state_all, unsigned n,
                                                         const __grid_constant__ rq_pauli_group G, double* __restrict__ partials) {
    const uint64_t N = 1ull << n;
    const rq_cplx* __restrict__ state = state_all + (uint64_t)blockIdx.y * N;
    const uint64_t stride = (uint64_t)gridDim.x * RT;
    double acc[TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) acc[t] = 0.0;
    // sign of term t at index i = parity of popcount(i & zmask_t): two 32-bit halves (the high one only for n > 32), and the
    // sign is applied by flipping the double's sign bit -- 4 integer instructions + 1 DADD per term and amplitude
    const bool wide = n > 32;
    auto add_signed = [&](int t, uint32_t il, uint32_t ih, double v) {
        uint32_t par = (uint32_t)__popc(il & (uint32_t)G.zmask[t]);
        if (wide) par += (uint32_t)__popc(ih & (uint32_t)(G.zmask[t] >> 32));
        acc[t] += __hiloint2double(__double2hiint(v) ^ (int)(par << 31), __double2loint(v));
    };
    if (G.xmask == 0) {
        for (uint64_t i = (uint64_t)blockIdx.x * RT + threadIdx.x; i < N; i += stride) {
            const double p = prob(state[i]);
            const uint32_t il = (uint32_t)i, ih = (uint32_t)(i >> 32);
#pragma unroll
            for (int t = 0; t < TT; ++t) add_signed(t, il, ih, p);
        }
    } else {
        const unsigned pv = 63u - (unsigned)__clzll((long long)G.xmask);
        const uint64_t low = (1ull << pv) - 1ull;
        for (uint64_t h = (uint64_t)blockIdx.x * RT + threadIdx.x; h < (N >> 1); h += stride) {
            const uint64_t i = ((h & ~low) << 1) | (h & low), j = i ^ G.xmask;
            const rq_cplx a = state[i], b = state[j];
            const double tr = (double)b.x * a.x + (double)b.y * a.y;
            const double ti = (double)b.x * a.y - (double)b.y * a.x;
            const uint32_t il = (uint32_t)i, ih = (uint32_t)(i >> 32);
#pragma unroll
            for (int t = 0; t < TT; ++t) add_signed(t, il, ih, (G.ny[t] & 1u) ? ti : tr);
        }
    }
    __shared__ double wsum[RT / 32][TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) {
        const double v = warp_sum(acc[t]);
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < TT) {
        double sum = 0.0;
        for (int w = 0; w < RT / 32; ++w) sum += wsum[w][threadIdx.x];
        partials[((uint64_t)blockIdx.y * gridDim.x + blockIdx.x) * TT + threadIdx.x] = sum;
    }
}
// The same sum for all-Z groups of several terms (the Z / ZZ terms of a Hamiltonian: one sweep for up to 16 of them), where one
// POPC -- a quarter-rate instruction -- per term and amplitude is the bound (1.46 ms for 16 terms at 28 qubits against 0.33 ms
// for the read).  A warp takes chunks of 2^10 consecutive indices i = chunk * 2^10 + j * 32 + lane, and the sign word (bit t =
// parity of popcount(i & z_t)) splits along those fields:  w = L ^ J[j] ^ C  with
//   L    the lane's part, constant per thread (TT popcounts per THREAD),
//   J[j] the part of the 32 values of j: a table in shared memory, read with a warp-uniform address that does not depend on
//        anything computed in the loop,
//   C    the chunk's part: lane t counts term t's bits once per chunk, one ballot gathers the word.
// Per term and amplitude there remain a shift, one LOP3 into the double's sign bit and the DADD (0.73 ms for 16 terms).
template <int TT>
__global__ void __launch_bounds__(RT) pauli_group_wide_kernel(const rq_cplx* __restrict__ state_all, unsigned n,
                                                              const __grid_constant__ rq_pauli_group G, double* __restrict__ partials) {
    const uint64_t N = 1ull << n, nchunks = N >> 10;          // (the launcher guarantees N >= 2^15 and an all-Z group)
    const rq_cplx* __restrict__ state = state_all + (uint64_t)blockIdx.y * N;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    __shared__ uint32_t J[32];
    if (threadIdx.x < 32) {
        uint32_t word = 0;
#pragma unroll
        for (int t = 0; t < TT; ++t) word |= (uint32_t)(__popcll(((uint64_t)threadIdx.x << 5) & G.zmask[t]) & 1) << t;
        J[threadIdx.x] = word;
    }
    uint32_t L = 0;
#pragma unroll
    for (int t = 0; t < TT; ++t) L |= (uint32_t)(__popcll((uint64_t)lane & G.zmask[t]) & 1) << t;
    const uint64_t myz = lane < (unsigned)TT ? G.zmask[lane] : 0ull;            // lane t serves term t in the ballot
    __syncthreads();
    double acc[TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) acc[t] = 0.0;
    const uint64_t gw = (uint64_t)blockIdx.x * (RT / 32) + warp, nw = (uint64_t)gridDim.x * (RT / 32);
    for (uint64_t c = gw; c < nchunks; c += nw) {
        const uint32_t LC = L ^ __ballot_sync(0xffffffffu, __popcll((c << 10) & myz) & 1);
        const rq_cplx* src = state + (c << 10) + lane;
#pragma unroll 4
        for (uint32_t j = 0; j < 32; ++j) {
            const uint32_t w = LC ^ J[j];
            const double p = prob(src[j << 5]);
#pragma unroll
            for (int t = 0; t < TT; ++t)
                acc[t] += __hiloint2double(__double2hiint(p) ^ (int)((w << (31 - t)) & 0x80000000u), __double2loint(p));
        }
    }
    __shared__ double wsum[RT / 32][TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) {
        const double v = warp_sum(acc[t]);
        if (lane == 0) wsum[warp][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < TT) {
        double sum = 0.0;
        for (int w_ = 0; w_ < RT / 32; ++w_) sum += wsum[w_][threadIdx.x];
        partials[((uint64_t)blockIdx.y * gridDim.x + blockIdx.x) * TT + threadIdx.x] = sum;
    }
}
// grid = (terms of the group, states): deterministic second stage, the i^ny coefficient, and the scatter to
// results[state * num_terms_total + G.index[t]]
__global__ void __launch_bounds__(RT) pauli_group_finish_kernel(const double* __restrict__ partials, unsigned nblocks, unsigned TT,
                                                                const __grid_constant__ rq_pauli_group G, unsigned num_terms_total,
                                                                double* __restrict__ results) {
    const unsigned t = blockIdx.x, st = blockIdx.y;
    __shared__ double wsum[RT / 32];
    double acc = 0.0;
    for (unsigned b = threadIdx.x; b < nblocks; b += RT) acc += partials[((uint64_t)st * nblocks + b) * TT + t];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double sum = 0.0;
        for (int w = 0; w < RT / 32; ++w) sum += wsum[w];
        if (G.xmask != 0) {
            const unsigned ny = G.ny[t];
            sum *= ((ny & 3u) == 0u || (ny & 3u) == 3u) ? 2.0 : -2.0;
        }
        results[(uint64_t)st * num_terms_total + G.index[t]] = sum;
    }
}

// ---------------------------------------------------------------------------------------------------
// exact probability masses of "bit q = 0" and "bit q = 1" (q >= n: everything counts as 0)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RT) fixed_masses_kernel(const rq_cplx* __restrict__ state, unsigned n, unsigned q,
                                                          uint64_t* __restrict__ partials) {
    const uint64_t N = 1ull << n;
    u128 a0 = {0, 0}, a1 = {0, 0};
    for (uint64_t i = (uint64_t)blockIdx.x * RT + threadIdx.x; i < N; i += (uint64_t)gridDim.x * RT) {
        const u128 v = fix88(prob(state[i]));
        if (q < n && ((i >> q) & 1ull)) add128(a1, v); else add128(a0, v);
    }
    a0 = warp_sum128(a0);
    a1 = warp_sum128(a1);
    __shared__ uint64_t ws[RT / 32][4];
    if ((threadIdx.x & 31) == 0) {
        ws[threadIdx.x >> 5][0] = a0.hi; ws[threadIdx.x >> 5][1] = a0.lo;
        ws[threadIdx.x >> 5][2] = a1.hi; ws[threadIdx.x >> 5][3] = a1.lo;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u128 s0 = {0, 0}, s1 = {0, 0};
        for (int w = 0; w < RT / 32; ++w) {
            add128(s0, u128{ws[w][0], ws[w][1]});
            add128(s1, u128{ws[w][2], ws[w][3]});
        }
        partials[4 * blockIdx.x + 0] = s0.hi; partials[4 * blockIdx.x + 1] = s0.lo;
        partials[4 * blockIdx.x + 2] = s1.hi; partials[4 * blockIdx.x + 3] = s1.lo;
    }
}
__global__ void sum_masses_kernel(const uint64_t* __restrict__ partials, unsigned nblocks, uint64_t* __restrict__ out4) {
    if (threadIdx.x != 0) return;
    u128 s0 = {0, 0}, s1 = {0, 0};
    for (unsigned b = 0; b < nblocks; ++b) {
        add128(s0, u128{partials[4 * b], partials[4 * b + 1]});
        add128(s1, u128{partials[4 * b + 2], partials[4 * b + 3]});
    }
    out4[0] = s0.hi; out4[1] = s0.lo; out4[2] = s1.hi; out4[3] = s1.lo;
}

// collapse (measurement_kernels.hip:37-58) and renormalise (:64-77) in one pass: the discarded half is
// only written, the kept half is read, scaled and written.
__global__ void __launch_bounds__(RT) collapse_kernel(rq_cplx* __restrict__ state, unsigned n, unsigned q, int outcome,
                                                      rq_real scale) {
    const uint64_t N = 1ull << n;
    for (uint64_t i = (uint64_t)blockIdx.x * RT + threadIdx.x; i < N; i += (uint64_t)gridDim.x * RT) {
        if ((int)((i >> q) & 1ull) != outcome) {
            state[i] = rq_cplx{0, 0};
        } else {
            rq_cplx a = state[i];
            a.x *= scale;
            a.y *= scale;
            state[i] = a;
        }
    }
}

// one warp per chunk of 2^chunk_bits amplitudes: exact mass of the chunk.  A lane moves 16 bytes per load (two complex64 / one
// complex128 amplitude) with four loads in flight, i.e. 2 KB per warp iteration: 64 resident warps keep 128 KB per SM in flight,
// enough to cover the HBM latency (the one-amplitude-per-load version ran at 0.43 of the copy peak).
struct __align__(16) amp16 { rq_cplx a[16 / sizeof(rq_cplx)]; };
__device__ __forceinline__ void add_amp16(u128& acc, const amp16& v) {
#pragma unroll
    for (unsigned e = 0; e < 16 / sizeof(rq_cplx); ++e) add128(acc, fix88(prob(v.a[e])));
}
__global__ void __launch_bounds__(RT) chunk_masses_kernel(const rq_cplx* __restrict__ state, unsigned n, unsigned chunk_bits,
                                                          uint64_t* __restrict__ hi, uint64_t* __restrict__ lo) {
    constexpr unsigned PER = 16 / sizeof(rq_cplx);                  // amplitudes per 16-byte element
    const uint64_t nchunks = 1ull << (n - chunk_bits);
    const unsigned lane = threadIdx.x & 31;
    const uint64_t warps = (uint64_t)gridDim.x * (RT / 32);
    const uint64_t len = 1ull << chunk_bits;
    for (uint64_t c = (uint64_t)blockIdx.x * (RT / 32) + (threadIdx.x >> 5); c < nchunks; c += warps) {
        const rq_cplx* p = state + (c << chunk_bits);
        u128 acc = {0, 0};
        if (len >= 128 * PER) {                                     // whole unrolled iterations only (chunks are powers of two)
            const amp16* v = reinterpret_cast<const amp16*>(p);
            const uint64_t nv = len / PER;
            for (uint64_t i = lane; i < nv; i += 128) {
                const amp16 v0 = v[i], v1 = v[i + 32], v2 = v[i + 64], v3 = v[i + 96];
                add_amp16(acc, v0); add_amp16(acc, v1); add_amp16(acc, v2); add_amp16(acc, v3);
            }
        } else {
            for (uint64_t i = lane; i < len; i += 32) add128(acc, fix88(prob(p[i])));
        }
        acc = warp_sum128(acc);
        if (lane == 0) { hi[c] = acc.hi; lo[c] = acc.lo; }
    }
}

// ---- exact inclusive scan of the chunk masses, on the device (128-bit integers: the result is order-independent) ----
// pass 1: block b scans its segment [b*seg, (b+1)*seg) in place and leaves the segment total in btot[b]
// pass 2: one block turns btot[] into exclusive offsets and writes the grand total to total2[0..1]
// pass 3: block b adds its offset to its segment
constexpr int ST = 256;
__device__ __forceinline__ u128 warp_incl_scan128(u128 v, unsigned lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u128 o;
        o.hi = __shfl_up_sync(0xffffffffu, v.hi, d);
        o.lo = __shfl_up_sync(0xffffffffu, v.lo, d);
        if ((int)lane >= d) add128(v, o);
    }
    return v;
}
// exclusive prefix of `mine` over the block's threads (in thread order); *block_total (optional) = sum over the block
__device__ __forceinline__ u128 block_excl_scan128(u128 mine, u128* block_total) {
    __shared__ uint64_t wt[ST / 32][2];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u128 incl = warp_incl_scan128(mine, lane);
    if (lane == 31) { wt[w][0] = incl.hi; wt[w][1] = incl.lo; }
    __syncthreads();
    u128 before = {0, 0}, all = {0, 0};
    for (unsigned k = 0; k < ST / 32; ++k) {
        const u128 t = {wt[k][0], wt[k][1]};
        if (k < w) add128(before, t);
        add128(all, t);
    }
    __syncthreads();
    if (block_total) *block_total = all;
    add128(before, sub128(incl, mine));
    return before;
}
__global__ void __launch_bounds__(ST) scan_segments_kernel(uint64_t* __restrict__ hi, uint64_t* __restrict__ lo, uint64_t count, uint64_t seg,
                                                           uint64_t* __restrict__ btot) {
    const uint64_t begin = (uint64_t)blockIdx.x * seg, end = begin + seg < count ? begin + seg : count;
    const uint64_t per = (seg + ST - 1) / ST;                       // consecutive items per thread
    const uint64_t b0 = begin + (uint64_t)threadIdx.x * per, b1 = b0 + per < end ? b0 + per : end;
    u128 mine = {0, 0};
    for (uint64_t i = b0; i < b1; ++i) add128(mine, u128{hi[i], lo[i]});
    u128 total;
    u128 run = block_excl_scan128(mine, &total);
    for (uint64_t i = b0; i < b1; ++i) {
        add128(run, u128{hi[i], lo[i]});
        hi[i] = run.hi; lo[i] = run.lo;
    }
    if (threadIdx.x == 0) { btot[2 * blockIdx.x] = total.hi; btot[2 * blockIdx.x + 1] = total.lo; }
}
__global__ void __launch_bounds__(ST) scan_totals_kernel(uint64_t* __restrict__ btot, unsigned nseg, uint64_t* __restrict__ total2) {
    const unsigned per = (nseg + ST - 1) / ST;
    const unsigned b0 = threadIdx.x * per, b1 = b0 + per < nseg ? b0 + per : nseg;
    u128 mine = {0, 0};
    for (unsigned i = b0; i < b1; ++i) add128(mine, u128{btot[2 * i], btot[2 * i + 1]});
    u128 total;
    u128 run = block_excl_scan128(mine, &total);
    for (unsigned i = b0; i < b1; ++i) {
        const u128 v = {btot[2 * i], btot[2 * i + 1]};
        btot[2 * i] = run.hi; btot[2 * i + 1] = run.lo;             // exclusive
        add128(run, v);
    }
    if (threadIdx.x == 0) { total2[0] = total.hi; total2[1] = total.lo; }
}
__global__ void __launch_bounds__(ST) scan_add_kernel(uint64_t* __restrict__ hi, uint64_t* __restrict__ lo, uint64_t count, uint64_t seg,
                                                      const uint64_t* __restrict__ btot) {
    if (blockIdx.x == 0) return;
    const u128 off = {btot[2 * blockIdx.x], btot[2 * blockIdx.x + 1]};
    const uint64_t begin = (uint64_t)blockIdx.x * seg, end = begin + seg < count ? begin + seg : count;
    for (uint64_t i = begin + threadIdx.x; i < end; i += ST) {
        u128 v = {hi[i], lo[i]};
        add128(v, off);
        hi[i] = v.hi; lo[i] = v.lo;
    }
}

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t& o0, uint32_t& o1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1;
}

// floor(S * U / 2^53), S < 2^125, U < 2^53 (oracle: orc_mul_u53)
__device__ __forceinline__ u128 mul_u53(const u128 S, const uint64_t U) {
    const uint64_t a_lo = S.lo * U, a_hi = __umul64hi(S.lo, U);      // A = S.lo * U
    const uint64_t b_lo = S.hi * U, b_hi = __umul64hi(S.hi, U);      // B = S.hi * U
    u128 r;                                                           // (B << 11) + (A >> 53)
    r.lo = b_lo << 11;
    r.hi = (b_hi << 11) | (b_lo >> 53);
    u128 a;
    a.lo = (a_lo >> 53) | (a_hi << 11);
    a.hi = a_hi >> 53;
    add128(r, a);
    return r;
}

// one warp per shot: index = min{ i : r < sum_{j<=i} q_j }, r = floor(U * S / 2^53)
__global__ void __launch_bounds__(RT) sample_kernel(const rq_cplx* __restrict__ state, unsigned n, unsigned chunk_bits,
                                                    const uint64_t* __restrict__ incl_hi, const uint64_t* __restrict__ incl_lo,
                                                    uint64_t nchunks, const uint64_t* __restrict__ totals4, uint64_t seed, uint64_t call,
                                                    unsigned shots, uint64_t shot_offset, const __grid_constant__ rq_shot_map map,
                                                    uint64_t* __restrict__ indices) {
    // totals4 = {total.hi, total.lo, win.hi, win.lo}: total = mass of the whole (possibly distributed) state; win = mass held
    // by lower ranks.  A shot whose threshold r falls outside [win, win + local mass) belongs to another rank: map.miss.
    const unsigned lane = threadIdx.x & 31;
    const unsigned warps = gridDim.x * (RT / 32);
    const uint64_t total_hi = totals4[0], total_lo = totals4[1], win_hi = totals4[2], win_lo = totals4[3];
    for (unsigned s = blockIdx.x * (RT / 32) + (threadIdx.x >> 5); s < shots; s += warps) {
        const uint64_t shot = shot_offset + s;
        uint32_t x0, x1;
        philox4x32_10((uint32_t)shot, (uint32_t)(shot >> 32), (uint32_t)call, (uint32_t)(call >> 32), (uint32_t)seed,
                      (uint32_t)(seed >> 32), x0, x1);
        const uint64_t U = (((uint64_t)x0 << 32) | x1) >> 11;
        u128 r = mul_u53(u128{total_hi, total_lo}, U);
        const u128 win = {win_hi, win_lo};
        const u128 local_total = {incl_hi[nchunks - 1], incl_lo[nchunks - 1]};
        bool mine = !lt128(r, win);
        if (mine) { r = sub128(r, win); mine = lt128(r, local_total); }
        if (!mine) { if (lane == 0) indices[s] = map.miss; continue; }
        uint64_t lo = 0, hi = nchunks - 1;                          // first chunk with r < incl[c]
        while (lo < hi) {
            const uint64_t mid = lo + ((hi - lo) >> 1);
            if (lt128(r, u128{incl_hi[mid], incl_lo[mid]})) hi = mid; else lo = mid + 1;
        }
        const uint64_t c = lo;
        u128 rr = r;
        if (c > 0) rr = sub128(r, u128{incl_hi[c - 1], incl_lo[c - 1]});
        const rq_cplx* p = state + (c << chunk_bits);
        const uint64_t len = 1ull << chunk_bits;
        u128 run = {0, 0};
        uint64_t found = len - 1;
        for (uint64_t seg = 0; seg < len; seg += 32) {
            const uint64_t i = seg + lane;
            u128 v = {0, 0};
            if (i < len) v = fix88(prob(p[i]));
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {                      // inclusive warp scan, exact
                u128 o;
                o.hi = __shfl_up_sync(0xffffffffu, v.hi, d);
                o.lo = __shfl_up_sync(0xffffffffu, v.lo, d);
                if ((int)lane >= d) add128(v, o);
            }
            add128(v, run);
            const unsigned hit = __ballot_sync(0xffffffffu, lt128(rr, v));
            if (hit) { found = seg + (unsigned)(__ffs((int)hit) - 1); break; }
            run.hi = __shfl_sync(0xffffffffu, v.hi, 31);
            run.lo = __shfl_sync(0xffffffffu, v.lo, 31);
        }
        if (lane == 0) {
            const uint64_t idx = map.high_base | ((c << chunk_bits) + found);
            uint64_t out = idx;
            if (map.nm != RQ_SHOT_RAW) {                            // result bit j = index bit pos[j] (hipStateVec.h:439-445)
                out = 0;
                for (unsigned j = 0; j < map.nm; ++j) out |= ((idx >> map.pos[j]) & 1ull) << j;
            }
            indices[s] = out;
        }
    }
}

}  // namespace

extern "C" unsigned rq_reduce_blocks(void) { return RBLOCKS; }

extern "C" int rq_launch_gather(rq_cplx* state, unsigned n, size_t batch, const unsigned* h_targets, unsigned k, uint64_t cmask,
                                const rq_cplx* d_matrix, void* stream) {
    gather_params P{};
    P.n = n; P.k = k; P.cmask = cmask;
    uint64_t fixed = cmask;
    for (unsigned b = 0; b < k; ++b) { P.tpos[b] = h_targets[b]; fixed |= 1ull << h_targets[b]; }
    unsigned nf = 0;
    for (unsigned p = 0; p < n; ++p) if ((fixed >> p) & 1ull) P.fix[nf++] = (unsigned char)p;
    P.nfix = nf;
    P.ngroups_per_state = 1ull << (n - nf);
    P.ngroups_total = P.ngroups_per_state * batch;
    const unsigned D = 1u << k, G = D >= (unsigned)RT ? 1u : (unsigned)RT / D;
    uint64_t blocks = (P.ngroups_total + G - 1) / G;
    if (blocks > RBLOCKS * 4ull) blocks = RBLOCKS * 4ull;
    if (blocks == 0) blocks = 1;
    const size_t smem = (size_t)G * D * sizeof(rq_cplx);
    gather_dense_kernel<<<(unsigned)blocks, RT, smem, (cudaStream_t)stream>>>(state, d_matrix, P);
    return (int)cudaGetLastError();
}

extern "C" int rq_launch_init_state(rq_cplx* state, size_t total_amps, int write_one, void* stream) {
    cudaError_t e = cudaMemsetAsync(state, 0, total_amps * sizeof(rq_cplx), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    if (write_one) set_one_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state);
    return (int)cudaGetLastError();
}

extern "C" int rq_launch_pauli_expect(const rq_cplx* state, unsigned n, uint64_t xmask, uint64_t zmask, unsigned ny,
                                      double* d_partials, unsigned nblocks, double* d_out, void* stream) {
    pauli_expect_kernel<<<nblocks, RT, 0, (cudaStream_t)stream>>>(state, n, xmask, zmask, ny, d_partials);
    sum_partials_kernel<<<1, RT, 0, (cudaStream_t)stream>>>(d_partials, nblocks, d_out);
    return (int)cudaGetLastError();
}

extern "C" int rq_launch_pauli_group(const rq_cplx* state, unsigned n, unsigned nstates, const rq_pauli_group* G, unsigned num_terms_total,
                                     double* d_partials, double* d_results, void* stream) {
    unsigned TT = 1;
    while (TT < G->nterms) TT <<= 1;
    if (TT > RQ_PAULI_GROUP_MAX || G->nterms == 0 || nstates == 0) return (int)cudaErrorInvalidValue;
    // small states: no more blocks than there is work for (the second stage reads nblocks partials per term)
    const uint64_t items = G->xmask ? ((1ull << n) >> 1) : (1ull << n);
    unsigned nb = (unsigned)((items + RT - 1) / RT);
    if (nb > RBLOCKS) nb = RBLOCKS;
    if (nb == 0) nb = 1;
    const dim3 grid(nb, nstates);
    cudaStream_t st = (cudaStream_t)stream;
    // several all-Z terms on a state of some size: the sign-word variant (0.50 / 0.55 / 0.73 ms for 4 / 8 / 16 terms at 28
    // qubits against 0.59 / 0.80 / 1.46 ms).  Groups with an x-mask stay on the POPC kernel: with two loads per item the
    // chunked loop was slower at 4 terms (0.57 against 0.44 ms per sweep, profiles/r02_expect_split.log).
    if (TT >= 4 && G->xmask == 0 && items >= (1ull << 15)) {
        switch (TT) {
            case 4: pauli_group_wide_kernel<4><<<grid, RT, 0, st>>>(state, n, *G, d_partials); break;
            case 8: pauli_group_wide_kernel<8><<<grid, RT, 0, st>>>(state, n, *G, d_partials); break;
            case 16: pauli_group_wide_kernel<16><<<grid, RT, 0, st>>>(state, n, *G, d_partials); break;
            default: pauli_group_kernel<32><<<grid, RT, 0, st>>>(state, n, *G, d_partials); break;     // (the engine groups <= 16 terms)
        }
    } else
    switch (TT) {
        case 1: pauli_group_kernel<1><<<grid, RT, 0, st>>>(state, n, *G, d_partials); break;
        case 2: pauli_group_kernel<2><<<grid, RT, 0, st>>>(state, n, *G, d_partials); break;
        case 4: pauli_group_kernel<4><<<grid, RT, 0, st>>>(state, n, *G, d_partials); break;
        case 8: pauli_group_kernel<8><<<grid, RT, 0, st>>>(state, n, *G, d_partials); break;
        case 16: pauli_group_kernel<16><<<grid, RT, 0, st>>>(state, n, *G, d_partials); break;
        default: pauli_group_kernel<32><<<grid, RT, 0, st>>>(state, n, *G, d_partials); break;
    }
    pauli_group_finish_kernel<<<dim3(G->nterms, nstates), RT, 0, st>>>(d_partials, nb, TT, *G, num_terms_total, d_results);
    return (int)cudaGetLastError();
}

extern "C" int rq_launch_fixed_masses(const rq_cplx* state, unsigned n, unsigned q, uint64_t* d_partials, unsigned nblocks,
                                      uint64_t* d_out4, void* stream) {
    fixed_masses_kernel<<<nblocks, RT, 0, (cudaStream_t)stream>>>(state, n, q, d_partials);
    sum_masses_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_partials, nblocks, d_out4);
    return (int)cudaGetLastError();
}

extern "C" int rq_launch_collapse(rq_cplx* state, unsigned n, unsigned q, int outcome, double scale, void* stream) {
    collapse_kernel<<<RBLOCKS, RT, 0, (cudaStream_t)stream>>>(state, n, q, outcome, (rq_real)scale);
    return (int)cudaGetLastError();
}

extern "C" int rq_launch_chunk_masses(const rq_cplx* state, unsigned n, unsigned chunk_bits, uint64_t* d_chunk_hi,
                                      uint64_t* d_chunk_lo, void* stream) {
    chunk_masses_kernel<<<RBLOCKS, RT, 0, (cudaStream_t)stream>>>(state, n, chunk_bits, d_chunk_hi, d_chunk_lo);
    return (int)cudaGetLastError();
}

extern "C" int rq_launch_scan_masses(uint64_t* d_hi, uint64_t* d_lo, uint64_t count, uint64_t* d_btot, uint64_t* d_total2, void* stream) {
    // <= RQ_SCAN_MAXSEG segments of at least ST items
    uint64_t seg = (count + RQ_SCAN_MAXSEG - 1) / RQ_SCAN_MAXSEG;
    if (seg < (uint64_t)ST) seg = ST;
    const unsigned nseg = (unsigned)((count + seg - 1) / seg);
    scan_segments_kernel<<<nseg, ST, 0, (cudaStream_t)stream>>>(d_hi, d_lo, count, seg, d_btot);
    scan_totals_kernel<<<1, ST, 0, (cudaStream_t)stream>>>(d_btot, nseg, d_total2);
    if (nseg > 1) scan_add_kernel<<<nseg, ST, 0, (cudaStream_t)stream>>>(d_hi, d_lo, count, seg, d_btot);
    return (int)cudaGetLastError();
}

extern "C" int rq_launch_sample(const rq_cplx* state, unsigned n, unsigned chunk_bits, const uint64_t* d_incl_hi,
                                const uint64_t* d_incl_lo, uint64_t nchunks, const uint64_t* d_totals4, uint64_t seed, uint64_t call,
                                unsigned shots, uint64_t shot_offset, const rq_shot_map* map, uint64_t* d_indices, void* stream) {
    unsigned blocks = (shots + (RT / 32) - 1) / (RT / 32);
    if (blocks > RBLOCKS) blocks = RBLOCKS;
    if (blocks == 0) blocks = 1;
    sample_kernel<<<blocks, RT, 0, (cudaStream_t)stream>>>(state, n, chunk_bits, d_incl_hi, d_incl_lo, nchunks, d_totals4, seed, call,
                                                           shots, shot_offset, *map, d_indices);
    return (int)cudaGetLastError();
}
