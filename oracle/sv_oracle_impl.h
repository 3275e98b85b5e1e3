/* oracle/sv_oracle_impl.h -- TEST INFRASTRUCTURE ONLY (included twice by sv_oracle.c).
 *
 * Precision-generic body of the CPU restatement.  Before including define
 *   REAL   float | double            (the reference's real_t,  hipStateVec.h:6-15)
 *   SFX(x) x##_c64 | x##_c128        (symbol suffix)
 *
 * Every function cites the reference file:line whose behaviour it restates
 * (paths relative to /root/reference).  Entry points the reference declares but
 * never defines are restated from the header contract; where the
 * reference ships the kernel such a call was meant to launch (ApplyMatrix k <= 4, Measure,
 * Z-product probabilities, the local bit swap) the restatement is pinned against that kernel
 * run through oracle/hip_shim/spec_driver.cpp; the rest is marked "parity unpinned" -- pinned
 * only by the analytic known-answer vectors in tests/golden/ (see the list in sv_oracle.c).
 */

typedef struct { REAL x, y; } SFX(cplx);
#define CPLX SFX(cplx)

/* complex helpers: the reference computes products as (ax*bx - ay*by, ax*by + ay*bx)
 * with separate mul/add in amplitude precision (single_qubit_kernels.hip:14-20). */
static inline CPLX SFX(cmul)(CPLX a, CPLX b) {
    CPLX r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
static inline CPLX SFX(cadd)(CPLX a, CPLX b) { CPLX r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }

/* hipStateVec.cpp:253-272 -- memset the WHOLE batched buffer, then write (1,0) at
 * absolute element 0 only (batch members >0 stay all-zero, as in the reference). */
void SFX(orc_init_state)(CPLX* s, unsigned n, size_t batch) {
    size_t total = batch * ((size_t)1 << n);
    memset(s, 0, total * sizeof(CPLX));
    s[0].x = (REAL)1; s[0].y = (REAL)0;
}

/* single_qubit_kernels.hip:28-72 (pair indexing :53-55, arithmetic :64-67).
 * m = {m00,m01,m10,m11} as (re,im) doubles, narrowed like make_complex, hipStateVec.cpp:72-78. */
void SFX(orc_apply_matrix1)(CPLX* s, unsigned n, size_t batch, unsigned t, const double* m) {
    CPLX m00 = {(REAL)m[0], (REAL)m[1]}, m01 = {(REAL)m[2], (REAL)m[3]};
    CPLX m10 = {(REAL)m[4], (REAL)m[5]}, m11 = {(REAL)m[6], (REAL)m[7]};
    size_t N = (size_t)1 << n, pairs = N >> 1, stride = (size_t)1 << t;
    if (pairs == 0) return;
    #pragma omp parallel for schedule(static)
    for (size_t p = 0; p < batch * pairs; ++p) {
        size_t b = p / pairs, q = p % pairs;
        size_t i0 = b * N + ((q & ~(stride - 1)) * 2 + (q & (stride - 1))), i1 = i0 + stride;
        CPLX a0 = s[i0], a1 = s[i1];
        s[i0] = SFX(cadd)(SFX(cmul)(m00, a0), SFX(cmul)(m01, a1));
        s[i1] = SFX(cadd)(SFX(cmul)(m10, a0), SFX(cmul)(m11, a1));
    }
}

/* single_qubit_kernels.hip:78-128 -- same, skipped where the control bit is 0 (:109-111). */
void SFX(orc_apply_cmatrix1)(CPLX* s, unsigned n, size_t batch, unsigned c, unsigned t, const double* m) {
    CPLX m00 = {(REAL)m[0], (REAL)m[1]}, m01 = {(REAL)m[2], (REAL)m[3]};
    CPLX m10 = {(REAL)m[4], (REAL)m[5]}, m11 = {(REAL)m[6], (REAL)m[7]};
    size_t N = (size_t)1 << n, pairs = N >> 1, stride = (size_t)1 << t, cm = (size_t)1 << c;
    if (pairs == 0) return;
    #pragma omp parallel for schedule(static)
    for (size_t p = 0; p < batch * pairs; ++p) {
        size_t b = p / pairs, q = p % pairs;
        size_t l0 = (q & ~(stride - 1)) * 2 + (q & (stride - 1));
        if ((l0 & cm) == 0) continue;
        size_t i0 = b * N + l0, i1 = i0 + stride;
        CPLX a0 = s[i0], a1 = s[i1];
        s[i0] = SFX(cadd)(SFX(cmul)(m00, a0), SFX(cmul)(m01, a1));
        s[i1] = SFX(cadd)(SFX(cmul)(m10, a0), SFX(cmul)(m11, a1));
    }
}

/* multi_qubit_kernels.hip:227-265 -- X on target where (idx & mask) == mask.
 * CNOT (two_qubit_kernels.hip:17-51) is the one-control case. */
void SFX(orc_mcx)(CPLX* s, unsigned n, size_t batch, unsigned long long mask, unsigned t) {
    size_t N = (size_t)1 << n, tm = (size_t)1 << t;
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < batch * N; ++i) {
        size_t l = i % N;
        if ((l & tm) || (l & mask) != mask) continue;
        CPLX tmp = s[i]; s[i] = s[i | tm]; s[i | tm] = tmp;
    }
}

/* two_qubit_kernels.hip:54-80 -- negate where both bits are 1. */
void SFX(orc_cz)(CPLX* s, unsigned n, size_t batch, unsigned a, unsigned b) {
    size_t N = (size_t)1 << n, m = ((size_t)1 << a) | ((size_t)1 << b);
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < batch * N; ++i) {
        if (((i % N) & m) != m) continue;
        s[i].x = -s[i].x; s[i].y = -s[i].y;
    }
}

/* two_qubit_kernels.hip:83-131 -- swap |01> <-> |10>;  with a control: multi_qubit_kernels.hip:268-307. */
void SFX(orc_cswap)(CPLX* s, unsigned n, size_t batch, unsigned long long cmask, unsigned a, unsigned b) {
    size_t N = (size_t)1 << n, am = (size_t)1 << a, bm = (size_t)1 << b;
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < batch * N; ++i) {
        size_t l = i % N;
        if ((l & cmask) != cmask) continue;
        if ((l & am) || !(l & bm)) continue;           /* only the a=0,b=1 branch moves */
        size_t j = (i | am) & ~bm;
        CPLX tmp = s[i]; s[i] = s[j]; s[j] = tmp;
    }
}

/* rocsvApplyMatrix / rocsvApplyControlledMatrix -- PARITY UNPINNED (declared hipStateVec.h:151-157,
 * 461-468; never defined).  Layout follows the only executable spec, multi_qubit_kernels.hip:37-115:
 * column-major M[i + j*dim] (:26), matrix-index bit b <-> targets[b] (:91-99), accumulate
 * sum_j M_ij * a_j in amplitude precision (:22-30).  Controls: act where all control bits are 1. */
void SFX(orc_apply_matrix)(CPLX* s, unsigned n, size_t batch, const unsigned* targets, unsigned k,
                           const unsigned* controls, unsigned nc, const CPLX* M) {
    size_t N = (size_t)1 << n, dim = (size_t)1 << k;
    size_t cmask = 0, fixed = 0;
    for (unsigned b = 0; b < k; ++b) fixed |= (size_t)1 << targets[b];
    for (unsigned b = 0; b < nc; ++b) cmask |= (size_t)1 << controls[b];
    fixed |= cmask;
    unsigned fixpos[64], nfix = 0;
    for (unsigned p = 0; p < n; ++p) if ((fixed >> p) & 1) fixpos[nfix++] = p;
    size_t* off = (size_t*)malloc(dim * sizeof(size_t));
    for (size_t j = 0; j < dim; ++j) {
        size_t o = 0;
        for (unsigned b = 0; b < k; ++b) if ((j >> b) & 1) o |= (size_t)1 << targets[b];
        off[j] = o;
    }
    const size_t groups = N >> nfix;                       /* one group per assignment of the free bits */
    #pragma omp parallel
    {
        CPLX* in = (CPLX*)malloc(dim * sizeof(CPLX));
        #pragma omp for schedule(static)
        for (size_t g = 0; g < batch * groups; ++g) {
            size_t base = g % groups;
            for (unsigned f = 0; f < nfix; ++f) {          /* deposit the free bits around the fixed positions */
                const unsigned p = fixpos[f];
                base = ((base >> p) << (p + 1)) | (base & (((size_t)1 << p) - 1));
            }
            const size_t i = (g / groups) * N + (base | cmask);
            for (size_t j = 0; j < dim; ++j) in[j] = s[i + off[j]];
            for (size_t r = 0; r < dim; ++r) {
                CPLX acc = {(REAL)0, (REAL)0};
                for (size_t j = 0; j < dim; ++j) {
                    CPLX mij = M[r + j * dim], v = in[j];
                    acc.x += mij.x * v.x - mij.y * v.y;
                    acc.y += mij.x * v.y + mij.y * v.x;
                }
                s[i + off[r]] = acc;
            }
        }
        free(in);
    }
    free(off);
}

/* rocsvSwapIndexBits, local case -- PARITY UNPINNED (declared hipStateVec.h:135-137).  The data
 * movement is swap_kernels.hip:95-114: out[swap_bits(i,a,b)] = in[i]; done here in place. */
void SFX(orc_swap_index_bits)(CPLX* s, unsigned n, size_t batch, unsigned a, unsigned b) {
    if (a == b) return;
    SFX(orc_cswap)(s, n, batch, 0ULL, a, b);
}

/* |a|^2 in double with one fused multiply-add: the summation spec shared with the GPU engine
 * (DESIGN.md "sampling spec"); fma() is exactly rounded so both sides agree bit for bit. */
static inline double SFX(prob)(CPLX a) { return fma((double)a.x, (double)a.x, (double)a.y * (double)a.y); }

double SFX(orc_norm2)(const CPLX* s, unsigned n) {
    size_t N = (size_t)1 << n; double acc = 0.0;
    #pragma omp parallel for reduction(+:acc) schedule(static)
    for (size_t i = 0; i < N; ++i) acc += SFX(prob)(s[i]);
    return acc;
}

/* exact 128-bit fixed-point masses: S0 = sum over bit q = 0, S1 = bit q = 1 (q >= n: everything in S0) */
void SFX(orc_fixed_masses)(const CPLX* s, unsigned n, unsigned q, u128* S0, u128* S1) {
    size_t N = (size_t)1 << n; u128 a0 = 0, a1 = 0;
    for (size_t i = 0; i < N; ++i) {
        u128 v = orc_fix88(SFX(prob)(s[i]));
        if (q < n && ((i >> q) & 1)) a1 += v; else a0 += v;
    }
    *S0 = a0; *S1 = a1;
}

/* rocsvMeasure -- PARITY UNPINNED (declared hipStateVec.h:172-177).  Algorithm: measurement_kernels.hip
 * (p0/p1 reduction :103-157, collapse :37-58, renormalise by 1/sqrt(p) :64-77) and
 * MULTI_GPU_GUIDE.md:61-78.  The draw is ours: U = philox(seed; call,0), outcome 0 iff
 * floor(U*(S0+S1)/2^53) < S0 on exact fixed-point masses. */
void SFX(orc_measure)(CPLX* s, unsigned n, unsigned q, uint64_t seed, uint64_t call, int* outcome, double* prob) {
    u128 S0, S1; SFX(orc_fixed_masses)(s, n, q, &S0, &S1);
    uint64_t U = orc_uniform53(seed, call, 0);
    u128 r = orc_mul_u53(S0 + S1, U);
    int out = (r < S0) ? 0 : 1;
    double tot = orc_u128_to_double(S0 + S1), mass = orc_u128_to_double(out ? S1 : S0);
    *outcome = out; *prob = mass / tot;
    REAL scale = (REAL)(1.0 / sqrt(mass * 0x1p-88));
    size_t N = (size_t)1 << n;
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; ++i) {
        int bit = (int)((i >> q) & 1);
        if (bit != out) { s[i].x = (REAL)0; s[i].y = (REAL)0; }
        else { s[i].x *= scale; s[i].y *= scale; }
    }
}

/* rocsvGetExpectationPauliString and the single-Pauli / Z-product special cases -- PARITY UNPINNED
 * (declared hipStateVec.h:340-423).  <psi|P|psi> with P = prod_k sigma_k on qubits[k]:
 * P|i> = i^{nY} (-1)^{popcount(i & zmask)} |i ^ xmask>, zmask = Z|Y positions, xmask = X|Y positions.
 * Z-parity form agrees with hipDensityMat.cpp:531-540; X/Y signs with hipDensityMat.cpp:99-113.
 * Non-destructive (hipStateVec.h:406). */
double SFX(orc_expect_pauli)(const CPLX* s, unsigned n, const char* paulis, const unsigned* qubits, unsigned k) {
    size_t N = (size_t)1 << n, xm = 0, zm = 0; unsigned ny = 0;
    for (unsigned j = 0; j < k; ++j) {
        char c = paulis[j]; size_t bit = (size_t)1 << qubits[j];
        if (c == 'X' || c == 'x') xm ^= bit;
        else if (c == 'Y' || c == 'y') { xm ^= bit; zm ^= bit; ++ny; }
        else if (c == 'Z' || c == 'z') zm ^= bit;
    }
    /* <psi|P|psi> = sum_i conj(psi_{i^x}) * phase(i) * psi_i,  phase(i) = i^ny * (-1)^{pc(i&z)} */
    double re = 0.0;
    #pragma omp parallel for reduction(+:re) schedule(static)
    for (size_t i = 0; i < N; ++i) {
        CPLX a = s[i], b = s[i ^ xm];
        /* t = conj(b) * a */
        double tr = (double)b.x * a.x + (double)b.y * a.y;
        double ti = (double)b.x * a.y - (double)b.y * a.x;
        double pr, pi;                       /* i^ny */
        switch (ny & 3) { case 0: pr = 1; pi = 0; break; case 1: pr = 0; pi = 1; break;
                          case 2: pr = -1; pi = 0; break; default: pr = 0; pi = -1; }
        double sgn = (__builtin_popcountll((unsigned long long)(i & zm)) & 1) ? -1.0 : 1.0;
        re += sgn * (pr * tr - pi * ti);
    }
    return re;
}

/* rocsvSample -- PARITY UNPINNED (declared hipStateVec.h:439-445; the reference fixes no RNG, SURVEY
 * section 8c).  Spec: q_i = fix88(|a_i|^2) exact integers, S = sum q_i; shot s draws
 * U = philox(seed; call, s); r = floor(U*S / 2^53); index = min{ i : r < sum_{j<=i} q_j };
 * result bit j = bit measured[j] of index (hipStateVec.h:427-445, examples/sampling_example.py:31-33). */
void SFX(orc_sample)(const CPLX* s, unsigned n, const unsigned* measured, unsigned nm, unsigned shots,
                     uint64_t seed, uint64_t call, uint64_t* out) {
    size_t N = (size_t)1 << n;
    u128* pre = (u128*)malloc(N * sizeof(u128));
    u128 acc = 0;
    for (size_t i = 0; i < N; ++i) { acc += orc_fix88(SFX(prob)(s[i])); pre[i] = acc; }
    for (unsigned sh = 0; sh < shots; ++sh) {
        uint64_t U = orc_uniform53(seed, call, sh);
        u128 r = orc_mul_u53(acc, U);
        size_t lo = 0, hi = N - 1;                     /* first i with r < pre[i] */
        while (lo < hi) { size_t mid = lo + ((hi - lo) >> 1); if (r < pre[mid]) hi = mid; else lo = mid + 1; }
        uint64_t bits = 0;
        for (unsigned j = 0; j < nm; ++j) bits |= (uint64_t)((lo >> measured[j]) & 1) << j;
        out[sh] = bits;
    }
    free(pre);
}

#undef CPLX
