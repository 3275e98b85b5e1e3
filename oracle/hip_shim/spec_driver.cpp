// oracle/hip_shim/spec_driver.cpp -- TEST INFRASTRUCTURE ONLY (compiled into oracle/_ref by oracle/Makefile).
//
// The reference declares rocsvApplyMatrix (hipStateVec.h:151-157) but never defines it; what it does ship is the kernel
// that call was meant to launch for up to four target qubits, apply_multi_qubit_generic_matrix_kernel
// (multi_qubit_kernels.hip:37-115) -- compiled here unmodified with the rest of that file, but launched by nothing.
// This driver supplies the missing launch: it runs the reference's kernel body once per "thread" (one thread per group of
// 2^m amplitudes, exactly as the kernel indexes itself: blockIdx.x * blockDim.x + threadIdx.x, :49-53) so the oracle's
// restatement of ApplyMatrix -- bit order of the matrix index, column-major storage, accumulation in amplitude
// precision -- can be pinned against the reference's own code instead of against a reading of it.
// Nothing of the reference is copied: only the kernel's prototype is repeated so that the linker can find it.
#include <hip/hip_runtime.h>

#include <cmath>
#include <cstdlib>

#include "rocquantum/hipStateVec.h"

__global__ void apply_multi_qubit_generic_matrix_kernel(rocComplex* state, unsigned numQubits, const unsigned* targetQubitIndices,
                                                        unsigned m, const rocComplex* matrixDevice);

extern "C" int refspec_apply_matrix(rocComplex* state, unsigned numQubits, const unsigned* targets, unsigned m,
                                    const rocComplex* matrixColumnMajor) {
    if (!state || !targets || !matrixColumnMajor || m < 1 || m > 4 || m > numQubits || numQubits > 31) return -1;   // its local arrays hold 16
    const long long groups = (long long)(((size_t)1 << numQubits) >> m);
#pragma omp parallel for schedule(static)
    for (long long t = 0; t < groups; ++t) {
        gridDim = dim3((unsigned)groups);
        blockDim = dim3(1);
        threadIdx = dim3(0);
        blockIdx = dim3((unsigned)t);
        apply_multi_qubit_generic_matrix_kernel(state, numQubits, targets, m, matrixColumnMajor);
    }
    return 0;
}

// ---- measurement_kernels.hip: defined by the reference, launched by nothing (rocsvMeasure and the expectation calls are
// declared only).  Prototypes repeated for the linker; bodies are the reference's, compiled unmodified. ------------------
__global__ void calculate_prob0_kernel(const rocComplex* state, unsigned numQubits, unsigned targetQubit, real_t* d_prob0_sum);
__global__ void collapse_state_kernel(rocComplex* state, unsigned numQubits, unsigned targetQubit, int measuredOutcome);
__global__ void renormalize_state_kernel(rocComplex* state, unsigned numQubits, real_t d_sum_sq_mag_inv_sqrt);
__global__ void sum_sq_magnitudes_kernel(const rocComplex* state, unsigned numQubits, real_t* d_sum_sq_mag);
__global__ void calculate_multi_z_probabilities_kernel(const rocComplex* local_slice_data, size_t local_slice_num_elements,
                                                       unsigned num_local_qubits, const unsigned* d_target_qubits,
                                                       unsigned num_target_paulis, real_t* d_outcome_probs_blocks);
__global__ void reduce_multi_z_block_probs_to_slice_total_kernel(const real_t* d_block_outcome_probs, unsigned num_prev_blocks,
                                                                 unsigned num_outcomes, real_t* d_slice_total_outcome_probs);

// the kernels' `extern __shared__` arrays: one thread per block, so a block's scratch is per thread (hip_runtime.h)
thread_local real_t sdata[2], sdata_sum[1], s_reduce_probs[2], s_reduce_sum_sq[1], s_prob_bins[256], s_final_probs[256];

static void as_thread(unsigned block, unsigned nblocks) {
    gridDim = dim3(nblocks);
    blockDim = dim3(1);
    threadIdx = dim3(0);
    blockIdx = dim3(block);
}

// The measurement sequence the reference's kernels and MULTI_GPU_GUIDE.md:61-78 describe, for a GIVEN outcome (the
// reference fixes no RNG): prob0 -> collapse -> sum of squares -> renormalise.  Returns prob0 through *prob0.
extern "C" int refspec_measure_with_outcome(rocComplex* state, unsigned numQubits, unsigned qubit, int outcome, double* prob0) {
    if (!state || !prob0 || qubit >= numQubits || numQubits > 31 || (outcome != 0 && outcome != 1)) return -1;
    const long long N = 1ll << numQubits;
    real_t p0 = 0, sum = 0;
    as_thread(0, 1);
    calculate_prob0_kernel(state, numQubits, qubit, &p0);
#pragma omp parallel for schedule(static)
    for (long long t = 0; t < N; ++t) { as_thread((unsigned)t, (unsigned)N); collapse_state_kernel(state, numQubits, qubit, outcome); }
    as_thread(0, 1);
    sum_sq_magnitudes_kernel(state, numQubits, &sum);
    if (!(sum > 0)) return -2;
    const real_t inv = (real_t)(1.0 / sqrt((double)sum));
#pragma omp parallel for schedule(static)
    for (long long t = 0; t < N; ++t) { as_thread((unsigned)t, (unsigned)N); renormalize_state_kernel(state, numQubits, inv); }
    *prob0 = (double)p0;
    return 0;
}

// Probabilities of the 2^k joint outcomes of k <= 8 qubits in the Z basis (bin bit j <-> targets[j]), by the reference's
// two kernels: per-block bins, then one thread per bin summing over blocks.  probs: 2^k doubles.
extern "C" int refspec_multi_z_probabilities(const rocComplex* state, unsigned numQubits, const unsigned* targets, unsigned k, double* probs) {
    if (!state || !targets || !probs || k < 1 || k > 8 || numQubits > 24) return -1;
    const size_t N = (size_t)1 << numQubits;
    const unsigned bins = 1u << k;
    real_t* blocks = (real_t*)calloc(N * bins, sizeof(real_t));
    real_t* total = (real_t*)calloc(bins, sizeof(real_t));
    if (!blocks || !total) { free(blocks); free(total); return -2; }
#pragma omp parallel for schedule(static)
    for (long long b = 0; b < (long long)N; ++b) {
        as_thread((unsigned)b, (unsigned)N);
        // the kernel clears bin t from thread t of the block (:313-315); this block has thread 0 only, so the clearing the
        // absent threads would do is done here -- binning, accumulation and the write-out below are the kernel's own
        for (unsigned t = 1; t < bins; ++t) s_prob_bins[t] = 0;
        calculate_multi_z_probabilities_kernel(state, N, numQubits, targets, k, blocks);
    }
    for (unsigned t = 0; t < bins; ++t) {
        as_thread(t, bins);
        reduce_multi_z_block_probs_to_slice_total_kernel(blocks, (unsigned)N, bins, total);
    }
    for (unsigned t = 0; t < bins; ++t) probs[t] = (double)total[t];
    free(blocks);
    free(total);
    return 0;
}

// ---- swap_kernels.hip:95-114: the local<->local case of rocsvSwapIndexBits (declared, never defined).  The kernel copies
// the slice to a scratch buffer, waits at a barrier, then scatters every element to the index with the two bits exchanged
// -- so its threads must really run together: one OpenMP thread per amplitude, __syncthreads an OpenMP barrier. -----------
__global__ void local_bit_swap_permutation_kernel(rocComplex* d_local_slice, rocComplex* d_temp_buffer_for_slice,
                                                  size_t local_slice_num_elements, unsigned local_qubit_idx1, unsigned local_qubit_idx2);

extern "C" int refspec_local_bit_swap(rocComplex* state, unsigned numQubits, unsigned q1, unsigned q2) {
    if (!state || numQubits > 8 || q1 >= numQubits || q2 >= numQubits) return -1;       // 2^n threads in one team
    const int N = 1 << numQubits;
    rocComplex* temp = (rocComplex*)calloc((size_t)N, sizeof(rocComplex));
    if (!temp) return -2;
    int team = 0;
    omp_set_dynamic(0);
    shim_block_is_a_team = true;
#pragma omp parallel num_threads(N)
    {
#pragma omp single
        team = omp_get_num_threads();
        if (omp_get_num_threads() == N) {                                                // uniform across the team
            gridDim = dim3(1);
            blockDim = dim3((unsigned)N);
            blockIdx = dim3(0);
            threadIdx = dim3((unsigned)omp_get_thread_num());
            local_bit_swap_permutation_kernel(state, temp, (size_t)N, q1, q2);
        }
    }
    shim_block_is_a_team = false;
    free(temp);
    return team == N ? 0 : -3;
}
