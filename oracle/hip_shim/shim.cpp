// TEST INFRASTRUCTURE ONLY: storage for the shim's per-thread launch geometry.
#include <hip/hip_runtime.h>
thread_local dim3 blockIdx, threadIdx, blockDim, gridDim;
bool shim_block_is_a_team = false;   // see __syncthreads in hip_runtime.h
