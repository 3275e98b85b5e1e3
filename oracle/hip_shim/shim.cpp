// TEST INFRASTRUCTURE ONLY: storage for the shim's per-thread launch geometry.
#include <hip/hip_runtime.h>
thread_local dim3 blockIdx, threadIdx, blockDim, gridDim;
