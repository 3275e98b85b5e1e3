// oracle/hip_shim/hip/hip_runtime.h  --  TEST INFRASTRUCTURE ONLY.
//
// A host stand-in for <hip/hip_runtime.h>, just large enough that the
// reference's own hipStateVec.cpp + {single,two,multi}_qubit_kernels.hip
// compile UNMODIFIED with g++ (see oracle/Makefile) and run on host threads.
// It is valid because every kernel the reference launches is a
// __syncthreads-free grid-stride loop.  A "launch" runs the
// kernel body once per OpenMP thread with gridDim = #threads, blockDim = 1.
// Nothing under rocquantum_b200/ may include or link this.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <omp.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct hipFloatComplex { float x, y; };
struct hipDoubleComplex { double x, y; };
typedef hipFloatComplex hipComplex;

typedef int hipError_t;
static const hipError_t hipSuccess = 0;
static const hipError_t hipErrorOutOfMemory = 2;
typedef void* hipStream_t;
enum hipMemcpyKind { hipMemcpyHostToHost, hipMemcpyHostToDevice, hipMemcpyDeviceToHost, hipMemcpyDeviceToDevice };

extern thread_local dim3 blockIdx, threadIdx, blockDim, gridDim;

// measurement_kernels.hip (compiled for oracle/hip_shim/spec_driver.cpp only; the reference launches none of it) uses
// block-shared scratch, barriers and shared-memory atomics.  With ONE thread per block -- the only geometry this shim
// ever runs -- a block's shared memory is that thread's own, a barrier has nobody to wait for, an atomic has no rival.
#define __shared__ thread_local
// (spec_driver.cpp also runs ONE kernel -- local_bit_swap_permutation_kernel, which has no shared memory -- as a real block
// of OpenMP threads; only then is the flag set and the barrier real.)
extern bool shim_block_is_a_team;
static inline void __syncthreads() {
    if (shim_block_is_a_team) {
        _Pragma("omp barrier")
    }
}
template <typename T, typename V> static inline T atomicAdd(T* p, V v) { const T old = *p; *p = old + (T)v; return old; }
// swap_kernels.hip's packing kernels (never run here: their cursor order is nondeterministic by design) need the name
template <typename T, typename V> static inline T hipAtomicAdd(T* p, V v) { return atomicAdd(p, v); }

template <typename T> static inline hipError_t hipMalloc(T** p, size_t bytes) {
    void* q = nullptr;
    if (posix_memalign(&q, 64, bytes ? bytes : 64) != 0) { *p = nullptr; return hipErrorOutOfMemory; }
    *p = static_cast<T*>(q);
    return hipSuccess;
}
static inline hipError_t hipFree(void* p) { free(p); return hipSuccess; }
static inline hipError_t hipMemset(void* p, int v, size_t bytes) { memset(p, v, bytes); return hipSuccess; }
static inline hipError_t hipMemcpy(void* d, const void* s, size_t bytes, hipMemcpyKind) { memcpy(d, s, bytes); return hipSuccess; }
static inline hipError_t hipStreamCreate(hipStream_t* s) { *s = nullptr; return hipSuccess; }
static inline hipError_t hipStreamDestroy(hipStream_t) { return hipSuccess; }
static inline hipError_t hipStreamSynchronize(hipStream_t) { return hipSuccess; }
static inline hipError_t hipDeviceSynchronize() { return hipSuccess; }
static inline hipError_t hipGetLastError() { return hipSuccess; }
static inline const char* hipGetErrorString(hipError_t) { return "host-shim"; }

// One "block" per OpenMP thread, one "thread" per block: every launched
// reference kernel strides by gridDim.x*blockDim.x, so this covers all work.
#define hipLaunchKernelGGL(kernel, grid, block, shmem, stream, ...)                 \
    do {                                                                            \
        _Pragma("omp parallel")                                                     \
        {                                                                           \
            gridDim = dim3((unsigned)omp_get_num_threads());                        \
            blockDim = dim3(1);                                                     \
            blockIdx = dim3((unsigned)omp_get_thread_num());                        \
            threadIdx = dim3(0);                                                    \
            kernel(__VA_ARGS__);                                                    \
        }                                                                           \
    } while (0)
