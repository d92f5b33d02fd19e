// Shared helpers for libmwx_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mwx_b200.h"

#define MWX_CUDA(call)                                   \
    do {                                                 \
        cudaError_t _e = (call);                         \
        if (_e != cudaSuccess) return (int)_e;           \
    } while (0)

#define MWX_LAUNCH_CHECK() MWX_CUDA(cudaGetLastError())

namespace mwx {

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

static inline unsigned grid_for(int64_t n, int block) {
    int64_t g = (n + block - 1) / block;
    return (unsigned)(g < 1 ? 1 : g);
}

// Number of SMs of the current device (148 on B200); cached per process.
int sm_count();

// Exclusive prefix sum of n int32 counts into int64 out[0..n] (out[n] = total).  Stream-ordered;
// temporaries come from cudaMallocAsync.
int exclusive_scan_i32_to_i64(const int32_t *in, int64_t n, int64_t *out, cudaStream_t st);
// int32 output variant (out[0..n]); total must fit in int32.
int exclusive_scan_i32_to_i32(const int32_t *in, int64_t n, int32_t *out, cudaStream_t st);

// The default memory pool hands freed blocks back to the OS at every synchronisation (release threshold 0): a 268 MB
// scratch array freed inside a step was unmapped at the step's final device synchronisation and mapped again in the next
// step - hundreds of milliseconds on some (virtualised) boxes.  Once per process the threshold is raised so that scratch
// memory stays in the pool.
void keep_pool_memory();

// Stream-ordered scratch buffer with RAII release.
template <typename T>
struct Scratch {
    T *p = nullptr;
    cudaStream_t st;
    explicit Scratch(cudaStream_t s) : st(s) {}
    cudaError_t alloc(size_t n) {
        keep_pool_memory();
        return cudaMallocAsync((void **)&p, (n ? n : 1) * sizeof(T), st);
    }
    ~Scratch() { if (p) cudaFreeAsync(p, st); }
    Scratch(const Scratch &) = delete;
    Scratch &operator=(const Scratch &) = delete;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int64_t warp_sum_i64(int64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum, result valid in thread 0.  `sh` holds >= 32 doubles.  Fixed order -> deterministic.
__device__ __forceinline__ double block_sum(double v, double *sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        r = lane < nw ? sh[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

}  // namespace mwx
