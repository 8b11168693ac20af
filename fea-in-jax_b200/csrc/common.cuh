// Shared host-side plumbing of libfeb200: error reporting, launch checks, small device helpers.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/feb200.h"

namespace feb {

void set_error(const char* fmt, ...);

#define FEB_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) {                          \
            feb::set_error(__VA_ARGS__);        \
            return (code);                      \
        }                                       \
    } while (0)

// Checks the launch itself (configuration errors); never synchronises.
#define FEB_CHECK_LAUNCH(name)                                                        \
    do {                                                                              \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) {                                                     \
            feb::set_error("%s: launch failed: %s", (name), cudaGetErrorString(e__)); \
            return FEB200_ERR_CUDA;                                                   \
        }                                                                             \
    } while (0)

#define FEB_CUDA(call)                                                               \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            feb::set_error("%s failed: %s", #call, cudaGetErrorString(e__));         \
            return FEB200_ERR_CUDA;                                                  \
        }                                                                            \
    } while (0)

static inline cudaStream_t as_stream(feb200_stream s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// number of SMs of the current device (cached); 148 on B200
int sm_count();

// grid size of a grid-stride kernel over n items: a multiple-of-SM-count cap (16 CTAs per SM)
static inline int grid_for(int64_t n, int block) {
    int64_t g = ceil_div(n, block);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// Deterministic grid-wide sum: every block writes its partial to part[blockIdx.x]; the last block
// to arrive (ticket counter) adds the partials in index order, so the result does not depend on
// block scheduling.  Layout of the reduce workspace (doubles): [0 .. MAX_PARTIALS*NRED) partials,
// then one uint32 ticket counter per "channel".
constexpr int MAX_PARTIALS = 4096;     // upper bound on gridDim.x of reducing kernels
constexpr int REDUCE_CHANNELS = 4;     // independent ticket counters
constexpr size_t REDUCE_WS_BYTES = (size_t)MAX_PARTIALS * 4 * sizeof(double) + REDUCE_CHANNELS * sizeof(unsigned int) + 64;

#if defined(__CUDACC__)
// material parameters of element e (layouts of fea.py:77-80): pointer to quadrature point 0 and the stride
// between quadrature points (0 when the parameters are constant over the element)
__device__ __forceinline__ const double* mp_of(const feb200_batch& b, int64_t e, int& q_stride) {
    if (b.layout == FEB200_LAYOUT_EQM) { q_stride = b.M; return b.mp + (e * b.Q) * (int64_t)b.M; }
    q_stride = 0;
    return (b.layout == FEB200_LAYOUT_EM) ? b.mp + e * (int64_t)b.M : b.mp;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of NV values per thread; result valid in thread 0. blockDim.x multiple of 32, <= 1024.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /* NV*32 doubles */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; ++k) smem[k * 32 + warp] = v[k];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double t = (lane < nwarp) ? smem[k * 32 + lane] : 0.0;
            v[k] = warp_sum(t);
        }
    }
}

// Called by all threads of every block after block_sum; returns true in thread 0 of the LAST block,
// with total[k] holding the grid-wide sums (fixed summation order).  part: NV * MAX_PARTIALS doubles.
template <int NV>
__device__ __forceinline__ bool grid_sum_last_block(const double (&v)[NV], double* part, unsigned int* ticket,
                                                    double (&total)[NV]) {
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) part[k * MAX_PARTIALS + blockIdx.x] = v[k];
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    // last block: ordered sum of the partials (strided per thread, then a fixed shuffle tree)
    __shared__ double red[NV * 32];
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double a = 0.0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x)
            a += __ldcg(&part[k * MAX_PARTIALS + i]);
        acc[k] = a;
    }
    block_sum<NV>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) total[k] = acc[k];
        *ticket = 0u;   // ready for the next launch on the same stream
        return true;
    }
    return false;
}
#endif

}  // namespace feb
