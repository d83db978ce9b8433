// mm_reduce.cuh - deterministic block / grid reductions with warp shuffles (fp64)
#pragma once
#include <cuda_runtime.h>

namespace mm {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return v;
}

// Sum NV per-thread values over the block; thread 0 writes them to out[0..NV).  Fixed order: shuffle tree inside
// each warp, then warp 0 adds the per-warp results in warp order -> bit-reproducible for a given launch shape.
template <int NV>
__device__ __forceinline__ void block_sum_store(const double (&vals)[NV], double *out) {
    __shared__ double red[32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        const double s = warp_sum(vals[k]);
        if (lane == 0) red[warp][k] = s;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double s = (lane < nwarps) ? red[lane][k] : 0.0;
            s = warp_sum(s);
            if (lane == 0) out[k] = s;
        }
    }
    __syncthreads();
}

// Sum per-block partials [nblocks][stride] for slots [0, nv) with one block of 256 threads (fixed order).
// Result for slot k is returned to thread 0 in res[k].
template <int NV>
__device__ __forceinline__ void partials_sum(const double *partials, int nblocks, int stride, double (&res)[NV]) {
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < NV; k++) acc[k] += partials[(size_t)b * stride + k];
    }
    __shared__ double out[NV];
    block_sum_store<NV>(acc, out);
#pragma unroll
    for (int k = 0; k < NV; k++) res[k] = out[k];
}

}  // namespace mm
