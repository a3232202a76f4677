// Per-graph masked reductions shared by the stand-alone mask kernels (s2v_mask.cu) and the fused head kernel
// (s2v_head.cu): one CTA of MK_THREADS threads per graph, warp shuffles with a FIXED combine order (deterministic),
// first-index tie break (np.argmax / torch.argmax semantics).  Both users run exactly this code, so a fused launch
// and the three-launch sequence give bit-identical results.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#define MK_THREADS 128

struct ArgPair { float v; int i; };

__device__ __forceinline__ ArgPair better(ArgPair a, ArgPair b) {
    // larger value wins; on ties the smaller index; i < 0 means "empty"
    if (b.i < 0) return a;
    if (a.i < 0) return b;
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}

__device__ __forceinline__ ArgPair block_argmax(ArgPair p, ArgPair* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgPair q;
        q.v = __shfl_down_sync(0xffffffffu, p.v, o);
        q.i = __shfl_down_sync(0xffffffffu, p.i, o);
        p = better(p, q);
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = p;
    __syncthreads();
    ArgPair r = sh[0];
    for (int w = 1; w < MK_THREADS / 32; ++w) r = better(r, sh[w]);
    __syncthreads();
    return r;
}

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = sh[0];
    for (int w = 1; w < MK_THREADS / 32; ++w) r += sh[w];
    __syncthreads();
    return r;
}

__device__ __forceinline__ float block_max(float v, float* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, o));
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = sh[0];
    for (int w = 1; w < MK_THREADS / 32; ++w) r = fmaxf(r, sh[w]);
    __syncthreads();
    return r;
}

