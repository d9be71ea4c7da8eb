// Shared host/device helpers for the occnerf_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/occnerf_b200.h"

void occnerf_set_error(const char *fmt, ...);

#define OCC_CHECK_ARG(cond, ...)                              \
    do {                                                      \
        if (!(cond)) {                                        \
            occnerf_set_error(__VA_ARGS__);                   \
            return OCCNERF_EINVAL;                            \
        }                                                     \
    } while (0)

#define OCC_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            occnerf_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return OCCNERF_ECUDA;                                                          \
        }                                                                                  \
    } while (0)

#define OCC_LAUNCH_CHECK() OCC_CUDA(cudaGetLastError())

static inline unsigned occ_div_up(long a, long b) { return (unsigned)((a + b - 1) / b); }

#define OCC_FULL 0xffffffffu

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(OCC_FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(OCC_FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(OCC_FULL, v, o));
    return v;
}

// vector reductions into global memory (sm_90+): one L2 atomic transaction for 2 / 4 floats
__device__ __forceinline__ void red_add_v2(float *addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
