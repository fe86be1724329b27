// csrc/common.cuh -- shared helpers for the sm_100a kernels of the CUDA layer.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vit_cuda_layer.h"

namespace vitcu {

// Records file:line + message for vitcu_last_error(); returns the code.
int set_error(int code, const char *file, int line, const char *what);
void count_launch();
// device flag raised by kernels whose mbarrier wait ran out of patience
uint32_t *watchdog_flag();

#define VITCU_TRY(expr)                                                        \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess)                                                 \
            return ::vitcu::set_error((int)_e, __FILE__, __LINE__,             \
                                      cudaGetErrorString(_e));                 \
    } while (0)

#define VITCU_REQUIRE(cond, msg)                                               \
    do {                                                                       \
        if (!(cond))                                                           \
            return ::vitcu::set_error(VITCU_E_ARG, __FILE__, __LINE__, msg);   \
    } while (0)

// after a <<<>>> launch
#define VITCU_LAUNCHED()                                                       \
    do {                                                                       \
        ::vitcu::count_launch();                                               \
        VITCU_TRY(cudaGetLastError());                                         \
    } while (0)

static inline cudaStream_t as_stream(vitcu_stream s) { return (cudaStream_t)s; }

constexpr int kEmbed = 768;
constexpr int kHeads = 12;
constexpr int kHeadDim = 64;
constexpr int kPatch = 16;

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// exact-erf GELU, the reference's form (R/ViT_seq.c:283-286, R/ll.cl:3-5)
__device__ __forceinline__ float gelu_erf(float x)
{
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}

} // namespace vitcu
