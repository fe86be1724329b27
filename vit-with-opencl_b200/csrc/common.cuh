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
// per-kernel-family launch counters (vitcu_launch_count_of): the tests assert which kernel a shape dispatched to
enum LaunchKind { LK_GEMM_PAIR = 0, LK_GEMM_1CTA, LK_ATTN_TC, LK_ATTN_FLASH, LK_ATTN_SIMT, LK_SGEMM, LK_LAYERNORM,
                  LK_PATCH_EMBED_TC, LK_ATTN_DUO, LK_OTHER, LK_COUNT };
void count_launch_kind(int kind);
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
#define VITCU_LAUNCHED_KIND(kind)                                              \
    do {                                                                       \
        ::vitcu::count_launch();                                               \
        ::vitcu::count_launch_kind(kind);                                      \
        VITCU_TRY(cudaGetLastError());                                         \
    } while (0)
#define VITCU_LAUNCHED() VITCU_LAUNCHED_KIND(::vitcu::LK_OTHER)

static inline cudaStream_t as_stream(vitcu_stream s) { return (cudaStream_t)s; }

// ---------------------------------------------------------------------------
// Programmatic dependent launch.  Every kernel of the layer starts with pdl_trigger() -- the next
// kernel in the stream may be launched and run its prologue (barrier init, TMEM allocation,
// descriptor prefetch) while this one is still working -- and calls pdl_wait() before its first
// global-memory access, which returns once the preceding kernel has completed and its writes are
// visible.  Nothing before pdl_wait() may touch global memory another kernel writes.  The forward
// is 89 launches; at batch 1 the launch-to-launch latency is most of its time.
// VITCU_PDL=0 launches everything with full stream serialisation (the wait is then a no-op).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled(); // runtime.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

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

// fp32 -> three bf16 pieces, x = p1 + p2 + p3 to 24 mantissa bits (see elementwise.cu: split3_kernel); a row of the
// split form is [p1 | p2 | p3], each K wide
__device__ __forceinline__ void split3(float x, __nv_bfloat16 &p1, __nv_bfloat16 &p2, __nv_bfloat16 &p3)
{
    p1 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(p1);
    p2 = __float2bfloat16_rn(r1);
    p3 = __float2bfloat16_rn(r1 - __bfloat162float(p2));
}
__device__ __forceinline__ void split3_store4(__nv_bfloat16 *row, int K, int col, float4 v)
{
    __nv_bfloat16 a[4], b[4], c[4];
    split3(v.x, a[0], b[0], c[0]);
    split3(v.y, a[1], b[1], c[1]);
    split3(v.z, a[2], b[2], c[2]);
    split3(v.w, a[3], b[3], c[3]);
    *reinterpret_cast<uint2 *>(row + col) = *reinterpret_cast<const uint2 *>(a);
    *reinterpret_cast<uint2 *>(row + K + col) = *reinterpret_cast<const uint2 *>(b);
    *reinterpret_cast<uint2 *>(row + 2 * K + col) = *reinterpret_cast<const uint2 *>(c);
}

} // namespace vitcu
