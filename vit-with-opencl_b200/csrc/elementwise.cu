// csrc/elementwise.cu -- the HBM-bound data-movement kernels of the path:
// weight packing, patch gather, class-token rows, LayerNorm, row softmax.
// All are vectorised (128-bit) and coalesced; none has data reuse, so none
// stages through shared memory.  R/ = /root/reference/MulticoreMainProject/.
#include "common.cuh"
#include <cuda_fp8.h>
#include <cuda_fp16.h>

using namespace vitcu;

// ---------------------------------------------------------------------------
// fp32 -> bf16 (one-time weight packing; also used by tests)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float4 *__restrict__ src,
                                                          uint2 *__restrict__ dst, size_t n4,
                                                          const float *__restrict__ src_tail,
                                                          __nv_bfloat16 *__restrict__ dst_tail, int tail)
{
    pdl_trigger();
    pdl_wait();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        float4 v = src[i];
        dst[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < tail)
        dst_tail[threadIdx.x] = __float2bfloat16_rn(src_tail[threadIdx.x]);
}

// ---------------------------------------------------------------------------
// fp32 -> three bf16 pieces (x = x1 + x2 + x3 to 24 mantissa bits).  The FP32 path of the engine
// runs its GEMMs on the BF16 tensor cores as six piece products with FP32 accumulation
// (a3 w1 + a2 w2 + a1 w3 + a2 w1 + a1 w2 + a1 w1: relative error ~1e-7, below a plain fp32 GEMM's
// summation error), see vitcu_gemm_bf16x3.  Output row r = [x1 | x2 | x3], each K wide.
// ---------------------------------------------------------------------------
template <bool kGelu>
__global__ void __launch_bounds__(256) split3_kernel(const float *__restrict__ x, size_t ld, __nv_bfloat16 *__restrict__ out,
                                                     size_t rows, int K)
{
    pdl_trigger();
    pdl_wait();
    const size_t k4 = K / 4, total = rows * k4;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const size_t r = i / k4;
        const int c = (int)(i - r * k4) * 4;
        float4 v = *reinterpret_cast<const float4 *>(x + r * ld + c);
        if (kGelu) // exact-erf form (R/ViT_seq.c:283-286): the GELU of an fc1 whose K slices met in global memory
            v = make_float4(gelu_erf(v.x), gelu_erf(v.y), gelu_erf(v.z), gelu_erf(v.w));
        split3_store4(out + r * 3 * (size_t)K, K, c, v);
    }
}

// ---------------------------------------------------------------------------
// patch gather: [B,3,img,img] -> [B*P, 3*ps*ps], column = (c*ps + kh)*ps + kw (ps = patch side, 16 or
// 32), the accumulation order of Conv2d_seq (R/ViT_seq.c:37-48).  One thread moves one 16-byte
// chunk (4 kw values): ps/4 threads cover a contiguous image run of one patch row.
// ---------------------------------------------------------------------------
template <bool kBf16>
__global__ void __launch_bounds__(256) patch_gather_kernel(const float *__restrict__ images,
                                                           void *__restrict__ patches, int batch,
                                                           int img, int side, int ps)
{
    pdl_trigger();
    pdl_wait();
    const int P = side * side;
    const int K4 = 3 * ps * ps / 4, ps4 = ps / 4; // float4 chunks per patch row of the output / per kernel row
    const size_t total4 = (size_t)batch * P * K4;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total4; i += stride) {
        const int col4 = (int)(i % K4);
        const size_t row = i / K4;
        const int p = (int)(row % P);
        const int b = (int)(row / P);
        const int c = col4 / (ps * ps4), kh = (col4 / ps4) % ps, kw4 = col4 % ps4;
        const int ph = p / side, pw = p % side;
        const float4 v = *reinterpret_cast<const float4 *>(
            images + (((size_t)b * 3 + c) * img + ph * ps + kh) * img + pw * ps + kw4 * 4);
        if (kBf16) {
            reinterpret_cast<uint2 *>(patches)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        } else {
            reinterpret_cast<float4 *>(patches)[i] = v;
        }
    }
}

// x[b*T, :] = cls + pos[0, :]   (R/ViT_seq.c:83-118; R/conv2d.cl:39-80, t == 0)
__global__ void __launch_bounds__(256) cls_rows_kernel(float *__restrict__ x, const float *__restrict__ cls,
                                                       const float *__restrict__ pos, int tokens, int cols)
{
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x, i = threadIdx.x; // cols/4 threads x float4 (192 for the 768-wide model)
    const float4 c = reinterpret_cast<const float4 *>(cls)[i];
    const float4 p = reinterpret_cast<const float4 *>(pos)[i];
    reinterpret_cast<float4 *>(x + (size_t)b * tokens * cols)[i] =
        make_float4(c.x + p.x, c.y + p.y, c.z + p.z, c.w + p.w);
}

// every token row before an accumulate-mode patch embedding (vitcu_token_rows_init): row 0 as above, row t > 0 = pos[t]
__global__ void __launch_bounds__(256) token_rows_init_kernel(float *__restrict__ x, const float *__restrict__ cls,
                                                              const float *__restrict__ pos, int tokens, int cols)
{
    pdl_trigger();
    pdl_wait();
    const int t = blockIdx.x, b = blockIdx.y, i = threadIdx.x;
    float4 p = reinterpret_cast<const float4 *>(pos + (size_t)t * cols)[i];
    if (t == 0) {
        const float4 c = reinterpret_cast<const float4 *>(cls)[i];
        p = make_float4(c.x + p.x, c.y + p.y, c.z + p.z, c.w + p.w);
    }
    reinterpret_cast<float4 *>(x + ((size_t)b * tokens + t) * cols)[i] = p;
}

// ---------------------------------------------------------------------------
// LayerNorm, one warp per row of NV*128 columns (NV = 6 for the 768-wide model, 3 / 8 for 384 / 1024):
// NV x float4 per lane held in registers,
// shuffle reductions, statistics in fp32 (mean first, then centred variance,
// which is at least as accurate as the reference's E[x^2]-mean^2 single pass,
// R/ViT_seq.c:126-135), eps 1e-6, output fp32 or bf16.
// Algorithmic bytes per row: 768*4 read + 768*(4|2) written.
// ---------------------------------------------------------------------------
// kEarly (a few hundred rows: batch-1 latency): gamma / beta are requested together with the row, ahead of the
// reductions, instead of as a second memory round trip behind them -- at batch 1 the weights of a forward do not stay
// in the L2, so that second trip went to DRAM: BF16 batch-1 forward 0.492 -> 0.460 ms, FP32 0.917 -> 0.819 ms (same box).
// Costs 12 more float4 registers per lane, which the bandwidth-bound many-row launches keep for occupancy.
// Measured on top of it and NOT kept (+-0.4 % or worse): L2 prefetches of gamma / beta, the head's weight rows, the
// position rows and the conv filter boxes ahead of griddepcontrol.wait; the head gemv and the softmax with all their
// loads up front, the zero-fill issued behind the row loads and tensor-map prefetches ahead of the wait (BF16 +1.4 %).
template <int kOut, int NV, bool kEarly = false> // kOut 0: fp32, 1: bf16, 2: three bf16 pieces [rows, 3*cols]
__global__ void __launch_bounds__(256) layernorm_kernel(const float *__restrict__ x, size_t x_row_stride,
                                                        void *__restrict__ y, const float *__restrict__ gamma,
                                                        const float *__restrict__ beta, int rows, int rev,
                                                        uint4 *__restrict__ zero_ptr, size_t zero_n16)
{
    pdl_trigger();
    pdl_wait();
    // vitcu_layernorm_zero: the launch also clears the output buffer of the accumulate-mode GEMM that follows
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < zero_n16; i += (size_t)gridDim.x * blockDim.x)
        zero_ptr[i] = make_uint4(0u, 0u, 0u, 0u);
    const int lane = threadIdx.x & 31;
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows)
        return;
    if (rev) // last rows first: they are what the producing GEMM wrote most recently, still in L2
        row = rows - 1 - row;
    const float4 *xr = reinterpret_cast<const float4 *>(x + (size_t)row * x_row_stride);
    constexpr int kCols = NV * 128;
    const float4 *g4 = reinterpret_cast<const float4 *>(gamma);
    const float4 *b4 = reinterpret_cast<const float4 *>(beta);
    float4 v[NV], ge[kEarly ? NV : 1], be[kEarly ? NV : 1];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++)
        v[i] = xr[lane + 32 * i];
    if (kEarly) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            ge[kEarly ? i : 0] = g4[lane + 32 * i];
            be[kEarly ? i : 0] = b4[lane + 32 * i];
        }
    }
#pragma unroll
    for (int i = 0; i < NV; i++)
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / kCols);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float var = warp_sum(q) * (1.0f / kCols);
    const float inv_std = 1.0f / sqrtf(var + 1e-6f);
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const float4 g = kEarly ? ge[kEarly ? i : 0] : g4[lane + 32 * i], bb = kEarly ? be[kEarly ? i : 0] : b4[lane + 32 * i];
        float4 o;
        o.x = (v[i].x - mean) * inv_std * g.x + bb.x;
        o.y = (v[i].y - mean) * inv_std * g.y + bb.y;
        o.z = (v[i].z - mean) * inv_std * g.z + bb.z;
        o.w = (v[i].w - mean) * inv_std * g.w + bb.w;
        if (kOut == 1) {
            reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(y) + (size_t)row * kCols)[lane + 32 * i] =
                make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        } else if (kOut == 2) {
            split3_store4(reinterpret_cast<__nv_bfloat16 *>(y) + (size_t)row * 3 * kCols, kCols, (lane + 32 * i) * 4, o);
        } else {
            reinterpret_cast<float4 *>(reinterpret_cast<float *>(y) + (size_t)row * kCols)[lane + 32 * i] = o;
        }
    }
}

// ---------------------------------------------------------------------------
// LayerNorm folded into the GEMM that follows it (BF16 path; include/vit_cuda_layer.h: vitcu_gemm_desc).
// One-time weight folding, one warp per output feature n:
//   w_folded[n,k] = bf16(gamma[k] W[n,k]);  colsum[n] = sum_k w_folded[n,k] (of the ROUNDED values: it multiplies
//   the mean that the tensor core's product of the rounded values contains);  bias_folded[n] = bias[n] + sum_k beta[k] W[n,k]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ln_fold_weights_kernel(const float *__restrict__ W, const float *__restrict__ gamma,
                                                              const float *__restrict__ beta, const float *__restrict__ bias,
                                                              __nv_bfloat16 *__restrict__ wf, float *__restrict__ colsum,
                                                              float *__restrict__ bias_f, int N, int K)
{
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= N)
        return;
    const float4 *w4 = reinterpret_cast<const float4 *>(W + (size_t)n * K);
    const float4 *g4 = reinterpret_cast<const float4 *>(gamma), *b4 = reinterpret_cast<const float4 *>(beta);
    uint2 *o = reinterpret_cast<uint2 *>(wf + (size_t)n * K);
    float cs = 0.f, bs = 0.f;
    for (int i = lane; i < K / 4; i += 32) {
        const float4 w = w4[i], g = g4[i], b = b4[i];
        const __nv_bfloat162 lo = __floats2bfloat162_rn(w.x * g.x, w.y * g.y), hi = __floats2bfloat162_rn(w.z * g.z, w.w * g.w);
        o[i] = make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
        cs += (__low2float(lo) + __high2float(lo)) + (__low2float(hi) + __high2float(hi));
        bs += (w.x * b.x + w.y * b.y) + (w.z * b.z + w.w * b.w);
    }
    cs = warp_sum(cs);
    bs = warp_sum(bs);
    if (lane == 0) {
        colsum[n] = cs;
        bias_f[n] = bias[n] + bs;
    }
}

// Entry of the folded chain: xb = bf16(x) and (sum x, sum x^2) of every row into slot 0 of the partial-sum table
// (the other slots are zeroed), one warp per row.  6 bytes per element like the LayerNorm kernel it replaces, but
// only once per forward: the out-proj / fc2 epilogues produce the same outputs for all later LayerNorms.
template <int NV>
__global__ void __launch_bounds__(256) rowstats_cast_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ xb,
                                                            float2 *__restrict__ stats, int rows, int slots)
{
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows)
        return;
    constexpr int kCols = NV * 128;
    const float4 *xr = reinterpret_cast<const float4 *>(x + (size_t)row * kCols);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const float4 v = xr[lane + 32 * i];
        s1 += (v.x + v.y) + (v.z + v.w);
        s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        reinterpret_cast<uint2 *>(xb + (size_t)row * kCols)[lane + 32 * i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane < slots)
        stats[(size_t)lane * rows + row] = lane == 0 ? make_float2(s1, s2) : make_float2(0.f, 0.f);
}

// ---------------------------------------------------------------------------
// FP8 (E4M3) path: per-tensor absolute maxima (scales are 448 / amax with head room) and weight quantisation
// ---------------------------------------------------------------------------
__device__ __forceinline__ void absmax_commit(float m, float *out)
{
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f)
        atomicMax(reinterpret_cast<int *>(out), __float_as_int(m)); // non-negative floats order like their bit patterns
}
__global__ void __launch_bounds__(256) absmax_f32_kernel(const float *__restrict__ x, const float *__restrict__ gamma, size_t rows,
                                                         int K, float *__restrict__ out)
{
    pdl_trigger();
    pdl_wait();
    float m = 0.f;
    const size_t n4 = rows * (size_t)(K / 4);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4 *>(x)[i];
        if (gamma) {
            const float4 g = reinterpret_cast<const float4 *>(gamma)[i % (size_t)(K / 4)];
            v = make_float4(v.x * g.x, v.y * g.y, v.z * g.z, v.w * g.w);
        }
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    absmax_commit(m, out);
}
__global__ void __launch_bounds__(256) absmax_bf16_kernel(const uint4 *__restrict__ x, size_t n8, float *__restrict__ out)
{
    pdl_trigger();
    pdl_wait();
    float m = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = x[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++)
            m = fmaxf(m, fmaxf(fabsf(__uint_as_float(w[k] << 16)), fabsf(__uint_as_float(w[k] & 0xffff0000u))));
    }
    absmax_commit(m, out);
}
__device__ __forceinline__ uint32_t tc_pack_e4m3x4(float a, float b, float c, float d)
{
    uint16_t lo, hi;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
    return static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
}
__device__ __forceinline__ float e4m3_to_float(uint32_t byte)
{
    __half_raw h = __nv_cvt_fp8_to_halfraw(static_cast<__nv_fp8_storage_t>(byte), __NV_E4M3);
    return __half2float(*reinterpret_cast<__half *>(&h));
}
// one warp per output feature n (see ln_fold_weights_kernel); q = e4m3(gamma W scale), colsum of the de-quantised row
__global__ void __launch_bounds__(256) fp8_quant_weights_kernel(const float *__restrict__ W, const float *__restrict__ gamma,
                                                                const float *__restrict__ beta, const float *__restrict__ bias,
                                                                float scale, uint8_t *__restrict__ q, float *__restrict__ colsum,
                                                                float *__restrict__ bias_f, int N, int K)
{
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= N)
        return;
    const float4 *w4 = reinterpret_cast<const float4 *>(W + (size_t)n * K);
    uint32_t *o = reinterpret_cast<uint32_t *>(q + (size_t)n * K);
    float cs = 0.f, bs = 0.f;
    for (int i = lane; i < K / 4; i += 32) {
        const float4 w = w4[i];
        float4 g = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gamma)
            g = reinterpret_cast<const float4 *>(gamma)[i];
        if (beta)
            b = reinterpret_cast<const float4 *>(beta)[i];
        const uint32_t pk = tc_pack_e4m3x4(w.x * g.x * scale, w.y * g.y * scale, w.z * g.z * scale, w.w * g.w * scale);
        o[i] = pk;
        cs += (e4m3_to_float(pk & 0xff) + e4m3_to_float((pk >> 8) & 0xff)) + (e4m3_to_float((pk >> 16) & 0xff) + e4m3_to_float(pk >> 24));
        bs += (w.x * b.x + w.y * b.y) + (w.z * b.z + w.w * b.w);
    }
    cs = warp_sum(cs);
    bs = warp_sum(bs);
    if (lane == 0) {
        if (colsum)
            colsum[n] = cs / scale;
        if (bias_f)
            bias_f[n] = bias[n] + bs;
    }
}

// ---------------------------------------------------------------------------
// row softmax (R/ViT_seq.c:372-397): one 256-thread block per row
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float *__restrict__ logits,
                                                           float *__restrict__ probs, int n)
{
    pdl_trigger();
    pdl_wait();
    __shared__ float red[8];
    const float *l = logits + (size_t)blockIdx.x * n;
    float *p = probs + (size_t)blockIdx.x * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float m = -INFINITY;
    for (int i = tid; i < n; i += 256)
        m = fmaxf(m, l[i]);
    m = warp_max(m);
    if (lane == 0)
        red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int i = 1; i < 8; i++)
        m = fmaxf(m, red[i]);
    __syncthreads();
    float s = 0.f;
    for (int i = tid; i < n; i += 256)
        s += expf(l[i] - m);
    s = warp_sum(s);
    if (lane == 0)
        red[warp] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++)
        s += red[i];
    for (int i = tid; i < n; i += 256)
        p[i] = expf(l[i] - m) / s;
}

// ---------------------------------------------------------------------------
// top-k per row: the reference's host-side argmax scan (R/Main.c:59-72) moved to the device so a
// caller that only wants labels reads k*8 bytes per image instead of 4 KB.  One warp per row; the
// row lives in registers (<= 32 values per lane up to 1024 columns, strided beyond that); round r
// picks the largest element that comes after round r-1's pick in the order (value descending,
// index ascending) -- the first maximum wins, as in Main.c's strict `>` scan.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) topk_rows_kernel(const float *__restrict__ x, int rows, int cols, int k,
                                                        int *__restrict__ idx, float *__restrict__ val)
{
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows)
        return;
    const float *r = x + (size_t)row * cols;
    float pv = INFINITY;
    int pi = -1;
    for (int t = 0; t < k; t++) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = lane; c < cols; c += 32) {
            const float v = r[c];
            // NaNs never win; "after the previous pick": smaller value, or same value and larger index
            const bool after = v < pv || (v == pv && c > pi);
            if (after && (v > bv || (v == bv && c < bi))) {
                bv = v;
                bi = c;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) {
                bv = ov;
                bi = oi;
            }
        }
        if (lane == 0) {
            idx[(size_t)row * k + t] = bi == 0x7fffffff ? -1 : bi;
            val[(size_t)row * k + t] = bv;
        }
        pv = bv;
        pi = bi;
    }
}

static inline int grid_for(size_t work_items, int block, int max_blocks = 148 * 16)
{
    size_t g = (work_items + block - 1) / block;
    if (g < 1)
        g = 1;
    if (g > (size_t)max_blocks)
        g = max_blocks;
    return (int)g;
}

extern "C" {

int vitcu_f32_to_bf16(const float *src, vitcu_bf16 *dst, size_t n, vitcu_stream s)
{
    VITCU_REQUIRE(src && dst, "NULL buffer");
    VITCU_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0, "unaligned buffer");
    const size_t n4 = n / 4;
    const int tail = (int)(n - n4 * 4);
    VITCU_TRY(launch_kernel(f32_to_bf16_kernel, grid_for(n4, 256), 256, 0, as_stream(s), reinterpret_cast<const float4 *>(src), reinterpret_cast<uint2 *>(dst), n4, src + n4 * 4, reinterpret_cast<__nv_bfloat16 *>(dst) + n4 * 4, tail));
    VITCU_LAUNCHED();
    return 0;
}

int vitcu_split3(const float *x, size_t ld, vitcu_bf16 *out, size_t rows, int K, vitcu_stream s)
{
    VITCU_REQUIRE(x && out && rows > 0 && K > 0 && K % 4 == 0 && ld % 4 == 0 && ld >= (size_t)K, "bad argument");
    VITCU_TRY(launch_kernel(split3_kernel<false>, grid_for(rows * (size_t)(K / 4), 256, 148 * 32), 256, 0, as_stream(s), x, ld, reinterpret_cast<__nv_bfloat16 *>(out), rows, K));
    VITCU_LAUNCHED();
    return 0;
}

int vitcu_split3_gelu(const float *x, size_t ld, vitcu_bf16 *out, size_t rows, int K, vitcu_stream s)
{
    VITCU_REQUIRE(x && out && rows > 0 && K > 0 && K % 4 == 0 && ld % 4 == 0 && ld >= (size_t)K, "bad argument");
    VITCU_TRY(launch_kernel(split3_kernel<true>, grid_for(rows * (size_t)(K / 4), 256, 148 * 32), 256, 0, as_stream(s), x, ld, reinterpret_cast<__nv_bfloat16 *>(out), rows, K));
    VITCU_LAUNCHED();
    return 0;
}

int vitcu_patch_gather_ex(const float *images, void *patches, int batch, int img, int patch, int out_bf16,
                          vitcu_stream s)
{
    VITCU_REQUIRE(images && patches, "NULL buffer");
    VITCU_REQUIRE(patch == 16 || patch == 32, "patch side must be 16 or 32");
    VITCU_REQUIRE(batch > 0 && img > 0 && img % patch == 0, "image side must be a positive multiple of the patch side");
    const int side = img / patch;
    const size_t total4 = (size_t)batch * side * side * (3 * patch * patch / 4);
    const int grid = grid_for(total4, 256, 148 * 32);
    if (out_bf16)
        VITCU_TRY(launch_kernel(patch_gather_kernel<true>, grid, 256, 0, as_stream(s), images, patches, batch, img, side, patch));
    else
        VITCU_TRY(launch_kernel(patch_gather_kernel<false>, grid, 256, 0, as_stream(s), images, patches, batch, img, side, patch));
    VITCU_LAUNCHED();
    return 0;
}
int vitcu_patch_gather(const float *images, void *patches, int batch, int img, int out_bf16, vitcu_stream s)
{
    return vitcu_patch_gather_ex(images, patches, batch, img, kPatch, out_bf16, s);
}

int vitcu_cls_rows_ex(float *x, const float *cls, const float *pos, int batch, int tokens, int cols, vitcu_stream s)
{
    VITCU_REQUIRE(x && cls && pos && batch > 0 && tokens > 0, "bad argument");
    VITCU_REQUIRE(cols > 0 && cols % 4 == 0 && cols <= 1024, "row width must be a multiple of 4, at most 1024");
    VITCU_TRY(launch_kernel(cls_rows_kernel, batch, cols / 4, 0, as_stream(s), x, cls, pos, tokens, cols));
    VITCU_LAUNCHED();
    return 0;
}
int vitcu_token_rows_init(float *x, const float *cls, const float *pos, int batch, int tokens, int cols, vitcu_stream s)
{
    VITCU_REQUIRE(x && cls && pos && batch > 0 && tokens > 0, "bad argument");
    VITCU_REQUIRE(cols > 0 && cols % 4 == 0 && cols <= 1024, "row width must be a multiple of 4, at most 1024");
    VITCU_TRY(launch_kernel(token_rows_init_kernel, dim3(tokens, batch), cols / 4, 0, as_stream(s), x, cls, pos, tokens, cols));
    VITCU_LAUNCHED();
    return 0;
}
int vitcu_cls_rows(float *x, const float *cls, const float *pos, int batch, int tokens, vitcu_stream s)
{
    return vitcu_cls_rows_ex(x, cls, pos, batch, tokens, kEmbed, s);
}

extern "C++" {
template <int NV>
static int launch_layernorm(const float *x, size_t x_row_stride, void *y, int y_bf16, const float *gamma, const float *beta,
                            int rows, vitcu_stream s, void *zero_ptr = nullptr, size_t zero_bytes = 0)
{
    int grid = (rows + 7) / 8;
    // extra CTAs (no rows of their own) when there is a buffer to clear: 16 stores of 16 bytes per thread, at most one wave
    const size_t zero_n16 = zero_bytes / 16;
    const int zgrid = (int)((zero_n16 + 256 * 16 - 1) / (256 * 16));
    if (zgrid > grid)
        grid = zgrid < 148 ? zgrid : (grid > 148 ? grid : 148);
    uint4 *zp = reinterpret_cast<uint4 *>(zero_ptr);
    static const int rev = !(getenv("VITCU_SERPENTINE") && atoi(getenv("VITCU_SERPENTINE")) == 0);
    const bool early = rows <= 148 * 8; // one wave of CTAs: a latency-bound launch
    if (y_bf16 == 1)
        VITCU_TRY(launch_kernel(early ? layernorm_kernel<1, NV, true> : layernorm_kernel<1, NV, false>, grid, 256, 0, as_stream(s), x, x_row_stride, y, gamma, beta, rows, rev, zp, zero_n16));
    else if (y_bf16 == 2)
        VITCU_TRY(launch_kernel(early ? layernorm_kernel<2, NV, true> : layernorm_kernel<2, NV, false>, grid, 256, 0, as_stream(s), x, x_row_stride, y, gamma, beta, rows, rev, zp, zero_n16));
    else
        VITCU_TRY(launch_kernel(early ? layernorm_kernel<0, NV, true> : layernorm_kernel<0, NV, false>, grid, 256, 0, as_stream(s), x, x_row_stride, y, gamma, beta, rows, rev, zp, zero_n16));
    VITCU_LAUNCHED_KIND(LK_LAYERNORM);
    return 0;
}
} // extern "C++"

int vitcu_layernorm_zero(const float *x, size_t x_row_stride, void *y, int y_bf16, const float *gamma, const float *beta,
                         int rows, int cols, void *zero_ptr, size_t zero_bytes, vitcu_stream s)
{
    VITCU_REQUIRE(x && y && gamma && beta && rows > 0, "bad argument");
    VITCU_REQUIRE(x_row_stride % 4 == 0 && x_row_stride >= (size_t)cols, "row stride must be >= the row width and a multiple of 4");
    VITCU_REQUIRE(zero_bytes == 0 || (zero_ptr && ((uintptr_t)zero_ptr & 15) == 0 && zero_bytes % 16 == 0),
                  "the buffer to clear must be 16-byte aligned and a multiple of 16 bytes long");
    switch (cols) {
    case 384:
        return launch_layernorm<3>(x, x_row_stride, y, y_bf16, gamma, beta, rows, s, zero_ptr, zero_bytes);
    case 768:
        return launch_layernorm<6>(x, x_row_stride, y, y_bf16, gamma, beta, rows, s, zero_ptr, zero_bytes);
    case 1024:
        return launch_layernorm<8>(x, x_row_stride, y, y_bf16, gamma, beta, rows, s, zero_ptr, zero_bytes);
    default:
        return set_error(VITCU_E_ARG, __FILE__, __LINE__, "LayerNorm row width must be 384, 768 or 1024");
    }
}
int vitcu_layernorm_ex(const float *x, size_t x_row_stride, void *y, int y_bf16, const float *gamma, const float *beta,
                       int rows, int cols, vitcu_stream s)
{
    return vitcu_layernorm_zero(x, x_row_stride, y, y_bf16, gamma, beta, rows, cols, nullptr, 0, s);
}
int vitcu_layernorm(const float *x, size_t x_row_stride, void *y, int y_bf16, const float *gamma,
                    const float *beta, int rows, vitcu_stream s)
{
    return vitcu_layernorm_ex(x, x_row_stride, y, y_bf16, gamma, beta, rows, kEmbed, s);
}

int vitcu_ln_fold_weights(const float *W, const float *gamma, const float *beta, const float *bias, vitcu_bf16 *w_folded,
                          float *colsum, float *bias_folded, int N, int K, vitcu_stream s)
{
    VITCU_REQUIRE(W && gamma && beta && bias && w_folded && colsum && bias_folded && N > 0 && K > 0 && K % 4 == 0, "bad argument");
    VITCU_TRY(launch_kernel(ln_fold_weights_kernel, (N + 7) / 8, 256, 0, as_stream(s), W, gamma, beta, bias,
                            reinterpret_cast<__nv_bfloat16 *>(w_folded), colsum, bias_folded, N, K));
    VITCU_LAUNCHED();
    return 0;
}

int vitcu_rowstats_cast(const float *x, vitcu_bf16 *xb, void *stats, int rows, int cols, int slots, vitcu_stream s)
{
    VITCU_REQUIRE(x && xb && stats && rows > 0 && slots > 0 && slots <= 32, "bad argument");
    float2 *st = reinterpret_cast<float2 *>(stats);
    __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(xb);
    const int grid = (rows + 7) / 8;
    switch (cols) {
    case 768:
        VITCU_TRY(launch_kernel(rowstats_cast_kernel<6>, grid, 256, 0, as_stream(s), x, o, st, rows, slots));
        break;
    case 1024:
        VITCU_TRY(launch_kernel(rowstats_cast_kernel<8>, grid, 256, 0, as_stream(s), x, o, st, rows, slots));
        break;
    case 256:
        VITCU_TRY(launch_kernel(rowstats_cast_kernel<2>, grid, 256, 0, as_stream(s), x, o, st, rows, slots));
        break;
    case 512:
        VITCU_TRY(launch_kernel(rowstats_cast_kernel<4>, grid, 256, 0, as_stream(s), x, o, st, rows, slots));
        break;
    default:
        return set_error(VITCU_E_ARG, __FILE__, __LINE__, "row width must be 256, 512, 768 or 1024");
    }
    VITCU_LAUNCHED();
    return 0;
}

int vitcu_absmax_f32(const float *x, const float *gamma, size_t rows, int K, float *out, vitcu_stream s)
{
    VITCU_REQUIRE(x && out && rows > 0 && K > 0 && K % 4 == 0, "bad argument");
    VITCU_TRY(launch_kernel(absmax_f32_kernel, grid_for(rows * (size_t)(K / 4), 256, 148 * 8), 256, 0, as_stream(s), x, gamma, rows, K, out));
    VITCU_LAUNCHED();
    return 0;
}
int vitcu_absmax_bf16(const vitcu_bf16 *x, size_t n, float *out, vitcu_stream s)
{
    VITCU_REQUIRE(x && out && n > 0 && n % 8 == 0 && ((uintptr_t)x & 15) == 0, "bad argument");
    VITCU_TRY(launch_kernel(absmax_bf16_kernel, grid_for(n / 8, 256, 148 * 8), 256, 0, as_stream(s), reinterpret_cast<const uint4 *>(x), n / 8, out));
    VITCU_LAUNCHED();
    return 0;
}
int vitcu_fp8_quant_weights(const float *W, const float *gamma, const float *beta, const float *bias, float scale, uint8_t *q,
                            float *colsum, float *bias_folded, int N, int K, vitcu_stream s)
{
    VITCU_REQUIRE(W && q && N > 0 && K > 0 && K % 4 == 0 && scale > 0.f, "bad argument");
    VITCU_REQUIRE(!bias_folded || bias, "bias_folded needs bias");
    VITCU_TRY(launch_kernel(fp8_quant_weights_kernel, (N + 7) / 8, 256, 0, as_stream(s), W, gamma, beta, bias, scale, q, colsum,
                            bias_folded, N, K));
    VITCU_LAUNCHED();
    return 0;
}

int vitcu_softmax_rows(const float *logits, float *probs, int rows, int n, vitcu_stream s)
{
    VITCU_REQUIRE(logits && probs && rows > 0 && n > 0, "bad argument");
    VITCU_TRY(launch_kernel(softmax_rows_kernel, rows, 256, 0, as_stream(s), logits, probs, n));
    VITCU_LAUNCHED();
    return 0;
}

int vitcu_topk_rows(const float *x, int rows, int cols, int k, int *idx, float *val, vitcu_stream s)
{
    VITCU_REQUIRE(x && idx && val && rows > 0 && cols > 0, "bad argument");
    VITCU_REQUIRE(k > 0 && k <= cols, "top-k needs 0 < k <= cols");
    VITCU_TRY(launch_kernel(topk_rows_kernel, (rows + 7) / 8, 256, 0, as_stream(s), x, rows, cols, k, idx, val));
    VITCU_LAUNCHED();
    return 0;
}

} // extern "C"
