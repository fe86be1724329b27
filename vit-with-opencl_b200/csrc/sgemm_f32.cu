// csrc/sgemm_f32.cu -- FP32 SIMT GEMM  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue).
//
// The FP32 path of the engine (BASELINE config 2: logits within 1e-4 of
// ViT_seq) keeps every product and sum in IEEE fp32 on the CUDA cores.  This
// replaces the reference's 8x8-tile OpenCL kernels linear_layer (R/ll.cl:7-70)
// and QKV (R/multihead.cl:3-63) and the direct conv (R/conv2d.cl:1-36, via the
// patch gather); the oracle is linear_layer_seq (R/ViT_seq.c:295-309).
// R/ = /root/reference/MulticoreMainProject/.
//
// Design: BMxBNx16 block tile, 256 threads, TMxTN register tile per thread
// (8x8 for the 128x128 tile: 64 FMAs per 16 shared-memory loads), both operands
// K-major in global memory, loaded as 128-bit vectors along K and transposed
// into shared memory, register-staged double buffering so the next tile's
// global loads overlap the current tile's FMAs.  The epilogue fuses bias, exact
// GELU, the residual add and the patch-embedding row remap + position add, so
// no separate elementwise kernel (R/layer_norm.cl:55-65) ever runs.
#include "common.cuh"

using namespace vitcu;

namespace {

struct EpiParams {
    int M, N, K;
    size_t lda, ldc;
    int epilogue;
    const float *bias;
    const float *residual;
    const float *pos;
    int patches, tokens;
    int out_bf16;
};

__device__ __forceinline__ void epilogue_store4(const EpiParams &p, void *C, int row, int col, float4 acc)
{
    // row < M and col + 3 < N guaranteed by the caller
    const float4 b = *reinterpret_cast<const float4 *>(p.bias + col);
    float4 v = make_float4(acc.x + b.x, acc.y + b.y, acc.z + b.z, acc.w + b.w);
    size_t orow = (size_t)row;
    if (p.epilogue == VITCU_EPI_BIAS_GELU) {
        v.x = gelu_erf(v.x);
        v.y = gelu_erf(v.y);
        v.z = gelu_erf(v.z);
        v.w = gelu_erf(v.w);
    } else if (p.epilogue == VITCU_EPI_BIAS_RESIDUAL) {
        const float4 r = *reinterpret_cast<const float4 *>(p.residual + (size_t)row * p.ldc + col);
        v.x += r.x;
        v.y += r.y;
        v.z += r.z;
        v.w += r.w;
    } else if (p.epilogue == VITCU_EPI_PATCH_EMBED) {
        const int img = row / p.patches, pi = row - img * p.patches;
        orow = (size_t)img * p.tokens + 1 + pi;
        const float4 e = *reinterpret_cast<const float4 *>(p.pos + (size_t)(1 + pi) * p.N + col);
        v.x += e.x;
        v.y += e.y;
        v.z += e.z;
        v.w += e.w;
    }
    if (p.out_bf16) {
        *reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(C) + orow * p.ldc + col) =
            make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    } else {
        *reinterpret_cast<float4 *>(reinterpret_cast<float *>(C) + orow * p.ldc + col) = v;
    }
}

constexpr int BK = 16;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) sgemm_kernel(const float *__restrict__ A, const float *__restrict__ W,
                                                    void *__restrict__ C, const EpiParams p)
{
    pdl_trigger();
    pdl_wait();
    static_assert(BM / TM * (BN / TN) == 256, "256 threads");
    static_assert(TM % 4 == 0 && TN % 4 == 0, "float4 register tiles");
    constexpr int PAD = 4;
    constexpr int A_LD4 = BM * BK / 4 / 256; // float4 loads per thread for the A tile
    constexpr int B_LD4 = BN * BK / 4 / 256;
    static_assert(A_LD4 >= 1 && B_LD4 >= 1, "tile too small for 256 threads");
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    float4 ra[A_LD4], rb[B_LD4];
    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_LD4; i++) {
            const int f = tid + i * 256, r = f >> 2, kq = f & 3;
            const int gr = m0 + r;
            ra[i] = gr < p.M ? *reinterpret_cast<const float4 *>(A + (size_t)gr * p.lda + k0 + kq * 4)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < B_LD4; i++) {
            const int f = tid + i * 256, r = f >> 2, kq = f & 3;
            const int gr = n0 + r;
            rb[i] = gr < p.N ? *reinterpret_cast<const float4 *>(W + (size_t)gr * p.K + k0 + kq * 4)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_LD4; i++) {
            const int f = tid + i * 256, r = f >> 2, kq = f & 3;
            As[buf][kq * 4 + 0][r] = ra[i].x;
            As[buf][kq * 4 + 1][r] = ra[i].y;
            As[buf][kq * 4 + 2][r] = ra[i].z;
            As[buf][kq * 4 + 3][r] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < B_LD4; i++) {
            const int f = tid + i * 256, r = f >> 2, kq = f & 3;
            Bs[buf][kq * 4 + 0][r] = rb[i].x;
            Bs[buf][kq * 4 + 1][r] = rb[i].y;
            Bs[buf][kq * 4 + 2][r] = rb[i].z;
            Bs[buf][kq * 4 + 3][r] = rb[i].w;
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++)
            acc[i][j] = 0.f;

    const int nk = p.K / BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; kt++) {
        const int buf = kt & 1;
        if (kt + 1 < nk)
            gload((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; k++) {
            float a[TM], b[TN];
            // register tile rows/cols are interleaved in groups of 4 across the
            // thread grid: conflict-free 128-bit shared loads
#pragma unroll
            for (int i = 0; i < TM / 4; i++)
                *reinterpret_cast<float4 *>(&a[i * 4]) =
                    *reinterpret_cast<const float4 *>(&As[buf][k][i * (BM / (TM / 4)) + ty * 4]);
#pragma unroll
            for (int j = 0; j < TN / 4; j++)
                *reinterpret_cast<float4 *>(&b[j * 4]) =
                    *reinterpret_cast<const float4 *>(&Bs[buf][k][j * (BN / (TN / 4)) + tx * 4]);
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++)
                    acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int row = m0 + (i / 4) * (BM / (TM / 4)) + ty * 4 + (i % 4);
        if (row >= p.M)
            continue;
#pragma unroll
        for (int j = 0; j < TN / 4; j++) {
            const int col = n0 + j * (BN / (TN / 4)) + tx * 4;
            if (col + 3 < p.N)
                epilogue_store4(p, C, row, col,
                                make_float4(acc[i][j * 4 + 0], acc[i][j * 4 + 1], acc[i][j * 4 + 2], acc[i][j * 4 + 3]));
        }
    }
}

// Small-M case (the classification head at batch <= 8): one warp per output feature, the weight
// row is read once (coalesced, 128-bit) and dotted with every input row; the 64x64 tile kernel
// would run M = 1 on 16 CTAs for 55 us.
__global__ void __launch_bounds__(256) gemv_rows_kernel(const float *__restrict__ A, const float *__restrict__ W,
                                                        float *__restrict__ C, const EpiParams p)
{
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= p.N)
        return;
    const float4 *w4 = reinterpret_cast<const float4 *>(W + (size_t)n * p.K);
    const int k4 = p.K / 4;
    if (k4 <= 256) {
        // the weight row (cold: the head's weights come from DRAM on every forward) is requested in one go -- a loop of
        // dependent load / fma iterations pays that latency once per iteration -- and held in registers for every input row
        float4 w[8];
        const float bias = p.bias[n];
#pragma unroll
        for (int i = 0; i < 8; i++)
            w[i] = lane + 32 * i < k4 ? __ldg(w4 + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int m = 0; m < p.M; m++) {
            const float4 *a4 = reinterpret_cast<const float4 *>(A + (size_t)m * p.lda);
            float4 a[8];
#pragma unroll
            for (int i = 0; i < 8; i++)
                a[i] = lane + 32 * i < k4 ? a4[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) { // same order of accumulation as the loop below
                acc = fmaf(a[i].x, w[i].x, acc);
                acc = fmaf(a[i].y, w[i].y, acc);
                acc = fmaf(a[i].z, w[i].z, acc);
                acc = fmaf(a[i].w, w[i].w, acc);
            }
            acc = warp_sum(acc);
            if (lane == 0)
                C[(size_t)m * p.ldc + n] = acc + bias;
        }
        return;
    }
    for (int m = 0; m < p.M; m++) {
        const float4 *a4 = reinterpret_cast<const float4 *>(A + (size_t)m * p.lda);
        float acc = 0.f;
        for (int k = lane; k < k4; k += 32) {
            const float4 a = a4[k], w = __ldg(w4 + k);
            acc = fmaf(a.x, w.x, acc);
            acc = fmaf(a.y, w.y, acc);
            acc = fmaf(a.z, w.z, acc);
            acc = fmaf(a.w, w.w, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0)
            C[(size_t)m * p.ldc + n] = acc + p.bias[n];
    }
}

} // namespace

extern "C" int vitcu_sgemm(const float *A, const float *W, void *C, const vitcu_gemm_desc *d, vitcu_stream s)
{
    VITCU_REQUIRE(A && W && C && d, "NULL argument");
    VITCU_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "empty GEMM");
    VITCU_REQUIRE(d->K % BK == 0 && d->N % 4 == 0 && d->lda % 4 == 0 && d->ldc % 4 == 0,
                  "sgemm needs K % 16 == 0 and N, lda, ldc % 4 == 0");
    VITCU_REQUIRE(d->bias, "bias is required");
    VITCU_REQUIRE(d->epilogue != VITCU_EPI_BIAS_RESIDUAL || d->residual, "residual pointer missing");
    VITCU_REQUIRE(d->epilogue != VITCU_EPI_PATCH_EMBED || (d->pos && d->patches > 0 && d->tokens > d->patches),
                  "patch-embed epilogue needs pos, patches, tokens");
    EpiParams p;
    p.M = d->M;
    p.N = d->N;
    p.K = d->K;
    p.lda = d->lda ? d->lda : (size_t)d->K;
    p.ldc = d->ldc ? d->ldc : (size_t)d->N;
    p.epilogue = d->epilogue;
    p.bias = d->bias;
    p.residual = d->residual;
    p.pos = d->pos;
    p.patches = d->patches;
    p.tokens = d->tokens;
    p.out_bf16 = d->out_bf16;
    if (p.M <= 8 && p.epilogue == VITCU_EPI_BIAS && !p.out_bf16) {
        VITCU_TRY(launch_kernel(gemv_rows_kernel, (p.N + 7) / 8, 256, 0, as_stream(s), A, W, reinterpret_cast<float *>(C), p));
        VITCU_LAUNCHED_KIND(LK_SGEMM);
        return 0;
    }
    // Big tile when it still fills the 148 SMs, small tile for the batch-1 /
    // head shapes where parallelism matters more than reuse.
    const long big_ctas = (long)((p.M + 127) / 128) * ((p.N + 127) / 128);
    if (big_ctas >= 148) {
        dim3 grid((p.N + 127) / 128, (p.M + 127) / 128);
        VITCU_TRY(launch_kernel(sgemm_kernel<128, 128, 8, 8>, grid, 256, 0, as_stream(s), A, W, C, p));
    } else {
        dim3 grid((p.N + 63) / 64, (p.M + 63) / 64);
        VITCU_TRY(launch_kernel(sgemm_kernel<64, 64, 4, 4>, grid, 256, 0, as_stream(s), A, W, C, p));
    }
    VITCU_LAUNCHED_KIND(LK_SGEMM);
    return 0;
}
