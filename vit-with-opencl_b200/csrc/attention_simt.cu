// csrc/attention_simt.cu -- fused multi-head attention core on the CUDA cores
// (FP32 math; fp32 or bf16 storage).  One CTA owns a 64-query tile of one
// (image, head): S = Q K^T, the row softmax and O = P V all stay in shared
// memory / registers, so scores never touch HBM.  Replaces QKV_TO_SCOREV
// (R/multihead.cl:65-137), which re-read all of K and V from global memory for
// every (query, head) work-group and was capped at 256 keys; the oracle is
// R/ViT_seq.c:192-262.  R/ = /root/reference/MulticoreMainProject/.
//
// This is the attention of the FP32 path (BASELINE config 2) and the
// bring-up attention of the BF16 path; any token count that fits shared
// memory works (197 and 577 both do).
//
//   phase 1  Q tile -> smem, transposed [d][q]
//   phase 2  per 64-key block: K block -> smem transposed [d][key];
//            4x4 register tiles of S = (Q.K) * 1/8 (the scale is applied after
//            the dot product, like R/ViT_seq.c:211) -> St[key][q]
//   phase 3  column softmax of St (max, expf, sum, divide -- R/ViT_seq.c:216-234)
//   phase 4  per 64-key block: V block -> smem; 4x4 register tiles of O += P V
//   phase 5  O -> out[(b*T + q), h*64 + d]
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

using namespace vitcu;

namespace {

// QT queries per CTA; 256 threads hold 4x4 register tiles of S, so a staged key block has
// KB = 4096 / QT keys.  QT = 64 (KB = 64) is the throughput shape; QT = 16 (KB = 256) gives a small
// batch four times as many CTAs (batch 1: 13 x 12 = 156 instead of 48, one wave on 148 SMs).
__host__ __device__ constexpr int kKeysPerBlock(int qt) { return 4096 / qt; }

template <typename T>
struct Elem;
template <>
struct Elem<float> {
    static __device__ __forceinline__ float4 load4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
    static __device__ __forceinline__ void store4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
};
template <>
struct Elem<__nv_bfloat16> {
    static __device__ __forceinline__ float4 load4(const __nv_bfloat16 *p)
    {
        const uint2 u = *reinterpret_cast<const uint2 *>(p);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&u.x);
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162 *>(&u.y);
        const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
    static __device__ __forceinline__ void store4(__nv_bfloat16 *p, float4 v)
    {
        *reinterpret_cast<uint2 *>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
};

// kSplit: write the output as three bf16 pieces [rows, 3*embed] (the A operand form of the FP32 tensor-core GEMM,
// vitcu_gemm_bf16x3) instead of T -- saves the separate split kernel in front of the output projection
template <typename T, int QT, bool kSplit = false>
__global__ void __launch_bounds__(256) attention_simt_kernel(const T *__restrict__ qkv, void *__restrict__ out_, int tokens, int embed)
{
    constexpr int KB = kKeysPerBlock(QT);
    constexpr int LDQ = QT + 4;       // padded row of the [d][q] / [key][q] tiles
    constexpr int LDK = KB + 4;       // padded row of the transposed K block [d][key]
    constexpr int LDV = kHeadDim + 4; // padded row of the V block [key][d]
    constexpr int KVF = kHeadDim * LDK > KB * LDV ? kHeadDim * LDK : KB * LDV;
    constexpr int NTX = QT / 4;       // threads along the query axis of the S tile
    constexpr int PARTS = 256 / QT;   // key-interleaved softmax parts per query
    constexpr int OQ = QT / 16;       // queries per thread in the P V phase
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float smem[];
    const int nkb = (tokens + KB - 1) / KB;
    float *Qs = smem;                // [64 d][LDQ]
    float *KV = Qs + kHeadDim * LDQ; // K block transposed [64 d][LDK], later V block [KB key][LDV]
    float *St = KV + KVF;            // [nkb*KB keys][LDQ]
    float *red = St + (size_t)nkb * KB * LDQ; // [PARTS][QT]

    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * QT, head = blockIdx.y, img = blockIdx.z;
    const size_t ld = 3 * (size_t)embed; // embed = heads * 64 (768 for ViT-B)
    const T *base = qkv + (size_t)img * tokens * ld + head * kHeadDim;

    // ROWS x 64 tile -> dst[d][row]; each thread moves (4 consecutive d) of ROWS*16/256 rows
    auto load_transposed = [&](float *dst, int ldd, const T *src, int row0, int rows) {
        for (int f = tid; f < rows * 16; f += 256) {
            const int r = f >> 4, d4 = (f & 15) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < tokens)
                v = Elem<T>::load4(src + (size_t)(row0 + r) * ld + d4);
            dst[(d4 + 0) * ldd + r] = v.x;
            dst[(d4 + 1) * ldd + r] = v.y;
            dst[(d4 + 2) * ldd + r] = v.z;
            dst[(d4 + 3) * ldd + r] = v.w;
        }
    };

    load_transposed(Qs, LDQ, base, q0, QT);

    // ---- phase 2: S = Q K^T / 8 ------------------------------------------------
    const int tx = tid % NTX, ty = tid / NTX; // tx -> 4 queries, ty -> 4 keys
    for (int kb = 0; kb < nkb; kb++) {
        __syncthreads(); // previous block's readers are done with KV (and Qs is visible)
        load_transposed(KV, LDK, base + embed, kb * KB, KB);
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
                acc[i][j] = 0.f;
#pragma unroll 16
        for (int d = 0; d < kHeadDim; d++) {
            const float4 a = *reinterpret_cast<const float4 *>(&Qs[d * LDQ + tx * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&KV[d * LDK + ty * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int i = 0; i < 4; i++)
                    acc[j][i] = fmaf(av[i], bv[j], acc[j][i]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float *row = St + (size_t)(kb * KB + ty * 4 + j) * LDQ + tx * 4;
            *reinterpret_cast<float4 *>(row) =
                make_float4(acc[j][0] * 0.125f, acc[j][1] * 0.125f, acc[j][2] * 0.125f, acc[j][3] * 0.125f);
        }
    }
    __syncthreads();

    // ---- phase 3: softmax down each query column -------------------------------
    {
        const int q = tid % QT, part = tid / QT; // PARTS key-interleaved parts per query
        float m = -INFINITY;
        for (int j = part; j < tokens; j += PARTS)
            m = fmaxf(m, St[(size_t)j * LDQ + q]);
        red[part * QT + q] = m;
        __syncthreads();
        m = red[q];
#pragma unroll
        for (int i = 1; i < PARTS; i++)
            m = fmaxf(m, red[i * QT + q]);
        __syncthreads();
        float s = 0.f;
        for (int j = part; j < tokens; j += PARTS) {
            const float e = expf(St[(size_t)j * LDQ + q] - m);
            St[(size_t)j * LDQ + q] = e;
            s += e;
        }
        red[part * QT + q] = s;
        __syncthreads();
        s = 0.f;
        if (PARTS == 4) { // the summation order the 64-query version has always used
            s = (red[q] + red[QT + q]) + (red[2 * QT + q] + red[3 * QT + q]);
        } else {
#pragma unroll
            for (int i = 0; i < PARTS; i++)
                s += red[i * QT + q];
        }
        for (int j = part; j < tokens; j += PARTS)
            St[(size_t)j * LDQ + q] /= s;
    }

    // ---- phase 4: O = P V -------------------------------------------------------
    // px -> 4 head dims, py -> OQ queries
    const int px = tid & 15, py = tid >> 4;
    float o[OQ][4];
#pragma unroll
    for (int i = 0; i < OQ; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            o[i][j] = 0.f;
    for (int kb = 0; kb < nkb; kb++) {
        __syncthreads(); // softmax writes visible / previous V block consumed
        for (int f = tid; f < KB * 16; f += 256) {
            const int r = f >> 4, d4 = (f & 15) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kb * KB + r < tokens)
                v = Elem<T>::load4(base + 2 * embed + (size_t)(kb * KB + r) * ld + d4);
            *reinterpret_cast<float4 *>(&KV[r * LDV + d4]) = v;
        }
        __syncthreads();
        const int jmax = min(KB, tokens - kb * KB);
        for (int j = 0; j < jmax; j++) {
            float av[OQ];
            if (OQ == 4) {
                const float4 a = *reinterpret_cast<const float4 *>(&St[(size_t)(kb * KB + j) * LDQ + py * 4]);
                av[0] = a.x;
                av[OQ > 1 ? 1 : 0] = a.y;
                av[OQ > 2 ? 2 : 0] = a.z;
                av[OQ > 3 ? 3 : 0] = a.w;
            } else {
#pragma unroll
                for (int i = 0; i < OQ; i++)
                    av[i] = St[(size_t)(kb * KB + j) * LDQ + py * OQ + i];
            }
            const float4 b = *reinterpret_cast<const float4 *>(&KV[j * LDV + px * 4]);
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < OQ; i++)
#pragma unroll
                for (int d = 0; d < 4; d++)
                    o[i][d] = fmaf(av[i], bv[d], o[i][d]);
        }
    }

    // ---- phase 5 ------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < OQ; i++) {
        const int q = q0 + py * OQ + i;
        if (q < tokens) {
            const float4 v = make_float4(o[i][0], o[i][1], o[i][2], o[i][3]);
            if (kSplit)
                split3_store4(reinterpret_cast<__nv_bfloat16 *>(out_) + ((size_t)img * tokens + q) * 3 * embed, embed,
                              head * kHeadDim + px * 4, v);
            else
                Elem<T>::store4(reinterpret_cast<T *>(out_) + ((size_t)img * tokens + q) * embed + head * kHeadDim + px * 4, v);
        }
    }
}

// ---------------------------------------------------------------------------
// Small-batch variant (batch-1 latency, FP32 storage, tokens <= 256).  The kernel above loads K and V in
// dependent load -> transposing-store loops, block after block; with a handful of images that chain of L2 round
// trips is most of the kernel (24 us at batch 1 for ~2 us of arithmetic).  Here one CTA owns 18 queries of one
// (image, head) -- 11 x 12 = 132 CTAs at batch 1: one wave, no SM with two (16-query tiles are 156 CTAs, and the
// eight SMs that get a second one set the kernel's time) -- and requests its operands with cp.async: Q + K rows in
// its first instructions, the V rows into the same buffer the moment the scores are done (they land under the
// softmax).  Everything stays row-major; what the profile showed to matter is shared-memory wavefronts
// (ncu: 11.5 k per CTA, half of the active cycles, in the first version), so both products are laid out to read every
// operand element once per CTA, 16 distinct bytes per lane, and to broadcast the other operand:
//   S    thread (tx, ty) = (tid / 64, tid % 64): queries tx + 4i, keys ty + 64j -- a warp reads 32 different K rows
//        (pitch 68 floats: conflict-free) and one broadcast Q row per instruction
//   softmax  a warp per query row (R/ViT_seq.c:216-234: max, expf, sum, divide)
//   P V  thread (px, g) = (tid % 16, tid / 16): head dims 4px..4px+3 of ALL 18 queries over the keys g, g + 16, ..;
//        the 16 key groups are summed by one shuffle + an 8-way pass through shared memory
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gmem_src, bool valid)
{
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    const int bytes = valid ? 16 : 0; // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(bytes) : "memory");
}

constexpr int kSmallQT = 18;            // queries per CTA
constexpr int kSmallLdr = kHeadDim + 4; // Q / K / V row pitch (floats)

template <int QT, bool kSplit>
__global__ void __launch_bounds__(256) attention_simt_small_kernel(const float *__restrict__ qkv, void *__restrict__ out_, int tokens,
                                                                   int embed)
{
    constexpr int NI = (QT + 3) / 4; // queries per thread in the score phase
    constexpr int QPAD = NI * 4;     // staged query rows (the pad rows are zero)
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float smem[];
    const int tok4 = (tokens + 3) & ~3;
    float *Qs = smem;                          // [QPAD][68]
    float *KV = Qs + QPAD * kSmallLdr;         // [tok4][68]: K, then V
    float *Ss = KV + (size_t)tok4 * kSmallLdr; // [QT][tok4]
    float *red = Ss + (size_t)QT * tok4;       // [8 warps][QT][64]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = blockIdx.x * QT, head = blockIdx.y, img = blockIdx.z;
    const size_t ld = 3 * (size_t)embed;
    const float *base = qkv + (size_t)img * tokens * ld + head * kHeadDim;

    for (int f = tid; f < QPAD * 16; f += 256) {
        const int r = f >> 4, c = (f & 15) * 4;
        const bool ok = r < QT && q0 + r < tokens;
        cp_async16_zfill(Qs + r * kSmallLdr + c, base + (size_t)(ok ? q0 + r : 0) * ld + c, ok);
    }
    auto load_rows = [&](const float *src) { // K or V rows of this head; rows past `tokens` are zero-filled
        for (int f = tid; f < tok4 * 16; f += 256) {
            const int r = f >> 4, c = (f & 15) * 4;
            const bool ok = r < tokens;
            cp_async16_zfill(KV + r * kSmallLdr + c, src + (size_t)(ok ? r : 0) * ld + c, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    load_rows(base + embed);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- S = Q K^T / 8 (scaled after the dot product, R/ViT_seq.c:211) ----
    {
        const int tx = tid >> 6, ty = tid & 63;
        int kr[4]; // key rows read (clamped into the staged range; results past `tokens` are not stored)
#pragma unroll
        for (int j = 0; j < 4; j++)
            kr[j] = min(ty + 64 * j, tok4 - 1);
        float acc[NI][4];
#pragma unroll
        for (int i = 0; i < NI; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
                acc[i][j] = 0.f;
#pragma unroll 4
        for (int d = 0; d < kHeadDim; d += 4) {
            float4 kv[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                kv[j] = *reinterpret_cast<const float4 *>(KV + kr[j] * kSmallLdr + d);
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const float4 qv = *reinterpret_cast<const float4 *>(Qs + (tx + 4 * i) * kSmallLdr + d);
#pragma unroll
                for (int j = 0; j < 4; j++)
                    acc[i][j] = fmaf(qv.w, kv[j].w, fmaf(qv.z, kv[j].z, fmaf(qv.y, kv[j].y, fmaf(qv.x, kv[j].x, acc[i][j]))));
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int k = ty + 64 * j;
            if (k < tokens) {
#pragma unroll
                for (int i = 0; i < NI; i++)
                    if (tx + 4 * i < QT)
                        Ss[(tx + 4 * i) * tok4 + k] = acc[i][j] * 0.125f;
            }
        }
    }
    __syncthreads(); // every read of K is done: its buffer takes V, requested now and landing under the softmax
    load_rows(base + 2 * embed);

    // ---- row softmax, one warp per query ----
    for (int q = warp; q < QT; q += 8) {
        float *row = Ss + q * tok4;
        float m = -INFINITY;
        for (int k = lane; k < tokens; k += 32)
            m = fmaxf(m, row[k]);
        m = warp_max(m);
        float s = 0.f;
        for (int k = lane; k < tokens; k += 32) {
            const float e = expf(row[k] - m);
            row[k] = e;
            s += e;
        }
        s = warp_sum(s);
        for (int k = lane; k < tokens; k += 32)
            row[k] = row[k] / s;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- O = P V: this thread's key group against all QT queries ----
    {
        const int px = tid & 15, g = tid >> 4;
        float o[QT][4];
#pragma unroll
        for (int i = 0; i < QT; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
                o[i][j] = 0.f;
        for (int k = g; k < tokens; k += 16) {
            const float4 v = *reinterpret_cast<const float4 *>(KV + k * kSmallLdr + px * 4);
#pragma unroll
            for (int i = 0; i < QT; i++) {
                const float p = Ss[i * tok4 + k];
                o[i][0] = fmaf(p, v.x, o[i][0]);
                o[i][1] = fmaf(p, v.y, o[i][1]);
                o[i][2] = fmaf(p, v.z, o[i][2]);
                o[i][3] = fmaf(p, v.w, o[i][3]);
            }
        }
        // the two key groups of a warp meet by shuffle, the eight warps through shared memory
#pragma unroll
        for (int i = 0; i < QT; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
                o[i][j] += __shfl_xor_sync(0xffffffffu, o[i][j], 16);
        if (lane < 16) {
#pragma unroll
            for (int i = 0; i < QT; i++)
                *reinterpret_cast<float4 *>(red + ((size_t)warp * QT + i) * kHeadDim + px * 4) = make_float4(o[i][0], o[i][1], o[i][2], o[i][3]);
        }
    }
    __syncthreads();
    for (int f = tid; f < QT * 16; f += 256) {
        const int i = f >> 4, c = (f & 15) * 4, q = q0 + i;
        if (q >= tokens)
            continue;
        float4 r = *reinterpret_cast<const float4 *>(red + (size_t)i * kHeadDim + c);
#pragma unroll
        for (int w = 1; w < 8; w++) {
            const float4 t = *reinterpret_cast<const float4 *>(red + ((size_t)w * QT + i) * kHeadDim + c);
            r.x += t.x;
            r.y += t.y;
            r.z += t.z;
            r.w += t.w;
        }
        if (kSplit)
            split3_store4(reinterpret_cast<__nv_bfloat16 *>(out_) + ((size_t)img * tokens + q) * 3 * embed, embed, head * kHeadDim + c, r);
        else
            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(out_) + ((size_t)img * tokens + q) * embed + head * kHeadDim + c) = r;
    }
}

template <int QT>
size_t attention_simt_small_smem(int tokens)
{
    const size_t tok4 = (size_t)(tokens + 3) & ~(size_t)3;
    return sizeof(float) * ((size_t)((QT + 3) / 4 * 4) * kSmallLdr + tok4 * kSmallLdr + (size_t)QT * tok4 + (size_t)8 * QT * kHeadDim);
}

template <int QT, bool kSplit>
int launch_attention_simt_small(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t st)
{
    const size_t smem = attention_simt_small_smem<QT>(tokens);
    auto k = attention_simt_small_kernel<QT, kSplit>;
    VITCU_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((tokens + QT - 1) / QT, heads, batch);
    VITCU_TRY(launch_kernel(k, grid, 256, smem, st, reinterpret_cast<const float *>(qkv), out, tokens, heads * kHeadDim));
    return 0;
}

template <int QT>
size_t attention_simt_smem(int tokens)
{
    constexpr int KB = kKeysPerBlock(QT), LDQ = QT + 4, LDK = KB + 4, LDV = kHeadDim + 4;
    constexpr int KVF = kHeadDim * LDK > KB * LDV ? kHeadDim * LDK : KB * LDV;
    const int nkb = (tokens + KB - 1) / KB;
    return sizeof(float) * ((size_t)kHeadDim * LDQ + KVF + (size_t)nkb * KB * LDQ + 256);
}

template <typename T, int QT, bool kSplit = false>
int launch_attention_simt(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t st)
{
    const size_t smem = attention_simt_smem<QT>(tokens);
    VITCU_REQUIRE(smem <= 227 * 1024, "token count too large for the shared-memory score tile");
    auto k = attention_simt_kernel<T, QT, kSplit>;
    VITCU_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((tokens + QT - 1) / QT, heads, batch);
    VITCU_TRY(launch_kernel(k, grid, 256, smem, st, reinterpret_cast<const T *>(qkv), out, tokens,
                            heads * kHeadDim));
    return 0;
}

} // namespace

namespace vitcu {
int attention_bf16_tc(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t st);       // attention_tc.cu
int attention_bf16_flash_tc(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t st); // attention_flash_tc.cu
int attention_bf16_duo_tc(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t st);   // attention_duo_tc.cu
int attention_bf16_flash_duo_tc(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t st); // attention_flash_duo_tc.cu
int device_sm_count(); // gemm_tc.cu
}

extern "C" int vitcu_attention_ex(const void *qkv, void *out, int batch, int tokens, int heads, int is_bf16, vitcu_stream s)
{
    VITCU_REQUIRE(qkv && out && batch > 0 && tokens > 0, "bad argument");
    VITCU_REQUIRE(heads > 0 && heads <= 32, "head count must be 1..32 (head dimension is 64)");
    // BF16 storage: tensor-core kernels -- all keys in one TMEM score buffer when they fit (<= 224),
    // key-blocked online softmax otherwise.  (VITCU_ATTN_SIMT=1 forces the CUDA-core kernel and
    // VITCU_ATTN_FLASH=1 the key-blocked kernel, for A/B measurements.)
    static const bool force_simt = getenv("VITCU_ATTN_SIMT") != nullptr;
    static const bool force_flash = getenv("VITCU_ATTN_FLASH") != nullptr;
    // single-block kernels (all keys in one TMEM score buffer): "duo" = two co-resident CTAs per SM, each running
    // its units serially (default); VITCU_ATTN_KERNEL=solo selects the one-CTA software-pipelined kernel
    const char *which = getenv("VITCU_ATTN_KERNEL"); // read per call: the tests exercise both kernels in one process
    const bool solo = which && !strcmp(which, "solo");
    if (is_bf16 == 1 && !force_simt) {
        if (tokens <= 208 && !force_flash)
            return solo ? attention_bf16_tc(qkv, out, batch, tokens, heads, as_stream(s))
                        : attention_bf16_duo_tc(qkv, out, batch, tokens, heads, as_stream(s));
        if (tokens <= 224 && !force_flash)
            return attention_bf16_tc(qkv, out, batch, tokens, heads, as_stream(s));
        return solo ? attention_bf16_flash_tc(qkv, out, batch, tokens, heads, as_stream(s))
                    : attention_bf16_flash_duo_tc(qkv, out, batch, tokens, heads, as_stream(s));
    }
    // 16-query tiles when 64-query tiles would leave most of the 148 SMs idle (small batches)
    // ... or when the 64-query score tile does not fit shared memory (more than ~620 tokens: 448x448 images)
    const bool small = ((long)batch * heads * ((tokens + 63) / 64) < 148 || attention_simt_smem<64>(tokens) > 227 * 1024) &&
                       attention_simt_smem<16>(tokens) <= 227 * 1024;
    int rc;
    // fp32 storage, a few images, all keys in one staged block: the cp.async kernel (18-query tiles);
    // VITCU_ATTN_SMALL=0 keeps the block-by-block kernel (A/B)
    static const bool no_small = getenv("VITCU_ATTN_SMALL") && !strcmp(getenv("VITCU_ATTN_SMALL"), "0");
    if (small && !no_small && is_bf16 != 1 && tokens <= 256) {
        rc = is_bf16 == 2 ? launch_attention_simt_small<kSmallQT, true>(qkv, out, batch, tokens, heads, as_stream(s))
                          : launch_attention_simt_small<kSmallQT, false>(qkv, out, batch, tokens, heads, as_stream(s));
        if (rc)
            return rc;
        VITCU_LAUNCHED_KIND(LK_ATTN_SIMT);
        return 0;
    }
    if (is_bf16 == 2)
        rc = small ? launch_attention_simt<float, 16, true>(qkv, out, batch, tokens, heads, as_stream(s))
                   : launch_attention_simt<float, 64, true>(qkv, out, batch, tokens, heads, as_stream(s));
    else if (is_bf16)
        rc = small ? launch_attention_simt<__nv_bfloat16, 16>(qkv, out, batch, tokens, heads, as_stream(s))
                   : launch_attention_simt<__nv_bfloat16, 64>(qkv, out, batch, tokens, heads, as_stream(s));
    else
        rc = small ? launch_attention_simt<float, 16>(qkv, out, batch, tokens, heads, as_stream(s))
                   : launch_attention_simt<float, 64>(qkv, out, batch, tokens, heads, as_stream(s));
    if (rc)
        return rc;
    VITCU_LAUNCHED_KIND(LK_ATTN_SIMT);
    return 0;
}

extern "C" int vitcu_attention(const void *qkv, void *out, int batch, int tokens, int is_bf16, vitcu_stream s)
{
    return vitcu_attention_ex(qkv, out, batch, tokens, kHeads, is_bf16, s);
}
