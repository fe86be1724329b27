// csrc/attention_tc.cu -- fused multi-head attention on the 5th-gen tensor cores
// (BF16 operands, FP32 accumulation and softmax), for token counts <= 256.
//
// Replaces QKV_TO_SCOREV (R/multihead.cl:65-137); oracle R/ViT_seq.c:192-262.
// R/ = /root/reference/MulticoreMainProject/.  S = Q K^T, the row softmax and
// O = P V never leave the SM: S, P and O all live in tensor memory.
//
// One persistent CTA per SM walks over (image, head) work items.  Per item the
// whole K and V of the head (KP = tokens rounded up to 16 rows) and up to two
// 128-query tiles sit in shared memory; shared memory is double-buffered so the
// TMA loads of item i+1 run under the math of item i.
//   warp 0       TMA producer: 3-D tensor map over qkv [B][T][2304]; rows past T
//                are zero-filled by TMA, so no neighbour image leaks in
//   warp 1       MMA issuer:  S_t = Q_t K^T  (M=128, N=KP, K=64, operands in smem)
//                             O_t = P_t V    (M=128, N=64, K=KP, A = P read from
//                                             TMEM, B = V used MN-major as loaded)
//   warp 2       TMEM allocation: one 256-column slot per query tile,
//                S in columns [0,KP), P (bf16 pairs) overwrites columns [0,KP/2)
//                of S as the softmax consumes it, O in columns [128,192)
//   warps 4-7    softmax + epilogue of query tile 0 (thread = one query row =
//                one TMEM lane, so S -> P in place needs no cross-thread sync)
//   warps 8-11   softmax + epilogue of query tile 1
// The two query tiles ping-pong: while one tile's softmax runs on the CUDA
// cores the other tile's MMAs run on the tensor cores.
//
// Softmax (R/ViT_seq.c:204-234): scores are scaled by 1/sqrt(64) after the dot
// product; p = exp(s/8 - max/8) is evaluated as exp2((s - max) * log2(e)/8);
// padded keys (j >= T) get p = 0; the division by the row sum is applied to O
// in the epilogue (same value, 64 instead of T divisions per row).
#include "tc_common.cuh"

using namespace vitcu;
using namespace vitcu::tc;

namespace {

constexpr int kThreadsAttn = 384;
constexpr int QT = 128;                   // queries per tile
constexpr uint32_t Q_BYTES = QT * 128;    // [128 x 64] bf16
constexpr int MAX_KP = 256;
constexpr uint32_t O_COL = 128;           // O accumulator columns inside a tile's TMEM slot

// barriers: per smem stage {QK_FULL, V_FULL, SMEM_FREE}; per query tile {S_FULL, P_FULL, O_FULL, O_READ}
enum Bar { QK_FULL = 0, V_FULL = 2, SMEM_FREE = 4, S_FULL = 6, P_FULL = 8, O_FULL = 10, O_READ = 12, NUM_BARS = 14 };

struct AttnParams {
    int batch, tokens, kp;     // kp = tokens rounded up to a multiple of 16
    int items;                 // batch * heads
    __nv_bfloat16 *out;        // [B*T, 768]
};

__device__ __forceinline__ float max3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// NCH = number of 32-column chunks of S (ceil(KP/32)); a template parameter so the register
// double-buffering below indexes statically (a run-time chunk loop made ptxas select between the two
// buffers with 32 SELs per chunk and pass, and the ALU pipe, not the MUFU, became the limiter:
// profiles/r01_v6_attention.md)
template <int NCH>
__global__ void __launch_bounds__(kThreadsAttn, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const AttnParams p, uint32_t *watchdog_flag)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t kv_bytes = static_cast<uint32_t>(p.kp) * 128u;
    const uint32_t stage_bytes = 2 * Q_BYTES + 2 * kv_bytes; // Q0 | Q1 | K | V
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * stage_bytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NUM_BARS);
    volatile uint32_t *cta_abort = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (p.tokens + QT - 1) / QT; // 1 or 2

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(&bars[QK_FULL + s], 1);
            mbar_init(&bars[V_FULL + s], 1);
            mbar_init(&bars[SMEM_FREE + s], 1);
            mbar_init(&bars[S_FULL + s], 1);
            mbar_init(&bars[P_FULL + s], 4);
            mbar_init(&bars[O_FULL + s], 1);
            mbar_init(&bars[O_READ + s], 4);
        }
        *cta_abort = 0;
        fence_barrier_init();
    }
    if (warp == 2)
        tmem_alloc(tmem_slot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const Watchdog wd{cta_abort, watchdog_flag};

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            prefetch_tensormap(&tmap_q);
            prefetch_tensormap(&tmap_kv);
        }
        uint32_t it = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, it++) {
            const int img = item / kHeads, head = item - img * kHeads;
            const uint32_t s = it & 1;
            // stage s was last used by item it-2: wait until its MMAs have retired
            if (it >= 2 && !mbar_wait_warp(&bars[SMEM_FREE + s], ((it >> 1) - 1) & 1, wd, 1))
                break;
            if (elect_one()) {
                uint8_t *sq = smem + s * stage_bytes;
                uint8_t *sk = sq + 2 * Q_BYTES;
                mbar_arrive_expect_tx(&bars[QK_FULL + s], ntiles * Q_BYTES + kv_bytes);
                for (int t = 0; t < ntiles; t++)
                    tma_load_3d(sq + t * Q_BYTES, &tmap_q, &bars[QK_FULL + s], head * kHeadDim, t * QT, img);
                tma_load_3d(sk, &tmap_kv, &bars[QK_FULL + s], kEmbed + head * kHeadDim, 0, img);
                mbar_arrive_expect_tx(&bars[V_FULL + s], kv_bytes);
                tma_load_3d(sk + kv_bytes, &tmap_kv, &bars[V_FULL + s], 2 * kEmbed + head * kHeadDim, 0, img);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_s = umma_idesc_bf16(QT, p.kp, false, false);
        const uint32_t idesc_o = umma_idesc_bf16(QT, kHeadDim, false, true);
        const uint32_t ksteps = static_cast<uint32_t>(p.kp) / 16u;
        uint32_t it = 0;
        bool ok = true;
        for (int item = blockIdx.x; item < p.items && ok; item += gridDim.x, it++) {
            const uint32_t s = it & 1, sph = (it >> 1) & 1, ph = it & 1;
            const uint32_t sq = smem_u32(smem + s * stage_bytes);
            const uint32_t sk = sq + 2 * Q_BYTES, sv = sk + kv_bytes;
            if (!(ok = mbar_wait_warp(&bars[QK_FULL + s], sph, wd, 2)))
                break;
            for (int t = 0; t < ntiles && ok; t++) {
                // S_t overwrites the TMEM slot whose O_t the previous item's epilogue is still reading
                if (it > 0 && !(ok = mbar_wait_warp(&bars[O_READ + t], (it - 1) & 1, wd, 3)))
                    break;
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint64_t q_desc = umma_desc_k_sw128(sq + t * Q_BYTES);
                    const uint64_t k_desc = umma_desc_k_sw128(sk);
#pragma unroll
                    for (int k = 0; k < kHeadDim / 16; k++)
                        umma_bf16_ss(tmem_base + t * 256, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
                    umma_commit(&bars[S_FULL + t]);
                }
                __syncwarp();
            }
            if (!ok || !(ok = mbar_wait_warp(&bars[V_FULL + s], sph, wd, 4)))
                break;
            for (int t = 0; t < ntiles && ok; t++) {
                if (!(ok = mbar_wait_warp(&bars[P_FULL + t], ph, wd, 5)))
                    break;
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t slot = tmem_base + t * 256;
                    for (uint32_t k = 0; k < ksteps; k++) {
                        // A: 16 keys = 8 packed columns of P in TMEM; B: 16 key rows of V (two 8-row swizzle atoms)
                        const uint64_t b_desc = umma_desc_mn_sw128(sv + k * 2048);
                        umma_bf16_ts(slot + O_COL, slot + k * 8, b_desc, idesc_o, k != 0);
                    }
                    umma_commit(&bars[O_FULL + t]);
                    if (t == ntiles - 1)
                        umma_commit(&bars[SMEM_FREE + s]); // Q/K/V of this stage may be overwritten
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===================== softmax + epilogue =====================
        const int tile = (warp - 4) >> 2;   // 0 or 1
        const int quad = warp & 3;          // TMEM lane quadrant of this warp
        const int row = quad * 32 + lane;   // query row inside the tile
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + tile * 256;
        const float sl2 = 0.125f * 1.4426950408889634f; // log2(e) / sqrt(64)
        uint32_t it = 0;
        if (tile < ntiles) {
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, it++) {
                const int img = item / kHeads, head = item - img * kHeads;
                const uint32_t ph = it & 1;
                if (!mbar_wait_warp(&bars[S_FULL + tile], ph, wd, 6))
                    break;
                tcgen05_fence_after();
                uint32_t buf[2][32];
                // pass 1: row maximum over the valid keys; tcgen05.ld double-buffered in registers.
                // Only the last chunk can hold padded keys (tokens > 32 * (NCH - 1) by construction).
                float mx = -INFINITY;
                tmem_ld_32x32b_x32(taddr, buf[0]);
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    tmem_ld_wait();
                    // next chunk, or chunk 0 again for pass 2
                    tmem_ld_32x32b_x32(taddr + (c + 1 < NCH ? (c + 1) * 32 : 0), buf[(c + 1) & 1]);
                    const uint32_t(&cur)[32] = buf[c & 1];
                    if (c + 1 < NCH) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2)
                            mx = max3(mx, __uint_as_float(cur[j]), __uint_as_float(cur[j + 1]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (c * 32 + j < p.tokens)
                                mx = fmaxf(mx, __uint_as_float(cur[j]));
                    }
                }
                // pass 2: p = exp2((s - max) * log2e/8), row sum, P (bf16 pairs) back into TMEM over S.
                // The row sum is taken over the fp32 values (3 instructions per element in total:
                // 1/2 FFMA2, 1 MUFU.EX2, 1/2 F2FP pack, 1/2 FADD2, + the max above).
                const f32x2 sl2v = pack2(sl2, sl2), nmx = pack2(-mx * sl2, -mx * sl2);
                f32x2 sum2 = pack2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    tmem_ld_wait();
                    if (c + 1 < NCH)
                        tmem_ld_32x32b_x32(taddr + (c + 1) * 32, buf[(NCH + c + 1) & 1]);
                    const uint32_t(&cur)[32] = buf[(NCH + c) & 1];
                    uint32_t packed[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        float a0, a1;
                        unpack2(fma2(pack2(__uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1])), sl2v, nmx), a0, a1);
                        float e0 = ex2_approx(a0), e1 = ex2_approx(a1);
                        if (c + 1 == NCH) { // padded keys contribute nothing
                            if (c * 32 + 2 * j >= p.tokens)
                                e0 = 0.f;
                            if (c * 32 + 2 * j + 1 >= p.tokens)
                                e1 = 0.f;
                        }
                        sum2 = add2(sum2, pack2(e0, e1));
                        packed[j] = pack_bf16x2(e0, e1);
                    }
                    // P columns [16c, 16c+16) overwrite S columns that are already consumed
                    // (chunk c/2 <= c) or not yet prefetched (chunk c+1 starts at column 32c+32)
                    tmem_st_32x32b_x16(taddr + c * 16, packed);
                }
                float sum, sum_hi;
                unpack2(sum2, sum, sum_hi);
                sum += sum_hi;
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&bars[P_FULL + tile]);

                // epilogue: O / sum -> bf16 -> out[(img*T + q), head*64 ..]
                if (!mbar_wait_warp(&bars[O_FULL + tile], ph, wd, 7))
                    break;
                tcgen05_fence_after();
                const float inv = 1.0f / sum;
                const int q = tile * QT + row;
                __nv_bfloat16 *dst = p.out + (static_cast<size_t>(img) * p.tokens + q) * kEmbed + head * kHeadDim;
                tmem_ld_32x32b_x32(taddr + O_COL, buf[0]);
                tmem_ld_32x32b_x32(taddr + O_COL + 32, buf[1]);
                tmem_ld_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&bars[O_READ + tile]); // the TMEM slot is free for the next item's S
                if (q < p.tokens) {
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const uint32_t(&v)[32] = buf[c];
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            reinterpret_cast<uint4 *>(dst + c * 32)[j] = make_uint4(
                                pack_bf16x2(__uint_as_float(v[8 * j + 0]) * inv, __uint_as_float(v[8 * j + 1]) * inv),
                                pack_bf16x2(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv),
                                pack_bf16x2(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv),
                                pack_bf16x2(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv));
                    }
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_qkv_map(CUtensorMap *map, const void *qkv, int batch, int tokens, uint32_t box_rows)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return set_error(VITCU_E_NODEVICE, __FILE__, __LINE__, "cuTensorMapEncodeTiled is unavailable");
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    const cuuint64_t ld = 3 * kEmbed;
    cuuint64_t dims[3] = {ld, (cuuint64_t)tokens, (cuuint64_t)batch};
    cuuint64_t strides[2] = {ld * 2, ld * 2 * (cuuint64_t)tokens};
    cuuint32_t box[3] = {kHeadDim, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(qkv), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(VITCU_E_ARG, __FILE__, __LINE__, "cuTensorMapEncodeTiled rejected the qkv tensor");
    return 0;
}

} // namespace

namespace vitcu {

int device_sm_count(); // gemm_tc.cu

// qkv [B*T, 2304] bf16 -> out [B*T, 768] bf16; tokens <= 256
int attention_bf16_tc(const void *qkv, void *out, int batch, int tokens, cudaStream_t st)
{
    const int kp = (tokens + 15) / 16 * 16;
    VITCU_REQUIRE(kp <= MAX_KP, "tensor-core attention handles at most 256 tokens");
    VITCU_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0, "buffers must be 16-byte aligned");
    CUtensorMap tq, tkv;
    int rc = make_qkv_map(&tq, qkv, batch, tokens, QT);
    if (rc)
        return rc;
    rc = make_qkv_map(&tkv, qkv, batch, tokens, (uint32_t)kp);
    if (rc)
        return rc;
    const size_t smem = 2 * (2 * (size_t)Q_BYTES + 2 * (size_t)kp * 128) + NUM_BARS * 8 + 16 + 1024;
    VITCU_REQUIRE(smem <= 227 * 1024, "attention tile does not fit shared memory");
    AttnParams p;
    p.batch = batch;
    p.tokens = tokens;
    p.kp = kp;
    p.items = batch * kHeads;
    p.out = reinterpret_cast<__nv_bfloat16 *>(out);
    const int sms = device_sm_count();
    const int grid = p.items < sms ? p.items : sms;
    int dev = 0;
    VITCU_TRY(cudaGetDevice(&dev));
    const int nch = (kp + 31) / 32;
#define VITCU_ATTN_CASE(N)                                                                                          \
    case N: {                                                                                                       \
        static int configured[64] = {0};                                                                            \
        if (dev < 64 && configured[dev] < (int)smem) {                                                              \
            VITCU_TRY(cudaFuncSetAttribute(attention_tc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                           (int)smem));                                                             \
            configured[dev] = (int)smem;                                                                            \
        }                                                                                                           \
        attention_tc_kernel<N><<<grid, kThreadsAttn, smem, st>>>(tq, tkv, p, watchdog_flag());                     \
        break;                                                                                                      \
    }
    switch (nch) {
        VITCU_ATTN_CASE(1)
        VITCU_ATTN_CASE(2)
        VITCU_ATTN_CASE(3)
        VITCU_ATTN_CASE(4)
        VITCU_ATTN_CASE(5)
        VITCU_ATTN_CASE(6)
        VITCU_ATTN_CASE(7)
        VITCU_ATTN_CASE(8)
    default:
        return set_error(VITCU_E_ARG, __FILE__, __LINE__, "unsupported key count");
    }
#undef VITCU_ATTN_CASE
    VITCU_LAUNCHED();
    return 0;
}

} // namespace vitcu
