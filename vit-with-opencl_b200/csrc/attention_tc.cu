// csrc/attention_tc.cu -- fused multi-head attention on the 5th-gen tensor cores
// (BF16 operands, FP32 accumulation and softmax), for token counts <= 256.
//
// Replaces QKV_TO_SCOREV (R/multihead.cl:65-137); oracle R/ViT_seq.c:192-262.
// R/ = /root/reference/MulticoreMainProject/.  S = Q K^T, the row softmax and
// O = P V never leave the SM: S and O live in TMEM, P goes through shared
// memory as the A operand of the second MMA.
//
// One persistent CTA per SM walks over (image, head) work items.  Per item the
// whole K and V of the head (KP = tokens rounded up to 16 rows) and up to two
// 128-query tiles are resident in shared memory:
//   warp 0       TMA producer: 3-D tensor map over qkv [B][T][2304]; rows past T
//                are zero-filled by TMA, so no neighbour image leaks in
//   warp 1       MMA issuer:  S_i = Q_i K^T   (M=128, N=KP, K=64, both K-major)
//                             O_i = P_i V     (M=128, N=64, K=KP, A K-major from
//                                              smem, B = V used MN-major as loaded)
//   warp 2       TMEM allocation (512 columns: S_0|O_0 at 0, S_1|O_1 at 256;
//                O_i reuses the columns of S_i once the softmax has consumed it)
//   warps 4-7    softmax + epilogue of query tile 0 (thread = one query row)
//   warps 8-11   softmax + epilogue of query tile 1
// so the two query tiles ping-pong: while one tile's softmax runs on the CUDA
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
constexpr uint32_t P_SLAB = QT * 128;     // one 64-key slab of P: [128 x 64] bf16
constexpr int MAX_KP = 256;

enum Bar { QK_FULL = 0, V_FULL, S_FULL0, S_FULL1, P_FULL0, P_FULL1, O_FULL0, O_FULL1, O_READ0, O_READ1, MMA_DONE, NUM_BARS };

struct AttnParams {
    int batch, tokens, kp;     // kp = tokens rounded up to a multiple of 16
    int items;                 // batch * heads
    __nv_bfloat16 *out;        // [B*T, 768]
};

__global__ void __launch_bounds__(kThreadsAttn, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const AttnParams p, uint32_t *watchdog_flag)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t kv_bytes = static_cast<uint32_t>(p.kp) * 128u;
    const uint32_t slabs = (static_cast<uint32_t>(p.kp) + 63u) / 64u;
    uint8_t *sQ = smem;                          // 2 x [128 x 64]
    uint8_t *sK = sQ + 2 * Q_BYTES;              // [kp x 64]
    uint8_t *sV = sK + kv_bytes;                 // [kp x 64]
    uint8_t *sP = sV + kv_bytes;                 // 2 x slabs x [128 x 64]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sP + 2 * slabs * P_SLAB);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NUM_BARS);
    volatile uint32_t *cta_abort = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (p.tokens + QT - 1) / QT; // 1 or 2

    if (threadIdx.x == 0) {
        mbar_init(&bars[QK_FULL], 1);
        mbar_init(&bars[V_FULL], 1);
        mbar_init(&bars[S_FULL0], 1);
        mbar_init(&bars[S_FULL1], 1);
        mbar_init(&bars[P_FULL0], 4);
        mbar_init(&bars[P_FULL1], 4);
        mbar_init(&bars[O_FULL0], 1);
        mbar_init(&bars[O_FULL1], 1);
        mbar_init(&bars[O_READ0], 4);
        mbar_init(&bars[O_READ1], 4);
        mbar_init(&bars[MMA_DONE], 1);
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
        if (lane == 0) {
            prefetch_tensormap(&tmap_q);
            prefetch_tensormap(&tmap_kv);
            uint32_t it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, it++) {
                const int img = item / kHeads, head = item - img * kHeads;
                if (it > 0 && !mbar_wait(&bars[MMA_DONE], (it - 1) & 1, wd, 1))
                    break;
                mbar_arrive_expect_tx(&bars[QK_FULL], ntiles * Q_BYTES + kv_bytes);
                for (int t = 0; t < ntiles; t++)
                    tma_load_3d(sQ + t * Q_BYTES, &tmap_q, &bars[QK_FULL], head * kHeadDim, t * QT, img);
                tma_load_3d(sK, &tmap_kv, &bars[QK_FULL], kEmbed + head * kHeadDim, 0, img);
                mbar_arrive_expect_tx(&bars[V_FULL], kv_bytes);
                tma_load_3d(sV, &tmap_kv, &bars[V_FULL], 2 * kEmbed + head * kHeadDim, 0, img);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_bf16(QT, p.kp, false, false);
            const uint32_t idesc_o = umma_idesc_bf16(QT, kHeadDim, false, true);
            const uint32_t ksteps = static_cast<uint32_t>(p.kp) / 16u;
            uint32_t it = 0;
            bool ok = true;
            for (int item = blockIdx.x; item < p.items && ok; item += gridDim.x, it++) {
                const uint32_t ph = it & 1;
                if (!(ok = mbar_wait(&bars[QK_FULL], ph, wd, 2)))
                    break;
                const uint64_t k_desc = umma_desc_k_sw128(smem_u32(sK));
                for (int t = 0; t < ntiles && ok; t++) {
                    // S_t overwrites the TMEM columns O_t of the previous item occupied
                    if (it > 0 && !(ok = mbar_wait(&bars[O_READ0 + t], (it - 1) & 1, wd, 3)))
                        break;
                    tcgen05_fence_after();
                    const uint64_t q_desc = umma_desc_k_sw128(smem_u32(sQ + t * Q_BYTES));
#pragma unroll
                    for (int k = 0; k < kHeadDim / 16; k++)
                        umma_bf16_ss(tmem_base + t * 256, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
                    umma_commit(&bars[S_FULL0 + t]);
                }
                if (!ok || !(ok = mbar_wait(&bars[V_FULL], ph, wd, 4)))
                    break;
                for (int t = 0; t < ntiles && ok; t++) {
                    if (!(ok = mbar_wait(&bars[P_FULL0 + t], ph, wd, 5)))
                        break;
                    tcgen05_fence_after();
                    const uint32_t p_base = smem_u32(sP + t * slabs * P_SLAB);
                    const uint32_t v_base = smem_u32(sV);
                    for (uint32_t k = 0; k < ksteps; k++) {
                        // A: slab k/4 of P, +32 B per 16-key step inside the slab's 128-byte rows
                        const uint64_t a_desc = umma_desc_k_sw128(p_base + (k >> 2) * P_SLAB + (k & 3) * 32);
                        // B: 16 key rows of V (= two 8-row swizzle atoms, 2048 B)
                        const uint64_t b_desc = umma_desc_mn_sw128(v_base + k * 2048);
                        umma_bf16_ss(tmem_base + t * 256, a_desc, b_desc, idesc_o, k != 0);
                    }
                    umma_commit(&bars[O_FULL0 + t]);
                }
                if (ok)
                    umma_commit(&bars[MMA_DONE]); // smem of this item may be overwritten
            }
        }
    } else if (warp >= 4) {
        // ===================== softmax + epilogue =====================
        const int tile = (warp - 4) >> 2;   // 0 or 1
        const int quad = warp & 3;          // TMEM lane quadrant of this warp
        const int row = quad * 32 + lane;   // query row inside the tile
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + tile * 256;
        const float sl2 = 0.125f * 1.4426950408889634f; // log2(e) / sqrt(64)
        const int nchunks = (p.kp + 31) / 32;
        uint8_t *prow = sP + tile * slabs * P_SLAB + row * 128;
        uint32_t it = 0;
        if (tile < ntiles) {
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, it++) {
                const int img = item / kHeads, head = item - img * kHeads;
                const uint32_t ph = it & 1;
                bool ok = mbar_wait(&bars[S_FULL0 + tile], ph, wd, 6);
                if (!__all_sync(0xffffffffu, ok))
                    break;
                tcgen05_fence_after();
                // pass 1: row maximum over the valid keys
                float mx = -INFINITY;
                for (int c = 0; c < nchunks; c++) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        if (c * 32 + j < p.tokens)
                            mx = fmaxf(mx, __uint_as_float(v[j]));
                }
                // pass 2: p = exp2((s - max) * log2e/8), row sum, P -> smem (bf16, 128B-swizzled K-major)
                const float mxs = mx * sl2;
                float sum = 0.f;
                for (int c = 0; c < nchunks; c++) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c * 32, v);
                    tmem_ld_wait();
                    float e[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float x = exp2f(fmaf(__uint_as_float(v[j]), sl2, -mxs));
                        e[j] = (c * 32 + j < p.tokens) ? x : 0.f;
                    }
                    // the bf16-rounded values are what the MMA sums, so sum those
                    uint32_t packed[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(e[2 * j], e[2 * j + 1]);
                        packed[j] = *reinterpret_cast<const uint32_t *>(&h);
                        const float2 f = __bfloat1622float2(h);
                        sum += f.x + f.y;
                    }
                    // 32 keys = 4 chunks of 16 B; slab = c/2, chunk index inside the 128-B row = (c&1)*4 + q
                    uint8_t *slab_row = prow + (c >> 1) * P_SLAB;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (c * 32 + q * 8 < p.kp) {
                            const int chunk = ((c & 1) * 4 + q) ^ (row & 7);
                            *reinterpret_cast<uint4 *>(slab_row + chunk * 16) =
                                make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
                        }
                    }
                }
                tcgen05_fence_before();
                fence_proxy_async_smem(); // generic-proxy smem writes -> visible to the MMA (async proxy)
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&bars[P_FULL0 + tile]);

                // epilogue: O / sum -> bf16 -> out[(img*T + q), head*64 ..]
                ok = mbar_wait(&bars[O_FULL0 + tile], ph, wd, 7);
                if (!__all_sync(0xffffffffu, ok))
                    break;
                tcgen05_fence_after();
                const float inv = 1.0f / sum;
                const int q = tile * QT + row;
                __nv_bfloat16 *dst = p.out + (static_cast<size_t>(img) * p.tokens + q) * kEmbed + head * kHeadDim;
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c * 32, v);
                    tmem_ld_wait();
                    if (q < p.tokens) {
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            reinterpret_cast<uint4 *>(dst + c * 32)[j] = make_uint4(
                                pack_bf16x2(__uint_as_float(v[8 * j + 0]) * inv, __uint_as_float(v[8 * j + 1]) * inv),
                                pack_bf16x2(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv),
                                pack_bf16x2(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv),
                                pack_bf16x2(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv));
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&bars[O_READ0 + tile]);
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
    const uint32_t slabs = ((uint32_t)kp + 63u) / 64u;
    const size_t smem = 2 * Q_BYTES + 2 * (size_t)kp * 128 + 2 * slabs * P_SLAB + NUM_BARS * 8 + 16 + 1024;
    VITCU_REQUIRE(smem <= 227 * 1024, "attention tile does not fit shared memory");
    static int configured[64] = {0};
    int dev = 0;
    VITCU_TRY(cudaGetDevice(&dev));
    if (dev < 64 && configured[dev] < (int)smem) {
        VITCU_TRY(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = (int)smem;
    }
    AttnParams p;
    p.batch = batch;
    p.tokens = tokens;
    p.kp = kp;
    p.items = batch * kHeads;
    p.out = reinterpret_cast<__nv_bfloat16 *>(out);
    const int sms = device_sm_count();
    const int grid = p.items < sms ? p.items : sms;
    attention_tc_kernel<<<grid, kThreadsAttn, smem, st>>>(tq, tkv, p, watchdog_flag());
    VITCU_LAUNCHED();
    return 0;
}

} // namespace vitcu
