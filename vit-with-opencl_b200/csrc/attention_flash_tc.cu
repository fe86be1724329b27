// csrc/attention_flash_tc.cu -- key-blocked ("flash") multi-head attention on the tensor cores for
// ANY token count (used for T > 256, e.g. the 384x384 / 577-token configuration that the
// reference's OpenCL kernel cannot run at all: R/multihead.cl:81-83 caps keys at 256).
// BF16 operands, FP32 accumulation and softmax.  Oracle: R/ViT_seq.c:192-262.
// R/ = /root/reference/MulticoreMainProject/.
//
// Work item = (image, head, pair of 128-query tiles).  Keys/values stream through a 4-stage TMA
// ring in blocks of 128; per block and query tile:
//     S = Q K_j^T            (tcgen05, M=128, N=128, K=64)          -> TMEM
//     online softmax         (one row per thread): m' = max(m, rowmax S), a = exp(m - m'),
//                            P = exp(S - m') as bf16 pairs back into TMEM over S, l = l a + sum P
//     O_j = P V_j            (tcgen05, A = P from TMEM, B = V_j MN-major, M=128, N=64, K=128)
//     o = o a + O_j          (64 fp32 accumulators per row in registers)
// and finally out = o / l.  Keys past T are zero-filled by TMA and masked to p = 0; scores are
// scaled by 1/sqrt(64) after the dot product (R/ViT_seq.c:211), folded into the exp2 argument.
//
//   warp 0       TMA producer (Q tiles double-buffered across items, K/V ring)
//   warp 1       MMA issuer; per tile the order is PV(j) then S(j+1), so the next scores are
//                already being computed while the softmax warps fold O_j into their registers
//   warp 2       TMEM allocation: per tile S/P at columns [0,128), O_j at [128,192)
//   warps 4-7    softmax + accumulation of query tile 0;  warps 8-11 of query tile 1
#include "tc_common.cuh"

using namespace vitcu;
using namespace vitcu::tc;

namespace {

constexpr int kThreadsFlash = 384;
constexpr int QT = 128, KB = 128, NS = 4;
constexpr uint32_t TILE_BYTES = 128 * 128;  // [128 rows x 64] bf16: one Q tile, one K block or one V block
constexpr uint32_t O_COL = 128;

// barrier indices
enum { Q_FULL = 0, Q_EMPTY = 2, KV_FULL = 4, KV_EMPTY = KV_FULL + NS, S_FULL = KV_EMPTY + NS, P_FULL = S_FULL + 2,
       O_FULL = P_FULL + 2, O_READ = O_FULL + 2, NUM_BARS = O_READ + 2 };

struct FlashParams {
    int batch, tokens;
    int qtiles;   // ceil(T / 128)
    int qpairs;   // ceil(qtiles / 2)
    int kblocks;  // ceil(T / 128)
    int items;    // batch * heads * qpairs
    int heads, embed; // embed = heads * 64: Q | K | V column blocks of qkv start at 0, embed, 2 * embed
    __nv_bfloat16 *out;
};

__device__ __forceinline__ float max3f(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__global__ void __launch_bounds__(kThreadsFlash, 1)
attention_flash_tc_kernel(const __grid_constant__ CUtensorMap tmap, const FlashParams p, uint32_t *watchdog_flag)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *sQ = smem;                          // [2 stages][2 tiles]
    uint8_t *sKV = sQ + 4 * TILE_BYTES;          // [NS stages][K | V]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sKV + NS * 2 * TILE_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NUM_BARS);
    volatile uint32_t *cta_abort = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&bars[Q_FULL + i], 1);
            mbar_init(&bars[Q_EMPTY + i], 1);
            mbar_init(&bars[S_FULL + i], 1);
            mbar_init(&bars[P_FULL + i], 4);
            mbar_init(&bars[O_FULL + i], 1);
            mbar_init(&bars[O_READ + i], 4);
        }
        for (int i = 0; i < NS; i++) {
            mbar_init(&bars[KV_FULL + i], 1);
            mbar_init(&bars[KV_EMPTY + i], 1);
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
    // every CTA holds its TMEM now: the next kernel may start its prologue; our own global-memory
    // traffic (TMA loads, epilogue stores) waits for the previous kernel to complete
    pdl_trigger();
    pdl_wait();
    const Watchdog wd{cta_abort, watchdog_flag};

    // item -> (image, head, first query tile, tiles in this item)
    auto decode = [&](int item, int &img, int &head, int &q0, int &nt) {
        const int pair = item % p.qpairs;
        const int bh = item / p.qpairs;
        img = bh / p.heads;
        head = bh - img * p.heads;
        q0 = pair * 2;
        nt = p.qtiles - q0 >= 2 ? 2 : 1;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one())
            prefetch_tensormap(&tmap);
        uint32_t it = 0, kv = 0; // items and K/V blocks issued so far
        bool ok = true;
        for (int item = blockIdx.x; item < p.items && ok; item += gridDim.x, it++) {
            int img, head, q0, nt;
            decode(item, img, head, q0, nt);
            const uint32_t qs = it & 1;
            if (!(ok = mbar_wait_warp(&bars[Q_EMPTY + qs], ((it >> 1) & 1) ^ 1, wd, 1)))
                break;
            if (elect_one()) {
                mbar_arrive_expect_tx(&bars[Q_FULL + qs], nt * TILE_BYTES);
                for (int t = 0; t < nt; t++)
                    tma_load_3d(sQ + (qs * 2 + t) * TILE_BYTES, &tmap, &bars[Q_FULL + qs], head * kHeadDim, (q0 + t) * QT, img);
            }
            __syncwarp();
            for (int j = 0; j < p.kblocks; j++, kv++) {
                const uint32_t s = kv % NS;
                if (!(ok = mbar_wait_warp(&bars[KV_EMPTY + s], ((kv / NS) & 1) ^ 1, wd, 2)))
                    break;
                if (elect_one()) {
                    uint8_t *dst = sKV + s * 2 * TILE_BYTES;
                    mbar_arrive_expect_tx(&bars[KV_FULL + s], 2 * TILE_BYTES);
                    tma_load_3d(dst, &tmap, &bars[KV_FULL + s], p.embed + head * kHeadDim, j * KB, img);
                    tma_load_3d(dst + TILE_BYTES, &tmap, &bars[KV_FULL + s], 2 * p.embed + head * kHeadDim, j * KB, img);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_s = umma_idesc_bf16(QT, KB, false, false);
        const uint32_t idesc_o = umma_idesc_bf16(QT, kHeadDim, false, true);
        uint32_t it = 0, kv = 0;
        uint32_t n_p[2] = {0, 0};  // P blocks consumed per tile (== PV MMAs issued)
        bool ok = true;
        auto issue_s = [&](int t, uint32_t qs, uint32_t stage) {
            if (elect_one()) {
                const uint64_t q_desc = umma_desc_k_sw128(smem_u32(sQ + (qs * 2 + t) * TILE_BYTES));
                const uint64_t k_desc = umma_desc_k_sw128(smem_u32(sKV + stage * 2 * TILE_BYTES));
#pragma unroll
                for (int k = 0; k < kHeadDim / 16; k++)
                    umma_bf16_ss(tmem_base + t * 256, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
                umma_commit(&bars[S_FULL + t]);
            }
            __syncwarp();
        };
        for (int item = blockIdx.x; item < p.items && ok; item += gridDim.x, it++) {
            int img, head, q0, nt;
            decode(item, img, head, q0, nt);
            const uint32_t qs = it & 1;
            if (!(ok = mbar_wait_warp(&bars[Q_FULL + qs], (it >> 1) & 1, wd, 3)))
                break;
            // scores of block 0 for every tile of the item
            if (!(ok = mbar_wait_warp(&bars[KV_FULL + kv % NS], (kv / NS) & 1, wd, 4)))
                break;
            tcgen05_fence_after();
            for (int t = 0; t < nt; t++)
                issue_s(t, qs, kv % NS);
            for (int j = 0; j < p.kblocks && ok; j++, kv++) {
                const uint32_t stage = kv % NS;
                if (j + 1 < p.kblocks) // next block's K (and V) must have landed before S(j+1)
                    if (!(ok = mbar_wait_warp(&bars[KV_FULL + (kv + 1) % NS], ((kv + 1) / NS) & 1, wd, 5)))
                        break;
                for (int t = 0; t < nt && ok; t++) {
                    // P_t(j) written by the softmax warps; O_t of the previous block folded into their registers
                    if (!(ok = mbar_wait_warp(&bars[P_FULL + t], n_p[t] & 1, wd, 6)))
                        break;
                    if (n_p[t] > 0 && !(ok = mbar_wait_warp(&bars[O_READ + t], (n_p[t] - 1) & 1, wd, 7)))
                        break;
                    tcgen05_fence_after();
                    if (elect_one()) {
                        const uint32_t slot = tmem_base + t * 256;
                        const uint32_t sv = smem_u32(sKV + stage * 2 * TILE_BYTES + TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < KB / 16; k++)
                            umma_bf16_ts(slot + O_COL, slot + k * 8, umma_desc_mn_sw128(sv + k * 2048), idesc_o, k != 0);
                        umma_commit(&bars[O_FULL + t]);
                    }
                    __syncwarp();
                    n_p[t]++;
                    if (j + 1 < p.kblocks)
                        issue_s(t, qs, (kv + 1) % NS); // in order after PV_t(j): P_t(j) is dead by then
                }
                if (ok && elect_one())
                    umma_commit(&bars[KV_EMPTY + stage]); // K_j and V_j are no longer read
                __syncwarp();
            }
            if (ok && elect_one())
                umma_commit(&bars[Q_EMPTY + qs]);
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===================== softmax + accumulation =====================
        const int tile = (warp - 4) >> 2;
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + tile * 256;
        const float sl2 = 0.125f * 1.4426950408889634f; // log2(e) / sqrt(64)
        uint32_t n_s = 0; // blocks processed by this warp's tile
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            int img, head, q0, nt;
            decode(item, img, head, q0, nt);
            if (tile >= nt)
                continue; // this query tile does not exist in the item (odd tile count)
            float o[64];
#pragma unroll
            for (int i = 0; i < 64; i++)
                o[i] = 0.f;
            float m = -INFINITY, l = 0.f;
            bool ok = true;
            for (int j = 0; j < p.kblocks; j++, n_s++) {
                const uint32_t ph = n_s & 1;
                if (!(ok = mbar_wait_warp(&bars[S_FULL + tile], ph, wd, 8)))
                    break;
                tcgen05_fence_after();
                const int valid = p.tokens - j * KB; // keys of this block that exist (>= 1)
                uint32_t v[32];
                // pass 1: block maximum
                float bm = -INFINITY;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    tmem_ld_32x32b_x32(taddr + c * 32, v);
                    tmem_ld_wait();
                    if (c * 32 + 32 <= valid) {
#pragma unroll
                        for (int i = 0; i < 32; i += 2)
                            bm = max3f(bm, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i++)
                            if (c * 32 + i < valid)
                                bm = fmaxf(bm, __uint_as_float(v[i]));
                    }
                }
                const float m_new = fmaxf(m, bm);
                const float alpha = ex2_approx((m - m_new) * sl2); // 0 for the first block (m = -inf)
                m = m_new;
                // pass 2: P = exp2((s - m) * log2e/8) -> bf16 pairs into TMEM over S; row sum
                const f32x2 sl2v = pack2(sl2, sl2), nmx = pack2(-m * sl2, -m * sl2);
                f32x2 sum2 = pack2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    tmem_ld_32x32b_x32(taddr + c * 32, v);
                    tmem_ld_wait();
                    uint32_t packed[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        float a0, a1;
                        unpack2(fma2(pack2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sl2v, nmx), a0, a1);
                        float e0 = ex2_approx(a0), e1 = ex2_approx(a1);
                        if (c * 32 + 32 > valid) {
                            if (c * 32 + 2 * i >= valid)
                                e0 = 0.f;
                            if (c * 32 + 2 * i + 1 >= valid)
                                e1 = 0.f;
                        }
                        sum2 = add2(sum2, pack2(e0, e1));
                        packed[i] = pack_bf16x2(e0, e1);
                    }
                    tmem_st_32x32b_x16(taddr + c * 16, packed); // columns [16c,16c+16) of S are consumed already
                }
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&bars[P_FULL + tile]);
                float s0, s1;
                unpack2(sum2, s0, s1);
                l = fmaf(l, alpha, s0 + s1);
                // while the PV MMA runs: scale the running output by alpha
#pragma unroll
                for (int i = 0; i < 64; i++)
                    o[i] *= alpha;
                if (!(ok = mbar_wait_warp(&bars[O_FULL + tile], ph, wd, 9)))
                    break;
                tcgen05_fence_after();
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    tmem_ld_32x32b_x32(taddr + O_COL + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; i++)
                        o[c * 32 + i] += __uint_as_float(v[i]);
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&bars[O_READ + tile]);
            }
            if (!ok)
                break;
            const int q = (q0 + tile) * QT + row;
            if (q < p.tokens) {
                const float inv = 1.0f / l;
                __nv_bfloat16 *dst = p.out + (static_cast<size_t>(img) * p.tokens + q) * p.embed + head * kHeadDim;
#pragma unroll
                for (int i = 0; i < 8; i++)
                    reinterpret_cast<uint4 *>(dst)[i] =
                        make_uint4(pack_bf16x2(o[8 * i + 0] * inv, o[8 * i + 1] * inv), pack_bf16x2(o[8 * i + 2] * inv, o[8 * i + 3] * inv),
                                   pack_bf16x2(o[8 * i + 4] * inv, o[8 * i + 5] * inv), pack_bf16x2(o[8 * i + 6] * inv, o[8 * i + 7] * inv));
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

} // namespace

namespace vitcu {

int device_sm_count(); // gemm_tc.cu

// qkv [B*T, 2304] bf16 -> out [B*T, 768] bf16; any token count
int attention_bf16_flash_tc(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t st)
{
    VITCU_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0, "buffers must be 16-byte aligned");
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return set_error(VITCU_E_NODEVICE, __FILE__, __LINE__, "cuTensorMapEncodeTiled is unavailable");
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    // one 3-D map over qkv [B][T][2304], box = 128 rows x 64 columns (Q tiles, K blocks and V blocks alike)
    CUtensorMap map;
    const cuuint64_t ld = 3 * (cuuint64_t)(heads * kHeadDim);
    cuuint64_t dims[3] = {ld, (cuuint64_t)tokens, (cuuint64_t)batch};
    cuuint64_t strides[2] = {ld * 2, ld * 2 * (cuuint64_t)tokens};
    cuuint32_t box[3] = {kHeadDim, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (fn(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(qkv), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return set_error(VITCU_E_ARG, __FILE__, __LINE__, "cuTensorMapEncodeTiled rejected the qkv tensor");

    FlashParams p;
    p.batch = batch;
    p.tokens = tokens;
    p.qtiles = (tokens + QT - 1) / QT;
    p.qpairs = (p.qtiles + 1) / 2;
    p.kblocks = (tokens + KB - 1) / KB;
    p.items = batch * heads * p.qpairs;
    p.heads = heads;
    p.embed = heads * kHeadDim;
    p.out = reinterpret_cast<__nv_bfloat16 *>(out);
    const size_t smem = (4 + 2 * NS) * (size_t)TILE_BYTES + NUM_BARS * 8 + 16 + 1024;
    static bool configured[64] = {false};
    int dev = 0;
    VITCU_TRY(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        VITCU_TRY(cudaFuncSetAttribute(attention_flash_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = true;
    }
    const int sms = device_sm_count();
    const int grid = p.items < sms ? p.items : sms;
    VITCU_TRY(launch_kernel(attention_flash_tc_kernel, grid, kThreadsFlash, smem, st, map, p, watchdog_flag()));
    VITCU_LAUNCHED_KIND(LK_ATTN_FLASH);
    return 0;
}

} // namespace vitcu
