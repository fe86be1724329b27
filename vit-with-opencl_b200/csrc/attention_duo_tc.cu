// csrc/attention_duo_tc.cu -- fused multi-head attention on the 5th-gen tensor cores for token
// counts <= 224, TWO co-resident CTAs per SM ("duo"): the default BF16 attention kernel.
//
// Replaces QKV_TO_SCOREV (R/multihead.cl:65-137); oracle R/ViT_seq.c:192-262.
// R/ = /root/reference/MulticoreMainProject/.  S = Q K^T, the row softmax and O = P V never leave
// the SM: S, P and O all live in tensor memory.
//
// Why two CTAs per SM.  Per 128-query unit the work is ~420 cycles of tcgen05.mma for S, 26.6 k
// exponentials (MUFU.EX2 runs at 16 per clock per SM on B200: a 1 664-cycle floor), ~420-830 cycles of
// P V and a TMEM read-out.  The one-CTA kernel (attention_tc.cu) software-pipelines those phases over
// two score buffers and four warp groups and still leaves the MUFU idle for half of every period:
// its phases wait on each other through a chain of seven barriers (profiles/r01_v6_attention.md).
// Here every CTA runs the phases of a unit strictly one after the other -- S, row maximum,
// exponentials, P V, read-out -- and the overlap comes from the hardware: each CTA takes 256 of the
// 512 TMEM columns and < 113 KB of shared memory, so two of them share an SM and while one is in its
// MUFU-bound exponential pass the other one issues MMAs, reads its output or loads its next item.
//
// One CTA = 6 warps, persistent over (image, head) items; a UNIT is one 128-query tile of an item.
//   warps 0-3   softmax + epilogue, thread = query row (warp w owns TMEM lanes 32w..32w+31):
//               pass 1 row maximum (scores streamed from TMEM), pass 2 p = exp2((s - max) log2(e)/8)
//               with fp32 row sum, bf16 P written back INTO TMEM over the scores already consumed;
//               then O / sum -> bf16 -> swizzled shared-memory tile -> TMA store (clipped at T)
//   warp 4      TMA producer (3-D tensor map over qkv [B][T][3*embed], rows past T zero-filled) and
//               TMEM allocation.  Shared memory is single-buffered per CTA; the next item's Q and K
//               are requested as soon as the last S of the current item has retired and its V as
//               soon as the last P V has, so the loads run under the remaining phases of the item
//   warp 5      MMA issuer: S = Q K^T (M=128, N=KP, K=64), O = P V (A = P from TMEM, V MN-major,
//               M=128, N=64, K=KP)
// TMEM columns (256 per CTA): scores at 0..KP-1; P (bf16 pairs) overwrites columns 0..KP/2-1; O is
// accumulated at columns 128..191, i.e. inside the score buffer, which is dead once P is complete.
//
// Softmax semantics (R/ViT_seq.c:204-234): scores are scaled by 1/sqrt(64) after the dot product;
// padded keys (j >= T) get p = 0; the division by the row sum is applied to O in the epilogue.
#include "tc_common.cuh"

using namespace vitcu;
using namespace vitcu::tc;

namespace {

constexpr int kThreadsDuo = 192;
constexpr int QT = 128;                   // queries per tile
constexpr uint32_t Q_BYTES = QT * 128;    // [128 x 64] bf16
constexpr uint32_t O_COL = 128;           // output accumulator, 64 columns, inside the dead score buffer
constexpr uint32_t TMEM_COLS = 256;

// barriers: per item {QK_FULL, V_FULL, K_FREE, V_FREE}; per unit {S_FULL, P_FULL, O_FULL, O_FREE}
enum Bar { QK_FULL = 0, V_FULL, K_FREE, V_FREE, S_FULL, P_FULL, O_FULL, O_FREE, NUM_BARS };

struct DuoParams {
    int tokens, kp; // kp = tokens rounded up to a multiple of 16
    int items;      // batch * heads
    int heads, embed;
    int rev;        // 1: walk the items from the last image down (the freshest QKV rows are still in L2)
};

__device__ __forceinline__ float max3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__device__ __forceinline__ int item_of(const DuoParams &p, int raw) { return p.rev ? p.items - 1 - raw : raw; }

// Row maximum over NC 16-column chunks of a score row (thread = row), columns >= valid_cols excluded.
// All loads of the part are issued back to back and waited for once; four independent running maxima.
template <int NC>
__device__ __forceinline__ float row_max_part(uint32_t taddr_s, int valid_cols, float mx)
{
    if (NC == 0)
        return mx;
    uint32_t sc[NC > 0 ? NC : 1][16];
#pragma unroll
    for (int c = 0; c < NC; c++)
        tmem_ld_32x32b_x16(taddr_s + c * 16, sc[c]);
    tmem_ld_wait();
    float m4[4] = {mx, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int c = 0; c < NC; c++) {
        if (c + 1 < NC) { // only the last chunk of a part can hold padded keys
#pragma unroll
            for (int j = 0; j < 16; j += 2)
                m4[(j >> 1) & 3] = max3(m4[(j >> 1) & 3], __uint_as_float(sc[c][j]), __uint_as_float(sc[c][j + 1]));
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (c * 16 + j < valid_cols)
                    m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(sc[c][j]));
        }
    }
    return max3(max3(m4[0], m4[1], m4[2]), m4[3], m4[3]);
}

// NCH = number of 16-column chunks of S (KP / 16)
template <int NCH>
__global__ void __launch_bounds__(kThreadsDuo, 2)
attention_duo_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                        const __grid_constant__ CUtensorMap tmap_out, const DuoParams p, uint32_t *watchdog_flag)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t kv_bytes = static_cast<uint32_t>(p.kp) * 128u;
    uint8_t *sq = smem;                          // Q0 | Q1
    uint8_t *sk = sq + 2 * Q_BYTES;              // K  [kp x 64] bf16, 128B-swizzled rows
    uint8_t *sv = sk + kv_bytes;                 // V
    uint8_t *ostage = sv + kv_bytes;             // [4 warps][32 rows x 128 B], 128B-swizzled, 1 KB aligned
    uint64_t *bars = reinterpret_cast<uint64_t *>(ostage + 4 * 4096);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NUM_BARS);
    volatile uint32_t *cta_abort = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (p.tokens + QT - 1) / QT; // 1 or 2
    const int n_items = blockIdx.x < (unsigned)p.items ? (p.items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int n_units = n_items * ntiles;

    if (threadIdx.x == 0) {
        mbar_init(&bars[QK_FULL], 1);
        mbar_init(&bars[V_FULL], 1);
        mbar_init(&bars[K_FREE], 1);
        mbar_init(&bars[V_FREE], 1);
        mbar_init(&bars[S_FULL], 1);
        mbar_init(&bars[P_FULL], 4);
        mbar_init(&bars[O_FULL], 1);
        mbar_init(&bars[O_FREE], 4);
        *cta_abort = 0;
        fence_barrier_init();
    }
    if (warp == 4)
        tmem_alloc(tmem_slot, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // every CTA holds its TMEM now: the next kernel may start its prologue; our own global-memory
    // traffic (TMA loads, epilogue stores) waits for the previous kernel to complete
    pdl_trigger();
    pdl_wait();
    const Watchdog wd{cta_abort, watchdog_flag};

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            prefetch_tensormap(&tmap_q);
            prefetch_tensormap(&tmap_kv);
        }
        for (int il = 0; il < n_items; il++) {
            const int item = item_of(p, blockIdx.x + il * gridDim.x);
            const int img = item / p.heads, head = item - img * p.heads;
            // Q and K of the previous item are dead once its last S has retired
            if (il > 0 && !mbar_wait_warp(&bars[K_FREE], (il - 1) & 1, wd, 1))
                break;
            if (elect_one()) {
                mbar_arrive_expect_tx(&bars[QK_FULL], ntiles * Q_BYTES + kv_bytes);
                for (int t = 0; t < ntiles; t++)
                    tma_load_3d(sq + t * Q_BYTES, &tmap_q, &bars[QK_FULL], head * kHeadDim, t * QT, img);
                tma_load_3d(sk, &tmap_kv, &bars[QK_FULL], p.embed + head * kHeadDim, 0, img);
            }
            __syncwarp();
            // ... and its V once its last P V has
            if (il > 0 && !mbar_wait_warp(&bars[V_FREE], (il - 1) & 1, wd, 2))
                break;
            if (elect_one()) {
                mbar_arrive_expect_tx(&bars[V_FULL], kv_bytes);
                tma_load_3d(sv, &tmap_kv, &bars[V_FULL], 2 * p.embed + head * kHeadDim, 0, img);
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_s = umma_idesc_bf16(QT, p.kp, false, false);
        const uint32_t idesc_o = umma_idesc_bf16(QT, kHeadDim, false, true);
        const uint32_t sq_a = smem_u32(sq), sk_a = smem_u32(sk), sv_a = smem_u32(sv);
        for (int k = 0; k < n_units; k++) {
            const int il = k / ntiles, t = k - il * ntiles;
            if (t == 0 && !mbar_wait_warp(&bars[QK_FULL], il & 1, wd, 3))
                break;
            // the score buffer also holds the previous unit's output: wait until it has been read out
            if (k > 0 && !mbar_wait_warp(&bars[O_FREE], (k - 1) & 1, wd, 4))
                break;
            tcgen05_fence_after();
            if (elect_one()) {
                const uint64_t q_desc = umma_desc_k_sw128(sq_a + t * Q_BYTES);
                const uint64_t k_desc = umma_desc_k_sw128(sk_a);
#pragma unroll
                for (int kk = 0; kk < kHeadDim / 16; kk++)
                    umma_bf16_ss(tmem_base, q_desc + 2 * kk, k_desc + 2 * kk, idesc_s, kk != 0);
                umma_commit(&bars[S_FULL]);
                if (t == ntiles - 1)
                    umma_commit(&bars[K_FREE]); // Q tiles and K may be overwritten
            }
            __syncwarp();
            if (t == 0 && !mbar_wait_warp(&bars[V_FULL], il & 1, wd, 5))
                break;
            if (!mbar_wait_warp(&bars[P_FULL], k & 1, wd, 6))
                break;
            tcgen05_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < NCH; kk++) // 16 keys per step: 8 packed P columns, 16 V rows of 128 B
                    umma_bf16_ts(tmem_base + O_COL, tmem_base + kk * 8, umma_desc_mn_sw128(sv_a + kk * 2048), idesc_o, kk != 0);
                umma_commit(&bars[O_FULL]);
                if (t == ntiles - 1)
                    umma_commit(&bars[V_FREE]);
            }
            __syncwarp();
        }
    } else {
        // ===================== softmax + epilogue (warps 0-3, thread = query row) =====================
        const int quad = warp;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const float sl2 = 0.125f * 1.4426950408889634f; // log2(e) / sqrt(64)
        uint8_t *tile = ostage + quad * 4096;
        for (int k = 0; k < n_units; k++) {
            const int il = k / ntiles, t = k - il * ntiles;
            // a warp whose 32 query rows all lie past T (second tile of a 197-token item: rows 224..255)
            // only keeps the barrier protocol going; its rows of S are exact zeros (Q zero-filled)
            const bool active = t * QT + quad * 32 < p.tokens;
            if (!mbar_wait_warp(&bars[S_FULL], k & 1, wd, 7))
                break;
            tcgen05_fence_after();
            float inv = 0.f;
            if (active) {
                // ---- pass 1: row maximum over the valid key columns, three rounds of <= 5 chunks ----
                constexpr int R0 = (NCH + 2) / 3, R1 = (NCH - R0 + 1) / 2, R2 = NCH - R0 - R1;
                float mx = row_max_part<R0>(lane_addr, min(p.tokens, R0 * 16), -INFINITY);
                mx = row_max_part<R1>(lane_addr + R0 * 16, max(0, min(p.tokens - R0 * 16, R1 * 16)), mx);
                mx = row_max_part<R2>(lane_addr + (R0 + R1) * 16, max(0, min(p.tokens - (R0 + R1) * 16, R2 * 16)), mx);
                // ---- pass 2: exponentials, chunk c + 1 in flight while chunk c is processed ----
                const f32x2 sl2v = pack2(sl2, sl2), nmx = pack2(-mx * sl2, -mx * sl2);
                f32x2 sum2 = pack2(0.f, 0.f);
                uint32_t sc[2][16];
                tmem_ld_32x32b_x16(lane_addr, sc[0]);
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    tmem_ld_wait();
                    if (c + 1 < NCH)
                        tmem_ld_32x32b_x16(lane_addr + (c + 1) * 16, sc[(c + 1) & 1]);
                    uint32_t packed[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const f32x2 arg = fma2(pack2(__uint_as_float(sc[c & 1][2 * j]), __uint_as_float(sc[c & 1][2 * j + 1])), sl2v, nmx);
                        float a0, a1;
                        unpack2(arg, a0, a1);
                        float e0 = ex2_approx(a0), e1 = ex2_approx(a1);
                        if (c + 1 == NCH) { // only the last chunk can hold padded keys
                            if (c * 16 + 2 * j >= p.tokens)
                                e0 = 0.f;
                            if (c * 16 + 2 * j + 1 >= p.tokens)
                                e1 = 0.f;
                        }
                        sum2 = add2(sum2, pack2(e0, e1));
                        packed[j] = pack_bf16x2(e0, e1);
                    }
                    // P (bf16 pairs) over score columns that are already in registers: 8c + 8 <= 16c + 16
                    tmem_st_32x32b_x8(lane_addr + c * 8, packed);
                }
                float s0, s1;
                unpack2(sum2, s0, s1);
                inv = 1.0f / (s0 + s1);
                tmem_st_wait();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&bars[P_FULL]);
            // ---- epilogue: O / sum -> bf16 -> swizzled tile -> TMA store ----
            if (!mbar_wait_warp(&bars[O_FULL], k & 1, wd, 8))
                break;
            tcgen05_fence_after();
            uint32_t vlo[32], vhi[32];
            if (active) {
                tmem_ld_32x32b_x32(lane_addr + O_COL, vlo);
                tmem_ld_32x32b_x32(lane_addr + O_COL + 32, vhi);
                tmem_ld_wait();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&bars[O_FREE]); // the accumulator (and with it the score buffer) may be overwritten
            if (active) {
                const int item = item_of(p, blockIdx.x + il * gridDim.x);
                const int img = item / p.heads, head = item - img * p.heads;
                if (lane == 0)
                    tma_wait_group_read<0>(); // the previous unit's store has read this tile
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    uint32_t w[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int e = 8 * i + 2 * q; // output columns e, e + 1
                        const uint32_t lo = e < 32 ? vlo[e] : vhi[e - 32], hi = e < 32 ? vlo[e + 1] : vhi[e - 31];
                        w[q] = pack_bf16x2(__uint_as_float(lo) * inv, __uint_as_float(hi) * inv);
                    }
                    *reinterpret_cast<uint4 *>(tile + lane * 128 + ((i ^ (lane & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(&tmap_out, tile, head * kHeadDim, t * QT + quad * 32, img);
                    tma_commit_group();
                }
            }
        }
        if (lane == 0)
            tma_wait_group<0>(); // this warp's output stores have landed before the CTA retires
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 4) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// 3-D map over a [B][T][width] bf16 tensor: box {64 columns, box_rows, 1}, 128-byte swizzle
int make_map3(CUtensorMap *map, const void *base, int batch, int tokens, int width, uint32_t box_rows, CUtensorMapL2promotion l2)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn)
        return set_error(VITCU_E_NODEVICE, __FILE__, __LINE__, "cuTensorMapEncodeTiled is unavailable");
    cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)tokens, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)width * 2, (cuuint64_t)width * 2 * (cuuint64_t)tokens};
    cuuint32_t box[3] = {kHeadDim, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return set_error(VITCU_E_ARG, __FILE__, __LINE__, "cuTensorMapEncodeTiled rejected the attention tensor");
    return 0;
}

} // namespace

namespace vitcu {

int device_sm_count(); // gemm_tc.cu

// qkv [B*T, 3*heads*64] bf16 -> out [B*T, heads*64] bf16; tokens <= 224
int attention_bf16_duo_tc(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t st)
{
    const int embed = heads * kHeadDim;
    const int kp = (tokens + 15) / 16 * 16;
    VITCU_REQUIRE(kp <= 224, "single-block tensor-core attention handles at most 224 tokens");
    VITCU_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0, "buffers must be 16-byte aligned");
    CUtensorMap tq, tkv, tout;
    int rc = make_map3(&tq, qkv, batch, tokens, 3 * embed, QT, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (!rc)
        rc = make_map3(&tkv, qkv, batch, tokens, 3 * embed, (uint32_t)kp, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (!rc) // out viewed as [B][T][embed]: box 32 rows x 64 columns (one head), rows past T are clipped
        rc = make_map3(&tout, out, batch, tokens, embed, 32, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc)
        return rc;
    const size_t smem = 2 * (size_t)Q_BYTES + 2 * (size_t)kp * 128 + 4 * 4096 + NUM_BARS * 8 + 16 + 1024;
    VITCU_REQUIRE(smem <= 113 * 1024, "attention tile does not fit two CTAs per SM");
    DuoParams p;
    p.tokens = tokens;
    p.kp = kp;
    p.items = batch * heads;
    p.heads = heads;
    p.embed = embed;
    static const bool serp = !(getenv("VITCU_SERPENTINE") && atoi(getenv("VITCU_SERPENTINE")) == 0);
    p.rev = serp;
    const int sms = device_sm_count();
    const int grid = p.items < 2 * sms ? p.items : 2 * sms;
    int dev = 0;
    VITCU_TRY(cudaGetDevice(&dev));
    const int nch = kp / 16;
#define VITCU_DUO_CASE(N)                                                                                               \
    case N: {                                                                                                           \
        static int configured[64] = {0};                                                                                \
        if (dev < 64 && configured[dev] < (int)smem) {                                                                  \
            VITCU_TRY(cudaFuncSetAttribute(attention_duo_tc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                           (int)smem));                                                                 \
            VITCU_TRY(cudaFuncSetAttribute(attention_duo_tc_kernel<N>, cudaFuncAttributePreferredSharedMemoryCarveout, \
                                           (int)cudaSharedmemCarveoutMaxShared));                                       \
            configured[dev] = (int)smem;                                                                                \
        }                                                                                                               \
        VITCU_TRY(launch_kernel(attention_duo_tc_kernel<N>, grid, kThreadsDuo, smem, st, tq, tkv, tout, p, watchdog_flag())); \
        break;                                                                                                          \
    }
    switch (nch) {
        VITCU_DUO_CASE(1)
        VITCU_DUO_CASE(2)
        VITCU_DUO_CASE(3)
        VITCU_DUO_CASE(4)
        VITCU_DUO_CASE(5)
        VITCU_DUO_CASE(6)
        VITCU_DUO_CASE(7)
        VITCU_DUO_CASE(8)
        VITCU_DUO_CASE(9)
        VITCU_DUO_CASE(10)
        VITCU_DUO_CASE(11)
        VITCU_DUO_CASE(12)
        VITCU_DUO_CASE(13)
        VITCU_DUO_CASE(14)
    default:
        return set_error(VITCU_E_ARG, __FILE__, __LINE__, "unsupported key count");
    }
#undef VITCU_DUO_CASE
    VITCU_LAUNCHED_KIND(LK_ATTN_DUO);
    return 0;
}

} // namespace vitcu
