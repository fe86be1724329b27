// csrc/attention_tc.cu -- fused multi-head attention on the 5th-gen tensor cores
// (BF16 operands, FP32 accumulation and softmax), for token counts <= 256.
//
// Replaces QKV_TO_SCOREV (R/multihead.cl:65-137); oracle R/ViT_seq.c:192-262.
// R/ = /root/reference/MulticoreMainProject/.  S = Q K^T, the row softmax and
// O = P V never leave the SM: S, P and O all live in tensor memory.
//
// One persistent CTA per SM walks over (image, head) items; a UNIT of work is one 128-query
// tile of an item.  Shared memory holds K, V (KP = tokens rounded up to 16 rows) and both query
// tiles of an item, double-buffered so the TMA loads of item i+1 run under the math of item i.
// Tensor memory holds TWO score buffers (units alternate between them) and one output buffer:
//
//   warp 0       TMA producer: 3-D tensor map over qkv [B][T][2304]; rows past T are zero-filled
//   warp 1       MMA issuer, software-pipelined across units:
//                    S(k+1) = Q K^T          (M=128, N=KP, K=64)   is issued BEFORE
//                    O(k)   = P(k) V         (M=128, N=64,  K=KP, A = P from TMEM, V MN-major)
//                so the next unit's scores are ready when the softmax warps finish unit k
//   warp 2       TMEM allocation: S buffers at columns 0 and 224, O at 448
//   warps 4-7    softmax of the LEFT half of the key columns of every unit (thread = query row)
//   warps 8-11   softmax of the RIGHT half; the two halves exchange the row maxima through shared
//                memory (mbarrier-synchronised) and the row sums for the epilogue
// P (bf16 pairs) is written back into TMEM over each half's own score columns (tcgen05.st), the
// epilogue of unit k (O / row sum -> bf16 -> global) is deferred until after the softmax of unit
// k+1 so the P V MMAs are hidden too.  Two warp groups on one tile halve the per-unit softmax
// latency; the earlier one-tile-per-warp-group version was bound by the serial chain
// S -> max -> exp -> P V -> epilogue (profiles/r01_v6_attention.md).
//
// Softmax (R/ViT_seq.c:204-234): scores are scaled by 1/sqrt(64) after the dot product;
// p = exp(s/8 - max/8) is evaluated as exp2((s - max) * log2(e)/8); padded keys (j >= T) get
// p = 0; the division by the row sum is applied to O in the epilogue.
#include "tc_common.cuh"

using namespace vitcu;
using namespace vitcu::tc;

namespace {

constexpr int kThreadsAttn = 384;
constexpr int QT = 128;                   // queries per tile
constexpr uint32_t Q_BYTES = QT * 128;    // [128 x 64] bf16
constexpr uint32_t S_STRIDE = 224;         // score buffers at columns 0 and 224 (<= 224 columns each)
constexpr uint32_t O_COL = 448;           // output accumulator, 64 columns
#ifndef VITCU_ATTN_POLY_MASK
// Bit j set = pair j of every 8 evaluates 2^x with an FMA-pipe polynomial instead of MUFU.EX2.
// Measured (profiles/r01_v6_attention.md): pass 2 takes ~1 950 cycles per unit for any mix from
// 0/8 to 3/8 -- the two pipes do not overlap well with two softmax warps per sub-partition -- so
// the default keeps everything on the MUFU.
#define VITCU_ATTN_POLY_MASK 0x00
#endif
constexpr uint32_t kPolyMask = VITCU_ATTN_POLY_MASK;
#ifndef VITCU_ATTN_SCALAR
#define VITCU_ATTN_SCALAR 0
#endif

// barriers: per smem stage {QK_FULL, V_FULL, SMEM_FREE}; per score buffer {S_FULL}; P_FULL, O_FULL, O_FREE, XCHG
enum Bar { QK_FULL = 0, V_FULL = 2, SMEM_FREE = 4, S_FULL = 6, P_FULL = 8, O_FULL = 9, O_FREE = 10, XCHG = 11, NUM_BARS = 12 };

struct AttnParams {
    int batch, tokens, kp;     // kp = tokens rounded up to a multiple of 16
    int items;                 // batch * heads
    int rev;                   // 1: walk the items from the last image down (the freshest QKV rows are still in L2)
    __nv_bfloat16 *out;        // [B*T, 768]
    unsigned long long *dbg;   // optional timeline of CTA 0 (clock64 stamps), see vitcu_attention_debug_timeline
};

__device__ __forceinline__ float max3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__device__ __forceinline__ int item_of(const AttnParams &p, int raw) { return p.rev ? p.items - 1 - raw : raw; }

__device__ __forceinline__ void stamp(const AttnParams &p, int role, int k, int ev)
{
    // role 0..3 (0: left softmax warp 4, 1: right softmax warp 8, 2: MMA issuer), 16 units x 8 events each
    if (p.dbg && blockIdx.x == 0 && k < 16 && (threadIdx.x & 31) == 0)
        p.dbg[(role * 16 + k) * 8 + ev] = clock64();
}

// Softmax of one warp group's NC 16-column chunks of a unit.  Returns the partial row sum.
// Pass 1 leaves the row maximum of the own columns in pmax; the caller exchanges it with the other
// warp group between the passes through `exchange`.
template <int NC, typename Exchange>
__device__ __forceinline__ float softmax_half(uint32_t taddr_s, int valid_cols, float sl2, Exchange &&exchange)
{
    if (NC == 0) {
        exchange(-INFINITY);
        return 0.f;
    }
    // The whole half-row (<= 112 scores) is pulled into registers with back-to-back tcgen05.ld and ONE
    // wait: the per-chunk load -> wait -> compute chain of the earlier versions paid the TMEM load
    // latency (~200 cycles under MMA traffic) 14 times per unit and bounded the kernel
    // (profiles/r01_v6_attention.md); both softmax passes now run from registers.
    uint32_t sc[NC > 0 ? NC : 1][16];
#pragma unroll
    for (int c = 0; c < NC; c++)
        tmem_ld_32x32b_x16(taddr_s + c * 16, sc[c]);
    tmem_ld_wait();
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        if (c + 1 < NC) { // only the last chunk of a half can hold padded keys
#pragma unroll
            for (int j = 0; j < 16; j += 2)
                mx = max3(mx, __uint_as_float(sc[c][j]), __uint_as_float(sc[c][j + 1]));
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (c * 16 + j < valid_cols)
                    mx = fmaxf(mx, __uint_as_float(sc[c][j]));
        }
    }
    mx = exchange(mx); // row maximum over BOTH halves
#if VITCU_ATTN_SCALAR
    const float nm = -mx * sl2;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        uint32_t packed[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float a0 = fmaf(__uint_as_float(sc[c][2 * j]), sl2, nm), a1 = fmaf(__uint_as_float(sc[c][2 * j + 1]), sl2, nm);
            float e0, e1;
            if (kPolyMask & (1 << j)) {
                e0 = exp2_poly(a0);
                e1 = exp2_poly(a1);
            } else {
                e0 = ex2_approx(a0);
                e1 = ex2_approx(a1);
            }
            if (c + 1 == NC) {
                if (c * 16 + 2 * j >= valid_cols)
                    e0 = 0.f;
                if (c * 16 + 2 * j + 1 >= valid_cols)
                    e1 = 0.f;
            }
            s0 += e0;
            s1 += e1;
            packed[j] = pack_bf16x2(e0, e1);
        }
        tmem_st_32x32b_x8(taddr_s + c * 8, packed);
    }
    return s0 + s1;
#else
    const f32x2 sl2v = pack2(sl2, sl2), nmx = pack2(-mx * sl2, -mx * sl2);
    f32x2 sum2 = pack2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NC; c++) {
        uint32_t packed[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const f32x2 arg = fma2(pack2(__uint_as_float(sc[c][2 * j]), __uint_as_float(sc[c][2 * j + 1])), sl2v, nmx);
            float e0, e1;
            if (kPolyMask & (1 << j)) { // this pair on the FMA pipe, the others on the MUFU
                exp2_poly2(arg, e0, e1);
            } else {
                float a0, a1;
                unpack2(arg, a0, a1);
                e0 = ex2_approx(a0);
                e1 = ex2_approx(a1);
            }
            if (c + 1 == NC) {
                if (c * 16 + 2 * j >= valid_cols)
                    e0 = 0.f;
                if (c * 16 + 2 * j + 1 >= valid_cols)
                    e1 = 0.f;
            }
            sum2 = add2(sum2, pack2(e0, e1));
            packed[j] = pack_bf16x2(e0, e1);
        }
        // P (bf16 pairs) back into TMEM over this half's own score columns
        tmem_st_32x32b_x8(taddr_s + c * 8, packed);
    }
    float s0, s1;
    unpack2(sum2, s0, s1);
    return s0 + s1;
#endif
}

// NCH = number of 16-column chunks of S (KP / 16); the left warp group takes ceil(NCH/2) chunks
template <int NCH>
__global__ void __launch_bounds__(kThreadsAttn, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const __grid_constant__ CUtensorMap tmap_out, const AttnParams p, uint32_t *watchdog_flag)
{
    constexpr int NC0 = (NCH + 1) / 2, NC1 = NCH / 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t kv_bytes = static_cast<uint32_t>(p.kp) * 128u;
    const uint32_t stage_bytes = 2 * Q_BYTES + 2 * kv_bytes; // Q0 | Q1 | K | V
    uint8_t *ostage = smem + 2 * stage_bytes;                   // [8 softmax warps][32 rows x 64 B], 64B-swizzled, 1 KB aligned
    uint64_t *bars = reinterpret_cast<uint64_t *>(ostage + 8 * 2048);
    float *xmax = reinterpret_cast<float *>(bars + NUM_BARS);   // [2 unit parity][2 halves][128 rows]
    float *xsum = xmax + 2 * 2 * QT;                            // [2 unit parity][2 halves][128 rows]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(xsum + 2 * 2 * QT);
    volatile uint32_t *cta_abort = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (p.tokens + QT - 1) / QT; // 1 or 2
    const int n_items = blockIdx.x < (unsigned)p.items ? (p.items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int n_units = n_items * ntiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(&bars[QK_FULL + s], 1);
            mbar_init(&bars[V_FULL + s], 1);
            mbar_init(&bars[SMEM_FREE + s], 1);
            mbar_init(&bars[S_FULL + s], 1);
        }
        mbar_init(&bars[P_FULL], 8);
        mbar_init(&bars[O_FULL], 1);
        mbar_init(&bars[O_FREE], 8);
        mbar_init(&bars[XCHG], 8);
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
        for (int il = 0; il < n_items; il++) {
            const int item = item_of(p, blockIdx.x + il * gridDim.x);
            const int img = item / kHeads, head = item - img * kHeads;
            const uint32_t s = il & 1;
            // stage s was last used by item il-2: wait until its MMAs have retired
            if (il >= 2 && !mbar_wait_warp(&bars[SMEM_FREE + s], ((il >> 1) - 1) & 1, wd, 1))
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
        bool ok = true;
        auto issue_s = [&](int k) { // scores of unit k into buffer k & 1
            const int il = k / ntiles, t = k - il * ntiles;
            const uint32_t s = il & 1;
            if (t == 0 && !(ok = mbar_wait_warp(&bars[QK_FULL + s], (il >> 1) & 1, wd, 2)))
                return;
            tcgen05_fence_after();
            if (elect_one()) {
                const uint32_t sq = smem_u32(smem + s * stage_bytes);
                const uint64_t q_desc = umma_desc_k_sw128(sq + t * Q_BYTES);
                const uint64_t k_desc = umma_desc_k_sw128(sq + 2 * Q_BYTES);
#pragma unroll
                for (int kk = 0; kk < kHeadDim / 16; kk++)
                    umma_bf16_ss(tmem_base + (k & 1) * S_STRIDE, q_desc + 2 * kk, k_desc + 2 * kk, idesc_s, kk != 0);
                umma_commit(&bars[S_FULL + (k & 1)]);
            }
            __syncwarp();
            stamp(p, 2, k, 0); // S(k) issued
        };
        if (n_units > 0)
            issue_s(0);
        for (int k = 0; k < n_units && ok; k++) {
            if (k + 1 < n_units) {
                issue_s(k + 1); // in order after P V of unit k-1, which read the same score buffer
                if (!ok)
                    break;
            }
            const int il = k / ntiles, t = k - il * ntiles;
            const uint32_t s = il & 1;
            if (t == 0 && !(ok = mbar_wait_warp(&bars[V_FULL + s], (il >> 1) & 1, wd, 3)))
                break;
            if (!(ok = mbar_wait_warp(&bars[P_FULL], k & 1, wd, 4)))
                break;
            stamp(p, 2, k, 1); // P(k) seen
            if (k > 0 && !(ok = mbar_wait_warp(&bars[O_FREE], (k - 1) & 1, wd, 5))) // epilogue of unit k-1 has read O
                break;
            tcgen05_fence_after();
            if (elect_one()) {
                const uint32_t sv = smem_u32(smem + s * stage_bytes) + 2 * Q_BYTES + kv_bytes;
                const uint32_t sb = tmem_base + (k & 1) * S_STRIDE;
#pragma unroll
                for (int kk = 0; kk < NCH; kk++) {
                    // 16 keys per step: 8 packed P columns, left half at the buffer start, right half at its own start
                    const uint32_t a = kk < NC0 ? sb + kk * 8 : sb + NC0 * 16 + (kk - NC0) * 8;
                    umma_bf16_ts(tmem_base + O_COL, a, umma_desc_mn_sw128(sv + kk * 2048), idesc_o, kk != 0);
                }
                umma_commit(&bars[O_FULL]);
                if (t == ntiles - 1)
                    umma_commit(&bars[SMEM_FREE + s]); // Q/K/V of this stage may be overwritten
            }
            __syncwarp();
            stamp(p, 2, k, 2); // PV(k) issued
        }
    } else if (warp >= 4) {
        // ===================== softmax + epilogue =====================
        const int wg = (warp - 4) >> 2;     // 0: left half of the key columns, 1: right half
        const int quad = warp & 3;          // TMEM lane quadrant of this warp
        const int row = quad * 32 + lane;   // query row inside the tile
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const float sl2 = 0.125f * 1.4426950408889634f; // log2(e) / sqrt(64)
        const int col0 = wg * NC0 * 16;                  // first score column of this half
        const int valid_cols = max(0, min(p.tokens - col0, (wg ? NC1 : NC0) * 16));
        bool ok = true;

        // epilogue of unit j: O / sum -> bf16; this warp group stores 32 of the 64 head dims
        auto epilogue = [&](int j) {
            if (!(ok = mbar_wait_warp(&bars[O_FULL], j & 1, wd, 7)))
                return;
            tcgen05_fence_after();
            uint32_t v[32];
            tmem_ld_32x32b_x32(lane_addr + O_COL + wg * 32, v);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&bars[O_FREE]); // O may be overwritten by the next P V
            const float *xs = xsum + (j & 1) * 2 * QT;
            const float inv = 1.0f / (xs[row] + xs[QT + row]);
            const int il = j / ntiles, t = j - il * ntiles;
            const int item = item_of(p, blockIdx.x + il * gridDim.x);
            const int img = item / kHeads, head = item - img * kHeads;
            // Row-per-thread 16-byte global stores touch 32 different lines per instruction (1 024 L1
            // wavefronts per unit, ~1 000-2 000 cycles on the timeline); instead every warp puts its
            // 32 rows x 64 B into a swizzled shared-memory tile and the TMA engine writes it, clipped at
            // T by the 3-D tensor map.
            uint8_t *tile = ostage + (warp - 4) * 2048;
            if (lane == 0)
                tma_wait_group_read<0>(); // the previous unit's store has read this tile
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; i++)
                *reinterpret_cast<uint4 *>(tile + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) = make_uint4(
                    pack_bf16x2(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv),
                    pack_bf16x2(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv),
                    pack_bf16x2(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv),
                    pack_bf16x2(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_3d(&tmap_out, tile, head * kHeadDim + wg * 32, t * QT + quad * 32, img);
                tma_commit_group();
            }
        };

        for (int k = 0; k < n_units && ok; k++) {
            stamp(p, wg, k, 0); // start waiting for S(k)
            if (!(ok = mbar_wait_warp(&bars[S_FULL + (k & 1)], (k >> 1) & 1, wd, 6)))
                break;
            tcgen05_fence_after();
            if (quad == 0)
                stamp(p, wg, k, 1); // S(k) ready
            float *xm = xmax + (k & 1) * 2 * QT;
            auto exchange = [&](float pmax) -> float {
                if (quad == 0)
                    stamp(p, wg, k, 2); // pass 1 done
                xm[wg * QT + row] = pmax;
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&bars[XCHG]);
                ok = mbar_wait_warp(&bars[XCHG], k & 1, wd, 8);
                if (quad == 0)
                    stamp(p, wg, k, 3); // exchange done
                return ok ? fmaxf(pmax, xm[(1 - wg) * QT + row]) : pmax;
            };
            const uint32_t ts = lane_addr + (k & 1) * S_STRIDE + col0;
            const float psum = wg == 0 ? softmax_half<NC0>(ts, valid_cols, sl2, exchange)
                                       : softmax_half<NC1>(ts, valid_cols, sl2, exchange);
            if (!ok)
                break;
            if (quad == 0)
                stamp(p, wg, k, 4); // pass 2 issued
            xsum[(k & 1) * 2 * QT + wg * QT + row] = psum;
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&bars[P_FULL]);
            if (quad == 0)
                stamp(p, wg, k, 5); // P(k) published
            if (k > 0)
                epilogue(k - 1); // deferred: the P V of unit k-1 ran under this unit's softmax
            if (quad == 0)
                stamp(p, wg, k, 6); // deferred epilogue done
        }
        if (ok && n_units > 0) {
            // the other half's row sums of the last unit: one more exchange round orders them
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&bars[XCHG]);
            if (mbar_wait_warp(&bars[XCHG], n_units & 1, wd, 9))
                epilogue(n_units - 1);
        }
        if (lane == 0)
            tma_wait_group<0>(); // this warp's output stores have landed before the CTA retires
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

unsigned long long *g_attn_dbg = nullptr; // set by vitcu_attention_debug_timeline

// qkv [B*T, 2304] bf16 -> out [B*T, 768] bf16; tokens <= 224 (two score buffers of <= 224 columns)
int attention_bf16_tc(const void *qkv, void *out, int batch, int tokens, cudaStream_t st)
{
    const int kp = (tokens + 15) / 16 * 16;
    VITCU_REQUIRE(kp <= 224, "single-block tensor-core attention handles at most 224 tokens");
    VITCU_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0, "buffers must be 16-byte aligned");
    CUtensorMap tq, tkv;
    int rc = make_qkv_map(&tq, qkv, batch, tokens, QT);
    if (rc)
        return rc;
    rc = make_qkv_map(&tkv, qkv, batch, tokens, (uint32_t)kp);
    if (rc)
        return rc;
    const size_t smem = 2 * (2 * (size_t)Q_BYTES + 2 * (size_t)kp * 128) + NUM_BARS * 8 + 2 * 2 * 2 * QT * sizeof(float) +
                        8 * 2048 + 16 + 1024;
    CUtensorMap tout; // out viewed as [B][T][768]: box 32 rows x 32 columns, rows past T are clipped
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
        cuuint64_t dims[3] = {kEmbed, (cuuint64_t)tokens, (cuuint64_t)batch};
        cuuint64_t strides[2] = {kEmbed * 2, (cuuint64_t)tokens * kEmbed * 2};
        cuuint32_t box[3] = {32, 32, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        if (fn(&tout, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return set_error(VITCU_E_ARG, __FILE__, __LINE__, "cuTensorMapEncodeTiled rejected the output tensor");
    }
    VITCU_REQUIRE(smem <= 227 * 1024, "attention tile does not fit shared memory");
    AttnParams p;
    p.batch = batch;
    p.tokens = tokens;
    p.kp = kp;
    p.items = batch * kHeads;
    p.out = reinterpret_cast<__nv_bfloat16 *>(out);
    p.dbg = g_attn_dbg;
    static const bool serp = !(getenv("VITCU_SERPENTINE") && atoi(getenv("VITCU_SERPENTINE")) == 0);
    p.rev = serp;
    const int sms = device_sm_count();
    const int grid = p.items < sms ? p.items : sms;
    int dev = 0;
    VITCU_TRY(cudaGetDevice(&dev));
    const int nch = kp / 16;
#define VITCU_ATTN_CASE(N)                                                                                          \
    case N: {                                                                                                       \
        static int configured[64] = {0};                                                                            \
        if (dev < 64 && configured[dev] < (int)smem) {                                                              \
            VITCU_TRY(cudaFuncSetAttribute(attention_tc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                           (int)smem));                                                             \
            configured[dev] = (int)smem;                                                                            \
        }                                                                                                           \
        attention_tc_kernel<N><<<grid, kThreadsAttn, smem, st>>>(tq, tkv, tout, p, watchdog_flag());                     \
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
        VITCU_ATTN_CASE(9)
        VITCU_ATTN_CASE(10)
        VITCU_ATTN_CASE(11)
        VITCU_ATTN_CASE(12)
        VITCU_ATTN_CASE(13)
        VITCU_ATTN_CASE(14)
    default:
        return set_error(VITCU_E_ARG, __FILE__, __LINE__, "unsupported key count");
    }
#undef VITCU_ATTN_CASE
    VITCU_LAUNCHED();
    return 0;
}

} // namespace vitcu

// Debug aid: when `buffer` (device memory, 3*16*8 u64) is non-NULL the next launches of the
// single-block attention kernel record clock64 stamps of CTA 0's first 16 units
// (tools/attn_timeline.py prints them).  Pass NULL to switch it off.
extern "C" int vitcu_attention_debug_timeline(unsigned long long *buffer)
{
    vitcu::g_attn_dbg = buffer;
    return 0;
}
