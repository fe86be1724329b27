// csrc/attention_tc.cu -- fused multi-head attention on the 5th-gen tensor cores
// (BF16 operands, FP32 accumulation and softmax), for token counts <= 224.
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
//   warps 4-7    exponentials of the LEFT half of the key columns of every unit (thread = query row)
//   warps 8-11   exponentials of the RIGHT half; row sums left in shared memory for the epilogue
//   warps 12-15  statistics + epilogue: read O(k-2) and free the accumulator, row maxima of S(k) while
//                warps 4-11 are still on unit k-1, then O(k-2) / row sum -> bf16 -> swizzled smem
//                tile -> TMA store
// P (bf16 pairs) is written back into TMEM over each half's own score columns (tcgen05.st).  Two
// warp groups on one tile halve the per-unit exp latency; the fourth warp group takes the max pass,
// the exchange of the half-row maxima and the ~900-cycle O read-out off their chain.  Measured on one
// box this layout is exactly as fast as the three-warp-group version it replaced (the warps of a
// sub-partition stretch each other's phases); it is kept because the roles are cleaner, not because
// it is faster (profiles/r01_v6_attention.md has the A/B and two further experiments).
// Registers: the kernel starts with 128 per thread (512 threads); setmaxnreg moves them to where
// the score rows live: warps 0-3 keep 40, the exp warp groups 168 each, the statistics warps 136
// (40 + 136 + 2 x 168 = 4 x 128).
//
// Softmax (R/ViT_seq.c:204-234): scores are scaled by 1/sqrt(64) after the dot product;
// p = exp(s/8 - max/8) is evaluated as exp2((s - max) * log2(e)/8); padded keys (j >= T) get
// p = 0; the division by the row sum is applied to O in the epilogue.
#include "tc_common.cuh"

using namespace vitcu;
using namespace vitcu::tc;

namespace {

constexpr int kThreadsAttn = 512;
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
#ifndef VITCU_ATTN_STAGGER_NS
#define VITCU_ATTN_STAGGER_NS 0 // measured: no effect (the warp groups re-synchronise through P_FULL)
#endif
#ifndef VITCU_ATTN_SCALAR
#define VITCU_ATTN_SCALAR 0
#endif

// barriers: per smem stage {QK_FULL, V_FULL, SMEM_FREE}; per score buffer {S_FULL, MAX_FULL}; P_FULL, O_FULL, O_FREE
enum Bar { QK_FULL = 0, V_FULL = 2, SMEM_FREE = 4, S_FULL = 6, P_FULL = 8, O_FULL = 9, O_FREE = 10, MAX_FULL = 11, NUM_BARS = 13 };

struct AttnParams {
    int batch, tokens, kp;     // kp = tokens rounded up to a multiple of 16
    int items;                 // batch * heads
    int heads, embed;          // embed = heads * 64: Q | K | V column blocks of qkv start at 0, embed, 2 * embed
    int dbg_first, dbg_cta;    // timeline window: units [dbg_first, dbg_first + 16) of CTA dbg_cta
    int order;                 // statistics warps: 0 = max(j) then whole epilogue(j-1); 1 = read O(j-2), max(j), store O(j-2)
    int stagger_ns;            // head start of the left exp warp group
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
    // role 0..3 (0: left exp warp 4, 1: right exp warp 8, 2: MMA issuer, 3: statistics/epilogue warp 12), 16 units x 8 events each
    k -= p.dbg_first;
    if (p.dbg && blockIdx.x == p.dbg_cta && k >= 0 && k < 16 && (threadIdx.x & 31) == 0)
        p.dbg[(role * 16 + k) * 8 + ev] = clock64();
}

// Row maximum over NC 16-column chunks of a score row (thread = row), columns >= valid_cols excluded.
// Four independent running maxima: one chain of 56 dependent 3-input max instructions would cost
// ~300 cycles of pure latency per part.
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

// Exponentials of one warp group's NC 16-column chunks of a unit.  Returns the partial row sum.
// The row maximum comes from the statistics warps (`row_max()` is called once the score loads are in
// flight, so waiting for it overlaps the TMEM load latency).
template <int NC, typename RowMax>
__device__ __forceinline__ float softmax_half(uint32_t taddr_s, int valid_cols, float sl2, RowMax &&row_max)
{
    if (NC == 0) {
        row_max();
        return 0.f;
    }
    // The whole half-row (<= 112 scores) is pulled into registers with back-to-back tcgen05.ld and ONE
    // wait: a per-chunk load -> wait -> compute chain pays the TMEM load latency (~200 cycles under
    // MMA traffic) once per chunk (profiles/r01_v6_attention.md).
    uint32_t sc[NC > 0 ? NC : 1][16];
#pragma unroll
    for (int c = 0; c < NC; c++)
        tmem_ld_32x32b_x16(taddr_s + c * 16, sc[c]);
    const float mx = row_max(); // row maximum over ALL key columns
    tmem_ld_wait();
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
    uint8_t *ostage = smem + 2 * stage_bytes;                   // [4 epilogue warps][32 rows x 128 B], 128B-swizzled, 1 KB aligned
    uint64_t *bars = reinterpret_cast<uint64_t *>(ostage + 8 * 2048);
    float *xmax = reinterpret_cast<float *>(bars + NUM_BARS);   // [2 unit parity][128 rows] (2 x 2 x 128 floats reserved)
    // row sums: 3 slots.  Slot reuse is ordered by the MMA chain: S(k+3) is issued after P V(k+1), which
    // waited for O_FREE(k), which the epilogue warps arrive on after reading the sums of unit k.
    float *xsum = xmax + 2 * 2 * QT;                            // [unit % 3][2 halves][128 rows]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(xsum + 3 * 2 * QT);
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
        mbar_init(&bars[O_FREE], 4);
        mbar_init(&bars[MAX_FULL + 0], 4);
        mbar_init(&bars[MAX_FULL + 1], 4);
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

    // register hand-over between the warp groups: every role branch below starts with its
    // setmaxnreg (all four warps of a warp group take the same one)
    if (warp < 4) {
      setmaxnreg_dec<40>();
      if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            prefetch_tensormap(&tmap_q);
            prefetch_tensormap(&tmap_kv);
        }
        for (int il = 0; il < n_items; il++) {
            const int item = item_of(p, blockIdx.x + il * gridDim.x);
            const int img = item / p.heads, head = item - img * p.heads;
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
                tma_load_3d(sk, &tmap_kv, &bars[QK_FULL + s], p.embed + head * kHeadDim, 0, img);
                mbar_arrive_expect_tx(&bars[V_FULL + s], kv_bytes);
                tma_load_3d(sk + kv_bytes, &tmap_kv, &bars[V_FULL + s], 2 * p.embed + head * kHeadDim, 0, img);
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
      }
    } else if (warp >= 12) {
        setmaxnreg_inc<136>();
        // ===================== statistics (row maxima) + epilogue (O / sum -> bf16 -> global) =====================
        const int quad = warp & 3;          // TMEM lane quadrant of this warp
        const int row = quad * 32 + lane;   // query row inside the tile
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        // Row-per-thread 16-byte global stores touch 32 different lines per instruction (1 024 L1
        // wavefronts per unit); instead every warp puts its 32 rows x 128 B into a swizzled
        // shared-memory tile and the TMA engine writes it, clipped at T by the 3-D tensor map.
        uint8_t *tile = ostage + quad * 4096;
        // The read-out of O(j) is split in two so that the row maxima (which the exp warps wait for) sit
        // between the halves: o_read frees the accumulator for P V(j+1) within ~300 cycles of O_FULL and
        // keeps the normalised row as 32 packed bf16 pairs; o_store writes them out after the max pass.
        uint32_t pk[32];
        auto o_read = [&](int j) -> bool {
            // O_FULL(j) follows P_FULL(j), which the exp warps arrive on after writing their sums
            if (!mbar_wait_warp(&bars[O_FULL], j & 1, wd, 7))
                return false;
            tcgen05_fence_after();
            uint32_t vlo[32], vhi[32];
            tmem_ld_32x32b_x32(lane_addr + O_COL, vlo);
            tmem_ld_32x32b_x32(lane_addr + O_COL + 32, vhi);
            const float *xs = xsum + (j % 3) * 2 * QT;
            const float inv = 1.0f / (xs[row] + xs[QT + row]);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&bars[O_FREE]); // O and this unit's sums have been read: P V(j+1) may overwrite O
#pragma unroll
            for (int i = 0; i < 16; i++) {
                pk[i] = pack_bf16x2(__uint_as_float(vlo[2 * i]) * inv, __uint_as_float(vlo[2 * i + 1]) * inv);
                pk[16 + i] = pack_bf16x2(__uint_as_float(vhi[2 * i]) * inv, __uint_as_float(vhi[2 * i + 1]) * inv);
            }
            return true;
        };
        auto o_store = [&](int j) {
            const int il = j / ntiles, t = j - il * ntiles;
            const int item = item_of(p, blockIdx.x + il * gridDim.x);
            const int img = item / p.heads, head = item - img * p.heads;
            if (lane == 0)
                tma_wait_group_read<0>(); // the previous unit's store has read this tile
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; i++)
                *reinterpret_cast<uint4 *>(tile + lane * 128 + ((i ^ (lane & 7)) << 4)) =
                    make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_3d(&tmap_out, tile, head * kHeadDim, t * QT + quad * 32, img);
                tma_commit_group();
            }
            if (quad == 0)
                stamp(p, 3, j, 6); // epilogue of unit j issued
        };
        // S(j) and O(j-2) complete at about the same time: both follow the issue of P V(j-2)
        bool ok = true;
        const int lag = p.order ? 2 : 1;
        for (int j = 0; j < n_units && ok; j++) {
            if (quad == 0)
                stamp(p, 3, j, 0);
            if (p.order && j > 1 && !(ok = o_read(j - 2)))
                break;
            // ---- row maxima of unit j (the exp warps are still on unit j-1) ----
            if (!(ok = mbar_wait_warp(&bars[S_FULL + (j & 1)], (j >> 1) & 1, wd, 6)))
                break;
            tcgen05_fence_after();
            const uint32_t ts = lane_addr + (j & 1) * S_STRIDE;
            constexpr int R0 = (NCH + 2) / 3, R1 = (NCH - R0 + 1) / 2, R2 = NCH - R0 - R1; // three rounds of <= 5 chunks
            float mx = row_max_part<R0>(ts, min(p.tokens, R0 * 16), -INFINITY);
            mx = row_max_part<R1>(ts + R0 * 16, max(0, min(p.tokens - R0 * 16, R1 * 16)), mx);
            mx = row_max_part<R2>(ts + (R0 + R1) * 16, max(0, min(p.tokens - (R0 + R1) * 16, R2 * 16)), mx);
            // slot j & 1 was last read for unit j-2, whose P_FULL preceded the issue of S(j)
            xmax[(j & 1) * QT + row] = mx;
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&bars[MAX_FULL + (j & 1)]);
            if (quad == 0)
                stamp(p, 3, j, 1);
            if (p.order) {
                if (j > 1)
                    o_store(j - 2);
            } else if (j > 0 && (ok = o_read(j - 1))) {
                o_store(j - 1);
            }
        }
        for (int j = max(0, n_units - lag); j < n_units && ok; j++) {
            if ((ok = o_read(j)))
                o_store(j);
        }
        if (lane == 0)
            tma_wait_group<0>(); // this warp's output stores have landed before the CTA retires
    } else {
        setmaxnreg_inc<168>();
        // ===================== exponentials: P = exp2((S - max) * log2(e)/8), row sums =====================
        const int wg = (warp - 4) >> 2;     // 0: left half of the key columns, 1: right half
        const int quad = warp & 3;          // TMEM lane quadrant of this warp
        const int row = quad * 32 + lane;   // query row inside the tile
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const float sl2 = 0.125f * 1.4426950408889634f; // log2(e) / sqrt(64)
        const int col0 = wg * NC0 * 16;                  // first score column of this half
        const int valid_cols = max(0, min(p.tokens - col0, (wg ? NC1 : NC0) * 16));
        bool ok = true;

        // Both halves share the MUFU of their sub-partition; started together they run their ~700 cycles
        // of per-unit latency (publish, barrier waits, TMEM loads) at the same time and the MUFU idles.
        // Half a period of head start for the left half interleaves the two.
        if (wg == 1 && p.stagger_ns > 0)
            __nanosleep(p.stagger_ns);
        for (int k = 0; k < n_units && ok; k++) {
            stamp(p, wg, k, 0); // start waiting for S(k)
            if (!(ok = mbar_wait_warp(&bars[S_FULL + (k & 1)], (k >> 1) & 1, wd, 6)))
                break;
            tcgen05_fence_after();
            if (quad == 0)
                stamp(p, wg, k, 1); // S(k) ready
            auto row_max = [&]() -> float {
                ok = mbar_wait_warp(&bars[MAX_FULL + (k & 1)], (k >> 1) & 1, wd, 8);
                if (quad == 0)
                    stamp(p, wg, k, 3); // row maxima available
                return ok ? xmax[(k & 1) * QT + row] : 0.f;
            };
            const uint32_t ts = lane_addr + (k & 1) * S_STRIDE + col0;
            const float psum = wg == 0 ? softmax_half<NC0>(ts, valid_cols, sl2, row_max)
                                       : softmax_half<NC1>(ts, valid_cols, sl2, row_max);
            if (!ok)
                break;
            if (quad == 0)
                stamp(p, wg, k, 4); // pass 2 issued
            xsum[(k % 3) * 2 * QT + wg * QT + row] = psum;
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&bars[P_FULL]);
            if (quad == 0)
                stamp(p, wg, k, 5); // P(k) published
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

int make_qkv_map(CUtensorMap *map, const void *qkv, int batch, int tokens, int embed, uint32_t box_rows)
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
    const cuuint64_t ld = 3 * (cuuint64_t)embed;
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
int attention_bf16_tc(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t st)
{
    const int embed = heads * kHeadDim;
    const int kp = (tokens + 15) / 16 * 16;
    VITCU_REQUIRE(kp <= 224, "single-block tensor-core attention handles at most 224 tokens");
    VITCU_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0, "buffers must be 16-byte aligned");
    CUtensorMap tq, tkv;
    int rc = make_qkv_map(&tq, qkv, batch, tokens, embed, QT);
    if (rc)
        return rc;
    rc = make_qkv_map(&tkv, qkv, batch, tokens, embed, (uint32_t)kp);
    if (rc)
        return rc;
    const size_t smem = 2 * (2 * (size_t)Q_BYTES + 2 * (size_t)kp * 128) + NUM_BARS * 8 + (2 + 3) * 2 * QT * sizeof(float) +
                        4 * 4096 + 16 + 1024;
    CUtensorMap tout; // out viewed as [B][T][768]: box 32 rows x 64 columns (one head), rows past T are clipped
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
        cuuint64_t dims[3] = {(cuuint64_t)embed, (cuuint64_t)tokens, (cuuint64_t)batch};
        cuuint64_t strides[2] = {(cuuint64_t)embed * 2, (cuuint64_t)tokens * embed * 2};
        cuuint32_t box[3] = {kHeadDim, 32, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        if (fn(&tout, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return set_error(VITCU_E_ARG, __FILE__, __LINE__, "cuTensorMapEncodeTiled rejected the output tensor");
    }
    VITCU_REQUIRE(smem <= 227 * 1024, "attention tile does not fit shared memory");
    AttnParams p;
    p.batch = batch;
    p.tokens = tokens;
    p.kp = kp;
    p.items = batch * heads;
    p.heads = heads;
    p.embed = embed;
    p.out = reinterpret_cast<__nv_bfloat16 *>(out);
    p.dbg = g_attn_dbg;
    static const bool serp = !(getenv("VITCU_SERPENTINE") && atoi(getenv("VITCU_SERPENTINE")) == 0);
    p.rev = serp;
    static const int order = getenv("VITCU_ATTN_ORDER") ? atoi(getenv("VITCU_ATTN_ORDER")) : 1;
    static const int stagger = getenv("VITCU_ATTN_STAGGER_NS") ? atoi(getenv("VITCU_ATTN_STAGGER_NS")) : VITCU_ATTN_STAGGER_NS;
    p.order = order;
    p.dbg_first = getenv("VITCU_ATTN_DBG_FIRST") ? atoi(getenv("VITCU_ATTN_DBG_FIRST")) : 0;
    p.dbg_cta = getenv("VITCU_ATTN_DBG_CTA") ? atoi(getenv("VITCU_ATTN_DBG_CTA")) : 0;
    p.stagger_ns = stagger;
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
        VITCU_TRY(launch_kernel(attention_tc_kernel<N>, grid, kThreadsAttn, smem, st, tq, tkv, tout, p, watchdog_flag()));                     \
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
    VITCU_LAUNCHED_KIND(LK_ATTN_TC);
    return 0;
}

} // namespace vitcu

// Debug aid: when `buffer` (device memory, 4*16*8 u64) is non-NULL the next launches of the
// single-block attention kernel record clock64 stamps of CTA 0's first 16 units
// (tools/attn_timeline.py prints them).  Pass NULL to switch it off.
extern "C" int vitcu_attention_debug_timeline(unsigned long long *buffer)
{
    vitcu::g_attn_dbg = buffer;
    return 0;
}
