// csrc/attention_duo_tc.cu -- fused multi-head attention on the 5th-gen tensor cores for token
// counts <= 224, TWO co-resident CTAs per SM ("duo"): the default BF16 attention kernel.
//
// Replaces QKV_TO_SCOREV (R/multihead.cl:65-137); oracle R/ViT_seq.c:192-262.
// R/ = /root/reference/MulticoreMainProject/.  S = Q K^T, the row softmax and O = P V never leave
// the SM: S, P and O all live in tensor memory.
//
// Why two CTAs per SM.  Per 128-query unit the work is ~420 cycles of tcgen05.mma for S, 26.6 k
// exponentials (MUFU.EX2 runs at 16 per clock per SM on B200: a 1 664-cycle floor), ~450 cycles of
// P V and a TMEM read-out.  The one-CTA kernel (attention_tc.cu) software-pipelines those phases over
// two score buffers and four warp groups and still leaves the MUFU idle for half of every period:
// its phases wait on each other through a chain of seven barriers (profiles/r01_v6_attention.md).
// Here every CTA runs the phases of a unit one after the other and the overlap comes from the
// hardware: each CTA takes 256 of the 512 TMEM columns and < 113 KB of shared memory, so two of them
// share an SM and while one is in its MUFU-bound exponential pass the other one runs MMAs, reads its
// output or loads its next item.  What is left on a CTA's own chain is kept short:
//   - there is no MMA warp: the softmax warps meet at a named barrier and one elected thread of warp
//     0 issues the MMAs on the spot (a hand-over to a parked issuer warp and back cost ~1 000 cycles
//     per MMA batch in the first version of this kernel, profiles/r02_attention.md);
//   - P V is issued in two halves, the first (keys 0..127) while the exponentials of the remaining
//     keys are still being computed;
//   - S of the next unit is issued right after the output has been read out of tensor memory and
//     runs under the normalise / store epilogue;
//   - the waits for an MMA batch spin instead of parking the warp.
//
// One CTA = 5 warps, persistent over (image, head) items; a UNIT is one 128-query tile of an item.
//   warps 0-3   softmax + epilogue, thread = query row (warp w owns TMEM lanes 32w..32w+31):
//               pass 1 row maximum (scores streamed from TMEM), pass 2 p = exp2((s - max) log2(e)/8)
//               with fp32 row sum, bf16 P written back INTO TMEM over the scores already consumed;
//               then O / sum -> bf16 -> swizzled shared-memory tile -> TMA store (clipped at T)
//   warp 4      TMA producer (3-D tensor map over qkv [B][T][3*embed], rows past T zero-filled) and
//               TMEM allocation.  Shared memory is single-buffered per CTA; the next item's Q and K
//               are requested as soon as the last S of the current item has retired and its V as
//               soon as the last P V has, so the loads run under the remaining phases of the item
// TMEM columns (256 per CTA): scores at 0..KP-1 (KP <= 224).  P (bf16 pairs, 8 columns per 16 keys)
// overwrites score columns that have already been consumed: keys 0..127 at columns 0..63, the rest at
// 128..; O is accumulated at columns 64..127, which are dead once the scores of keys 0..127 are gone.
//
// Softmax semantics (R/ViT_seq.c:204-234): scores are scaled by 1/sqrt(64) after the dot product;
// padded keys (j >= T) get p = 0; the division by the row sum is applied to O in the epilogue.
#include "tc_common.cuh"

using namespace vitcu;
using namespace vitcu::tc;

namespace {

#ifndef VITCU_CTL_SPIN
#define VITCU_CTL_SPIN 0 // measured: spinning costs 10 % (flash) / 1.5 % (single-block) -- control warp: 1 = spin on its mbarriers (a parked warp wakes up late and everything it issues is on the CTA's chain)
#endif
#if VITCU_CTL_SPIN
#define ctl_wait mbar_wait_spin_warp
#else
#define ctl_wait mbar_wait_warp
#endif

constexpr int kThreadsDuo = 160;
constexpr int QT = 128;                   // queries per tile
constexpr uint32_t Q_BYTES = QT * 128;    // [128 x 64] bf16
constexpr uint32_t O_COL = 64;            // output accumulator, 64 columns, inside the dead part of the score buffer
constexpr uint32_t TMEM_COLS = 256;
constexpr int kHalf = 8;                  // chunks (of 16 keys) in the first P V batch
#ifndef VITCU_DUO_POLY_MASK
// bit j set = pair j of every 8 evaluates 2^x with an FMA-pipe polynomial (relative error 1e-4, far below the bf16
// rounding of P) instead of MUFU.EX2, which is what the exponential pass is bound by
#define VITCU_DUO_POLY_MASK 0x11 // same-box A/B: 0x00 75.0 us, 0x01 72.0, 0x11 70.4, 0x49 70.7, 0x55 74.8 (B=256, T=197)
#endif
constexpr uint32_t kDuoPolyMask = VITCU_DUO_POLY_MASK;
#ifndef VITCU_DUO_P1_ROUNDS
#define VITCU_DUO_P1_ROUNDS 3 // rounds of the row-maximum pass (loads of a round are issued back to back, one wait per round)
#endif

// barriers: per unit {Q_FULL, S_FULL, O_FULL}; per item {K_FULL, V_FULL[item & 1]}
enum Bar { Q_FULL = 0, K_FULL, V_FULL, V_FULL1, S_FULL, O_FULL, NUM_BARS };

struct DuoParams {
    int tokens, kp; // kp = tokens rounded up to a multiple of 16
    int items;      // batch * heads
    int heads, embed;
    int rev;        // 1: walk the items from the last image down (the freshest QKV rows are still in L2)
    int unit_split; // 1: an item is ONE query tile of an (image, head) -- items = batch * heads * tiles -- so that a handful
                    // of images (batch-1 latency: 12 items of two tiles) spreads over twice as many CTAs; K and V are then
                    // loaded once per tile instead of once per (image, head), which only matters when there are many items
};

// (image * heads + head, query tile) of the unit with item slot `raw` and tile `t` inside the item
struct UnitId {
    int item, tile;
};

__device__ __forceinline__ float max3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__device__ __forceinline__ int item_of(const DuoParams &p, int raw) { return p.rev ? p.items - 1 - raw : raw; }
__device__ __forceinline__ UnitId unit_of(const DuoParams &p, int raw, int t, int tiles)
{
    const int it = item_of(p, raw);
    UnitId u;
    u.item = p.unit_split ? it / tiles : it;
    u.tile = p.unit_split ? it - u.item * tiles : t;
    return u;
}

// TMEM column of the 8 packed-bf16 P columns of key chunk c
__device__ __forceinline__ constexpr uint32_t p_col(int c) { return c < kHalf ? 8u * c : 128u + 8u * (c - kHalf); }

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

// Exponentials of key chunks [C0, C1) of one row: chunk c + 1 is in flight while chunk c is processed (the
// load of chunk C0 has been issued by the caller; the load of chunk C1, if it exists, is left in flight).
template <int NCH, int C0, int C1>
__device__ __forceinline__ void exp_chunks(uint32_t lane_addr, uint32_t (&sc)[2][16], f32x2 sl2v, f32x2 nmx, f32x2 &sum2, int tokens)
{
#pragma unroll
    for (int c = C0; c < C1; c++) {
        tmem_ld_wait();
        if (c + 1 < NCH)
            tmem_ld_32x32b_x16(lane_addr + (c + 1) * 16, sc[(c + 1) & 1]);
        uint32_t packed[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const f32x2 arg = fma2(pack2(__uint_as_float(sc[c & 1][2 * j]), __uint_as_float(sc[c & 1][2 * j + 1])), sl2v, nmx);
            float e0, e1;
            if (kDuoPolyMask & (1u << j)) { // this pair on the FMA pipe (degree-3 polynomial), the others on the MUFU
                exp2_poly2(arg, e0, e1);
            } else {
                float a0, a1;
                unpack2(arg, a0, a1);
                e0 = ex2_approx(a0);
                e1 = ex2_approx(a1);
            }
            if (c + 1 == NCH) { // only the last chunk can hold padded keys
                if (c * 16 + 2 * j >= tokens)
                    e0 = 0.f;
                if (c * 16 + 2 * j + 1 >= tokens)
                    e1 = 0.f;
            }
            sum2 = add2(sum2, pack2(e0, e1));
            packed[j] = pack_bf16x2(e0, e1);
        }
        // P (bf16 pairs) over score columns that every warp has consumed already
        tmem_st_32x32b_x8(lane_addr + p_col(c), packed);
    }
}

// NCH = number of 16-column chunks of S (KP / 16)
template <int NCH>
__global__ void __launch_bounds__(kThreadsDuo, 2)
attention_duo_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                        const __grid_constant__ CUtensorMap tmap_out, const DuoParams p, uint32_t *watchdog_flag)
{
    constexpr int H = NCH > kHalf ? kHalf : NCH; // chunks of the first P V batch
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t kv_bytes = static_cast<uint32_t>(p.kp) * 128u;
    uint8_t *sq = smem;                          // Q tile of the current unit [128 x 64] bf16, 128B-swizzled rows
    uint8_t *sk = sq + Q_BYTES;                  // K  [kp x 64]
    uint8_t *sv = sk + kv_bytes;                 // V  [2][kp x 64]: items alternate, so V is requested a whole item ahead
    uint8_t *ostage = sv + 2 * kv_bytes;         // [4 warps][32 rows x 128 B], 128B-swizzled, 1 KB aligned
    uint64_t *bars = reinterpret_cast<uint64_t *>(ostage + 4 * 4096);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NUM_BARS);
    volatile uint32_t *cta_abort = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = (p.tokens + QT - 1) / QT;  // query tiles of an (image, head): 1 or 2
    const int ntiles = p.unit_split ? 1 : tiles; // ... and of an item
    const int n_items = blockIdx.x < (unsigned)p.items ? (p.items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int n_units = n_items * ntiles;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NUM_BARS; i++)
            mbar_init(&bars[i], 1);
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
        // ===================== control warp: TMA loads + MMA issue =====================
        // It meets the softmax warps at the three named barriers of every unit and issues the next MMA
        // batch the moment the barrier opens.  A failed wait raises cta_abort and skips the issue; the
        // warp keeps attending the barriers until the softmax warps leave at their next mbarrier wait.
        const uint32_t idesc_s = umma_idesc_bf16(QT, p.kp, false, false);
        const uint32_t idesc_o = umma_idesc_bf16(QT, kHeadDim, false, true);
        const uint32_t sq_a = smem_u32(sq), sk_a = smem_u32(sk), sv_a = smem_u32(sv);
        if (elect_one()) {
            prefetch_tensormap(&tmap_q);
            prefetch_tensormap(&tmap_kv);
        }
        // operands of unit k: its Q tile, and K and V of its item when it is the item's first tile.  Called once
        // S(k-1) has retired (Q and K buffers dead); the V buffer of item il was last read by item il-2.
        auto load_unit = [&](int k) {
            const int il = k / ntiles, t = k - il * ntiles;
            const UnitId u = unit_of(p, blockIdx.x + il * gridDim.x, t, tiles);
            const int img = u.item / p.heads, head = u.item - img * p.heads;
            if (elect_one()) {
                if (t == 0) { // K first: S needs it together with Q
                    mbar_arrive_expect_tx(&bars[K_FULL], kv_bytes);
                    tma_load_3d(sk, &tmap_kv, &bars[K_FULL], p.embed + head * kHeadDim, 0, img);
                }
                mbar_arrive_expect_tx(&bars[Q_FULL], Q_BYTES);
                tma_load_3d(sq, &tmap_q, &bars[Q_FULL], head * kHeadDim, u.tile * QT, img);
            }
            __syncwarp();
        };
        // V of item il into buffer il & 1, whose last reader was the final P V of item il - 2: requested a whole item
        // ahead (after the first read-out of item il - 1), because under load the 26 KB take several microseconds
        auto load_v = [&](int il) {
            const int item = unit_of(p, blockIdx.x + il * gridDim.x, 0, tiles).item;
            const int img = item / p.heads, head = item - img * p.heads;
            if (elect_one()) {
                mbar_arrive_expect_tx(&bars[V_FULL + (il & 1)], kv_bytes);
                tma_load_3d(sv + (il & 1) * kv_bytes, &tmap_kv, &bars[V_FULL + (il & 1)], 2 * p.embed + head * kHeadDim, 0, img);
            }
            __syncwarp();
        };
        auto issue_s = [&](int k) {
            const int il = k / ntiles, t = k - il * ntiles;
            bool ok = ctl_wait(&bars[Q_FULL], k & 1, wd, 3);
            if (ok && t == 0)
                ok = ctl_wait(&bars[K_FULL], il & 1, wd, 4);
            if (!ok)
                return;
            tcgen05_fence_after();
            if (elect_one()) {
                const uint64_t q_desc = umma_desc_k_sw128(sq_a);
                const uint64_t k_desc = umma_desc_k_sw128(sk_a);
#pragma unroll
                for (int kk = 0; kk < kHeadDim / 16; kk++)
                    umma_bf16_ss(tmem_base, q_desc + 2 * kk, k_desc + 2 * kk, idesc_s, kk != 0);
                umma_commit(&bars[S_FULL]);
            }
            __syncwarp();
        };
        // P V over key chunks [c0, c1): 16 keys per step = 8 packed P columns and 16 V rows of 128 B
        auto issue_pv = [&](int k, int c0, int c1, bool last) {
            const int il = k / ntiles, t = k - il * ntiles;
            if (c0 == 0 && t == 0 && !ctl_wait(&bars[V_FULL + (il & 1)], (il >> 1) & 1, wd, 5))
                return;
            tcgen05_fence_after();
            if (elect_one()) {
                const uint32_t v_a = sv_a + (il & 1) * kv_bytes;
#pragma unroll
                for (int kk = 0; kk < NCH; kk++)
                    if (kk >= c0 && kk < c1)
                        umma_bf16_ts(tmem_base + O_COL, tmem_base + p_col(kk), umma_desc_mn_sw128(v_a + kk * 2048), idesc_o, kk != 0);
                if (last)
                    umma_commit(&bars[O_FULL]);
            }
            __syncwarp();
        };
        if (n_units > 0) {
            load_unit(0);
            load_v(0);
            issue_s(0);
        }
        for (int k = 0; k < n_units; k++) {
            // S(k) retired: the Q buffer (and K, after the item's last tile) can take the next unit's operands
            if (!ctl_wait(&bars[S_FULL], k & 1, wd, 6))
                break; // the softmax warps wait on the same barrier and leave with us
            if (k + 1 < n_units)
                load_unit(k + 1);
            named_bar_sync(1, kThreadsDuo); // scores of keys 0..127 consumed, their P written
            issue_pv(k, 0, H, NCH <= kHalf);
            if (NCH > kHalf) {
                named_bar_sync(1, kThreadsDuo); // all of P written
                issue_pv(k, H, NCH, true);
            }
            named_bar_sync(1, kThreadsDuo); // O read out of tensor memory
            if (k + 1 < n_units)
                issue_s(k + 1);
            if (k % ntiles == 0 && k / ntiles + 1 < n_items)
                load_v(k / ntiles + 1);
            if (*cta_abort)
                break; // the softmax warps leave at their next mbarrier wait, before any further named barrier
        }
    } else {
        // ===================== softmax + epilogue (warps 0-3, thread = query row) =====================
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        const float sl2 = 0.125f * 1.4426950408889634f; // log2(e) / sqrt(64)
        uint8_t *tile = ostage + warp * 4096;
        for (int k = 0; k < n_units; k++) {
            const int il = k / ntiles;
            const UnitId u = unit_of(p, blockIdx.x + il * gridDim.x, k - il * ntiles, tiles);
            const int t = u.tile;
            // a warp whose 32 query rows all lie past T (second tile of a 197-token item: rows 224..255)
            // only keeps the barrier protocol going; its rows of S are exact zeros (Q zero-filled)
            const bool active = t * QT + warp * 32 < p.tokens;
            if (*cta_abort || !mbar_wait_spin_warp(&bars[S_FULL], k & 1, wd, 7))
                break;
            tcgen05_fence_after();
            float inv = 0.f;
            f32x2 sum2 = pack2(0.f, 0.f);
            f32x2 sl2v = pack2(sl2, sl2), nmx = pack2(0.f, 0.f);
            uint32_t sc[2][16];
            if (active) {
                // ---- pass 1: row maximum over the valid key columns, in rounds of <= 5 (or 7) chunks ----
#if VITCU_DUO_P1_ROUNDS == 2
                constexpr int R0 = (NCH + 1) / 2, R1 = NCH - R0, R2 = 0;
#else
                constexpr int R0 = (NCH + 2) / 3, R1 = (NCH - R0 + 1) / 2, R2 = NCH - R0 - R1;
#endif
                float mx = row_max_part<R0>(lane_addr, min(p.tokens, R0 * 16), -INFINITY);
                mx = row_max_part<R1>(lane_addr + R0 * 16, max(0, min(p.tokens - R0 * 16, R1 * 16)), mx);
                mx = row_max_part<R2>(lane_addr + (R0 + R1) * 16, max(0, min(p.tokens - (R0 + R1) * 16, R2 * 16)), mx);
                nmx = pack2(-mx * sl2, -mx * sl2);
                // ---- pass 2, keys 0..127 ----
                tmem_ld_32x32b_x16(lane_addr, sc[0]);
                exp_chunks<NCH, 0, H>(lane_addr, sc, sl2v, nmx, sum2, p.tokens);
                tmem_st_wait();
            }
            tcgen05_fence_before();
            named_bar_sync(1, kThreadsDuo); // -> control warp issues P V over keys 0..127
            if (NCH > kHalf) {
                // ---- pass 2, remaining keys, under the first P V batch ----
                if (active) {
                    exp_chunks<NCH, H, NCH>(lane_addr, sc, sl2v, nmx, sum2, p.tokens);
                    tmem_st_wait();
                }
                tcgen05_fence_before();
                named_bar_sync(1, kThreadsDuo); // -> P V over the remaining keys
            }
            if (active) {
                float s0, s1;
                unpack2(sum2, s0, s1);
                inv = 1.0f / (s0 + s1);
            }
            // ---- read-out: O leaves tensor memory, the next unit's S may overwrite the buffer ----
            const bool o_ok = mbar_wait_spin_warp(&bars[O_FULL], k & 1, wd, 8);
            tcgen05_fence_after();
            uint32_t vlo[32], vhi[32];
            if (active && o_ok) {
                tmem_ld_32x32b_x32(lane_addr + O_COL, vlo);
                tmem_ld_32x32b_x32(lane_addr + O_COL + 32, vhi);
                tmem_ld_wait();
            }
            tcgen05_fence_before();
            named_bar_sync(1, kThreadsDuo); // -> S of the next unit, which runs under the epilogue below
            if (!o_ok)
                break;
            // ---- epilogue: O / sum -> bf16 -> swizzled tile -> TMA store ----
            if (active) {
                const int img = u.item / p.heads, head = u.item - img * p.heads;
                if (lane == 0)
                    tma_wait_group_read<0>(); // the previous unit's store has read this tile
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    uint32_t w[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int e = 8 * i + 2 * q; // output columns e, e + 1
                        const uint32_t lo = e < 32 ? vlo[e & 31] : vhi[e & 31], hi = e < 32 ? vlo[(e + 1) & 31] : vhi[(e + 1) & 31];
                        w[q] = pack_bf16x2(__uint_as_float(lo) * inv, __uint_as_float(hi) * inv);
                    }
                    sts128(tile + lane * 128 + ((i ^ (lane & 7)) << 4), make_uint4(w[0], w[1], w[2], w[3]));
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(&tmap_out, tile, head * kHeadDim, t * QT + warp * 32, img);
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
    VITCU_REQUIRE(kp <= 208, "the two-CTA attention kernel handles at most 208 tokens (double-buffered V in 113 KB)");
    VITCU_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0, "buffers must be 16-byte aligned");
    CUtensorMap tq, tkv, tout;
    int rc = make_map3(&tq, qkv, batch, tokens, 3 * embed, QT, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (!rc)
        rc = make_map3(&tkv, qkv, batch, tokens, 3 * embed, (uint32_t)kp, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (!rc) // out viewed as [B][T][embed]: box 32 rows x 64 columns (one head), rows past T are clipped
        rc = make_map3(&tout, out, batch, tokens, embed, 32, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc)
        return rc;
    const size_t smem = (size_t)Q_BYTES + 3 * (size_t)kp * 128 + 4 * 4096 + NUM_BARS * 8 + 16 + 1024;
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
    // a handful of images: one query tile per item, so that every tile has a CTA of its own
    // (VITCU_ATTN_UNIT_SPLIT=0: off; read per call, the tests run both forms)
    const int tiles = (tokens + QT - 1) / QT;
    const char *us = getenv("VITCU_ATTN_UNIT_SPLIT");
    p.unit_split = tiles > 1 && p.items * tiles <= 2 * sms && !(us && !strcmp(us, "0"));
    if (p.unit_split)
        p.items *= tiles;
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
    default:
        return set_error(VITCU_E_ARG, __FILE__, __LINE__, "unsupported key count");
    }
#undef VITCU_DUO_CASE
    VITCU_LAUNCHED_KIND(LK_ATTN_DUO);
    return 0;
}

} // namespace vitcu
