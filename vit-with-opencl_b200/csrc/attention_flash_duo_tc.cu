// csrc/attention_flash_duo_tc.cu -- key-blocked ("flash") multi-head attention on the tensor cores for ANY token
// count, two co-resident CTAs per SM: the default for more than 208 tokens (384x384 images -> 577 tokens, which the
// reference's OpenCL kernel cannot run at all: R/multihead.cl:81-83 caps keys at 256).
// BF16 operands, FP32 accumulation and softmax.  Oracle: R/ViT_seq.c:192-262.  R/ = /root/reference/MulticoreMainProject/.
//
// Same idea as the single-block "duo" kernel (attention_duo_tc.cu, profiles/r02_attention.md): a CTA owns 256 TMEM
// columns and < 113 KB of shared memory and the hardware interleaves two such CTAs per SM -- while one is in its
// MUFU-bound exponential pass the other issues MMAs, reads scores or waits for a load.
//
// Work item = (image, head, 128-query tile).  Per item the keys / values stream through TMA rings in blocks of 64:
//     S = Q K_j^T            (tcgen05, M=128, N=64, K=64)   -> one of TWO score buffers (TMEM columns 0..63 / 64..127)
//     softmax                (thread = query row; the 64 scores of the block are read from TMEM ONCE and stay in
//                            registers): p = exp2((s - m) log2e/8) as bf16 pairs back into TMEM over the scores, l += sum p
//     O += P V_j             (tcgen05, A = P from TMEM, B = V_j MN-major, M=128, N=64, K=64) -> TMEM columns 128..191
// O accumulates IN tensor memory over all key blocks of the item: the reference maximum m of a row is only moved (and O and l
// rescaled, a read-modify-write of the row's 64 output columns) when a block's maximum exceeds it by more than 2^8 in the
// exp2 domain -- softmax does not care which reference is subtracted, p <= 256 is harmless in bf16 / fp32, and after the first
// block that practically never happens -- so the softmax warps neither wait for P V nor read O per block.  With two score
// buffers S(j+2) is issued right behind P V(j) (the tensor pipe is in order, so it overwrites buffer j & 1 only after P V(j)
// has read P from it) and S(j+1) is already complete when the softmax warps come back from the barrier: their only
// per-block waits are the TMEM load and one named barrier.
// Finally out = O / l through a swizzled shared-memory tile and a TMA store (rows past T clipped).  Keys past T are
// zero-filled by TMA and masked to p = 0; scores are scaled by 1/sqrt(64) after the dot product (R/ViT_seq.c:211),
// folded into the exp2 argument.
//
//   warps 0-3   softmax + epilogue (warp w owns TMEM lanes 32w..32w+31)
//   warp 4      control: TMA loads (Q double-buffered across items, K ring of 4, V ring of 3 blocks of 64 keys) and MMA
//               issue; it meets the softmax warps at ONE named barrier per block ("P is written").
#include "tc_common.cuh"

using namespace vitcu;
using namespace vitcu::tc;

namespace vitcu {
int device_sm_count(); // gemm_tc.cu
}

namespace {

#ifndef VITCU_CTL_SPIN
#define VITCU_CTL_SPIN 0 // measured: spinning costs 10 % (flash) / 1.5 % (single-block) -- control warp: 1 = spin on its mbarriers (a parked warp wakes up late and everything it issues is on the CTA's chain)
#endif
#if VITCU_CTL_SPIN
#define ctl_wait mbar_wait_spin_warp
#else
#define ctl_wait mbar_wait_warp
#endif

constexpr int kThreadsFD = 160;
constexpr int QT = 128;
constexpr uint32_t Q_BYTES = 128 * 128;     // [128 rows x 64] bf16
constexpr uint32_t O_COL = 128;
constexpr uint32_t TMEM_COLS = 256;
#ifndef VITCU_FD_POLY_MASK
#define VITCU_FD_POLY_MASK 0x00 // pairs (of every 8) whose 2^x runs on the FMA pipe instead of the MUFU
#endif
constexpr uint32_t kFDPolyMask = VITCU_FD_POLY_MASK;
constexpr float kRescaleLog2 = 8.0f; // a row's reference maximum moves only when a block exceeds it by more than 2^8

// Shape of a variant: KB keys per block, NBUF score buffers (NBUF * KB + 64 <= 256 TMEM columns), K / V ring depths, Q buffers.
//   <128, 1, 2, 2, 1>  one 128-column score buffer: S(j+1) is issued behind P V(j) and the softmax warps wait for it (default)
//   <64, 2, 4, 3, 2>   two 64-column score buffers, the block is read from TMEM once: no wait for S, but twice the
//                      per-block fixed cost (barrier, tcgen05.ld / st round trips) -- slower, see profiles/r02_flash_attention.md
template <int KB_, int NBUF_, int KS_, int VS_, int QB_>
struct FDShape {
    static constexpr int KB = KB_, NBUF = NBUF_, KS = KS_, VS = VS_, QB = QB_;
    static constexpr uint32_t KV_BYTES = KB_ * 128; // [KB rows x 64] bf16: one K block or one V block
    // barriers: Q_FULL[QB] per item; K_FULL[KS], V_FULL[VS] per block; S_FULL[NBUF] per score buffer; O_FULL per block
    static constexpr int Q_FULL = 0, K_FULL = QB_, V_FULL = K_FULL + KS_, S_FULL = V_FULL + VS_, O_FULL = S_FULL + NBUF_,
                         NUM_BARS = O_FULL + 1;
    static constexpr size_t SMEM = QB_ * (size_t)Q_BYTES + (KS_ + VS_) * (size_t)KV_BYTES + 4 * 4096 + NUM_BARS * 8 + 16 + 1024;
};

struct FDParams {
    int tokens;
    int qtiles;   // ceil(T / 128)
    int kblocks;  // ceil(T / 64)
    int items;    // batch * heads * qtiles
    int heads, embed;
};

__device__ __forceinline__ float max3f(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <class SH>
__global__ void __launch_bounds__(kThreadsFD, 2)
attention_flash_duo_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                              const __grid_constant__ CUtensorMap tmap_out, const FDParams p, uint32_t *watchdog_flag)
{
    constexpr int KB = SH::KB, NBUF = SH::NBUF, KS = SH::KS, VS = SH::VS, QB = SH::QB;
    constexpr uint32_t KV_BYTES = SH::KV_BYTES;
    constexpr int Q_FULL = SH::Q_FULL, K_FULL = SH::K_FULL, V_FULL = SH::V_FULL, S_FULL = SH::S_FULL, O_FULL = SH::O_FULL,
                  NUM_BARS = SH::NUM_BARS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *sQ = smem;                          // Q tiles [QB]: items alternate
    uint8_t *sK = sQ + QB * Q_BYTES;             // K ring [KS]
    uint8_t *sV = sK + KS * KV_BYTES;            // V ring [VS]
    uint8_t *ostage = sV + VS * KV_BYTES;        // [4 warps][32 rows x 128 B], 128B-swizzled (the rings keep it 1 KB aligned)
    uint64_t *bars = reinterpret_cast<uint64_t *>(ostage + 4 * 4096);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NUM_BARS);
    volatile uint32_t *cta_abort = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = blockIdx.x < (unsigned)p.items ? (p.items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int n_blocks = n_items * p.kblocks; // key blocks this CTA walks over, across its items

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
    pdl_trigger();
    pdl_wait();
    const Watchdog wd{cta_abort, watchdog_flag};

    // item (walked in steps of gridDim.x) -> (image, head, query tile); consecutive items of one (image, head) land on
    // neighbouring CTAs at the same time, so its K / V blocks are read from L2 by all of them
    auto decode = [&](int il, int &img, int &head, int &qt) {
        const int item = blockIdx.x + il * gridDim.x;
        qt = item % p.qtiles;
        const int bh = item / p.qtiles;
        img = bh / p.heads;
        head = bh - img * p.heads;
    };

    if (warp == 4) {
        // ===================== control warp: TMA loads + MMA issue =====================
        // g = block counter of this CTA (item il = g / kblocks, key block j = g % kblocks); score buffer g & 1,
        // K stage g % KS, V stage g % VS, Q buffer il & 1
        const uint32_t idesc_s = umma_idesc_bf16(QT, KB, false, false);
        const uint32_t idesc_o = umma_idesc_bf16(QT, kHeadDim, false, true);
        const uint32_t sq_a = smem_u32(sQ), sk_a = smem_u32(sK), sv_a = smem_u32(sV);
        if (elect_one()) {
            prefetch_tensormap(&tmap_q);
            prefetch_tensormap(&tmap_kv);
        }
        auto load_q = [&](int il) {
            int img, head, qt;
            decode(il, img, head, qt);
            if (elect_one()) {
                mbar_arrive_expect_tx(&bars[Q_FULL + il % QB], Q_BYTES);
                tma_load_3d(sQ + (il % QB) * Q_BYTES, &tmap_q, &bars[Q_FULL + il % QB], head * kHeadDim, qt * QT, img);
            }
            __syncwarp();
        };
        auto load_k = [&](int g) {
            int img, head, qt;
            decode(g / p.kblocks, img, head, qt);
            if (elect_one()) {
                mbar_arrive_expect_tx(&bars[K_FULL + g % KS], KV_BYTES);
                tma_load_3d(sK + (g % KS) * KV_BYTES, &tmap_kv, &bars[K_FULL + g % KS], p.embed + head * kHeadDim, (g % p.kblocks) * KB, img);
            }
            __syncwarp();
        };
        auto load_v = [&](int g) {
            int img, head, qt;
            decode(g / p.kblocks, img, head, qt);
            if (elect_one()) {
                mbar_arrive_expect_tx(&bars[V_FULL + g % VS], KV_BYTES);
                tma_load_3d(sV + (g % VS) * KV_BYTES, &tmap_kv, &bars[V_FULL + g % VS], 2 * p.embed + head * kHeadDim, (g % p.kblocks) * KB, img);
            }
            __syncwarp();
        };
        auto issue_s = [&](int g) { // scores of block g into buffer g % NBUF: needs K(g) and the item's Q tile
            const int il = g / p.kblocks;
            bool ok = ctl_wait(&bars[K_FULL + g % KS], (g / KS) & 1, wd, 3);
            if (ok && g % p.kblocks == 0)
                ok = ctl_wait(&bars[Q_FULL + il % QB], (il / QB) & 1, wd, 4);
            if (!ok)
                return;
            tcgen05_fence_after();
            if (elect_one()) {
                const uint64_t q_desc = umma_desc_k_sw128(sq_a + (il % QB) * Q_BYTES);
                const uint64_t k_desc = umma_desc_k_sw128(sk_a + (g % KS) * KV_BYTES);
#pragma unroll
                for (int k = 0; k < kHeadDim / 16; k++)
                    umma_bf16_ss(tmem_base + (g % NBUF) * KB, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
                umma_commit(&bars[S_FULL + g % NBUF]);
            }
            __syncwarp();
        };
        if (n_blocks > 0) {
            for (int il = 0; il < QB && il < n_items; il++)
                load_q(il);
            for (int g = 0; g < KS && g < n_blocks; g++)
                load_k(g);
            for (int g = 0; g < VS && g < n_blocks; g++)
                load_v(g);
            for (int g = 0; g < NBUF && g < n_blocks; g++)
                issue_s(g);
        }
        for (int g = 0; g < n_blocks; g++) {
            // P V(g-1) retired (a whole block ago): its V stage takes block g - 1 + VS
            if (g > 0) {
                if (!ctl_wait(&bars[O_FULL], (g - 1) & 1, wd, 7))
                    break;
                if (g - 1 + VS < n_blocks)
                    load_v(g - 1 + VS);
            }
            // S(g) retired: its K stage takes block g + KS; at the first block of an item the Q buffer of the item
            // before it is long dead and takes the tile of the item after it
            if (!ctl_wait(&bars[S_FULL + g % NBUF], (g / NBUF) & 1, wd, 5))
                break; // the softmax warps wait on the same barrier and leave with us
            if (g + KS < n_blocks)
                load_k(g + KS);
            // Q: with two buffers the tile of the NEXT item is requested at the first block of an item (the buffer of the
            // item before is long dead); with one buffer after the item's last S has retired
            if (QB == 2 && g % p.kblocks == 0 && g > 0 && g / p.kblocks + 1 < n_items)
                load_q(g / p.kblocks + 1);
            if (QB == 1 && (g + 1) % p.kblocks == 0 && g + 1 < n_blocks)
                load_q((g + 1) / p.kblocks);
            named_bar_sync(1, kThreadsFD); // P(g) is written (and a rescale of O, if any, is complete)
            if (ctl_wait(&bars[V_FULL + g % VS], (g / VS) & 1, wd, 6)) {
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t v_a = sv_a + (g % VS) * KV_BYTES;
#pragma unroll
                    for (int k = 0; k < KB / 16; k++)
                        umma_bf16_ts(tmem_base + O_COL, tmem_base + (g % NBUF) * KB + k * 8, umma_desc_mn_sw128(v_a + k * 2048), idesc_o,
                                     (g % p.kblocks != 0) | (k != 0)); // O accumulates over the item's key blocks
                    umma_commit(&bars[O_FULL]);
                }
                __syncwarp();
                if (g + NBUF < n_blocks)
                    issue_s(g + NBUF); // in order behind P V(g): overwrites buffer g % NBUF once P(g) has been read
            }
            if (*cta_abort)
                break; // the softmax warps leave at their next mbarrier wait, before any further named barrier
        }
    } else {
        // ===================== softmax + epilogue (warps 0-3, thread = query row) =====================
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        const float sl2 = 0.125f * 1.4426950408889634f; // log2(e) / sqrt(64)
        uint8_t *tile = ostage + warp * 4096;
        int g = 0;
        bool alive = true;
        for (int il = 0; il < n_items && alive; il++) {
            int img, head, qt;
            decode(il, img, head, qt);
            // a warp whose 32 query rows all lie past T only keeps the barrier protocol going
            const bool active = qt * QT + warp * 32 < p.tokens;
            float m = -INFINITY, l = 0.f;
            for (int j = 0; j < p.kblocks; j++, g++) {
                if (*cta_abort || !mbar_wait_spin_warp(&bars[S_FULL + g % NBUF], (g / NBUF) & 1, wd, 8)) {
                    alive = false;
                    break;
                }
                tcgen05_fence_after();
                if (active) {
                    const int valid = p.tokens - j * KB; // keys of this block that exist (>= 1)
                    const uint32_t ts = taddr + (g % NBUF) * KB;
                    uint32_t v[2][32];
                    // block maximum: rounds of two 32-column loads with one wait each (KB = 64: the whole block, which
                    // then stays in registers for the exponentials)
                    float bm = -INFINITY;
#pragma unroll
                    for (int r = 0; r < KB / 64; r++) {
                        tmem_ld_32x32b_x32(ts + (2 * r) * 32, v[0]);
                        tmem_ld_32x32b_x32(ts + (2 * r + 1) * 32, v[1]);
                        tmem_ld_wait();
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int c = 2 * r + h;
                            if (c * 32 + 32 <= valid) {
#pragma unroll
                                for (int i = 0; i < 32; i += 2)
                                    bm = max3f(bm, __uint_as_float(v[h][i]), __uint_as_float(v[h][i + 1]));
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; i++)
                                    if (c * 32 + i < valid)
                                        bm = fmaxf(bm, __uint_as_float(v[h][i]));
                            }
                        }
                    }
                    if (j == 0) {
                        m = bm; // the first P V of an item overwrites O: no history to rescale
                    } else {
                        // move the reference maximum only when this block exceeds it by more than 2^kRescaleLog2
                        const bool need = (bm - m) * sl2 > kRescaleLog2;
                        if (__any_sync(0xffffffffu, need)) {
                            // rare: O and l of the rows concerned are rescaled in tensor memory, after P V(g-1) has retired
                            if (!mbar_wait_spin_warp(&bars[O_FULL], (g - 1) & 1, wd, 9)) {
                                alive = false;
                                break;
                            }
                            tcgen05_fence_after();
                            const float a = need ? ex2_approx((m - bm) * sl2) : 1.0f;
#pragma unroll
                            for (int c = 0; c < 2; c++) {
                                uint32_t w[32], lo[16], hi[16];
                                tmem_ld_32x32b_x32(taddr + O_COL + c * 32, w);
                                tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 16; i++) {
                                    lo[i] = __float_as_uint(__uint_as_float(w[i]) * a);
                                    hi[i] = __float_as_uint(__uint_as_float(w[16 + i]) * a);
                                }
                                tmem_st_32x32b_x16(taddr + O_COL + c * 32, lo);
                                tmem_st_32x32b_x16(taddr + O_COL + c * 32 + 16, hi);
                            }
                            l *= a;
                            if (need)
                                m = bm;
                        }
                    }
                    // P = exp2((s - m) * log2e/8) -> bf16 pairs into TMEM over the scores; row sum.  KB = 64: from the
                    // registers of the maximum pass; larger blocks are read from tensor memory a second time
                    const f32x2 sl2v = pack2(sl2, sl2), nmx = pack2(-m * sl2, -m * sl2);
                    f32x2 sum2 = pack2(0.f, 0.f);
#pragma unroll
                    for (int c = 0; c < KB / 32; c++) {
                        if (KB > 64) {
                            tmem_ld_32x32b_x32(ts + c * 32, v[0]);
                            tmem_ld_wait();
                        }
                        const uint32_t(&vc)[32] = v[KB > 64 ? 0 : c & 1];
                        uint32_t packed[16];
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const f32x2 arg = fma2(pack2(__uint_as_float(vc[2 * i]), __uint_as_float(vc[2 * i + 1])), sl2v, nmx);
                            float e0, e1;
                            if (kFDPolyMask & (1u << (i & 7))) { // this pair on the FMA pipe (degree-3 polynomial)
                                exp2_poly2(arg, e0, e1);
                            } else {
                                float a0, a1;
                                unpack2(arg, a0, a1);
                                e0 = ex2_approx(a0);
                                e1 = ex2_approx(a1);
                            }
                            if (c * 32 + 32 > valid) {
                                if (c * 32 + 2 * i >= valid)
                                    e0 = 0.f;
                                if (c * 32 + 2 * i + 1 >= valid)
                                    e1 = 0.f;
                            }
                            sum2 = add2(sum2, pack2(e0, e1));
                            packed[i] = pack_bf16x2(e0, e1);
                        }
                        tmem_st_32x32b_x16(ts + c * 16, packed); // columns [16c, 16c+16) of the buffer are consumed already
                    }
                    tmem_st_wait();
                    float s0, s1;
                    unpack2(sum2, s0, s1);
                    l += s0 + s1;
                }
                tcgen05_fence_before();
                named_bar_sync(1, kThreadsFD); // -> control warp issues P V(g) (accumulating into O) and S(g+2)
            }
            if (!alive)
                break;
            // ---- the item's last P V has retired: O out of tensor memory ----
            if (!mbar_wait_spin_warp(&bars[O_FULL], (g - 1) & 1, wd, 10))
                break;
            tcgen05_fence_after();
            float o[64];
            if (active) {
                uint32_t w[32];
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    tmem_ld_32x32b_x32(taddr + O_COL + c * 32, w);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; i++)
                        o[c * 32 + i] = __uint_as_float(w[i]);
                }
            }
            tcgen05_fence_before(); // the next item's first P V (issued after the next named barrier) overwrites O
            // ---- epilogue: o / l -> bf16 -> swizzled tile -> TMA store (rows past T clipped by the tensor map) ----
            if (active) {
                const float inv = 1.0f / l;
                if (lane == 0)
                    tma_wait_group_read<0>(); // the previous item's store has read this tile
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; i++)
                    sts128(tile + lane * 128 + ((i ^ (lane & 7)) << 4),
                           make_uint4(pack_bf16x2(o[8 * i + 0] * inv, o[8 * i + 1] * inv), pack_bf16x2(o[8 * i + 2] * inv, o[8 * i + 3] * inv),
                                      pack_bf16x2(o[8 * i + 4] * inv, o[8 * i + 5] * inv), pack_bf16x2(o[8 * i + 6] * inv, o[8 * i + 7] * inv)));
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(&tmap_out, tile, head * kHeadDim, qt * QT + warp * 32, img);
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

// 3-D map over a [B][T][width] bf16 tensor: box {64 columns, box_rows, 1}, 128-byte swizzle
int make_map3(CUtensorMap *map, const void *base, int batch, int tokens, int width, uint32_t box_rows, CUtensorMapL2promotion l2)
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
    cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)tokens, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)width * 2, (cuuint64_t)width * 2 * (cuuint64_t)tokens};
    cuuint32_t box[3] = {kHeadDim, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return set_error(VITCU_E_ARG, __FILE__, __LINE__, "cuTensorMapEncodeTiled rejected the attention tensor");
    return 0;
}

template <class SH>
int launch_fd(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t st)
{
    const int embed = heads * kHeadDim;
    CUtensorMap mapq, mapkv, tout;
    int rc = make_map3(&mapq, qkv, batch, tokens, 3 * embed, QT, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (!rc)
        rc = make_map3(&mapkv, qkv, batch, tokens, 3 * embed, SH::KB, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (!rc)
        rc = make_map3(&tout, out, batch, tokens, embed, 32, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc)
        return rc;
    FDParams p;
    p.tokens = tokens;
    p.qtiles = (tokens + QT - 1) / QT;
    p.kblocks = (tokens + SH::KB - 1) / SH::KB;
    p.items = batch * heads * p.qtiles;
    p.heads = heads;
    p.embed = embed;
    static_assert(SH::SMEM <= 113 * 1024 && SH::NBUF * SH::KB + 64 <= (int)TMEM_COLS, "two CTAs per SM");
    static bool configured[64] = {false};
    int dev = 0;
    VITCU_TRY(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        VITCU_TRY(cudaFuncSetAttribute(attention_flash_duo_tc_kernel<SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SH::SMEM));
        VITCU_TRY(cudaFuncSetAttribute(attention_flash_duo_tc_kernel<SH>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared));
        configured[dev] = true;
    }
    const int sms = vitcu::device_sm_count();
    const int grid = p.items < 2 * sms ? p.items : 2 * sms;
    VITCU_TRY(launch_kernel(attention_flash_duo_tc_kernel<SH>, grid, kThreadsFD, SH::SMEM, st, mapq, mapkv, tout, p, watchdog_flag()));
    VITCU_LAUNCHED_KIND(LK_ATTN_FLASH);
    return 0;
}

} // namespace

namespace vitcu {

// qkv [B*T, 3*heads*64] bf16 -> out [B*T, heads*64] bf16; any token count
int attention_bf16_flash_duo_tc(const void *qkv, void *out, int batch, int tokens, int heads, cudaStream_t st)
{
    VITCU_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0, "buffers must be 16-byte aligned");
    // VITCU_FD_SHAPE=64: the two-buffer / 64-key variant (A/B measurements)
    const char *shape = getenv("VITCU_FD_SHAPE");
    if (shape && atoi(shape) == 64)
        return launch_fd<FDShape<64, 2, 4, 3, 2>>(qkv, out, batch, tokens, heads, st);
    return launch_fd<FDShape<128, 1, 2, 2, 1>>(qkv, out, batch, tokens, heads, st);
}

} // namespace vitcu
