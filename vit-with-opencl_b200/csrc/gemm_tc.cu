// csrc/gemm_tc.cu -- BF16 tensor-core GEMM for sm_100a:
//     C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue),  FP32 accumulation.
//
// This is the B200 replacement for the reference's 8x8-tile OpenCL GEMMs
// linear_layer (R/ll.cl:7-70) and QKV (R/multihead.cl:3-63), which carry 96 % of
// the FLOPs of the path; the oracle for the computation is
// linear_layer_seq (R/ViT_seq.c:295-309).  R/ = /root/reference/MulticoreMainProject/.
//
// Structure (one persistent CTA per SM, 320 threads, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor 128x64 A tiles and BNx64 W
//               tiles (both K-major, 128-byte swizzle) into a STAGES-deep ring,
//               completion on mbarriers
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma
//               (cta_group::1, kind::f16, M=128, N=BN, K=16) with the FP32
//               accumulator in TMEM; tcgen05.commit releases ring slots and
//               publishes finished accumulators.  Also owns TMEM alloc/dealloc.
//   warps 2-9   epilogue: tcgen05.ld the accumulator (two warps per 32-lane
//               TMEM quadrant, half of the columns each), fuse bias / GELU /
//               residual / patch-embed remap, store bf16 or fp32 rows.
// The accumulator is double-buffered in TMEM (2 x BN columns), so the epilogue
// of tile i overlaps the MMAs of tile i+1.  Ragged M (197*B is rarely a
// multiple of 128) is handled by TMA zero-fill on load and a row predicate on
// store.  Every mbarrier wait is watchdog-guarded (tc_common.cuh).
#include "tc_common.cuh"
#ifndef VITCU_GELU_FORM
#define VITCU_GELU_FORM 2 // 0 = (3,3) rational, 1 / 2 = MUFU.TANH form scalar / packed (default), 3 = sigmoid form (EX2 + RCP), packed
#endif
#ifndef VITCU_EXIT_WAIT_WRITES
#define VITCU_EXIT_WAIT_WRITES 0 // 1: the single-CTA GEMM waits for its TMA stores to land in global memory before it retires (A/B)
#endif
#ifndef VITCU_BIAS_AHEAD
#define VITCU_BIAS_AHEAD 1 // narrow tiles: bias of a chunk requested a chunk ahead (0: at its use, A/B)
#endif
#ifndef VITCU_GELU_SCALAR
#define VITCU_GELU_SCALAR 0
#endif
#include <stdlib.h>
#include <string.h>

using namespace vitcu;
using namespace vitcu::tc;

namespace {

struct EpiParams {
    int M, N, K;
    size_t ldc;
    int epilogue;
    const float *bias;
    const float *residual;
    const float *pos;
    int patches, tokens;
    int out_bf16;
    int tma_out; // 0: epilogue writes with LSU stores; 1: bf16 tiles by TMA store; 2: fp32 tiles by TMA reduce-add (C += tile); 3: fp32 tiles by TMA store
    int splits;     // split-K factor (1-CTA kernel, TMA reduce-add epilogue only): work item = (tile, K slice)
    int prefetch_w; // 1-CTA kernel, one work item per CTA: prefetch the item's W boxes into the L2 before griddepcontrol.wait
    int exact_gelu; // erff instead of the polynomial (fp32 outputs of the split-bf16 FP32 path)
    // K loop as a list of segments (split-bf16 FP32 path: six piece products over one K range)
    int nseg, seg_kb;        // segments, k-blocks per segment
    int a_seg[6], b_seg[6];  // column offset (elements) of each segment in A and in W
    // LayerNorm folded into the GEMM (BF16 path, see vitcu_gemm_desc): consumer side
    const float2 *ln_stats;  // [ln_slots][M] partial (sum, sum of squares) of the fp32 rows the A operand was cast from
    const float *ln_colsum;  // [N] column sums of the folded weight
    int ln_slots;
    float ln_inv_d;          // 1 / row length
    // ... and producer side (tma_out == 4): the residual epilogue also emits bf16(x_new) and the row partials
    float2 *emit_stats;      // [N / 128][M]
    void *emit_ptr;          // bf16 [M,N] (or e4m3 [M,N]) copy of the updated rows
    // FP8 (E4M3) path: operands are quantised with per-tensor scales, acc_scale = 1 / (scale_A * scale_W) brings the
    // accumulator back; out_scale / emit_scale quantise what this GEMM hands to the next FP8 GEMM
    int kb_elems;            // operand elements per 128-byte k-block row: 64 (bf16) or 128 (e4m3)
    float acc_scale;         // multiplies the accumulator (1 for bf16 operands)
    float out_scale;         // tma_out == 5: y * out_scale -> e4m3
    int emit_fp8;            // emit mode: the second output is e4m3(x * emit_scale) instead of bf16(x)
    float emit_scale;
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 320;
constexpr int kEpiWarps = 8;
constexpr int kStageLd = 36;                // floats per staged epilogue row (32 + 4 pad), LSU path
constexpr int kStageFloats = 2048;          // per epilogue warp: 2 x 4 KB TMA-store buffers (>= 32 * kStageLd floats)

// PIECES = 3: one ring slot holds the three bf16 pieces of BOTH operands for one logical k-block (split-bf16 FP32 path
// at small M), so that the six piece products are issued from tiles loaded once; the epilogue staging shrinks to 4 KB
// per warp (one fp32 tile) to make room for two such slots
template <int BN, int STAGES, int PIECES = 1>
struct SmemLayout {
    static constexpr uint32_t A_BYTES = PIECES * BM * BK * 2;
    static constexpr uint32_t B_BYTES = PIECES * BN * BK * 2;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr uint32_t EPI_WARP_BYTES = PIECES == 1 ? kStageFloats * 4 : 4096;
    static constexpr uint32_t EPI_OFFSET = STAGES * STAGE_BYTES;              // per-warp transpose tiles
    static constexpr uint32_t BAR_OFFSET = EPI_OFFSET + kEpiWarps * EPI_WARP_BYTES;
    static constexpr uint32_t FULLS = STAGES * PIECES; // PIECES = 3: one "landed" barrier per piece pair (a_i, w_i) of a slot
    static constexpr uint32_t NUM_BARS = FULLS + STAGES + 4;
    static constexpr uint32_t TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024; // + alignment slack
    static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

// Epilogue of one warp's share of an accumulator tile: 32 rows (its TMEM lane quadrant) x BN/2
// columns, in 32-column chunks.
//
// tcgen05.ld hands every thread one ROW (32 consecutive columns); storing that straight to global
// memory makes each warp instruction touch 32 different 128-byte lines (32 L1 wavefronts), and that
// LSU traffic competes with the tensor core's operand reads for the L1/shared-memory data pipe
// (profiles/r01_v4_epilogue.md: lsu wavefronts 59 % + tensor 33 % of the pipe).  So every chunk is
// transposed through a per-warp shared-memory tile ([32][36] floats, conflict-free for 128-bit
// accesses both ways) and leaves the SM as fully coalesced 128-bit accesses: 8 lanes per fp32 row
// (4 per bf16 row).  Latency hiding: the tcgen05.ld of chunk c+1 and the residual / position rows
// of chunk c+1 are in flight while chunk c is processed, and the first residual rows are requested
// before the wait for the accumulator.
// Folded LayerNorm, consumer side: mean / rstd of one row from the producer's partial sums, slots in fixed order
// (deterministic).  Single-pass variance like the reference (R/ViT_seq.c:126-135).
struct LnRow {
    float rstd, nrm; // y = rstd * acc + nrm * colsum[n] + bias'[n],  nrm = -rstd * mean
};
constexpr int kMaxLnSlots = 8;
__device__ __forceinline__ void ln_row_fetch(const EpiParams &p, int row, float2 (&part)[kMaxLnSlots])
{
#pragma unroll
    for (int sl = 0; sl < kMaxLnSlots; sl++) {
        part[sl] = make_float2(0.f, 0.f);
        if (sl < p.ln_slots && row < p.M)
            part[sl] = __ldg(p.ln_stats + static_cast<size_t>(sl) * p.M + row);
    }
}
__device__ __forceinline__ LnRow ln_row_finish(const EpiParams &p, const float2 (&part)[kMaxLnSlots])
{
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int sl = 0; sl < kMaxLnSlots; sl++) {
        s1 += part[sl].x;
        s2 += part[sl].y;
    }
    const float mu = s1 * p.ln_inv_d;
    const float var = fmaxf(s2 * p.ln_inv_d - mu * mu, 0.0f);
    LnRow r;
    r.rstd = 1.0f / sqrtf(var + 1e-6f);
    r.nrm = -r.rstd * mu;
    return r;
}

// Per-column epilogue coefficients in shared memory (last 2 KB of a warp's 8 KB staging area, two buffers of 256 floats:
// bias at [0, 128), colsum at [128, 256)): available when the bf16 TMA-store path leaves that part free.
constexpr uint32_t kCoefOffset = 6144;
#ifndef VITCU_STORE_BUFS
#define VITCU_STORE_BUFS 2 // bf16 / e4m3 TMA-store tiles in flight per epilogue warp; 3 was measured and is SLOWER (same box: qkv + LN 130.5 -> 135.7 us, fc1 + LN + GELU 204 -> 211 us): the stores are not what the K = 768 epilogue waits for
#endif
template <bool LN, int STAGE_BYTES_PER_WARP, int NCHUNK>
__device__ __forceinline__ bool coef_in_smem(const EpiParams &p)
{
    return LN || (STAGE_BYTES_PER_WARP >= 8192 && NCHUNK <= 4 && (p.tma_out == 1 || p.tma_out == 5));
}
// asynchronous copy (16 B per lane) of this warp's NCHUNK * 32 coefficients starting at column col_base
template <bool LN, int NCHUNK>
__device__ __forceinline__ void ln_coef_copy(const EpiParams &p, float *dst, int lane, int col_base)
{
    if (lane < NCHUNK * 8) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + 4 * lane)), "l"(p.bias + col_base + 4 * lane) : "memory");
        if (LN)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + 128 + 4 * lane)), "l"(p.ln_colsum + col_base + 4 * lane) : "memory");
    }
}

// ln: statistics of this thread's row for THIS tile (filled by the previous call, or by the caller before the first
// tile); next_row0 >= 0: first row of this warp's share of the NEXT tile -- its partials are requested during the last
// chunk of this tile (when the accumulator prefetch registers are free) and reduced after it, so that the L2 latency
// of those loads never sits between two tiles of an epilogue-bound GEMM.
template <int NCHUNK, int STAGE_BYTES_PER_WARP, bool PREFETCH, bool LN = false>
__device__ __forceinline__ bool epilogue_tile(const EpiParams &p, void *C, const CUtensorMap *tmap_c, uint32_t &chunk_ctr,
                                              float *stage, int lane, int row0, int col_base, uint32_t taddr,
                                              uint64_t *tfull, uint32_t parity, const Watchdog &wd, LnRow &ln, int next_row0,
                                              int next_col_base, uint32_t tile_it, float bias_on = 1.0f)
{
    const bool fp32_add = !p.tma_out && !p.out_bf16 && p.epilogue != VITCU_EPI_BIAS; // residual or position rows to fetch
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
    const float *addsrc = p.epilogue == VITCU_EPI_PATCH_EMBED ? p.pos : p.residual;
    // element offset of (row, first column of this lane) in the output / in the rows to add
    auto out_off = [&](int row) -> size_t {
        if (p.epilogue == VITCU_EPI_PATCH_EMBED) {
            const int img = row / p.patches, pi = row - img * p.patches;
            return (static_cast<size_t>(img) * p.tokens + 1 + pi) * p.ldc + col_base + c4;
        }
        return static_cast<size_t>(row) * p.ldc + col_base + c4;
    };
    auto add_off = [&](int row) -> size_t {
        if (p.epilogue == VITCU_EPI_PATCH_EMBED)
            return static_cast<size_t>(1 + row % p.patches) * p.N + col_base + c4;
        return static_cast<size_t>(row) * p.ldc + col_base + c4;
    };
    float4 add[8];
    auto fetch_add = [&](int chunk) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int row = row0 + i * 4 + rsub;
            add[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (fp32_add && row < p.M)
                add[i] = *reinterpret_cast<const float4 *>(addsrc + add_off(row) + chunk * 32);
        }
    };
    fetch_add(0); // does not depend on the accumulator: in flight during the wait below
    // Narrow tiles (the single-CTA kernel at BN = 128: small M, batch-1 latency, one tile per CTA) whose bias is not
    // staged in shared memory: the bias of a chunk is requested a chunk ahead -- the first one before the wait for the
    // accumulator -- instead of sitting as an L2 round trip between the accumulator and the store
    constexpr bool kBiasAhead = VITCU_BIAS_AHEAD && !LN && NCHUNK <= 2;
    float4 bnext[kBiasAhead ? 8 : 1];
    auto fetch_bias = [&](int chunk) {
        if (kBiasAhead) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                bnext[kBiasAhead ? j : 0] = __ldg(reinterpret_cast<const float4 *>(p.bias + col_base + chunk * 32) + j);
        }
    };
    fetch_bias(0);
    const float rstd = ln.rstd * p.acc_scale, nrm = ln.nrm; // FP8 operands: the de-quantisation rides on the row factor
    float2 part[kMaxLnSlots];
    // Folded LayerNorm: the per-column coefficients of this warp's columns (bias' and colsum, NCHUNK * 32 floats each)
    // sit in shared memory, double-buffered by tile: read through __ldg in the chunk loop they cost an L2 round trip
    // per chunk on the epilogue's critical path (profiles/r02_layernorm_fold.md).  They are copied asynchronously
    // (cp.async, 16 B per lane) one tile ahead into the upper half of the warp's staging area, which the bf16 TMA-store
    // path does not use.
    float *coef = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(stage) + kCoefOffset) + (tile_it & 1) * 256;
    static_assert(!LN || (STAGE_BYTES_PER_WARP >= 8192 && NCHUNK <= 4), "coefficient buffers need the last 2 KB of an 8 KB staging area");
    // the plain bias of the bf16-output GEMMs takes the same route (their FADDs waited on the same loads)
    const bool coef_on = coef_in_smem<LN, STAGE_BYTES_PER_WARP, NCHUNK>(p);
    if (coef_on) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        if (next_col_base >= 0)
            ln_coef_copy<LN, NCHUNK>(p, reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(stage) + kCoefOffset) + ((tile_it + 1) & 1) * 256,
                                     lane, next_col_base);
    }

    bool ok = mbar_wait(tfull, parity, wd, 4);
    ok = __all_sync(0xffffffffu, ok);
    if (!ok)
        return false;
    tcgen05_fence_after();

    // with 8 epilogue warps the accumulator chunks are double-buffered in registers; with 16 warps
    // (96 registers per thread) the other warps hide the tcgen05.ld latency instead
    constexpr int NACC = PREFETCH ? 2 : 1;
    uint32_t acc[NACC][32];
    if (PREFETCH)
        tmem_ld_32x32b_x32(taddr, acc[0]);
#pragma unroll
    for (int c = 0; c < NCHUNK; c++) {
        if (!PREFETCH)
            tmem_ld_32x32b_x32(taddr + c * 32, acc[0]);
        tmem_ld_wait();
        if (PREFETCH && c + 1 < NCHUNK)
            tmem_ld_32x32b_x32(taddr + (c + 1) * 32, acc[(c + 1) % NACC]);
        if (LN && c + 1 == NCHUNK && next_row0 >= 0)
            ln_row_fetch(p, next_row0 + lane, part); // in flight during the last chunk
        if (row0 >= p.M) // warp-uniform: nothing of this warp's rows exists
            continue;
        const int col0 = col_base + c * 32;
        // ---- row-per-thread part: bias (+ GELU), then into the staging tile ----
        float v[32];
        if (LN) {
            const f32x2 r2 = pack2(rstd, rstd), n2 = pack2(nrm, nrm);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b = lds128(coef + c * 32 + j);
                const float4 cs = lds128(coef + 128 + c * 32 + j);
                unpack2(fma2(pack2(__uint_as_float(acc[c % NACC][j + 0]), __uint_as_float(acc[c % NACC][j + 1])), r2,
                             fma2(pack2(cs.x, cs.y), n2, pack2(b.x, b.y))), v[j + 0], v[j + 1]);
                unpack2(fma2(pack2(__uint_as_float(acc[c % NACC][j + 2]), __uint_as_float(acc[c % NACC][j + 3])), r2,
                             fma2(pack2(cs.z, cs.w), n2, pack2(b.z, b.w))), v[j + 2], v[j + 3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b = kBiasAhead ? bnext[kBiasAhead ? j / 4 : 0]
                                            : (coef_on ? lds128(coef + c * 32 + j)
                                                       : __ldg(reinterpret_cast<const float4 *>(p.bias + col0 + j)));
                v[j + 0] = fmaf(b.x, bias_on, __uint_as_float(acc[c % NACC][j + 0]));
                v[j + 1] = fmaf(b.y, bias_on, __uint_as_float(acc[c % NACC][j + 1]));
                v[j + 2] = fmaf(b.z, bias_on, __uint_as_float(acc[c % NACC][j + 2]));
                v[j + 3] = fmaf(b.w, bias_on, __uint_as_float(acc[c % NACC][j + 3]));
            }
            if (kBiasAhead && c + 1 < NCHUNK)
                fetch_bias(c + 1); // in flight during this chunk's GELU / staging / store
        }
        if (p.epilogue == VITCU_EPI_BIAS_GELU) {
            if (p.exact_gelu) {
#pragma unroll
                for (int j = 0; j < 32; j++)
                    v[j] = gelu_erf(v[j]);
            } else {
#if VITCU_GELU_FORM == 3
#pragma unroll
                for (int j = 0; j < 32; j += 2)
                    unpack2(gelu_erf_sigmoid2(pack2(v[j], v[j + 1])), v[j], v[j + 1]);
#elif VITCU_GELU_FORM == 2
#pragma unroll
                for (int j = 0; j < 32; j += 2)
                    unpack2(gelu_erf_tanh2(pack2(v[j], v[j + 1])), v[j], v[j + 1]);
#elif VITCU_GELU_FORM == 1
#pragma unroll
                for (int j = 0; j < 32; j++)
                    v[j] = gelu_erf_tanh1(v[j]);
#elif VITCU_GELU_SCALAR
#pragma unroll
                for (int j = 0; j < 32; j++)
                    v[j] = gelu_erf_fast1(v[j]);
#else
#pragma unroll
                for (int j = 0; j < 32; j += 2)
                    unpack2(gelu_erf_fast2(pack2(v[j], v[j + 1])), v[j], v[j + 1]);
#endif
            }
        }
        if (p.tma_out) {
            // ---- TMA path: the row-per-thread registers go straight into a swizzled shared-memory
            // tile and the TMA engine writes (bf16) or reduce-adds (fp32 residual stream) it to global
            // memory, clipped at M.  No global-memory latency is left on this warp's critical path.
            // staging tiles: 2 KB (bf16) or 4 KB (fp32) each, double-buffered when the warp's share allows
            const uint32_t tile_bytes = p.tma_out == 5 ? 1024u : (p.tma_out == 1 ? 2048u : 4096u);
            // staging tiles per warp: VITCU_STORE_BUFS (2) of the 16-bit / 8-bit tiles in an 8 KB share, whose last 2 KB
            // hold the per-column coefficients; one or two otherwise
            const bool small_tile = p.tma_out == 1 || p.tma_out == 5;
            const int nbuf = (STAGE_BYTES_PER_WARP >= 8192 && small_tile) ? VITCU_STORE_BUFS
                             : ((STAGE_BYTES_PER_WARP >= 8192 || (STAGE_BYTES_PER_WARP >= 4096 && small_tile)) ? 2 : 1);
            uint8_t *buf = reinterpret_cast<uint8_t *>(stage) + (chunk_ctr % (uint32_t)nbuf) * tile_bytes;
            if (lane == 0) { // the store that used this buffer before has finished reading it
                if (nbuf == 3)
                    tma_wait_group_read<2>();
                else if (nbuf == 2)
                    tma_wait_group_read<1>();
                else
                    tma_wait_group_read<0>();
            }
            __syncwarp();
            if (p.tma_out == 5) { // 32 x 32 B rows of e4m3, no swizzle
                const float os = p.out_scale;
#pragma unroll
                for (int q = 0; q < 2; q++)
                    sts128(buf + lane * 32 + (q << 4),
                        make_uint4(pack_e4m3x4(v[16 * q + 0] * os, v[16 * q + 1] * os, v[16 * q + 2] * os, v[16 * q + 3] * os),
                                   pack_e4m3x4(v[16 * q + 4] * os, v[16 * q + 5] * os, v[16 * q + 6] * os, v[16 * q + 7] * os),
                                   pack_e4m3x4(v[16 * q + 8] * os, v[16 * q + 9] * os, v[16 * q + 10] * os, v[16 * q + 11] * os),
                                   pack_e4m3x4(v[16 * q + 12] * os, v[16 * q + 13] * os, v[16 * q + 14] * os, v[16 * q + 15] * os)));
            } else if (p.tma_out == 1) { // 32 x 64 B rows, 64-byte swizzle: 16-byte chunk ^= (row / 2) % 4
#pragma unroll
                for (int q = 0; q < 4; q++)
                    sts128(buf + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4),
                           make_uint4(pack_bf16x2(v[8 * q + 0], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                      pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7])));
            } else { // 32 x 128 B rows, 128-byte swizzle: 16-byte chunk ^= row % 8
#pragma unroll
                for (int q = 0; q < 8; q++)
                    sts128(buf + lane * 128 + ((q ^ (lane & 7)) << 4), make_float4(v[4 * q + 0], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (p.tma_out == 2)
                    tma_reduce_add_2d(tmap_c, buf, col0, row0);
                else
                    tma_store_2d(tmap_c, buf, col0, row0);
                tma_commit_group();
            }
            chunk_ctr++;
            continue;
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
            sts128(stage + lane * kStageLd + 4 * j, make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        __syncwarp();
        // ---- coalesced part ----
        if (p.out_bf16) {
            // 4 lanes x 16 B cover the 64-byte bf16 row segment; 8 rows per instruction
            const int rs = lane >> 2, c8 = (lane & 3) * 8;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int r = i * 8 + rs, row = row0 + r;
                const float4 x = lds128(stage + r * kStageLd + c8);
                const float4 y = lds128(stage + r * kStageLd + c8 + 4);
                if (row < p.M)
                    *reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(C) + static_cast<size_t>(row) * p.ldc + col0 + c8) =
                        make_uint4(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w), pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
            }
        } else {
            // 8 lanes x 16 B cover the 128-byte fp32 row segment; 4 rows per instruction
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int r = i * 4 + rsub, row = row0 + r;
                const float4 x = lds128(stage + r * kStageLd + c4);
                if (row < p.M)
                    *reinterpret_cast<float4 *>(reinterpret_cast<float *>(C) + out_off(row) + c * 32) =
                        make_float4(x.x + add[i].x, x.y + add[i].y, x.z + add[i].z, x.w + add[i].w);
            }
            if (c + 1 < NCHUNK)
                fetch_add(c + 1); // in flight during the next chunk's tcgen05.ld wait, bias and staging
        }
        __syncwarp(); // the staging tile is reused by the next chunk
    }
    if (LN && next_row0 >= 0)
        ln = ln_row_finish(p, part);
    return true;
}

// Producer side of the folded LayerNorm (tma_out == 4; out-proj and fc2 of the BF16 path): the residual update
//     x_new = x_old + acc + bias
// is computed in registers instead of by TMA reduce-add, because the row is needed twice more: as bf16(x_new), the
// A operand of the next GEMM (which applies the LayerNorm in its own epilogue), and in the row's partial sums
// (sum x, sum x^2) over this warp's 128 columns -> emit_stats[col_base / 128][row].  x_old arrives by TMA (a 32 x 32
// fp32 tile per chunk, same tensor map as the store), the thread that owns the row adds its accumulators, puts
// x_new back IN PLACE and the TMA engine stores the tile; a second, bf16 tile goes out beside it.  The layernorm
// kernel's launch and its re-read of the fp32 stream (4 of its 6 bytes per element) disappear.
// Per-warp staging (12 KB): R[2] fp32 in/out tiles at 0 / 4096, B[2] bf16 tiles at 8192 / 10240; rbar[2] signal the loads.
#ifndef VITCU_EMIT_VARIANT
#define VITCU_EMIT_VARIANT 0 // 0: residual tiles and outputs through the TMA engine (default); 4: cp.async + coalesced stores (A/B, not faster)
#endif
template <int NCHUNK>
__device__ __forceinline__ bool epilogue_tile_emit(const EpiParams &p, const CUtensorMap *tmap_c, const CUtensorMap *tmap_d,
                                                   uint8_t *stage, uint64_t *rbar, uint32_t (&rphase)[2], int lane, int row0,
                                                   int col_base, uint32_t taddr, uint64_t *tfull, uint32_t parity,
                                                   const Watchdog &wd)
{
    const bool rows_exist = row0 < p.M; // warp-uniform
    if (rows_exist && lane == 0) {
        tma_wait_group_read<0>(); // this warp's stores of the previous tile have read the staging buffers
        for (int c = 0; c < 2 && c < NCHUNK; c++) {
            mbar_arrive_expect_tx(&rbar[c], 4096);
            tma_load_2d(stage + c * 4096, tmap_c, &rbar[c], col_base + c * 32, row0);
        }
    }
    __syncwarp();
    bool ok = mbar_wait(tfull, parity, wd, 4);
    ok = __all_sync(0xffffffffu, ok);
    if (!ok)
        return false;
    tcgen05_fence_after();
    uint32_t acc[2][32];
    tmem_ld_32x32b_x32(taddr, acc[0]);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCHUNK; c++) {
        tmem_ld_wait();
        if (c + 1 < NCHUNK)
            tmem_ld_32x32b_x32(taddr + (c + 1) * 32, acc[(c + 1) & 1]);
        if (!rows_exist)
            continue;
        const int col0 = col_base + c * 32, b = c & 1;
        uint8_t *R = stage + b * 4096, *B = stage + 8192 + b * 2048;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4 *>(p.bias + col0 + j));
            v[j + 0] = fmaf(__uint_as_float(acc[c & 1][j + 0]), p.acc_scale, bb.x);
            v[j + 1] = fmaf(__uint_as_float(acc[c & 1][j + 1]), p.acc_scale, bb.y);
            v[j + 2] = fmaf(__uint_as_float(acc[c & 1][j + 2]), p.acc_scale, bb.z);
            v[j + 3] = fmaf(__uint_as_float(acc[c & 1][j + 3]), p.acc_scale, bb.w);
        }
        // the residual tile of this chunk has landed
        ok = mbar_wait(&rbar[b], rphase[b], wd, 9);
        ok = __all_sync(0xffffffffu, ok);
        if (!ok)
            return false;
        rphase[b] ^= 1;
#pragma unroll
        for (int q = 0; q < 8; q++) { // 32 x 128 B rows, 128-byte swizzle: 16-byte chunk ^= row % 8
            uint8_t *slot = R + lane * 128 + ((q ^ (lane & 7)) << 4);
            const float4 xo = lds128(slot);
            const float a0 = xo.x + v[4 * q + 0], a1 = xo.y + v[4 * q + 1], a2 = xo.z + v[4 * q + 2], a3 = xo.w + v[4 * q + 3];
            v[4 * q + 0] = a0;
            v[4 * q + 1] = a1;
            v[4 * q + 2] = a2;
            v[4 * q + 3] = a3;
            sts128(slot, make_float4(a0, a1, a2, a3));
            s1 += (a0 + a1) + (a2 + a3);
            s2 = fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, fmaf(a3, a3, s2))));
        }
        if (p.emit_fp8) { // 32 x 32 B rows of e4m3(x * emit_scale), no swizzle: the A operand of an FP8 GEMM
            const float es = p.emit_scale;
#pragma unroll
            for (int q = 0; q < 2; q++)
                sts128(B + lane * 32 + (q << 4),
                       make_uint4(pack_e4m3x4(v[16 * q + 0] * es, v[16 * q + 1] * es, v[16 * q + 2] * es, v[16 * q + 3] * es),
                                  pack_e4m3x4(v[16 * q + 4] * es, v[16 * q + 5] * es, v[16 * q + 6] * es, v[16 * q + 7] * es),
                                  pack_e4m3x4(v[16 * q + 8] * es, v[16 * q + 9] * es, v[16 * q + 10] * es, v[16 * q + 11] * es),
                                  pack_e4m3x4(v[16 * q + 12] * es, v[16 * q + 13] * es, v[16 * q + 14] * es, v[16 * q + 15] * es)));
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) // 32 x 64 B rows, 64-byte swizzle: 16-byte chunk ^= (row / 2) % 4
                sts128(B + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4),
                       make_uint4(pack_bf16x2(v[8 * q + 0], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                  pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7])));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(tmap_c, R, col0, row0);
            tma_store_2d(tmap_d, B, col0, row0);
            tma_commit_group();
            if (c + 2 < NCHUNK) { // the buffers are reloaded once the stores just issued have read them
                tma_wait_group_read<0>();
                mbar_arrive_expect_tx(&rbar[b], 4096);
                tma_load_2d(R, tmap_c, &rbar[b], col_base + (c + 2) * 32, row0);
            }
        }
        __syncwarp();
    }
    if (rows_exist && row0 + lane < p.M)
        p.emit_stats[static_cast<size_t>(col_base / (32 * NCHUNK)) * p.M + row0 + lane] = make_float2(s1, s2);
    return true;
}

// The same producer epilogue without the TMA engine (VITCU_EMIT_VARIANT 4): the epilogue's TMA loads and the read-waits
// of its TMA stores queue behind the main loop's operand loads (profiles/r02_layernorm_fold.md), so here the residual tile
// arrives by cp.async (16 B per lane, 8 lanes per 128-byte row, written into the same swizzled layout) two chunks ahead
// into a ring of three 4 KB buffers, and both outputs leave through coalesced 128-bit stores of the warp itself: x_new goes
// back into its buffer, is read out with the 8-lanes-per-row mapping and stored; the bf16 / e4m3 rows take the same route
// through the same buffer.  Every read of a buffer is synchronous, so it can be refilled after a __syncwarp.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
template <int NCHUNK>
__device__ __forceinline__ bool epilogue_tile_emit_lsu(const EpiParams &p, uint8_t *stage, int lane, int row0, int col_base,
                                                       uint32_t taddr, uint64_t *tfull, uint32_t parity, const Watchdog &wd)
{
    const bool rows_exist = row0 < p.M; // warp-uniform
    const int rs = lane >> 3, q8 = lane & 7; // fp32 tiles: 8 lanes per 128-byte row, 4 rows per instruction
    float *xg = const_cast<float *>(p.residual);
    auto load_chunk = [&](int c) {
        uint8_t *R = stage + (c % 3) * 4096;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = 4 * i + rs;
            if (row0 + r < p.M)
                cp_async16(R + r * 128 + ((q8 ^ (r & 7)) << 4), xg + static_cast<size_t>(row0 + r) * p.ldc + col_base + c * 32 + 4 * q8);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (rows_exist) {
        load_chunk(0);
        if (NCHUNK > 1)
            load_chunk(1);
    }
    bool ok = mbar_wait(tfull, parity, wd, 4);
    ok = __all_sync(0xffffffffu, ok);
    if (!ok)
        return false;
    tcgen05_fence_after();
    uint32_t acc[2][32];
    tmem_ld_32x32b_x32(taddr, acc[0]);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCHUNK; c++) {
        tmem_ld_wait();
        if (c + 1 < NCHUNK)
            tmem_ld_32x32b_x32(taddr + (c + 1) * 32, acc[(c + 1) & 1]);
        if (!rows_exist)
            continue;
        const int col0 = col_base + c * 32;
        uint8_t *R = stage + (c % 3) * 4096;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4 *>(p.bias + col0 + j));
            v[j + 0] = fmaf(__uint_as_float(acc[c & 1][j + 0]), p.acc_scale, bb.x);
            v[j + 1] = fmaf(__uint_as_float(acc[c & 1][j + 1]), p.acc_scale, bb.y);
            v[j + 2] = fmaf(__uint_as_float(acc[c & 1][j + 2]), p.acc_scale, bb.z);
            v[j + 3] = fmaf(__uint_as_float(acc[c & 1][j + 3]), p.acc_scale, bb.w);
        }
        if (c + 2 < NCHUNK)
            load_chunk(c + 2); // its buffer was last used by chunk c - 1, read out synchronously
        // chunk c has landed: at most the groups of chunks c + 1, c + 2 may still be in flight
        if (c + 2 < NCHUNK)
            cp_async_wait<2>();
        else if (c + 1 < NCHUNK)
            cp_async_wait<1>();
        else
            cp_async_wait<0>();
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; q++) { // thread = row: 32 x 128 B rows, 128-byte swizzle: 16-byte chunk ^= row % 8
            uint8_t *slot = R + lane * 128 + ((q ^ (lane & 7)) << 4);
            const float4 xo = lds128(slot);
            const float a0 = xo.x + v[4 * q + 0], a1 = xo.y + v[4 * q + 1], a2 = xo.z + v[4 * q + 2], a3 = xo.w + v[4 * q + 3];
            v[4 * q + 0] = a0;
            v[4 * q + 1] = a1;
            v[4 * q + 2] = a2;
            v[4 * q + 3] = a3;
            sts128(slot, make_float4(a0, a1, a2, a3));
            s1 += (a0 + a1) + (a2 + a3);
            s2 = fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, fmaf(a3, a3, s2))));
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; i++) { // x_new out: 8 lanes per row, 4 rows (4 x 128 contiguous bytes) per instruction
            const int r = 4 * i + rs;
            const float4 xn = lds128(R + r * 128 + ((q8 ^ (r & 7)) << 4));
            if (row0 + r < p.M)
                *reinterpret_cast<float4 *>(xg + static_cast<size_t>(row0 + r) * p.ldc + col0 + 4 * q8) = xn;
        }
        __syncwarp(); // the fp32 tile has been read: its buffer takes the 16-bit / 8-bit copy
        if (p.emit_fp8) { // 32 x 32 B rows of e4m3(x * emit_scale)
            const float es = p.emit_scale;
#pragma unroll
            for (int q = 0; q < 2; q++)
                sts128(R + lane * 32 + (q << 4),
                       make_uint4(pack_e4m3x4(v[16 * q + 0] * es, v[16 * q + 1] * es, v[16 * q + 2] * es, v[16 * q + 3] * es),
                                  pack_e4m3x4(v[16 * q + 4] * es, v[16 * q + 5] * es, v[16 * q + 6] * es, v[16 * q + 7] * es),
                                  pack_e4m3x4(v[16 * q + 8] * es, v[16 * q + 9] * es, v[16 * q + 10] * es, v[16 * q + 11] * es),
                                  pack_e4m3x4(v[16 * q + 12] * es, v[16 * q + 13] * es, v[16 * q + 14] * es, v[16 * q + 15] * es)));
            __syncwarp();
            uint8_t *og = reinterpret_cast<uint8_t *>(p.emit_ptr);
#pragma unroll
            for (int i = 0; i < 2; i++) { // 2 lanes per 32-byte row, 16 rows per instruction
                const int r = 16 * i + (lane >> 1);
                const float4 t = lds128(R + r * 32 + ((lane & 1) << 4));
                if (row0 + r < p.M)
                    *reinterpret_cast<float4 *>(og + static_cast<size_t>(row0 + r) * p.N + col0 + 16 * (lane & 1)) = t;
            }
        } else { // 32 x 64 B rows of bf16, 64-byte swizzle: 16-byte chunk ^= (row / 2) % 4
#pragma unroll
            for (int q = 0; q < 4; q++)
                sts128(R + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4),
                       make_uint4(pack_bf16x2(v[8 * q + 0], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                  pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7])));
            __syncwarp();
            __nv_bfloat16 *og = reinterpret_cast<__nv_bfloat16 *>(p.emit_ptr);
#pragma unroll
            for (int i = 0; i < 4; i++) { // 4 lanes per 64-byte row, 8 rows per instruction
                const int r = 8 * i + (lane >> 2), q4 = lane & 3;
                const float4 t = lds128(R + r * 64 + ((q4 ^ ((r >> 1) & 3)) << 4));
                if (row0 + r < p.M)
                    *reinterpret_cast<float4 *>(og + static_cast<size_t>(row0 + r) * p.N + col0 + 8 * q4) = t;
            }
        }
        __syncwarp(); // the buffer may be refilled (by the cp.async of chunk c + 3, issued one iteration later)
    }
    if (rows_exist && row0 + lane < p.M)
        p.emit_stats[static_cast<size_t>(col_base / (32 * NCHUNK)) * p.M + row0 + lane] = make_float2(s1, s2);
    return true;
}

template <int BN, int STAGES, bool LN = false, int PIECES = 1>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, void *C, const EpiParams p, uint32_t *watchdog_flag)
{
    using L = SmemLayout<BN, STAGES, PIECES>;
    static_assert(PIECES == 1 || (PIECES == 3 && !LN), "PIECES is 1 or 3 (split-bf16 operands)");
    static_assert(BN % 64 == 0 && BN <= 256, "BN must be 64..256 in steps of 64");
    constexpr uint32_t TMEM_COLS = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
    constexpr uint32_t IDESC = umma_idesc_bf16(BM, BN, false, false);

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L::BAR_OFFSET);
    uint64_t *empty_bar = full_bar + L::FULLS;
    uint64_t *tfull_bar = empty_bar + STAGES;
    uint64_t *tempty_bar = tfull_bar + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + 2);
    volatile uint32_t *cta_abort = tmem_slot + 1;

    const int warp = threadIdx.x >> 5; // warp-uniform
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < (int)L::FULLS; i++)
            mbar_init(&full_bar[i], 1);
        for (int i = 0; i < STAGES; i++)
            mbar_init(&empty_bar[i], 1);
        for (int i = 0; i < 2; i++) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], kEpiWarps);
        }
        *cta_abort = 0;
        fence_barrier_init();
    }
    if (warp == 1)
        tmem_alloc(tmem_slot, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int num_m = (p.M + BM - 1) / BM, num_n = p.N / BN;
    // work item = (output tile, K slice): with splits > 1 every slice reduce-adds its partial tile
    // into C through the TMA engine (only slice 0 carries the bias), so small-M GEMMs fill the SMs
    // (PIECES = 3: a k-block is a LOGICAL one -- all six piece products of 64 columns of K)
    const int total_kb = PIECES == 3 ? p.seg_kb : p.nseg * p.seg_kb, kb_per = (total_kb + p.splits - 1) / p.splits;
    const int num_tiles = num_m * num_n * p.splits;
    // every CTA holds its TMEM now: the next kernel may start its prologue; our own global-memory
    // traffic (TMA loads, epilogue stores) waits for the previous kernel to complete
    pdl_trigger();
    // ... except for this: with one work item per CTA (batch-1 latency) the W tiles of the item come straight from
    // DRAM (the weights of a forward do not fit the L2), and the wait below is idle time -- the CTA is resident while
    // the kernel in front of it (a LayerNorm, the attention) still runs.  W does not depend on that kernel, so its
    // boxes are prefetched into the L2 now; A, which does, is only requested after the wait.  An L2 prefetch cannot
    // return stale data (the L2 is the coherence point), so this is safe even when W was written by the preceding kernel.
    if (p.prefetch_w && warp == 0 && (int)blockIdx.x < num_tiles && elect_one()) {
        prefetch_tensormap(&tmap_b);
        const int tile = blockIdx.x, split = tile % p.splits, n_blk = (tile / p.splits) % num_n;
        const int kb_end = min(total_kb, (split + 1) * kb_per);
        for (int kb = split * kb_per; kb < kb_end; kb++) {
            const int seg = kb / p.seg_kb, kk = (kb - seg * p.seg_kb) * BK;
            tma_prefetch_l2_2d(&tmap_b, p.b_seg[seg] + kk, n_blk * BN);
        }
    }
    pdl_wait();
    const Watchdog wd{cta_abort, watchdog_flag};


    if (warp == 0) {
        // ===================== TMA producer =====================
        // The whole warp runs the loop and one elected lane issues: with warp-uniform
        // control flow the descriptors/addresses stay in uniform registers (a lane-0-only
        // branch made ptxas emit ELECT + R2UR.BROADCAST chains per instruction, which made
        // the single issuing thread the bottleneck -- profiles/r01_v4_issue_loop.md).
        if (elect_one()) {
            prefetch_tensormap(&tmap_a);
            prefetch_tensormap(&tmap_b);
        }
        uint32_t stage = 0, phase = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < num_tiles && ok; tile += gridDim.x) {
            const int split = tile % p.splits, t2 = tile / p.splits;
            const int m_blk = t2 / num_n, n_blk = t2 - m_blk * num_n;
            const int kb_end = min(total_kb, (split + 1) * kb_per);
            for (int kb = split * kb_per; kb < kb_end; kb++) {
                if (!(ok = mbar_wait_warp(&empty_bar[stage], phase ^ 1, wd, 1)))
                    break;
                if (elect_one()) {
                    uint8_t *sa = smem + stage * L::STAGE_BYTES;
                    if (PIECES == 3) {
                        // piece i of an operand starts at column i * K of its [rows, 3K] matrix; the pair (a_i, w_i)
                        // has its own barrier, so the issuer starts on a1 w1 when a third of the slot has landed
#pragma unroll
                        for (int i = 0; i < 3; i++) {
                            uint64_t *bar = &full_bar[stage * 3 + i];
                            mbar_arrive_expect_tx(bar, L::STAGE_BYTES / 3);
                            tma_load_2d(sa + i * (L::A_BYTES / 3), &tmap_a, bar, i * p.K + kb * BK, m_blk * BM);
                            tma_load_2d(sa + L::A_BYTES + i * (L::B_BYTES / 3), &tmap_b, bar, i * p.K + kb * BK, n_blk * BN);
                        }
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                        const int seg = kb / p.seg_kb, kk = (kb - seg * p.seg_kb) * BK;
                        tma_load_2d(sa, &tmap_a, &full_bar[stage], p.a_seg[seg] + kk, m_blk * BM);
                        tma_load_2d(sa + L::A_BYTES, &tmap_b, &full_bar[stage], p.b_seg[seg] + kk, n_blk * BN);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        uint32_t stage = 0, phase = 0, it = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < num_tiles && ok; tile += gridDim.x, it++) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            if (!(ok = mbar_wait_warp(&tempty_bar[acc], acc_phase ^ 1, wd, 2)))
                break;
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            const int split = tile % p.splits;
            const int kb_begin = split * kb_per, kb_end = min(total_kb, (split + 1) * kb_per);
            for (int kb = kb_begin; kb < kb_end; kb++) {
                const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
                if (PIECES == 3) {
                    // the six significant piece products of this logical k-block from tiles that were loaded once,
                    // in the order their operands land: a1 w1 | a1 w2, a2 w1, a2 w2 | a1 w3, a3 w1
                    constexpr int ai[6] = {0, 0, 1, 1, 0, 2}, bj[6] = {0, 1, 0, 1, 2, 0}, first_of[3] = {0, 1, 4}, end_of[3] = {1, 4, 6};
#pragma unroll
                    for (int g = 0; g < 3 && ok; g++) {
                        if (!(ok = mbar_wait_warp(&full_bar[stage * 3 + g], phase, wd, 3)))
                            break;
                        tcgen05_fence_after();
                        if (elect_one()) {
#pragma unroll
                            for (int c = first_of[g]; c < end_of[g]; c++) {
                                const uint64_t a_desc = umma_desc_k_sw128(sa + ai[c] * (L::A_BYTES / 3));
                                const uint64_t b_desc = umma_desc_k_sw128(sa + L::A_BYTES + bj[c] * (L::B_BYTES / 3));
#pragma unroll
                                for (int k = 0; k < BK / 16; k++)
                                    umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, IDESC, (kb != kb_begin) | (c != 0) | (k != 0));
                            }
                        }
                        __syncwarp();
                    }
                    if (!ok)
                        break;
                } else {
                    if (!(ok = mbar_wait_warp(&full_bar[stage], phase, wd, 3)))
                        break;
                    tcgen05_fence_after();
                }
                if (elect_one()) {
                    if (PIECES == 1) {
                        const uint64_t a_desc = umma_desc_k_sw128(sa);
                        const uint64_t b_desc = umma_desc_k_sw128(sa + L::A_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; k++) // +32 bytes per K=16 step inside the 128B swizzle row
                            umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, IDESC, (kb != kb_begin) | (k != 0));
                    }
                    umma_commit(&empty_bar[stage]); // ring slot reusable once these MMAs retire
                    if (kb == kb_end - 1)
                        umma_commit(&tfull_bar[acc]); // accumulator complete
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int quad = warp & 3;          // TMEM lanes [32*quad, 32*quad+32) are this warp's
        const int half = (warp - 2) >> 2;   // which half of the BN columns
        uint32_t it = 0, chunk_ctr = 0;
        if (warp == 2 && lane == 0)
            prefetch_tensormap(&tmap_c);
        LnRow ln = {1.0f, 0.0f};
        if (LN && (int)blockIdx.x < num_tiles) { // first tile: fetched here; later tiles one tile ahead
            float2 part[kMaxLnSlots];
            ln_row_fetch(p, ((int)blockIdx.x / p.splits / num_n) * BM + quad * 32 + lane, part);
            ln = ln_row_finish(p, part);
        }
        if (coef_in_smem<LN, L::EPI_WARP_BYTES, BN / 64>(p) && (int)blockIdx.x < num_tiles)
            ln_coef_copy<LN, BN / 64>(p, reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(smem + L::EPI_OFFSET) + (warp - 2) * L::EPI_WARP_BYTES + kCoefOffset),
                                      lane, (((int)blockIdx.x / p.splits) % num_n) * BN + half * (BN / 2));
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
            const int split = tile % p.splits, t2 = tile / p.splits;
            const int m_blk = t2 / num_n, n_blk = t2 - m_blk * num_n;
            const int tile_n = tile + (int)gridDim.x;
            const int next_row0 = tile_n < num_tiles ? (tile_n / p.splits / num_n) * BM + quad * 32 : -1;
            const int next_col = tile_n < num_tiles ? ((tile_n / p.splits) % num_n) * BN + half * (BN / 2) : -1;
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            float *stage_tile = reinterpret_cast<float *>(smem + L::EPI_OFFSET + (warp - 2) * L::EPI_WARP_BYTES);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + half * (BN / 2);
            if (!epilogue_tile<BN / 64, L::EPI_WARP_BYTES, true, LN>(p, C, &tmap_c, chunk_ctr, stage_tile, lane, m_blk * BM + quad * 32,
                                   n_blk * BN + half * (BN / 2), taddr, &tfull_bar[acc], acc_phase, wd, ln, next_row0, next_col, it,
                                   split == 0 ? 1.0f : 0.0f))
                break;
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&tempty_bar[acc]);
        }
        // the staging buffers must outlive the TMA engine's reads of them; the writes themselves are complete and
        // visible once the grid has completed, which is what the next kernel's griddepcontrol.wait (or the stream)
        // orders on -- waiting here for the global-memory side as well costs a round trip per launch, which
        // matters only where launches are a few microseconds long (batch-1 latency)
        if (lane == 0) {
#if VITCU_EXIT_WAIT_WRITES
            tma_wait_group<0>();
#else
            tma_wait_group_read<0>();
#endif
        }
    }

    // ===================== teardown =====================
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs on the two SMs of a TPC own one
// 256x256 output tile.  Each CTA stages its own 128 A rows and HALF of the W
// tile (128 rows), so a k-block costs 32 KB of shared-memory writes + 32 KB of
// reads per SM instead of 48 + 48 -- the 1-CTA kernel is shared-memory-port
// bound at ~62 % tensor-pipe utilisation (profiles/r01_v1_summary.md).  The
// leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2 with M = 256; the
// accumulator rows of each CTA live in that CTA's TMEM, and each CTA runs its
// own epilogue.  Barrier topology:
//   full[s]    leader's barrier, 2 arrivals (each CTA's producer: expect_tx of its
//              own bytes) + the TMA bytes of both CTAs
//   empty[s]   one per CTA, released by the leader's multicast tcgen05.commit
//   tfull[a]   one per CTA, multicast commit when the accumulator is complete
//   tempty[a]  leader's barrier, 16 arrivals (8 epilogue warps of each CTA)
// ---------------------------------------------------------------------------
template <int STAGES, uint32_t EPI_BYTES = 65536>
struct SmemLayout2 {
    static constexpr uint32_t A_BYTES = BM * BK * 2;       // this CTA's 128 A rows
    static constexpr uint32_t B_BYTES = 128 * BK * 2;      // this CTA's half of the 256-row W tile
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr uint32_t EPI_OFFSET = STAGES * STAGE_BYTES;
    static constexpr uint32_t BAR_OFFSET = EPI_OFFSET + EPI_BYTES; // epilogue staging, split evenly over the warps
    static constexpr uint32_t NUM_BARS = 2 * STAGES + 4 + 2 * 16; // + two residual-tile barriers per epilogue warp (emit mode)
    static constexpr uint32_t TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024;
    static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

// EW = epilogue warps per CTA: 8 (two per TMEM lane quadrant, 128 columns each, 168 registers) or
// 16 (four per quadrant, 64 columns each, 96 registers): the heavier epilogues (GELU) are bound by
// the latency of one warp's chunk chain, which more warps overlap
// EPI_BYTES = epilogue staging per CTA: 64 KB with 5 operand stages, or 32 KB (4 KB per warp: two bf16 tiles
// or one fp32 tile, TMA output path only) which makes room for a sixth stage
// EMIT = residual epilogue that also emits bf16(x) and the row partial sums for the folded LayerNorm (epilogue_tile_emit)
// FP8 = E4M3 operands (tcgen05.mma kind::f8f6f4): the same byte layout in shared memory, 128 elements per k-block row
template <int STAGES, int EW, uint32_t EPI_BYTES = 65536, bool EMIT = false, bool LN = false, bool FP8 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
gemm_bf16_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_d, void *C,
                     const EpiParams p, uint32_t *watchdog_flag)
{
    using L = SmemLayout2<STAGES, EPI_BYTES>;
    constexpr int BN = 256, BM2 = 256;
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t IDESC = FP8 ? umma_idesc_e4m3(BM2, BN) : umma_idesc_bf16(BM2, BN, false, false);

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L::BAR_OFFSET);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tfull_bar = empty_bar + STAGES;
    uint64_t *tempty_bar = tfull_bar + 2;
    uint64_t *res_bar = tempty_bar + 2; // [EW][2] residual-tile barriers (emit mode)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(res_bar + 2 * 16);
    volatile uint32_t *cta_abort = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; i++) {
            mbar_init(&full_bar[i], 2);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 2 * EW);
        }
        for (int i = 0; i < 2 * 16; i++)
            mbar_init(&res_bar[i], 1);
        *cta_abort = 0;
        fence_barrier_init();
    }
    if (warp == 1)
        tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tcgen05_fence_before();
    cluster_sync_all(); // both CTAs' barriers are initialised before any remote arrive / TMA signal
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // every CTA holds its TMEM now: the next kernel may start its prologue; our own global-memory
    // traffic (TMA loads, epilogue stores) waits for the previous kernel to complete
    pdl_trigger();
    pdl_wait();
    const Watchdog wd{cta_abort, watchdog_flag};

    const int num_m = (p.M + BM2 - 1) / BM2, num_n = p.N / BN;
    const int num_kb = p.nseg * p.seg_kb;
    // work item = (output tile, K slice).  splits > 1 only with the TMA reduce-add epilogue at small M (batch-1
    // latency): every slice adds its partial tile into C, slice 0 carries the bias, and a [197 x 768] x [768 x 2304]
    // product occupies 72 CTA pairs instead of 9.  Against the single-CTA kernel's 128 x 128 items the pair moves
    // 57 KB instead of 98 KB from L2 per 256 x 256 x 64 block, and those launches are bound by L2 -> SM traffic.
    const int kb_per = (num_kb + p.splits - 1) / p.splits;
    const int num_tiles = num_m * num_n * p.splits;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (elect_one()) {
            prefetch_tensormap(&tmap_a);
            prefetch_tensormap(&tmap_b);
        }
        uint32_t stage = 0, phase = 0;
        bool ok = true;
        // the leader's full barriers as shared::cluster addresses (consecutive 8-byte slots)
        const uint32_t full_leader0 = mapa_u32(smem_u32(&full_bar[0]), 0);
        for (int tile = pair; tile < num_tiles && ok; tile += num_pairs) {
            const int t2 = tile / p.splits, split = tile - t2 * p.splits;
            const int m_blk = t2 / num_n, n_blk = t2 - m_blk * num_n;
            const int arow = m_blk * BM2 + (int)rank * BM, brow = n_blk * BN + (int)rank * 128;
            const int kb_begin = split * kb_per, kb_end = min(num_kb, kb_begin + kb_per);
            // running (segment, column) instead of a division per k-block: this warp's instruction
            // stream is what feeds the tensor core (the MMA issuer spends most of its time on the full barrier)
            int seg = kb_begin / p.seg_kb, kin = kb_begin - seg * p.seg_kb;
            int acol = p.a_seg[seg] + kin * p.kb_elems, bcol = p.b_seg[seg] + kin * p.kb_elems;
            for (int kb = kb_begin; kb < kb_end; kb++) {
                if (!(ok = mbar_wait_warp(&empty_bar[stage], phase ^ 1, wd, 1)))
                    break;
                if (elect_one()) {
                    uint8_t *sa = smem + stage * L::STAGE_BYTES;
                    const uint32_t full_leader = full_leader0 + stage * 8;
                    mbar_arrive_expect_tx_cluster(full_leader, L::STAGE_BYTES);
                    tma_load_2d_2sm(sa, &tmap_a, full_leader, acol, arow);
                    tma_load_2d_2sm(sa + L::A_BYTES, &tmap_b, full_leader, bcol, brow);
                }
                __syncwarp();
                acol += p.kb_elems;
                bcol += p.kb_elems;
                if (++kin == p.seg_kb && kb + 1 < kb_end) {
                    kin = 0;
                    seg++;
                    acol = p.a_seg[seg];
                    bcol = p.b_seg[seg];
                }
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            uint32_t stage = 0, phase = 0, it = 0;
            bool ok = true;
            for (int tile = pair; tile < num_tiles && ok; tile += num_pairs, it++) {
                const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                if (!(ok = mbar_wait_warp(&tempty_bar[acc], acc_phase ^ 1, wd, 2)))
                    break;
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                const int split = tile % p.splits;
                const int kb_begin = split * kb_per, kb_end = min(num_kb, kb_begin + kb_per);
                for (int kb = kb_begin; kb < kb_end; kb++) {
                    if (!(ok = mbar_wait_warp(&full_bar[stage], phase, wd, 3)))
                        break;
                    tcgen05_fence_after();
                    if (elect_one()) {
                        const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
                        const uint64_t a_desc = umma_desc_k_sw128(sa);
                        const uint64_t b_desc = umma_desc_k_sw128(sa + L::A_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; k++) { // four instructions per k-block, 32 bytes of K each
                            if (FP8)
                                umma_e4m3_ss_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, IDESC, ((kb - kb_begin) | k) != 0);
                            else
                                umma_bf16_ss_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, IDESC, ((kb - kb_begin) | k) != 0);
                        }
                        umma_commit_2sm(&empty_bar[stage], 0x3); // frees the slot in both CTAs
                        if (kb == kb_end - 1)
                            umma_commit_2sm(&tfull_bar[acc], 0x3); // accumulator complete in both CTAs
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2.., both CTAs) =====================
        constexpr int CW = BN / (EW / 4);          // columns per warp
        constexpr int SBW = EPI_BYTES / EW;        // staging bytes per warp
        const int quad = warp & 3;
        const int cgrp = (warp - 2) >> 2;
        uint32_t it = 0, chunk_ctr = 0;
        uint32_t rphase[2] = {0, 0};
        if (warp == 2 && lane == 0) {
            prefetch_tensormap(&tmap_c);
            if (EMIT)
                prefetch_tensormap(&tmap_d);
        }
        LnRow ln = {1.0f, 0.0f};
        if (LN && pair < num_tiles) { // first tile: fetched here; later tiles one tile ahead
            float2 part[kMaxLnSlots];
            ln_row_fetch(p, (pair / p.splits / num_n) * BM2 + (int)rank * BM + quad * 32 + lane, part);
            ln = ln_row_finish(p, part);
        }
        if (!EMIT && coef_in_smem<LN, SBW, CW / 32>(p) && pair < num_tiles)
            ln_coef_copy<LN, CW / 32>(p, reinterpret_cast<float *>(smem + L::EPI_OFFSET + (warp - 2) * SBW + kCoefOffset), lane,
                                      ((pair / p.splits) % num_n) * BN + cgrp * CW);
        for (int tile = pair; tile < num_tiles; tile += num_pairs, it++) {
            const int t2 = tile / p.splits, split = tile - t2 * p.splits;
            const int m_blk = t2 / num_n, n_blk = t2 - m_blk * num_n;
            const int t2n = (tile + num_pairs) / p.splits;
            const int next_row0 = tile + num_pairs < num_tiles ? (t2n / num_n) * BM2 + (int)rank * BM + quad * 32 : -1;
            const int next_col = tile + num_pairs < num_tiles ? (t2n % num_n) * BN + cgrp * CW : -1;
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            float *stage_tile = reinterpret_cast<float *>(smem + L::EPI_OFFSET + (warp - 2) * SBW);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + cgrp * CW;
            bool tile_ok;
            if (EMIT) {
                static_assert(!EMIT || SBW >= 12288, "emit mode needs 12 KB of staging per epilogue warp");
#if VITCU_EMIT_VARIANT == 4
                tile_ok = epilogue_tile_emit_lsu<CW / 32>(p, reinterpret_cast<uint8_t *>(stage_tile), lane,
                                                          m_blk * BM2 + (int)rank * BM + quad * 32, n_blk * BN + cgrp * CW, taddr,
                                                          &tfull_bar[acc], acc_phase, wd);
#else
                tile_ok = epilogue_tile_emit<CW / 32>(p, &tmap_c, &tmap_d, reinterpret_cast<uint8_t *>(stage_tile),
                                                      &res_bar[2 * (warp - 2)], rphase, lane,
                                                      m_blk * BM2 + (int)rank * BM + quad * 32, n_blk * BN + cgrp * CW, taddr,
                                                      &tfull_bar[acc], acc_phase, wd);
#endif
            } else {
                tile_ok = epilogue_tile<CW / 32, SBW, EW == 8, LN>(p, C, &tmap_c, chunk_ctr, stage_tile, lane,
                                                               m_blk * BM2 + (int)rank * BM + quad * 32, n_blk * BN + cgrp * CW,
                                                               taddr, &tfull_bar[acc], acc_phase, wd, ln, next_row0, next_col, it,
                                                               split == 0 ? 1.0f : 0.0f);
            }
            if (!tile_ok)
                break;
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0)); // the leader's barrier
        }
        if (lane == 0)
            tma_wait_group<0>();
    }

    // ===================== teardown =====================
    tcgen05_fence_before();
    cluster_sync_all(); // neither CTA may exit (or free TMEM) while its peer can still touch it
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
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

template <int BN, int STAGES, bool LN = false, int PIECES = 1>
int launch(const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &tc, void *C, const EpiParams &p, int sms,
           cudaStream_t st)
{
    using L = SmemLayout<BN, STAGES, PIECES>;
    auto kernel = gemm_bf16_tc_kernel<BN, STAGES, LN, PIECES>;
    static bool configured[64] = {false};
    int dev = 0;
    VITCU_TRY(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        VITCU_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL));
        configured[dev] = true;
    }
    const int num_tiles = ((p.M + BM - 1) / BM) * (p.N / BN) * p.splits;
    const int grid = num_tiles < sms ? num_tiles : sms;
    EpiParams q = p;
    static const bool no_prefetch = getenv("VITCU_W_PREFETCH") && !strcmp(getenv("VITCU_W_PREFETCH"), "0");
    // one work item per CTA: small M.  Measured (same box, p50 of the batch-1 forward): BF16 0.5438 -> 0.5374 ms; the FP32
    // chain (six piece products, bound by L2 -> SM bytes) 0.9941 -> 1.0003 ms, so it stays off there (VITCU_W_PREFETCH=0: off)
    q.prefetch_w = num_tiles <= sms && p.nseg == 1 && !no_prefetch;
    VITCU_TRY(launch_kernel(kernel, grid, kThreads, L::TOTAL, st, ta, tb, tc, C, q, watchdog_flag()));
    VITCU_LAUNCHED_KIND(LK_GEMM_1CTA);
    return 0;
}

template <int STAGES, int EW, uint32_t EPI_BYTES = 65536, bool EMIT = false, bool LN = false, bool FP8 = false>
int launch_pair(const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &tc, const CUtensorMap &td, void *C,
                const EpiParams &p, int sms, cudaStream_t st)
{
    using L = SmemLayout2<STAGES, EPI_BYTES>;
    auto kernel = gemm_bf16_tc2_kernel<STAGES, EW, EPI_BYTES, EMIT, LN, FP8>;
    static bool configured[64] = {false};
    int dev = 0;
    VITCU_TRY(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        VITCU_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL));
        configured[dev] = true;
    }
    const int num_tiles = ((p.M + 255) / 256) * (p.N / 256) * p.splits;
    const int pairs = num_tiles < sms / 2 ? num_tiles : sms / 2;
    VITCU_TRY(launch_kernel(kernel, 2 * pairs, 64 + 32 * EW, L::TOTAL, st, ta, tb, tc, td, C, p, watchdog_flag()));
    VITCU_LAUNCHED_KIND(LK_GEMM_PAIR);
    return 0;
}

} // namespace

namespace vitcu {

int make_tensor_map_2d(CUtensorMap *map, const void *base, int elem_bytes, uint64_t rows, uint64_t cols,
                       uint64_t ld_bytes, uint32_t box_rows, uint32_t box_cols, int swizzle_bytes)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn)
        return set_error(VITCU_E_NODEVICE, __FILE__, __LINE__, "cuTensorMapEncodeTiled is unavailable");
    const CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                   : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUresult r = fn(map, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(VITCU_E_ARG, __FILE__, __LINE__, "cuTensorMapEncodeTiled rejected the tensor");
    return 0;
}

int device_sm_count()
{
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
        return 148;
    if (!sms[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        sms[dev] = n;
    }
    return sms[dev];
}

} // namespace vitcu

// CTA pairs on 256x256 tiles when N allows and every pair gets work
static bool pair_eligible(int M, int N)
{
    static const bool force_1cta = getenv("VITCU_GEMM_MODE") && !strcmp(getenv("VITCU_GEMM_MODE"), "1cta");
    return !force_1cta && N % 256 == 0 && ((M + 255) / 256) * (N / 256) >= device_sm_count() / 2;
}

extern "C" int vitcu_gemm_bf16_emit_supported(int M, int N) { return M > 0 && N > 0 && pair_eligible(M, N) ? 1 : 0; }

extern "C" int vitcu_gemm_split_k_pays(int M, int N)
{
    return M > 0 && N > 0 && N % 128 == 0 && !pair_eligible(M, N) && ((M + BM - 1) / BM) * (N / 128) * 2 <= device_sm_count() ? 1 : 0;
}

static int gemm_dispatch(const vitcu_bf16 *A, const vitcu_bf16 *W, void *C, const vitcu_gemm_desc *d, bool split3,
                         vitcu_stream s)
{
    VITCU_REQUIRE(A && W && C && d, "NULL argument");
    VITCU_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "empty GEMM");
    VITCU_REQUIRE(d->K % BK == 0, "bf16 GEMM needs K % 64 == 0");
    VITCU_REQUIRE(d->N % 128 == 0, "bf16 GEMM needs N % 128 == 0");
    VITCU_REQUIRE(d->bias && ((uintptr_t)d->bias & 15) == 0, "bias is required (16-byte aligned)");
    VITCU_REQUIRE(d->epilogue != VITCU_EPI_BIAS_RESIDUAL || d->residual, "residual pointer missing");
    VITCU_REQUIRE(d->epilogue != VITCU_EPI_PATCH_EMBED || (d->pos && d->patches > 0 && d->tokens > d->patches),
                  "patch-embed epilogue needs pos, patches, tokens");
    EpiParams p;
    memset(&p, 0, sizeof(p));
    p.M = d->M;
    p.N = d->N;
    p.K = d->K;
    p.ldc = d->ldc ? d->ldc : (size_t)d->N;
    p.epilogue = d->epilogue;
    p.bias = d->bias;
    p.residual = d->residual;
    p.pos = d->pos;
    p.patches = d->patches;
    p.tokens = d->tokens;
    p.out_bf16 = d->out_bf16;
    p.exact_gelu = split3 && !d->out_bf16;
    p.kb_elems = BK;
    p.acc_scale = 1.0f;
    // LayerNorm folded into the GEMM (see vitcu_gemm_desc)
    if (d->ln_stats) {
        VITCU_REQUIRE(!split3 && d->ln_colsum && d->ln_slots > 0, "LayerNorm-folded GEMM needs bf16 operands, column sums and slots");
        VITCU_REQUIRE(d->epilogue == VITCU_EPI_BIAS || d->epilogue == VITCU_EPI_BIAS_GELU, "LayerNorm fold applies to bias / GELU epilogues");
        VITCU_REQUIRE(d->out_bf16 && ((uintptr_t)C & 15) == 0 && ((uintptr_t)d->bias & 15) == 0 && ((uintptr_t)d->ln_colsum & 15) == 0,
                      "LayerNorm-folded GEMM writes bf16 through the TMA store path and needs 16-byte aligned coefficient vectors");
        p.ln_stats = reinterpret_cast<const float2 *>(d->ln_stats);
        p.ln_colsum = d->ln_colsum;
        p.ln_slots = d->ln_slots;
        p.ln_inv_d = 1.0f / (float)d->K;
    }
    const bool emit = d->emit_bf16 != nullptr;
    if (emit) {
        VITCU_REQUIRE(!split3 && d->emit_stats && d->epilogue == VITCU_EPI_BIAS_RESIDUAL && d->residual == (const float *)C &&
                          !d->out_bf16 && ((uintptr_t)C & 15) == 0 && ((uintptr_t)d->emit_bf16 & 15) == 0,
                      "emit mode needs the in-place fp32 residual epilogue");
        VITCU_REQUIRE(vitcu_gemm_bf16_emit_supported(d->M, d->N), "emit mode needs the CTA-pair kernel (N % 256 == 0, enough tiles)");
        p.emit_stats = reinterpret_cast<float2 *>(d->emit_stats);
        p.emit_ptr = d->emit_bf16;
        p.emit_fp8 = d->emit_fp8;
        p.emit_scale = d->emit_scale;
        VITCU_REQUIRE(!d->emit_fp8 || d->emit_scale > 0.f, "emit_fp8 needs a positive emit_scale");
    }
    VITCU_REQUIRE(p.ldc % 8 == 0, "ldc must be a multiple of 8");
    const uint64_t kphys = split3 ? 3 * (uint64_t)d->K : (uint64_t)d->K; // physical operand width
    const size_t lda = split3 ? (size_t)kphys : (d->lda ? d->lda : (size_t)d->K);
    VITCU_REQUIRE(lda % 8 == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0, "operands must be 16-byte aligned");
    p.seg_kb = d->K / BK;
    p.splits = 1;
    if (split3) {
        // small products first: a3 w1, a2 w2, a1 w3, a2 w1, a1 w2, a1 w1
        const int K = d->K;
        const int as[6] = {2 * K, K, 0, K, 0, 0}, bs[6] = {0, K, 2 * K, 0, K, 0};
        p.nseg = 6;
        for (int i = 0; i < 6; i++) {
            p.a_seg[i] = as[i];
            p.b_seg[i] = bs[i];
        }
    } else {
        p.nseg = 1;
    }

    const int sms = device_sm_count();
    // VITCU_GEMM_MODE=1cta forces the single-CTA kernels, VITCU_GEMM_EPI=lsu the LSU-store epilogue (A/B measurements)
    static const bool force_lsu = getenv("VITCU_GEMM_EPI") && !strcmp(getenv("VITCU_GEMM_EPI"), "lsu");
    // Output through the TMA engine: bf16 / fp32 tiles are stored; the in-place residual update
    // C = C + (acc + bias) becomes a TMA reduce-add, so the fp32 residual stream is never read by the SMs.
    p.tma_out = 0;
    if (d->accumulate) {
        // C was zeroed by the caller: every work item reduce-adds its partial tile (the first K slice carries the bias),
        // which is what lets a small-M product be cut along K (see the split computation below)
        VITCU_REQUIRE(!d->out_bf16 && d->epilogue == VITCU_EPI_BIAS && !d->ln_stats && !emit && ((uintptr_t)C & 15) == 0,
                      "accumulate mode: fp32 output, bias epilogue, 16-byte aligned C");
        p.tma_out = 2;
    } else if (!force_lsu && ((uintptr_t)C & 15) == 0) {
        const bool plain = p.epilogue == VITCU_EPI_BIAS || p.epilogue == VITCU_EPI_BIAS_GELU;
        if (p.out_bf16 && plain)
            p.tma_out = 1;
        else if (!p.out_bf16 && p.epilogue == VITCU_EPI_BIAS_RESIDUAL && p.residual == (const float *)C)
            p.tma_out = 2;
        else if (!p.out_bf16 && plain)
            p.tma_out = 3;
    }
    CUtensorMap ta, tb, tc, td;
    memset(&td, 0, sizeof(td));
    int rc = 0;
    if (emit) {
        p.tma_out = 4;
        rc = d->emit_fp8 ? make_tensor_map_2d(&td, d->emit_bf16, 1, (uint64_t)d->M, (uint64_t)d->N, (size_t)d->N, 32, 32, 0)
                         : make_tensor_map_2d(&td, d->emit_bf16, 2, (uint64_t)d->M, (uint64_t)d->N, (size_t)d->N * 2, 32, 32, 64);
        if (rc)
            return rc;
    }
    if (p.tma_out == 1)
        rc = make_tensor_map_2d(&tc, C, 2, (uint64_t)d->M, (uint64_t)d->N, p.ldc * 2, 32, 32, 64);
    else if (p.tma_out >= 2)
        rc = make_tensor_map_2d(&tc, C, 4, (uint64_t)d->M, (uint64_t)d->N, p.ldc * 4, 32, 32, 128);
    else
        memset(&tc, 0, sizeof(tc));
    if (rc)
        return rc;
    // CTA pairs on 256x256 tiles when N allows and every pair gets work
    const bool pair = pair_eligible(d->M, d->N);
    rc = make_tensor_map_2d(&ta, A, 2, (uint64_t)d->M, kphys, lda * 2, BM, BK);
    if (rc)
        return rc;
    if (pair) {
        rc = make_tensor_map_2d(&tb, W, 2, (uint64_t)d->N, kphys, kphys * 2, 128, BK);
        if (rc)
            return rc;
        if (emit) // 4 operand stages: the emit epilogue stages 12 KB per warp
            return launch_pair<4, 8, 98304, true>(ta, tb, tc, td, C, p, sms, as_stream(s));
        // 16 epilogue warps (VITCU_GEMM_EW=16; they need the TMA output path, their staging share is 4 KB) paid
        // off for the 17-instruction rational GELU (1195 vs 1127 TFLOP/s); with the 9-instruction packed
        // MUFU.TANH form 8 warps with double-buffered tcgen05.ld are as fast or faster (0.195 vs 0.198 ms)
        static const int force_ew = getenv("VITCU_GEMM_EW") ? atoi(getenv("VITCU_GEMM_EW")) : 0;
        const bool ew16 = p.tma_out != 0 && (force_ew ? force_ew == 16 : (VITCU_GELU_FORM == 0 && p.epilogue == VITCU_EPI_BIAS_GELU));
        if (ew16 && !p.ln_stats)
            return launch_pair<5, 16>(ta, tb, tc, td, C, p, sms, as_stream(s));
        // VITCU_GEMM_STAGES=6: a sixth operand stage in exchange for single-buffered epilogue staging (A/B)
        // measured (M=50432): bf16-output launches unchanged, fp32 reduce-add launches (out_proj, fc2) +1.5 %
        static const int stages = getenv("VITCU_GEMM_STAGES") ? atoi(getenv("VITCU_GEMM_STAGES")) : 0;
        if (p.tma_out != 0 && (stages == 6 || (stages == 0 && p.tma_out == 2)))
            return launch_pair<6, 8, 32768>(ta, tb, tc, td, C, p, sms, as_stream(s));
        if (p.ln_stats)
            return launch_pair<5, 8, 65536, false, true>(ta, tb, tc, td, C, p, sms, as_stream(s));
        return launch_pair<5, 8>(ta, tb, tc, td, C, p, sms, as_stream(s));
    }
    // small M with the reduce-add epilogue (batch-1 latency: out-proj / fc2, and qkv / fc1 / patch embedding of the
    // FP32 chain in accumulate mode) on CTA pairs with split-K -- VITCU_GEMM_MODE=pair, an A/B variant that is NOT the
    // default: the 256 x 256 pair tile moves 57 KB instead of 98 KB from L2 per 256 x 256 x 64 block and its k-loop
    // is 1 - 1.5 us shorter per launch, but a cluster launch with 512 TMEM columns per CTA costs 4.5 us more before
    // the first load (same box: FP32 batch-1 forward 1.040 vs 0.995 ms, BF16 0.648 vs 0.548 ms)
    const char *mode = getenv("VITCU_GEMM_MODE"); // read per call: the tests run both kernels in one process
    const bool force_1cta_small = !(mode && !strcmp(mode, "pair"));
    if (!force_1cta_small && p.tma_out == 2 && d->N % 256 == 0 && !p.ln_stats) {
        const int tiles = ((d->M + 255) / 256) * (d->N / 256), total_kb = p.nseg * p.seg_kb;
        int splits = (sms / 2) / tiles;
        if (splits > total_kb / 2)
            splits = total_kb / 2;
        if (splits < 1)
            splits = 1;
        const int kb_per = (total_kb + splits - 1) / splits;
        p.splits = (total_kb + kb_per - 1) / kb_per; // no empty slice (its accumulator would never be committed)
        rc = make_tensor_map_2d(&tb, W, 2, (uint64_t)d->N, kphys, kphys * 2, 128, BK);
        if (rc)
            return rc;
        return launch_pair<6, 8, 32768>(ta, tb, tc, td, C, p, sms, as_stream(s));
    }
    // 128x256 tiles when they divide N and still give every SM work; else 128x128
    const bool wide = d->N % 256 == 0 && ((d->M + BM - 1) / BM) * (d->N / 256) >= sms;
    // split-bf16 operands on the 128 x 128 kernel: a ring slot holds all three pieces of both operands for one logical
    // k-block and the six piece products are issued from tiles loaded ONCE (six tile loads per 64 columns of K instead
    // of twelve).  These launches are bound by L2 -> SM bytes (profiles/r02_batch1_latency.md).  Needs the TMA output
    // path (4 KB of epilogue staging per warp).  VITCU_FP32_FUSED=0: one product per ring slot (A/B).
    const char *fenv = getenv("VITCU_FP32_FUSED"); // read per call: the tests run both forms in one process
    const bool fused3 = split3 && !wide && p.tma_out != 0 && !p.ln_stats && !(fenv && !strcmp(fenv, "0"));
    if (!wide && p.tma_out == 2) {
        // small M (batch-1 latency): split K so that the work items roughly fill the SMs, at least
        // two k-blocks per slice (one logical k-block = six products in the fused form); the slices meet in C
        // through TMA reduce-add
        const int tiles = ((d->M + BM - 1) / BM) * (d->N / 128), total_kb = fused3 ? p.seg_kb : p.nseg * p.seg_kb;
        const int cap = fused3 ? total_kb : total_kb / 2;
        int splits = sms / tiles;
        if (splits > cap)
            splits = cap;
        if (splits < 1)
            splits = 1;
        // no empty slice: a work item without k-blocks would never commit its accumulator and its
        // epilogue would wait forever (K = 4096 over 9 slices of 8 k-blocks leaves the ninth empty)
        const int kb_per = (total_kb + splits - 1) / splits;
        p.splits = (total_kb + kb_per - 1) / kb_per;
    }
    // (128 x 64 tiles for the outputs that cannot be summed across K slices -- bf16 qkv / fc1 of the BF16 chain at batch 1,
    // 72 / 96 CTAs instead of 36 / 48 -- were measured and are slower: 0.555 vs 0.538 ms per forward.  Twice the CTAs
    // re-read the A tile twice as often and these launches are bound by L2 -> SM bytes.)
    rc = make_tensor_map_2d(&tb, W, 2, (uint64_t)d->N, kphys, kphys * 2, wide ? 256 : 128, BK);
    if (rc)
        return rc;
    if (p.ln_stats)
        return wide ? launch<256, 3, true>(ta, tb, tc, C, p, sms, as_stream(s)) : launch<128, 5, true>(ta, tb, tc, C, p, sms, as_stream(s));
    if (fused3)
        return launch<128, 2, false, 3>(ta, tb, tc, C, p, sms, as_stream(s));
    if (wide)
        return launch<256, 3>(ta, tb, tc, C, p, sms, as_stream(s));
    return launch<128, 5>(ta, tb, tc, C, p, sms, as_stream(s));
}

// E4M3 operands (per-tensor scales), CTA-pair kernel only.  Two uses: the LayerNorm-folded fc1 (GELU, e4m3 or bf16 out)
// and the residual fc2 that emits bf16(x) / e4m3(x) + row sums.
extern "C" int vitcu_gemm_e4m3(const uint8_t *A, const uint8_t *W, void *C, const vitcu_gemm_desc *d, vitcu_stream s)
{
    VITCU_REQUIRE(A && W && C && d, "NULL argument");
    VITCU_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0 && d->K % 128 == 0 && d->N % 256 == 0, "e4m3 GEMM needs K % 128 == 0, N % 256 == 0");
    VITCU_REQUIRE(pair_eligible(d->M, d->N), "e4m3 GEMM runs on the CTA-pair kernel only (M too small)");
    VITCU_REQUIRE(d->bias && ((uintptr_t)d->bias & 15) == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)C & 15) == 0,
                  "16-byte aligned operands, output and bias are required");
    VITCU_REQUIRE(d->acc_scale > 0.f, "acc_scale = 1 / (scale_A * scale_W) must be positive");
    EpiParams p;
    memset(&p, 0, sizeof(p));
    p.M = d->M;
    p.N = d->N;
    p.K = d->K;
    p.ldc = (size_t)d->N;
    p.epilogue = d->epilogue;
    p.bias = d->bias;
    p.residual = d->residual;
    p.out_bf16 = d->out_bf16;
    p.kb_elems = 128;
    p.acc_scale = d->acc_scale;
    p.nseg = 1;
    p.seg_kb = d->K / 128;
    p.splits = 1;
    const int sms = device_sm_count();
    CUtensorMap ta, tb, tc, td;
    memset(&td, 0, sizeof(td));
    int rc = make_tensor_map_2d(&ta, A, 1, (uint64_t)d->M, (uint64_t)d->K, (size_t)d->K, BM, 128);
    if (!rc)
        rc = make_tensor_map_2d(&tb, W, 1, (uint64_t)d->N, (uint64_t)d->K, (size_t)d->K, 128, 128);
    if (rc)
        return rc;
    if (d->emit_bf16) { // residual update in place + bf16 / e4m3 copy + row sums
        VITCU_REQUIRE(d->emit_stats && d->epilogue == VITCU_EPI_BIAS_RESIDUAL && d->residual == (const float *)C && !d->out_bf16 &&
                          ((uintptr_t)d->emit_bf16 & 15) == 0, "emit mode needs the in-place fp32 residual epilogue");
        VITCU_REQUIRE(!d->emit_fp8 || d->emit_scale > 0.f, "emit_fp8 needs a positive emit_scale");
        p.emit_stats = reinterpret_cast<float2 *>(d->emit_stats);
        p.emit_ptr = d->emit_bf16;
        p.emit_fp8 = d->emit_fp8;
        p.emit_scale = d->emit_scale;
        p.tma_out = 4;
        rc = make_tensor_map_2d(&tc, C, 4, (uint64_t)d->M, (uint64_t)d->N, (size_t)d->N * 4, 32, 32, 128);
        if (!rc)
            rc = d->emit_fp8 ? make_tensor_map_2d(&td, d->emit_bf16, 1, (uint64_t)d->M, (uint64_t)d->N, (size_t)d->N, 32, 32, 0)
                             : make_tensor_map_2d(&td, d->emit_bf16, 2, (uint64_t)d->M, (uint64_t)d->N, (size_t)d->N * 2, 32, 32, 64);
        if (rc)
            return rc;
        return launch_pair<4, 8, 98304, true, false, true>(ta, tb, tc, td, C, p, sms, as_stream(s));
    }
    VITCU_REQUIRE(d->ln_stats && d->ln_colsum && d->ln_slots > 0 && ((uintptr_t)d->ln_colsum & 15) == 0,
                  "the non-residual e4m3 GEMM is the LayerNorm-folded one: ln_stats / ln_colsum / ln_slots are required");
    VITCU_REQUIRE(d->epilogue == VITCU_EPI_BIAS || d->epilogue == VITCU_EPI_BIAS_GELU, "LayerNorm fold applies to bias / GELU epilogues");
    VITCU_REQUIRE(d->out_fp8 ? d->out_scale > 0.f : d->out_bf16, "output must be e4m3 (out_fp8 + out_scale) or bf16");
    p.ln_stats = reinterpret_cast<const float2 *>(d->ln_stats);
    p.ln_colsum = d->ln_colsum;
    p.ln_slots = d->ln_slots;
    p.ln_inv_d = 1.0f / (float)d->K;
    if (d->out_fp8) {
        p.tma_out = 5;
        p.out_scale = d->out_scale;
        rc = make_tensor_map_2d(&tc, C, 1, (uint64_t)d->M, (uint64_t)d->N, (size_t)d->N, 32, 32, 0);
    } else {
        p.tma_out = 1;
        rc = make_tensor_map_2d(&tc, C, 2, (uint64_t)d->M, (uint64_t)d->N, (size_t)d->N * 2, 32, 32, 64);
    }
    if (rc)
        return rc;
    return launch_pair<5, 8, 65536, false, true, true>(ta, tb, tc, td, C, p, sms, as_stream(s));
}

extern "C" int vitcu_gemm_bf16(const vitcu_bf16 *A, const vitcu_bf16 *W, void *C, const vitcu_gemm_desc *d,
                               vitcu_stream s)
{
    return gemm_dispatch(A, W, C, d, false, s);
}

extern "C" int vitcu_gemm_bf16x3(const vitcu_bf16 *A3, const vitcu_bf16 *W3, void *C, const vitcu_gemm_desc *d,
                                 vitcu_stream s)
{
    return gemm_dispatch(A3, W3, C, d, true, s);
}
