// csrc/patch_embed_tc.cu -- patch embedding as ONE tensor-core GEMM whose patch gather is done by TMA.
//
// Replaces conv2d_kernel + postprocess (R/conv2d.cl:1-80; host side Conv2d / postConv2d,
// R/ViT_opencl.c:361-442) on the BF16 path; oracle R/ViT_seq.c:25-118.
// R/ = /root/reference/MulticoreMainProject/.
//
// The 16x16 / stride-16 convolution is the GEMM  tokens[p, oc] = sum_k patch[p, k] * w[oc, k],
// k = (c, kh, kw).  Nothing is gathered in a separate kernel: the image [B,3,S,S] is described to
// the TMA engine as the 5-D tensor (kw:16, pw:S/16, kh:16, ph:S/16, b*c), and one box
// {16, S/16, 1, PH, 1} lands in shared memory as [PH*S/16 patches] x [16 kw] fp32 -- exactly a
// K-major operand tile with 64-byte rows (64-byte swizzle) for one (c, kh) slice of K.  The MMA
// consumes the fp32 pixels and fp32 weights directly as TF32 (tcgen05.mma.kind::tf32, M=128,
// N=256, K=8), FP32 accumulation in TMEM; 48 k-blocks (3 channels x 16 kernel rows) per tile.
// Epilogue: + conv bias + position embedding, rows remapped past the class token
// (token = 1 + patch), fp32 residual stream.  The class-token rows are written by cls_rows_kernel.
//
// TF32 keeps 10 mantissa bits of pixels and weights (bf16 would keep 7), and the embedding is
// 0.66 % of the FLOPs, so half-rate TF32 costs nothing measurable.
#include "tc_common.cuh"

using namespace vitcu;
using namespace vitcu::tc;

namespace {

constexpr int kThreadsPE = 192;
constexpr int BN = 256;
constexpr int STAGES = 6;
constexpr uint32_t A_BYTES = 128 * 64;      // 128 patch rows x 16 fp32
constexpr uint32_t B_BYTES = BN * 64;       // 256 filters x 16 fp32
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t BAR_OFFSET = STAGES * STAGE_BYTES;
constexpr uint32_t SMEM_TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;

struct PEParams {
    int batch, side, tokens;   // side = S/16 patches per image row, tokens = side*side + 1
    int ph_box, tiles_per_img; // patch rows per tile, tiles per image
    int embed, num_n;          // output features (768 for ViT-B) and embed / BN column tiles
    uint32_t a_box_bytes;      // bytes one A box delivers
    const float *bias;         // [768]
    const float *pos;          // [tokens, 768]
    float *x;                  // [batch*tokens, 768]
    int splits;                // K slices per tile (accumulate form: x already holds the position rows, slices add)
};

__device__ __forceinline__ void tma_load_5d(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2,
                                            int c3, int c4)
{
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
                 "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
// K-major operand tile with 64-byte rows and 64-byte swizzle: 8-row groups are 512 B apart
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(512 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(4) << 61; // SWIZZLE_64B
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
                 : "memory");
}

__global__ void __launch_bounds__(kThreadsPE, 1)
patch_embed_tc_kernel(const __grid_constant__ CUtensorMap tmap_img, const __grid_constant__ CUtensorMap tmap_w,
                      const PEParams p, uint32_t *watchdog_flag)
{
    constexpr uint32_t IDESC = umma_idesc_tf32(128, BN);
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + BAR_OFFSET);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tfull_bar = empty_bar + STAGES;
    uint64_t *tempty_bar = tfull_bar + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + 2);
    volatile uint32_t *cta_abort = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4);
        }
        *cta_abort = 0;
        fence_barrier_init();
    }
    if (warp == 1)
        tmem_alloc(tmem_slot, 512);
    // the A tile has 128 rows but a box fills only ph_box*side of them: clear the rest once so the
    // unused accumulator rows stay finite
    for (uint32_t i = threadIdx.x; i < STAGES * (A_BYTES / 16); i += kThreadsPE) {
        const uint32_t st = i / (A_BYTES / 16), off = i % (A_BYTES / 16);
        *reinterpret_cast<uint4 *>(smem + st * STAGE_BYTES + off * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // every CTA holds its TMEM now: the next kernel may start its prologue; our own global-memory
    // traffic (TMA loads, epilogue stores) waits for the previous kernel to complete
    pdl_trigger();
    pdl_wait();
    const Watchdog wd{cta_abort, watchdog_flag};

    const int NUM_N = p.num_n; // 3 for the 768-wide model
    constexpr int NUM_KB = 3 * kPatch; // (channel, kernel row) slices
    // work item = (tile, K slice).  splits > 1 is the accumulate form (vitcu_patch_embed_tc_acc, a handful of images):
    // the token rows already hold the position embedding and every slice adds its partial product with red.global
    // (slice 0 carries the conv bias), so that one image is 96 work items instead of 6
    const int kb_per = NUM_KB / p.splits; // splits divides 48
    const int num_tiles = p.batch * p.tiles_per_img * NUM_N * p.splits;

    if (warp == 0) {
        // ===================== TMA producer: the gather happens here =====================
        if (elect_one()) {
            prefetch_tensormap(&tmap_img);
            prefetch_tensormap(&tmap_w);
        }
        uint32_t stage = 0, phase = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < num_tiles && ok; tile += gridDim.x) {
            const int split = tile % p.splits, t2 = tile / p.splits;
            const int n_blk = t2 % NUM_N, mt = t2 / NUM_N;
            const int img = mt / p.tiles_per_img, ph0 = (mt - img * p.tiles_per_img) * p.ph_box;
            for (int kb = split * kb_per; kb < (split + 1) * kb_per; kb++) {
                if (!(ok = mbar_wait_warp(&empty_bar[stage], phase ^ 1, wd, 1)))
                    break;
                if (elect_one()) {
                    uint8_t *sa = smem + stage * STAGE_BYTES;
                    const int c = kb / kPatch, kh = kb - c * kPatch;
                    mbar_arrive_expect_tx(&full_bar[stage], p.a_box_bytes + B_BYTES);
                    // box {16 kw, side pw, 1 kh, ph_box ph, 1 image-channel}: [patch][kw] rows of 64 bytes
                    tma_load_5d(sa, &tmap_img, &full_bar[stage], 0, 0, kh, ph0, img * 3 + c);
                    tma_load_2d(sa + A_BYTES, &tmap_w, &full_bar[stage], kb * kPatch, n_blk * BN);
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
            for (int kb = 0; kb < kb_per; kb++) {
                if (!(ok = mbar_wait_warp(&full_bar[stage], phase, wd, 3)))
                    break;
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t a_desc = umma_desc_k_sw64(sa), b_desc = umma_desc_k_sw64(sa + A_BYTES);
                    // 16 fp32 per row = two K=8 TF32 steps, +32 bytes inside the 64-byte swizzled row
                    umma_tf32_ss(d_tmem, a_desc, b_desc, IDESC, kb != 0);
                    umma_tf32_ss(d_tmem, a_desc + 2, b_desc + 2, IDESC, 1);
                    umma_commit(&empty_bar[stage]);
                    if (kb == kb_per - 1)
                        umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quad = warp & 3;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
            const int split = tile % p.splits, t2 = tile / p.splits;
            const int n_blk = t2 % NUM_N, mt = t2 / NUM_N;
            const int img = mt / p.tiles_per_img, ph0 = (mt - img * p.tiles_per_img) * p.ph_box;
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            const int rows_here = min(p.ph_box, p.side - ph0) * p.side; // patches of this tile that exist
            const int r = quad * 32 + lane;
            const int patch = ph0 * p.side + r;
            const bool valid = r < rows_here;
            float *dst = p.x + (static_cast<size_t>(img) * p.tokens + 1 + patch) * p.embed + n_blk * BN;
            const float *pos = p.pos + static_cast<size_t>(1 + patch) * p.embed + n_blk * BN;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
            // conv bias and position rows of a chunk are requested a chunk ahead, the first before the wait for the
            // accumulator: read at their use they were an L2 / DRAM round trip per 32 columns on the epilogue's chain
            const bool acc_form = p.splits > 1;
            const float bias_on = (!acc_form || split == 0) ? 1.0f : 0.0f;
            float4 bn[8], en[8];
            auto fetch_coef = [&](int c) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    bn[j] = __ldg(reinterpret_cast<const float4 *>(p.bias + n_blk * BN + c * 32) + j);
                    en[j] = (valid && !acc_form) ? __ldg(reinterpret_cast<const float4 *>(pos + c * 32) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            fetch_coef(0);
            if (!mbar_wait_warp(&tfull_bar[acc], acc_phase, wd, 4))
                break;
            tcgen05_fence_after();
#pragma unroll 1
            for (int c = 0; c < BN / 32; c++) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(taddr + c * 32, v);
                tmem_ld_wait();
                float4 o[8];
#pragma unroll
                for (int j = 0; j < 8; j++)
                    o[j] = make_float4(fmaf(bn[j].x, bias_on, __uint_as_float(v[4 * j + 0])) + en[j].x,
                                       fmaf(bn[j].y, bias_on, __uint_as_float(v[4 * j + 1])) + en[j].y,
                                       fmaf(bn[j].z, bias_on, __uint_as_float(v[4 * j + 2])) + en[j].z,
                                       fmaf(bn[j].w, bias_on, __uint_as_float(v[4 * j + 3])) + en[j].w);
                if (c + 1 < BN / 32)
                    fetch_coef(c + 1);
                if (valid && acc_form) { // accumulate form: x += partial (+ bias once); the position rows are in x already
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        atomicAdd(reinterpret_cast<float4 *>(dst + c * 32) + j, o[j]);
                } else if (valid) {
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        reinterpret_cast<float4 *>(dst + c * 32)[j] = o[j];
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&tempty_bar[acc]);
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
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
}

// images [B,3,S,S] fp32, conv_w [768, 3*16*16] fp32 -> x[b*T + 1 + p, :] = conv + bias + pos[1 + p, :]
// accumulate != 0: x[b*T + 1 + p, :] += conv + bias, K cut into slices (the rows hold pos[1 + p, :] already)
static int patch_embed_tc_launch(const float *images, const float *conv_w, const float *conv_b, const float *pos,
                                 float *x, int batch, int img, int embed, int accumulate, vitcu_stream s)
{
    VITCU_REQUIRE(images && conv_w && conv_b && pos && x, "NULL argument");
    VITCU_REQUIRE(embed > 0 && embed % BN == 0, "embedding width must be a multiple of the 256-column tile");
    VITCU_REQUIRE(batch > 0 && img > 0 && img % kPatch == 0, "image side must be a positive multiple of 16");
    const int side = img / kPatch;
    VITCU_REQUIRE(side <= 128, "image too large for one 128-row tile per patch row");
    VITCU_REQUIRE(((uintptr_t)images & 15) == 0 && ((uintptr_t)conv_w & 15) == 0 && ((uintptr_t)x & 15) == 0,
                  "buffers must be 16-byte aligned");
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return set_error(VITCU_E_NODEVICE, __FILE__, __LINE__, "cuTensorMapEncodeTiled is unavailable");
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    PEParams p;
    p.batch = batch;
    p.side = side;
    p.tokens = side * side + 1;
    p.ph_box = 128 / side < side ? 128 / side : side;
    p.tiles_per_img = (side + p.ph_box - 1) / p.ph_box;
    p.a_box_bytes = (uint32_t)(kPatch * side * p.ph_box * sizeof(float));
    p.bias = conv_b;
    p.pos = pos;
    p.x = x;
    p.embed = embed;
    p.num_n = embed / BN;

    CUtensorMap timg, tw;
    {
        // (kw, pw, kh, ph, image*channel) view of the NCHW image
        cuuint64_t dims[5] = {kPatch, (cuuint64_t)side, kPatch, (cuuint64_t)side, (cuuint64_t)batch * 3};
        cuuint64_t strides[4] = {kPatch * sizeof(float), (cuuint64_t)img * sizeof(float),
                                 (cuuint64_t)kPatch * img * sizeof(float), (cuuint64_t)img * img * sizeof(float)};
        cuuint32_t box[5] = {kPatch, (cuuint32_t)side, 1, (cuuint32_t)p.ph_box, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        if (fn(&timg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float *>(images), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return set_error(VITCU_E_ARG, __FILE__, __LINE__, "cuTensorMapEncodeTiled rejected the image tensor");
    }
    int rc = make_tensor_map_2d(&tw, conv_w, 4, embed, 3 * kPatch * kPatch, 3 * kPatch * kPatch * sizeof(float), BN, kPatch, 64);
    if (rc)
        return rc;
    static bool configured[64] = {false};
    int dev = 0;
    VITCU_TRY(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        VITCU_TRY(cudaFuncSetAttribute(patch_embed_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL));
        configured[dev] = true;
    }
    const int sms = device_sm_count();
    // accumulate form: as many K slices (a divisor of the 48 k-blocks, at least 3 k-blocks each) as fill the SMs
    p.splits = 1;
    if (accumulate) {
        const int tiles = batch * p.tiles_per_img * (embed / BN);
        static const int cand[5] = {16, 12, 8, 4, 2};
        for (int i = 0; i < 5 && p.splits == 1; i++)
            if (tiles * cand[i] <= sms)
                p.splits = cand[i];
    }
    const int num_tiles = batch * p.tiles_per_img * (embed / BN) * p.splits;
    VITCU_TRY(launch_kernel(patch_embed_tc_kernel, num_tiles < sms ? num_tiles : sms, kThreadsPE, SMEM_TOTAL, as_stream(s),
                            timg, tw, p, watchdog_flag()));
    VITCU_LAUNCHED_KIND(LK_PATCH_EMBED_TC);
    return 0;
}

extern "C" int vitcu_patch_embed_tc_ex(const float *images, const float *conv_w, const float *conv_b, const float *pos,
                                       float *x, int batch, int img, int embed, vitcu_stream s)
{
    return patch_embed_tc_launch(images, conv_w, conv_b, pos, x, batch, img, embed, 0, s);
}

extern "C" int vitcu_patch_embed_tc_acc(const float *images, const float *conv_w, const float *conv_b, const float *pos,
                                        float *x, int batch, int img, int embed, vitcu_stream s)
{
    return patch_embed_tc_launch(images, conv_w, conv_b, pos, x, batch, img, embed, 1, s);
}

extern "C" int vitcu_patch_embed_tc(const float *images, const float *conv_w, const float *conv_b, const float *pos,
                                    float *x, int batch, int img, vitcu_stream s)
{
    return vitcu_patch_embed_tc_ex(images, conv_w, conv_b, pos, x, batch, img, kEmbed, s);
}
