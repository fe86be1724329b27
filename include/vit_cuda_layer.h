/* include/vit_cuda_layer.h -- the thin C-ABI device layer of the B200 engine.
 *
 * This header is what replaces the reference's OpenCL glue
 * (/root/reference/MulticoreMainProject/kernelHandler.h:6-16, kernelHandler.c and
 * the clCreate* / clEnqueue* calls all over ViT_opencl.c).  R/ = that directory.
 * The C host (vit-with-opencl_b200/host/) calls ONLY these functions; they are
 * implemented in vit-with-opencl_b200/csrc/ as extern "C" wrappers around
 * hand-written sm_100a kernels.  Plain pointers and sizes only -- no C++,
 * CUDA-runtime or torch types cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success and a non-zero code on failure
 *     (a cudaError_t value, or VITCU_E_* below); vitcu_last_error() then
 *     describes it with the failing source line.  The reference's convention
 *     of print + exit (CHECK_ERROR, R/kernelHandler.h:6-10) is applied by the
 *     C host around these calls (VIT_CHECK in host/vit_engine.c), not in here.
 *   - `stream` is an opaque handle from vitcu_stream_create (NULL = default).
 *   - device pointers are `void *` / typed pointers into device memory of the
 *     current device; "bf16" buffers are uint16_t bit patterns.
 *   - all matrices are row-major; weights are [out_features, in_features]
 *     exactly as the reference stores them (R/ViT_seq.c:304), i.e. both GEMM
 *     operands are K-major, which is what tcgen05 wants.
 */
#ifndef VIT_CUDA_LAYER_H
#define VIT_CUDA_LAYER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VITCU_E_ARG 90001      /* bad argument / unsupported shape */
#define VITCU_E_NODEVICE 90002 /* no sm_100 device */
#define VITCU_E_WATCHDOG 90003 /* a kernel-side pipeline wait timed out */

typedef void *vitcu_stream;
typedef void *vitcu_event;
typedef void *vitcu_graph;
typedef uint16_t vitcu_bf16;

/* ---- runtime (replaces clGetPlatformIDs/clCreateContext/queues, R/ViT_opencl.c:799-861) */
const char *vitcu_last_error(void);
int vitcu_device_count(int *count);
int vitcu_set_device(int device);
/* one-time per-device set-up (shared-memory opt-ins, watchdog flag); must run
 * on the current device before any stream capture */
int vitcu_prepare_device(void);
/* fills name (<=255 chars), SM count, compute capability major*10+minor, total bytes */
int vitcu_device_info(int device, char *name, int *sms, int *cc, size_t *mem_bytes);
int vitcu_stream_create(vitcu_stream *s);
int vitcu_stream_destroy(vitcu_stream s);
int vitcu_stream_sync(vitcu_stream s);
int vitcu_device_sync(void);
int vitcu_event_create(vitcu_event *e);
int vitcu_event_destroy(vitcu_event e);
int vitcu_event_record(vitcu_event e, vitcu_stream s);
int vitcu_event_sync(vitcu_event e);
int vitcu_stream_wait_event(vitcu_stream s, vitcu_event e);
int vitcu_event_elapsed_ms(vitcu_event start, vitcu_event stop, float *ms);

/* ---- memory (replaces clCreateBuffer/clEnqueueWriteBuffer/ReadBuffer, R/ViT_opencl.c:125-330) */
int vitcu_malloc(void **dptr, size_t bytes);
int vitcu_free(void *dptr);
int vitcu_host_alloc(void **hptr, size_t bytes); /* pinned */
int vitcu_host_free(void *hptr);
int vitcu_host_register(void *hptr, size_t bytes); /* pin caller memory in place */
int vitcu_host_unregister(void *hptr);
/* *pinned = 1 when hptr is page-locked (cudaHostAlloc / cudaHostRegister) memory a DMA can read in place */
int vitcu_host_is_pinned(const void *hptr, int *pinned);
int vitcu_memcpy_h2d(void *dst, const void *src, size_t bytes, vitcu_stream s);
int vitcu_memcpy_d2h(void *dst, const void *src, size_t bytes, vitcu_stream s);
int vitcu_memcpy_d2d(void *dst, const void *src, size_t bytes, vitcu_stream s);
int vitcu_memset(void *dst, int value, size_t bytes, vitcu_stream s);

/* ---- CUDA graphs (the 12-layer launch chain is captured once and replayed) */
int vitcu_graph_begin(vitcu_stream s);
int vitcu_graph_end(vitcu_stream s, vitcu_graph *g);
int vitcu_graph_launch(vitcu_graph g, vitcu_stream s);
int vitcu_graph_destroy(vitcu_graph g);

/* ---- counters: kernels launched by this layer since the last reset */
void vitcu_launch_count_reset(void);
unsigned long long vitcu_launch_count(void);
/* launches of one kernel family since the last reset, by kernel name: "gemm_bf16_tc2_kernel" (CTA-pair GEMM),
 * "gemm_bf16_tc_kernel", "attention_tc_kernel", "attention_duo_tc_kernel", "attention_flash_tc_kernel", "attention_simt_kernel",
 * "sgemm_kernel", "layernorm_kernel", "patch_embed_tc_kernel".  A launch captured into a graph counts once. */
unsigned long long vitcu_launch_count_of(const char *kernel);
/* device-side watchdog flag set by a kernel whose mbarrier wait timed out */
int vitcu_watchdog_check(void);

/* ================= kernels ================================================= */

/* fp32 -> bf16 (round-to-nearest-even), n elements.  One-time weight packing. */
int vitcu_f32_to_bf16(const float *src, vitcu_bf16 *dst, size_t n, vitcu_stream s);

/* fp32 [rows,K] (row stride ld) -> bf16 [rows, 3K] = [x1 | x2 | x3] with x = x1 + x2 + x3 to 24
 * mantissa bits: the operand format of vitcu_gemm_bf16x3 (FP32 path on the tensor cores). */
int vitcu_split3(const float *x, size_t ld, vitcu_bf16 *out, size_t rows, int K, vitcu_stream s);
/* same with the exact-erf GELU (R/ViT_seq.c:283-286) applied first: out = split3(gelu(x)).  The fc1 output of the
 * accumulate-mode chain (vitcu_gemm_desc.accumulate) is complete only when every K slice has landed, so its GELU moves
 * from the GEMM epilogue into the split pass in front of fc2. */
int vitcu_split3_gelu(const float *x, size_t ld, vitcu_bf16 *out, size_t rows, int K, vitcu_stream s);

/* Patch gather (replaces the data movement of conv2d_kernel, R/conv2d.cl:1-36):
 * images [B,3,img,img] fp32 -> patches [B*P, 768] with column order (c,kh,kw),
 * the order Conv2d_seq accumulates in (R/ViT_seq.c:37-48).  out_bf16 selects
 * bf16 or fp32 output. */
int vitcu_patch_gather(const float *images, void *patches, int batch, int img,
                       int out_bf16, vitcu_stream s);
/* same with a patch side of 16 or 32: patches [B*P, 3*patch*patch] (the /32 variants of the model) */
int vitcu_patch_gather_ex(const float *images, void *patches, int batch, int img, int patch,
                          int out_bf16, vitcu_stream s);

/* Patch embedding as one TF32 tensor-core GEMM whose patch gather is staged by TMA (5-D tensor map
 * over the NCHW image; replaces conv2d_kernel + postprocess, R/conv2d.cl:1-80): for every image b
 * and patch p, x[b*T + 1 + p, :] = conv(patch) + conv_b + pos[1 + p, :].  images [B,3,S,S] fp32,
 * conv_w [768, 3*16*16] fp32.  Class-token rows are left to vitcu_cls_rows. */
int vitcu_patch_embed_tc(const float *images, const float *conv_w, const float *conv_b, const float *pos,
                         float *x, int batch, int img, vitcu_stream s);
/* same for an embedding width other than 768 (a multiple of 256; 16x16 patches) */
int vitcu_patch_embed_tc_ex(const float *images, const float *conv_w, const float *conv_b, const float *pos,
                            float *x, int batch, int img, int embed, vitcu_stream s);

/* Accumulate form for a handful of images (batch-1 latency): the token rows of x already hold the position embedding
 * (vitcu_token_rows_init) and the kernel ADDS conv(patch) + conv_b, with K cut into as many slices as fill the SMs
 * (one 224 x 224 image: 96 work items instead of 6).  Same TF32 arithmetic; the slices add in arrival order. */
int vitcu_patch_embed_tc_acc(const float *images, const float *conv_w, const float *conv_b, const float *pos,
                             float *x, int batch, int img, int embed, vitcu_stream s);

/* Row 0 of every image: x[b*T + 0, :] = cls + pos[0, :]  (R/conv2d.cl:39-80 t==0
 * branch; R/ViT_seq.c:83-118). */
int vitcu_cls_rows(float *x, const float *cls, const float *pos, int batch, int tokens,
                   vitcu_stream s);
/* same for rows of `cols` features (a multiple of 4, at most 1024) */
int vitcu_cls_rows_ex(float *x, const float *cls, const float *pos, int batch, int tokens, int cols,
                      vitcu_stream s);

/* Every token row of every image before an ACCUMULATE-mode patch embedding (vitcu_gemm_desc.accumulate; batch-1
 * latency): x[b*T + 0, :] = cls + pos[0, :] and x[b*T + t, :] = pos[t, :] for t > 0; the GEMM then adds conv(patch) +
 * conv_b into rows 1.. of the image.  Same result as the EPI_PATCH_EMBED epilogue + vitcu_cls_rows. */
int vitcu_token_rows_init(float *x, const float *cls, const float *pos, int batch, int tokens, int cols,
                          vitcu_stream s);

/* LayerNorm over 768 features (replaces layerNorm, R/layer_norm.cl:3-53; oracle
 * R/ViT_seq.c:120-142).  rows = number of rows normalised; row r is read at
 * x + r*x_row_stride (elements) so the final LN can visit only the class-token
 * rows; output rows are dense: y_bf16 = 0 fp32 [rows,768], 1 bf16 [rows,768], 2 three bf16
 * pieces [rows,3*768] (see vitcu_split3). */
int vitcu_layernorm(const float *x, size_t x_row_stride, void *y, int y_bf16,
                    const float *gamma, const float *beta, int rows, vitcu_stream s);
/* same over `cols` features: 384, 768 or 1024 (embed_dim of the model variants, R/ViT_seq.c:14) */
int vitcu_layernorm_ex(const float *x, size_t x_row_stride, void *y, int y_bf16, const float *gamma,
                       const float *beta, int rows, int cols, vitcu_stream s);
/* same, and the launch also zero-fills zero_bytes (a multiple of 16) at zero_ptr (16-byte aligned): the output buffer
 * of the accumulate-mode GEMM that consumes y (vitcu_gemm_desc.accumulate) without a launch of its own */
int vitcu_layernorm_zero(const float *x, size_t x_row_stride, void *y, int y_bf16, const float *gamma,
                         const float *beta, int rows, int cols, void *zero_ptr, size_t zero_bytes, vitcu_stream s);

/* Epilogue selector for both GEMM families */
enum {
    VITCU_EPI_BIAS = 0,          /* y = acc + b                               */
    VITCU_EPI_BIAS_GELU = 1,     /* y = gelu_erf(acc + b)   (R/ll.cl:3-5)      */
    VITCU_EPI_BIAS_RESIDUAL = 2, /* y = acc + b + residual  (R/layer_norm.cl:55) */
    VITCU_EPI_PATCH_EMBED = 3    /* row r -> (r/P)*T + 1 + r%P, y = acc + b + pos */
};

typedef struct {
    int M, N, K;          /* C[M,N] = A[M,K] * W[N,K]^T                         */
    size_t lda;           /* A row stride in elements (>= K)                    */
    int epilogue;         /* VITCU_EPI_*                                        */
    const float *bias;    /* [N]                                                */
    const float *residual;/* [M,N] fp32, may alias out (EPI_BIAS_RESIDUAL)      */
    const float *pos;     /* [T,N] fp32 (EPI_PATCH_EMBED)                       */
    int patches, tokens;  /* P and T (EPI_PATCH_EMBED)                          */
    int out_bf16;         /* output element type: 0 fp32, 1 bf16                */
    size_t ldc;           /* output row stride in elements                      */
    /* LayerNorm folded into the GEMM (BF16 path; 0 / NULL = off).  Consumer (the GEMM that follows a
     * LayerNorm, R/ViT_opencl.c:718,736): A holds bf16(x), the un-normalised fp32 rows; W holds
     * bf16(gamma * W) and bias holds b + W beta (vitcu_ln_fold_weights); the epilogue computes
     * y = rstd * acc - rstd * mean * ln_colsum[n] + bias[n] with mean / rstd of row r from the partial
     * sums ln_stats[slot][r] = (sum x, sum x^2), slot = 0 .. ln_slots-1 (float2, [ln_slots][M]). */
    const void *ln_stats;
    int ln_slots;
    const float *ln_colsum; /* [N] column sums of the folded bf16 weight                */
    /* Producer (EPI_BIAS_RESIDUAL in place, CTA-pair kernel only: vitcu_gemm_bf16_emit_supported): besides
     * x += acc + bias the epilogue writes bf16(x) to emit_bf16 [M,N] and the partial sums of every row
     * over each block of 128 columns to emit_stats (float2, [N / 128][M]). */
    void *emit_bf16;
    void *emit_stats;
    /* FP8 (E4M3) path, per-tensor scales (vitcu_gemm_e4m3 and the emit epilogue of vitcu_gemm_bf16):
     *   acc_scale  = 1 / (scale_A * scale_W): multiplies the accumulator of an e4m3 x e4m3 product
     *   out_fp8    : the output is e4m3(y * out_scale) [M,N] (the A operand of the next e4m3 GEMM)
     *   emit_fp8   : the emit epilogue writes e4m3(x * emit_scale) to emit_bf16 instead of bf16(x) */
    float acc_scale;
    int out_fp8;
    float out_scale;
    int emit_fp8;
    float emit_scale;
    /* accumulate (FP32 path at small M, the batch-1 latency shape): C [M,N] fp32 has been ZEROED by the caller and the
     * GEMM adds acc + bias into it through TMA reduce-add (EPI_BIAS only).  With no value to finish per element the K
     * range can be cut into slices that run on different SMs (split-K, chosen so that the work items fill the device):
     * a [197 x 768] x [768 x 2304] product is 36 tiles, 36 of 148 SMs busy, unless sliced. */
    int accumulate;
} vitcu_gemm_desc;

/* FP32 SIMT GEMM (replaces linear_layer, R/ll.cl:7-70 and QKV, R/multihead.cl:3-63
 * on the FP32 path; oracle R/ViT_seq.c:295-309).  A and W fp32. */
int vitcu_sgemm(const float *A, const float *W, void *C, const vitcu_gemm_desc *d,
                vitcu_stream s);

/* BF16 tensor-core GEMM: tcgen05.mma with TMEM accumulators, TMA-staged
 * operands, fused epilogue (same computation, BF16 inputs, FP32 accumulate).
 * A [M,K] bf16 (lda == K), W [N,K] bf16.  Requires K % 64 == 0, N % 16 == 0. */
int vitcu_gemm_bf16(const vitcu_bf16 *A, const vitcu_bf16 *W, void *C,
                    const vitcu_gemm_desc *d, vitcu_stream s);

/* FP8 tensor-core GEMM (tcgen05.mma kind::f8f6f4, E4M3 x E4M3 -> FP32): A [M,K] and W [N,K] are e4m3 bytes quantised
 * with per-tensor scales, d->acc_scale = 1 / (scale_A * scale_W).  CTA-pair kernel only (vitcu_gemm_bf16_emit_supported
 * (M, N) must hold, K % 128 == 0).  Two forms: (a) LayerNorm-folded consumer (ln_stats / ln_colsum, bias or GELU
 * epilogue) writing e4m3 (out_fp8) or bf16; (b) in-place fp32 residual update that emits bf16(x) or e4m3(x) and the row
 * sums (emit_bf16 / emit_stats), like the bf16 kernel's producer epilogue. */
int vitcu_gemm_e4m3(const uint8_t *A, const uint8_t *W, void *C, const vitcu_gemm_desc *d, vitcu_stream s);

/* Per-tensor absolute maximum into *out (device float, must be zeroed by the caller; the kernel max-combines):
 * of an fp32 matrix [rows,K], optionally with column factors gamma[K] (the folded weight gamma * W), or of a bf16 buffer. */
int vitcu_absmax_f32(const float *x, const float *gamma, size_t rows, int K, float *out, vitcu_stream s);
int vitcu_absmax_bf16(const vitcu_bf16 *x, size_t n, float *out, vitcu_stream s);

/* fp32 weights -> e4m3 with one scale: q[n,k] = e4m3(W[n,k] * gamma[k] * scale) (gamma NULL = 1).  When colsum is given:
 * colsum[n] = sum_k deq(q[n,k]) / scale and bias_folded[n] = bias[n] + sum_k beta[k] W[n,k] (LayerNorm fold, see
 * vitcu_ln_fold_weights). */
int vitcu_fp8_quant_weights(const float *W, const float *gamma, const float *beta, const float *bias, float scale, uint8_t *q,
                            float *colsum, float *bias_folded, int N, int K, vitcu_stream s);

/* 1 when an [M,N] product would leave SMs idle without split-K (fewer 128 x 128 tiles than half the SMs): the shapes
 * for which the engine's FP32 chain uses accumulate mode */
int vitcu_gemm_split_k_pays(int M, int N);

/* 1 when vitcu_gemm_bf16 can run the LayerNorm-producer epilogue (emit_bf16 / emit_stats) for an [M,N] output */
int vitcu_gemm_bf16_emit_supported(int M, int N);

/* One-time weight folding for a GEMM that follows a LayerNorm (gamma, beta over the K input features):
 *   w_folded[n,k] = bf16(gamma[k] * W[n,k]),  colsum[n] = sum_k w_folded[n,k],  bias_folded[n] = bias[n] + sum_k beta[k] W[n,k]
 * so that LN(x) W^T + bias = rstd * (x w_folded^T) - rstd * mean * colsum + bias_folded. */
int vitcu_ln_fold_weights(const float *W, const float *gamma, const float *beta, const float *bias, vitcu_bf16 *w_folded,
                          float *colsum, float *bias_folded, int N, int K, vitcu_stream s);

/* Entry of the folded-LayerNorm chain (the rows the patch embedding wrote): xb = bf16(x) and the row sums
 * (sum x, sum x^2) into slot 0 of stats (float2, [slots][rows]; the other slots are zeroed).  cols % 128 == 0, <= 1024. */
int vitcu_rowstats_cast(const float *x, vitcu_bf16 *xb, void *stats, int rows, int cols, int slots, vitcu_stream s);

/* FP32-accurate GEMM on the BF16 tensor cores: A3 [M,3K] and W3 [N,3K] hold the three bf16 pieces
 * of the fp32 operands (vitcu_split3); C = sum of the six significant piece products, FP32
 * accumulation in TMEM, relative error ~1e-7.  d->K is the logical K (K % 64 == 0); epilogues as
 * above, GELU evaluated with erff when the output is fp32. */
int vitcu_gemm_bf16x3(const vitcu_bf16 *A3, const vitcu_bf16 *W3, void *C, const vitcu_gemm_desc *d,
                      vitcu_stream s);

/* Multi-head attention core over a fused QKV buffer (replaces QKV_TO_SCOREV,
 * R/multihead.cl:65-137; oracle R/ViT_seq.c:192-262): qkv [B*T, 2304] with
 * Q|K|V column blocks of 768 (R/ViT_seq.c:150), 12 heads of 64; scores are
 * scaled by 1/sqrt(64) after the dot product; out [B*T,768].  Element type of
 * qkv/out: fp32 (is_bf16 = 0) or bf16 (is_bf16 = 1; softmax stays fp32). */
int vitcu_attention(const void *qkv, void *out, int batch, int tokens, int is_bf16,
                    vitcu_stream s);
/* same with `heads` heads of 64 (num_heads, R/ViT_seq.c:16): qkv [B*T, 3*heads*64], out [B*T, heads*64].
 * is_bf16 = 2: fp32 qkv, fp32 math, output as three bf16 pieces [B*T, 3*heads*64] (the A operand of
 * vitcu_gemm_bf16x3, see vitcu_split3) */
int vitcu_attention_ex(const void *qkv, void *out, int batch, int tokens, int heads, int is_bf16,
                       vitcu_stream s);

/* Debug aid for the single-block attention kernel: record clock64 stamps of CTA 0's first 16 units
 * into `buffer` (device memory, 4*16*8 uint64); NULL switches it off.  See tools/attn_timeline.py. */
int vitcu_attention_debug_timeline(unsigned long long *buffer);

/* Row softmax over `n` logits per row (replaces softMax, R/miniSoftMax.cl:1-50;
 * oracle R/ViT_seq.c:372-397). */
int vitcu_softmax_rows(const float *logits, float *probs, int rows, int n, vitcu_stream s);

/* The k largest entries of every row, largest first, ties to the lower index: idx/val [rows,k]
 * (device).  k = 1 is the argmax scan of R/Main.c:59-72 (first maximum wins) without its stale
 * pred_idx carry-over between images. */
int vitcu_topk_rows(const float *x, int rows, int cols, int k, int *idx, float *val, vitcu_stream s);

#ifdef __cplusplus
}
#endif
#endif
