/* include/vit_b200.h -- public C interface of the B200 ViT-B/16 inference engine.
 *
 * R/ = /root/reference/MulticoreMainProject/.
 *
 * 1. The drop-in entry point.  libvit_b200.so exports the symbol the reference
 *    declares in R/ViT_opencl.h:6 and calls from R/Main.c:54:
 *
 *        void ViT_opencl(ImageData *image, Network *networks, float **prb);
 *
 *    so the reference's unmodified Main.c, Network.c and comparator.c link
 *    against this library instead of ViT_opencl.c + kernelHandler.c + the five
 *    .cl files (INTEGRATION.md shows the link line).  Semantics kept from the
 *    reference: n = image->n images of c*h*w floats each, 152 weight blobs,
 *    caller-allocated prb[i][0..999] receives softmax PROBABILITIES; the call
 *    is synchronous; all device state is released before returning; any
 *    failure prints "[file:line] CUDA error N (...)" and exit(EXIT_FAILURE)s,
 *    the CHECK_ERROR convention of R/kernelHandler.h:6-10.  Differences, all
 *    deliberate: results are complete on return (the reference relied on an
 *    implicit drain, R/ViT_opencl.c:978-983); there is no 100-image cap
 *    (R/ViT_opencl.c:107-111); image side is read from the struct so 384x384
 *    works; missing blobs are reported instead of dereferenced.
 *    Environment knobs: VITB200_PRECISION=fp32|bf16|fp8 (default fp32, the
 *    reference's arithmetic), VITB200_GPUS=<n> (default: 1 GPU per 256 images,
 *    capped by the visible devices), VITB200_BATCH=<images per chunk>,
 *    VITB200_PERSIST=1 (keep the engines and the packed weights across calls;
 *    weights are re-uploaded only when their signature changes; released by
 *    vitb200_release_persistent()).
 *
 * 2. The engine API underneath it, for callers that keep the model resident
 *    (the bench, the tests, a serving loop).  The structs mirror the
 *    reference's so its own loaders can feed them.
 */
#ifndef VIT_B200_H
#define VIT_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Layout-identical mirrors of R/Network.h:7-14 and :19-23.  When this header is
 * used together with the reference's Network.h, define VITB200_USE_REFERENCE_TYPES
 * first and the reference's own typedefs are used instead. */
#ifndef VITB200_USE_REFERENCE_TYPES
typedef struct {
    int n, c, h, w;
    float *data;
} vitb200_image; /* == ImageData */
typedef struct {
    float *data;
    size_t size;
} vitb200_blob; /* == Network */
#else
typedef ImageData vitb200_image;
typedef Network vitb200_blob;
#endif

#define VITB200_NBLOBS 152
#define VITB200_CLASSES 1000

/* VITB200_FP8: the BF16 path with the two MLP GEMMs (fc1, fc2: 63.5 % of the FLOPs) on E4M3 operands
 * (tcgen05.mma kind::f8f6f4, per-tensor scales; activations calibrated on the engine's first chunk).  It applies to
 * chunks large enough for the CTA-pair GEMM (>= 32 images at 224x224); smaller chunks run the BF16 kernels.
 * Accuracy contract: INTEGRATION.md section "FP8". */
enum { VITB200_FP32 = 0, VITB200_BF16 = 1, VITB200_FP8 = 2 };

typedef struct vitb200_engine vitb200_engine;

/* the reference's entry point (R/ViT_opencl.h:6) */
void ViT_opencl(vitb200_image *image, vitb200_blob *networks, float **prb);

/* The image split ViT_opencl uses over several GPUs: contiguous shards first[g] .. first[g] + count[g]
 * of the n images (arrays of at least `gpus` entries); returns the number of shards (<= gpus). */
int vitb200_shard_plan(int n, int gpus, int *first, int *count);

/* The chunk schedule of one forward call over n images on an engine whose full chunk is `chunk` images: size of the
 * next chunk when `done` images have been issued.  Full chunks, except that (a) with head_split the call opens with
 * a quarter chunk -- nothing can be computed before the first chunk has arrived -- and (b) with tail_split (pageable
 * sources) a last chunk of more than half a chunk is cut so that only a quarter chunk is left to run after the last
 * upload.  The engine enables the splits only where a quarter chunk runs the same kernels as a full one, so that an
 * image's result does not depend on the chunk it travels in. */
int vitb200_next_chunk(int n, int done, int chunk, int head_split, int tail_split);

/* what the last ViT_opencl call of this process spent where: wall time of the whole call and, per
 * phase, the slowest shard's time (the shards run concurrently, one host thread per GPU) */
typedef struct {
    int images, gpus;
    double wall_s, create_s, weights_s, forward_s;
    double teardown_s; /* releasing the engine (default semantics: every device and pinned allocation of the call is freed) */
} vitb200_call_stats;
void vitb200_last_call_stats(vitb200_call_stats *out);

/* frees the engines VITB200_PERSIST=1 kept alive (no-op otherwise) */
void vitb200_release_persistent(void);

/* ---- engine API: every int function returns 0 or an error code whose text is
 * in vitb200_last_error(). ---- */
const char *vitb200_last_error(void);
int vitb200_device_count(void);

/* One engine = one GPU, one image size, one precision.  max_batch = images per
 * forward chunk (activation buffers are sized for it).  vitb200_create builds the
 * reference's model, ViT-B/16 (R/ViT_seq.c:10-17). */
int vitb200_create(vitb200_engine **out, int device, int img, int precision, int max_batch);

/* Other members of the family with the same blob order: the values of the
 * reference's macros img_size, patch_size, embed_dim, depth, num_heads and embed_dim * mlp_ratio
 * (R/ViT_seq.c:10-17) as run-time numbers.  Supported: patch 16 or 32; embed 384, 768 or 1024 with
 * heads = embed / 64; depth 1..32; hidden a multiple of 128; img a multiple of patch.  The model
 * has 8 + 12 * depth blobs (152 for depth 12): cls, conv w [embed, 3*patch*patch], conv b,
 * pos [(img/patch)^2 + 1, embed], 12 per layer, final LN, head. */
typedef struct {
    int img, patch, embed, depth, heads, hidden;
} vitb200_model;
int vitb200_create_model(vitb200_engine **out, int device, const vitb200_model *model, int precision, int max_batch);
/* Reads the dims of a 152-blob (depth 12) model off the blob sizes: embed from the class token,
 * patch from the conv filters, hidden from the fc1 bias, heads = embed / 64; img from `image` (or,
 * when image is NULL, from the position table).  Falls back to ViT-B/16 when the sizes are
 * inconsistent, so that vitb200_load_weights names the offending blob.  ViT_opencl uses this, which
 * makes ViT-B/32 or ViT-S/16 weights work through the unchanged Main.c. */
int vitb200_model_from_blobs(const vitb200_blob *networks, const vitb200_image *image, vitb200_model *model);
void vitb200_destroy(vitb200_engine *e);

/* Upload + pack the 8 + 12 * depth blobs (152 for the reference's model; validates presence and sizes). */
int vitb200_load_weights(vitb200_engine *e, const vitb200_blob *networks);

/* Forward n images from a contiguous host array [n,3,img,img] (pinned memory
 * from vitb200_host_alloc streams fastest).  probs [n,1000] required; logits
 * [n,1000] optional (NULL to skip).  Synchronous. */
int vitb200_forward(vitb200_engine *e, const float *images_host, int n, float *probs_host,
                    float *logits_host);

/* Same, from the reference's per-image structs into per-image rows (the
 * ViT_opencl calling convention). */
int vitb200_forward_structs(vitb200_engine *e, const vitb200_image *images, int n, float **prb);

/* Labels only: the k most probable classes of every image, most probable first, ties to the lower
 * class index (k = 1 is the argmax scan of R/Main.c:59-72 without its pred_idx carried over from the
 * previous image).  labels/probs are [n,k] host arrays; the 1000-way probabilities stay on the GPU. */
#define VITB200_TOPK_MAX 8
int vitb200_forward_topk(vitb200_engine *e, const float *images_host, int n, int k, int *labels, float *probs);

/* Device-resident variant for kernel-only timing: stage n <= max_batch images
 * once, then run the forward on them any number of times. */
int vitb200_stage_images(vitb200_engine *e, const float *images_host, int n);
int vitb200_forward_resident(vitb200_engine *e, int n);
int vitb200_read_probs(vitb200_engine *e, int n, float *probs_host, float *logits_host);
/* elapsed device time of the last vitb200_forward_resident, CUDA events on the
 * engine's compute stream */
int vitb200_last_forward_ms(vitb200_engine *e, float *ms);
/* times `iters` back-to-back resident forwards with one event pair */
int vitb200_time_resident(vitb200_engine *e, int n, int iters, float *total_ms);

/* in-situ time of the dense-layer (GEMM) launches: runs `iters` eager forwards of the n staged images
 * with one CUDA-event pair around every GEMM launch on the compute stream and returns the summed GEMM
 * time per forward (bench.py's roofline: same data, cache state and clocks as the timed forward) */
int vitb200_profile_gemms(vitb200_engine *e, int n, int iters, float *gemm_ms_per_forward, int *gemm_launches);

/* whole-forward timeline: one eager forward of the n staged images with an event in front of every
 * launch; the span up to the next launch (kernel + gap) is booked under 0 = GEMM, 1 = attention,
 * 2 = LayerNorm, 3 = everything else (patch embedding, class rows, split kernels, final LN + head + softmax) */
int vitb200_profile_timeline(vitb200_engine *e, int n, float ms_by_kind[4], int launches_by_kind[4]);

/* debugging / per-stage parity: copy an internal activation of the last
 * resident forward to the host.  what: 0 = residual stream x [n*T,768] fp32
 * after the last executed stage; stop_after_layer (set before the forward)
 * limits execution: -1 = full, 0 = embedding only, L = after encoder layer L. */
int vitb200_set_stop_after_layer(vitb200_engine *e, int layer);
int vitb200_read_tokens(vitb200_engine *e, int n, float *x_host);

/* pinned host memory for image batches */
int vitb200_host_alloc(void **ptr, size_t bytes);
int vitb200_host_free(void *ptr);

/* kernels launched per forward chunk of the current configuration */
int vitb200_kernels_per_forward(const vitb200_engine *e);
int vitb200_tokens(const vitb200_engine *e);
/* embedding width of the engine's model (row length of vitb200_read_tokens) */
int vitb200_embed(const vitb200_engine *e);

#ifdef __cplusplus
}
#endif
#endif
