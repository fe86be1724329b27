/* host/vit_engine.c -- C host orchestration of the ViT-B/16 forward on one B200.
 *
 * Takes the place of the host side of R/ViT_opencl.c (R/ = /root/reference/
 * MulticoreMainProject/): weight staging (R/ViT_opencl.c:125-330), the per-image
 * kernel chain (Conv2d :361, postConv2d :407, Encoder :710, layer_norm :444,
 * linear_layer :622, Softmax :750) and read-back.  It talks to the GPU only
 * through include/vit_cuda_layer.h.  Differences in kind, not just in speed:
 * images are processed in batches (M = batch*tokens rows per GEMM) instead of
 * one by one; residual adds, bias, GELU, class-token/position adds are fused
 * into GEMM epilogues so a layer is 7 launches instead of 9 per image; the
 * whole chunk forward is captured once in a CUDA graph and replayed; image
 * upload of chunk i+1 overlaps the compute of chunk i.
 */
#include "vit_engine_internal.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static _Thread_local char g_msg[640] = "no error";

const char *vitb200_last_error(void) { return g_msg; }

int vit_fail(const char *file, int line, int rc, const char *what)
{
    const char *base = strrchr(file, '/');
    if (what)
        snprintf(g_msg, sizeof(g_msg), "[%s:%d] CUDA error %d (%s)", base ? base + 1 : file, line, rc, what);
    else
        snprintf(g_msg, sizeof(g_msg), "[%s:%d] %s", base ? base + 1 : file, line, vitcu_last_error());
    return rc;
}

int vitb200_device_count(void)
{
    int n = 0;
    if (vitcu_device_count(&n) != 0)
        return 0;
    return n;
}

int vitb200_host_alloc(void **ptr, size_t bytes)
{
    VIT_TRY(vitcu_host_alloc(ptr, bytes));
    return 0;
}
int vitb200_host_free(void *ptr)
{
    VIT_TRY(vitcu_host_free(ptr));
    return 0;
}

/* expected element count of blob idx (index map: R/ViT_seq.c:437-513; torchvision state_dict() order) */
static size_t blob_elems(const vitb200_engine *e, int idx)
{
    const size_t D = (size_t)e->D, HID = (size_t)e->HID;
    const int head0 = e->nblobs - 4;
    if (idx == 0 || idx == 2)
        return D;
    if (idx == 1)
        return D * 3 * e->patch * e->patch;
    if (idx == 3)
        return (size_t)e->T * D;
    if (idx >= 4 && idx < head0) {
        switch ((idx - 4) % 12) {
        case 2:
            return 3 * D * D;
        case 3:
            return 3 * D;
        case 4:
            return D * D;
        case 8:
            return HID * D;
        case 9:
            return HID;
        case 10:
            return D * HID;
        default:
            return D;
        }
    }
    if (idx == head0 || idx == head0 + 1)
        return D;
    if (idx == head0 + 2)
        return (size_t)VITB200_CLASSES * D;
    return VITB200_CLASSES;
}

/* GEMM weight matrices (everything else stays fp32 in both precisions) */
static int is_gemm_weight(const vitb200_engine *e, int idx)
{
    if (idx == 1)
        return 1;
    if (idx >= 4 && idx < e->nblobs - 4) {
        int k = (idx - 4) % 12;
        return k == 2 || k == 4 || k == 8 || k == 10;
    }
    return 0;
}

/* GEMM weights whose input is a LayerNorm output: in_proj (LN1) and fc1 (LN2) of every layer (R/ViT_opencl.c:718,736) */
static int follows_layernorm(const vitb200_engine *e, int idx)
{
    if (idx >= 4 && idx < e->nblobs - 4) {
        int k = (idx - 4) % 12;
        return k == 2 || k == 8;
    }
    return 0;
}

/* the two MLP matrices of a layer (fc1: base + 8, fc2: base + 10): 63.5 % of the FLOPs, the FP8 candidates */
static int is_mlp_weight(const vitb200_engine *e, int idx)
{
    if (idx >= 4 && idx < e->nblobs - 4) {
        int k = (idx - 4) % 12;
        return k == 8 || k == 10;
    }
    return 0;
}

/* K (row length) of GEMM weight idx */
static int gemm_weight_k(const vitb200_engine *e, int idx)
{
    if (idx == 1)
        return 3 * e->patch * e->patch;
    return (idx - 4) % 12 == 10 ? e->HID : e->D;
}

int vitb200_model_from_blobs(const vitb200_blob *net, const vitb200_image *image, vitb200_model *m)
{
    if (!net || !m)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "NULL argument");
    /* class token = embed_dim; conv filters = embed * 3 * patch^2; fc1 bias = hidden; 64 per head;
     * position rows = (img / patch)^2 + 1.  Anything absent or inconsistent -> the ViT-B/16 defaults,
     * so that load_weights reports exactly which blob is wrong. */
    m->img = image ? image->h : 224;
    m->patch = 16;
    m->embed = 768;
    m->depth = 12;
    m->heads = 12;
    m->hidden = 3072;
    if (!net[0].data || !net[1].data || !net[3].data || !net[13].data)
        return 0;
    const size_t D = net[0].size;
    if (D == 0 || D % 64 != 0 || net[1].size % (3 * D) != 0)
        return 0;
    const size_t pp = net[1].size / (3 * D);
    int patch = 0;
    while ((size_t)patch * patch < pp)
        patch++;
    if ((size_t)patch * patch != pp || net[3].size % D != 0 || net[13].size == 0)
        return 0;
    m->patch = patch;
    m->embed = (int)D;
    m->heads = (int)(D / 64);
    m->hidden = (int)net[13].size;
    if (!image) { /* image side from the position table */
        const size_t P = net[3].size / D - 1;
        int side = 0;
        while ((size_t)side * side < P)
            side++;
        if ((size_t)side * side == P)
            m->img = side * patch;
    }
    return 0;
}

int vitb200_create(vitb200_engine **out, int device, int img, int precision, int max_batch)
{
    const vitb200_model m = {img, 16, 768, 12, 12, 3072}; /* ViT-B/16, the reference's macros (R/ViT_seq.c:10-17) */
    return vitb200_create_model(out, device, &m, precision, max_batch);
}

int vitb200_create_model(vitb200_engine **out, int device, const vitb200_model *m, int precision, int max_batch)
{
    if (!out)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "engine out-pointer is NULL");
    *out = NULL;
    if (!m)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "model description is NULL");
    const int img = m->img;
    if (m->patch != 16 && m->patch != 32)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "patch side must be 16 or 32");
    if (img <= 0 || img % m->patch != 0)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "image side must be a positive multiple of 16");
    if (m->embed != 384 && m->embed != 768 && m->embed != 1024)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "embedding width must be 384, 768 or 1024");
    if (m->heads * 64 != m->embed)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "heads must be embed / 64 (head dimension 64)");
    if (m->depth < 1 || m->depth > VIT_MAX_DEPTH)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "depth must be 1..32");
    if (m->hidden <= 0 || m->hidden % 128 != 0)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "MLP width must be a positive multiple of 128");
    if (precision != VITB200_FP32 && precision != VITB200_BF16 && precision != VITB200_FP8)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "precision must be VITB200_FP32, VITB200_BF16 or VITB200_FP8");
    if (precision == VITB200_FP8 && (m->embed % 256 != 0 || m->hidden % 256 != 0))
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "FP8 precision needs embed and MLP widths that are multiples of 256");
    if (max_batch <= 0)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "max_batch must be positive");
    int ndev = 0;
    VIT_TRY(vitcu_device_count(&ndev));
    if (device < 0 || device >= ndev)
        return vit_fail(__FILE__, __LINE__, VITCU_E_NODEVICE, "no such CUDA device");
    int cc = 0;
    VIT_TRY(vitcu_device_info(device, NULL, NULL, &cc, NULL));
    if (cc < 100)
        return vit_fail(__FILE__, __LINE__, VITCU_E_NODEVICE, "device is not sm_100 (B200) class");
    VIT_TRY(vitcu_set_device(device));
    VIT_TRY(vitcu_prepare_device());

    vitb200_engine *e = (vitb200_engine *)calloc(1, sizeof(*e));
    if (!e)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "out of host memory");
    e->device = device;
    e->img = img;
    e->patch = m->patch;
    e->D = m->embed;
    e->HID = m->hidden;
    e->depth = m->depth;
    e->heads = m->heads;
    e->nblobs = 8 + 12 * m->depth;
    e->side = img / m->patch;
    e->P = e->side * e->side;
    e->T = e->P + 1;
    e->precision = precision;
    e->B = max_batch;
    e->stop_after = -1;
    e->no_graph = getenv("VITB200_NO_GRAPH") != NULL;
    /* FP32 precision runs its GEMMs on the tensor cores as split-bf16 (error ~1e-7 relative);
     * VITB200_FP32_SIMT=1 selects the CUDA-core FFMA GEMM instead */
    e->fp32_tc = precision == VITB200_FP32 && getenv("VITB200_FP32_SIMT") == NULL;
    e->fp32_splitk = !(getenv("VITB200_FP32_SPLITK") && atoi(getenv("VITB200_FP32_SPLITK")) == 0);
    e->pe_splitk = !(getenv("VITB200_PE_SPLITK") && atoi(getenv("VITB200_PE_SPLITK")) == 0);
    /* VITB200_PE_GATHER=1: separate gather kernel + BF16 GEMM instead of the TMA-gather TF32 GEMM */
    e->pe_gather = getenv("VITB200_PE_GATHER") != NULL || e->patch != 16 || e->D % 256 != 0;
    const int bf = precision != VITB200_FP32; /* FP8 = the BF16 path with fc1 / fc2 on e4m3 operands */
    e->fp8 = precision == VITB200_FP8;
    /* LayerNorm folded into the qkv / fc1 GEMMs and produced by the out-proj / fc2 epilogues: BF16 path, widths the
     * CTA-pair GEMM tiles (VITB200_LN_FOLD=0 keeps the separate LayerNorm kernel) */
    e->ln_fold = bf && e->D % 256 == 0 && (e->fp8 || !(getenv("VITB200_LN_FOLD") && atoi(getenv("VITB200_LN_FOLD")) == 0));
    const size_t act = bf ? 2 : 4;
    const size_t rows = (size_t)e->B * e->T;
    const size_t img_elems = (size_t)3 * img * img;

    int rc = 0;
#define ENG_TRY(x)                                                              \
    do {                                                                        \
        rc = (x);                                                               \
        if (rc) {                                                               \
            vit_fail(__FILE__, __LINE__, rc, NULL);                             \
            vitb200_destroy(e);                                                 \
            return rc;                                                          \
        }                                                                       \
    } while (0)
    ENG_TRY(vitcu_stream_create(&e->stream));
    ENG_TRY(vitcu_stream_create(&e->copy_stream));
    for (int i = 0; i < 2; i++) {
        ENG_TRY(vitcu_event_create(&e->ev_h2d[i]));
        ENG_TRY(vitcu_event_create(&e->ev_done[i]));
        ENG_TRY(vitcu_event_create(&e->ev_out[i]));
    }
    ENG_TRY(vitcu_event_create(&e->ev_t0));
    ENG_TRY(vitcu_event_create(&e->ev_t1));
    /* ONE device allocation for every activation buffer (sixteen cudaMalloc / cudaFree pairs otherwise: the default
     * ViT_opencl call creates and releases an engine per call, and the driver-side cost of those calls is what makes a
     * cold call's time jump, profiles/r02_batch1_latency.md); 256-byte aligned pieces (TMA tensor maps need 16) */
    {
        size_t sz[16], off[16], total = 0;
        int k = 0;
        const size_t a3_elems = (size_t)(e->HID > 3 * e->patch * e->patch ? e->HID : 3 * e->patch * e->patch);
        sz[k++] = (size_t)e->B * img_elems * sizeof(float);                               /* 0, 1: d_images */
        sz[k++] = (size_t)e->B * img_elems * sizeof(float);
        sz[k++] = (size_t)e->B * VITB200_CLASSES * sizeof(float);                         /* 2, 3: d_probs */
        sz[k++] = (size_t)e->B * VITB200_CLASSES * sizeof(float);
        sz[k++] = (size_t)e->B * VITB200_CLASSES * sizeof(float);                         /* 4, 5: d_logits */
        sz[k++] = (size_t)e->B * VITB200_CLASSES * sizeof(float);
        sz[k++] = (size_t)e->B * e->P * 3 * e->patch * e->patch * act;                    /* 6: d_patches */
        sz[k++] = rows * e->D * sizeof(float);                                            /* 7: d_x */
        sz[k++] = rows * e->D * (e->fp32_tc ? 6 : act);                                   /* 8: d_ln */
        sz[k++] = e->fp32_tc ? rows * a3_elems * 3 * sizeof(vitcu_bf16) : 0;              /* 9: d_a3 */
        sz[k++] = rows * 3 * e->D * act;                                                  /* 10: d_qkv */
        sz[k++] = rows * e->D * (e->fp32_tc ? 6 : act);                                   /* 11: d_att */
        sz[k++] = rows * e->HID * act;                                                    /* 12: d_hid */
        sz[k++] = (size_t)e->B * e->D * sizeof(float);                                    /* 13: d_cls */
        sz[k++] = e->ln_fold ? rows * (size_t)(e->D / 128) * 2 * sizeof(float) : 0;       /* 14: d_lnstats */
        sz[k++] = e->fp8 ? (size_t)2 * VIT_MAX_DEPTH * sizeof(float) : 0;                 /* 15: d_amax */
        for (int i = 0; i < k; i++) {
            off[i] = total;
            total += (sz[i] + 255) & ~(size_t)255;
        }
        ENG_TRY(vitcu_malloc(&e->d_arena, total));
        char *base = (char *)e->d_arena;
#define PIECE(i) (sz[i] ? (void *)(base + off[i]) : NULL)
        e->d_images[0] = (float *)PIECE(0);
        e->d_images[1] = (float *)PIECE(1);
        e->d_probs[0] = (float *)PIECE(2);
        e->d_probs[1] = (float *)PIECE(3);
        e->d_logits[0] = (float *)PIECE(4);
        e->d_logits[1] = (float *)PIECE(5);
        e->d_patches = PIECE(6);
        e->d_x = (float *)PIECE(7);
        e->d_ln = PIECE(8);
        e->d_a3 = PIECE(9);
        e->d_qkv = PIECE(10);
        e->d_att = PIECE(11);
        e->d_hid = PIECE(12);
        e->d_cls = (float *)PIECE(13);
        e->d_lnstats = PIECE(14);
        e->d_amax = (float *)PIECE(15);
#undef PIECE
    }
    ENG_TRY(vitcu_host_alloc((void **)&e->h_probs, (size_t)e->B * VITB200_CLASSES * sizeof(float) * 2));
    ENG_TRY(vitcu_host_alloc((void **)&e->h_logits, (size_t)e->B * VITB200_CLASSES * sizeof(float) * 2));
#undef ENG_TRY
    *out = e;
    return 0;
}

static double destroy_now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

void vitb200_destroy(vitb200_engine *e)
{
    if (!e)
        return;
    /* VITB200_DEBUG_TEARDOWN=1: where the tear-down spends its time (device sync / device frees / pinned frees /
     * copy threads / events, streams, graphs) */
    const int dbg = getenv("VITB200_DEBUG_TEARDOWN") != NULL;
    double t[6];
    t[0] = destroy_now();
    vitcu_set_device(e->device);
    vitcu_device_sync();
    t[1] = destroy_now();
    vitcu_free(e->w_arena); /* every w32[] / w16[] pointer lives in it */
    vitcu_free(e->d_arena); /* ... and every activation buffer in this one */
    vitcu_free(e->d_topi);
    vitcu_free(e->d_topv);
    t[2] = destroy_now();
    vitcu_host_free(e->h_probs);
    vitcu_host_free(e->h_logits);
    vitcu_host_free(e->h_topi);
    vitcu_host_free(e->h_topv);
    vitcu_host_free(e->h_stage);
    t[3] = destroy_now();
    vit_stager_destroy(e->stager);
    t[4] = destroy_now();
    for (int i = 0; i < 2; i++) {
        if (e->ev_h2d[i])
            vitcu_event_destroy(e->ev_h2d[i]);
        if (e->ev_done[i])
            vitcu_event_destroy(e->ev_done[i]);
        if (e->ev_out[i])
            vitcu_event_destroy(e->ev_out[i]);
        if (e->graph[i])
            vitcu_graph_destroy(e->graph[i]);
    }
    if (e->ev_t0)
        vitcu_event_destroy(e->ev_t0);
    if (e->ev_t1)
        vitcu_event_destroy(e->ev_t1);
    for (int i = 0; i < VIT_STAGE_SLOTS; i++)
        if (e->ev_slot[i])
            vitcu_event_destroy(e->ev_slot[i]);
    if (e->stream)
        vitcu_stream_destroy(e->stream);
    if (e->copy_stream)
        vitcu_stream_destroy(e->copy_stream);
    t[5] = destroy_now();
    if (dbg)
        printf("ViT_b200 tear-down: sync %.4f, device frees %.4f, pinned frees %.4f, copy threads %.4f, events / streams / graphs %.4f s\n",
               t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4]);
    free(e);
}

/* The pinned staging ring (VIT_STAGE_SLOTS slots of stage_group images) and the copy threads, created on first
 * use: by the first pageable image upload or by the first weight upload from pageable blobs. */
static int ensure_stage(vitb200_engine *e)
{
    if (e->stager)
        return 0;
    const size_t img_bytes = (size_t)3 * e->img * e->img * sizeof(float);
    /* small slots: page-locking costs ~0.4 ms per MB and is paid inside the first call */
    const char *mb = getenv("VITB200_STAGE_SLOT_MB");
    e->stage_group = (int)(((size_t)(mb && atoi(mb) > 0 ? atoi(mb) : 4) << 20) / img_bytes);
    if (e->stage_group < 1)
        e->stage_group = 1;
    if (e->stage_group > e->B)
        e->stage_group = e->B;
    /* a retry after a partial failure reuses what the first attempt got (nothing is allocated twice) */
    if (!e->h_stage)
        VIT_TRY(vitcu_host_alloc((void **)&e->h_stage, (size_t)VIT_STAGE_SLOTS * e->stage_group * img_bytes));
    for (int i = 0; i < VIT_STAGE_SLOTS; i++)
        if (!e->ev_slot[i])
            VIT_TRY(vitcu_event_create(&e->ev_slot[i]));
    e->stager = vit_stager_create(vit_stager_threads_default());
    if (!e->stager)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "out of host memory");
    return 0;
}

/* One weight blob to the device.  The reference hands over 152 malloc'd (pageable) blobs (R/Network.c:134-215):
 * a cudaMemcpy from pageable memory is staged by the driver on the calling thread at a few GB/s and does not
 * return before it is done, which made the 346 MB upload the longest part of a cold ViT_opencl call.  Here the
 * copy threads move the blob piece by piece into the pinned ring (host/vit_stage.c) and every
 * piece leaves by asynchronous DMA, so the host copy of piece i+1 overlaps the transfer of piece i.  Pinned
 * sources are DMA'd in place.  VITB200_WEIGHT_STAGE=0: plain copies (A/B). */
static int upload_blob(vitb200_engine *e, void *dst, const void *src, size_t bytes, int staged)
{
    int pinned = 0;
    if (!staged || (vitcu_host_is_pinned(src, &pinned) == 0 && pinned) || bytes < ((size_t)64 << 10))
        return vitcu_memcpy_h2d(dst, src, bytes, e->stream);
    const size_t slot_bytes = (size_t)e->stage_group * 3 * e->img * e->img * sizeof(float);
    for (size_t off = 0; off < bytes; off += slot_bytes) {
        const size_t len = bytes - off < slot_bytes ? bytes - off : slot_bytes;
        const int slot = e->stage_next++ % VIT_STAGE_SLOTS;
        char *h = e->h_stage + (size_t)slot * slot_bytes;
        if (e->stage_used[slot]) {
            int rc = vitcu_event_sync(e->ev_slot[slot]);
            if (rc)
                return rc;
        }
        vit_stager_copy(e->stager, h, NULL, (const char *)src + off, len, 1);
        int rc = vitcu_memcpy_h2d((char *)dst + off, h, len, e->stream);
        if (!rc)
            rc = vitcu_event_record(e->ev_slot[slot], e->stream);
        if (rc)
            return rc;
        e->stage_used[slot] = 1;
    }
    return 0;
}

int vitb200_load_weights(vitb200_engine *e, const vitb200_blob *net)
{
    if (!e || !net)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "NULL argument");
    VIT_TRY(vitcu_set_device(e->device));
    /* validate first: the reference would dereference a NULL blob (R/Network.c:144-148
     * leaves absent files as {NULL,0}) */
    for (int i = 0; i < e->nblobs; i++) {
        if (!net[i].data || net[i].size == 0) {
            char what[96];
            snprintf(what, sizeof(what), "weight blob %d is missing", i);
            return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, what);
        }
        if (net[i].size != blob_elems(e, i)) {
            char what[128];
            snprintf(what, sizeof(what), "weight blob %d has %zu elements, expected %zu", i, net[i].size,
                     blob_elems(e, i));
            return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, what);
        }
    }
    const int bf = e->precision != VITB200_FP32;
    /* One device allocation for all weights (152 cudaMalloc + 49 cudaFree calls cost more than the
     * upload itself) and two fp32 scratch buffers through which the GEMM weights pass on their way
     * to the packed form: bf16 [N,K] on the BF16 path, three bf16 pieces [N,3K] on the FP32
     * tensor-core path.  Everything else (biases, LayerNorm, class token, position embedding, head,
     * and the conv filters of the TF32 patch embedding) stays fp32. */
    size_t off32[VIT_MAX_BLOBS], off16[VIT_MAX_BLOBS], offf[VIT_MAX_BLOBS], off8[VIT_MAX_BLOBS], total = 0, scratch_elems = 0;
    for (int i = 0; i < e->nblobs; i++) {
        const size_t n = net[i].size;
        const int packed = is_gemm_weight(e, i) && (bf || e->fp32_tc);
        offf[i] = off8[i] = (size_t)-1;
        if (e->fp8 && is_mlp_weight(e, i)) { /* e4m3 copy [N,K] (+ colsum [N] | bias' [N] for the folded fc1) */
            const size_t N = n / (size_t)gemm_weight_k(e, i);
            off8[i] = total;
            total += ((n + 255) & ~(size_t)255) + 2 * ((N * sizeof(float) + 255) & ~(size_t)255);
        }
        if (e->ln_fold && follows_layernorm(e, i)) { /* folded copy: bf16 [N,K] | colsum [N] | bias' [N] */
            const size_t N = n / (size_t)e->D;
            offf[i] = total;
            total += ((n * sizeof(vitcu_bf16) + 255) & ~(size_t)255) + 2 * ((N * sizeof(float) + 255) & ~(size_t)255);
        }
        const int keep32 = !packed || (bf && i == 1 && !e->pe_gather);
        const int need16 = packed && !(bf && i == 1 && !e->pe_gather);
        off32[i] = off16[i] = (size_t)-1;
        if (keep32) {
            off32[i] = total;
            total += (n * sizeof(float) + 255) & ~(size_t)255;
        }
        if (need16) {
            off16[i] = total;
            total += (n * (e->fp32_tc ? 3 : 1) * sizeof(vitcu_bf16) + 255) & ~(size_t)255;
            if (n > scratch_elems)
                scratch_elems = n;
        }
    }
    /* The new arena is built beside the old one and swapped in only when the whole upload succeeded: a
     * failed reload leaves the engine serving its previous weights.  A successful one invalidates
     * everything that holds the old arena's addresses -- the captured graphs have device pointers and
     * tensor maps baked into their kernel nodes -- so they are dropped with it. */
    void *arena = NULL;
    float *scratch[2] = {NULL, NULL};
    float **w32 = (float **)calloc(VIT_MAX_BLOBS, sizeof(float *));
    vitcu_bf16 **w16 = (vitcu_bf16 **)calloc(VIT_MAX_BLOBS, sizeof(vitcu_bf16 *));
    vitcu_bf16 **wf16 = (vitcu_bf16 **)calloc(VIT_MAX_BLOBS, sizeof(vitcu_bf16 *));
    float **wf_cs = (float **)calloc(VIT_MAX_BLOBS, sizeof(float *)), **wf_b = (float **)calloc(VIT_MAX_BLOBS, sizeof(float *));
    unsigned char **wq8 = (unsigned char **)calloc(VIT_MAX_BLOBS, sizeof(unsigned char *));
    float **wq8_cs = (float **)calloc(VIT_MAX_BLOBS, sizeof(float *)), **wq8_b = (float **)calloc(VIT_MAX_BLOBS, sizeof(float *));
    float *wq8_scale = (float *)calloc(VIT_MAX_BLOBS, sizeof(float));
    int rc = (w32 && w16 && wf16 && wf_cs && wf_b && wq8 && wq8_cs && wq8_b && wq8_scale) ? 0 : VITCU_E_ARG;
    if (rc)
        vit_fail(__FILE__, __LINE__, rc, "out of host memory");
    const int staged = !(getenv("VITB200_WEIGHT_STAGE") && atoi(getenv("VITB200_WEIGHT_STAGE")) == 0);
    if (!rc && staged)
        rc = ensure_stage(e);
    /* the two scratch buffers ride at the end of the arena (19 MB that stay allocated): one cudaMalloc and no cudaFree
     * per load instead of three and two -- a cold ViT_opencl call pays for every one of them */
    const size_t scratch_bytes = (scratch_elems * sizeof(float) + 255) & ~(size_t)255;
    if (!rc && (rc = vitcu_malloc(&arena, total + 2 * scratch_bytes)) != 0)
        vit_fail(__FILE__, __LINE__, rc, NULL);
    for (int k = 0; k < 2 && !rc && scratch_elems; k++)
        scratch[k] = (float *)((char *)arena + total + (size_t)k * scratch_bytes);
    int k = 0;
    for (int i = 0; i < e->nblobs && !rc; i++) {
        const size_t n = net[i].size;
        w32[i] = off32[i] == (size_t)-1 ? NULL : (float *)((char *)arena + off32[i]);
        w16[i] = off16[i] == (size_t)-1 ? NULL : (vitcu_bf16 *)((char *)arena + off16[i]);
        if (w32[i])
            rc = upload_blob(e, w32[i], net[i].data, n * sizeof(float), staged);
        if (!rc && w16[i]) {
            /* stream order makes the scratch reuse safe: the conversion that read it two blobs ago
             * precedes this copy on the same stream */
            float *src = w32[i] ? w32[i] : scratch[k++ & 1];
            if (src != w32[i])
                rc = upload_blob(e, src, net[i].data, n * sizeof(float), staged);
            if (!rc && bf)
                rc = vitcu_f32_to_bf16(src, w16[i], n, e->stream);
            if (!rc && e->fp32_tc) { /* [N,K] fp32 -> [N,3K] bf16 pieces */
                const int K = gemm_weight_k(e, i);
                rc = vitcu_split3(src, (size_t)K, w16[i], n / (size_t)K, K, e->stream);
            }
            if (!rc && offf[i] != (size_t)-1) {
                /* the LayerNorm in front of this GEMM folded into its weights (gamma / beta: blobs i-2, i-1;
                 * bias: blob i+1, uploaded right here because the fold needs it before its own turn) */
                const int N = (int)(n / (size_t)e->D);
                char *base = (char *)arena + offf[i];
                wf16[i] = (vitcu_bf16 *)base;
                wf_cs[i] = (float *)(base + ((n * sizeof(vitcu_bf16) + 255) & ~(size_t)255));
                wf_b[i] = (float *)((char *)wf_cs[i] + (((size_t)N * sizeof(float) + 255) & ~(size_t)255));
                float *bias_dev = (float *)((char *)arena + off32[i + 1]);
                rc = vitcu_memcpy_h2d(bias_dev, net[i + 1].data, (size_t)N * sizeof(float), e->stream);
                if (!rc)
                    rc = vitcu_ln_fold_weights(src, w32[i - 2], w32[i - 1], bias_dev, wf16[i], wf_cs[i], wf_b[i], N, e->D, e->stream);
            }
            if (!rc && off8[i] != (size_t)-1) {
                /* FP8: one scale per weight matrix, 448 / max|gamma W| (e4m3 is a floating-point format: 3 mantissa bits
                 * over 2^15 of range, so a per-tensor scale loses nothing a per-channel one would keep) */
                const int K = gemm_weight_k(e, i), N = (int)(n / (size_t)K);
                const int folded = follows_layernorm(e, i);
                char *base = (char *)arena + off8[i];
                wq8[i] = (unsigned char *)base;
                wq8_cs[i] = (float *)(base + ((n + 255) & ~(size_t)255));
                wq8_b[i] = (float *)((char *)wq8_cs[i] + (((size_t)N * sizeof(float) + 255) & ~(size_t)255));
                float amax = 0.f;
                rc = vitcu_memset(e->d_amax, 0, sizeof(float), e->stream);
                if (!rc)
                    rc = vitcu_absmax_f32(src, folded ? w32[i - 2] : NULL, (size_t)N, K, e->d_amax, e->stream);
                if (!rc)
                    rc = vitcu_memcpy_d2h(&amax, e->d_amax, sizeof(float), e->stream);
                if (!rc)
                    rc = vitcu_stream_sync(e->stream);
                if (!rc) {
                    wq8_scale[i] = amax > 0.f ? 448.0f / amax : 1.0f;
                    float *bias_dev = (float *)((char *)arena + off32[i + 1]);
                    if (folded)
                        rc = vitcu_fp8_quant_weights(src, w32[i - 2], w32[i - 1], bias_dev, wq8_scale[i], wq8[i], wq8_cs[i], wq8_b[i], N, K,
                                                     e->stream);
                    else
                        rc = vitcu_fp8_quant_weights(src, NULL, NULL, NULL, wq8_scale[i], wq8[i], NULL, NULL, N, K, e->stream);
                }
            }
        }
        if (rc)
            vit_fail(__FILE__, __LINE__, rc, NULL);
    }
    if (!rc && (rc = vitcu_stream_sync(e->stream)) != 0)
        vit_fail(__FILE__, __LINE__, rc, NULL);
    for (int i = 0; i < VIT_STAGE_SLOTS; i++) /* the ring is idle again (the image uploads record on another stream) */
        e->stage_used[i] = 0;
    if (!rc) {
        for (int i = 0; i < 2; i++) {
            if (e->graph[i])
                vitcu_graph_destroy(e->graph[i]);
            e->graph[i] = NULL;
        }
        e->warmed = 0;
        vitcu_free(e->w_arena); /* the sync above also covers every forward that read it */
        e->w_arena = arena;
        memcpy(e->w32, w32, sizeof(e->w32));
        memcpy(e->w16, w16, sizeof(e->w16));
        memcpy(e->wf16, wf16, sizeof(e->wf16));
        memcpy(e->wf_cs, wf_cs, sizeof(e->wf_cs));
        memcpy(e->wf_b, wf_b, sizeof(e->wf_b));
        memcpy(e->wq8, wq8, sizeof(e->wq8));
        memcpy(e->wq8_cs, wq8_cs, sizeof(e->wq8_cs));
        memcpy(e->wq8_b, wq8_b, sizeof(e->wq8_b));
        memcpy(e->wq8_scale, wq8_scale, sizeof(e->wq8_scale));
        e->fp8_calibrated = 0; /* activation scales belong to the weights: recalibrate on the next chunk */
        e->weights_loaded = 1;
    } else {
        vitcu_free(arena);
    }
    free(w32);
    free(w16);
    free(wf16);
    free(wf_cs);
    free(wf_b);
    free(wq8);
    free(wq8_cs);
    free(wq8_b);
    free(wq8_scale);
    return rc;
}

/* one GEMM of the chain, dispatched on the engine precision.  a_split: A already holds the three
 * bf16 pieces (FP32 tensor-core path; the LayerNorm kernel emits them directly). */
/* timeline profiling: an event in front of the next launch, tagged with what that launch is */
static int mark(vitb200_engine *e, int kind)
{
    if (!e->mark_ev || e->mark_n >= e->mark_cap)
        return 0;
    VIT_TRY(vitcu_event_record(e->mark_ev[e->mark_n], e->stream));
    e->mark_kind[e->mark_n++] = (unsigned char)kind;
    return 0;
}

/* fold: 0 = plain; 1 = this GEMM follows a LayerNorm that is folded into it (A = bf16 residual rows, folded weights,
 * statistics from d_lnstats); 2 = residual GEMM that also emits bf16(x) into d_ln and the row partial sums;
 * 3 + layer = the same with an e4m3 copy for that layer's FP8 fc1; -1 = accumulate mode (FP32 chain at small M: C was
 * zeroed by the preceding LayerNorm launch, the K range is cut into slices that reduce-add into it).
 * a_split: 0 = A is fp32, split here; 1 = A already holds the three bf16 pieces; 2 = A is the fp32 output of an
 * accumulate-mode fc1: GELU, then split */
static int gemm(vitb200_engine *e, const void *A, int a_split, int widx, int bidx, void *C, int M, int N, int K, int epi,
                int out_bf16, int fold)
{
    vitcu_gemm_desc d;
    memset(&d, 0, sizeof(d));
    d.M = M;
    d.N = N;
    d.K = K;
    d.lda = (size_t)K;
    d.ldc = (size_t)N;
    d.epilogue = epi;
    d.bias = fold == 1 ? e->wf_b[widx] : e->w32[bidx];
    d.out_bf16 = out_bf16;
    if (fold == 1) {
        d.ln_stats = e->d_lnstats;
        d.ln_slots = e->D / 128;
        d.ln_colsum = e->wf_cs[widx];
    } else if (fold >= 2) {
        d.emit_bf16 = e->d_ln;
        d.emit_stats = e->d_lnstats;
        if (fold > 2) { /* 3 + layer: the copy is e4m3(x * scale), the A operand of that layer's FP8 fc1 */
            d.emit_fp8 = 1;
            d.emit_scale = e->act_scale_x[fold - 3];
        }
    }
    if (fold == -1)
        d.accumulate = 1;
    if (epi == VITCU_EPI_BIAS_RESIDUAL)
        d.residual = (const float *)C;
    if (epi == VITCU_EPI_PATCH_EMBED) {
        d.pos = e->w32[3];
        d.patches = e->P;
        d.tokens = e->T;
    }
    e->launches++;
    if (e->fp32_tc && a_split != 1) {
        VIT_TRY_RC(mark(e, VIT_K_OTHER));
        int rc = a_split == 2 ? vitcu_split3_gelu((const float *)A, (size_t)K, e->d_a3, (size_t)M, K, e->stream)
                              : vitcu_split3((const float *)A, (size_t)K, e->d_a3, (size_t)M, K, e->stream);
        if (rc)
            return rc;
        e->launches++;
        A = e->d_a3;
    }
    VIT_TRY_RC(mark(e, VIT_K_GEMM));
    const int prof = e->prof_ev && e->prof_n < e->prof_cap;
    if (prof)
        VIT_TRY(vitcu_event_record(e->prof_ev[2 * e->prof_n], e->stream));
    int rc;
    if (e->precision != VITB200_FP32)
        rc = vitcu_gemm_bf16((const vitcu_bf16 *)A, fold == 1 ? e->wf16[widx] : e->w16[widx], C, &d, e->stream);
    else if (e->fp32_tc)
        rc = vitcu_gemm_bf16x3((const vitcu_bf16 *)A, e->w16[widx], C, &d, e->stream);
    else
        rc = vitcu_sgemm((const float *)A, e->w32[widx], C, &d, e->stream);
    if (!rc && prof)
        VIT_TRY(vitcu_event_record(e->prof_ev[2 * e->prof_n++ + 1], e->stream));
    return rc;
}

/* FP8 GEMMs of the MLP (vitcu_gemm_e4m3): fc1 = LayerNorm-folded consumer with GELU writing e4m3, fc2 = residual
 * update emitting the bf16 rows + row sums the next layer's qkv reads */
static int gemm_fp8(vitb200_engine *e, int layer, int is_fc2, int M)
{
    const int w = 4 + 12 * layer;
    vitcu_gemm_desc d;
    memset(&d, 0, sizeof(d));
    d.M = M;
    e->launches++;
    VIT_TRY_RC(mark(e, VIT_K_GEMM));
    const int prof = e->prof_ev && e->prof_n < e->prof_cap;
    if (prof)
        VIT_TRY(vitcu_event_record(e->prof_ev[2 * e->prof_n], e->stream));
    int rc;
    if (!is_fc2) {
        d.N = e->HID;
        d.K = e->D;
        d.epilogue = VITCU_EPI_BIAS_GELU;
        d.bias = e->wq8_b[w + 8];
        d.ln_stats = e->d_lnstats;
        d.ln_slots = e->D / 128;
        d.ln_colsum = e->wq8_cs[w + 8];
        d.acc_scale = 1.0f / (e->act_scale_x[layer] * e->wq8_scale[w + 8]);
        d.out_fp8 = 1;
        d.out_scale = e->act_scale_h[layer];
        rc = vitcu_gemm_e4m3((const uint8_t *)e->d_ln, e->wq8[w + 8], e->d_hid, &d, e->stream);
    } else {
        d.N = e->D;
        d.K = e->HID;
        d.epilogue = VITCU_EPI_BIAS_RESIDUAL;
        d.bias = e->w32[w + 11];
        d.residual = e->d_x;
        d.emit_bf16 = e->d_ln;
        d.emit_stats = e->d_lnstats;
        d.acc_scale = 1.0f / (e->act_scale_h[layer] * e->wq8_scale[w + 10]);
        rc = vitcu_gemm_e4m3((const uint8_t *)e->d_hid, e->wq8[w + 10], e->d_x, &d, e->stream);
    }
    d.lda = (size_t)d.K;
    d.ldc = (size_t)d.N;
    if (!rc && prof)
        VIT_TRY(vitcu_event_record(e->prof_ev[2 * e->prof_n++ + 1], e->stream));
    if (rc)
        return vit_fail(__FILE__, __LINE__, rc, NULL);
    return 0;
}

/* Enqueue the forward of b images resident in d_images[buf] on e->stream. */
#define MARK(kind) VIT_TRY_RC(mark(e, kind))
static int enqueue_forward_mode(vitb200_engine *e, int buf, int b, int calibrate);
static int enqueue_forward(vitb200_engine *e, int buf, int b) { return enqueue_forward_mode(e, buf, b, 0); }

/* calibrate = 1 (FP8 engines, once): run the BF16 chain and record the per-layer maxima of the two tensors the FP8
 * GEMMs will read -- the residual rows in front of LN2 and the GELU output -- into d_amax */
static int enqueue_forward_mode(vitb200_engine *e, int buf, int b, int calibrate)
{
    const int bf = e->precision != VITB200_FP32;
    const int M = b * e->T;
    const int ln_mode = bf ? 1 : (e->fp32_tc ? 2 : 0); /* LayerNorm output: bf16 / three bf16 pieces / fp32 */
    vitcu_stream s = e->stream;
    e->launches = 0;

    /* FP32 chain at small M (batch-1 latency): qkv and fc1 are a few dozen 128 x 128 tiles with the whole K range each
     * -- 36 and 48 of 148 SMs busy for 72 k-blocks.  In accumulate mode their K range is cut into slices that meet in
     * the output through TMA reduce-add, like out-proj and fc2 always did; the output is zeroed by the LayerNorm launch
     * in front of the GEMM and fc1's GELU moves into the split pass in front of fc2 (VITB200_FP32_SPLITK=0: off) */
    const int acc = e->fp32_tc && e->fp32_splitk && vitcu_gemm_split_k_pays(M, 3 * e->D) && vitcu_gemm_split_k_pays(M, e->HID);

    /* patch embedding (Conv2d + postConv2d, R/ViT_opencl.c:361-442).  BF16 path: one TF32
     * tensor-core GEMM that gathers the patches by TMA straight from the NCHW image; FP32 path:
     * gather kernel + FP32-accurate GEMM with the class/position epilogue */
    MARK(VIT_K_OTHER);
    /* BF16 chain, a handful of images: the TF32 patch embedding with K cut into slices that add into token rows which
     * already hold the position embedding (one image: 6 tiles of 48 k-blocks -> 96 work items of 3; VITB200_PE_SPLITK=0: off) */
    const int pe_acc = bf && !e->pe_gather && e->pe_splitk && b * ((e->side * e->side + 127) / 128) * (e->D / 256) * 4 <= 148;
    if (bf && !e->pe_gather) {
        if (pe_acc) {
            VIT_TRY(vitcu_token_rows_init(e->d_x, e->w32[0], e->w32[3], b, e->T, e->D, s));
            e->launches++;
            VIT_TRY(vitcu_patch_embed_tc_acc(e->d_images[buf], e->w32[1], e->w32[2], e->w32[3], e->d_x, b, e->img, e->D, s));
        } else {
            VIT_TRY(vitcu_patch_embed_tc_ex(e->d_images[buf], e->w32[1], e->w32[2], e->w32[3], e->d_x, b, e->img, e->D, s));
        }
        e->launches++;
    } else {
        VIT_TRY(vitcu_patch_gather_ex(e->d_images[buf], e->d_patches, b, e->img, e->patch, bf, s));
        e->launches++;
        if (acc && b == 1) {
            /* one image: its patch rows are rows 1..P of x, a plain 2-D tile target -- the token rows start as the
             * position rows (class row included) and the 12-tile GEMM becomes 144 K slices that add into them */
            VIT_TRY(vitcu_token_rows_init(e->d_x, e->w32[0], e->w32[3], b, e->T, e->D, s));
            e->launches++;
            VIT_TRY(gemm(e, e->d_patches, 0, 1, 2, e->d_x + e->D, e->P, e->D, 3 * e->patch * e->patch, VITCU_EPI_BIAS, 0, -1));
        } else {
            VIT_TRY(gemm(e, e->d_patches, 0, 1, 2, e->d_x, b * e->P, e->D, 3 * e->patch * e->patch, VITCU_EPI_PATCH_EMBED, 0, 0));
        }
    }
    if (!(acc && b == 1) && !pe_acc) { /* (acc implies the FP32 chain, which always takes the gather branch above) */
        MARK(VIT_K_OTHER);
        VIT_TRY(vitcu_cls_rows_ex(e->d_x, e->w32[0], e->w32[3], b, e->T, e->D, s));
        e->launches++;
    }

    const int layers = e->stop_after < 0 ? e->depth : (e->stop_after < e->depth ? e->stop_after : e->depth);
    /* LayerNorm folded into the GEMMs when the chunk is large enough for the CTA-pair kernel, whose residual
     * epilogue can emit the bf16 rows and the row sums: 5 launches per layer instead of 7 and the fp32 stream is
     * not re-read by a LayerNorm kernel.  Small chunks (batch-1 latency) keep the separate kernel. */
    const int fold = e->ln_fold && vitcu_gemm_bf16_emit_supported(M, e->D);
    const int fp8 = e->fp8 && fold && e->fp8_calibrated && !calibrate;
    if (calibrate && !fold)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "FP8 calibration needs a chunk that runs the CTA-pair GEMM");
    if (calibrate)
        VIT_TRY(vitcu_memset(e->d_amax, 0, (size_t)2 * VIT_MAX_DEPTH * sizeof(float), s));
    if (fold && layers > 0) {
        MARK(VIT_K_LAYERNORM);
        VIT_TRY(vitcu_rowstats_cast(e->d_x, (vitcu_bf16 *)e->d_ln, e->d_lnstats, M, e->D, e->D / 128, s));
        e->launches++;
    }
    for (int l = 0; l < layers; l++) {
        const int w = 4 + 12 * l; /* blob base of the layer (R/ViT_seq.c:446-504) */
        /* x -> LN1 -> QKV -> attention -> out-proj (+x)  (R/ViT_opencl.c:710-730) */
        if (!fold) {
            MARK(VIT_K_LAYERNORM);
            VIT_TRY(vitcu_layernorm_zero(e->d_x, e->D, e->d_ln, ln_mode, e->w32[w + 0], e->w32[w + 1], M, e->D,
                                         acc ? e->d_qkv : NULL, acc ? (size_t)M * 3 * e->D * sizeof(float) : 0, s));
            e->launches++;
        }
        VIT_TRY(gemm(e, e->d_ln, 1, w + 2, w + 3, e->d_qkv, M, 3 * e->D, e->D, VITCU_EPI_BIAS, bf, acc ? -1 : fold));
        MARK(VIT_K_ATTENTION);
        /* FP32 tensor-core path: the attention kernel writes its output already split into three bf16 pieces */
        VIT_TRY(vitcu_attention_ex(e->d_qkv, e->d_att, b, e->T, e->heads, e->fp32_tc ? 2 : bf, s));
        e->launches++;
        VIT_TRY(gemm(e, e->d_att, e->fp32_tc, w + 4, w + 5, e->d_x, M, e->D, e->D, VITCU_EPI_BIAS_RESIDUAL, 0,
                     fp8 ? 3 + l : (fold ? 2 : 0)));
        if (calibrate) {
            VIT_TRY(vitcu_absmax_bf16((const vitcu_bf16 *)e->d_ln, (size_t)M * e->D, e->d_amax + 2 * l, s));
            e->launches++;
        }
        /* -> LN2 -> fc1+GELU -> fc2 (+r1)  (R/ViT_opencl.c:732-746) */
        if (!fold) {
            MARK(VIT_K_LAYERNORM);
            VIT_TRY(vitcu_layernorm_zero(e->d_x, e->D, e->d_ln, ln_mode, e->w32[w + 6], e->w32[w + 7], M, e->D,
                                         acc ? e->d_hid : NULL, acc ? (size_t)M * e->HID * sizeof(float) : 0, s));
            e->launches++;
        }
        if (fp8) {
            VIT_TRY_RC(gemm_fp8(e, l, 0, M));
            VIT_TRY_RC(gemm_fp8(e, l, 1, M));
            continue;
        }
        VIT_TRY(gemm(e, e->d_ln, 1, w + 8, w + 9, e->d_hid, M, e->HID, e->D, acc ? VITCU_EPI_BIAS : VITCU_EPI_BIAS_GELU, bf,
                     acc ? -1 : fold));
        if (calibrate) {
            VIT_TRY(vitcu_absmax_bf16((const vitcu_bf16 *)e->d_hid, (size_t)M * e->HID, e->d_amax + 2 * l + 1, s));
            e->launches++;
        }
        /* the last layer's fc2 has no LayerNorm consumer over all rows (the final one visits the class rows only) */
        VIT_TRY(gemm(e, e->d_hid, acc ? 2 : 0, w + 10, w + 11, e->d_x, M, e->D, e->HID, VITCU_EPI_BIAS_RESIDUAL, 0,
                     fold && l + 1 < layers ? 2 : 0));
    }
    if (e->stop_after >= 0)
        return 0;

    /* final LN on the class-token rows only, head, softmax
     * (R/ViT_opencl.c:951-959; R/ViT_seq.c:506-515 normalises all rows but uses row 0) */
    MARK(VIT_K_OTHER); /* final LN + head + softmax: one 'other' span */
    VIT_TRY(vitcu_layernorm_ex(e->d_x, (size_t)e->T * e->D, e->d_cls, 0, e->w32[e->nblobs - 4], e->w32[e->nblobs - 3], b, e->D, s));
    e->launches++;
    vitcu_gemm_desc d;
    memset(&d, 0, sizeof(d));
    d.M = b;
    d.N = VITB200_CLASSES;
    d.K = e->D;
    d.lda = e->D;
    d.ldc = VITB200_CLASSES;
    d.epilogue = VITCU_EPI_BIAS;
    d.bias = e->w32[e->nblobs - 1];
    /* the head stays FP32 on both paths: 0.004 % of the FLOPs, all of the logit precision */
    const float *head_w = e->w32[e->nblobs - 2];
    VIT_TRY(vitcu_sgemm(e->d_cls, head_w, e->d_logits[buf], &d, s));
    e->launches++;
    VIT_TRY(vitcu_softmax_rows(e->d_logits[buf], e->d_probs[buf], b, VITB200_CLASSES, s));
    e->launches++;
    return 0;
}

/* forward of a chunk, through the captured graph when the chunk is full-size */
/* FP8 engines: activation scales from the maxima of the first eligible chunk, with a factor 2 of head room (the
 * conversion saturates at +-448).  One extra BF16 forward per weight load. */
static int fp8_calibrate(vitb200_engine *e, int buf, int b)
{
    float amax[2 * VIT_MAX_DEPTH];
    VIT_TRY_RC(enqueue_forward_mode(e, buf, b, 1));
    VIT_TRY(vitcu_memcpy_d2h(amax, e->d_amax, sizeof(amax), e->stream));
    VIT_TRY(vitcu_stream_sync(e->stream));
    for (int l = 0; l < e->depth; l++) {
        e->act_scale_x[l] = amax[2 * l] > 0.f ? 224.0f / amax[2 * l] : 1.0f;
        e->act_scale_h[l] = amax[2 * l + 1] > 0.f ? 224.0f / amax[2 * l + 1] : 1.0f;
    }
    e->fp8_calibrated = 1;
    return 0;
}

static int run_chunk(vitb200_engine *e, int buf, int b)
{
    if (e->fp8 && !e->fp8_calibrated && e->stop_after < 0 && vitcu_gemm_bf16_emit_supported(b * e->T, e->D))
        VIT_TRY_RC(fp8_calibrate(e, buf, b));
    /* the first full-size chunk runs eagerly (module load, attribute set-up);
     * later ones are captured once per input buffer and replayed */
    if (b == e->B && e->stop_after < 0 && !e->no_graph && e->warmed) {
        if (!e->graph[buf]) {
            VIT_TRY(vitcu_graph_begin(e->stream));
            int rc = enqueue_forward(e, buf, b);
            vitcu_graph g = NULL;
            int rc2 = vitcu_graph_end(e->stream, &g);
            if (rc)
                return rc;
            if (rc2)
                return vit_fail(__FILE__, __LINE__, rc2, NULL);
            e->graph[buf] = g;
            e->kernels_per_forward = e->launches;
        }
        VIT_TRY(vitcu_graph_launch(e->graph[buf], e->stream));
        return 0;
    }
    int rc = enqueue_forward(e, buf, b);
    if (!rc && e->stop_after < 0) {
        e->kernels_per_forward = e->launches;
        if (b == e->B)
            e->warmed = 1;
    }
    return rc;
}

static int check_ready(vitb200_engine *e)
{
    if (!e)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "engine is NULL");
    if (!e->weights_loaded)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "weights not loaded");
    VIT_TRY(vitcu_set_device(e->device));
    return 0;
}

/* Upload of b images from pageable host memory into d_images[buf]: the pool gathers a group of
 * images into a pinned slot, the copy stream DMAs the slot, and a slot is refilled once the DMA
 * that read it (VIT_STAGE_SLOTS groups ago) has finished -- so gathering group g+1 overlaps the
 * DMA of group g.  The ring and the threads are created on first use. */
static int upload_pageable(vitb200_engine *e, int buf, const float *contig, const vitb200_image *structs, int b)
{
    const size_t img_elems = (size_t)3 * e->img * e->img, img_bytes = img_elems * sizeof(float);
    VIT_TRY_RC(ensure_stage(e));
    for (int g0 = 0; g0 < b; g0 += e->stage_group) {
        const int g = b - g0 < e->stage_group ? b - g0 : e->stage_group;
        const int slot = e->stage_next++ % VIT_STAGE_SLOTS;
        char *h = e->h_stage + (size_t)slot * e->stage_group * img_bytes;
        if (e->stage_used[slot])
            VIT_TRY(vitcu_event_sync(e->ev_slot[slot]));
        vit_stager_copy(e->stager, h, structs ? structs + g0 : NULL, contig ? contig + (size_t)g0 * img_elems : NULL,
                        img_bytes, g);
        VIT_TRY(vitcu_memcpy_h2d(e->d_images[buf] + (size_t)g0 * img_elems, h, (size_t)g * img_bytes,
                                 e->copy_stream));
        VIT_TRY(vitcu_event_record(e->ev_slot[slot], e->copy_stream));
        e->stage_used[slot] = 1;
    }
    return 0;
}

int vitb200_next_chunk(int n, int done, int chunk, int head_split, int tail_split)
{
    if (n <= 0 || done < 0 || done >= n || chunk <= 0)
        return 0;
    const int left = n - done;
    int b = left < chunk ? left : chunk;
    /* when the host-side staging is the slower side (a call sharded over many GPUs shares the host's copy bandwidth),
     * everything after the last upload is exposed -- the whole forward of the last chunk: leave a quarter chunk for it */
    if (tail_split && done > 0 && left <= chunk && left > chunk / 2 && chunk >= 4)
        b = left - chunk / 4;
    /* at the head of a call nothing can be computed before the first chunk has arrived: a quarter chunk first (0.8 ms
     * of upload instead of 3 ms at 256 images), the rest in full chunks whose upload hides under the forward before them */
    if (head_split && done == 0 && n >= chunk && chunk >= 4)
        b = chunk / 4;
    return b;
}

/* Shared chunk pipeline.  Source of chunk c is either a contiguous host array
 * (images_host) or per-image structs.  While chunk c computes, the host copies
 * chunk c-1's results out of pinned staging and the copy stream uploads chunk
 * c+1, so the GPU never waits for the host. */
static int forward_pipeline(vitb200_engine *e, const float *images_host, const vitb200_image *structs, int n,
                            float *probs_host, float **prob_rows, float *logits_host, int topk, int *top_idx,
                            float *top_val)
{
    const size_t img_elems = (size_t)3 * e->img * e->img;
    const size_t pbytes = (size_t)VITB200_CLASSES * sizeof(float);
    const size_t stage = (size_t)e->B * VITB200_CLASSES;
    int done = 0, chunk = 0;
    int pend_n = 0, pend_off = 0, pend_buf = 0;
    /* a pinned (or registered) source is DMA'd in place; pageable memory goes through the stager */
    int pinned = 0;
    VIT_TRY(vitcu_host_is_pinned(images_host ? (const void *)images_host : (const void *)structs[0].data, &pinned));
    if (getenv("VITB200_NO_STAGER"))
        pinned = 1; /* let the driver stage pageable copies itself */
    /* Chunks are cut (below) only where a quarter chunk still runs the kernels a full chunk runs -- the CTA-pair GEMMs,
     * whose tiles are anchored at absolute rows, so an image's result does not depend on the chunk it travels in; a
     * smaller piece would take the single-CTA / split-K kernels and differ in rounding from the same image in a full
     * chunk (tests/test_gpu_forward.py::test_vit_opencl_multi_gpu_split compares calls bit for bit) */
    const int split_ok = e->B >= 64 && vitcu_gemm_bf16_emit_supported((e->B / 4) * e->T, e->D);
    const int tail_split = split_ok && !(getenv("VITB200_TAIL_SPLIT") && atoi(getenv("VITB200_TAIL_SPLIT")) == 0);
    const int head_split = split_ok && !(getenv("VITB200_HEAD_SPLIT") && atoi(getenv("VITB200_HEAD_SPLIT")) == 0);
    while (done < n || pend_n) {
        const int buf = chunk & 1;
        int b = 0;
        if (done < n) {
            /* full chunks, with a quarter chunk cut off the head of the call and (pageable sources, where the host-side
             * staging can be the slower side of the pipeline) off its tail: vitb200_next_chunk */
            b = vitb200_next_chunk(n, done, e->B, head_split, !pinned && tail_split);
            /* d_images[buf] is free once the forward that read it (chunk-2) finished */
            VIT_TRY(vitcu_stream_wait_event(e->copy_stream, e->ev_done[buf]));
            if (!pinned) {
                VIT_TRY_RC(upload_pageable(e, buf, images_host ? images_host + (size_t)done * img_elems : NULL,
                                           structs ? structs + done : NULL, b));
            } else if (images_host) {
                VIT_TRY(vitcu_memcpy_h2d(e->d_images[buf], images_host + (size_t)done * img_elems,
                                         (size_t)b * img_elems * sizeof(float), e->copy_stream));
            } else {
                /* per-image malloc'd buffers (R/Network.c:84-105): one copy per image */
                for (int i = 0; i < b; i++)
                    VIT_TRY(vitcu_memcpy_h2d(e->d_images[buf] + (size_t)i * img_elems, structs[done + i].data,
                                             img_elems * sizeof(float), e->copy_stream));
            }
            VIT_TRY(vitcu_event_record(e->ev_h2d[buf], e->copy_stream));
            VIT_TRY(vitcu_stream_wait_event(e->stream, e->ev_h2d[buf]));
            VIT_TRY_RC(run_chunk(e, buf, b));
            VIT_TRY(vitcu_event_record(e->ev_done[buf], e->stream));
            if (probs_host || prob_rows)
                VIT_TRY(vitcu_memcpy_d2h(e->h_probs + buf * stage, e->d_probs[buf], (size_t)b * pbytes, e->stream));
            if (topk) {
                /* labels only: k*8 bytes per image come back instead of 4 KB */
                const size_t tk = (size_t)e->B * VITB200_TOPK_MAX;
                VIT_TRY(vitcu_topk_rows(e->d_probs[buf], b, VITB200_CLASSES, topk, e->d_topi + buf * tk,
                                        e->d_topv + buf * tk, e->stream));
                VIT_TRY(vitcu_memcpy_d2h(e->h_topi + buf * tk, e->d_topi + buf * tk, (size_t)b * topk * sizeof(int),
                                         e->stream));
                VIT_TRY(vitcu_memcpy_d2h(e->h_topv + buf * tk, e->d_topv + buf * tk,
                                         (size_t)b * topk * sizeof(float), e->stream));
            }
            if (logits_host)
                VIT_TRY(vitcu_memcpy_d2h(e->h_logits + buf * stage, e->d_logits[buf], (size_t)b * pbytes,
                                         e->stream));
            VIT_TRY(vitcu_event_record(e->ev_out[buf], e->stream));
        }
        if (pend_n) {
            VIT_TRY(vitcu_event_sync(e->ev_out[pend_buf]));
            const float *src = e->h_probs + pend_buf * stage;
            if (probs_host)
                memcpy(probs_host + (size_t)pend_off * VITB200_CLASSES, src, (size_t)pend_n * pbytes);
            if (prob_rows)
                for (int i = 0; i < pend_n; i++)
                    memcpy(prob_rows[pend_off + i], src + (size_t)i * VITB200_CLASSES, pbytes);
            if (logits_host)
                memcpy(logits_host + (size_t)pend_off * VITB200_CLASSES, e->h_logits + pend_buf * stage,
                       (size_t)pend_n * pbytes);
            if (topk) {
                const size_t tk = (size_t)e->B * VITB200_TOPK_MAX;
                memcpy(top_idx + (size_t)pend_off * topk, e->h_topi + pend_buf * tk,
                       (size_t)pend_n * topk * sizeof(int));
                memcpy(top_val + (size_t)pend_off * topk, e->h_topv + pend_buf * tk,
                       (size_t)pend_n * topk * sizeof(float));
            }
            pend_n = 0;
        }
        if (b) {
            pend_n = b;
            pend_off = done;
            pend_buf = buf;
            done += b;
            chunk++;
        }
    }
    VIT_TRY(vitcu_stream_sync(e->stream));
    VIT_TRY(vitcu_watchdog_check());
    return 0;
}

int vitb200_forward(vitb200_engine *e, const float *images_host, int n, float *probs_host, float *logits_host)
{
    VIT_TRY_RC(check_ready(e));
    if (n < 0 || (n > 0 && (!images_host || !probs_host)))
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "bad images/probs argument");
    if (n == 0)
        return 0;
    return forward_pipeline(e, images_host, NULL, n, probs_host, NULL, logits_host, 0, NULL, NULL);
}

int vitb200_forward_topk(vitb200_engine *e, const float *images_host, int n, int k, int *labels, float *probs)
{
    VIT_TRY_RC(check_ready(e));
    if (n < 0 || (n > 0 && (!images_host || !labels || !probs)))
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "bad images/labels/probs argument");
    if (k < 1 || k > VITB200_TOPK_MAX)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "k must be in 1..VITB200_TOPK_MAX");
    if (n == 0)
        return 0;
    if (!e->h_topv) { /* first use: all four buffers or none (h_topv is committed last) */
        const size_t tk = (size_t)e->B * VITB200_TOPK_MAX * 2;
        int *di = NULL, *hi = NULL;
        float *dv = NULL, *hv = NULL;
        int rc = vitcu_malloc((void **)&di, tk * sizeof(int));
        if (!rc)
            rc = vitcu_malloc((void **)&dv, tk * sizeof(float));
        if (!rc)
            rc = vitcu_host_alloc((void **)&hi, tk * sizeof(int));
        if (!rc)
            rc = vitcu_host_alloc((void **)&hv, tk * sizeof(float));
        if (rc) {
            vit_fail(__FILE__, __LINE__, rc, NULL);
            vitcu_free(di);
            vitcu_free(dv);
            vitcu_host_free(hi);
            return rc;
        }
        e->d_topi = di;
        e->d_topv = dv;
        e->h_topi = hi;
        e->h_topv = hv;
    }
    return forward_pipeline(e, images_host, NULL, n, NULL, NULL, NULL, k, labels, probs);
}

int vitb200_forward_structs(vitb200_engine *e, const vitb200_image *images, int n, float **prb)
{
    VIT_TRY_RC(check_ready(e));
    if (n < 0 || (n > 0 && (!images || !prb)))
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "bad images/prb argument");
    for (int i = 0; i < n; i++) {
        if (!images[i].data || !prb[i])
            return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "NULL image data or probability row");
        if (images[i].c != 3 || images[i].h != e->img || images[i].w != e->img)
            return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "image shape does not match the engine");
    }
    if (n == 0)
        return 0;
    return forward_pipeline(e, NULL, images, n, NULL, prb, NULL, 0, NULL, NULL);
}

int vitb200_stage_images(vitb200_engine *e, const float *images_host, int n)
{
    VIT_TRY_RC(check_ready(e));
    if (n <= 0 || n > e->B || !images_host)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "stage_images: need 0 < n <= max_batch");
    VIT_TRY(vitcu_memcpy_h2d(e->d_images[0], images_host, (size_t)n * 3 * e->img * e->img * sizeof(float),
                             e->stream));
    VIT_TRY(vitcu_stream_sync(e->stream));
    return 0;
}

int vitb200_forward_resident(vitb200_engine *e, int n)
{
    VIT_TRY_RC(check_ready(e));
    if (n <= 0 || n > e->B)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "forward_resident: need 0 < n <= max_batch");
    VIT_TRY(vitcu_event_record(e->ev_t0, e->stream));
    VIT_TRY_RC(run_chunk(e, 0, n));
    VIT_TRY(vitcu_event_record(e->ev_t1, e->stream));
    VIT_TRY(vitcu_stream_sync(e->stream));
    VIT_TRY(vitcu_watchdog_check());
    return 0;
}

int vitb200_time_resident(vitb200_engine *e, int n, int iters, float *total_ms)
{
    VIT_TRY_RC(check_ready(e));
    if (n <= 0 || n > e->B || iters <= 0 || !total_ms)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "time_resident: bad argument");
    VIT_TRY(vitcu_event_record(e->ev_t0, e->stream));
    for (int i = 0; i < iters; i++)
        VIT_TRY_RC(run_chunk(e, 0, n));
    VIT_TRY(vitcu_event_record(e->ev_t1, e->stream));
    VIT_TRY(vitcu_stream_sync(e->stream));
    VIT_TRY(vitcu_event_elapsed_ms(e->ev_t0, e->ev_t1, total_ms));
    VIT_TRY(vitcu_watchdog_check());
    return 0;
}

int vitb200_profile_gemms(vitb200_engine *e, int n, int iters, float *gemm_ms_per_forward, int *gemm_launches)
{
    VIT_TRY_RC(check_ready(e));
    if (n <= 0 || n > e->B || iters <= 0 || !gemm_ms_per_forward)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "profile_gemms: bad argument");
    const int cap = 4 * e->depth + 1; /* qkv, out_proj, fc1, fc2 per layer (+ the gather-path patch embedding) */
    vitcu_event *ev = (vitcu_event *)calloc((size_t)2 * cap, sizeof(vitcu_event));
    if (!ev)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "out of host memory");
    int rc = 0;
    for (int i = 0; i < 2 * cap && !rc; i++)
        rc = vitcu_event_create(&ev[i]);
    double total = 0.0;
    int launches = 0;
    for (int it = 0; it < iters && !rc; it++) {
        e->prof_ev = ev;
        e->prof_cap = cap;
        e->prof_n = 0;
        rc = enqueue_forward(e, 0, n); /* eager: events cannot sit inside the captured graph */
        e->prof_ev = NULL;
        if (!rc)
            rc = vitcu_stream_sync(e->stream);
        launches = e->prof_n;
        for (int k = 0; k < launches && !rc; k++) {
            float ms = 0.f;
            rc = vitcu_event_elapsed_ms(ev[2 * k], ev[2 * k + 1], &ms);
            total += ms;
        }
    }
    for (int i = 0; i < 2 * cap; i++)
        if (ev[i])
            vitcu_event_destroy(ev[i]);
    free(ev);
    if (rc)
        return vit_fail(__FILE__, __LINE__, rc, NULL);
    *gemm_ms_per_forward = (float)(total / iters);
    if (gemm_launches)
        *gemm_launches = launches;
    VIT_TRY(vitcu_watchdog_check());
    return 0;
}

int vitb200_profile_timeline(vitb200_engine *e, int n, float ms_by_kind[4], int launches_by_kind[4])
{
    VIT_TRY_RC(check_ready(e));
    if (n <= 0 || n > e->B || !ms_by_kind)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "profile_timeline: bad argument");
    const int cap = 16 * e->depth + 16;
    vitcu_event *ev = (vitcu_event *)calloc((size_t)cap, sizeof(vitcu_event));
    unsigned char *kind = (unsigned char *)calloc((size_t)cap, 1);
    if (!ev || !kind) {
        free(ev);
        free(kind);
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "out of host memory");
    }
    int rc = 0;
    for (int i = 0; i < cap && !rc; i++)
        rc = vitcu_event_create(&ev[i]);
    e->mark_ev = ev;
    e->mark_kind = kind;
    e->mark_cap = cap - 1;
    e->mark_n = 0;
    if (!rc)
        rc = enqueue_forward(e, 0, n);
    const int marks = e->mark_n;
    if (!rc)
        rc = vitcu_event_record(ev[marks], e->stream); /* end of the last span */
    e->mark_ev = NULL;
    if (!rc)
        rc = vitcu_stream_sync(e->stream);
    for (int k = 0; k < 4; k++) {
        ms_by_kind[k] = 0.f;
        if (launches_by_kind)
            launches_by_kind[k] = 0;
    }
    for (int k = 0; k < marks && !rc; k++) {
        float ms = 0.f;
        rc = vitcu_event_elapsed_ms(ev[k], ev[k + 1], &ms);
        ms_by_kind[kind[k] & 3] += ms;
        if (launches_by_kind)
            launches_by_kind[kind[k] & 3]++;
    }
    for (int i = 0; i < cap; i++)
        if (ev[i])
            vitcu_event_destroy(ev[i]);
    free(ev);
    free(kind);
    if (rc)
        return vit_fail(__FILE__, __LINE__, rc, NULL);
    return 0;
}

int vitb200_last_forward_ms(vitb200_engine *e, float *ms)
{
    if (!e || !ms)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "NULL argument");
    VIT_TRY(vitcu_event_elapsed_ms(e->ev_t0, e->ev_t1, ms));
    return 0;
}

int vitb200_read_probs(vitb200_engine *e, int n, float *probs_host, float *logits_host)
{
    VIT_TRY_RC(check_ready(e));
    if (n <= 0 || n > e->B)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "read_probs: need 0 < n <= max_batch");
    const size_t bytes = (size_t)n * VITB200_CLASSES * sizeof(float);
    if (probs_host)
        VIT_TRY(vitcu_memcpy_d2h(probs_host, e->d_probs[0], bytes, e->stream));
    if (logits_host)
        VIT_TRY(vitcu_memcpy_d2h(logits_host, e->d_logits[0], bytes, e->stream));
    VIT_TRY(vitcu_stream_sync(e->stream));
    return 0;
}

int vitb200_set_stop_after_layer(vitb200_engine *e, int layer)
{
    if (!e || layer < -1 || layer > e->depth)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "stop_after_layer must be in [-1,12]");
    e->stop_after = layer;
    return 0;
}

int vitb200_read_tokens(vitb200_engine *e, int n, float *x_host)
{
    VIT_TRY_RC(check_ready(e));
    if (n <= 0 || n > e->B || !x_host)
        return vit_fail(__FILE__, __LINE__, VITCU_E_ARG, "read_tokens: bad argument");
    VIT_TRY(vitcu_memcpy_d2h(x_host, e->d_x, (size_t)n * e->T * e->D * sizeof(float), e->stream));
    VIT_TRY(vitcu_stream_sync(e->stream));
    return 0;
}

int vitb200_kernels_per_forward(const vitb200_engine *e) { return e ? e->kernels_per_forward : 0; }
int vitb200_tokens(const vitb200_engine *e) { return e ? e->T : 0; }
int vitb200_embed(const vitb200_engine *e) { return e ? e->D : 0; }
