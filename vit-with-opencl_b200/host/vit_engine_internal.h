/* host/vit_engine_internal.h -- engine state shared by the C host files. */
#ifndef VIT_ENGINE_INTERNAL_H
#define VIT_ENGINE_INTERNAL_H

#include "../../include/vit_b200.h"
#include "../../include/vit_cuda_layer.h"

/* blob slots: 4 embedding blobs + 12 per encoder layer + 4 head blobs (R/ViT_seq.c:437-513 for depth 12) */
#define VIT_MAX_DEPTH 32
#define VIT_MAX_BLOBS (8 + 12 * VIT_MAX_DEPTH)

struct vitb200_engine {
    int device, img, side, P, T, precision, B;
    int patch, D, HID, depth, heads; /* model dims (ViT-B/16: 16, 768, 3072, 12, 12) */
    int nblobs;                      /* 8 + 12 * depth */
    int weights_loaded, stop_after, no_graph, warmed;
    int pe_gather;                   /* BF16 path: use the gather kernel + BF16 GEMM for the patch embedding */
    int fp32_tc;                     /* FP32 precision computed as split-bf16 (x3 pieces, 6 products) on the tensor cores */
    int fp32_splitk;                 /* ... with qkv / fc1 in accumulate mode (split-K) when the chunk is a few dozen tiles */
    int pe_splitk;                   /* BF16 chain: K-sliced (accumulate) TF32 patch embedding for a handful of images */
    int launches, kernels_per_forward;
    vitcu_stream stream, copy_stream;
    vitcu_event ev_h2d[2], ev_done[2], ev_out[2], ev_t0, ev_t1;
    vitcu_graph graph[2];
    void *w_arena;                   /* one device allocation holding every weight below */
    void *d_arena;                   /* one device allocation holding every activation buffer below */
    float *w32[VIT_MAX_BLOBS];         /* fp32 blobs on the device (pointers into w_arena) */
    vitcu_bf16 *w16[VIT_MAX_BLOBS]; /* BF16 path: bf16 GEMM weights [N,K]; FP32 tensor-core path: three bf16 pieces [N,3K] */
    /* LayerNorm folded into the GEMMs (BF16 path, big chunks): folded copies of the weights that follow a LayerNorm
     * (in_proj: blob base + 2, fc1: base + 8): bf16(gamma * W), column sums, b + W beta; the bf16 copy of the residual
     * rows lives in d_ln, the rows' partial sums in d_lnstats */
    int ln_fold;
    vitcu_bf16 *wf16[VIT_MAX_BLOBS];
    float *wf_cs[VIT_MAX_BLOBS], *wf_b[VIT_MAX_BLOBS];
    void *d_lnstats;                 /* float2 [D / 128][B*T] */
    /* FP8 precision: fc1 (folded) and fc2 weights as e4m3 with one scale each; activation scales (the e4m3 copy of the
     * residual rows that fc1 reads, the GELU output that fc2 reads) from a calibration pass over the first chunk */
    int fp8, fp8_calibrated;
    unsigned char *wq8[VIT_MAX_BLOBS];
    float *wq8_cs[VIT_MAX_BLOBS], *wq8_b[VIT_MAX_BLOBS];
    float wq8_scale[VIT_MAX_BLOBS];
    float act_scale_x[VIT_MAX_DEPTH], act_scale_h[VIT_MAX_DEPTH];
    float *d_amax;                   /* [2 * depth] calibration maxima */
    vitcu_bf16 *d_a3;                /* FP32 tensor-core path: split form [rows,3K] of the current GEMM A operand */
    float *d_images[2];              /* double-buffered input chunk [B,3,img,img] */
    void *d_patches;                 /* [B*P,768] gathered patches */
    float *d_x;                      /* [B*T,768] fp32 residual stream */
    void *d_ln, *d_qkv, *d_att, *d_hid;
    float *d_cls;                    /* [B,768] final-LN class tokens */
    float *d_logits[2], *d_probs[2]; /* [B,1000] */
    float *h_probs, *h_logits;       /* pinned staging [2][B,1000] */
    int *d_topi, *h_topi;            /* [2][B,VITB200_TOPK_MAX] labels of vitb200_forward_topk (allocated on first use) */
    float *d_topv, *h_topv;
    /* in-situ GEMM timing (vitb200_profile_gemms): one event pair per GEMM launch of an eager forward */
    vitcu_event *prof_ev;            /* [2 * prof_cap], NULL when not profiling */
    int prof_cap, prof_n;
    /* whole-timeline variant (vitb200_profile_timeline): one event before every launch, tagged by kind */
    vitcu_event *mark_ev;
    unsigned char *mark_kind;
    int mark_cap, mark_n;
    /* pageable sources: worker threads gather images into a ring of pinned slots (vit_stage.c) */
    struct vit_stager *stager;
    char *h_stage;                   /* [VIT_STAGE_SLOTS][stage_group images] pinned */
    int stage_group, stage_next, stage_used[4];
    vitcu_event ev_slot[4];
};

enum { VIT_K_GEMM = 0, VIT_K_ATTENTION = 1, VIT_K_LAYERNORM = 2, VIT_K_OTHER = 3 };
#define VIT_STAGE_SLOTS 3
typedef struct vit_stager vit_stager;
int vit_stager_threads_default(void);
vit_stager *vit_stager_create(int threads);
void vit_stager_destroy(vit_stager *s);
/* copies `count` blocks of `bytes` into dst (block i from structs[i].data, or contig + i*bytes when
 * structs is NULL) on all threads of the pool plus the caller; returns when every block is copied */
void vit_stager_copy(vit_stager *s, void *dst, const vitb200_image *structs, const void *contig, size_t bytes,
                     int count);

/* records "[file:line] ..." for vitb200_last_error(); what == NULL takes the
 * device layer's message */
int vit_fail(const char *file, int line, int rc, const char *what);

/* device-layer call: on failure record its message and return the code */
#define VIT_TRY(call)                                                          \
    do {                                                                       \
        int _rc = (call);                                                      \
        if (_rc)                                                               \
            return vit_fail(__FILE__, __LINE__, _rc, NULL);                    \
    } while (0)
/* host-level call whose message is already recorded */
#define VIT_TRY_RC(call)                                                       \
    do {                                                                       \
        int _rc = (call);                                                      \
        if (_rc)                                                               \
            return _rc;                                                        \
    } while (0)

#endif
