/* oracle/vit_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see vit_oracle.h).
 *
 * Restates /root/reference/MulticoreMainProject/ViT_seq.c.  R/ below means that
 * directory.  The rule followed everywhere: an output element is the result of
 * exactly the float operations the reference performs for it, in the same
 * order (accumulator starts at the bias, products are added in increasing-k
 * order, no fused multiply-add, no re-association).  What differs is only which
 * independent elements are computed side by side: the weight matrix is
 * transposed once so 32 neighbouring outputs advance together in SIMD lanes,
 * and independent tokens go to different OpenMP threads.
 *
 * Build: gcc -O2 -mavx2 -fopenmp -ffp-contract=off (oracle/Makefile).  -mavx2
 * without -mfma cannot emit FMA; -ffp-contract=off says so explicitly.
 */
#pragma GCC optimize("fp-contract=off")
#include "vit_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define D VIT_ORACLE_EMBED
#define NH VIT_ORACLE_HEADS
#define HD (VIT_ORACLE_EMBED / VIT_ORACLE_HEADS)
#define HID VIT_ORACLE_HIDDEN
#define NCLS VIT_ORACLE_CLASSES
#define PS VIT_ORACLE_PATCH
#define LANES 32

static int g_threads = 0;

void vit_oracle_set_threads(int n)
{
    g_threads = n;
#ifdef _OPENMP
    if (n > 0)
        omp_set_num_threads(n);
#endif
}

int vit_oracle_get_threads(void)
{
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}

int vit_oracle_tokens(int img)
{
    int side = img / PS;
    return side * side + 1;
}

static float *xmalloc_f(size_t n)
{
    float *p = (float *)malloc(n * sizeof(float));
    if (!p)
        abort();
    return p;
}

/* [rows, cols] -> [cols, rows] */
static float *transpose(const float *w, int rows, int cols)
{
    float *t = xmalloc_f((size_t)rows * cols);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < cols; c++)
        for (int r = 0; r < rows; r++)
            t[(size_t)c * rows + r] = w[(size_t)r * cols + c];
    return t;
}

/* exact-erf GELU, R/ViT_seq.c:283-286: 0.5f*x*(1.0f+erff(x/sqrtf(2.0f))) */
static inline float gelu_ref(float x)
{
    return 0.5f * x * (1.0f + erff(x / sqrtf(2.0f)));
}

/* y[t,o] = b[o] + sum_i x[t,i]*w[o,i], i ascending -- R/ViT_seq.c:295-309.
 * wt is w transposed ([in_f, out_f]) so the LANES outputs of a block are
 * contiguous; each lane keeps its own sequential accumulator. */
static void linear_t(const float *x, float *y, int tokens, int in_f, int out_f,
                     const float *wt, const float *b, int gelu)
{
#pragma omp parallel for schedule(static)
    for (int t = 0; t < tokens; t++) {
        const float *xr = x + (size_t)t * in_f;
        float *yr = y + (size_t)t * out_f;
        int ob = 0;
        for (; ob + LANES <= out_f; ob += LANES) {
            float acc[LANES];
            for (int j = 0; j < LANES; j++)
                acc[j] = b[ob + j];
            for (int i = 0; i < in_f; i++) {
                const float xi = xr[i];
                const float *wr = wt + (size_t)i * out_f + ob;
                for (int j = 0; j < LANES; j++)
                    acc[j] += xi * wr[j];
            }
            for (int j = 0; j < LANES; j++)
                yr[ob + j] = acc[j];
        }
        for (; ob < out_f; ob++) { /* ragged tail (out_f = 1000) */
            float acc = b[ob];
            for (int i = 0; i < in_f; i++)
                acc += xr[i] * wt[(size_t)i * out_f + ob];
            yr[ob] = acc;
        }
        if (gelu) /* R/ViT_seq.c:320-323 applies GELU after the whole fc1 */
            for (int o = 0; o < out_f; o++)
                yr[o] = gelu_ref(yr[o]);
    }
}

void vit_oracle_linear(const float *x, float *y, int tokens, int in_f, int out_f,
                       const float *w, const float *b, int gelu)
{
    float *wt = transpose(w, out_f, in_f);
    linear_t(x, y, tokens, in_f, out_f, wt, b, gelu);
    free(wt);
}

/* R/ViT_seq.c:25-57 (conv), :59-81 (flatten+transpose), :83-105 (class token),
 * :107-118 (position embedding).  The conv accumulates bias + products in
 * (ic, kh, kw) order, which is the row order of the [768, 3*16*16] weight
 * blob, so it is a linear layer over the gathered patch. */
void vit_oracle_patch_embed(const float *image, int img, const float *cls,
                            const float *conv_w, const float *conv_b,
                            const float *pos, float *tokens)
{
    const int side = img / PS, np = side * side, kdim = 3 * PS * PS;
    float *patches = xmalloc_f((size_t)np * kdim);
    float *emb = xmalloc_f((size_t)np * D);
    for (int p = 0; p < np; p++) {
        const int oh = p / side, ow = p % side;
        for (int ic = 0; ic < 3; ic++)
            for (int kh = 0; kh < PS; kh++)
                memcpy(patches + (size_t)p * kdim + (ic * PS + kh) * PS,
                       image + ((size_t)ic * img + oh * PS + kh) * img + ow * PS,
                       PS * sizeof(float));
    }
    vit_oracle_linear(patches, emb, np, kdim, D, conv_w, conv_b, 0);
    for (int j = 0; j < D; j++)
        tokens[j] = cls[j] + pos[j];
    for (size_t i = 0; i < (size_t)np * D; i++)
        tokens[D + i] = emb[i] + pos[D + i];
    free(patches);
    free(emb);
}

/* R/ViT_seq.c:120-142.  eps is the double literal 1e-6 (R/ViT_seq.c:21), so
 * var + eps is a double addition narrowed to float by sqrtf's parameter. */
void vit_oracle_layer_norm(const float *x, float *y, int tokens,
                           const float *gamma, const float *beta)
{
#pragma omp parallel for schedule(static)
    for (int t = 0; t < tokens; t++) {
        const float *xr = x + (size_t)t * D;
        float *yr = y + (size_t)t * D;
        float sum = 0.0f, sum_sq = 0.0f;
        for (int i = 0; i < D; i++) {
            float v = xr[i];
            sum += v;
            sum_sq += v * v;
        }
        float mean = sum / D;
        float var = sum_sq / D - mean * mean;
        float inv_std = 1.0f / sqrtf(var + 1e-6);
        for (int i = 0; i < D; i++)
            yr[i] = (xr[i] - mean) * inv_std * gamma[i] + beta[i];
    }
}

/* R/ViT_seq.c:192-262: per head, scores = (q.k summed over d ascending) then
 * divided by sqrtf(64); softmax with max subtraction, expf, float sum, one
 * division per element; out = sum_j p_j * v_j, j ascending. */
void vit_oracle_attention_core(const float *q, const float *k, const float *v,
                               float *o, int tokens)
{
    const int T = tokens;
    const float scale = sqrtf((float)HD);
#pragma omp parallel
    {
        float *kt = xmalloc_f((size_t)HD * T);
        float *s = xmalloc_f((size_t)T);
#pragma omp for schedule(static) collapse(1)
        for (int h = 0; h < NH; h++) {
            const int ho = h * HD;
            for (int d = 0; d < HD; d++)
                for (int j = 0; j < T; j++)
                    kt[(size_t)d * T + j] = k[(size_t)j * D + ho + d];
            for (int i = 0; i < T; i++) {
                const float *qr = q + (size_t)i * D + ho;
                for (int j = 0; j < T; j++)
                    s[j] = 0.0f;
                for (int d = 0; d < HD; d++) {
                    const float qd = qr[d];
                    const float *kr = kt + (size_t)d * T;
                    for (int j = 0; j < T; j++)
                        s[j] += qd * kr[j];
                }
                for (int j = 0; j < T; j++)
                    s[j] = s[j] / scale;
                float mx = s[0];
                for (int j = 1; j < T; j++)
                    if (s[j] > mx)
                        mx = s[j];
                float sum = 0.0f;
                for (int j = 0; j < T; j++) {
                    s[j] = expf(s[j] - mx);
                    sum += s[j];
                }
                for (int j = 0; j < T; j++)
                    s[j] /= sum;
                float acc[HD];
                for (int d = 0; d < HD; d++)
                    acc[d] = 0.0f;
                for (int j = 0; j < T; j++) {
                    const float pj = s[j];
                    const float *vr = v + (size_t)j * D + ho;
                    for (int d = 0; d < HD; d++)
                        acc[d] += pj * vr[d];
                }
                for (int d = 0; d < HD; d++)
                    o[(size_t)i * D + ho + d] = acc[d];
            }
        }
        free(kt);
        free(s);
    }
}

/* R/ViT_seq.c:144-281.  Q, K, V are rows [0,768), [768,1536), [1536,2304) of
 * in_proj (R/ViT_seq.c:150,161-166): one linear layer with 2304 outputs. */
void vit_oracle_mha(const float *x, float *y, int tokens, const float *w_in,
                    const float *b_in, const float *w_out, const float *b_out)
{
    const int T = tokens;
    float *qkv = xmalloc_f((size_t)T * 3 * D);
    float *q = xmalloc_f((size_t)T * D), *k = xmalloc_f((size_t)T * D), *v = xmalloc_f((size_t)T * D);
    float *att = xmalloc_f((size_t)T * D);
    vit_oracle_linear(x, qkv, T, D, 3 * D, w_in, b_in, 0);
    for (int t = 0; t < T; t++) {
        memcpy(q + (size_t)t * D, qkv + (size_t)t * 3 * D, D * sizeof(float));
        memcpy(k + (size_t)t * D, qkv + (size_t)t * 3 * D + D, D * sizeof(float));
        memcpy(v + (size_t)t * D, qkv + (size_t)t * 3 * D + 2 * D, D * sizeof(float));
    }
    vit_oracle_attention_core(q, k, v, att, T);
    vit_oracle_linear(att, y, T, D, D, w_out, b_out, 0);
    free(qkv);
    free(q);
    free(k);
    free(v);
    free(att);
}

/* R/ViT_seq.c:330-370: x -> LN1 -> MHA -> +x -> LN2 -> fc1/GELU/fc2 -> +r1 */
void vit_oracle_encoder(const float *x, float *y, int tokens, const float *const *w)
{
    const size_t n = (size_t)tokens * D;
    float *ln = xmalloc_f(n), *att = xmalloc_f(n), *r1 = xmalloc_f(n);
    float *hid = xmalloc_f((size_t)tokens * HID), *mlp = xmalloc_f(n);
    vit_oracle_layer_norm(x, ln, tokens, w[0], w[1]);
    vit_oracle_mha(ln, att, tokens, w[2], w[3], w[4], w[5]);
    for (size_t i = 0; i < n; i++)
        r1[i] = x[i] + att[i];
    vit_oracle_layer_norm(r1, ln, tokens, w[6], w[7]);
    vit_oracle_linear(ln, hid, tokens, D, HID, w[8], w[9], 1);
    vit_oracle_linear(hid, mlp, tokens, HID, D, w[10], w[11], 0);
    for (size_t i = 0; i < n; i++)
        y[i] = r1[i] + mlp[i];
    free(ln);
    free(att);
    free(r1);
    free(hid);
    free(mlp);
}

/* R/ViT_seq.c:372-397 */
void vit_oracle_softmax(const float *logits, float *probs, int n)
{
    float mx = logits[0];
    for (int i = 1; i < n; i++)
        if (logits[i] > mx)
            mx = logits[i];
    float sum = 0.0f;
    for (int i = 0; i < n; i++) {
        probs[i] = expf(logits[i] - mx);
        sum += probs[i];
    }
    for (int i = 0; i < n; i++)
        probs[i] /= sum;
}

/* R/ViT_seq.c:402-517.  Blob indices: 0 cls, 1/2 conv w/b, 3 pos, 4+12L.. the
 * layer blobs, 148/149 final LN, 150/151 head (R/ViT_seq.c:437-513).  The
 * final LN runs on every token though only row 0 feeds the head (:506-513). */
int vit_oracle_forward(const float *images, int n, int img, const float *const *w,
                       float *probs, float *logits, float *stage_dump)
{
    if (!images || !w || !probs || n < 0 || img <= 0 || img % PS)
        return -1;
    for (int i = 0; i < VIT_ORACLE_NBLOBS; i++)
        if (!w[i])
            return -1;
    const int T = vit_oracle_tokens(img);
    const size_t tok = (size_t)T * D, px = (size_t)3 * img * img;
    float *a = xmalloc_f(tok), *b = xmalloc_f(tok), *lg = xmalloc_f(NCLS);
    for (int im = 0; im < n; im++) {
        vit_oracle_patch_embed(images + (size_t)im * px, img, w[0], w[1], w[2], w[3], a);
        if (stage_dump && im == 0)
            memcpy(stage_dump, a, tok * sizeof(float));
        for (int l = 0; l < VIT_ORACLE_DEPTH; l++) {
            vit_oracle_encoder(a, b, T, w + 4 + 12 * l);
            float *t = a;
            a = b;
            b = t;
            if (stage_dump && im == 0)
                memcpy(stage_dump + (size_t)(l + 1) * tok, a, tok * sizeof(float));
        }
        const int hb = 4 + 12 * VIT_ORACLE_DEPTH; /* 148: final LN gamma/beta, head weight/bias */
        vit_oracle_layer_norm(a, b, T, w[hb], w[hb + 1]);
        vit_oracle_linear(b, lg, 1, D, NCLS, w[hb + 2], w[hb + 3], 0);
        if (logits)
            memcpy(logits + (size_t)im * NCLS, lg, NCLS * sizeof(float));
        vit_oracle_softmax(lg, probs + (size_t)im * NCLS, NCLS);
    }
    free(a);
    free(b);
    free(lg);
    return 0;
}
