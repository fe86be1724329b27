/* host/vit_stage.c -- worker threads that gather pageable host images into pinned staging.
 *
 * The reference hands ViT_opencl one malloc'd buffer per image (R/Network.c:84-105, R/ =
 * /root/reference/MulticoreMainProject/).  A DMA engine cannot read pageable memory, so something
 * has to copy every image into pinned memory first; the CUDA driver does that on the calling
 * thread at ~10 GB/s, which is slower than the GPU consumes images.  This pool spreads the copy
 * over several cores so the upload of a chunk costs less than its forward.
 */
#include "vit_engine_internal.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define VIT_STAGE_STREAMING 1
#else
#define VIT_STAGE_STREAMING 0
#endif

#define VIT_STAGE_MAX_THREADS 32
#define VIT_STAGE_PIECE ((size_t)128 << 10)

struct vit_stager {
    pthread_t th[VIT_STAGE_MAX_THREADS];
    int nth; /* workers besides the calling thread */
    pthread_mutex_t mu;
    pthread_cond_t cv_job, cv_done;
    unsigned gen;
    int stop, running;
    /* the current job: `count` blocks of `bytes`, block i from structs[i].data or contig + i*bytes,
     * handed out in pieces of VIT_STAGE_PIECE bytes so a small group still occupies every thread */
    char *dst;
    const vitb200_image *structs;
    const char *contig;
    size_t bytes;
    int count, pieces; /* pieces per block */
    atomic_int next;
};

/* One piece into the pinned slot.  The destination is written once and next read by a DMA engine, never by this
 * core, which is the textbook case for streaming (non-temporal) stores: no read-for-ownership of the destination
 * lines, two bytes over the memory bus per byte copied instead of three.  MEASURED, it loses: one ViT_opencl call over
 * 4 096 images on 8 GPUs (8 x 8 copy threads, the saturated case) 87.4 k images/s with streaming stores against 92.5 k
 * with memcpy, 4 GPUs 74.8 k against 79.9 k, and a single copy thread is 3 % slower too (profiles/r02_batch1_latency.md)
 * -- glibc's memcpy already picks the fastest path of this host for 128 KB pieces.  VITB200_STAGE_STREAMING=1 selects
 * the streaming copy for hosts where that is not so. */
static void stage_copy(char *dst, const char *src, size_t n)
{
#if VIT_STAGE_STREAMING
    static int streaming = -1;
    if (streaming < 0) {
        const char *v = getenv("VITB200_STAGE_STREAMING");
        streaming = v && atoi(v) != 0;
    }
    if (streaming && ((uintptr_t)dst & 15) == 0 && n >= 256) {
        size_t i = 0;
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128((const __m128i *)(src + i));
            const __m128i b = _mm_loadu_si128((const __m128i *)(src + i + 16));
            const __m128i c = _mm_loadu_si128((const __m128i *)(src + i + 32));
            const __m128i d = _mm_loadu_si128((const __m128i *)(src + i + 48));
            _mm_stream_si128((__m128i *)(dst + i), a);
            _mm_stream_si128((__m128i *)(dst + i + 16), b);
            _mm_stream_si128((__m128i *)(dst + i + 32), c);
            _mm_stream_si128((__m128i *)(dst + i + 48), d);
        }
        if (i < n)
            memcpy(dst + i, src + i, n - i);
        _mm_sfence(); /* the streamed lines are globally visible before the job is reported done */
        return;
    }
#endif
    memcpy(dst, src, n);
}

static void run_job(vit_stager *s)
{
    for (;;) {
        const int u = atomic_fetch_add(&s->next, 1);
        if (u >= s->count * s->pieces)
            break;
        const int i = u / s->pieces;
        const size_t off = (size_t)(u % s->pieces) * VIT_STAGE_PIECE;
        const size_t len = s->bytes - off < VIT_STAGE_PIECE ? s->bytes - off : VIT_STAGE_PIECE;
        const char *src = s->structs ? (const char *)s->structs[i].data : s->contig + (size_t)i * s->bytes;
        stage_copy(s->dst + (size_t)i * s->bytes + off, src + off, len);
    }
}

static void *worker(void *arg)
{
    vit_stager *s = (vit_stager *)arg;
    unsigned seen = 0;
    pthread_mutex_lock(&s->mu);
    for (;;) {
        while (!s->stop && s->gen == seen)
            pthread_cond_wait(&s->cv_job, &s->mu);
        if (s->stop)
            break;
        seen = s->gen;
        pthread_mutex_unlock(&s->mu);
        run_job(s);
        pthread_mutex_lock(&s->mu);
        if (--s->running == 0)
            pthread_cond_signal(&s->cv_done);
    }
    pthread_mutex_unlock(&s->mu);
    return NULL;
}

int vit_stager_threads_default(void)
{
    const char *env = getenv("VITB200_STAGE_THREADS");
    long n = env ? atol(env) : 0;
    if (n <= 0) {
        n = sysconf(_SC_NPROCESSORS_ONLN) / 2;
        if (n > 8)
            n = 8;
    }
    if (n < 1)
        n = 1;
    if (n > VIT_STAGE_MAX_THREADS)
        n = VIT_STAGE_MAX_THREADS;
    return (int)n;
}

vit_stager *vit_stager_create(int threads)
{
    vit_stager *s = (vit_stager *)calloc(1, sizeof(*s));
    if (!s)
        return NULL;
    pthread_mutex_init(&s->mu, NULL);
    pthread_cond_init(&s->cv_job, NULL);
    pthread_cond_init(&s->cv_done, NULL);
    if (threads > VIT_STAGE_MAX_THREADS)
        threads = VIT_STAGE_MAX_THREADS;
    /* the caller copies too, so `threads` total means threads-1 workers; a failed
     * pthread_create only lowers the count */
    for (int i = 0; i < threads - 1; i++) {
        if (pthread_create(&s->th[s->nth], NULL, worker, s) != 0)
            break;
        s->nth++;
    }
    return s;
}

void vit_stager_destroy(vit_stager *s)
{
    if (!s)
        return;
    pthread_mutex_lock(&s->mu);
    s->stop = 1;
    pthread_cond_broadcast(&s->cv_job);
    pthread_mutex_unlock(&s->mu);
    for (int i = 0; i < s->nth; i++)
        pthread_join(s->th[i], NULL);
    pthread_cond_destroy(&s->cv_job);
    pthread_cond_destroy(&s->cv_done);
    pthread_mutex_destroy(&s->mu);
    free(s);
}

void vit_stager_copy(vit_stager *s, void *dst, const vitb200_image *structs, const void *contig, size_t bytes,
                     int count)
{
    pthread_mutex_lock(&s->mu);
    s->dst = (char *)dst;
    s->structs = structs;
    s->contig = (const char *)contig;
    s->bytes = bytes;
    s->count = count;
    s->pieces = (int)((bytes + VIT_STAGE_PIECE - 1) / VIT_STAGE_PIECE);
    atomic_store(&s->next, 0);
    s->running = s->nth;
    s->gen++;
    pthread_cond_broadcast(&s->cv_job);
    pthread_mutex_unlock(&s->mu);
    run_job(s);
    pthread_mutex_lock(&s->mu);
    while (s->running)
        pthread_cond_wait(&s->cv_done, &s->mu);
    pthread_mutex_unlock(&s->mu);
}
