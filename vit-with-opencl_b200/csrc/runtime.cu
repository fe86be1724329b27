// csrc/runtime.cu -- device/stream/memory/graph plumbing of the C-ABI layer.
// Replaces the reference's OpenCL runtime glue: platform/context/queue set-up
// (R/ViT_opencl.c:799-861), buffer creation and transfers (R/ViT_opencl.c:125-330)
// and kernelHandler.c.  R/ = /root/reference/MulticoreMainProject/.
#include "common.cuh"

#include <atomic>
#include <stdlib.h>
#include <string.h>

namespace vitcu {

static thread_local char g_err[512] = "no error";
static std::atomic<unsigned long long> g_launches{0};

int set_error(int code, const char *file, int line, const char *what)
{
    const char *base = strrchr(file, '/');
    snprintf(g_err, sizeof(g_err), "[%s:%d] CUDA error %d (%s)", base ? base + 1 : file, line, code,
             what ? what : "?");
    return code;
}

bool pdl_enabled()
{
    static const bool on = !(getenv("VITCU_PDL") && atoi(getenv("VITCU_PDL")) == 0);
    return on;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static std::atomic<unsigned long long> g_kind_launches[LK_COUNT];
void count_launch_kind(int kind)
{
    if (kind >= 0 && kind < LK_COUNT)
        g_kind_launches[kind].fetch_add(1, std::memory_order_relaxed);
}
static const char *const kKindNames[LK_COUNT] = {"gemm_bf16_tc2_kernel", "gemm_bf16_tc_kernel", "attention_tc_kernel",
                                                 "attention_flash_tc_kernel", "attention_simt_kernel", "sgemm_kernel",
                                                 "layernorm_kernel", "patch_embed_tc_kernel", "attention_duo_tc_kernel", "other"};

// one flag per device, lazily allocated
static uint32_t *g_watchdog[64] = {nullptr};
uint32_t *watchdog_flag()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
        return nullptr;
    if (!g_watchdog[dev]) {
        uint32_t *p = nullptr;
        if (cudaMalloc(&p, sizeof(uint32_t)) != cudaSuccess)
            return nullptr;
        cudaMemset(p, 0, sizeof(uint32_t));
        g_watchdog[dev] = p;
    }
    return g_watchdog[dev];
}

} // namespace vitcu

using namespace vitcu;

extern "C" {

const char *vitcu_last_error(void) { return g_err; }

int vitcu_device_count(int *count)
{
    VITCU_REQUIRE(count, "count is NULL");
    *count = 0;
    VITCU_TRY(cudaGetDeviceCount(count));
    return 0;
}

int vitcu_set_device(int device)
{
    VITCU_TRY(cudaSetDevice(device));
    return 0;
}

int vitcu_prepare_device(void)
{
    VITCU_REQUIRE(watchdog_flag() != nullptr, "cannot allocate the watchdog flag");
    return 0;
}

int vitcu_device_info(int device, char *name, int *sms, int *cc, size_t *mem_bytes)
{
    cudaDeviceProp p;
    VITCU_TRY(cudaGetDeviceProperties(&p, device));
    if (name) {
        strncpy(name, p.name, 255);
        name[255] = 0;
    }
    if (sms)
        *sms = p.multiProcessorCount;
    if (cc)
        *cc = p.major * 10 + p.minor;
    if (mem_bytes)
        *mem_bytes = p.totalGlobalMem;
    return 0;
}

int vitcu_stream_create(vitcu_stream *s)
{
    VITCU_REQUIRE(s, "stream out-pointer is NULL");
    cudaStream_t st;
    VITCU_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *s = (vitcu_stream)st;
    return 0;
}
int vitcu_stream_destroy(vitcu_stream s)
{
    VITCU_TRY(cudaStreamDestroy(as_stream(s)));
    return 0;
}
int vitcu_stream_sync(vitcu_stream s)
{
    VITCU_TRY(cudaStreamSynchronize(as_stream(s)));
    return 0;
}
int vitcu_device_sync(void)
{
    VITCU_TRY(cudaDeviceSynchronize());
    return 0;
}
int vitcu_event_create(vitcu_event *e)
{
    VITCU_REQUIRE(e, "event out-pointer is NULL");
    cudaEvent_t ev;
    VITCU_TRY(cudaEventCreate(&ev));
    *e = (vitcu_event)ev;
    return 0;
}
int vitcu_event_destroy(vitcu_event e)
{
    VITCU_TRY(cudaEventDestroy((cudaEvent_t)e));
    return 0;
}
int vitcu_event_record(vitcu_event e, vitcu_stream s)
{
    VITCU_TRY(cudaEventRecord((cudaEvent_t)e, as_stream(s)));
    return 0;
}
int vitcu_event_sync(vitcu_event e)
{
    VITCU_TRY(cudaEventSynchronize((cudaEvent_t)e));
    return 0;
}
int vitcu_stream_wait_event(vitcu_stream s, vitcu_event e)
{
    VITCU_TRY(cudaStreamWaitEvent(as_stream(s), (cudaEvent_t)e, 0));
    return 0;
}
int vitcu_event_elapsed_ms(vitcu_event start, vitcu_event stop, float *ms)
{
    VITCU_REQUIRE(ms, "ms is NULL");
    VITCU_TRY(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return 0;
}

int vitcu_malloc(void **dptr, size_t bytes)
{
    VITCU_REQUIRE(dptr, "dptr is NULL");
    VITCU_TRY(cudaMalloc(dptr, bytes ? bytes : 1));
    return 0;
}
int vitcu_free(void *dptr)
{
    if (dptr)
        VITCU_TRY(cudaFree(dptr));
    return 0;
}
int vitcu_host_alloc(void **hptr, size_t bytes)
{
    VITCU_REQUIRE(hptr, "hptr is NULL");
    VITCU_TRY(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocPortable));
    return 0;
}
int vitcu_host_free(void *hptr)
{
    if (hptr)
        VITCU_TRY(cudaFreeHost(hptr));
    return 0;
}
int vitcu_host_register(void *hptr, size_t bytes)
{
    VITCU_TRY(cudaHostRegister(hptr, bytes, cudaHostRegisterPortable));
    return 0;
}
int vitcu_host_unregister(void *hptr)
{
    VITCU_TRY(cudaHostUnregister(hptr));
    return 0;
}
int vitcu_host_is_pinned(const void *hptr, int *pinned)
{
    VITCU_REQUIRE(hptr && pinned, "NULL argument");
    cudaPointerAttributes a;
    VITCU_TRY(cudaPointerGetAttributes(&a, hptr));
    *pinned = a.type == cudaMemoryTypeHost;
    return 0;
}
int vitcu_memcpy_h2d(void *dst, const void *src, size_t bytes, vitcu_stream s)
{
    VITCU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(s)));
    return 0;
}
int vitcu_memcpy_d2h(void *dst, const void *src, size_t bytes, vitcu_stream s)
{
    VITCU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(s)));
    return 0;
}
int vitcu_memcpy_d2d(void *dst, const void *src, size_t bytes, vitcu_stream s)
{
    VITCU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(s)));
    return 0;
}
int vitcu_memset(void *dst, int value, size_t bytes, vitcu_stream s)
{
    VITCU_TRY(cudaMemsetAsync(dst, value, bytes, as_stream(s)));
    return 0;
}

int vitcu_graph_begin(vitcu_stream s)
{
    VITCU_TRY(cudaStreamBeginCapture(as_stream(s), cudaStreamCaptureModeRelaxed));
    return 0;
}
int vitcu_graph_end(vitcu_stream s, vitcu_graph *g)
{
    VITCU_REQUIRE(g, "graph out-pointer is NULL");
    cudaGraph_t graph = nullptr;
    VITCU_TRY(cudaStreamEndCapture(as_stream(s), &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    VITCU_TRY(e);
    *g = (vitcu_graph)exec;
    return 0;
}
int vitcu_graph_launch(vitcu_graph g, vitcu_stream s)
{
    VITCU_TRY(cudaGraphLaunch((cudaGraphExec_t)g, as_stream(s)));
    return 0;
}
int vitcu_graph_destroy(vitcu_graph g)
{
    if (g)
        VITCU_TRY(cudaGraphExecDestroy((cudaGraphExec_t)g));
    return 0;
}

void vitcu_launch_count_reset(void)
{
    g_launches.store(0);
    for (int k = 0; k < LK_COUNT; k++)
        g_kind_launches[k].store(0);
}
unsigned long long vitcu_launch_count_of(const char *kernel)
{
    for (int k = 0; kernel && k < LK_COUNT; k++)
        if (!strcmp(kernel, kKindNames[k]))
            return g_kind_launches[k].load();
    return 0;
}
unsigned long long vitcu_launch_count(void) { return g_launches.load(); }

int vitcu_watchdog_check(void)
{
    uint32_t *flag = watchdog_flag();
    if (!flag)
        return 0;
    uint32_t v = 0;
    VITCU_TRY(cudaMemcpy(&v, flag, sizeof(v), cudaMemcpyDeviceToHost));
    if (v) {
        cudaMemset(flag, 0, sizeof(v));
        return set_error(VITCU_E_WATCHDOG, __FILE__, __LINE__, "kernel pipeline wait timed out");
    }
    return 0;
}

} // extern "C"
