// runtime.cu -- contexts, buffers, events, errors: the C-ABI replacement of the OpenCL runtime
// calls the reference drivers make (include/b200spmv.h cites each one).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

// one name per b200_opt, in enum order: the environment variable read at b200_ctx_create and the
// name b200_ctx_set_option takes
const char *const b200_opt_names[OPT_COUNT] = {
    "B200_CSR_LANES", "B200_ELL_LANES", "B200_CSR_UNROLL", "B200_ELL_UNROLL", "B200_CSR_STREAM", "B200_SELL_WPC",
    "B200_SELL_UNROLL", "B200_SELL_TMA", "B200_SELL_TMA_BLOCKS", "B200_SELL_TMA_SUSPEND_NS", "B200_COO_U", "B200_CMRS_U",
    "B200_CMRS_WPS", "B200_CMRS_STREAM", "B200_ELLCM_Q", "B200_RING_FLUSH", "B200_RING_POLL", "B200_RING_SLEEP_NS",
    "B200_BCAST_U", "B200_CSR_STREAM_G", "B200_SELL_PIPE",
};

static thread_local char g_last_error[512] = "";

void b200_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof g_last_error, fmt, ap);
    va_end(ap);
}

int b200_cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    b200_set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    // clear sticky-free errors so that the next call reports its own failure
    (void)cudaGetLastError();
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice)
        return B200_ERR_NO_DEVICE;
    if (e == cudaErrorMemoryAllocation) return B200_ERR_OUT_OF_MEMORY;
    return B200_ERR_CUDA;
}

extern "C" {

const char *b200_status_string(int status)
{
    switch (status) {
    case B200_SUCCESS: return "success";
    case B200_ERR_NO_DEVICE: return "no CUDA device";
    case B200_ERR_CUDA: return "CUDA runtime error";
    case B200_ERR_INVALID_VALUE: return "invalid value";
    case B200_ERR_OUT_OF_MEMORY: return "out of device memory";
    case B200_ERR_UNSUPPORTED: return "unsupported";
    case B200_ERR_DOMAIN: return "input outside the builder's domain";
    case B200_ERR_COMM: return "communicator (NCCL) error";
    default: return "unknown status";
    }
}

const char *b200_last_error(void) { return g_last_error; }

int b200_version(void) { return B200SPMV_VERSION; }

int b200_get_device_count(int *count)
{
    B200_REQUIRE(count, "null count");
    *count = 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return b200_cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    if (n <= 0) {
        b200_set_error("no CUDA device present");
        return B200_ERR_NO_DEVICE;
    }
    *count = n;
    return B200_SUCCESS;
}

int b200_device_name(int device, char *buf, size_t buf_len)
{
    B200_REQUIRE(buf && buf_len > 0, "null buffer");
    cudaDeviceProp p;
    B200_CUDA(cudaGetDeviceProperties(&p, device));
    snprintf(buf, buf_len, "%s", p.name);
    return B200_SUCCESS;
}

int b200_device_sm_count(int device, int *sm_count)
{
    B200_REQUIRE(sm_count, "null sm_count");
    B200_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, device));
    return B200_SUCCESS;
}

static int ctx_create_impl(int device, cudaStream_t stream, bool own, b200_ctx **out)
{
    B200_REQUIRE(out, "null ctx out");
    *out = nullptr;
    int n = 0;
    int rc = b200_get_device_count(&n);
    if (rc) return rc;
    if (device < 0 || device >= n) {
        b200_set_error("device %d out of range (have %d)", device, n);
        return B200_ERR_NO_DEVICE;
    }
    B200_CUDA(cudaSetDevice(device));
    int major = 0;
    B200_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major < 10) {
        b200_set_error("device %d is sm_%d0; libb200spmv carries sm_100a code only", device, major);
        return B200_ERR_UNSUPPORTED;
    }
    b200_ctx *c = new b200_ctx();
    c->device = device;
    c->owns_stream = own;
    c->stream = stream;
    c->scratch = nullptr;
    c->host_scratch = nullptr;
    c->watch_flag = false;
    c->watch_saved = false;
    c->persist_set = false;
    // launch overlap is on by default on the library's own queue (only library calls enqueue there,
    // and every one of them but the SpMV launches sets needs_order); a caller-owned stream may carry
    // foreign kernels that write matrix arrays, so there it stays opt-in
    c->overlap = own;
    c->needs_order = true;
    for (int o = 0; o < OPT_COUNT; ++o) {  // the only getenv calls of the library
        const char *e = getenv(b200_opt_names[o]);
        c->opt[o] = (e && *e) ? atoi(e) : kOptUnset;
    }
    if (own) {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete c;
            return b200_cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__);
        }
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&c->l2_bytes, cudaDevAttrL2CacheSize, device);
    cudaDeviceGetAttribute(&c->max_persist_l2, cudaDevAttrMaxPersistingL2CacheSize, device);
    cudaError_t e = cudaMalloc(&c->scratch, 4096);
    if (e == cudaSuccess) e = cudaMemset(c->scratch, 0, 4096);
    if (e == cudaSuccess) e = cudaMallocHost(&c->host_scratch, 4096);
    if (e != cudaSuccess) {
        if (c->scratch) cudaFree(c->scratch);
        if (own) cudaStreamDestroy(c->stream);
        delete c;
        return b200_cuda_fail(e, "scratch allocation", __FILE__, __LINE__);
    }
    *out = c;
    return B200_SUCCESS;
}

int b200_ctx_create(int device, b200_ctx **ctx) { return ctx_create_impl(device, nullptr, true, ctx); }

int b200_ctx_create_on_stream(int device, void *cuda_stream, b200_ctx **ctx)
{
    return ctx_create_impl(device, reinterpret_cast<cudaStream_t>(cuda_stream), false, ctx);
}

int b200_ctx_destroy(b200_ctx *ctx)
{
    if (!ctx) return B200_SUCCESS;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->persist_set) {  // do not leave a device-wide L2 carve-out behind
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    }
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->host_scratch) cudaFreeHost(ctx->host_scratch);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return B200_SUCCESS;
}

int b200_ctx_device(const b200_ctx *ctx, int *device)
{
    B200_REQUIRE(ctx && device, "null argument");
    *device = ctx->device;
    return B200_SUCCESS;
}

int b200_ctx_set_l2_persist(b200_ctx *ctx, const void *dptr, size_t bytes)
{
    B200_ENTER(ctx);
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    if (bytes == 0 || !dptr) {
        attr.accessPolicyWindow.num_bytes = 0;
        B200_CUDA(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
        // the set-aside is DEVICE-wide: left in place it takes up to 79 of the 126 MB of L2 away from
        // every other kernel on the device (measured: the Laplacian kernel 0.150 -> 0.197 ms behind a
        // stale carve-out).  Give it back and turn the persisting lines into normal ones.
        if (ctx->persist_set) {
            B200_CUDA(cudaStreamSynchronize(ctx->stream));
            B200_CUDA(cudaCtxResetPersistingL2Cache());
            B200_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
            ctx->persist_set = false;
        }
        return B200_SUCCESS;
    }
    if (ctx->max_persist_l2 <= 0) return B200_SUCCESS;  // nothing to set aside: best effort
    size_t carve = bytes < (size_t)ctx->max_persist_l2 ? bytes : (size_t)ctx->max_persist_l2;
    B200_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
    ctx->persist_set = true;
    int max_window = 0;
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
    size_t win = bytes;
    if (max_window > 0 && win > (size_t)max_window) win = (size_t)max_window;
    attr.accessPolicyWindow.base_ptr = const_cast<void *>(dptr);
    attr.accessPolicyWindow.num_bytes = win;
    attr.accessPolicyWindow.hitRatio = win <= carve ? 1.0f : (float)carve / (float)win;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    B200_CUDA(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    return B200_SUCCESS;
}

int b200_ctx_set_option(b200_ctx *ctx, const char *name, const char *value)
{
    B200_REQUIRE(ctx && name, "null argument");
    for (int o = 0; o < OPT_COUNT; ++o) {
        if (strcmp(name, b200_opt_names[o]) == 0) {
            ctx->opt[o] = (value && *value) ? atoi(value) : kOptUnset;
            return B200_SUCCESS;
        }
    }
    b200_set_error("b200_ctx_set_option: unknown tuning hook '%s'", name);
    return B200_ERR_INVALID_VALUE;
}

static int check_watch_flag(b200_ctx *ctx);

int b200_ctx_set_launch_overlap(b200_ctx *ctx, int enable)
{
    B200_REQUIRE(ctx, "null context");
    ctx->overlap = enable != 0;
    return B200_SUCCESS;
}

int b200_malloc(b200_ctx *ctx, size_t bytes, void **dptr)
{
    B200_ENTER(ctx);
    B200_REQUIRE(dptr, "null dptr");
    *dptr = nullptr;
    // B200_MALLOC_PAD spare bytes: the vector kernels read whole 16-byte-aligned groups of four entries
    // (32 bytes for fp64 values), so the last group of an array may reach up to 24 bytes past its end
    B200_CUDA(cudaMalloc(dptr, (bytes ? bytes : 1) + B200_MALLOC_PAD));
    return B200_SUCCESS;
}

int b200_free(b200_ctx *ctx, void *dptr)
{
    B200_ENTER(ctx);
    if (dptr) B200_CUDA(cudaFree(dptr));
    return B200_SUCCESS;
}

int b200_memcpy_h2d_async(b200_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    B200_TRACE("b200 H2D");
    B200_ENTER(ctx);
    if (bytes) B200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return B200_SUCCESS;
}

int b200_memcpy_d2h(b200_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    B200_TRACE("b200 D2H");
    B200_ENTER(ctx);
    if (bytes) B200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    return check_watch_flag(ctx);
}

int b200_memcpy_d2h_async(b200_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    B200_ENTER(ctx);
    if (bytes) B200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return B200_SUCCESS;
}

int b200_memcpy_d2d_async(b200_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    B200_ENTER(ctx);
    if (bytes) B200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return B200_SUCCESS;
}

int b200_memset_async(b200_ctx *ctx, void *dst, int byte_value, size_t bytes)
{
    B200_ENTER(ctx);
    if (bytes) B200_CUDA(cudaMemsetAsync(dst, byte_value, bytes, ctx->stream));
    return B200_SUCCESS;
}

// after a stream sync: did a bulk-copy kernel give up on an mbarrier wait?
static int check_watch_flag(b200_ctx *ctx)
{
    if (!ctx->watch_flag) return B200_SUCCESS;
    ctx->watch_flag = false;
    int *h = static_cast<int *>(ctx->host_scratch);
    B200_CUDA(cudaMemcpyAsync(h, ctx->scratch + kWatchFlag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*h != 0) {
        cudaMemsetAsync(ctx->scratch + kWatchFlag, 0, sizeof(int), ctx->stream);
        b200_set_error("a bulk-copy (TMA) kernel timed out waiting for its data: results are invalid");
        return B200_ERR_CUDA;
    }
    return B200_SUCCESS;
}

int b200_sync(b200_ctx *ctx)
{
    B200_ENTER(ctx);
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    return check_watch_flag(ctx);
}

int b200_host_alloc_pinned(size_t bytes, void **hptr)
{
    B200_REQUIRE(hptr, "null hptr");
    *hptr = nullptr;
    B200_CUDA(cudaMallocHost(hptr, bytes ? bytes : 1));
    return B200_SUCCESS;
}

int b200_host_free_pinned(void *hptr)
{
    if (hptr) B200_CUDA(cudaFreeHost(hptr));
    return B200_SUCCESS;
}

int b200_event_create(b200_ctx *ctx, b200_event **ev)
{
    B200_ENTER(ctx);
    B200_REQUIRE(ev, "null event out");
    b200_event *e = new b200_event();
    e->device = ctx->device;
    cudaError_t ce = cudaEventCreate(&e->ev);
    if (ce != cudaSuccess) {
        delete e;
        return b200_cuda_fail(ce, "cudaEventCreate", __FILE__, __LINE__);
    }
    *ev = e;
    return B200_SUCCESS;
}

int b200_event_record(b200_ctx *ctx, b200_event *ev)
{
    B200_ENTER(ctx);
    B200_REQUIRE(ev, "null event");
    B200_CUDA(cudaEventRecord(ev->ev, ctx->stream));
    return B200_SUCCESS;
}

int b200_event_elapsed_ms(b200_event *start, b200_event *stop, float *ms)
{
    B200_REQUIRE(start && stop && ms, "null argument");
    B200_CUDA(cudaSetDevice(stop->device));
    B200_CUDA(cudaEventSynchronize(stop->ev));
    B200_CUDA(cudaEventElapsedTime(ms, start->ev, stop->ev));
    return B200_SUCCESS;
}

// ---- launch graphs: record a sequence of launches once, replay it with one call ----------------
// A cant-sized SpMV lasts ~10 us, less than the host needs to issue it; a solver loop or a
// rotation over several matrices is therefore captured into a CUDA graph and replayed.

struct b200_graph {
    int device;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    bool watched;  // a kernel that can raise kWatchFlag (TMA wait, ring flag wait) was recorded
};

int b200_graph_begin(b200_ctx *ctx)
{
    B200_ENTER(ctx);
    B200_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    // launches recorded from here on set watch_flag without having run: keep the live state aside
    ctx->watch_saved = ctx->watch_flag;
    ctx->watch_flag = false;
    return B200_SUCCESS;
}

int b200_graph_end(b200_ctx *ctx, b200_graph **graph)
{
    B200_ENTER(ctx);
    B200_REQUIRE(graph, "null graph out");
    *graph = nullptr;
    cudaGraph_t g = nullptr;
    const bool watched = ctx->watch_flag;
    ctx->watch_flag = ctx->watch_saved;
    B200_CUDA(cudaStreamEndCapture(ctx->stream, &g));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, g, 0);
    if (e != cudaSuccess) {
        cudaGraphDestroy(g);
        return b200_cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
    }
    b200_graph *out = new b200_graph();
    out->device = ctx->device;
    out->graph = g;
    out->exec = exec;
    out->watched = watched;
    *graph = out;
    return B200_SUCCESS;
}

int b200_graph_launch(b200_ctx *ctx, b200_graph *graph)
{
    B200_ENTER(ctx);
    B200_REQUIRE(graph && graph->device == ctx->device, "graph was recorded on another device");
    B200_CUDA(cudaGraphLaunch(graph->exec, ctx->stream));
    if (graph->watched) ctx->watch_flag = true;  // every replay can time out, not only the recording
    return B200_SUCCESS;
}

int b200_graph_destroy(b200_graph *graph)
{
    if (!graph) return B200_SUCCESS;
    cudaSetDevice(graph->device);
    cudaGraphExecDestroy(graph->exec);
    cudaGraphDestroy(graph->graph);
    delete graph;
    return B200_SUCCESS;
}

int b200_event_destroy(b200_event *ev)
{
    if (!ev) return B200_SUCCESS;
    cudaSetDevice(ev->device);
    cudaEventDestroy(ev->ev);
    delete ev;
    return B200_SUCCESS;
}

}  // extern "C"
